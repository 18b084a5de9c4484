"""Per-iteration record keeper with the reference's interface (SCvx/utils/logging.py:6-52): `records`, `log`, `save_csv`,
`save_json`, `clear`.  The file formats are the reference's (CSV: header taken from the first record, one row per record;
JSON: a list of objects, indent 2) so that downstream scripts such as SCvx/examples/compare_admm_vs_nash.py read both alike.

`log_metrics` is an addition for the batched engine: it turns one agent's column of the device metrics table
(`scvx_outer_update`: nu_norm, slack_norm, dx, du, ds, sigma per outer iteration) into the same records.
"""
import csv
import json
from typing import Dict, Iterable, List

METRIC_KEYS = ("nu_norm", "slack_norm", "dx", "du", "ds", "sigma")     # column order of scvx_outer_update's metrics


class Logger:
    def __init__(self):
        self.records: List[Dict] = []

    def log(self, record: Dict) -> None:
        self.records += [record]

    def log_metrics(self, rows: Iterable, first_iter: int = 0) -> None:
        """rows: iterable of 6-vectors in METRIC_KEYS order, one per outer iteration."""
        for offset, row in enumerate(rows):
            entry = {"iter": first_iter + offset}
            entry.update((key, float(val)) for key, val in zip(METRIC_KEYS, row))
            self.log(entry)

    def save_csv(self, filepath: str) -> None:
        if len(self.records) == 0:
            return                      # nothing to write: the reference leaves no file behind either
        columns = [*self.records[0]]
        with open(filepath, "w", newline="") as handle:
            out = csv.writer(handle)
            out.writerow(columns)
            for rec in self.records:
                unknown = set(rec) - set(columns)
                if unknown:             # same failure mode as csv.DictWriter(extrasaction="raise")
                    raise ValueError("dict contains fields not in fieldnames: " + ", ".join(repr(k) for k in sorted(unknown)))
                out.writerow([rec.get(col, "") for col in columns])

    def save_json(self, filepath: str) -> None:
        text = json.dumps(self.records, indent=2)
        with open(filepath, "w") as handle:
            handle.write(text)

    def clear(self) -> None:
        self.records = []
