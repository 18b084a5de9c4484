"""CPU tests of the host-side utilities that carry the reference's public names (admm_utils, Logger, console printers)
and of the game model's horizon inference.  Expected strings / values were produced by the unmodified reference modules
(SCvx/utils/multi_agent_logging.py, SCvx/utils/logging.py, SCvx/optimization/admm_utils.py) in the build container."""
import csv
import io
import json
from contextlib import redirect_stdout

import numpy as np
import pytest

from scvx_b200.optimization import admm_utils
from scvx_b200.utils import multi_agent_logging as mal
from scvx_b200.utils.logging import METRIC_KEYS, Logger


def _printed(fn, *args):
    buf = io.StringIO()
    with redirect_stdout(buf):
        fn(*args)
    return buf.getvalue()


def test_console_formats_are_the_references():
    line = _printed(mal.print_iteration, 3, 1.2e-3, 4.5, 0.1, 2e-5, 3.3, 1e-9, 24.14159, 1.728)
    assert line == ("Iter  3 | v=1.200e-03 | slack=4.500e+00 | p_res=1.000e-01 | d_res=2.000e-05 "
                    "| Δx=3.30e+00 | Δs=1.00e-09 | o=24.142 | tr= 1.728\n")
    assert _printed(mal.print_summary, 10, 24.1) == (
        "\n=== SCvx+ADMM Summary ===\n  Total iterations: 10\n  Final time scale o: 24.100\n=========================\n\n")
    assert "  Total runtime:    3.14s\n" in _printed(mal.print_summary, 10, 24.1, 3.14159)


def test_admm_utils_known_answers():
    """SCvx/multi_agent_tests/test_admm_utils.py:7-45."""
    assert admm_utils.WEIGHT_COLLISION_SLACK == 1e5
    p = np.array([[1.0, 2.0], [3.0, 4.0]]); Y = np.array([[1.0, 0.0], [0.0, 4.0]])
    assert admm_utils.primal_residual(p, Y) == np.linalg.norm(p - Y) == pytest.approx(np.sqrt(13.0))
    assert admm_utils.dual_residual(Y, np.zeros_like(Y)) == np.linalg.norm(Y)
    assert admm_utils.update_rho_admm(1.0, 20.0, 1.0) == 2.0           # primal residual dominates: grow
    assert admm_utils.update_rho_admm(1.0, 1.0, 20.0) == 0.5           # dual residual dominates: shrink
    assert admm_utils.update_rho_admm(1.0, 1.0, 1.0) == 1.0
    assert admm_utils.update_rho_admm(3.0, 1.0, 20.0, tau_dec=3.0) == 3.0 / 3.0
    rng = np.random.default_rng(0)
    for _ in range(50):                                                # bit-level agreement with numpy's Frobenius norm
        a, b = rng.normal(size=(3, 50)), rng.normal(size=(3, 50))
        assert admm_utils.primal_residual(a, b) == np.linalg.norm(a - b)


def test_logger_files(tmp_path):
    lg = Logger()
    lg.save_csv(str(tmp_path / "none.csv"))
    assert not (tmp_path / "none.csv").exists()                        # nothing logged, nothing written
    lg.log({"iter": 0, "nu_norm": 1.5e-3, "note": "a,b"})
    lg.log_metrics([[1.0, 2.0, 3.0, 4.0, 5.0, 6.0]], first_iter=1)
    assert set(lg.records[1]) == {"iter", *METRIC_KEYS} and lg.records[1]["iter"] == 1 and lg.records[1]["sigma"] == 6.0
    lg.records.pop()
    lg.log({"iter": 1, "nu_norm": 0.0, "note": None})
    lg.save_csv(str(tmp_path / "r.csv")); lg.save_json(str(tmp_path / "r.json"))
    rows = list(csv.DictReader(open(tmp_path / "r.csv", newline="")))
    assert rows == [{"iter": "0", "nu_norm": "0.0015", "note": "a,b"}, {"iter": "1", "nu_norm": "0.0", "note": ""}]
    assert json.load(open(tmp_path / "r.json")) == lg.records
    assert open(tmp_path / "r.json").read().startswith("[\n  {\n    \"iter\": 0,")
    lg.log({"iter": 2, "surprise": 1})
    with pytest.raises(ValueError):
        lg.save_csv(str(tmp_path / "bad.csv"))
    lg.clear()
    assert lg.records == []


def test_game_model_horizon_without_obstacles():
    """GameUnicycleModel(obstacles=[]).get_cost_function crashed on the holders AgentBestResponse passes."""
    from scvx_b200.models.game_model import GameUnicycleModel
    from scvx_b200.optimization.sc_problem import _Holder
    m = GameUnicycleModel(r_init=np.array([0.0, 0.0, 0.0]), r_final=np.array([1.0, 0.0, 0.0]), obstacles=[])
    cost = m.get_cost_function(neighbour_pos=[_Holder((2, 37))])
    assert cost["control_weight"] == 1.0 and len(m.z_params) == 1 and m.z_params[0].shape == (2, 37)
    m2 = GameUnicycleModel(r_init=np.array([0.0, 0.0, 0.0]), r_final=np.array([1.0, 0.0, 0.0]), obstacles=[])
    m2.get_cost_function(neighbour_pos=[np.zeros((2, 21)), np.zeros((2, 21))])
    assert [z.shape for z in m2.z_params] == [(2, 21), (2, 21)]
