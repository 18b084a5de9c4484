"""bench.py --impl reference: the reference arm (oracle port on the host cores) prints ONE JSON line with the contract's keys;
under torchrun rank 0 alone runs and prints it, the other ranks exit 0 without work.  (The CUDA arm needs a GPU: -m gpu tests
and the driver's own runs cover it.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e"}


def _check(line, n_gpus):
    d = json.loads(line)
    assert KEYS <= set(d), sorted(KEYS - set(d))
    assert d["impl"] == "reference" and d["metric"] == "scvx_agent_iterations_per_sec" and d["unit"] == "agent-iterations/s"
    assert d["n_gpus"] == n_gpus and d["steps"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["dtype"] == "f64" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    # stages 1-2 run the reference's own package when baseline/_ref holds it (baseline/install_reference.py), else the oracle port
    has_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "SCvx"))
    assert cb["kind"] == ("reference" if has_ref else "port") and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"]
    assert "SAMPLE" in cb["sample"] and "outer iterations 0-3" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_single_process():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, check=True, timeout=600).stdout
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    _check(lines[0], 1)


def test_reference_arm_under_torchrun_prints_once():
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29641", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, check=True, timeout=600).stdout
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    _check(lines[0], 2)
