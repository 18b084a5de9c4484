"""The ncu report tool reads the committed metric capture of every kernel (profiles/r01k_all_kernels_metrics.csv) and the
table it prints covers every kernel the C-ABI library launches (north_star: roofline fractions for every kernel)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_all_kernel_report_covers_every_global_kernel():
    # round 2's capture (every kernel the drivers launch now) + round 1's for kernels that did not change and are no longer on
    # the batched drivers' path (linearize_collision_kernel: still launched by the literal per-agent mirrors)
    rows = []
    for name in ("r02_all_kernels_metrics.csv", "r01k_all_kernels_metrics.csv"):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_all_report.py"), os.path.join(ROOT, "profiles", name)],
                             capture_output=True, text=True, check=True).stdout
        rows += [ln for ln in out.splitlines() if ln.startswith("| `")]
    reported = {re.match(r"\| `([A-Za-z0-9_]+)", ln).group(1) for ln in rows}
    # every __global__ kernel defined under csrc/ (probe / flush helpers of bench.py aside)
    defined = set()
    csrc = os.path.join(ROOT, "dynamic-programming-multiagent-trajectory-optimiziation_b200", "csrc")
    for f in os.listdir(csrc):
        if f.endswith((".cu", ".cuh")):
            src = open(os.path.join(csrc, f)).read()
            defined |= set(re.findall(r"__global__\s+void(?:\s+__launch_bounds__\([^)]*\))?\s+([A-Za-z0-9_]+)\s*\(", src))
    # not step kernels: bench.py's probe / flush helpers, the constant fill of the distance matrices, and sbar_qp_kernel
    # (the nq > 8 path of the consensus QP, not reached by any shipped scenario)
    # and mu0_from_iters_kernel (n-element map of the opt-in adaptive barrier start, added after the r01k capture)
    defined -= {"fp64_fma_probe_kernel", "l2_flush_kernel", "fill_kernel", "sbar_qp_kernel", "mu0_from_iters_kernel"}
    defined -= {"scvx_user_foh", "scvx_user_piecewise", "scvx_user_full"}      # generated at run time per user model (codegen.py)
    assert defined, "no kernels found"
    assert defined <= reported, sorted(defined - reported)
    for ln in rows:                                                                # every row carries both roofline fractions
        assert ln.count("%") >= 5, ln
