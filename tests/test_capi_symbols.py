"""The C-ABI library loads and exports every symbol include/scvx_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "scvx_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scvx_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_entry_points():
    names = _declared()
    for must in ("scvx_foh_batched", "scvx_linearize_obstacles_batched", "scvx_linearize_collision_batched",
                 "scvx_solve_batched", "scvx_consensus_update", "scvx_outer_update"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from scvx_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    try:
        lib = ctypes.CDLL(_lib.LIB_PATH)
    except OSError as e:   # e.g. libcudart not resolvable on a CPU-only box
        pytest.skip(f"cannot dlopen without CUDA runtime: {e}")
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(_lib.SIGNATURES) == set(_declared())
    assert lib.scvx_abi_version() == 1
    nx, nu, d = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.scvx_model_dims(0, ctypes.byref(nx), ctypes.byref(nu), ctypes.byref(d)) == 0
    assert (nx.value, nu.value, d.value) == (3, 2, 2)
    assert lib.scvx_model_dims(99, None, None, None) == -1


@pytest.mark.parametrize("cname,pyname", [("scvx_solve_args", "SolveArgs"), ("scvx_lti_args", "LtiArgs")])
def test_args_structs_match_header(cname, pyname):
    """Field order of the ctypes mirrors == field order of the structs in the header."""
    from scvx_b200 import _lib
    src = open(os.path.join(ROOT, "include", "scvx_b200.h")).read()
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        stmt = re.sub(r"^(const\s+)?(unsigned\s+long\s+long|unsigned\s+char|double|int|void)\s*", "", stmt)
        fields += [f.strip().lstrip("*").strip() for f in stmt.split(",")]
    assert fields == [f[0] for f in getattr(_lib, pyname)._fields_]


def test_integration_doc_lists_every_entry_point():
    """INTEGRATION.md's entry-point table names every function include/scvx_b200.h declares."""
    import os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "scvx_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    names = set(re.findall(r"^(?:int|unsigned long long|const char\*)\s+(scvx_[a-z0-9_]+)\(", hdr, flags=re.M))
    assert len(names) >= 25
    missing = sorted(n for n in names if n not in doc)
    assert not missing, missing
