"""Place the UNMODIFIED reference package under baseline/_ref (git-ignored; it travels to the GPU box with the snapshot).

`python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference` fails:
"Directory '/root/reference' is not installable. Neither 'setup.py' nor 'pyproject.toml' found."  The reference is a plain
source tree of pure-Python packages, so the install is a copy of its `SCvx` package (Python files only: no docs, caches or
figures).  Nothing here is edited; `bench.py --impl reference` imports it through oracle/refshim.py with the inert cvxpy stub
(cvxpy/ECOS are absent from the image, so only stages 1-2 of the reference can run: its FirstOrderHold and model lambdas).

usage: python baseline/install_reference.py        (called by __graft_entry__.build() when /root/reference exists)"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SCVX_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")


def install(force=False):
    pkg = os.path.join(SRC, "SCvx")
    if not os.path.isdir(pkg):
        return None
    out = os.path.join(DST, "SCvx")
    if os.path.isdir(out) and not force:
        return DST
    if os.path.isdir(out):
        shutil.rmtree(out)
    keep = lambda d, names: [n for n in names if n in ("docs", "__pycache__", "visualization", "examples")   # noqa: E731
                             or (os.path.isfile(os.path.join(d, n)) and not n.endswith(".py"))]
    shutil.copytree(pkg, out, ignore=keep)
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
