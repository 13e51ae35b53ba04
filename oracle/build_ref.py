"""Recipe for ``oracle/_ref``: the UNMODIFIED reference modules of the hot path, taken from where they lie under
``/root/reference`` (read-only) so that ``bench.py --impl reference`` can time the reference ITSELF on the GPU box's host
cores (``cpu_baseline.kind = "reference"``).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference is pure Python (no build system, no ``setup.py``: the
``pip install --target`` recipe of the bench contract has nothing to install), so "building" it is a file copy.  The output
directory is git-ignored -- reference sources never enter this repository's history -- but not gpurun-ignored, so it
travels to the GPU box like the built ``.so``.  ``/root/reference`` does not exist there; when ``oracle/_ref`` is absent
the reference arm falls back to the oracle port (``kind = "port"``).

    python oracle/build_ref.py            (also run by __graft_entry__.build() when /root/reference is present)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
MODULES = ["FunctionManager.py", "NewtonSolver.py", "NewtonSolverInfeasibleStart.py", "PhaseOneSolver.py", "PhaseOne.py",
           "LPSolver.py", "QPSolver.py", "SOCPSolver.py", "LassoSolver.py"]


def build(src="/root/reference", quiet=False):
    if not os.path.isdir(src):
        return None
    os.makedirs(OUT, exist_ok=True)
    for m in MODULES:
        shutil.copyfile(os.path.join(src, m), os.path.join(OUT, m))
    if not quiet:
        print(f"oracle/_ref: {len(MODULES)} reference modules copied from {src}")
    return OUT


def import_reference():
    """Import the reference's solver classes from oracle/_ref (cvxpy / matplotlib, which it imports for its optional
    pre-check and plots, are stubbed as in tests/golden/generate_golden.py).  Returns a dict of classes or None."""
    if not os.path.exists(os.path.join(OUT, "LPSolver.py")):
        return None
    import types

    for name in ("cvxpy", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, OUT)
    try:
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()):  # "Not able to run with GPU" banners
            from LassoSolver import LassoSolver
            from LPSolver import LPSolver
            from QPSolver import QPSolver
            from SOCPSolver import SOCPSolver
    finally:
        sys.path.remove(OUT)
    return {"LPSolver": LPSolver, "QPSolver": QPSolver, "SOCPSolver": SOCPSolver, "LassoSolver": LassoSolver}


if __name__ == "__main__":
    build(*sys.argv[1:2])
