"""Benchmark / CSV harness compatible with the reference's ``testSolver.py`` and ``parseAndPlot.py`` (SURVEY.md
8(f)-2): the same seeded problem streams (``np.random.seed(1)``, ``testSolver.py:30-91, 453-522, 822-880,
1053-1104``), the same solver settings, and result files in the layout ``parse_csv`` reads
(``parseAndPlot.py:7-141``): first line ``num_tests,N``, then a header and one row per (problem size, repetition),
zeros for runs that were skipped (the parser turns them into NaN).

The ``*_gpu_*`` columns are this engine on the B200.  The ``*_cpu_*`` columns are the reference's own NumPy arm when
a checkout of the reference is handed in (``reference_dir``; its modules are imported from there, cvxpy / matplotlib
stubbed) and zeros otherwise; the ``cvxpy_*`` / ``jax_*`` columns are always zeros (neither package is a dependency).

    python -m ipm_b200.harness --out results/b200_ --n 100 200 400 --N 3 [--reference-dir /path/to/reference]
"""

import argparse
import sys
import time
import types

import numpy as np

try:
    from .LassoSolver import LassoSolver
    from .LPSolver import LPSolver
    from .QPSolver import QPSolver
    from .SOCPSolver import SOCPSolver
except ImportError:  # flat-module use
    from LassoSolver import LassoSolver
    from LPSolver import LPSolver
    from QPSolver import QPSolver
    from SOCPSolver import SOCPSolver

LP_SETTINGS = dict(epsilon=1e-4, mu=15, t0=1, max_inner_iters=20, max_outer_iters=10, beta=0.5, alpha=0.05)
QP_SETTINGS = dict(epsilon=1e-8, mu=15, t0=0.01, max_inner_iters=100, max_outer_iters=10, beta=0.6, alpha=0.4)
SOCP_SETTINGS = dict(epsilon=1e-4, mu=15, t0=0.1, max_inner_iters=500, max_outer_iters=10, beta=0.5, alpha=0.05)
LASSO_SETTINGS = dict(rho=0.4, check_stop=10, add_bias=True, normalize_A=False, positive=False, compute_loss=False,
                      adaptive_rho=False, num_chunks=0, eps_rel=1e-6, eps_abs=1e-6, check_cvxpy=False)
LASSO_PROBLEMS = 30  # fixed in the reference (testSolver.py:1066, parseAndPlot.py:38)


def _repetitions(n, N):
    """testSolver.py:66-71."""
    return N if n < 1000 else (int(N / 2) if n < 2500 else 3)


def load_reference(reference_dir):
    """Import the reference's solver classes (NumPy arm) from a checkout; None when no directory is given."""
    if reference_dir is None:
        return None
    for name in ("cvxpy", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    saved = list(sys.path)
    sys.path.insert(0, reference_dir)
    try:
        import importlib

        mods = {}
        for m in ("LPSolver", "QPSolver", "SOCPSolver", "LassoSolver"):
            sys.modules.pop(m, None)
            mods[m] = getattr(importlib.import_module(m), m)
    finally:
        sys.path[:] = saved
    return mods


def _timed_solve(solver):
    tik = time.time()
    out = solver.solve()
    return time.time() - tik, out


def _write(filename, num_tests, N, columns):
    """pandas-free writer of the reference's layout (df.to_csv(index=False) + prepended ``num_tests,N`` line)."""
    names = list(columns)
    rows = len(columns[names[0]])
    with open(filename, "w") as f:
        f.write(f"{num_tests},{N}\n")
        f.write(",".join(names) + "\n")
        for r in range(rows):
            f.write(",".join(repr(float(columns[k][r])) if k != "n_values" else str(int(columns[k][r]))
                             for k in names) + "\n")


def parse_csv(filename, origin):
    """Mirror of parseAndPlot.parse_csv for LP / QP / SOCP files: (N, num_tests, n_values, {column: [num_tests, N]})
    with zeros mapped to NaN."""
    with open(filename) as f:
        num_tests, N = (int(v) for v in f.readline().split(","))
        names = f.readline().strip().split(",")
        data = np.loadtxt(f, delimiter=",", ndmin=2)
    cols = {k: data[:, i] for i, k in enumerate(names)}
    n_values = cols.pop("n_values").reshape(-1, N)[:, 0].astype(int)
    prefix = {"LP": "ls", "QP": "qp", "SOCP": "socp"}[origin]
    need = ["cvxpy_times", "cvxpy_values"] + [f"{prefix}_{a}_{b}" for a in ("gpu", "cpu") for b in ("times", "values")]
    out = {}
    for k in need + [k for k in cols if k.startswith("jax")]:
        v = cols[k].reshape(num_tests, N).copy()
        v[v == 0] = np.nan
        out[k] = v
    return N, num_tests, n_values, out


def _run(kind, n_values, N, filename, verbose, ref, generate, settings, prefix, make, extra_cols=()):
    np.random.seed(1)
    n_values = np.array(n_values, dtype=np.int32)
    num_tests = len(n_values)
    t_gpu, v_gpu = np.zeros((num_tests, N)), np.zeros((num_tests, N))
    t_cpu, v_cpu = np.zeros((num_tests, N)), np.zeros((num_tests, N))
    for count, n in enumerate(n_values):
        m, k = int(0.8 * n), int(0.2 * n)
        for i in range(_repetitions(n, N)):
            prob = generate(int(n), m, k)
            s = make(LPSolver if kind == "LP" else QPSolver if kind == "QP" else SOCPSolver, prob, settings, True)
            t_gpu[count, i], _ = _timed_solve(s)
            v_gpu[count, i] = s.value
            if verbose:
                print(f"{kind} n={n} rep {i}: B200 {t_gpu[count, i]:.3f} s, value {s.value}")
            del s
            if ref is not None:
                r = make(ref[kind + "Solver"], prob, settings, False)
                t_cpu[count, i], _ = _timed_solve(r)
                v_cpu[count, i] = r.value
    if filename is not None:
        zeros = np.zeros(num_tests * N)
        cols = {"n_values": np.repeat(n_values, N), "cvxpy_times": zeros, "cvxpy_values": zeros,
                f"{prefix}_gpu_times": t_gpu.ravel(), f"{prefix}_gpu_values": v_gpu.ravel(),
                f"{prefix}_cpu_times": t_cpu.ravel(), f"{prefix}_cpu_values": v_cpu.ravel()}
        for c in extra_cols:
            cols[c] = zeros
        _write(filename, num_tests, N, cols)
    return t_gpu, v_gpu, t_cpu, v_cpu


def _make(cls, prob, settings, gpu):
    return cls(**prob, use_gpu=gpu, suppress_print=True, check_cvxpy=False, **settings)


def test_LP(n_values, verbose=False, N=10, filename=None, reference=None):
    """testSolver.py:15-276."""
    def generate(n, m, k):
        A = np.random.uniform(low=-2, high=2, size=(m, n))
        C = np.random.uniform(low=-2, high=2, size=(k, n))
        x_feas = np.random.uniform(low=-2, high=2, size=(n))
        c = np.random.uniform(low=-2, high=2, size=(n))
        return dict(c=c, A=A, b=A @ x_feas, C=C, d=C @ x_feas, lower_bound=-3, upper_bound=3)

    return _run("LP", n_values, N, filename, verbose, reference, generate, LP_SETTINGS, "ls", _make,
                extra_cols=("jax_times", "jax_values"))


def test_QP(n_values, verbose=False, N=10, filename=None, reference=None):
    """testSolver.py:437-808 (k = 20 inequality rows, :487)."""
    def generate(n, m, k):
        k = 20
        Pp = np.random.uniform(low=-2, high=2, size=(m, n))
        P = Pp.T @ Pp + np.eye(n)
        A = np.random.uniform(low=-2, high=2, size=(m, n))
        C = np.random.uniform(low=-2, high=2, size=(k, n))
        x_feas = np.random.uniform(low=-2, high=2, size=(n))
        q = np.random.uniform(low=-2, high=2, size=(n))
        return dict(P=P, q=q, A=A, b=A @ x_feas, C=C, d=C @ x_feas, lower_bound=-3, upper_bound=3)

    return _run("QP", n_values, N, filename, verbose, reference, generate, QP_SETTINGS, "qp", _make,
                extra_cols=("jax_times", "jax_values"))


def test_SOCP(n_values, verbose=False, N=10, filename=None, reference=None):
    """testSolver.py:810-1034 (5 cones of m rows, k = 50 equalities; d_j uses the LIST b, i.e. the spectral norm of
    the stacked residuals so far, exactly as :878 does)."""
    def generate(n, m, k):
        k = 50
        Pp = np.random.uniform(low=-2, high=2, size=(m, n))
        P = Pp.T @ Pp + np.eye(n)
        q = np.random.uniform(low=-2, high=2, size=(n))
        A, b, c, d = [], [], [], []
        x0 = np.random.randn(n)
        for j in range(5):
            A.append(np.random.randn(m, n))
            b.append(np.random.randn(m))
            c.append(np.random.randn(n))
            d.append(float(np.linalg.norm(A[j] @ x0 + b, 2) - c[j] @ x0))
        F = np.random.randn(k, n)
        return dict(P=P, q=q, A=A, b=b, c=c, d=d, F=F, g=F @ x0, lower_bound=None, upper_bound=None)

    return _run("SOCP", n_values, N, filename, verbose, reference, generate, SOCP_SETTINGS, "socp", _make)


def test_LASSO(n_values, verbose=False, N=10, filename=None, reference=None):
    """testSolver.py:1036-1292: two files, ``<name>Times.csv`` and ``<name>Values.csv`` (30 objectives per run)."""
    np.random.seed(1)
    n_values = np.array(n_values, dtype=np.int32)
    num_tests, P = len(n_values), LASSO_PROBLEMS
    t_gpu, t_cpu = np.zeros((num_tests, N)), np.zeros((num_tests, N))
    v_gpu, v_cpu = np.zeros((num_tests, N, P)), np.zeros((num_tests, N, P))
    for count, n in enumerate(n_values):
        m = int(0.8 * n)
        for i in range(_repetitions(n, N)):
            rows, nnz = 3 * m, int(n * P / 4)
            A = np.random.rand(rows, int(n))
            x_true = np.zeros((int(n), P))
            x_true[np.unravel_index(np.random.randint(0, n * P, nnz), (int(n), P))] = np.random.uniform(0, 50, nnz)
            reg = 0.05 + 0.01 * np.random.randn(P)
            b = A @ x_true + np.random.randn(rows, P)
            s = LassoSolver(A=A, b=b, reg=reg, max_iters=5000, use_gpu=True, **LASSO_SETTINGS)
            t_gpu[count, i], out = _timed_solve(s)
            v_gpu[count, i] = np.asarray(out[1])
            if verbose:
                print(f"LASSO n={n} rep {i}: B200 {t_gpu[count, i]:.3f} s, {out[3]} ADMM iterations")
            if reference is not None:
                r = reference["LassoSolver"](A=A.copy(), b=b, reg=reg, max_iters=1000, use_gpu=False, **LASSO_SETTINGS)
                t_cpu[count, i], out = _timed_solve(r)
                v_cpu[count, i] = np.asarray(out[1])
    if filename is not None:
        z = np.zeros(num_tests * N)
        _write(filename[:-4] + "Times.csv", num_tests, N,
               {"n_values": np.repeat(n_values, N), "cvxpy_times": z, "lasso_gpu_times": t_gpu.ravel(),
                "lasso_cpu_times": t_cpu.ravel(), "lasso_jax_times": z})
        zz = np.zeros(num_tests * N * P)
        with open(filename[:-4] + "Values.csv", "w") as f:  # no num_tests,N line in the values file (:1277-1287)
            f.write("cvxpy_values,lasso_gpu_values,lasso_cpu_values,lasso_jax_values\n")
            for a, g, c_, j in zip(zz, v_gpu.ravel(), v_cpu.ravel(), zz):
                f.write(f"{a!r},{float(g)!r},{float(c_)!r},{j!r}\n")
    return t_gpu, v_gpu, t_cpu, v_cpu


def test_all_solvers(n_values, verbose=False, N=10, filename=None, reference=None):
    """testSolver.py:1294-1301 (all four enabled)."""
    test_LP(n_values, verbose, N, filename + "LP.csv", reference)
    test_QP(n_values, verbose, N, filename + "QP.csv", reference)
    test_SOCP(n_values, verbose, N, filename + "SOCP.csv", reference)
    test_LASSO(n_values, verbose, N, filename + "LASSO.csv", reference)


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--out", required=True, help="file name prefix, e.g. results/b200_")
    ap.add_argument("--n", type=int, nargs="+", default=[100, 200, 300, 400, 500])
    ap.add_argument("--N", type=int, default=3)
    ap.add_argument("--reference-dir", default=None)
    ap.add_argument("--quiet", action="store_true")
    a = ap.parse_args()
    test_all_solvers(a.n, not a.quiet, a.N, a.out, load_reference(a.reference_dir))


if __name__ == "__main__":
    main()
