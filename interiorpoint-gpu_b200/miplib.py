"""Loader for the reference's MIPLIB ``.npy`` problem files (``testSolver.py:278-300``): seven consecutive
``np.save`` arrays ``c, A, b, C, d, up_bnd, lo_bnd`` in one file, all dense float64 (the matrices are >99 % zeros),

    minimise c'x   s.t.  A x = b,  C x <= d,  lo_bnd <= x <= up_bnd.

``load_lp`` returns the keyword arguments of ``LPSolver`` (the reference passes exactly these, testSolver.py:338-356);
``save_lp`` writes the same format (used by the tests for the synthetic stand-in of the missing
``example_data/aflow40b.npy``, SURVEY.md 8(d) cfg 1).  ``LPSolver(sparse="auto")`` then keeps ``C`` as CSR on the
device and forms the Hessian entry-wise (``engine.SparseRows``).
"""

import numpy as np

FIELDS = ("c", "A", "b", "C", "d", "upper_bound", "lower_bound")


def load_lp(path):
    """-> dict(c, A, b, C, d, upper_bound, lower_bound); empty equality / inequality blocks become None."""
    out = {}
    with open(path, "rb") as f:
        for name in FIELDS:
            out[name] = np.load(f)
    n = out["c"].shape[0]
    for M, v in (("A", "b"), ("C", "d")):
        if out[M].size == 0:
            out[M] = out[v] = None
        elif out[M].ndim != 2 or out[M].shape[1] != n or out[M].shape[0] != out[v].shape[0]:
            raise ValueError(f"{path}: {M} / {v} do not match the {n} variables of c")
    for bnd in ("upper_bound", "lower_bound"):
        if out[bnd].shape not in ((), (n,)):
            raise ValueError(f"{path}: {bnd} must be a scalar or have one entry per variable")
    return out


def save_lp(path, c, A, b, C, d, upper_bound, lower_bound):
    n = len(c)
    with open(path, "wb") as f:
        for a in (c, np.zeros((0, n)) if A is None else A, np.zeros(0) if b is None else b,
                  np.zeros((0, n)) if C is None else C, np.zeros(0) if d is None else d, upper_bound, lower_bound):
            np.save(f, np.asarray(a, dtype=np.float64))


def synthetic_network_lp(seed=40, n=2728, p=78, m=1364, density=0.0017):
    """Stand-in with the published shape of MIPLIB aflow40b (n = 2728 variables, ~1442 constraint rows, ~0.17 %
    non-zeros; the split into p equalities and m inequalities is arbitrary): sparse rows with +-1 /
    uniform coefficients, box [0, 1], feasible by construction with a strictly interior point for C."""
    rs = np.random.RandomState(seed)

    def sparse_rows(rows):
        Mx = np.zeros((rows, n))
        k = max(2, int(round(density * n)))
        for r in range(rows):
            cols = rs.choice(n, size=k, replace=False)
            Mx[r, cols] = rs.choice([-1.0, 1.0], size=k) * rs.uniform(0.5, 2.0, size=k)
        return Mx

    A, C = sparse_rows(p), sparse_rows(m)
    x_feas = rs.uniform(0.2, 0.8, n)
    c = rs.uniform(-1.0, 1.0, n)
    return dict(c=c, A=A, b=A @ x_feas, C=C, d=C @ x_feas + rs.uniform(0.05, 0.5, m), upper_bound=np.ones(n),
                lower_bound=np.zeros(n))
