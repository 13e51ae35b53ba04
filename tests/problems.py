"""Seeded synthetic problem generators shared by the golden generator, the tests and bench.py.

The distributions mirror the reference's own harness: ``testSolver.py:76-91`` (LP),
``:502-522`` (QP), ``:1096-1104`` (Lasso), ``demo.ipynb`` cell 33 / ``time_profiling.py:54-70``
(SOCP) and the BASELINE-family generators of SURVEY.md section 8(d) / Appendix A.
Everything is a pure function of the seed, so fixtures store only the seed and the expected outputs.
"""

import numpy as np

LP_TEST_SETTINGS = dict(epsilon=1e-4, mu=15, t0=1, max_inner_iters=20, max_outer_iters=10, beta=0.5,
                        alpha=0.05)  # testSolver.py:172-190
QP_TEST_SETTINGS = dict(epsilon=1e-8, mu=15, t0=0.01, max_inner_iters=100, max_outer_iters=10, beta=0.6,
                        alpha=0.4)  # testSolver.py:607-626
SOCP_TEST_SETTINGS = dict(epsilon=1e-4, mu=15, t0=0.1, max_inner_iters=500, max_outer_iters=10, beta=0.5,
                          alpha=0.05)  # testSolver.py:924-945
LASSO_TEST_SETTINGS = dict(rho=0.4, max_iters=1000, check_stop=10, add_bias=True, normalize_A=False, positive=False,
                           eps_rel=1e-6, eps_abs=1e-6)  # testSolver.py:1186-1203


def lp_testsolver(seed=1, n=100, m=80, k=20, count=1):
    """``count`` consecutive instances of test_LP's stream (A, C, x_feas, c drawn in that order)."""
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(count):
        A = rs.uniform(-2, 2, (m, n))
        C = rs.uniform(-2, 2, (k, n))
        x_feas = rs.uniform(-2, 2, n)
        c = rs.uniform(-2, 2, n)
        out.append(dict(c=c, A=A, b=A @ x_feas, C=C, d=C @ x_feas, lower_bound=-3, upper_bound=3))
    return out


def qp_testsolver(seed=1, n=100, m=80, k=20, count=1):
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(count):
        Pp = rs.uniform(-2, 2, (m, n))
        P = Pp.T @ Pp + np.eye(n)
        A = rs.uniform(-2, 2, (m, n))
        C = rs.uniform(-2, 2, (k, n))
        x_feas = rs.uniform(-2, 2, n)
        q = rs.uniform(-2, 2, n)
        out.append(dict(P=P, q=q, A=A, b=A @ x_feas, C=C, d=C @ x_feas, lower_bound=-3, upper_bound=3))
    return out


def lasso_testsolver(seed=1, n=100, m=80, num_problems=30):
    rs = np.random.RandomState(seed)
    rows = 3 * m
    nnz = int(n * num_problems / 4)
    A = rs.rand(rows, n)
    x_true = np.zeros((n, num_problems))
    x_true[np.unravel_index(rs.randint(0, n * num_problems, nnz), (n, num_problems))] = rs.uniform(0, 50, nnz)
    reg = 0.05 + 0.01 * rs.randn(num_problems)
    b = A @ x_true + rs.randn(rows, num_problems)
    return dict(A=A, b=b, reg=reg)


def lasso_cfg5(K, n=512, rows=2048, seed=5):
    """BASELINE configs[4]: A 2048 x 512 (+ bias added by the solver), K problems, generator of testSolver.py:1096-1104."""
    rs = np.random.RandomState(seed)
    A = rs.rand(rows, n)
    nnz = int(n * K / 4)
    x_true = np.zeros((n, K))
    x_true[np.unravel_index(rs.randint(0, n * K, nnz), (n, K))] = rs.uniform(0, 50, nnz)
    reg = 0.05 + 0.01 * rs.randn(K)
    b = A @ x_true + rs.randn(rows, K)
    return A, b, reg


def lp_dense_family(seed, n, m=None, warm=False):
    """BASELINE cfg-2 family: inequality-only dense LP with box +-3."""
    m = 2 * n if m is None else m
    rs = np.random.RandomState(seed)
    C = rs.uniform(-2, 2, (m, n))
    x_feas = rs.uniform(-2, 2, n)
    c = rs.uniform(-2, 2, n)
    d = C @ x_feas + rs.uniform(0.1, 1.0, m)
    p = dict(c=c, C=C, d=d, lower_bound=-3, upper_bound=3)
    if warm:
        p["x0"] = x_feas.copy()
    return p


def qp_dense_family(seed, n, p, k, gram=None, with_feasible_point=False):
    """BASELINE cfg-3 family: P = Pp'Pp/n + I, p equalities, k inequalities.
    ``gram`` (optional) computes Pp'Pp -- the full-size tests pass a device matmul, NumPy needs ~10 s at n = 8192."""
    rs = np.random.RandomState(seed)
    Pp = rs.uniform(-2, 2, (n // 2, n))
    P = (Pp.T @ Pp if gram is None else gram(Pp)) / n + np.eye(n)
    A = rs.uniform(-2, 2, (p, n))
    C = rs.uniform(-2, 2, (k, n))
    x_feas = rs.uniform(-2, 2, n)
    q = rs.uniform(-2, 2, n)
    d = C @ x_feas + rs.uniform(0.1, 1, k)
    out = dict(P=P, q=q, A=A, b=A @ x_feas, C=C, d=d, lower_bound=-3, upper_bound=3)
    if with_feasible_point:
        out["x_feas"] = x_feas  # strictly feasible by construction (not a solver argument: pop it before the call)
    return out


def socp_family(seed, n, M, k, p=0, margin=1.0, identity_P=True, warm=True):
    """BASELINE cfg-4 family: M dense cones of k rows, optional p equalities."""
    rs = np.random.RandomState(seed)
    q = rs.randn(n)
    x0 = rs.randn(n)
    A, b, c, d = [], [], [], []
    for _ in range(M):
        Ai = rs.randn(k, n)
        bi = rs.randn(k)
        ci = rs.randn(n)
        A.append(Ai), b.append(bi), c.append(ci)
        d.append(float(np.linalg.norm(Ai @ x0 + bi, 2) - ci @ x0 + margin))
    out = dict(P=np.eye(n) if identity_P else None, q=q, A=A, b=b, c=c, d=d, lower_bound=None, upper_bound=None)
    if p:
        F = rs.randn(p, n)
        out.update(F=F, g=F @ x0)
    if warm:
        out["x0"] = x0.copy()
    return out


def lp_bounds_only(seed, n, p=0):
    """Bounds-only LP (diagonal Hessian paths, LPSolver.py:442-446), optional equalities."""
    rs = np.random.RandomState(seed)
    c = rs.uniform(-2, 2, n)
    out = dict(c=c, lower_bound=-3, upper_bound=3)
    if p:
        A = rs.uniform(-2, 2, (p, n))
        out.update(A=A, b=A @ rs.uniform(-2, 2, n))
    return out


def lp_small_polytope(seed=3, n=12, infeasible=False):
    """Small LP over random half-spaces and the box [-3, 3]^n; `infeasible` adds  sum x <= 1  and  sum x >= 2  (empty set:
    the phase-I ValueError of LPSolver.py:553-558).  Without them the default x0 (box midpoint 0) is strictly feasible
    (d >= 1), phase-I is skipped, and the main loop updates the solver's own iterate in place."""
    rs = np.random.RandomState(seed)
    C = rs.uniform(-1, 1, (6, n))
    d = rs.uniform(1, 2, 6)
    c = rs.uniform(-1, 1, n)
    if infeasible:
        C = np.vstack([C, np.ones((1, n)), -np.ones((1, n))])
        d = np.concatenate([d, [1.0], [-2.0]])
    return dict(c=c, C=C, d=d, lower_bound=-3, upper_bound=3)
