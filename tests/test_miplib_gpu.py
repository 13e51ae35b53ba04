"""the MIPLIB `.npy` loader and the sparse-aware LP path (testSolver.py:278-300).

`example_data/aflow40b.npy` is missing from the reference checkout, so the file-format round trip and the solver run
on a synthetic stand-in with aflow40b's published shape (flagged as such; ipm_b200/miplib.py).  The sparse path is
checked against (i) the CPU oracle at a size it finishes quickly and (ii) our own dense path on the same file at full
stand-in size."""

import numpy as np
import pytest

import problems

pytestmark = pytest.mark.gpu

SETTINGS = problems.LP_TEST_SETTINGS  # test_LP_sparse uses the same (testSolver.py:338-356)


def test_loader_round_trip_and_sparse_vs_oracle(tmp_path):
    from ipm_b200 import miplib
    from ipm_b200.LPSolver import LPSolver
    from oracle import OracleLP

    prob = miplib.synthetic_network_lp(seed=7, n=400, p=12, m=200, density=0.01)
    path = tmp_path / "stand_in.npy"
    miplib.save_lp(path, **prob)
    loaded = miplib.load_lp(path)
    assert all(np.array_equal(prob[k], loaded[k]) for k in prob)
    o = OracleLP(**loaded, **SETTINGS)
    ref = o.solve()
    s = LPSolver(**loaded, check_cvxpy=False, suppress_print=True, **SETTINGS)
    assert s.data.sparse is not None and s.data.C is None      # auto-detected; C never densified on the device
    val = s.solve()
    print(val, ref, s.inner_iters, o.inner_iters, s.phase1_solver.inner_iters, o.phase1.inner_iters if o.phase1 else None)
    assert val == pytest.approx(ref, rel=1e-6, abs=1e-9)
    assert len(s.inner_iters) == len(o.inner_iters)
    assert all(abs(a - b) <= 2 or b >= SETTINGS["max_inner_iters"] for a, b in zip(s.inner_iters, o.inner_iters))
    assert np.linalg.norm(np.asarray(s.xstar) - np.asarray(o.xstar)) <= 1e-4 * (1 + np.linalg.norm(o.xstar))


def test_aflow40b_shaped_stand_in_sparse_equals_dense():
    from ipm_b200 import miplib
    from ipm_b200.LPSolver import LPSolver

    prob = miplib.synthetic_network_lp()  # n = 2728, 78 equalities, 1364 inequalities, 0.17 % non-zeros
    sp = LPSolver(**prob, check_cvxpy=False, suppress_print=True, **SETTINGS)
    dn = LPSolver(**prob, check_cvxpy=False, suppress_print=True, sparse=False, **SETTINGS)
    assert sp.data.sparse is not None and dn.data.sparse is None
    v_sp, v_dn = sp.solve(), dn.solve()
    print("sparse", v_sp, sp.inner_iters, "dense", v_dn, dn.inner_iters, "Hessian entries", sp.data.sparse.nout)
    assert v_sp == pytest.approx(v_dn, rel=1e-8)
    assert len(sp.inner_iters) == len(dn.inner_iters)      # same arithmetic up to summation order: counts +-2
    assert all(abs(a - b) <= 2 for a, b in zip(sp.inner_iters, dn.inner_iters))
    x = np.asarray(sp.xstar)
    assert np.all(prob["C"] @ x < prob["d"]) and np.all(x > 0) and np.all(x < 1)
    assert np.linalg.norm(prob["A"] @ x - prob["b"]) < 1e-6


def test_qp_with_sparse_inequalities_equals_dense():
    """QPSolver takes the same sparse-rows path (t*P is copied first, the entry-wise C' diag(w) C is added on top)."""
    from ipm_b200.QPSolver import QPSolver

    rs = np.random.RandomState(5)
    n, m, p = 320, 400, 24
    Pp = rs.uniform(-1, 1, (n // 2, n))
    P = Pp.T @ Pp / n + np.eye(n)
    C = np.where(rs.rand(m, n) < 0.01, rs.uniform(-2, 2, (m, n)), 0.0)
    A = rs.uniform(-2, 2, (p, n))
    x_feas = rs.uniform(-2, 2, n)
    prob = dict(P=P, q=rs.uniform(-2, 2, n), A=A, b=A @ x_feas, C=C, d=C @ x_feas + rs.uniform(0.1, 1, m),
                lower_bound=-3, upper_bound=3)
    sp = QPSolver(**prob, check_cvxpy=False, suppress_print=True, **problems.QP_TEST_SETTINGS)
    dn = QPSolver(**prob, check_cvxpy=False, suppress_print=True, sparse=False, **problems.QP_TEST_SETTINGS)
    assert sp.data.sparse is not None and dn.data.sparse is None
    v_sp, v_dn = sp.solve(), dn.solve()
    print("sparse", v_sp, sp.inner_iters, "dense", v_dn, dn.inner_iters)
    assert v_sp == pytest.approx(v_dn, rel=1e-8)
    assert len(sp.inner_iters) == len(dn.inner_iters)
    # Same arithmetic up to summation order: counts +-2.  The last centering steps of an equality-constrained solve
    # stop on a residual that is close to its own rounding noise (on this instance the reference's
    # NumPy arm needs 12 Newton steps in the last centering, the dense device path 16, the sparse one 11), so one
    # step may differ by more -- optimum and iterate must not.
    diffs = [abs(a - b) for a, b in zip(sp.inner_iters, dn.inner_iters)]
    assert sum(d > 2 for d in diffs) <= 1 and max(diffs) <= 6, (sp.inner_iters, dn.inner_iters)
    assert np.linalg.norm(np.asarray(sp.xstar) - np.asarray(dn.xstar)) <= 1e-6 * (1 + np.linalg.norm(dn.xstar))
