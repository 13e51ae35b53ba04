"""Parity at BASELINE.json's FULL sizes (configs[1]..[4]) through size-independent properties.

The oracle cannot run these sizes in test time, so each solve is certified by convex duality: an independent
textbook barrier method (tests/barrier_polish.py, torch/cuBLAS/cuSOLVER float64, exact Hessians, none of the
reference's quirks) follows the central path from the generator's strictly feasible point and yields a dual-feasible
point, i.e. a rigorous lower bound on the optimum (tests/certificates.py); the product's point is checked to be
strictly feasible (upper bound).  `value - lower <= 1e-6 |value|` then proves the product's optimum is within
north_star's 1e-6 relative of the true one.  Everything goes through the drop-in classes, i.e. the C-ABI."""

import json
import os
import time

import numpy as np
import pytest
import torch

import barrier_polish as polish
import certificates as cert
import problems

pytestmark = pytest.mark.gpu

REL = 1e-6  # north_star: optimal objective within 1e-6 relative


def _gram_on_device(Pp):
    t = torch.as_tensor(Pp).to("cuda")
    return (t.T @ t).cpu().numpy()


def _dev(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64)).to("cuda")


def _certify(label, val, own, bar, xc, tc, w=None, rel=REL):
    """`own`: feasibility / objective of the product's point; (xc, tc, w): centred point of the independent barrier
    solve, whose bracket [lower, upper] contains the true optimum."""
    upper, lower = polish.bracket(bar, tc, xc, w, refine=bar.M == 0)
    print(label, "product value %.12g | independent bracket [%.12g, %.12g] | value - lower = %.3g (%.2g relative)" % (
        val, lower, upper, val - lower, (val - lower) / abs(val)))
    assert own["min_slack"] > 0.0                           # the product's point is strictly feasible ...
    assert own["primal"] == pytest.approx(val, rel=1e-10)   # ... and its reported value is its true objective
    assert upper - lower <= 0.5 * REL * abs(val)            # the independent bracket itself is tight
    assert val >= lower - 1e-9 * abs(val)                   # weak duality
    assert val - lower <= rel * abs(val)                    # north_star: optimum within 1e-6 relative


def test_cfg2_dense_lp_full_size():
    """configs[1]: dense LP n = 8192, m = 16384 (+ box), reference defaults; warm start (bench.py's workload) and
    cold start (default x0 = box midpoint is infeasible for C x <= d, so phase-I with its (n+1)-systems runs)."""
    from ipm_b200.LPSolver import LPSolver

    prob = problems.lp_dense_family(seed=8192, n=8192, m=16384, warm=True)
    s = LPSolver(**prob, check_cvxpy=False, suppress_print=True)
    t0 = time.perf_counter()
    val = s.solve()
    print("cfg2 warm: %.2f s, main %s, value %.12g" % (time.perf_counter() - t0, s.inner_iters, val))
    assert s.optimality_gap < s.epsilon
    cold = {k: v for k, v in prob.items() if k != "x0"}
    sc = LPSolver(**cold, check_cvxpy=False, suppress_print=True)
    t0 = time.perf_counter()
    val_cold = sc.solve()
    print("cfg2 cold: %.2f s, phase-I %s, main %s, value %.12g" % (time.perf_counter() - t0,
                                                                    sc.phase1_solver.inner_iters, sc.inner_iters, val_cold))
    assert sum(sc.phase1_solver.inner_iters) > 0
    assert val_cold == pytest.approx(val, rel=1e-9)     # two different trajectories, one optimum
    # independent textbook barrier solve from the generator's strictly feasible point
    n = 8192
    lo, hi = torch.full((n,), -3.0, dtype=torch.float64, device="cuda"), torch.full((n,), 3.0, dtype=torch.float64,
                                                                                       device="cuda")
    bar = polish.Barrier(n, q=_dev(prob["c"]), C=_dev(prob["C"]), d=_dev(prob["d"]), lo=lo, hi=hi)
    t_end = s.num_constraints / (0.3 * REL * abs(val))
    xc, _, tc = polish.follow_path(bar, _dev(prob["x0"]), 1.0, t_end, verbose=True)
    _certify("cfg2", val, cert.lp_point(prob, np.asarray(s.xstar)), bar, xc, tc)
    _certify("cfg2-cold", val_cold, cert.lp_point(prob, np.asarray(sc.xstar)), bar, xc, tc)


def test_cfg3_qp_equalities_full_size():
    """configs[2]: QP n = 8192, p = 2048 equalities (infeasible-start Newton + Schur complement), 20 inequalities
    (phase-I), box; test_QP settings (testSolver.py:563-582)."""
    from ipm_b200.QPSolver import QPSolver

    prob = problems.qp_dense_family(seed=3, n=8192, p=2048, k=20, gram=_gram_on_device, with_feasible_point=True)
    x_feas = prob.pop("x_feas")
    s = QPSolver(**prob, check_cvxpy=False, suppress_print=True, get_dual_variables=True, **problems.QP_TEST_SETTINGS)
    t0 = time.perf_counter()
    val = s.solve()
    print("cfg3: %.2f s, phase-I %s, main %s, value %.12g" % (time.perf_counter() - t0, s.phase1_solver.inner_iters,
                                                               s.inner_iters, val))
    own = cert.qp_point(prob, np.asarray(s.xstar))
    assert own["eq_residual"] < 1e-6 * (1 + np.linalg.norm(prob["b"]))
    n = 8192
    lo, hi = torch.full((n,), -3.0, dtype=torch.float64, device="cuda"), torch.full((n,), 3.0, dtype=torch.float64,
                                                                                       device="cuda")
    bar = polish.Barrier(n, P=_dev(prob["P"]), q=_dev(prob["q"]), C=_dev(prob["C"]), d=_dev(prob["d"]), lo=lo, hi=hi,
                         E=_dev(prob["A"]), e=_dev(prob["b"]))
    t_end = s.num_constraints / (0.3 * REL * abs(val))
    xc, w, tc = polish.follow_path(bar, _dev(x_feas), 1.0, t_end, verbose=True)
    _certify("cfg3", val, own, bar, xc, tc, w)


def test_cfg4_socp_full_size():
    """configs[3]: SOCP n = 16384, 256 cones of 64 rows, P = I, warm start, test_SOCP settings
    (testSolver.py:924-945).  The reference cannot hold this size at all (O(M n^2) caches)."""
    from ipm_b200.SOCPSolver import SOCPSolver

    n, M, k = 16384, 256, 64
    prob = problems.socp_family(seed=4, n=n, M=M, k=k)
    s = SOCPSolver(**prob, check_cvxpy=False, suppress_print=True, **problems.SOCP_TEST_SETTINGS)
    t0 = time.perf_counter()
    val = s.solve()
    print("cfg4: %.2f s, main %s, value %.12g" % (time.perf_counter() - t0, s.inner_iters, val))
    own = cert.socp_point(prob, np.asarray(s.xstar))
    cones = (_dev(np.concatenate(prob["A"], axis=0)), _dev(np.concatenate(prob["b"])), _dev(np.stack(prob["c"])),
             _dev(np.asarray(prob["d"])), k)
    bar = polish.Barrier(n, q=_dev(prob["q"]), cones=cones, identity_P=True)
    t_end = 2 * M / (0.3 * REL * abs(val))
    xc, _, tc = polish.follow_path(bar, _dev(prob["x0"]), 1.0, t_end, verbose=True)
    # The REFERENCE's own SOCP arm stops ~1.5e-4 (relative) above the true optimum on this family -- its Newton
    # decrement is measured with the +cc' Hessian and ends the centering steps early; measured with the
    # oracle at n = 512 in tests/test_oracle_golden.py::test_reference_socp_stops_above_true_optimum.  Parity with
    # the reference therefore means the same ~1e-4 distance from the certified optimum, not 1e-6.
    _certify("cfg4", val, own, bar, xc, tc, rel=5e-4)
    upper, lower = polish.bracket(bar, tc, xc)
    assert val - lower >= 2e-5 * abs(val)   # ... and not suspiciously better than the reference's algorithm either


def test_cfg5_lasso_batch_full_size():
    """configs[4]: 4096 Lasso problems, A 2048 x 512 (+ bias), GPU-arm settings (testSolver.py:1142-1159)."""
    from ipm_b200.LassoSolver import LassoSolver

    A, b, reg = problems.lasso_cfg5(4096)
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "lasso_full.json")))
    s = LassoSolver(A, b, reg, rho=0.4, check_stop=10, add_bias=True, check_cvxpy=False, eps_abs=1e-6, eps_rel=1e-6,
                    max_iters=5000)
    X, sol, _, its = s.solve()
    primal, gap = cert.lasso_gaps(A, b, reg, np.asarray(X))
    print("cfg5: %d ADMM iterations, max relative duality gap %.3g, median %.3g" % (
        its, np.max(gap / primal), np.median(gap / primal)))
    np.testing.assert_allclose(np.asarray(sol), primal, rtol=1e-9)   # reported objectives are the true objectives
    assert np.all(gap >= -1e-9 * primal)
    assert np.max(gap / primal) < 1e-2 and np.median(gap / primal) < 1e-3   # ADMM stops on its residual tests
    assert its < 5000
    # against the REAL reference at full size (tests/golden/lasso_full.json, generate_golden.py --lasso-full-only): the
    # batch-coupled stop test fires at the same check, and the solutions agree
    assert its == gold["iterations"]
    np.testing.assert_allclose(np.asarray(sol)[:16], gold["solutions_head"], rtol=1e-9)
    assert float(np.sum(sol)) == pytest.approx(gold["solutions_sum"], rel=1e-10)
    assert np.linalg.norm(X) == pytest.approx(gold["X_frob"], rel=1e-9)
    assert np.abs(np.asarray(X)).sum() == pytest.approx(gold["X_abs_sum"], rel=1e-9)
    np.testing.assert_allclose(np.asarray(X)[:, 0], gold["X_col0"], rtol=1e-6, atol=1e-9)
