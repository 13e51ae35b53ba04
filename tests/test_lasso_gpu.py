"""LassoSolver on the B200 engine vs the golden fixtures recorded from the real reference (and its frozen CSV)."""

import numpy as np
import pytest

import problems
from conftest import load_golden

pytestmark = pytest.mark.gpu

LASSO = load_golden("lasso_cases.json")


@pytest.mark.parametrize("case", LASSO, ids=[c["name"] for c in LASSO])
def test_lasso_matches_reference(case):
    from ipm_b200.LassoSolver import LassoSolver

    prob = problems.lasso_testsolver(**case["generator_kwargs"])
    s = LassoSolver(prob["A"], prob["b"], prob["reg"], check_cvxpy=False, **case["settings"])
    X, sol, _, its = s.solve()
    print(case["name"], its, case["iterations"], np.asarray(sol)[:3], case["solutions"][:3])
    # iteration counts: the stop test runs every check_stop iterations, so parity means the SAME check index
    assert its == case["iterations"]
    np.testing.assert_allclose(np.asarray(sol), case["solutions"], rtol=1e-9)
    np.testing.assert_allclose(np.asarray(X)[:, 0], case["X_col0"], rtol=1e-7, atol=1e-10)
    assert np.linalg.norm(X) == pytest.approx(case["X_frob"], rel=1e-9)
    assert np.abs(np.asarray(X)).sum() == pytest.approx(case["X_abs_sum"], rel=1e-9)


def test_lasso_ragged_single_problem_and_scalar_reg():
    """K = 1 (1-D b), scalar reg, no bias: shapes the reference mishandles must still work."""
    from ipm_b200.LassoSolver import LassoSolver
    from oracle import OracleLasso

    rs = np.random.RandomState(9)
    A = rs.rand(50, 7)
    b = A @ rs.rand(7) + 0.01 * rs.randn(50)
    s = LassoSolver(A, b, reg=0.01, add_bias=False, check_cvxpy=False, eps_abs=1e-8, eps_rel=1e-8, max_iters=2000)
    X, sol, _, its = s.solve()
    o = OracleLasso(A, b, [0.01], add_bias=False, eps_abs=1e-8, eps_rel=1e-8, max_iters=2000)
    Xo, solo, itso = o.solve()
    assert its == itso
    np.testing.assert_allclose(np.asarray(sol), solo, rtol=1e-10)
    np.testing.assert_allclose(np.asarray(X), Xo, rtol=1e-7, atol=1e-10)
