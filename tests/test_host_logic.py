"""Host-side logic that needs no GPU: the reference's step-size sequence, default starting points, input
validation (same ValueErrors as LPSolver.py:226-318 / QPSolver / SOCPSolver) and method dispatch."""

import numpy as np
import pytest

from ipm_b200 import _solver_base as sb
from ipm_b200.engine import step_table


def test_step_table_is_the_reference_sequence():
    for beta in (0.5, 0.6, 0.7, 0.9):
        tab = step_table(beta)
        a, ref = 1, [1.0]
        while True:  # NewtonSolver.py:175-176
            a *= beta
            ref.append(a)
            if a < 1e-13:
                break
        assert tab == ref
        assert tab[-1] < 1e-13 <= tab[-2]
    with pytest.raises(ValueError):
        step_table(1.0)


def test_default_x0_matches_reference_rules():
    n = 4
    np.testing.assert_array_equal(sb.default_x0(n, np.array(-3.0), np.array(3.0)), np.zeros(n))
    np.testing.assert_allclose(sb.default_x0(n, np.array(0.0), None), 0.1 * np.ones(n))
    np.testing.assert_allclose(sb.default_x0(n, None, np.array(1e9)), (1e2 - 1e-1) * np.ones(n))
    np.random.seed(3)
    want = np.random.rand(n)
    np.random.seed(3)
    np.testing.assert_array_equal(sb.default_x0(n, None, None), want)


def test_lp_validation_errors_come_before_any_device_work():
    from ipm_b200.LPSolver import LPSolver

    c = np.ones(3)
    with pytest.raises(ValueError, match="Both A and b"):
        LPSolver(c=c, A=np.eye(3), check_cvxpy=False)
    with pytest.raises(ValueError, match="Both C and d"):
        LPSolver(c=c, d=np.ones(3), check_cvxpy=False)
    with pytest.raises(ValueError, match="agreeing dimensions"):
        LPSolver(c=c, C=np.eye(3), d=np.ones(2), check_cvxpy=False)
    with pytest.raises(ValueError, match="1-dimensional"):
        LPSolver(c=np.ones((3, 1)), check_cvxpy=False)
    with pytest.raises(ValueError, match="Lower bound must be lower"):
        LPSolver(c=c, C=np.eye(3), d=np.ones(3), lower_bound=2, upper_bound=1, check_cvxpy=False)
    with pytest.raises(ValueError, match="same number of entries"):
        LPSolver(c=c, C=np.ones((2, 4)), d=np.ones(2), check_cvxpy=False)


def test_qp_and_socp_validation():
    from ipm_b200.QPSolver import QPSolver
    from ipm_b200.SOCPSolver import SOCPSolver

    with pytest.raises(ValueError, match="just an LP"):
        QPSolver(P=None, q=np.ones(2))
    with pytest.raises(ValueError, match="square"):
        QPSolver(P=np.ones((2, 3)), q=np.ones(3))
    with pytest.raises(ValueError, match="No cone"):
        SOCPSolver(P=np.eye(2), q=np.ones(2), A=None)
    with pytest.raises(ValueError, match="equal number of A and b"):
        SOCPSolver(q=np.ones(2), A=[np.ones((3, 2)), np.ones((3, 2))], b=[np.ones(3)] * 3, c=[np.ones(2)] * 2,
                   d=[1.0, 1.0])
    with pytest.raises(ValueError, match="d must be a scalar"):
        SOCPSolver(q=np.ones(2), A=[np.ones((3, 2))], b=[np.ones(3)], c=[np.ones(2)], d=[np.ones(2)])


def test_method_dispatch_rules():
    chk = sb.BarrierSolverBase._check_method
    for m in ("cholesky", "np_solve", "np_lstsq", "direct"):
        chk(m, False)
        chk(m, True)
    chk("kkt", True)
    with pytest.raises(ValueError, match="No KKT"):  # LPSolver.py:423-430
        chk("kkt", False)
    with pytest.raises(ValueError, match="valid linear solve"):  # LPSolver.py:447-448
        chk("qr", False)
    with pytest.raises(NotImplementedError):  # NewtonSolverInfeasibleStart.py:604
        chk("cg", True)


def test_host_array_answers_get():
    a = sb.HostArray(np.arange(3.0))
    np.testing.assert_array_equal(a.get(), np.arange(3.0))
    assert float(a.sum()) == 3.0
