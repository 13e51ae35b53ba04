"""Outer barrier loops (LP / QP / SOCP) and the integrated phase-I driver -- oracle restatement.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``LPSolver.solve`` (LPSolver.py:514-653), ``QPSolver.solve`` (QPSolver.py:500-638),
``SOCPSolver.solve`` (SOCPSolver.py:616-753), the Newton-class dispatch (LPSolver.py:371-469,
QPSolver.py:385-455, SOCPSolver.py:485-555; only the default ``"cholesky"`` method and its
diagonal specialisations) and ``PhaseOneSolver`` (PhaseOneSolver.py:6-154).
"""

import numpy as np

from .barrier import ConeBarrier, LinearBarrier, PhaseOneConeBarrier, PhaseOneLinearBarrier
from .newton import FeasibleNewton, InfeasibleNewton


def default_x0(n, lb, ub):
    """LPSolver.py:131-143 (same in QPSolver.py:139-151, SOCPSolver.py:154-166)."""
    if lb is not None and ub is not None:
        return (np.maximum(lb, -1e2) + np.minimum(ub, 1e2)) / 2 * np.ones(n)
    if lb is not None:
        return (np.maximum(lb, -1e2) + 1e-1) * np.ones(n)
    if ub is not None:
        return (np.minimum(ub, 1e2) - 1e-1) * np.ones(n)
    return np.random.rand(n)


class OraclePhaseOne:
    """PhaseOneSolver.py:6-154."""

    def __init__(self, barrier, x0, n, max_outer_iters, max_inner_iters, epsilon, inner_epsilon, alpha, beta, mu,
                 t0, tol, use_psd_condition=False, update_slacks_every=0, trace=None):
        self.fm = barrier
        self.n = n
        self.max_outer_iters = max_outer_iters
        self.epsilon, self.mu, self.t0, self.tol = epsilon, mu, t0, tol
        self.z = np.append(x0, barrier.s)  # PhaseOneSolver.py:86-89
        self.ns = FeasibleNewton(barrier, max_iters=max_inner_iters, epsilon=inner_epsilon, alpha=alpha, beta=beta,
                                 phase1_flag=True, phase1_tol=tol, use_psd_condition=use_psd_condition,
                                 update_slacks_every=update_slacks_every, trace=trace)
        self.outer_iters = 0
        self.inner_iters = []

    def solve(self):
        t = self.t0
        self.outer_iters = 0
        self.inner_iters = []
        obj = None
        for _ in range(self.max_outer_iters):
            self.z, k, _, _ = self.ns.solve(self.z)
            self.outer_iters += 1
            self.inner_iters.append(k)
            self.fm.move(self.z)
            obj = self.fm.objective()
            if obj < -self.tol:
                break
            t = min(t * self.mu, (self.n + 1.0) / self.epsilon)  # PhaseOneSolver.py:151
            self.fm.set_t(t)
        return self.z[:-1], obj


class _BarrierSolver:
    """Shared outer loop; subclasses provide the barrier, the equality pair and the tolerance."""

    eq_tol_scale_n = False

    def _common(self, n, lb, ub, t0, max_outer_iters, max_inner_iters, phase1_max_inner_iters, epsilon,
                inner_epsilon, alpha, beta, mu, phase1_tol, phase1_t0, x0, update_slacks_every, use_psd_condition,
                trace):
        self.n = n
        self.lb = None if lb is None else np.array(lb)
        self.ub = None if ub is None else np.array(ub)
        self.x = default_x0(n, self.lb, self.ub) if x0 is None else x0
        self.t0, self.mu = t0, mu
        self.max_outer_iters, self.max_inner_iters = max_outer_iters, max_inner_iters
        self.phase1_max_inner_iters = phase1_max_inner_iters
        self.epsilon, self.inner_epsilon = epsilon, inner_epsilon
        self.alpha, self.beta = alpha, beta
        self.phase1_tol, self.phase1_t0 = phase1_tol, phase1_t0
        self.update_slacks_every = update_slacks_every
        self.use_psd_condition = use_psd_condition
        self.trace = trace
        self.phase1 = None
        self.outer_iters, self.inner_iters = 0, []
        self.value = self.xstar = self.optimality_gap = None

    def _make_phase1(self, barrier):
        return OraclePhaseOne(barrier, self.x, self.n, self.max_outer_iters, self.phase1_max_inner_iters,
                              self.epsilon, self.inner_epsilon, self.alpha, self.beta, self.mu, self.phase1_t0,
                              self.phase1_tol, self.use_psd_condition, self.update_slacks_every,
                              None if self.trace is None else self.trace.setdefault("phase1", []))

    def _make_newton(self, diagonal=False):
        tr = None if self.trace is None else self.trace.setdefault("main", [])
        if self.E is not None:
            return InfeasibleNewton(self.fm, self.E, self.e, max_iters=self.max_inner_iters, epsilon=self.inner_epsilon,
                                    alpha=self.alpha, beta=self.beta, use_psd_condition=self.use_psd_condition,
                                    update_slacks_every=self.update_slacks_every, diagonal=diagonal, trace=tr)
        return FeasibleNewton(self.fm, max_iters=self.max_inner_iters, epsilon=self.inner_epsilon, alpha=self.alpha,
                              beta=self.beta, use_psd_condition=self.use_psd_condition,
                              update_slacks_every=self.update_slacks_every, diagonal=diagonal, trace=tr,
                              linear_solver="cg" if getattr(self, "linear_solve_method", "") == "cg" else "cholesky",
                              max_cg_iters=getattr(self, "max_cg_iters", 50))

    def solve(self, t0=None, max_outer_iters=None):
        t = self.t0 if t0 is None else t0
        max_outer = self.max_outer_iters if max_outer_iters is None else max_outer_iters
        x = self.x
        if self.phase1 is not None and self.phase1.fm.s >= 1:  # LPSolver.py:546-560
            x, s = self.phase1.solve()
            if s > -self.phase1_tol:
                raise ValueError("Phase 1 Solver did not successfully find a feasible point!")
        self.outer_iters, self.inner_iters = 0, []
        self.objective_vals = []
        self.fm.move(x)
        self.fm.set_t(t)
        v = np.zeros(self.E.shape[0]) if self.E is not None else None
        gap = self.num_constraints
        best_x, best_obj = x.copy(), np.inf
        for _ in range(max_outer):
            if self.E is not None:
                x, v, k, _, ok = self.ns.solve(x, v0=v)
            else:
                x, k, _, ok = self.ns.solve(x)
            self.outer_iters += 1
            self.inner_iters.append(k)
            tol = 1e-4 * self.n if self.eq_tol_scale_n else 1e-3  # LPSolver.py:600-602 vs QPSolver.py:585-587
            if self.E is None or np.linalg.norm(np.matmul(self.E, x) - self.e) < tol:
                obj = self.fm.objective()
                self.objective_vals.append(obj)
                if obj < best_obj:
                    best_obj, best_x = obj, x.copy()
                elif ok:
                    break
            gap = self.num_constraints / t
            if gap < self.epsilon:
                break
            t = t * self.mu
            self.fm.set_t(t)
        self.xstar, self.value, self.optimality_gap, self.v = best_x, best_obj, gap, v
        self.t_final = t
        return self.value

    def dual_variables(self):
        """(lam_star, v_star) of LPSolver.py:641-646 / QPSolver.py:626-631: 1 / (t slacks(x*)) in the slack layout
        [C rows | upper bounds | lower bounds] and v / t, both with the LAST t of the outer loop (already multiplied by
        mu when the loop ran out of iterations instead of meeting the gap test)."""
        lam = None
        if self.num_constraints > 0:
            self.fm.move(self.xstar)
            lam = 1 / (self.t_final * self.fm.slacks)
        nu = None if self.E is None else self.v / self.t_final
        return lam, nu


class OracleLP(_BarrierSolver):
    eq_tol_scale_n = True

    def __init__(self, c=None, A=None, b=None, C=None, d=None, lower_bound=0, upper_bound=None, t0=0.1,
                 max_outer_iters=20, max_inner_iters=50, phase1_max_inner_iters=500, epsilon=1e-10,
                 inner_epsilon=1e-5, alpha=0.2, beta=0.6, mu=15, try_diag=True, phase1_tol=0, phase1_t0=0.01,
                 x0=None, update_slacks_every=0, trace=None, linear_solve_method="cholesky", max_cg_iters=50):
        self.linear_solve_method, self.max_cg_iters = linear_solve_method, max_cg_iters  # "cg": LPSolver.py:412-420
        n = len(c) if c is not None else (A.shape[1] if A is not None else C.shape[1])
        self._common(n, lower_bound, upper_bound, t0, max_outer_iters, max_inner_iters, phase1_max_inner_iters,
                     epsilon, inner_epsilon, alpha, beta, mu, phase1_tol, phase1_t0, x0, update_slacks_every, False,
                     trace)
        self.E, self.e = A, b
        self.num_constraints = (0 if d is None else len(d)) + (n if self.lb is not None else 0) + (
            n if self.ub is not None else 0)
        if C is not None:
            self.phase1 = self._make_phase1(PhaseOneLinearBarrier(C, d, self.x, self.lb, self.ub, t=phase1_t0))
        self.fm = LinearBarrier(n, c=c, C=C, d=d, lb=self.lb, ub=self.ub, t=1, try_diag=try_diag)
        self.ns = self._make_newton(diagonal=(C is None and try_diag))  # LPSolver.py:436-446


class OracleQP(_BarrierSolver):
    def __init__(self, P=None, q=None, A=None, b=None, C=None, d=None, lower_bound=0, upper_bound=None, t0=0.1,
                 max_outer_iters=20, max_inner_iters=50, phase1_max_inner_iters=500, epsilon=1e-10,
                 inner_epsilon=1e-5, alpha=0.2, beta=0.6, mu=15, phase1_tol=0, phase1_t0=0.01, x0=None,
                 update_slacks_every=0, trace=None):
        if P is None:
            raise ValueError("Setting P to None is just an LP! Please use LP solver or set a value to P.")
        n = len(q) if q is not None else P.shape[1]
        self._common(n, lower_bound, upper_bound, t0, max_outer_iters, max_inner_iters, phase1_max_inner_iters,
                     epsilon, inner_epsilon, alpha, beta, mu, phase1_tol, phase1_t0, x0, update_slacks_every, False,
                     trace)
        self.E, self.e = A, b
        self.num_constraints = (0 if d is None else len(d)) + (n if self.lb is not None else 0) + (
            n if self.ub is not None else 0)
        if C is not None:
            self.phase1 = self._make_phase1(PhaseOneLinearBarrier(C, d, self.x, self.lb, self.ub, t=phase1_t0))
        self.fm = LinearBarrier(n, P=P, q=q, C=C, d=d, lb=self.lb, ub=self.ub, t=1)
        self.ns = self._make_newton()


class OracleSOCP(_BarrierSolver):
    def __init__(self, P=None, q=None, A=None, b=None, c=None, d=None, F=None, g=None, lower_bound=0,
                 upper_bound=None, t0=0.1, phase1_t0=0.01, max_outer_iters=20, max_inner_iters=50,
                 phase1_max_inner_iters=500, epsilon=1e-10, inner_epsilon=1e-5, alpha=0.2, beta=0.6, mu=15,
                 phase1_tol=0, use_psd_condition=False, x0=None, update_slacks_every=0, trace=None):
        A, b, c, d = self._normalise_cones(A, b, c, d)
        n = len(q) if q is not None else (P.shape[1] if P is not None else (A[0].shape[1] if A[0].ndim > 1 else len(A[0])))
        self._common(n, lower_bound, upper_bound, t0, max_outer_iters, max_inner_iters, phase1_max_inner_iters,
                     epsilon, inner_epsilon, alpha, beta, mu, phase1_tol, phase1_t0, x0, update_slacks_every,
                     use_psd_condition, trace)
        self.E, self.e = F, g
        self.num_constraints = len(A) + (n if self.lb is not None else 0) + (n if self.ub is not None else 0)
        self.phase1 = self._make_phase1(PhaseOneConeBarrier(A, b, c, d, self.x, self.lb, self.ub, t=phase1_t0))
        self.fm = ConeBarrier(n, P=P, q=q, A=A, b=b, c=c, d=d, lb=self.lb, ub=self.ub, t=1)
        self.ns = self._make_newton()

    @staticmethod
    def _normalise_cones(A, b, c, d):
        """List handling and diagonal compression of ``SOCPSolver.__check_inputs`` (SOCPSolver.py:255-385)."""
        if A is None:
            raise ValueError("No cone contraints detected. Run with LPSolver or QPSolver for better performance.")
        A = list(A) if isinstance(A, list) else [A]
        for i, Ai in enumerate(A):
            if Ai.ndim == 2:
                off = Ai.copy()
                np.fill_diagonal(off, 0)
                if (off == 0).all():
                    A[i] = np.diag(Ai).copy()
        if b is not None:
            b = list(b) if isinstance(b, list) else [b]
            if len(b) == 1:
                b = b * len(A)
        if c is not None:
            c = list(c) if isinstance(c, list) else [c]
        if d is not None:
            d = list(d) if isinstance(d, list) else [d]
            if len(d) == 1:
                d = d * len(A)
        return A, b, c, d
