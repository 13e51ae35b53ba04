"""``QPSolver`` -- drop-in for the reference's ``QPSolver.py`` (constructor :20-50, ``solve`` :500-638) on the
B200 engine:  minimise 1/2 x'Px + q'x  s.t.  Ax = b,  Cx <= d,  lb <= x <= ub."""

import numpy as np
import torch
import torch.distributed as tdist

try:
    from ._solver_base import BarrierSolverBase, as_bound, check_bounds, check_pair, default_x0, HostArray
    from .engine import F64, Launcher, LinearNewton, LinearProblemData
    from .PhaseOneSolver import PhaseOneSolver
    from .dist import row_range
    from .sharded_engine import ShardedLinearNewton
except ImportError:  # flat-module use
    from _solver_base import BarrierSolverBase, as_bound, check_bounds, check_pair, default_x0, HostArray
    from engine import F64, Launcher, LinearNewton, LinearProblemData
    from PhaseOneSolver import PhaseOneSolver
    from dist import row_range
    from sharded_engine import ShardedLinearNewton


class QPSolver(BarrierSolverBase):
    def __init__(self, P=None, q=None, A=None, b=None, C=None, d=None, lower_bound=0, upper_bound=None, t0=0.1,
                 max_outer_iters=20, max_inner_iters=50, phase1_max_inner_iters=500, epsilon=1e-10,
                 inner_epsilon=1e-5, check_cvxpy=True, linear_solve_method="cholesky", max_cg_iters=50, alpha=0.2,
                 beta=0.6, mu=15, suppress_print=False, use_gpu=False, track_loss=False, get_dual_variables=False,
                 phase1_tol=0, phase1_t0=0.01, x0=None, update_slacks_every=0, shard_rows=False, sparse="auto"):
        if P is None:
            raise ValueError("Setting P to None is just an LP! Please use LP solver or set a value to P.")
        self.P, self.q, self.A, self.C, self.b, self.d = P, q, A, C, b, d
        if P.ndim != 2 or P.shape[0] != P.shape[1]:
            raise ValueError("P must be a symmetric, square PSD matrix!")
        if q is not None and (q.ndim != 1 or len(q) != P.shape[1]):
            raise ValueError("P and q must have the same dimension")
        check_pair(A, b, "A", "b")
        check_pair(C, d, "C", "d")
        self.n = len(q) if q is not None else P.shape[1]
        for M, nm in ((A, "A"), (C, "C")):
            if M is not None and M.shape[1] != self.n:
                raise ValueError(f"q must have the same number of entries as {nm} has columns!")
        self.lb, self.ub = as_bound(lower_bound, "Lower"), as_bound(upper_bound, "Upper")
        check_bounds(self.lb, self.ub, self.n)
        self.equality_constrained = A is not None
        self.bounded = self.lb is not None or self.ub is not None
        self.x = default_x0(self.n, self.lb, self.ub) if x0 is None else np.asarray(x0, dtype=np.float64)
        self._init_common(t0, mu, max_outer_iters, max_inner_iters, phase1_max_inner_iters, epsilon, inner_epsilon,
                          max_cg_iters, alpha, beta, suppress_print, track_loss, linear_solve_method,
                          get_dual_variables, phase1_t0, phase1_tol, update_slacks_every, use_gpu)
        self._check_method(linear_solve_method, self.equality_constrained)
        self.num_constraints = (0 if d is None else len(d)) + (self.n if self.lb is not None else 0) + (
            self.n if self.ub is not None else 0)
        self._eq_tol = 1e-3  # QPSolver.py:585-587
        self.launcher = Launcher(self.device)
        # optional extension (not in the reference): shard the inequality rows of ONE problem over the ranks of an
        # initialised torch.distributed group (partial Hessian + NCCL all-reduce, see sharded_engine.py)
        C_loc, d_loc, lb_loc, ub_loc, newton_cls = C, d, self.lb, self.ub, LinearNewton
        self.sharded = bool(shard_rows) and tdist.is_available() and tdist.is_initialized() and \
            tdist.get_world_size() > 1
        if self.sharded:
            if C is None:
                raise NotImplementedError("shard_rows needs inequality rows to shard")
            lo, hi = row_range(C.shape[0], tdist.get_rank(), tdist.get_world_size())
            C_loc, d_loc = C[lo:hi], d[lo:hi]
            if tdist.get_rank() != 0:
                lb_loc = ub_loc = None  # bound rows belong to rank 0
            newton_cls = ShardedLinearNewton
        self.data = LinearProblemData(self.n, self.device, P=P, q=q, C=C_loc, d=d_loc, lb=lb_loc, ub=ub_loc, A=A, b=b,
                                      sparse=False if self.sharded else sparse)
        self.x_dev = torch.as_tensor(self.x).to(device=self.device, dtype=F64).clone()
        if C is not None:
            self.phase1_solver = PhaseOneSolver(
                C=C, d=d, lower_bound=self.lb, upper_bound=self.ub, x0=self.x, max_outer_iters=max_outer_iters,
                max_inner_iters=phase1_max_inner_iters, epsilon=epsilon, inner_epsilon=inner_epsilon, alpha=alpha,
                beta=beta, mu=mu, suppress_print=suppress_print, n=self.n, tol=phase1_tol, t0=phase1_t0,
                update_slacks_every=update_slacks_every, _data=self.data, _launcher=self.launcher,
                _newton_cls=newton_cls)
        self.ns = newton_cls(self.data, phase1=False, max_iters=max_inner_iters, epsilon=inner_epsilon, alpha=alpha,
                               beta=beta, update_slacks_every=update_slacks_every, launcher=self.launcher,
                               max_cg_iters=max_cg_iters,
                               linear_solver="cg" if linear_solve_method == "cg" else "cholesky")

    def _objective_value(self, x):
        return self.ns.qp_objective(x)

    def _equality_residual(self, x):
        return self.ns.equality_residual(x)

    def _dual_variables(self, best_x, t):
        if self.C is not None or self.bounded:
            self.lam_star = HostArray((1.0 / (t * self.ns.slacks_at(best_x))).cpu().numpy())
        if self.A is not None:
            self.v_star = HostArray((self.ns.ws.v / t).cpu().numpy())
