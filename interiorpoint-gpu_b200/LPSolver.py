"""``LPSolver`` -- drop-in for the reference's ``LPSolver.py`` (constructor :20-50, ``solve`` :514-653) running
on the B200 engine:  minimise c'x  s.t.  Ax = b,  Cx <= d,  lb <= x <= ub  by a log-barrier interior-point
method with Cholesky Newton steps (feasible start, or infeasible start when ``A`` is given) and a phase-I."""

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


class LPSolver(BarrierSolverBase):
    def __init__(self, c=None, A=None, b=None, C=None, d=None, lower_bound=0, upper_bound=None, t0=0.1,
                 max_outer_iters=20, max_inner_iters=50, phase1_max_inner_iters=500, epsilon=1e-10,
                 inner_epsilon=1e-5, check_cvxpy=True, linear_solve_method="cholesky", max_cg_iters=50, alpha=0.2,
                 beta=0.6, mu=15, suppress_print=False, use_gpu=False, try_diag=True, track_loss=False,
                 get_dual_variables=False, phase1_tol=0, phase1_t0=0.01, x0=None, update_slacks_every=0, shard_rows=False, sparse="auto"):
        self.A, self.c, self.C, self.b, self.d = A, c, C, b, d
        if c is not None and c.ndim != 1:
            raise ValueError("c must be 1-dimensional!")
        check_pair(A, b, "A", "b")
        check_pair(C, d, "C", "d")
        if c is not None:
            self.n = len(c)
        elif A is not None:
            self.n = A.shape[1]
        elif C is not None:
            self.n = C.shape[1]
        else:
            raise ValueError("At least one of c, A, C must be given")
        for M, nm in ((A, "A"), (C, "C")):
            if M is not None and M.shape[1] != self.n:
                raise ValueError(f"c must have the same number of entries as {nm} has columns!")
        self.lb, self.ub = as_bound(lower_bound, "Lower"), as_bound(upper_bound, "Upper")
        check_bounds(self.lb, self.ub, self.n)
        self.equality_constrained = A is not None
        self.bounded = self.lb is not None or self.ub is not None
        self.x = default_x0(self.n, self.lb, self.ub) if x0 is None else np.asarray(x0, dtype=np.float64)
        self._init_common(t0, mu, max_outer_iters, max_inner_iters, phase1_max_inner_iters, epsilon, inner_epsilon,
                          max_cg_iters, alpha, beta, suppress_print, track_loss, linear_solve_method,
                          get_dual_variables, phase1_t0, phase1_tol, update_slacks_every, use_gpu)
        self.try_diag = try_diag
        self._check_method(linear_solve_method, self.equality_constrained)
        if check_cvxpy:
            self._cvxpy_precheck()
        self.num_constraints = (0 if d is None else len(d)) + (self.n if self.lb is not None else 0) + (
            self.n if self.ub is not None else 0)
        self._eq_tol = 1e-4 * self.n  # LPSolver.py:600-602
        # host -> device (LPSolver.py:160-176): the only bulk transfer of a solve
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
        # optional extension: `sparse` ("auto" | True | False) -- inequality rows that are almost all zeros (the MIPLIB
        # .npy problems of testSolver.py:278-300) are kept as CSR and the Hessian is formed entry-wise (engine.SparseRows)
        self.data = LinearProblemData(self.n, self.device, c=c, C=C_loc, d=d_loc, lb=lb_loc, ub=ub_loc, A=A, b=b,
                                      sparse=False if self.sharded else sparse)
        self.x_dev = torch.as_tensor(self.x).to(device=self.device, dtype=F64).clone()
        if C is not None:
            self.phase1_solver = PhaseOneSolver(
                C=C, d=d, lower_bound=self.lb, upper_bound=self.ub, x0=self.x, max_outer_iters=max_outer_iters,
                max_inner_iters=phase1_max_inner_iters, epsilon=epsilon, inner_epsilon=inner_epsilon, alpha=alpha,
                beta=beta, mu=mu, suppress_print=suppress_print, n=self.n, tol=phase1_tol, t0=phase1_t0,
                update_slacks_every=update_slacks_every, _data=self.data, _launcher=self.launcher,
                _newton_cls=newton_cls)
        diagonal = C is None and try_diag  # LPSolver.py:436-446
        if diagonal and not self.bounded:
            raise ValueError("LP without inequality constraints or bounds has no barrier Hessian")
        self.ns = newton_cls(self.data, phase1=False, max_iters=max_inner_iters, epsilon=inner_epsilon, alpha=alpha,
                               beta=beta, update_slacks_every=update_slacks_every, diagonal=diagonal,
                               launcher=self.launcher, max_cg_iters=max_cg_iters,
                               linear_solver="cg" if linear_solve_method == "cg" else "cholesky")

    def _cvxpy_precheck(self):
        """Optional CVXPY/Clarabel feasibility pre-check (LPSolver.py:471-505); skipped when cvxpy is absent."""
        try:
            import cvxpy as cvx
        except Exception:
            return
        x = cvx.Variable(self.n)
        obj = cvx.Minimize(self.c.T @ x) if self.c is not None else cvx.Minimize(cvx.sum(x))
        cons = []
        if self.A is not None:
            cons.append(self.A @ x == self.b)
        if self.C is not None:
            cons.append(self.C @ x <= self.d)
        if self.lb is not None:
            cons.append(x >= self.lb)
        if self.ub is not None:
            cons.append(self.ub >= x)
        prob = cvx.Problem(obj, cons)
        try:
            prob.solve(solver="CLARABEL")
        except Exception as e:  # pragma: no cover
            print(e)
        self.feasible, self.cvxpy_val, self.cvxpy_sol = prob.status, prob.value, x.value
        if self.feasible == "infeasible":
            raise ValueError("Provided problem instance is infeasible!")
        elif self.feasible == "unbounded":
            raise ValueError("Provided problem instance is unbounded!")

    def _objective_value(self, x):
        return float(self.ns.dot(self.data.c, x))

    def _equality_residual(self, x):
        return self.ns.equality_residual(x)

    def _dual_variables(self, best_x, t):
        """LPSolver.py:641-646."""
        if self.C is not None or self.bounded:
            self.lam_star = HostArray((1.0 / (t * self.ns.slacks_at(best_x))).cpu().numpy())
        if self.A is not None:
            self.v_star = HostArray((self.ns.ws.v / t).cpu().numpy())
