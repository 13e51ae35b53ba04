"""``SOCPSolver`` -- drop-in for the reference's ``SOCPSolver.py`` (constructor :20-54, ``solve`` :616-753) on
the B200 engine:

    minimise 1/2 x'Px + q'x   s.t.  ||A_i x + b_i||_2 <= c_i'x + d_i  (i < M),   F x = g,   lb <= x <= ub

``A, b, c, d`` are lists (one entry per cone); a 1-D ``A_i`` (or a diagonal 2-D one) is a diagonal matrix
(SOCPSolver.py:282-292)."""

import numpy as np
import torch
import torch.distributed as tdist

try:
    from ._solver_base import BarrierSolverBase, as_bound, check_bounds, default_x0, HostArray
    from .cone_engine import ConeNewton, ConeProblemData
    from .dist import row_range
    from .engine import F64, Launcher
    from .PhaseOneSolver import PhaseOneSolver
    from .sharded_engine import ShardedConeNewton
except ImportError:  # flat-module use
    from _solver_base import BarrierSolverBase, as_bound, check_bounds, default_x0, HostArray
    from cone_engine import ConeNewton, ConeProblemData
    from dist import row_range
    from engine import F64, Launcher
    from PhaseOneSolver import PhaseOneSolver
    from sharded_engine import ShardedConeNewton


def _as_list(v):
    if v is None:
        return None
    return list(v) if isinstance(v, (list, tuple)) else [v]


class SOCPSolver(BarrierSolverBase):
    def __init__(self, P=None, q=None, A=None, b=None, c=None, d=None, F=None, g=None, lower_bound=0,
                 upper_bound=None, t0=0.1, phase1_t0=0.01, max_outer_iters=20, max_inner_iters=50,
                 phase1_max_inner_iters=500, epsilon=1e-10, inner_epsilon=1e-5, check_cvxpy=True,
                 linear_solve_method="cholesky", max_cg_iters=50, alpha=0.2, beta=0.6, mu=15, suppress_print=False,
                 use_gpu=False, try_diag=True, track_loss=False, get_dual_variables=False, phase1_tol=0,
                 use_psd_condition=False, x0=None, update_slacks_every=0, shard_rows=False):
        self.P, self.q, self.F, self.g = P, q, F, g
        if A is None:
            raise ValueError("No cone contraints detected. Run with LPSolver or QPSolver for better performance.")
        A, b, c, d = _as_list(A), _as_list(b), _as_list(c), _as_list(d)
        # -- validation (SOCPSolver.py:255-385) --
        if P is not None and (P.ndim != 2 or P.shape[0] != P.shape[1]):
            raise ValueError("P must be a symmetric, square PSD matrix!")
        if q is not None and q.ndim != 1:
            raise ValueError("q must be q-dimensional!")
        if P is not None and q is not None and P.shape[1] != len(q):
            raise ValueError("P and q must have the same dimension")
        for i, Ai in enumerate(A):
            if Ai.ndim > 2:
                raise ValueError("A must be 1- or 2-dimensional!")
            if Ai.ndim == 2 and Ai.shape[0] == Ai.shape[1]:
                off = Ai - np.diag(np.diag(Ai))
                if not off.any():
                    A[i] = np.diag(Ai).copy()  # diagonal compression
        if q is not None:
            self.n = len(q)
        elif P is not None:
            self.n = P.shape[1]
        else:
            self.n = A[0].shape[-1]
        for Ai in A:
            if Ai.shape[-1] != self.n:
                raise ValueError("q must have the same number of entries as A has columns!")
        if b is not None:
            if len(b) == 1:
                b = b * len(A)
            if len(A) != len(b):
                raise ValueError("Must provide an equal number of A and b")
            for Ai, bi in zip(A, b):
                if bi.ndim != 1:
                    raise ValueError("b must be 1-dimensional!")
                if len(bi) != (Ai.shape[0]):
                    raise ValueError("A and b must have agreeing dimensions!")
        if F is not None and F.ndim != 2:
            raise ValueError("F must be 2-dimensional!")
        if (F is None) ^ (g is None):
            raise ValueError("Both F and g must be defined, or neither!")
        if F is not None and (g.ndim != 1 or len(g) != F.shape[0]):
            raise ValueError("F and g must have agreeing dimensions!")
        if F is not None and F.shape[1] != self.n:
            raise ValueError("A and F must have the same number of columns!")
        if c is not None:
            for ci in c:
                if ci.ndim != 1:
                    raise ValueError("c must be 1-dimensional!")
                if len(ci) != self.n:
                    raise ValueError("c must have the same number of entries as A has columns!")
            if len(A) != len(c):
                raise ValueError("Must provide equal number of c and A")
        if d is not None:
            for di in d:
                if not np.isscalar(di):
                    raise ValueError("d must be a scalar!")
            if len(d) == 1:
                d = d * len(A)
            if len(d) != len(A):
                raise ValueError("Must provide equal number of A and d")
        self.A, self.b, self.c, self.d = A, b, c, d
        self.lb, self.ub = as_bound(lower_bound, "Lower"), as_bound(upper_bound, "Upper")
        check_bounds(self.lb, self.ub, self.n)
        self.equality_constrained = F is not None
        self.inequality_constrained = True
        self.bounded = self.lb is not None or self.ub is not None
        self.x = default_x0(self.n, self.lb, self.ub) if x0 is None else np.asarray(x0, dtype=np.float64)
        self._init_common(t0, mu, max_outer_iters, max_inner_iters, phase1_max_inner_iters, epsilon, inner_epsilon,
                          max_cg_iters, alpha, beta, suppress_print, track_loss, linear_solve_method,
                          get_dual_variables, phase1_t0, phase1_tol, update_slacks_every, use_gpu)
        self.use_psd_condition = use_psd_condition
        self._check_method(linear_solve_method, self.equality_constrained)
        self.num_constraints = len(A) + (self.n if self.lb is not None else 0) + (self.n if self.ub is not None else 0)
        self._eq_tol = 1e-3  # SOCPSolver.py:699-703
        self.launcher = Launcher(self.device)
        # optional extension (not in the reference): shard whole cones of ONE problem over the ranks of an initialised
        # torch.distributed group (partial Hessian + NCCL all-reduce, see sharded_engine.py; BASELINE configs[3])
        newton_cls, lb_loc, ub_loc = ConeNewton, self.lb, self.ub
        self.sharded = bool(shard_rows) and tdist.is_available() and tdist.is_initialized() and \
            tdist.get_world_size() > 1
        if self.sharded:
            x_all = torch.as_tensor(self.x).to(device=self.device, dtype=F64).clone()
            tdist.broadcast(x_all, src=0)  # default_x0 may be random (SOCPSolver.py:166): one x0 for all ranks
            self.x = x_all.cpu().numpy()
            lo, hi = row_range(len(A), tdist.get_rank(), tdist.get_world_size())
            if hi <= lo:
                raise ValueError("shard_rows needs at least one cone per rank")
            A, b, c, d = A[lo:hi], (None if b is None else b[lo:hi]), (None if c is None else c[lo:hi]), (
                None if d is None else d[lo:hi])
            if tdist.get_rank() != 0:
                lb_loc = ub_loc = None  # bound rows belong to rank 0
            newton_cls = ShardedConeNewton
        self.data = ConeProblemData(self.n, self.device, P, q, A, b, c, d, lb=lb_loc, ub=ub_loc, F=F, g=g)
        self.x_dev = torch.as_tensor(self.x).to(device=self.device, dtype=F64).clone()
        self.phase1_solver = PhaseOneSolver(
            socp=True, socp_params=(A, b, c, d), _newton_cls=newton_cls, lower_bound=self.lb, upper_bound=self.ub, x0=self.x,
            max_outer_iters=max_outer_iters, max_inner_iters=phase1_max_inner_iters, epsilon=epsilon,
            inner_epsilon=inner_epsilon, alpha=alpha, beta=beta, mu=mu, suppress_print=suppress_print, n=self.n,
            tol=phase1_tol, use_psd_condition=use_psd_condition, t0=phase1_t0,
            update_slacks_every=update_slacks_every, _data=self.data, _launcher=self.launcher)
        self.ns = newton_cls(self.data, phase1=False, max_iters=max_inner_iters, epsilon=inner_epsilon, alpha=alpha,
                             beta=beta, use_psd_condition=use_psd_condition, update_slacks_every=update_slacks_every,
                             launcher=self.launcher, max_cg_iters=max_cg_iters,
                             linear_solver="cg" if linear_solve_method == "cg" else "cholesky")

    def _objective_value(self, x):
        if self.P is not None:
            return self.ns.qp_objective(x)
        return float(self.ns.dot(self.data.c, x))

    def _equality_residual(self, x):
        return self.ns.equality_residual(x)

    def _dual_variables(self, best_x, t):
        """lam = 1/(t * slack) per barrier term (the reference's own SOCP branch dereferences a non-existent
        attribute, SOCPSolver.py:742)."""
        self.lam_star = HostArray((1.0 / (t * self.ns.slacks_at(best_x))).cpu().numpy())
        if self.F is not None:
            self.v_star = HostArray((self.ns.ws.v / t).cpu().numpy())
