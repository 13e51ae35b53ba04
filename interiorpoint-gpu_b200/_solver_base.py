"""Shared front-end logic of the drop-in ``LPSolver`` / ``QPSolver`` / ``SOCPSolver`` classes: input
validation, default starting point, phase-I hand-off and the outer barrier loop -- the reference's
``LPSolver.py:105-224,514-653`` (QP/SOCP: same structure) with the Newton centering delegated to the
device engine.  Only scalars cross the PCIe bus inside ``solve()``.
"""

import numpy as np
import torch

try:
    from . import _abi
except ImportError:  # flat-module use
    import _abi


class HostArray(np.ndarray):
    """NumPy array that also answers ``.get()`` (CuPy idiom used by callers of the reference's GPU arm,
    e.g. testSolver.py:1171) so result attributes work for code written against either arm."""

    def __new__(cls, a):
        return np.asarray(a).view(cls)

    def get(self):
        return np.asarray(self)


def as_bound(v, name):
    if v is None:
        return None
    try:
        return np.array(v, dtype=np.float64)
    except Exception:
        raise ValueError(f"{name} bound must be a scalar or list!")


def default_x0(n, lb, ub):
    """LPSolver.py:131-143."""
    if lb is not None and ub is not None:
        return (np.maximum(lb, -1e2) + np.minimum(ub, 1e2)) / 2 * np.ones(n)
    if lb is not None:
        return (np.maximum(lb, -1e2) + 1e-1) * np.ones(n)
    if ub is not None:
        return (np.minimum(ub, 1e2) - 1e-1) * np.ones(n)
    return np.random.rand(n)


def check_bounds(lb, ub, n):
    """LPSolver.py:271-314."""
    if lb is not None and lb.ndim > 0 and len(lb) != n:
        raise ValueError("Lower bound must be a scalar or have the same number of dimensions as other parameters!")
    if ub is not None and ub.ndim > 0 and len(ub) != n:
        raise ValueError("Upper bound must be a scalar or have the same number of dimensions as other parameters!")
    if lb is not None and ub is not None and np.any(ub - lb < 0):
        raise ValueError("Lower bound must be lower than upper bound")


def check_pair(M, v, Mname, vname):
    """LPSolver.py:237-269."""
    if (M is not None) ^ (v is not None):
        raise ValueError(f"Both {Mname} and {vname} must be defined, or neither!")
    if M is not None:
        if M.ndim != 2:
            raise ValueError(f"{Mname} must be 2-dimensional!")
        if v.ndim != 1:
            raise ValueError(f"{vname} must be 1-dimensional!")
        if len(v) != M.shape[0]:
            raise ValueError(f"{Mname} and {vname} must have agreeing dimensions!")


class BarrierSolverBase:
    """Outer barrier loop.  Subclasses set: ``self.n``, ``self.x`` (host x0), ``self.device``, ``self.ns``
    (main Newton engine), ``self.phase1_solver`` (or None), ``self.num_constraints``, ``self._eq_tol``."""

    def _init_common(self, t0, mu, max_outer_iters, max_inner_iters, phase1_max_inner_iters, epsilon, inner_epsilon,
                     max_cg_iters, alpha, beta, suppress_print, track_loss, linear_solve_method, get_dual_variables,
                     phase1_t0, phase1_tol, update_slacks_every, use_gpu):
        _abi.require_device()  # no CPU fallback, whatever `use_gpu` says
        self.use_gpu = True
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.alpha, self.beta = alpha, beta
        self.t0, self.mu = t0, mu
        self.outer_iters, self.inner_iters = 0, []
        self.max_outer_iters, self.max_inner_iters = max_outer_iters, max_inner_iters
        self.epsilon, self.inner_epsilon = epsilon, inner_epsilon
        self.max_cg_iters = max_cg_iters
        self.optimal = False
        self.value = self.optimality_gap = self.xstar = self.lam_star = self.vstar = self.v_star = None
        self.suppress_print = suppress_print
        self.track_loss = track_loss
        self.linear_solve_method = linear_solve_method
        self.get_dual_variables = get_dual_variables
        self.phase1_t0, self.phase1_tol = phase1_t0, phase1_tol
        self.phase1_max_inner_iters = phase1_max_inner_iters
        self.update_slacks_every = update_slacks_every
        self.feasible, self.cvxpy_val, self.cvxpy_sol = None, None, None
        self.phase1_solver = None
        self.timings = {}

    @staticmethod
    def _check_method(method, equality_constrained):
        """Newton-class dispatch of the reference (LPSolver.py:371-448).  Every direct method solves the same SPD
        system; the device engine implements them with its Cholesky kernels.  ``cg`` for equality-constrained
        problems raises NotImplementedError in the reference too (NewtonSolverInfeasibleStart.py:604); without equality
        constraints it selects the device CG (engine.LinearNewton._cg_direction, NewtonSolver.py:365-400)."""
        if method not in ("cholesky", "np_solve", "np_lstsq", "direct", "cg", "kkt"):
            raise ValueError("Please enter a valid linear solve method!")
        if method == "kkt" and not equality_constrained:
            raise ValueError("No KKT System non-equality-constrained problems! Please choose another solver")
        if method == "cg" and equality_constrained:
            # NewtonSolverCGInfeasibleStart / ...Diagonal...: NewtonSolverInfeasibleStart.py:604,874
            raise NotImplementedError("conjugate-gradient Newton solves are not implemented for equality-constrained "
                                      "problems (neither in the reference)")

    # -- helpers on the device -------------------------------------------------------------------------
    def _objective_value(self, x):
        raise NotImplementedError

    def _equality_residual(self, x):
        raise NotImplementedError

    def _check_x0(self, x):
        """LPSolver.py:655-682."""
        lb, ub = self.lb, self.ub
        if lb is not None and (x <= lb).any():
            raise ValueError("Initial x must be in domain of problem (all entries greater than lower bound)")
        elif ub is not None and (x >= ub).any():
            raise ValueError("Initial x must be in domain of problem (all entries less than upper bound)")
        if len(x) != self.n:
            raise ValueError("Initial x must have the problem dimension!")

    def __repr__(self):
        opt_val = "Not yet solved" if self.optimal is False else self.value
        return f"{type(self).__name__}(Optimal Value: {opt_val})"

    # -- the outer loop ----------------------------------------------------------------------------------
    def solve(self, resolve=True, **kwargs):
        """Same contract as the reference's ``solve`` (LPSolver.py:514-653): returns the optimal value; sets
        ``value, xstar, optimality_gap, outer_iters, inner_iters, objective_vals`` (+ ``lam_star, v_star``)."""
        if not resolve and self.optimal:
            return self.value
        t = kwargs.get("t0", self.t0)
        max_outer_iters = kwargs.get("max_outer_iters", self.max_outer_iters)
        self.track_loss = kwargs.get("track_loss", self.track_loss)
        if "x0" in kwargs:
            x_host = np.asarray(kwargs["x0"], dtype=np.float64)
            self._check_x0(x_host)
            self.x_dev.copy_(torch.as_tensor(x_host))
        x = self.x_dev
        if self.phase1_solver is not None and self.phase1_solver.phase1_fm.s >= 1:
            if not self.suppress_print:
                print("running phase 1 solver")
            x, s = self.phase1_solver.solve()
            if s > -self.phase1_tol:
                raise ValueError("Phase 1 Solver did not successfully find a feasible point!")
            if not self.suppress_print:
                print(f"found a feasible point with slack {s}")
        if not self.suppress_print:
            print("proceeding to solve method")
        self.outer_iters = 0
        objective_vals = []
        self.inner_iters = []
        ns = self.ns
        ns.set_t(t)
        ns.shift = ns.base_shift  # a regularisation found necessary in an earlier solve() does not carry over
        if ns.equality:
            ns.reset_dual()
        dual_gap = self.num_constraints
        best_x = x.clone()
        best_obj = np.inf
        for it in range(max_outer_iters):
            numiters_t, _, success_flag = ns.solve(x)
            self.outer_iters += 1
            self.inner_iters.append(numiters_t)
            if not ns.equality or self._equality_residual(x) < self._eq_tol:
                obj_val = self._objective_value(x)
                if not self.suppress_print:
                    print(f"Objective value is now {obj_val}")
                if self.track_loss:
                    objective_vals.append(obj_val)
                if obj_val < best_obj:
                    best_obj = obj_val
                    best_x = x.clone()
                elif success_flag:
                    break
            else:
                if not self.suppress_print:
                    print(f"Newton step at iteration {it+1} did not converge")
                if len(objective_vals) > 0:
                    objective_vals.append(objective_vals[-1])
            if not self.suppress_print and numiters_t >= self.max_inner_iters:
                print(f"Reached max Newton steps during {it+1}th centering step (t={t})")
            dual_gap = self.num_constraints / t
            if dual_gap < self.epsilon:
                break
            t = t * self.mu
            ns.set_t(t)
        self.xstar_device = best_x
        self.t_final = t  # the t the dual variables are formed with (LPSolver.py:641-646)
        self.xstar = HostArray(best_x.cpu().numpy())
        if self.get_dual_variables:
            self._dual_variables(best_x, t)
        self.optimal = True
        self.value = best_obj
        self.optimality_gap = dual_gap
        self.objective_vals = objective_vals
        return self.value

    def _dual_variables(self, best_x, t):
        pass

    def plot(self, subtract_cvxpy=True):
        """Convergence plot (LPSolver.py:684-705); matplotlib is imported lazily."""
        if not (self.optimal and self.track_loss):
            raise ValueError("Need to solve problem with track_loss set to True to be able to plot convergence!")
        import matplotlib.pyplot as plt

        obj_vals = np.asarray(self.objective_vals, dtype=float)
        ref = self.cvxpy_val if (subtract_cvxpy and self.cvxpy_val is not None) else obj_vals.min()
        ax = plt.subplot()
        ax.step(np.cumsum(self.inner_iters[-len(obj_vals):]), obj_vals - ref, where="post")
        ax.set_xlabel("Cumulative Newton iterations")
        ax.set_ylabel("Optimality gap")
        ax.set_title(f"Convergence of {type(self).__name__}")
        ax.set_yscale("log")
        return ax
