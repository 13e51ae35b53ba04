"""Integrated phase-I driver (drop-in for the reference's ``PhaseOneSolver.py:6-154``): minimise s subject to
``Cx - d <= s`` and relaxed bounds (or the second-order-cone analogue) on the device, stopping as soon as
``s < -tol``."""

import numpy as np
import torch

try:
    from . import _abi
    from .engine import F64, LinearNewton, LinearProblemData
except ImportError:  # flat-module use (directory on sys.path, like the reference)
    import _abi
    from engine import F64, LinearNewton, LinearProblemData


class _PhaseOneState:
    """What callers read from ``phase1_solver.phase1_fm`` (LPSolver.py:546): the current slack variable."""

    def __init__(self):
        self.s = 0.0
        self.t = None


class PhaseOneSolver:
    def __init__(self, C=None, d=None, lower_bound=0, upper_bound=None, x0=None, max_outer_iters=50,
                 max_inner_iters=20, epsilon=1e-8, inner_epsilon=1e-5, linear_solve_method="cholesky",
                 max_cg_iters=50, alpha=0.2, beta=0.6, mu=15, t0=1, suppress_print=False, use_gpu=False,
                 track_loss=False, n=None, tol=0.1, socp=False, socp_params=None, use_psd_condition=False,
                 update_slacks_every=0, _data=None, _launcher=None, _newton_cls=None):
        _abi.require_device()
        self.C, self.d, self.lb, self.ub = C, d, lower_bound, upper_bound
        self.n = n if n is not None else len(x0)
        self.max_outer_iters, self.max_inner_iters = max_outer_iters, max_inner_iters
        self.epsilon, self.inner_epsilon = epsilon, inner_epsilon
        self.alpha, self.beta, self.mu, self.t0, self.tol = alpha, beta, mu, t0, tol
        self.suppress_print = suppress_print
        self.use_gpu = True
        self.phase1_fm = _PhaseOneState()
        self.outer_iters, self.inner_iters = 0, []
        device = torch.device("cuda", torch.cuda.current_device())
        if socp:
            try:
                from .cone_engine import ConeNewton, ConeProblemData
            except ImportError:
                from cone_engine import ConeNewton, ConeProblemData
            data = _data if _data is not None else ConeProblemData(self.n, device, None, None, *socp_params,
                                                                   lb=lower_bound, ub=upper_bound)
            self.ns = (_newton_cls or ConeNewton)(data, phase1=True, max_iters=max_inner_iters, epsilon=inner_epsilon, alpha=alpha,
                                 beta=beta, phase1_tol=tol, use_psd_condition=use_psd_condition,
                                 update_slacks_every=update_slacks_every, launcher=_launcher)
        else:
            data = _data if _data is not None else LinearProblemData(self.n, device, C=C, d=d, lb=lower_bound,
                                                                     ub=upper_bound)
            self.ns = (_newton_cls or LinearNewton)(data, phase1=True, max_iters=max_inner_iters, epsilon=inner_epsilon, alpha=alpha,
                                   beta=beta, phase1_tol=tol, use_psd_condition=use_psd_condition,
                                   update_slacks_every=update_slacks_every, launcher=_launcher)
        self.data = data
        # z = (x0, s0),  s0 = 1 - min slack at s = 0  (FunctionManager.py:390-393, PhaseOneSolver.py:86-89)
        self.z = torch.zeros(self.n + 1, dtype=F64, device=device)
        self.z[: self.n].copy_(torch.as_tensor(np.asarray(x0, dtype=np.float64)))
        self.phase1_fm.s = float(-self.ns.min_slack(self.z) + 1)
        self.z[self.n] = self.phase1_fm.s
        self.x = self.z

    def solve(self, x0=None):
        """Returns (x view of z on the device, final s).  ``x0`` is accepted and ignored, as in the reference
        (it updates the function manager but not the iterate, PhaseOneSolver.py:114-127)."""
        t = self.t0
        self.outer_iters, self.inner_iters = 0, []
        ns = self.ns
        ns.set_t(t)
        obj_val = self.phase1_fm.s
        for _ in range(self.max_outer_iters):
            if not self.suppress_print:
                print(f"Current slack: {self.phase1_fm.s}")
            k, _, _ = ns.solve(self.z)
            self.outer_iters += 1
            self.inner_iters.append(k)
            obj_val = float(self.z[self.n])
            self.phase1_fm.s = obj_val
            if obj_val < -self.tol:
                break
            t = min(t * self.mu, (self.n + 1.0) / self.epsilon)  # PhaseOneSolver.py:151
            ns.set_t(t)
        return self.z[: self.n], obj_val
