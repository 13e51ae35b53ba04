"""Stand-alone phase-I -- drop-in for the reference's ``PhaseOne.PhaseOneSolver`` (PhaseOne.py:18-395) on the
B200 engine: find a strictly interior point of ``G x <= h`` by minimising s subject to ``G x - h <= s``.

Differences in mechanism, not in contract: the Newton system ``(H + 0.01 I) d = -g`` (PhaseOne.py:120-134) is
solved with the engine's Cholesky (the matrix is SPD) instead of ``numpy.linalg.solve``; ``linear_solver="cg"`` runs
conjugate gradients on the device with the reference's start vector ``[x, s]`` and iteration cap (PhaseOne.py:137-150).  The line search is the reference's plain Armijo rule
(PhaseOne.py:187-218: slope ``g.d``, alpha 0.2, beta 0.7, barrier re-evaluated at every trial)."""

import numpy as np
import torch

try:
    from . import _abi
    from ._solver_base import HostArray
    from .engine import F64, LinearNewton, LinearProblemData
except ImportError:  # flat-module use
    import _abi
    from _solver_base import HostArray
    from engine import F64, LinearNewton, LinearProblemData


class PhaseOneSolver:
    def __init__(self, G, h, mu, x0=None, eps=1e-8, max_iter_interior=200, max_iter_newton=200, use_cupy=False,
                 linear_solver="solve", max_cg_iters=50):
        if linear_solver not in ["solve", "cg"]:
            raise RuntimeError("Invalid linear solver")
        _abi.require_device()
        self.G, self.h = G, h
        self.mu, self.eps = mu, eps
        self.max_iter_interior, self.max_iter_newton = max_iter_interior, max_iter_newton
        self.solver, self.max_iter_cg = linear_solver, max_cg_iters
        self.use_cupy = use_cupy
        self.warn = False
        m, n = G.shape
        self.m, self.n = m, n
        device = torch.device("cuda", torch.cuda.current_device())
        self.data = LinearProblemData(n, device, C=G, d=np.asarray(h, dtype=np.float64).ravel())
        self.ns = LinearNewton(self.data, phase1=True, max_iters=max_iter_newton, epsilon=eps, alpha=0.2, beta=0.7,
                               linear_solver="cg" if linear_solver == "cg" else "cholesky", max_cg_iters=max_cg_iters)
        self.ns.base_shift = self.ns.shift = 0.01  # "some conditioning", PhaseOne.py:123-127
        self.z = torch.zeros(n + 1, dtype=F64, device=device)
        self.z[:n].copy_(torch.as_tensor(np.ones(n) if x0 is None else np.asarray(x0, dtype=np.float64)))
        # s = max(Gx - h) + 1 (PhaseOne.py:96): min slack of the s = 0 problem is -max(Gx - h)
        self.max_violation = -self.ns.min_slack(self.z)
        self.z[n] = self.max_violation + 1
        self.newton_steps = 0

    @property
    def x(self):
        return HostArray(self.z[: self.n].cpu().numpy())

    @property
    def s(self):
        return float(self.z[self.n])

    # ---- the reference's public pieces (exercised by its known-answer tests, AutomatedTestsPhaseOne.py:15-232) ----
    def phase_one_objective(self, x, s, t):
        """t s - sum log(s + h - G x) at the GIVEN point (PhaseOne.py:171-185)."""
        ns, ws = self.ns, self.ns.ws
        z = torch.empty(self.n + 1, dtype=F64, device=self.z.device)
        z[: self.n].copy_(torch.as_tensor(np.asarray(x, dtype=np.float64)))
        z[self.n] = float(s)
        ns._eval(z, ws.tri)
        return float(t) * float(s) - float(ws.tri.red[0])

    def phase_one_check_feasibility(self, x, s):
        """max(G x - h) < s (PhaseOne.py:220-238)."""
        z = torch.zeros(self.n + 1, dtype=F64, device=self.z.device)
        z[: self.n].copy_(torch.as_tensor(np.asarray(x, dtype=np.float64)))
        return bool(-self.ns.min_slack(z) < float(s))

    def phase_one_gradient(self, t):
        """[sum_i g_i / f_i ; t - sum_i 1 / f_i] at the current (x, s) (PhaseOne.py:240-272)."""
        ns, ws = self.ns, self.ns.ws
        ns.set_t(t)
        ns._eval(self.z, ws.cur)
        ns._gradient(float(t), None, ws.cur, ws.g, want_border=True)
        return HostArray(ws.g[: self.n + 1].cpu().numpy())

    def phase_one_hessian(self):
        """Full symmetric (n+1) x (n+1) Hessian WITHOUT the 0.01 I conditioning (PhaseOne.py:274-328)."""
        ns, ws = self.ns, self.ns.ws
        ns._eval(self.z, ws.cur)
        ns._gradient(ns.t, None, ws.cur, ws.g, want_border=True)
        shift, ns.shift = ns.shift, 0.0
        try:
            ns._hessian(ns.t)
        finally:
            ns.shift = shift
        U = torch.triu(ws.H[: ns.nz, : ns.nz])
        return HostArray((U + torch.triu(U, 1).T).cpu().numpy())

    def phase_one_newtons_method(self, t):
        """PhaseOne.py:109-162.  Returns True if the Newton iteration cap was reached."""
        ns, ws, L, z = self.ns, self.ns.ws, self.ns.L, self.z
        ns.set_t(t)
        it = 0
        for it in range(self.max_iter_newton):
            ns._eval(z, ws.cur)
            ns._gradient(t, None, ws.cur, ws.g, want_border=True)
            ns._hessian(t)
            if self.solver == "cg":
                # spcg(hess, -grad, x0=[x, s], maxiter=max_cg_iters)  (PhaseOne.py:143-150)
                L("ipm_symmetrize_upper_f64", ws.H.data_ptr(), ws.ldh, ns.nz)
                L("ipm_gemv_n_f64", ws.H.data_ptr(), ws.ldh, ns.nz, ns.nz, z.data_ptr(), ws.cg_hx0.data_ptr(), 1.0, 0.0)
                L("ipm_lincomb3_f64", ns.nz, -1.0, ws.g.data_ptr(), 0.0, None, 0.0, None, ws.gtrial.data_ptr())
                ws.dz.copy_(z)
                L("ipm_cg_solve_f64", ws.H.data_ptr(), ws.ldh, ns.nz, ws.gtrial.data_ptr(), ws.dz.data_ptr(),
                  ws.cg_hx0.data_ptr(), 1.0, self.max_iter_cg, 1e-5, ws.cg_ws.data_ptr())
            else:
                ns._factor()
                L("ipm_lincomb3_f64", ns.nz, -1.0, ws.g.data_ptr(), 0.0, None, 0.0, None, ws.dz.data_ptr())
                ns._chol_solve_vec(ws.dz)
            pairs = ns._objective_pairs(z, None) + [(ws.g, z, ns.nz), (ws.g, ws.dz, ns.nz)]
            ns._dots(pairs)
            lam_sq = -ns._read_terms(5)[4]
            if lam_sq / 2 <= self.eps:  # PhaseOne.py:152-155
                break
            ns._feasibility(z)
            L("ipm_ls_armijo_f64", self.data.n_slacks, ws.slacks.data_ptr(), ws.p1.data_ptr(), None,
              ns.table.data_ptr(), ns.table_len, ws.kmax.data_ptr(), ws.red.data_ptr(), ws.terms.data_ptr(), float(t),
              0.2, 0, None, None, 1, 0, ws.ls_out.data_ptr())
            L("ipm_axpy_dev_f64", ns.nz, ws.ls_out.data_ptr(), ws.dz.data_ptr(), z.data_ptr())
            self.newton_steps += 1
            if self.s < 0:  # PhaseOne.py:160-161
                break
        return it == self.max_iter_newton - 1

    def execute_phase_one(self):
        """PhaseOne.py:330-375."""
        if self.max_violation <= 0:  # already feasible: s = -1 (PhaseOne.py:341-344)
            self.z[self.n] = -1.0
            return
        t = 1
        for _ in range(self.max_iter_interior):
            if self.phase_one_newtons_method(t):
                print("Warning, Newtons method ran its maximum number of steps")
                self.warn = True
            if self.m / t <= self.eps:
                break
            if self.s < 0:
                break
            t *= self.mu

    def solve(self):
        """Returns ``(x, s, warn)``: s < 0 strictly feasible, s > 0 the set is empty (PhaseOne.py:377-395)."""
        self.execute_phase_one()
        return self.x, self.s, self.warn
