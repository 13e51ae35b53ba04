"""Device engine for second-order-cone barriers (SOCP main phase and its phase-I).

Mirrors ``FunctionManagerSOCP`` / ``FunctionManagerSOCPPhase1`` (FunctionManager.py:834-1460) in the stacked
form of SURVEY.md K2: cone rows are one row-major matrix, the Hessian is ONE weighted SYRK over
``W = [A rows ; c rows ; g rows]`` (no per-cone n x n caches), the line search uses the exact quadratic of each
cone slack along the ray.  Everything outside the cone-specific pieces (Cholesky, solves, Armijo / residual
searches, infeasible-start block elimination for ``F x = g``) is inherited from ``LinearNewton``.
"""

import numpy as np
import torch

try:
    from . import _abi
    from .engine import F64, LinearNewton, _round_up, hess_i8_slices, to_dev_matrix, to_dev_vector
except ImportError:  # flat-module use
    import _abi
    from engine import F64, LinearNewton, _round_up, hess_i8_slices, to_dev_matrix, to_dev_vector


class ConeProblemData:
    """SOCP data in HBM.  ``W`` holds [stacked A_i rows | c_i rows | g_i rows (rewritten every iteration)]."""

    def __init__(self, n, device, P, q, A, b, c, d, lb=None, ub=None, F=None, g=None):
        self.n, self.device = n, device
        A = [np.diag(Ai) if Ai.ndim == 1 else Ai for Ai in A]  # compressed diagonals (SOCPSolver.py:282-292)
        self.M = len(A)
        ks = [Ai.shape[0] for Ai in A]
        self.cone_off_host = np.concatenate([[0], np.cumsum(ks)]).astype(np.int32)
        self.ktot = int(self.cone_off_host[-1])
        self.cone_off = torch.as_tensor(self.cone_off_host).to(device)
        self.rows_w = self.ktot + 2 * self.M
        self.ldw = _round_up(n, 16)
        self.W = torch.zeros((self.rows_w, self.ldw), dtype=F64, device=device)
        for i, Ai in enumerate(A):
            self.W[self.cone_off_host[i]:self.cone_off_host[i + 1], :n].copy_(torch.as_tensor(Ai))
        if c is not None:
            for i, ci in enumerate(c):
                self.W[self.ktot + i, :n].copy_(torch.as_tensor(ci))
        self.has_c = c is not None
        bst = np.zeros(self.ktot) if b is None else np.concatenate([np.asarray(bi, dtype=np.float64) for bi in b])
        self.bstack = to_dev_vector(bst, device)
        dv = np.zeros(self.M) if d is None else np.asarray([float(di) for di in d], dtype=np.float64)
        self.dvec = to_dev_vector(dv, device)
        self.lb = to_dev_vector(lb, device, n)
        self.ub = to_dev_vector(ub, device, n)
        self.nbounds = (n if ub is not None else 0) + (n if lb is not None else 0)
        # fields LinearNewton expects
        self.m, self.C, self.ldc, self.d = 0, None, 0, None
        self.is_qp = P is not None
        self.P, self.ldp = (None, 0) if P is None else to_dev_matrix(P, device)
        self.q = to_dev_vector(q, device)
        if not self.is_qp:
            self.c = to_dev_vector(np.zeros(n) if q is None else q, device)  # linear objective q'x
        self.p = 0 if F is None else F.shape[0]
        self.A, self.lda = (None, 0) if F is None else to_dev_matrix(F, device)
        self.At, self.ldat = (None, 0) if F is None else to_dev_matrix(torch.as_tensor(F).T, device)
        self.b = to_dev_vector(g, device)
        self.n_slacks = self.M + self.nbounds   # barrier ("constraint") entries, FunctionManager.py:921
        self.n_tail = self.M                    # right-hand sides, feasibility test only
        self.h2d_bytes = sum(t.numel() * 8 for t in (self.W, self.bstack, self.dvec, self.lb, self.ub, self.P, self.q,
                                                     self.A, self.At, self.b) if t is not None)


class ConeNewton(LinearNewton):
    def __init__(self, data, **kw):
        super().__init__(data, **kw)
        d, dev = data, data.device
        z = lambda *s: torch.zeros(*s, dtype=F64, device=dev)  # noqa: E731
        ws = self.ws
        ws.lhs, ws.dlhs, ws.coefA = z(d.ktot), z(d.ktot), z(d.ktot)
        ws.rhs, ws.drhs, ws.coefC, ws.plog, ws.pinv = z(d.M), z(d.M), z(d.M), z(d.M), z(d.M)
        ws.wts = z(d.rows_w)
        ws.red_b = z(8)
        ws.p2 = z(d.n_slacks + d.n_tail)
        ws.Vc = z(2, d.M)
        ws.Vc[0].fill_(1.0)
        self.guard = 1e-15 if self.phase1 else 1e-12  # Q5
        self.direct_trial = True
        self.tail_off = d.n_slacks
        nws = _abi.lib().ipm_gemv_t_ws_doubles(d.M, d.n, 2)
        if nws > ws.gt_ws_n:
            ws.gt_ws, ws.gt_ws_n = z(nws), nws

    # -- cone-specific pieces ----------------------------------------------------------------------
    def _A(self):
        return self.d.W.data_ptr()

    def _Cc(self):
        return self.d.W.data_ptr() + 8 * self.d.ktot * self.d.ldw

    def _G(self):
        return self.d.W.data_ptr() + 8 * (self.d.ktot + self.d.M) * self.d.ldw

    def _bound_inv_ptrs(self, inv):
        d = self.d
        off = d.M
        ub_p = lb_p = None
        if d.ub is not None:
            ub_p = inv.data_ptr() + 8 * off
            off += d.n
        if d.lb is not None:
            lb_p = inv.data_ptr() + 8 * off
        return ub_p, lb_p

    def _p2_ptr(self):
        return self.ws.p2.data_ptr()

    def _verify_trial(self, z):
        """Evaluate the barrier at x + table[kmax]*dx itself: cone slacks rhs^2 - |lhs|^2 cancel catastrophically
        near the boundary, so the frozen log-sum of the Armijo test and the feasibility verdict must come from the
        same formula the next iteration uses (as in the reference, NewtonSolver.py:172-183)."""
        ws, L = self.ws, self.L
        L("ipm_table_lookup_f64", self.table.data_ptr(), self.table_len, ws.kmax.data_ptr(), ws.ls_out.data_ptr() + 48)
        L("ipm_trial_point_f64", self.nz, ws.ls_out.data_ptr() + 48, z.data_ptr(), ws.dz.data_ptr(),
          ws.trial.data_ptr())
        self._eval(ws.trial, ws.tri)
        return ws.red_t.data_ptr(), ws.red_t.data_ptr() + 32

    def _eval(self, z, slot):
        """FunctionManager.py:933-994 (+ 1258-1262 in phase-I)."""
        d, ws, L = self.d, self.ws, self.L
        n, M = d.n, d.M
        ws.lhs.copy_(d.bstack)
        L("ipm_gemv_n_f64", self._A(), d.ldw, d.ktot, n, z.data_ptr(), ws.lhs.data_ptr(), 1.0, 1.0)
        ws.rhs.copy_(d.dvec)
        if d.has_c:
            L("ipm_gemv_n_f64", self._Cc(), d.ldw, M, n, z.data_ptr(), ws.rhs.data_ptr(), 1.0, 1.0)
        s_ptr = z.data_ptr() + 8 * n if self.phase1 else None
        red_b = None
        if d.nbounds:
            L("ipm_lin_barrier_eval_f64", 0, n, None, None, z.data_ptr(), _abi.ptr(d.ub), _abi.ptr(d.lb), s_ptr,
              int(self.phase1), self.guard, slot.slacks.data_ptr() + 8 * M, slot.inv.data_ptr() + 8 * M, None,
              slot.hdiag.data_ptr(), ws.red_b.data_ptr(), ws.ev_ws.data_ptr())
            red_b = ws.red_b.data_ptr()
        L("ipm_cone_eval_f64", M, d.cone_off.data_ptr(), d.ktot, ws.lhs.data_ptr(), ws.rhs.data_ptr(), s_ptr, self.guard,
          self.tail_off, slot.slacks.data_ptr(), slot.inv.data_ptr(), ws.wts.data_ptr(), ws.coefA.data_ptr(),
          ws.coefC.data_ptr(), ws.plog.data_ptr(), ws.pinv.data_ptr(), red_b, slot.red.data_ptr())

    def _gradient(self, t, lin, slot, g, want_border):
        """FunctionManager.py:1055-1102 (main), 1323-1374 (phase-I).  The per-cone rows
        g_i = 2/(s_i+eps) (A_i' lhs_i - c_i rhs_i) land in W (they are also SYRK operand rows)."""
        d, ws, L = self.d, self.ws, self.L
        n, M = d.n, d.M
        L("ipm_cone_grad_rows_f64", M, n, d.cone_off.data_ptr(), self._A(), d.ldw, self._Cc(), d.ldw,
          ws.coefA.data_ptr(), ws.coefC.data_ptr(), self._G(), d.ldw)
        nv = 2 if (self.phase1 and want_border) else 1
        if nv == 2:
            ws.Vc[1].copy_(slot.inv[:M])
        L("ipm_gemv_t_f64", self._G(), d.ldw, M, n, ws.Vc.data_ptr(), nv, M, ws.CtV.data_ptr(), n, 1.0, 0.0,
          ws.gt_ws.data_ptr(), ws.gt_ws_n)
        ub_p, lb_p = self._bound_inv_ptrs(slot.inv)
        L("ipm_lin_grad_f64", n, t, _abi.ptr(lin), ws.CtV.data_ptr(), ub_p, lb_p, int(self.phase1),
          slot.red.data_ptr() + 16, (ws.CtV.data_ptr() + 8 * n) if nv == 2 else None, g.data_ptr(), ws.hxs.data_ptr())

    def _hessian(self, t):
        """H = t P + sum_i [ 2/(s_i+eps) (A_i'A_i + c_i c_i') + g_i g_i' ] + bound diagonal (+ phase-I border)
        (FunctionManager.py:1104-1158, 1376-1455) as one weighted SYRK over W."""
        d, ws, L = self.d, self.ws, self.L
        n = d.n
        beta = 0.0
        if d.is_qp and not self.phase1:
            L("ipm_scale_copy_upper_f64", ws.H.data_ptr(), ws.ldh, d.P.data_ptr(), d.ldp, n, t)
            beta = 1.0
        L.tag = "hessian"
        slices = hess_i8_slices(d.rows_w, n)  # all SYRK weights are positive (2 / slack, 1): csrc/cone.cu
        if slices:
            L("ipm_hess_i8_f64", d.W.data_ptr(), d.ldw, d.rows_w, n, ws.wts.data_ptr(), beta, ws.H.data_ptr(), ws.ldh, slices,
              self._hess_i8_ws(slices, d.rows_w).data_ptr())
        else:
            L("ipm_gemm_tn_f64", d.W.data_ptr(), d.ldw, d.W.data_ptr(), d.ldw, ws.wts.data_ptr(), 1.0, beta,
              ws.H.data_ptr(), ws.ldh, n, n, d.rows_w, 1)
        L.tag = None
        shift = self.shift + (1e-9 if self.use_psd_condition else 0.0)
        L("ipm_hess_finish_f64", ws.H.data_ptr(), ws.ldh, n, ws.hdiag.data_ptr() if d.nbounds else None,
          ws.hxs.data_ptr() if self.phase1 else None, (ws.red.data_ptr() + 24) if self.phase1 else None, shift)

    def _feasibility(self, z):
        """Step-size back-off over [cone slacks | bounds | cone right-hand sides] (NewtonSolver.py:170-183 with
        the slack vector of FunctionManager.py:962-988)."""
        d, ws, L = self.d, self.ws, self.L
        n, M = d.n, d.M
        L("ipm_gemv_n_f64", self._A(), d.ldw, d.ktot, n, ws.dz.data_ptr(), ws.dlhs.data_ptr(), 1.0, 0.0)
        if d.has_c:
            L("ipm_gemv_n_f64", self._Cc(), d.ldw, M, n, ws.dz.data_ptr(), ws.drhs.data_ptr(), 1.0, 0.0)
        if d.nbounds:
            L("ipm_ls_feas_lin_f64", 0, n, ws.slacks.data_ptr() + 8 * M, None, ws.dz.data_ptr(), int(d.ub is not None),
              int(d.lb is not None), int(self.phase1), self.table.data_ptr(), self.table_len,
              ws.p1.data_ptr() + 8 * M, ws.kmax.data_ptr())
        ds_ptr = ws.dz.data_ptr() + 8 * n if self.phase1 else None
        L("ipm_cone_ls_coeffs_f64", M, d.cone_off.data_ptr(), ws.lhs.data_ptr(), ws.rhs.data_ptr(), ws.dlhs.data_ptr(),
          ws.drhs.data_ptr(), ds_ptr, self.tail_off, ws.p1.data_ptr(), ws.p2.data_ptr())
        L("ipm_ls_feas_poly_f64", M, ws.slacks.data_ptr(), ws.p1.data_ptr(), ws.p2.data_ptr(), self.table.data_ptr(),
          self.table_len, ws.kmax.data_ptr(), 0 if d.nbounds else 1)
        t8 = 8 * self.tail_off
        L("ipm_ls_feas_poly_f64", M, ws.slacks.data_ptr() + t8, ws.p1.data_ptr() + t8, ws.p2.data_ptr() + t8,
          self.table.data_ptr(), self.table_len, ws.kmax.data_ptr(), 0)
