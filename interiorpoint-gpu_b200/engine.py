"""Device-resident Newton / barrier engine for the linear-inequality family (LP, QP and their phase-I).

Host side of the hot path: Python + torch tensors for memory, every numerical op is a call into
``libipm_b200.so`` (hand-written sm_100a CUDA, see ``csrc/``).  One Newton iteration is a fixed sequence of
asynchronous launches on one stream; the host reads back a handful of scalars ONCE per iteration (step size,
stuck flag, g.dx, potrf info) to take the reference's control-flow decisions (NewtonSolver.py:129-133).

Mirrors, on the device, ``FunctionManagerLP/QP/Phase1`` (FunctionManager.py:197-831) and
``NewtonSolverCholesky`` / ``NewtonSolverCholeskyInfeasibleStart`` (+ diagonal variants)
(NewtonSolver.py:250-341,403-420; NewtonSolverInfeasibleStart.py:356-538,757-809).
"""

import ctypes as C
import os
from types import SimpleNamespace

import numpy as np
import torch

try:
    from . import _abi
except ImportError:  # flat-module use (directory on sys.path, like the reference)
    import _abi

STUCK = 1e-13
F64 = torch.float64


def _round_up(v, m):
    return (v + m - 1) // m * m


def step_table(beta):
    """The reference's step sequence 1, beta, beta*beta, ... (repeated multiplication, NewtonSolver.py:175,195)
    up to and including the first entry below 1e-13."""
    if not (0.0 < beta < 1.0):
        raise ValueError("beta must lie in (0, 1)")
    tab = [1.0]
    a = 1
    while True:
        a *= beta
        tab.append(float(a))
        if a < STUCK:
            break
    return tab


# Hessian on the INT8 tensor pipe (csrc/hess_i8.cu) instead of the FP64 DMMA contraction.  Measured on a B200
# (profiles/hess_i8_sizes_r02.jsonl): 12.8 ms against 32.5 ms at n = 8192, m = 16384 with 8 digits per entry, identical
# accuracy class (2e-15 of sum |x||x|); 1.59x at n = 2048, m = 4096; 0.84x at n = 1024 (three kernels and a 148-CTA
# persistent grid only pay off on large operands), so the default is size-gated.  IPM_HESSIAN_I8 = 0: never; 1: whenever the shape
# is supported; 5..8: that many 7-bit digits, whenever supported.
HESS_I8_MIN_N, HESS_I8_MIN_M, HESS_I8_SLICES = 2048, 1024, 8


def hess_i8_slices(m, n):
    """Digits per entry for an m x n dense constraint matrix, or 0 for the FP64 DMMA kernel."""
    env = os.environ.get("IPM_HESSIAN_I8", "")
    if env == "0" or m <= 0:
        return 0
    slices = int(env) if env in ("5", "6", "7", "8") else HESS_I8_SLICES
    if env == "" and (n < HESS_I8_MIN_N or m < HESS_I8_MIN_M):
        return 0
    return slices if _abi.lib().ipm_hess_i8_ws_bytes(m, n, slices) > 0 else 0


class Launcher:
    """Counts launches of our own kernels (bench.py reports it) and owns the stream."""

    def __init__(self, device):
        self.device = device
        self.calls = 0
        self._base = _abi.lib().ipm_launch_count()
        self.timed_ops = None  # {"op name": [(start_event, end_event, tag), ...]} when bench.py profiles live
        self.tag = None
        self.stream = None     # raw cudaStream_t handed to every call (None = legacy default stream = torch's default)

    def kernel_launches(self):
        """Kernels launched by libipm_b200 since this launcher was created (process-wide counter)."""
        return int(_abi.lib().ipm_launch_count() - self._base)

    def timed_range(self, tag):
        """Context manager: CUDA events around a group of launches/collectives (only while bench.py profiles)."""
        launcher = self

        class _R:
            def __enter__(self_r):
                self_r.rec = launcher.timed_ops.get("range:" + tag) if launcher.timed_ops is not None else None
                if self_r.rec is not None:
                    self_r.e0 = torch.cuda.Event(enable_timing=True)
                    self_r.e0.record()
                return self_r

            def __exit__(self_r, *exc):
                if self_r.rec is not None:
                    e1 = torch.cuda.Event(enable_timing=True)
                    e1.record()
                    self_r.rec.append((self_r.e0, e1, tag))
                return False

        return _R()

    def __call__(self, name, *args):
        self.calls += 1
        rec = self.timed_ops.get(name) if self.timed_ops is not None else None
        if rec is None:
            _abi.call(name, *args, self.stream)
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _abi.call(name, *args, self.stream)
        e1.record()
        rec.append((e0, e1, self.tag))


def to_dev_matrix(a, device):
    """Host (or device) 2-D array -> device FP64 row-major with ld rounded up to 16 doubles (TMA needs a 16-byte
    row stride; the padding columns are zero)."""
    t = torch.as_tensor(a)
    r, c = t.shape
    ld = _round_up(max(c, 1), 16)
    out = torch.zeros((r, ld), dtype=F64, device=device)
    out[:, :c].copy_(t, non_blocking=True)
    return out, ld


def to_dev_vector(a, device, n=None):
    if a is None:
        return None
    t = torch.as_tensor(np.asarray(a, dtype=np.float64))
    if t.ndim == 0:
        t = t.repeat(n)
    return t.to(device=device, dtype=F64).contiguous()


SPARSE_DENSITY = 0.02   # C is handled as sparse rows below this fill (auto mode)
SPARSE_MIN_N = 256      # ... and only when the problem is large enough for the dense SYRK to matter


def _looks_sparse(C):
    """Fill below SPARSE_DENSITY?  A strided sample of <= 128 rows decides first, so that a dense 1 GB matrix is not
    scanned on the host inside the constructor (bench.py's end-to-end arm times it)."""
    sample = C[:: max(1, C.shape[0] // 128)]
    if np.count_nonzero(sample) >= SPARSE_DENSITY * sample.size:
        return False
    return np.count_nonzero(C) < SPARSE_DENSITY * C.size


class SparseRows:
    """CSR(C), CSR(C^T) and the segment map of the Hessian pattern, resident in HBM.

    C^T diag(w) C only has entries (i, j) where some row holds both columns.  For every such upper-triangle entry
    the host lists, once, the contributing rows r (ascending) with their products c_ri * c_rj; the device kernel
    ``ipm_sparse_syrk_f64`` then forms each entry with one thread in a fixed order.  C itself is never materialised
    as a dense device matrix."""

    def __init__(self, C, n, device):
        import scipy.sparse as sp

        S = sp.csr_matrix(np.asarray(C, dtype=np.float64))
        S.sum_duplicates()
        S.sort_indices()
        St = S.T.tocsr()
        St.sort_indices()
        i32 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32)).to(device)  # noqa: E731
        f64 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(device)  # noqa: E731
        self.m, self.n, self.nnz = S.shape[0], n, int(S.nnz)
        self.rowptr, self.col, self.val = i32(S.indptr), i32(S.indices), f64(S.data)
        self.t_rowptr, self.t_col, self.t_val = i32(St.indptr), i32(St.indices), f64(St.data)
        # all pairs (i <= j) of the non-zero columns of every row
        counts = np.diff(S.indptr)
        rows_i, rows_j, rows_r, prods = [], [], [], []
        for k in np.unique(counts):
            if k == 0:
                continue
            rr = np.nonzero(counts == k)[0]
            starts = S.indptr[rr]
            idx = starts[:, None] + np.arange(k)[None, :]
            cols, vals = S.indices[idx], S.data[idx]          # [len(rr), k], columns ascending
            a, b = np.triu_indices(k)
            rows_i.append(cols[:, a].ravel()), rows_j.append(cols[:, b].ravel())
            rows_r.append(np.repeat(rr, len(a))), prods.append((vals[:, a] * vals[:, b]).ravel())
        if rows_i:
            I, J = np.concatenate(rows_i).astype(np.int64), np.concatenate(rows_j).astype(np.int64)
            R, Pv = np.concatenate(rows_r), np.concatenate(prods)
            key = I * n + J
            order = np.lexsort((R, key))                      # by output entry, then ascending row
            key, R, Pv, I, J = key[order], R[order], Pv[order], I[order], J[order]
            first = np.concatenate([[True], key[1:] != key[:-1]])
            segptr = np.concatenate([np.nonzero(first)[0], [len(key)]])
            out_i, out_j = I[first], J[first]
        else:
            segptr, R, Pv, out_i, out_j = np.zeros(1), np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0)
        self.nout = len(out_i)
        self.segptr, self.seg_row, self.seg_prod = i32(segptr), i32(R), f64(Pv)
        self.out_i, self.out_j = i32(out_i), i32(out_j)
        self.h2d_bytes = 4 * (len(S.indptr) + len(St.indptr) + 2 * S.nnz + len(segptr) + len(R)) + 8 * (
            2 * S.nnz + len(Pv) + self.nout)


class LinearProblemData:
    """Problem data resident in HBM.  C: m x n inequality rows, A: p x n equality rows (also kept transposed,
    n x p, because every contraction kernel wants the contracted index as the row index), P: n x n."""

    def __init__(self, n, device, c=None, P=None, q=None, C=None, d=None, lb=None, ub=None, A=None, b=None,
                 sparse="auto"):
        self.n, self.device = n, device
        self.m = 0 if C is None else C.shape[0]
        self.p = 0 if A is None else A.shape[0]
        self.sparse = None
        if C is not None and sparse:
            if sparse is True or (n >= SPARSE_MIN_N and isinstance(C, np.ndarray) and _looks_sparse(C)):
                self.sparse = SparseRows(C, n, device)
        self.C, self.ldc = (None, 0) if (C is None or self.sparse is not None) else to_dev_matrix(C, device)
        self.d = to_dev_vector(d, device)
        self.lb = to_dev_vector(lb, device, n)
        self.ub = to_dev_vector(ub, device, n)
        self.is_qp = P is not None
        self.P, self.ldp = (None, 0) if P is None else to_dev_matrix(P, device)
        self.q = to_dev_vector(q, device)
        if not self.is_qp:
            self.c = to_dev_vector(np.ones(n) if c is None else c, device)
        self.A, self.lda = (None, 0) if A is None else to_dev_matrix(A, device)
        self.At, self.ldat = (None, 0) if A is None else to_dev_matrix(torch.as_tensor(A).T, device)
        self.b = to_dev_vector(b, device)
        self.n_slacks = self.m + (n if ub is not None else 0) + (n if lb is not None else 0)
        self.h2d_bytes = sum(t.numel() * 8 for t in (self.C, self.d, self.lb, self.ub, self.P, self.q, self.A,
                                                     self.At, self.b) if t is not None)
        if self.sparse is not None:
            self.h2d_bytes += self.sparse.h2d_bytes


class NewtonWorkspace:
    """All per-iteration buffers, allocated once (nz = n for the main phase, n + 1 for phase-I)."""

    def __init__(self, data, nz):
        dev, n, m, p = data.device, data.n, data.m, data.p
        z = lambda *s: torch.zeros(*s, dtype=F64, device=dev)  # noqa: E731
        self.nz = nz
        self.ldh = _round_up(nz, 16)
        self.H = z(nz, self.ldh)
        self.g, self.dz, self.trial, self.gtrial = z(nz), z(nz), z(nz), z(nz)
        ms = max(data.n_slacks + getattr(data, "n_tail", 0), 1)
        self.slacks, self.inv, self.p1 = z(ms), z(ms), z(ms)
        self.slacks_t, self.inv_t = z(ms), z(ms)
        self.w, self.Cx, self.Cdx = z(max(m, 1)), z(max(m, 1)), z(max(m, 1))
        self.w_t = z(max(m, 1))
        self.hdiag, self.hxs, self.lin, self.hq, self.Pdx = z(n), z(n), z(n), z(n), z(n)
        self.hdiag_t = z(n)
        self.CtV = z(2, n)
        self.CtV_in = z(2, max(m, 1))
        self.red = z(8)
        self.tr_ws = z(max(nz, p, 1))
        self.red_t = z(8)
        self.terms = z(8)  # [obj0, dobj, quad, g.z, g.dz, ...]
        self.ls_out = z(8)
        self.consts = torch.tensor([1.0, 0.0], dtype=F64, device=dev)
        self.kmax = torch.zeros(1, dtype=torch.int32, device=dev)
        self.info = torch.zeros(2, dtype=torch.int32, device=dev)
        self.ev_ws = z(_abi.lib().ipm_lin_barrier_ws_doubles())
        nws = max(_abi.lib().ipm_gemv_t_ws_doubles(max(m, 1), n, 2),
                  _abi.lib().ipm_gemv_t_ws_doubles(max(p, 1), n, 2),
                  _abi.lib().ipm_gemv_t_ws_doubles(n, max(p, 1), 2))
        self.gt_ws, self.gt_ws_n = z(nws), nws
        if p:
            self.ldy = _round_up(p + 1, 16)
            self.Y = z(n, self.ldy)      # U^{-T} [A^T | g]
            self.lds = _round_up(p, 16)
            self.S = z(p, self.lds)      # Schur complement
            self.v, self.dv, self.wv = z(p), z(p), z(p)
            self.Axb, self.Adx, self.rhs_p = z(p), z(p), z(p)
            self.ATv, self.ATdv, self.yv = z(n), z(n), z(n)
            self.r0d, self.u0, self.u1 = z(n), z(n), z(n)
            self.hinv = z(n)
        self.cur = SimpleNamespace(slacks=self.slacks, inv=self.inv, w=self.w, hdiag=self.hdiag, red=self.red)
        self.tri = SimpleNamespace(slacks=self.slacks_t, inv=self.inv_t, w=self.w_t, hdiag=self.hdiag_t, red=self.red_t)
        # pinned host mirror for the once-per-iteration readback
        self.host = torch.zeros(24, dtype=F64).pin_memory()
        self.host_info = torch.zeros(2, dtype=torch.int32).pin_memory()


class LinearNewton:
    """Damped Newton centering for the LP / QP barrier (``phase1=False``) or its phase-I problem
    (``phase1=True``; iterate z = (x, s)).  Equality constraints (``data.A``) select the infeasible-start
    method with block elimination."""

    def __init__(self, data, phase1=False, max_iters=50, epsilon=1e-5, alpha=0.2, beta=0.6, phase1_tol=0.1,
                 use_psd_condition=False, update_slacks_every=0, diagonal=False, launcher=None,
                 linear_solver="cholesky", max_cg_iters=50):
        self.d = data
        self.phase1 = phase1
        self.nz = data.n + (1 if phase1 else 0)
        self.max_iters, self.eps, self.alpha, self.beta = max_iters, epsilon, alpha, beta
        self.phase1_tol = phase1_tol
        self.use_psd_condition = use_psd_condition
        self.update_slacks_every = update_slacks_every
        self.diagonal = diagonal
        self.equality = (data.A is not None) and not phase1
        # "cg": the reference's NewtonSolverCG (NewtonSolver.py:365-400) -- at most max_cg_iters conjugate-gradient steps
        # on (-H) dx = g instead of a factorisation (feasible start only; the reference's equality-constrained CG classes
        # raise NotImplementedError, NewtonSolverInfeasibleStart.py:604,874)
        self.linear_solver, self.max_cg_iters = linear_solver, max_cg_iters
        if linear_solver == "cg" and self.equality:
            raise NotImplementedError("conjugate gradients are not implemented for equality-constrained problems")
        self.L = launcher or Launcher(data.device)
        self.ws = NewtonWorkspace(data, self.nz)
        tab = step_table(beta)
        self.table = torch.tensor(tab, dtype=F64, device=data.device)
        self.table_len = len(tab)
        self.t = 1.0
        self.use_backup = False  # sticky, like the reference (NewtonSolver.py:319)
        self.direct_trial = False  # cones: re-evaluate the barrier AT the proposed trial point
        # diagonal shift of the Newton systems: base_shift = deliberate conditioning (0.01 in the stand-alone phase-I,
        # PhaseOne.py:123-127); shift grows from it only after a failed factorisation (_retry_regularised), is capped
        # relative to diag(H) and returns to base_shift at the start of every front-end solve()
        self.base_shift = 0.0
        self.shift = 0.0
        self._zsave = torch.zeros(self.nz, dtype=F64, device=data.device)
        self._vsave = torch.zeros(max(data.p, 1), dtype=F64, device=data.device)
        self.newton_steps = 0
        self.trace = None
        if linear_solver == "cg":
            zf = lambda k: torch.zeros(k, dtype=F64, device=data.device)  # noqa: E731
            self.ws.cg_hx, self.ws.cg_hx0 = zf(self.nz), zf(self.nz)
            self.ws.cg_ws = zf(_abi.lib().ipm_cg_ws_doubles(self.nz))

    # ---------------------------------------------------------------- barrier pieces
    def set_t(self, t):
        self.t = float(t)

    def _eval(self, z, slot):
        """Slacks, reciprocals, SYRK weights, diagonal terms and scalar reductions at z into `slot`
        (ws.cur for the iterate, ws.tri for a line-search trial point)."""
        d, ws, L = self.d, self.ws, self.L
        slacks, inv, w, hdiag, red = slot.slacks, slot.inv, slot.w, slot.hdiag, slot.red
        n, m = d.n, d.m
        if m:
            self._C_times(z, ws.Cx)
        s_ptr = z.data_ptr() + 8 * n if self.phase1 else None
        L("ipm_lin_barrier_eval_f64", m, n, _abi.ptr(ws.Cx) if m else None, _abi.ptr(d.d), z.data_ptr(),
          _abi.ptr(d.ub), _abi.ptr(d.lb), s_ptr, int(self.phase1),
          1e-15 if (self.phase1 or self.diagonal) else 0.0, slacks.data_ptr(), inv.data_ptr(), w.data_ptr(),
          hdiag.data_ptr(), red.data_ptr(), ws.ev_ws.data_ptr())

    def _C_times(self, x, out):
        """out = C x (dense row GEMV, or CSR for sparse C)."""
        d, L = self.d, self.L
        if d.sparse is not None:
            sp = d.sparse
            L("ipm_csr_gemv_f64", sp.rowptr.data_ptr(), sp.col.data_ptr(), sp.val.data_ptr(), d.m, x.data_ptr(),
              out.data_ptr(), 1.0, 0.0)
        else:
            L("ipm_gemv_n_f64", d.C.data_ptr(), d.ldc, d.m, d.n, x.data_ptr(), out.data_ptr(), 1.0, 0.0)

    def _bound_inv_ptrs(self, inv):
        d = self.d
        off = d.m
        ub_p = lb_p = None
        if d.ub is not None:
            ub_p = inv.data_ptr() + 8 * off
            off += d.n
        if d.lb is not None:
            lb_p = inv.data_ptr() + 8 * off
        return ub_p, lb_p

    def _lin_term(self, z):
        """lin = c (LP) or P x + q (QP): the objective gradient without t."""
        d, ws, L = self.d, self.ws, self.L
        if self.phase1:
            return None
        if not d.is_qp:
            return d.c
        if d.q is not None:
            ws.lin.copy_(d.q)
            beta = 1.0
        else:
            beta = 0.0
        L("ipm_gemv_n_f64", d.P.data_ptr(), d.ldp, d.n, d.n, z.data_ptr(), ws.lin.data_ptr(), 1.0, beta)
        return ws.lin

    def _gradient(self, t, lin, slot, g, want_border):
        """g from reciprocal slacks (FunctionManager.py:232-265, 509-545, 741-781)."""
        d, ws, L = self.d, self.ws, self.L
        inv, w, red = slot.inv, slot.w, slot.red
        n, m = d.n, d.m
        nv = 2 if (self.phase1 and want_border) else 1
        if m:
            if nv == 2:  # V = [inv_ineq ; w] as two rows of one buffer
                ws.CtV_in[0].copy_(inv[:m])
                ws.CtV_in[1].copy_(w[:m])
                vptr, ldv = ws.CtV_in.data_ptr(), m
            else:
                vptr, ldv = inv.data_ptr(), m
            if d.sparse is not None:  # C^T v through CSR(C^T): fixed summation order, no atomics
                sp = d.sparse
                for v in range(nv):
                    L("ipm_csr_gemv_f64", sp.t_rowptr.data_ptr(), sp.t_col.data_ptr(), sp.t_val.data_ptr(), n,
                      vptr + 8 * v * ldv, ws.CtV.data_ptr() + 8 * v * n, 1.0, 0.0)
            else:
                L("ipm_gemv_t_f64", d.C.data_ptr(), d.ldc, m, n, vptr, nv, ldv, ws.CtV.data_ptr(), n, 1.0, 0.0,
                  ws.gt_ws.data_ptr(), ws.gt_ws_n)
        ub_p, lb_p = self._bound_inv_ptrs(inv)
        L("ipm_lin_grad_f64", n, t, _abi.ptr(lin), ws.CtV.data_ptr() if m else None, ub_p, lb_p, int(self.phase1),
          red.data_ptr() + 16, (ws.CtV.data_ptr() + 8 * n) if (m and nv == 2) else None, g.data_ptr(),
          ws.hxs.data_ptr())

    def _hessian(self, t):
        """H (upper) = [t P +] C' diag(w) C + diag(hdiag) [+ phase-I border] (FunctionManager.py:267-326,
        547-611, 783-827)."""
        d, ws, L = self.d, self.ws, self.L
        n, m = d.n, d.m
        beta = 0.0
        if d.is_qp and not self.phase1:
            L("ipm_scale_copy_upper_f64", ws.H.data_ptr(), ws.ldh, d.P.data_ptr(), d.ldp, n, t)
            beta = 1.0
        elif m == 0 or d.sparse is not None:
            L("ipm_scale_copy_upper_f64", ws.H.data_ptr(), ws.ldh, None, 0, n, 0.0)
        if m and d.sparse is not None:
            sp = d.sparse
            L.tag = "hessian"
            L("ipm_sparse_syrk_f64", sp.nout, sp.segptr.data_ptr(), sp.seg_row.data_ptr(), sp.seg_prod.data_ptr(),
              sp.out_i.data_ptr(), sp.out_j.data_ptr(), ws.w.data_ptr(), ws.H.data_ptr(), ws.ldh)
            L.tag = None
        elif m:
            L.tag = "hessian"
            slices = hess_i8_slices(m, n)
            if slices:
                L("ipm_hess_i8_f64", d.C.data_ptr(), d.ldc, m, n, ws.w.data_ptr(), beta, ws.H.data_ptr(), ws.ldh, slices,
                  self._hess_i8_ws(slices).data_ptr())
            else:
                L("ipm_gemm_tn_f64", d.C.data_ptr(), d.ldc, d.C.data_ptr(), d.ldc, ws.w.data_ptr(), 1.0, beta,
                  ws.H.data_ptr(), ws.ldh, n, n, m, 1)
            L.tag = None
        shift = self.shift + (1e-9 if self.use_psd_condition else 0.0)
        L("ipm_hess_finish_f64", ws.H.data_ptr(), ws.ldh, n, ws.hdiag.data_ptr(),
          ws.hxs.data_ptr() if self.phase1 else None, (ws.red.data_ptr() + 24) if self.phase1 else None, shift)

    def _hess_i8_ws(self, slices, rows=None):
        """Slice buffer of the INT8 Hessian kernel (as many bytes per entry of the operand as digits): one per problem,
        shared by the phase-I and the main-phase solver (same operand shape)."""
        d = self.d
        rows = d.m if rows is None else rows
        cached = getattr(d, "hess_i8_ws", None)
        if cached is None or cached[0] != (slices, rows):
            nbytes = _abi.lib().ipm_hess_i8_ws_bytes(rows, d.n, slices)
            buf = torch.empty(nbytes, dtype=torch.uint8, device=d.device)
            self.L("ipm_hess_i8_prepare", buf.data_ptr(), rows, d.n, slices)
            d.hess_i8_ws = cached = ((slices, rows), buf)
        return cached[1]

    def _schur_i8(self, slices):
        """(unit weights, slice buffer) of the INT8 kernel for the Schur complement Y'Y (Y: n x p)."""
        d = self.d
        cached = getattr(d, "schur_i8_ws", None)
        if cached is None or cached[0] != slices:
            buf = torch.empty(_abi.lib().ipm_hess_i8_ws_bytes(d.n, d.p, slices), dtype=torch.uint8, device=d.device)
            self.L("ipm_hess_i8_prepare", buf.data_ptr(), d.n, d.p, slices)
            d.schur_i8_ws = cached = (slices, (torch.ones(d.n, dtype=F64, device=d.device), buf))
        return cached[1]

    def _p2_ptr(self):
        """Quadratic line-search coefficients (second-order cones only)."""
        return None

    def _cg_direction(self, z):
        """dz = cg(-H, g, x0, maxiter=max_cg_iters) with x0 = -(x.g) x / (x.Hx) if x.g < 0 else 0 (NewtonSolver.py:376-398).
        All of it stays on the device: no read-back until the line search."""
        ws, L, nz = self.ws, self.L, self.nz
        L("ipm_symmetrize_upper_f64", ws.H.data_ptr(), ws.ldh, nz)
        L("ipm_gemv_n_f64", ws.H.data_ptr(), ws.ldh, nz, nz, z.data_ptr(), ws.cg_hx.data_ptr(), 1.0, 0.0)
        L("ipm_cg_descent_x0_f64", nz, z.data_ptr(), ws.g.data_ptr(), ws.cg_hx.data_ptr(), ws.dz.data_ptr(),
          ws.cg_hx0.data_ptr())
        L("ipm_cg_solve_f64", ws.H.data_ptr(), ws.ldh, nz, ws.g.data_ptr(), ws.dz.data_ptr(), ws.cg_hx0.data_ptr(), -1.0,
          self.max_cg_iters, 1e-5, ws.cg_ws.data_ptr())
        ws.info.zero_()

    def _factor(self):
        """ws.H (upper) <- U with H = U'U; ws.info[0] = LAPACK-style info (NewtonSolver.py:286,303)."""
        ws = self.ws
        self.L("ipm_potrf_upper_f64", ws.H.data_ptr(), ws.ldh, self.nz, ws.info.data_ptr())

    def _fuse_forward(self):
        """Forward solve of the Newton right-hand side inside the factorisation launch: from the size where the
        single-launch tile-DAG Cholesky is the default (csrc/chol.cu, dag::kAutoMinN).  IPM_FUSE_FORWARD=0 switches it off."""
        return self.nz >= 2048 and os.environ.get("IPM_FUSE_FORWARD", "1") != "0"

    def _rhs_column(self):
        ws = self.ws
        if getattr(ws, "rhs1", None) is None:
            ws.rhs1 = torch.zeros((self.nz, 16), dtype=F64, device=self.d.device)  # TMA: 16-byte row stride, even ld
        return ws.rhs1

    def _chol_solve_vec(self, vec):
        """vec <- H^{-1} vec using the factor in ws.H."""
        ws, L = self.ws, self.L
        L("ipm_trsv_upper_f64", ws.H.data_ptr(), ws.ldh, self.nz, vec.data_ptr(), 1, ws.tr_ws.data_ptr())
        L("ipm_trsv_upper_f64", ws.H.data_ptr(), ws.ldh, self.nz, vec.data_ptr(), 0, ws.tr_ws.data_ptr())

    def _dots(self, pairs, out_offset=0):
        k = len(pairs)
        pa = (C.c_void_p * k)(*[a.data_ptr() if hasattr(a, "data_ptr") else a for a, _, _ in pairs])
        pb = (C.c_void_p * k)(*[b.data_ptr() if hasattr(b, "data_ptr") else b for _, b, _ in pairs])
        nn = (C.c_int * k)(*[c for _, _, c in pairs])
        self.L("ipm_dots_f64", k, pa, pb, nn, self.ws.terms.data_ptr() + 8 * out_offset)

    # ---------------------------------------------------------------- scalar queries used by the outer loops
    def _read_terms(self, k):
        ws = self.ws
        ws.host[:k].copy_(ws.terms[:k], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return [float(v) for v in ws.host[:k]]

    def dot(self, a, b):
        self._dots([(a, b, a.numel())])
        return self._read_terms(1)[0]

    def min_slack(self, z):
        """Smallest slack at z (used for the phase-I start s0 = 1 - min slack, FunctionManager.py:390-393)."""
        ws = self.ws
        self._eval(z, ws.cur)
        ws.host[:1].copy_(ws.red[1:2], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(ws.host[0])

    def slacks_at(self, x):
        ws = self.ws
        self._eval(x, ws.cur)
        return ws.slacks[: self.d.n_slacks].clone()

    def equality_residual(self, x):
        """||A x - b||_2 (LPSolver.py:600-602)."""
        d, ws = self.d, self.ws
        ws.Axb.copy_(d.b)
        self.L("ipm_gemv_n_f64", d.A.data_ptr(), d.lda, d.p, d.n, x.data_ptr(), ws.Axb.data_ptr(), 1.0, -1.0)
        return float(np.sqrt(self.dot(ws.Axb, ws.Axb)))

    def qp_objective(self, x):
        """1/2 x'Px + q'x (FunctionManager.py:683-704)."""
        d, ws = self.d, self.ws
        self.L("ipm_gemv_n_f64", d.P.data_ptr(), d.ldp, d.n, d.n, x.data_ptr(), ws.Pdx.data_ptr(), 1.0, 0.0)
        pairs = [(x, ws.Pdx, d.n)] + ([(d.q, x, d.n)] if d.q is not None else [])
        self._dots(pairs)
        vals = self._read_terms(len(pairs))
        return 0.5 * vals[0] + (vals[1] if d.q is not None else 0.0)

    # ---------------------------------------------------------------- feasible-start iteration
    def _objective_pairs(self, z, lin):
        """dot pairs producing [obj(x), d obj.dx, dx'P dx] -- the exact polynomial of obj along the ray."""
        d, ws = self.d, self.ws
        n = d.n
        one, zero = ws.consts.data_ptr(), ws.consts.data_ptr() + 8
        if self.phase1:  # obj = s (FunctionManager.py:472-482)
            return [(z.data_ptr() + 8 * n, one, 1), (ws.dz.data_ptr() + 8 * n, one, 1), (zero, zero, 1)]
        if not d.is_qp:
            return [(d.c, z, n), (d.c, ws.dz, n), (zero, zero, 1)]
        # obj = x.(0.5 P x + q);  hq = lin - 0.5 P x = 0.5 P x + q
        self.L("ipm_gemv_n_f64", d.P.data_ptr(), d.ldp, n, n, ws.dz.data_ptr(), ws.Pdx.data_ptr(), 1.0, 0.0)
        self.L("ipm_lincomb3_f64", n, 0.5, lin.data_ptr(), 0.5 if d.q is not None else 0.0, _abi.ptr(d.q), 0.0,
               None, ws.hq.data_ptr())
        return [(ws.hq, z, n), (lin, ws.dz, n), (ws.dz, ws.Pdx, n)]

    def _feasibility(self, z):
        d, ws, L = self.d, self.ws, self.L
        n, m = d.n, d.m
        if m:
            self._C_times(ws.dz, ws.Cdx)
        L("ipm_ls_feas_lin_f64", m, n, ws.slacks.data_ptr(), ws.Cdx.data_ptr() if m else None, ws.dz.data_ptr(),
          int(d.ub is not None), int(d.lb is not None), int(self.phase1), self.table.data_ptr(), self.table_len,
          ws.p1.data_ptr(), ws.kmax.data_ptr())

    def _iterate_feasible(self, z):
        d, ws, L = self.d, self.ws, self.L
        t = self.t
        self._eval(z, ws.cur)
        lin = self._lin_term(z)
        self._gradient(t, lin, ws.cur, ws.g, want_border=True)
        if self.diagonal:
            # bounds-only LP: H is diagonal, dx = -g / h (NewtonSolver.py:415-420)
            L("ipm_vec_op_f64", 1, d.n, ws.g.data_ptr(), ws.hdiag.data_ptr(), ws.dz.data_ptr(), -1.0)
        elif self.linear_solver == "cg":
            self._hessian(t)
            self._cg_direction(z)
        else:
            self._hessian(t)
            L("ipm_lincomb3_f64", self.nz, -1.0, ws.g.data_ptr(), 0.0, None, 0.0, None, ws.dz.data_ptr())
            if self._fuse_forward():
                # ONE launch: H = U'U and f = U^{-T}(-g) -- the right-hand side rides along as an extra block column of
                # the tile DAG (as [A' | g] does in the equality-constrained step); only the backward solve remains
                rhs = self._rhs_column()
                rhs[:, 0].copy_(ws.dz)
                L("ipm_potrf_trsm_upper_f64", ws.H.data_ptr(), ws.ldh, self.nz, rhs.data_ptr(), rhs.stride(0), 1,
                  ws.info.data_ptr())
                ws.dz.copy_(rhs[:, 0])
                L("ipm_trsv_upper_f64", ws.H.data_ptr(), ws.ldh, self.nz, ws.dz.data_ptr(), 0, ws.tr_ws.data_ptr())
            else:
                self._factor()
                self._chol_solve_vec(ws.dz)
        self._feasibility(z)
        pairs = self._objective_pairs(z, lin) + [(ws.g, z, self.nz), (ws.g, ws.dz, self.nz)]
        self._dots(pairs)
        resume = 0
        while True:
            L_direct, nneg = self._verify_trial(z)
            L("ipm_ls_armijo_f64", d.n_slacks, ws.slacks.data_ptr(), ws.p1.data_ptr(), self._p2_ptr(),
              self.table.data_ptr(), self.table_len, ws.kmax.data_ptr(), ws.red.data_ptr(), ws.terms.data_ptr(), t,
              self.alpha, self.update_slacks_every, L_direct, nneg, 0, resume, ws.ls_out.data_ptr())
            # one readback per Newton iteration (the axpy below is skipped by the kernel-side flag only on retry)
            ws.host[:5].copy_(ws.ls_out[:5], non_blocking=True)
            ws.host[5:10].copy_(ws.terms[:5], non_blocking=True)
            ws.host_info.copy_(ws.info, non_blocking=True)
            if L_direct is None:
                break
            torch.cuda.current_stream().synchronize()
            status = float(ws.host[1])
            if status == 3.0:
                ws.kmax.add_(1)  # rare: the polynomial proposal is infeasible when evaluated directly -> next step
                resume = 0
            elif status == 4.0:
                resume = 1       # update_slacks_every: the kernel left kmax at the lagging point; re-evaluate there
            else:
                break
        L("ipm_axpy_dev_f64", self.nz, ws.ls_out.data_ptr(), ws.dz.data_ptr(), z.data_ptr())
        ws.host[10:11].copy_(z[self.nz - 1:self.nz], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        _abi.check_device_fault()
        h = ws.host
        return dict(step=float(h[0]), stuck=h[1] != 0, g_dz=float(h[9]), z_last=float(h[10]),
                    info=int(ws.host_info[0]))

    def _verify_trial(self, z):
        """Hook for barriers whose slacks must be re-evaluated AT the proposed trial point (second-order cones).
        Returns device pointers (frozen log-sum, #negative slacks) or (None, None)."""
        return None, None

    def solve(self, z):
        """Centering at the current t.  ``z`` is updated in place on the device.  Returns
        (iters, decrement_or_residual, success) like NewtonSolver.solve / NewtonSolverInfeasibleStart.solve."""
        if self.equality:
            return self._solve_infeasible(z)
        nd = None
        it = 0
        for it in range(self.max_iters):
            zsave = None
            if not self.diagonal:
                zsave = self._zsave  # persistent buffer: only read if the factorisation fails
                zsave.copy_(z)
            r = self._iterate_feasible(z)
            if r["info"] != 0 and not self.diagonal:
                r = self._retry_regularised(z, zsave)
            self.newton_steps += 1
            nd = -r["g_dz"] / 2
            if self.trace is not None:
                self.trace.append((r["step"], nd))
            if self.phase1 and r["z_last"] < -self.phase1_tol:  # NewtonSolver.py:105-107
                return it + 1, None, True
            if r["step"] < STUCK:
                return it + 1, nd, False
            elif nd < self.eps:
                return it + 1, nd, True
        return it + 1, nd, False

    def _retry_regularised(self, z, zsave):
        """Cholesky hit a non-positive pivot.  The reference switches (sticky) to an SVD least-squares solve
        (NewtonSolver.py:314-341); the device engine re-factorises H + shift*I with a growing shift instead
        (documented deviation: this path is only reached on numerically singular Hessians)."""
        self.use_backup = True
        if self.shift == self.base_shift:
            # scale of the shift from a fresh (unfactored) Hessian: the failed potrf left NaNs in ws.H
            z.copy_(zsave)
            self._eval(z, self.ws.cur)
            self._gradient(self.t, self._lin_term(z), self.ws.cur, self.ws.g, want_border=True)
            self._hessian(self.t)
            diag = torch.diagonal(self.ws.H[:, : self.nz])
            base = float(diag[torch.isfinite(diag)].abs().mean())
            scale = base if base > 0 else 1.0
            self._shift_cap = max(1e-6 * scale, 1e4 * self.base_shift)  # beyond this it is no longer the Newton system
            shift = max(100.0 * self.base_shift, 1e-14 * scale)
        else:
            shift = 100.0 * self.shift  # the step that just failed already used self.shift
        while shift <= self._shift_cap:
            self.shift = shift
            z.copy_(zsave)
            r = self._iterate_feasible(z)
            if r["info"] == 0:
                return r
            shift *= 100.0
        raise np.linalg.LinAlgError("Hessian is not positive definite even after a regularisation of 1e-6 of its "
                                    "diagonal")

    # ---------------------------------------------------------------- infeasible-start (equality constrained)
    def _direction_infeasible(self, z, lin):
        """Block elimination (NewtonSolverInfeasibleStart.py:386-490):
        H = U'U;  [Y | f] = U^{-T} [A' | g];  S = Y'Y;  w = S^{-1}(Ax - b - Y'f);  dx = -H^{-1}(g + A'w)
        (the reference's  y = H^{-1} g;  w = S^{-1}(Ax - b - A y)  with A y = Y'f, the forward solves of A' and g folded
        into the factorisation launch)."""
        d, ws, L = self.d, self.ws, self.L
        n, p, t = d.n, d.p, self.t
        # b2 = A x - b
        ws.Axb.copy_(d.b)
        L("ipm_gemv_n_f64", d.A.data_ptr(), d.lda, p, n, z.data_ptr(), ws.Axb.data_ptr(), 1.0, -1.0)
        if self.diagonal:
            # S = A diag(1/h) A'  (NewtonSolverInfeasibleStart.py:774-809)
            L("ipm_vec_op_f64", 2, n, ws.hdiag.data_ptr(), None, ws.hinv.data_ptr(), 1.0)
            L("ipm_gemm_tn_f64", d.At.data_ptr(), d.ldat, d.At.data_ptr(), d.ldat, ws.hinv.data_ptr(), 1.0, 0.0,
              ws.S.data_ptr(), ws.lds, p, p, n, 1)
            L("ipm_vec_op_f64", 0, n, ws.hinv.data_ptr(), ws.g.data_ptr(), ws.yv.data_ptr(), 1.0)
        else:
            self._hessian(t)
            # ONE launch: H = U'U and [Y | f] = U^{-T} [A' | g] (the right-hand sides are extra block columns of the tile
            # DAG):   A H^{-1} A' = Y'Y,   A H^{-1} g = Y'f
            ws.Y[:, :p].copy_(d.At[:, :p])
            ws.Y[:, p].copy_(ws.g)
            L("ipm_potrf_trsm_upper_f64", ws.H.data_ptr(), ws.ldh, n, ws.Y.data_ptr(), ws.ldy, p + 1, ws.info.data_ptr())
            ws.yv.copy_(ws.Y[:, p])  # f = U^{-T} g
            slices = hess_i8_slices(n, p)  # S = Y'Y is the same contraction as the Hessian (unit weights)
            if slices:
                L("ipm_hess_i8_f64", ws.Y.data_ptr(), ws.ldy, n, p, self._schur_i8(slices)[0].data_ptr(), 0.0,
                  ws.S.data_ptr(), ws.lds, slices, self._schur_i8(slices)[1].data_ptr())
            else:
                L("ipm_gemm_tn_f64", ws.Y.data_ptr(), ws.ldy, ws.Y.data_ptr(), ws.ldy, None, 1.0, 0.0, ws.S.data_ptr(),
                  ws.lds, p, p, n, 1)
        if self.shift:
            L("ipm_hess_finish_f64", ws.S.data_ptr(), ws.lds, p, None, None, None, self.shift)
        L("ipm_potrf_upper_f64", ws.S.data_ptr(), ws.lds, p, ws.info.data_ptr() + 4)
        # rhs = b2 - A H^{-1} g ; w = S^{-1} rhs
        ws.rhs_p.copy_(ws.Axb)
        if self.diagonal:
            L("ipm_gemv_n_f64", d.A.data_ptr(), d.lda, p, n, ws.yv.data_ptr(), ws.rhs_p.data_ptr(), -1.0, 1.0)
        else:
            L("ipm_gemv_t_f64", ws.Y.data_ptr(), ws.ldy, n, p, ws.yv.data_ptr(), 1, n, ws.rhs_p.data_ptr(), p, -1.0, 1.0,
              ws.gt_ws.data_ptr(), ws.gt_ws_n)
        ws.wv.copy_(ws.rhs_p)
        L("ipm_trsv_upper_f64", ws.S.data_ptr(), ws.lds, p, ws.wv.data_ptr(), 1, ws.tr_ws.data_ptr())
        L("ipm_trsv_upper_f64", ws.S.data_ptr(), ws.lds, p, ws.wv.data_ptr(), 0, ws.tr_ws.data_ptr())
        # dx = -H^{-1}(g + A'w).  g and A'w nearly cancel close to the central path, so the sum is formed FIRST and then
        # solved (forward + backward), as the reference does; -U^{-1}(f + Y w) would save the forward solve but lets the
        # cancellation happen after the amplification by U^{-T} (measured: late centering steps of qp_dense_n2048 then
        # need 13 instead of 8 Newton steps).
        ws.dz.copy_(ws.g)
        L("ipm_gemv_t_f64", d.A.data_ptr(), d.lda, p, n, ws.wv.data_ptr(), 1, p, ws.dz.data_ptr(), n, 1.0, 1.0,
          ws.gt_ws.data_ptr(), ws.gt_ws_n)
        if self.diagonal:
            L("ipm_vec_op_f64", 0, n, ws.hinv.data_ptr(), ws.dz.data_ptr(), ws.dz.data_ptr(), -1.0)
        else:
            self._chol_solve_vec(ws.dz)
            L("ipm_lincomb3_f64", n, -1.0, ws.dz.data_ptr(), 0.0, None, 0.0, None, ws.dz.data_ptr())
        # dv = w - v
        L("ipm_lincomb3_f64", p, 1.0, ws.wv.data_ptr(), -1.0, ws.v.data_ptr(), 0.0, None, ws.dv.data_ptr())

    def _iterate_infeasible(self, z):
        d, ws, L = self.d, self.ws, self.L
        n, p, t = d.n, d.p, self.t
        self._eval(z, ws.cur)
        lin = self._lin_term(z)
        self._gradient(t, lin, ws.cur, ws.g, want_border=False)
        self._direction_infeasible(z, lin)
        # feasibility back-off, then the barrier part of the gradient at the first feasible trial (frozen, Q4)
        self._feasibility(z)
        # cached products (NewtonSolverInfeasibleStart.py:196-205)
        L("ipm_gemv_t_f64", d.A.data_ptr(), d.lda, p, n, ws.v.data_ptr(), 1, p, ws.ATv.data_ptr(), n, 1.0, 0.0,
          ws.gt_ws.data_ptr(), ws.gt_ws_n)
        L("ipm_gemv_t_f64", d.A.data_ptr(), d.lda, p, n, ws.dv.data_ptr(), 1, p, ws.ATdv.data_ptr(), n, 1.0, 0.0,
          ws.gt_ws.data_ptr(), ws.gt_ws_n)
        L("ipm_gemv_n_f64", d.A.data_ptr(), d.lda, p, n, ws.dz.data_ptr(), ws.Adx.data_ptr(), 1.0, 0.0)
        # r0 dual part = g + A'v ;  u1 = t*P dx + A'dv
        L("ipm_lincomb3_f64", n, 1.0, ws.g.data_ptr(), 1.0, ws.ATv.data_ptr(), 0.0, None, ws.r0d.data_ptr())
        if d.is_qp:
            L("ipm_gemv_n_f64", d.P.data_ptr(), d.ldp, n, n, ws.dz.data_ptr(), ws.Pdx.data_ptr(), 1.0, 0.0)
            L("ipm_lincomb3_f64", n, t, ws.Pdx.data_ptr(), 1.0, ws.ATdv.data_ptr(), 0.0, None, ws.u1.data_ptr())
        else:
            L("ipm_lincomb3_f64", n, 1.0, ws.ATdv.data_ptr(), 0.0, None, 0.0, None, ws.u1.data_ptr())
        resume = 0
        host_loop = self.direct_trial or self.update_slacks_every > 0
        while True:
            L("ipm_table_lookup_f64", self.table.data_ptr(), self.table_len, ws.kmax.data_ptr(),
              ws.ls_out.data_ptr() + 48)  # a_k -> ls_out[6]
            L("ipm_trial_point_f64", n, ws.ls_out.data_ptr() + 48, z.data_ptr(), ws.dz.data_ptr(),
              ws.trial.data_ptr())
            self._eval(ws.trial, ws.tri)
            self._gradient(0.0, None, ws.tri, ws.gtrial, want_border=False)  # barrier part only
            # u0 = t*lin + gbar + A'v
            L("ipm_lincomb3_f64", n, t, lin.data_ptr(), 1.0, ws.gtrial.data_ptr(), 1.0, ws.ATv.data_ptr(),
              ws.u0.data_ptr())
            nneg = (ws.red_t.data_ptr() + 32) if self.direct_trial else None
            L("ipm_ls_residual_f64", n, p, ws.r0d.data_ptr(), ws.u0.data_ptr(), ws.u1.data_ptr(), ws.Axb.data_ptr(),
              ws.Adx.data_ptr(), self.table.data_ptr(), self.table_len, ws.kmax.data_ptr(), self.alpha, nneg,
              self.update_slacks_every, resume, ws.ls_out.data_ptr())
            ws.host[:5].copy_(ws.ls_out[:5], non_blocking=True)
            ws.host_info.copy_(ws.info, non_blocking=True)
            if not host_loop:
                break
            torch.cuda.current_stream().synchronize()
            status = float(ws.host[1])
            if status == 3.0:
                ws.kmax.add_(1)
                resume = 0
            elif status == 4.0:
                resume = 1  # update_slacks_every (NewtonSolverInfeasibleStart.py:249-255): refresh the barrier gradient
            else:
                break
        L("ipm_axpy_dev_f64", n, ws.ls_out.data_ptr(), ws.dz.data_ptr(), z.data_ptr())
        L("ipm_axpy_dev_f64", p, ws.ls_out.data_ptr(), ws.dv.data_ptr(), ws.v.data_ptr())
        torch.cuda.current_stream().synchronize()
        _abi.check_device_fault()
        h = ws.host
        return dict(step=float(h[0]), stuck=int(h[1]), r0=float(h[3]), rnorm=float(h[4]),
                    info=int(ws.host_info[0]) or int(ws.host_info[1]))

    def reset_dual(self):
        self.ws.v.zero_()

    def _solve_infeasible(self, z):
        rn = None
        it = 0
        for it in range(self.max_iters):
            zsave, vsave = self._zsave, self._vsave[: self.d.p]
            zsave.copy_(z)
            vsave.copy_(self.ws.v)
            r = self._iterate_infeasible(z)
            if r["info"] != 0:
                r = self._retry_regularised_infeasible(z, zsave, vsave)
            self.newton_steps += 1
            rn = None if r["stuck"] == 2 else r["rnorm"]
            if self.trace is not None:
                self.trace.append((r["step"], rn))
            if r["step"] < STUCK:
                return it + 1, rn, False
            elif rn < self.eps:
                return it + 1, rn, True
        return it + 1, rn, False

    def _retry_regularised_infeasible(self, z, zsave, vsave):
        """Reference: sticky switch to LU solves (NewtonSolverInfeasibleStart.py:491-538).  Device engine:
        regularised re-factorisation (same documented deviation as the feasible-start path)."""
        self.use_backup = True
        ws = self.ws
        if self.shift == self.base_shift:
            hd = ws.hdiag[torch.isfinite(ws.hdiag)]
            base = float(hd.abs().mean()) if hd.numel() else 1.0
            scale = base if base > 0 else 1.0
            self._shift_cap = max(1e-6 * scale, 1e4 * self.base_shift)
            shift = max(100.0 * self.base_shift, 1e-14 * scale)
        else:
            shift = 100.0 * self.shift
        while shift <= self._shift_cap:
            self.shift = shift
            z.copy_(zsave)
            ws.v.copy_(vsave)
            r = self._iterate_infeasible(z)
            if r["info"] == 0:
                return r
            shift *= 100.0
        raise np.linalg.LinAlgError("KKT system is not solvable even after a regularisation of 1e-6 of its diagonal")
