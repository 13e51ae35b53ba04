"""Row-sharded Newton centering for ONE large LP / QP / SOCP across the GPUs of a node.

Every rank keeps a contiguous block of the inequality rows ``C[lo:hi], d[lo:hi]`` (SOCP: a contiguous range of whole
cones, i.e. of the rows of the stacked matrix W) resident in its HBM; the iterate, the bounds and the objective are
replicated (bound / objective terms are contributed by rank 0 only).  Per Newton
iteration the ranks exchange, over NCCL (NVLink 5 / NVSwitch):

    barrier sums (4 doubles, SUM; min slack, MIN) -> gradient partial (n doubles, SUM)
    -> partial Hessian C_r' diag(w_r) C_r (the n x ld buffer, SUM) -> step index kmax (int, MAX)
    -> frozen log-sum of the Armijo test (1 double, SUM)

and then run the factorisation, triangular solves and the scalar line-search logic REPLICATED on identical data,
so every rank takes the same control-flow decisions without a broadcast.  The reference has no multi-GPU path;
this follows BASELINE.json's north_star ("partial A'DA per GPU + NCCL allreduce before a replicated factorisation").
"""

import torch
import torch.distributed as dist

try:
    from . import _abi
    from .cone_engine import ConeNewton
    from .engine import F64, LinearNewton
except ImportError:  # flat-module use
    import _abi
    from cone_engine import ConeNewton
    from engine import F64, LinearNewton


class _RowSharded:
    """Mixin (placed before the engine class in the MRO): wraps the barrier pieces of the single-GPU engine with the
    collectives that make their outputs global."""

    def _init_sharding(self, group):
        if self.equality:
            raise NotImplementedError("row sharding is implemented for the feasible-start method (no A x = b)")
        if self.update_slacks_every > 0 or self.diagonal:
            raise NotImplementedError("row sharding supports the default dense Cholesky path only")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.ws.Lsum = torch.zeros(1, dtype=F64, device=self.d.device)
        self.ws.mn = torch.zeros(1, dtype=F64, device=self.d.device)
        self.comm_bytes = 0

    def _sum(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        self.comm_bytes += t.numel() * t.element_size()

    def _eval(self, z, slot):
        super()._eval(z, slot)
        self.ws.mn.copy_(slot.red[1:2])
        self._sum(slot.red)  # sum log, (min: fixed below), sum inv, sum inv^2, #negative
        dist.all_reduce(self.ws.mn, op=dist.ReduceOp.MIN, group=self.group)
        slot.red[1:2].copy_(self.ws.mn)

    def _gradient(self, t, lin, slot, g, want_border):
        ws = self.ws
        super()._gradient(t, lin if self.rank == 0 else None, slot, g, want_border)
        if self.phase1 and self.rank != 0:
            g[self.d.n:].zero_()  # g_s = t - sum(inv) is global already: contributed once
        self._sum(g)
        if self.phase1 and want_border:
            self._sum(ws.hxs)

    def _hessian(self, t):
        with self.L.timed_range("hessian_formation"):
            self._hessian_impl(t)

    def _feasibility(self, z):
        super()._feasibility(z)
        dist.all_reduce(self.ws.kmax, op=dist.ReduceOp.MAX, group=self.group)

    def slacks_at(self, x):
        raise NotImplementedError("dual variables are not gathered in row-sharded mode")


class ShardedLinearNewton(_RowSharded, LinearNewton):
    def __init__(self, data, group=None, **kw):
        super().__init__(data, **kw)
        self._init_sharding(group)

    def _hessian_impl(self, t):
        d, ws, L = self.d, self.ws, self.L
        n, m = d.n, d.m
        beta = 0.0
        if d.is_qp and not self.phase1 and self.rank == 0:
            L("ipm_scale_copy_upper_f64", ws.H.data_ptr(), ws.ldh, d.P.data_ptr(), d.ldp, n, t)
            beta = 1.0
        L.tag = "hessian"
        L("ipm_gemm_tn_f64", d.C.data_ptr(), d.ldc, d.C.data_ptr(), d.ldc, ws.w.data_ptr(), 1.0, beta, ws.H.data_ptr(),
          ws.ldh, n, n, m, 1)
        L.tag = None
        if self.rank == 0:  # bound diagonal: rank 0 owns the bound rows
            L("ipm_hess_finish_f64", ws.H.data_ptr(), ws.ldh, n, ws.hdiag.data_ptr(), None, None, 0.0)
        self._sum(ws.H[:n] if not self.phase1 else ws.H)
        shift = self.shift + (1e-9 if self.use_psd_condition else 0.0)
        if self.phase1 or shift:
            L("ipm_hess_finish_f64", ws.H.data_ptr(), ws.ldh, n, None, ws.hxs.data_ptr() if self.phase1 else None,
              (ws.red.data_ptr() + 24) if self.phase1 else None, shift)

    def _verify_trial(self, z):
        """Global frozen log-sum for the Armijo test: local partial at a = table[kmax], SUM all-reduce."""
        d, ws, L = self.d, self.ws, self.L
        L("ipm_ls_logsum_f64", d.n_slacks, ws.slacks.data_ptr(), ws.p1.data_ptr(), self._p2_ptr(),
          self.table.data_ptr(), self.table_len, ws.kmax.data_ptr(), ws.Lsum.data_ptr())
        self._sum(ws.Lsum)
        return ws.Lsum.data_ptr(), None


class ShardedConeNewton(_RowSharded, ConeNewton):
    """SOCP: every rank owns a contiguous range of whole cones (its slice of the stacked W).  The Armijo test
    already evaluates the barrier AT the trial point (ConeNewton._verify_trial -> _eval), whose reductions the mixin
    all-reduces, so frozen log-sum and feasibility verdict are global without further exchange."""

    def __init__(self, data, group=None, **kw):
        super().__init__(data, **kw)
        self._init_sharding(group)

    def _hessian_impl(self, t):
        d, ws, L = self.d, self.ws, self.L
        n = d.n
        beta = 0.0
        if d.is_qp and not self.phase1 and self.rank == 0:
            L("ipm_scale_copy_upper_f64", ws.H.data_ptr(), ws.ldh, d.P.data_ptr(), d.ldp, n, t)
            beta = 1.0
        L.tag = "hessian"
        L("ipm_gemm_tn_f64", d.W.data_ptr(), d.ldw, d.W.data_ptr(), d.ldw, ws.wts.data_ptr(), 1.0, beta, ws.H.data_ptr(),
          ws.ldh, n, n, d.rows_w, 1)
        L.tag = None
        if d.nbounds:  # only rank 0 holds bounds
            L("ipm_hess_finish_f64", ws.H.data_ptr(), ws.ldh, n, ws.hdiag.data_ptr(), None, None, 0.0)
        self._sum(ws.H[:n] if not self.phase1 else ws.H)
        shift = self.shift + (1e-9 if self.use_psd_condition else 0.0)
        if self.phase1 or shift:
            L("ipm_hess_finish_f64", ws.H.data_ptr(), ws.ldh, n, None, ws.hxs.data_ptr() if self.phase1 else None,
              (ws.red.data_ptr() + 24) if self.phase1 else None, shift)
