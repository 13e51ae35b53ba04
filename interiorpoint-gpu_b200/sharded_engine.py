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

import ctypes as C
import os

import torch
import torch.distributed as dist

try:
    from . import _abi
    from .cone_engine import ConeNewton
    from .engine import F64, LinearNewton, hess_i8_slices
except ImportError:  # flat-module use
    import _abi
    from cone_engine import ConeNewton
    from engine import F64, LinearNewton, hess_i8_slices


_PEER_CACHE = {}  # (H shape, world, group, device) -> peer-mapped buffers and their bookkeeping


PEER_POTRF_MIN_N = 12288  # crossover of the two measured sizes, see _setup_peer


class _UseCached(Exception):
    pass


class _RowSharded:
    """Mixin (placed before the engine class in the MRO): wraps the barrier pieces of the single-GPU engine with the
    collectives that make their outputs global."""

    def _init_sharding(self, group):
        # Equality constraints (infeasible-start method, NewtonSolverInfeasibleStart.py:386-511): the equality rows A, the
        # dual iterate and the block elimination (TRSM of A', Schur complement, its factorisation) are replicated; only the
        # barrier pieces -- slacks, gradient, Hessian, feasibility back-off -- are sharded, through the same overrides.
        if self.update_slacks_every > 0 or self.diagonal or self.linear_solver != "cholesky":
            raise NotImplementedError("row sharding supports the default dense Cholesky path only")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.ws.Lsum = torch.zeros(1, dtype=F64, device=self.d.device)
        self.ws.mn = torch.zeros(1, dtype=F64, device=self.d.device)
        self.comm_bytes = 0
        self.peer = None
        if os.environ.get("IPM_PEER_HESSIAN", "1") != "0" and self.world <= 8:
            self._setup_peer_memory()

    # ------------------------------------------------------------------ Hessian exchange over peer memory
    def _setup_peer_memory(self):
        """Symmetric (peer-mapped) allocations for the fused SYRK + reduce-scatter + all-gather (csrc/gemm_tn.cu,
        ipm_syrk_scatter_f64 / ipm_hess_reduce_bcast_f64): the Hessian buffer itself, an inbox of partial tiles and the
        arrival flags / completion counter.  Falls back to the NCCL all-reduce when the build of torch or the box has
        no symmetric memory (all ranks take the same decision: the probe result is MIN-reduced)."""
        ok = torch.ones(1, dtype=torch.int32, device=self.d.device)
        st = None
        # The peer-mapped buffers are kept per (shape, group) for the life of the process and handed to the next engine
        # of the same shape: allocating and exchanging handles costs ~0.1 s per buffer, which is what a user who builds
        # a new solver object per problem would otherwise pay every time.  Epochs and counters simply run on.
        key = (tuple(self.ws.H.shape), self.world, id(self.group), self.d.device.index)
        cached = _PEER_CACHE.get(key)
        try:
            if cached is not None:
                raise _UseCached
            import torch.distributed._symmetric_memory as symm

            grp = self.group if self.group is not None else dist.group.WORLD
            n, ws = self.d.n, self.ws
            T = (n + 127) // 128
            slots = (T * (T + 1) // 2 + self.world - 1) // self.world
            Hs = symm.empty((ws.H.shape[0], ws.H.shape[1]), dtype=F64, device=self.d.device)
            inbox = symm.empty((self.world * slots * 128 * 128,), dtype=F64, device=self.d.device)
            sig = symm.empty((slots * self.world + 64,), dtype=torch.int32, device=self.d.device)
            # distributed factorisation (ipm_potrf_upper_peer_f64): progress counters and the info word of every rank
            words = _abi.lib().ipm_potrf_peer_prog_words()
            prog = symm.empty((words,), dtype=torch.int64, device=self.d.device)
            info = symm.empty((4,), dtype=torch.int32, device=self.d.device)
            Hs.zero_()
            sig.zero_()
            prog.zero_()
            info.zero_()
            hH, hI, hS, hP, hN = (symm.rendezvous(t, grp.group_name) for t in (Hs, inbox, sig, prog, info))
            st = dict(slots=slots, tiles=T * (T + 1) // 2, H=Hs, inbox=inbox, sig=sig, prog=prog, info=info,
                      handles=(hH, hI, hS, hP, hN), done_off=slots * self.world, epoch=0, target=0, potrf_epoch=0)
            R = self.world
            st["p_inbox"] = (C.c_void_p * R)(*[int(p) for p in hI.buffer_ptrs])
            st["p_flags"] = (C.c_void_p * R)(*[int(p) for p in hS.buffer_ptrs])
            st["p_H"] = (C.c_void_p * R)(*[int(p) for p in hH.buffer_ptrs])
            st["p_done"] = (C.c_void_p * R)(*[int(p) + 4 * st["done_off"] for p in hS.buffer_ptrs])
            st["p_prog"] = (C.c_void_p * R)(*[int(p) for p in hP.buffer_ptrs])
            st["p_info"] = (C.c_void_p * R)(*[int(p) for p in hN.buffer_ptrs])
        except _UseCached:
            st = cached
        except Exception as e:  # noqa: BLE001 -- any failure means "no peer path on this box"
            ok.zero_()
            self.peer_error = repr(e)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 1:
            _PEER_CACHE[key] = st
            self.peer = st
            self.ws.H = st["H"]  # the factorisation runs in the peer-mapped buffer the owners write into
            self.ws.info = st["info"][:2]  # same layout as the engine's own info pair
            # block columns of the factorisation dealt over the ranks (csrc/chol.cu, struct dag::Peers).  Measured on
            # 8 B200s (profiles/scale_r02_n8_*.json): 26.1 ms against 45.7 ms replicated at n = 16384, but 11.8 ms
            # against 7.0 ms at n = 8192, where one block column's panel latency plus its NVLink pushes exceed the
            # trailing update they overlap with -- so the replicated factorisation stays the default below
            # PEER_POTRF_MIN_N.  IPM_PEER_POTRF=1 / 0 forces it on / off.
            env = os.environ.get("IPM_PEER_POTRF", "")
            self.peer_potrf = self.nz > 384 and (env == "1" or (env != "0" and self.nz >= PEER_POTRF_MIN_N))
            torch.cuda.synchronize()
            dist.barrier(group=self.group)

    def _peer_hessian(self, rows_ptr, ld, K, w_ptr, P=None, ldp=0, tP=0.0):
        """H (upper, n x n block of ws.H) = sum over ranks of rows' diag(w) rows [+ tP * P], identical bits on every
        rank, without NCCL."""
        pr, ws, L, d = self.peer, self.ws, self.L, self.d
        pr["epoch"] += 1
        pr["target"] = (pr["target"] + pr["tiles"]) & 0xFFFFFFFF
        L.tag = "hessian"
        slices = self._shard_i8_slices(K)
        if slices:
            # this rank's rows on the INT8 tensor pipe (csrc/hess_i8.cu): the partial tiles stay in this rank's exchange
            # buffer, filed under their owners, who pull them over NVLink
            L("ipm_hess_i8_scatter_f64", rows_ptr, ld, K, d.n, w_ptr, slices, self._hess_i8_ws(slices, K).data_ptr(),
              pr["p_inbox"], pr["p_flags"], self.rank, self.world, pr["slots"], pr["epoch"])
            L.tag = None
            L("ipm_hess_reduce_bcast_pull_f64", pr["p_inbox"], pr["sig"].data_ptr(), pr["p_H"], pr["p_done"], ws.ldh, d.n,
              self.rank, self.world, pr["slots"], pr["epoch"], pr["target"], _abi.ptr(P), ldp, tP)
        else:
            L("ipm_syrk_scatter_f64", rows_ptr, ld, w_ptr, d.n, K, 1.0, None, 0, pr["p_inbox"], pr["p_flags"], self.rank,
              self.world, pr["slots"], pr["epoch"])
            L.tag = None
            L("ipm_hess_reduce_bcast_f64", pr["inbox"].data_ptr(), pr["sig"].data_ptr(), pr["p_H"], pr["p_done"], ws.ldh,
              d.n, self.rank, self.world, pr["slots"], pr["epoch"], pr["target"], _abi.ptr(P), ldp, tP)
        self.comm_bytes += 2 * pr["tiles"] * 128 * 128 * 8 * (self.world - 1) // self.world

    def _shard_i8_slices(self, K):
        """Digits of the INT8 Hessian kernel for this rank's K rows, or 0 -- agreed on by all ranks (push and pull
        exchanges do not mix; the row blocks differ by one row at most, but a threshold may fall between them)."""
        cache = self.__dict__.setdefault("_i8_agreed", {})
        if K not in cache:
            t = torch.tensor([hess_i8_slices(K, self.d.n)], dtype=torch.int32, device=self.d.device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
            cache[K] = int(t.item())
        return cache[K]

    def _fuse_forward(self):
        return super()._fuse_forward() and not getattr(self, "peer_potrf", False)  # the distributed kernel has no RHS

    def _factor(self):
        """Distributed tile-DAG Cholesky over the ranks: block column j on rank j % R, finished rows pushed into every
        rank's copy of H over NVLink, so every rank ends with the whole factor (the triangular solves stay replicated)."""
        pr = self.peer
        if pr is None or not getattr(self, "peer_potrf", False):
            return super()._factor()
        pr["potrf_epoch"] = pr["potrf_epoch"] % 0x7FFFFFFF + 1
        with self.L.timed_range("factorisation"):
            self.L("ipm_potrf_upper_peer_f64", pr["p_H"], self.ws.ldh, self.nz, pr["p_info"], pr["p_prog"], self.rank,
                   self.world, pr["potrf_epoch"], 0)

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
        """All slacks in the reference's layout [inequality rows / cones of rank 0, 1, ... | bound rows] gathered on every
        rank (``get_dual_variables``: lam = 1 / (t slacks), LPSolver.py:641-646)."""
        d, ws = self.d, self.ws
        super()._eval(x, ws.cur)  # the engine's own evaluation: this rank's rows only, no reduction
        n_rows = getattr(d, "M", None) if d.m == 0 and hasattr(d, "M") else d.m
        sizes = torch.tensor([n_rows, d.n_slacks - n_rows], dtype=torch.int64, device=d.device)
        allsz = [torch.zeros_like(sizes) for _ in range(self.world)]
        dist.all_gather(allsz, sizes, group=self.group)
        allsz = [t.tolist() for t in allsz]
        width = max(max(a) for a in allsz)
        pad = torch.zeros(2, max(width, 1), dtype=F64, device=d.device)
        pad[0, :n_rows] = ws.slacks[:n_rows]
        pad[1, :d.n_slacks - n_rows] = ws.slacks[n_rows:d.n_slacks]
        parts = [torch.zeros_like(pad) for _ in range(self.world)]
        dist.all_gather(parts, pad, group=self.group)
        rows = [parts[r][0, :allsz[r][0]] for r in range(self.world)]
        bounds = [parts[r][1, :allsz[r][1]] for r in range(self.world)]
        return torch.cat(rows + bounds)


class ShardedLinearNewton(_RowSharded, LinearNewton):
    def __init__(self, data, group=None, **kw):
        super().__init__(data, **kw)
        self._init_sharding(group)

    def _hessian_impl(self, t):
        d, ws, L = self.d, self.ws, self.L
        n, m = d.n, d.m
        if self.peer is not None and m > 0:
            qp = d.is_qp and not self.phase1
            if self.rank != 0:
                ws.hdiag.zero_()  # bound rows belong to rank 0
            self._sum(ws.hdiag)
            self._peer_hessian(d.C.data_ptr(), d.ldc, m, ws.w.data_ptr(), d.P if qp else None, d.ldp if qp else 0,
                               t if qp else 0.0)
            shift = self.shift + (1e-9 if self.use_psd_condition else 0.0)
            L("ipm_hess_finish_f64", ws.H.data_ptr(), ws.ldh, n, ws.hdiag.data_ptr(),
              ws.hxs.data_ptr() if self.phase1 else None, (ws.red.data_ptr() + 24) if self.phase1 else None, shift)
            return
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
        flag = torch.tensor([1 if data.nbounds else 0], dtype=torch.int32, device=data.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        self.bounds_anywhere = bool(int(flag.item()))

    def _hessian_impl(self, t):
        d, ws, L = self.d, self.ws, self.L
        n = d.n
        if self.peer is not None:
            qp = d.is_qp and not self.phase1
            if self.bounds_anywhere:
                if not d.nbounds:
                    ws.hdiag.zero_()
                self._sum(ws.hdiag)
            self._peer_hessian(d.W.data_ptr(), d.ldw, d.rows_w, ws.wts.data_ptr(), d.P if qp else None,
                               d.ldp if qp else 0, t if qp else 0.0)
            shift = self.shift + (1e-9 if self.use_psd_condition else 0.0)
            L("ipm_hess_finish_f64", ws.H.data_ptr(), ws.ldh, n, ws.hdiag.data_ptr() if self.bounds_anywhere else None,
              ws.hxs.data_ptr() if self.phase1 else None, (ws.red.data_ptr() + 24) if self.phase1 else None, shift)
            return
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
