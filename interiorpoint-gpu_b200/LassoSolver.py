"""``LassoSolver`` -- drop-in for the reference's ``LassoSolver.py`` (constructor :18-36, ``solve`` :226-238) on
the B200 engine: batched ADMM for

    minimise_x  1/(2m) ||A x - b_j||^2 + reg_j ||x[1:]||_1        for many columns b_j / reg_j sharing one A.

Setup (AtA, Cholesky, explicit (A'A + m rho I)^{-1}, A'b, Q A'b; LassoSolver.py:126-222) runs on the DMMA
GEMM / Cholesky kernels; every ADMM iteration is ONE fused kernel (csrc/lasso.cu); the batch-coupled stop test
(:273-298) reads four scalars every ``check_stop`` iterations."""

import ctypes as C
import os

import numpy as np
import torch

try:
    from . import _abi
    from ._solver_base import HostArray
    from .engine import F64, Launcher, _round_up, to_dev_matrix
except ImportError:  # flat-module use
    import _abi
    from _solver_base import HostArray
    from engine import F64, Launcher, _round_up, to_dev_matrix


class LassoSolver:
    def __init__(self, A, b, reg=1, rho=0.4, max_iters=1000, check_stop=10, add_bias=False, normalize_A=False,
                 positive=False, compute_loss=False, adaptive_rho=False, eps_abs=1e-4, eps_rel=3e-2, use_gpu=False,
                 num_chunks=0, check_cvxpy=True, _columns=None):
        _abi.require_device()
        self.use_gpu = True
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_chunks = max(1, num_chunks)  # 180 GB of HBM: no automatic chunking (LassoSolver.py:80-94)
        b = np.asarray(b, dtype=np.float64)
        if b.ndim < 2:
            b = b[:, None]
        reg = np.atleast_1d(np.asarray(reg, dtype=np.float64))  # the reference needs a sequence
        assert len(reg) == b.shape[1] or len(reg) == 1 or b.shape[1] == 1
        self.num_samples = max(b.shape[1], len(reg))
        if b.shape[1] == 1 and self.num_samples > 1:
            b = np.repeat(b, self.num_samples, axis=1)
        if len(reg) == 1 and self.num_samples > 1:
            reg = np.repeat(reg, self.num_samples)
        self.b_host, self.reg_host = b, reg
        self.rho, self.max_iters, self.check_stop = rho, max_iters, check_stop
        self.compute_loss = compute_loss
        self.EPS_ABS, self.EPS_REL = eps_abs, eps_rel
        self.positive, self.add_bias = positive, add_bias
        if adaptive_rho:
            raise NotImplementedError("adaptive_rho is marked NOT IMPLEMENTED in the reference (LassoSolver.py:62)")
        A = np.array(A, dtype=np.float64)
        self.m = A.shape[0]
        if normalize_A:
            A = A / A.std(axis=0)  # LassoSolver.py:120-121
        if add_bias:
            A = np.hstack((np.ones((self.m, 1)), A))  # LassoSolver.py:123-129
        self.n = A.shape[1]
        self.L = Launcher(self.device)
        self._setup(A)
        self.gaps = np.zeros((self.max_iters, self.num_samples))
        self.X = np.zeros((self.n, self.num_samples))
        self.feasible, self.cvxpy_vals, self.cvxpy_sols = None, None, None
        self.h2d_bytes = (A.size + b.size + reg.size) * 8
        # keep the ADMM state L2-resident between iterations (ipm_l2_persist); IPM_LASSO_L2=0 switches it off
        self.l2_resident = os.environ.get("IPM_LASSO_L2", "1") != "0"
        self.l2_hit_ratio = 0.0
        # the whole ADMM loop in persistent multi-iteration launches (csrc/lasso_multi.cu); IPM_LASSO_MULTI=0 falls back to
        # one launch per iteration (always used with compute_loss=True, which wants the objective after every iteration)
        self.multi_iteration = os.environ.get("IPM_LASSO_MULTI", "1") != "0"
        self.lookahead = 4  # launches enqueued before the host looks at the device-side stop flag of an older one
        # one chunk: everything is precomputed and resident before solve() (LassoSolver.py:193-222)
        self._prep = self._prepare(np.arange(self.num_samples)) if self.num_chunks == 1 else None

    # ------------------------------------------------------------------------------------------------
    def _gemm(self, A, lda, B, ldb, D, ldd, M, N, K, alpha=1.0, beta=0.0, upper=0):
        self.L("ipm_gemm_tn_f64", A.data_ptr(), lda, B.data_ptr(), ldb, None, alpha, beta, D.data_ptr(), ldd, M, N, K,
               upper)

    def _setup(self, A):
        """Q = (A'A + m rho I)^{-1} explicit, Q~ = -m rho Q (LassoSolver.py:158-222)."""
        dev, n, m, L = self.device, self.n, self.m, self.L
        self.A_dev, self.lda = to_dev_matrix(A, dev)                 # m x n   (contracted index m = rows)
        self.At_dev, self.ldat = to_dev_matrix(np.ascontiguousarray(A.T), dev)  # n x m (for A alpha)
        ldn = _round_up(n, 16)
        H = torch.zeros((n, ldn), dtype=F64, device=dev)
        self._gemm(self.A_dev, self.lda, self.A_dev, self.lda, H, ldn, n, n, m)
        L("ipm_scale_shift_f64", H.data_ptr(), ldn, H.data_ptr(), ldn, n, n, 1.0, m * self.rho)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        L("ipm_potrf_upper_f64", H.data_ptr(), ldn, n, info.data_ptr())
        Y = torch.zeros((n, ldn), dtype=F64, device=dev)
        Y[:, :n].fill_diagonal_(1.0)
        L("ipm_trsm_upper_t_f64", H.data_ptr(), ldn, n, Y.data_ptr(), ldn, n)   # Y = U^{-T}
        self.Qinv = torch.zeros((n, ldn), dtype=F64, device=dev)
        self._gemm(Y, ldn, Y, ldn, self.Qinv, ldn, n, n, n)                      # Q = Y'Y = U^{-1} U^{-T}
        self.Qt = torch.zeros((n, ldn), dtype=F64, device=dev)
        L("ipm_scale_shift_f64", self.Qinv.data_ptr(), ldn, self.Qt.data_ptr(), ldn, n, n, -m * self.rho, 0.0)
        self.ldn = ldn
        if int(info.item()) != 0:
            raise np.linalg.LinAlgError("A'A + m*rho*I is not positive definite")

    def _prepare(self, cols):
        """Host->device transfer of the chunk's b / reg and the cached products A'b, Q A'b (LassoSolver.py:193-216
        for one chunk -- done once in the constructor, like the reference -- and :352-390 per chunk)."""
        dev, n, m = self.device, self.n, self.m
        whole = len(cols) == self.num_samples  # one chunk: no fancy-index copy of the 67 MB right-hand side
        b = self.b_host if whole else np.ascontiguousarray(self.b_host[:, cols])
        reg = self.reg_host if whole else np.ascontiguousarray(self.reg_host[cols])
        K = b.shape[1]
        ld = _round_up(K, 16)
        b_dev, _ = to_dev_matrix(b, dev)
        reg_dev = torch.as_tensor(reg).to(dev)
        eta = reg_dev / self.rho
        # bA and the four state arrays live in ONE allocation so that a single L2 access-policy window covers what
        # every ADMM iteration re-reads (5 n ld doubles = 84 MB at n = 513, K = 4096; see _run)
        block = torch.zeros((5, n, ld), dtype=F64, device=dev)
        z = lambda: torch.zeros((n, ld), dtype=F64, device=dev)  # noqa: E731
        Atb, bA = z(), block[0]
        self._gemm(self.A_dev, self.lda, b_dev, ld, Atb, ld, n, K, m)          # A'b
        self._gemm(self.Qinv, self.ldn, Atb, ld, bA, ld, n, K, n)              # Q A'b   (Q symmetric)
        return dict(K=K, ld=ld, b_dev=b_dev, reg_dev=reg_dev, eta=eta, bA=bA, block=block,
                    state=[block[1], block[2], block[3], block[4]],
                    R=torch.zeros((m, ld), dtype=F64, device=dev), fvals=torch.zeros(K, dtype=F64, device=dev))

    def _run(self, cols, prep=None):
        """ADMM on the column subset ``cols`` (one chunk).  Returns (alpha device [n x K], objective [K], iters)."""
        dev, n, m, L = self.device, self.n, self.m, self.L
        prep = prep or self._prepare(cols)
        K, ld, b_dev, reg_dev, eta, bA = (prep[k] for k in ("K", "ld", "b_dev", "reg_dev", "eta", "bA"))
        alpha, u, z0, z1 = prep["state"]
        for t_ in (alpha, u, z0, z1):
            t_.zero_()  # LassoSolver.py:226-236: every solve() restarts from zero
        npart = _abi.lib().ipm_lasso_partials_doubles(n, K)
        partials = torch.zeros(npart, dtype=F64, device=dev)
        norms = torch.zeros(4, dtype=F64, device=dev)
        host = torch.zeros(4, dtype=F64).pin_memory()
        stop_mult = self.EPS_ABS * np.sqrt(n * K)  # LassoSolver.py:200, 361
        zin, zout = z0, z1
        it = 0
        R, fvals = prep["R"], prep["fvals"]
        block = prep["block"]
        ratio = C.c_double(0.0)
        if self.l2_resident:
            _abi.call("ipm_l2_persist", block.data_ptr(), block.numel() * 8, C.byref(ratio), None)
        try:
            if self.multi_iteration and not self.compute_loss:
                it = self._iterate_multi(prep, alpha, u, z0, z1, stop_mult)
            else:
                it = self._iterate(prep, alpha, u, zin, zout, partials, norms, host, stop_mult, cols, R, fvals)
        finally:
            if self.l2_resident:
                _abi.call("ipm_l2_persist", None, 0, None, None)
        self.l2_hit_ratio = ratio.value
        self._objective(alpha, b_dev, reg_dev, ld, K, R, fvals)
        return alpha, fvals, it

    def _iterate(self, prep, alpha, u, zin, zout, partials, norms, host, stop_mult, cols, R, fvals):
        """The ADMM loop (LassoSolver.py:240-337): one fused kernel per iteration, one host read-back per
        ``check_stop`` iterations.  (Replaying blocks of ``check_stop`` iterations as a CUDA graph was measured on
        B200 and gave nothing -- 100.6 vs 99.4 us per iteration: the gap between the dependent 95 us kernels is
        device-side launch latency, which programmatic dependent launch addresses instead, csrc/lasso.cu.)"""
        n, L = self.n, self.L
        K, ld, b_dev, reg_dev, eta, bA = (prep[k] for k in ("K", "ld", "b_dev", "reg_dev", "eta", "bA"))
        it = 0
        for it in range(self.max_iters):
            check = it % self.check_stop == self.check_stop - 1
            L("ipm_lasso_admm_step_f64", self.Qt.data_ptr(), self.ldn, n, K, bA.data_ptr(), eta.data_ptr(), self.rho,
              alpha.data_ptr(), u.data_ptr(), zin.data_ptr(), zout.data_ptr(), ld, int(self.add_bias),
              int(self.positive), int(check), partials.data_ptr(), norms.data_ptr())
            zin, zout = zout, zin
            if self.compute_loss:
                self._objective(alpha, b_dev, reg_dev, ld, K, R, fvals)
                self.gaps[it, cols] = fvals.cpu().numpy()
            if check:
                host.copy_(norms, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                _abi.check_device_fault()
                r_norm, d_norm, a_norm, u_norm = (float(np.sqrt(v)) for v in host.tolist())
                tol_primal = stop_mult + self.EPS_REL * a_norm
                tol_dual = stop_mult + self.EPS_REL * self.rho * u_norm
                if r_norm < tol_primal and d_norm < tol_dual:
                    break
        return it

    def _iterate_multi(self, prep, alpha, u, z0, z1, stop_mult):
        """The ADMM loop in launches of ``check_stop`` iterations (``ipm_lasso_admm_steps_f64``): the stop test of
        LassoSolver.py:273-298 is evaluated on the device at the end of every launch, a launch that finds the stop flag up
        does nothing, so the host enqueues ``lookahead`` launches ahead and only reads (flag, iterations done) of older
        ones -- no host round trip on the critical path.  Returns the index of the last iteration performed."""
        n, L = self.n, self.L
        K, ld, eta, bA = (prep[k] for k in ("K", "ld", "eta", "bA"))
        nbytes = _abi.lib().ipm_lasso_steps_ws_bytes(n, K)
        ws = prep.get("steps_ws")
        if ws is None or ws.numel() < nbytes:
            ws = prep["steps_ws"] = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        ws.zero_()
        state = ws[:8].view(torch.int32)
        full, rem = divmod(self.max_iters, self.check_stop)
        batches = [(self.check_stop, 1)] * full + ([(rem, 0)] if rem else [])
        LA = self.lookahead
        pinned = torch.zeros((LA, 2), dtype=torch.int32).pin_memory()
        events = [None] * LA
        for b, (n_iters, want_norms) in enumerate(batches):
            if b >= LA:
                events[b % LA].synchronize()
                if int(pinned[b % LA, 0]) != 0:
                    break
            L("ipm_lasso_admm_steps_f64", self.Qt.data_ptr(), self.ldn, n, K, bA.data_ptr(), eta.data_ptr(), self.rho,
              alpha.data_ptr(), u.data_ptr(), z0.data_ptr(), z1.data_ptr(), ld, int(self.add_bias), int(self.positive),
              n_iters, want_norms, stop_mult, self.EPS_REL, ws.data_ptr())
            pinned[b % LA].copy_(state[:2], non_blocking=True)
            events[b % LA] = torch.cuda.Event()
            events[b % LA].record()
        pinned[0].copy_(state[:2], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        _abi.check_device_fault()
        return int(pinned[0, 1]) - 1

    def _objective(self, alpha, b_dev, reg_dev, ld, K, R, out):
        """f = 1/(2m) ||A alpha - b||^2 + reg ||alpha[1:]||_1 per column (LassoSolver.py:314-325)."""
        R.copy_(b_dev)
        self._gemm(self.At_dev, self.ldat, alpha, ld, R, ld, self.m, K, self.n, alpha=1.0, beta=-1.0)
        self.L("ipm_lasso_objective_f64", R.data_ptr(), ld, self.m, alpha.data_ptr(), ld, self.n, K,
               reg_dev.data_ptr(), int(self.add_bias), int(self.positive), out.data_ptr())

    # ------------------------------------------------------------------------------------------------
    def solve(self):
        """Returns ``(X, solutions, gaps, iterations)`` like the reference (LassoSolver.py:337, 485): one chunk
        reports ``iteration + 1``, several chunks report the list of ``iteration`` per chunk (:479)."""
        Ktot = self.num_samples
        if self.num_chunks == 1:
            alpha, f, it = self._run(np.arange(Ktot), self._prep)
            self.alpha = alpha[:, :Ktot]
            self.X = HostArray(self.alpha.cpu().numpy())
            self.solutions = HostArray(f.cpu().numpy())
            self.num_iterations = [it + 1]
            return self.X, self.solutions, self.gaps[: it + 1], it + 1
        X = np.zeros((self.n, Ktot))
        sol = np.empty(Ktot)
        its = []
        idx = np.arange(Ktot)
        for i in range(self.num_chunks):
            cols = idx[i:: self.num_chunks]  # LassoSolver.py:349-351
            alpha, f, it = self._run(cols)
            X[:, cols] = alpha[:, : len(cols)].cpu().numpy()
            sol[cols] = f.cpu().numpy()
            its.append(it)
        self.X, self.solutions, self.num_iterations = HostArray(X), HostArray(sol), its
        return self.X, self.solutions, self.gaps, its

    def objective(self):
        """Objective per problem at the current solution (the reference's version has its ``positive`` branches
        swapped, LassoSolver.py:503-508; this returns the documented quantity = ``solutions``)."""
        return self.solutions

    def plot(self):
        import matplotlib.pyplot as plt

        if not self.compute_loss:
            raise ValueError("Need compute_loss=True to plot convergence")
        ax = plt.subplot()
        ax.plot(self.gaps[: self.num_iterations[0]])
        ax.set_yscale("log")
        return ax
