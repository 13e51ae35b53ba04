"""Library context for the hand-written kernels (tools only -- nothing here is on the product path): what cuBLAS /
cuSOLVER, i.e. the reference's own GPU arm (CuPy -> cuBLAS DGEMM, cuSOLVER DPOTRF), achieve for the same operations on the
same box, next to the C-ABI kernels.

  Hessian  C' diag(w) C,  C: m x n      reference arm: elementwise scale + cublasDgemm (cp.matmul(C.T, w[:,None]*C),
                                        FunctionManager.py:301-312); also scale + cublasDsyrk (half the flops)
  Cholesky n x n                        cusolverDnDpotrf through torch.linalg.cholesky (cp.linalg.cholesky, NewtonSolver.py:286)
  TRSM     U^{-T} B, B: n x p           cublasDtrsm through torch.linalg.solve_triangular
  Lasso    Q (n x n) @ Z (n x K)        cublasDgemm (LassoSolver.py:245-249)

    python tools/lib_context.py [n] [m]         -> one JSON line (default n = 8192, m = 16384)"""
import ctypes as C
import glob
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ipm_b200 import _abi  # noqa: E402

PEAK = 37.1
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * n
_abi.require_device()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(n)
Cm = torch.rand((m, n), dtype=torch.float64, device=dev, generator=g) * 4 - 2
w = torch.rand(m, dtype=torch.float64, device=dev, generator=g) + 0.5
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)


def timed(fn, reps=4):
    ts = []
    for _ in range(reps + 1):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[1:]))


def entry(ms, flop):
    tf = flop / (ms * 1e-3) / 1e12
    return {"ms": ms, "tflops": tf, "frac_fp64_peak": tf / PEAK}


out = {"n": n, "m": m, "peak_tflops": PEAK}
syrk_flop = float(m) * n * (n + 1)

# ---- Hessian: ours
H = torch.zeros((n, n), dtype=torch.float64, device=dev)
ms = timed(lambda: _abi.call("ipm_gemm_tn_f64", Cm.data_ptr(), n, Cm.data_ptr(), n, w.data_ptr(), 1.0, 0.0, H.data_ptr(),
                             n, n, n, m, 1, None))
out["hessian_ipm_gemm_tn"] = dict(entry(ms, syrk_flop), what="ipm_gemm_tn_f64, weights fused, upper tiles only")
# ---- Hessian: the reference arm's formulation on cuBLAS (scale + full DGEMM)
scaled = torch.empty_like(Cm)


def ref_hessian():
    torch.mul(Cm, w[:, None], out=scaled)
    torch.mm(Cm.T, scaled, out=H)


ms = timed(ref_hessian)
out["hessian_cublas_dgemm"] = dict(entry(ms, syrk_flop), what="w[:,None]*C then cublasDgemm C'(.) (2x the SYRK flops; "
                                                                "TFLOP/s quoted on the SYRK flop count)")
# ---- Hessian: scale by sqrt(w) + cublasDsyrk
try:
    lib = None
    for pat in (os.path.join(os.path.dirname(torch.__file__), "lib", "libcublas.so*"),
                os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cublas", "lib", "libcublas.so*"),
                "/usr/local/cuda/lib64/libcublas.so*"):
        hits = sorted(glob.glob(pat))
        if hits:
            lib = C.CDLL(hits[0])
            break
    handle = C.c_void_p()
    assert lib.cublasCreate_v2(C.byref(handle)) == 0
    lib.cublasSetStream_v2(handle, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    one, zero = C.c_double(1.0), C.c_double(0.0)
    sw = torch.sqrt(w)

    def syrk():
        torch.mul(Cm, sw[:, None], out=scaled)
        # row-major m x n == column-major n x m (lda = n): C_colmajor = A A' with A = scaled' (n x m): op N, k = m
        rc = lib.cublasDsyrk_v2(handle, 1, 0, n, m, C.byref(one), C.c_void_p(scaled.data_ptr()), n, C.byref(zero),
                                C.c_void_p(H.data_ptr()), n)
        assert rc == 0, rc

    ms = timed(syrk)
    out["hessian_cublas_dsyrk"] = dict(entry(ms, syrk_flop), what="sqrt(w)[:,None]*C then cublasDsyrk")
    lib.cublasDestroy_v2(handle)
except Exception as e:  # noqa: BLE001
    out["hessian_cublas_dsyrk"] = {"error": repr(e)}

# ---- Cholesky
_abi.call("ipm_gemm_tn_f64", Cm.data_ptr(), n, Cm.data_ptr(), n, w.data_ptr(), 1.0, 0.0, H.data_ptr(), n, n, n, m, 1, None)
Hfull = torch.triu(H) + torch.triu(H, 1).T
Hfull.diagonal().add_(1.0)
work = torch.empty_like(Hfull)
info = torch.zeros(1, dtype=torch.int32, device=dev)


def ours_potrf():
    _abi.call("ipm_potrf_upper_f64", work.data_ptr(), n, n, info.data_ptr(), None)


def run_with_copy(fn):
    def f():
        work.copy_(Hfull)
        fn()
    return f


copy_ms = timed(lambda: work.copy_(Hfull))
ms = timed(run_with_copy(ours_potrf)) - copy_ms
out["potrf_ipm"] = dict(entry(ms, n ** 3 / 3.0), what="ipm_potrf_upper_f64 (pipelined tile-DAG kernel)")
ms = timed(run_with_copy(lambda: torch.linalg.cholesky(work, upper=True, out=work))) - copy_ms
out["potrf_cusolver"] = dict(entry(ms, n ** 3 / 3.0), what="torch.linalg.cholesky -> cusolverDnDpotrf (includes its "
                                                             "workspace query and info check)")
# ---- TRSM
p = n // 4
U = torch.linalg.cholesky(Hfull, upper=True)
B = torch.rand((n, p), dtype=torch.float64, device=dev, generator=g)
Bw = torch.empty_like(B)
ms = timed(lambda: (Bw.copy_(B), _abi.call("ipm_trsm_upper_t_f64", U.data_ptr(), n, n, Bw.data_ptr(), p, p, None)))
out["trsm_ipm_stream_ordered"] = dict(entry(ms, float(n) * n * p), what="ipm_trsm_upper_t_f64 (stream-ordered panels)")
ms = timed(lambda: torch.linalg.solve_triangular(U.T, B, upper=False, out=Bw))
out["trsm_cublas"] = dict(entry(ms, float(n) * n * p), what="torch.linalg.solve_triangular -> cublasDtrsm")
Hw = torch.empty_like(Hfull)
ms = timed(lambda: (Hw.copy_(Hfull), Bw.copy_(B), _abi.call("ipm_potrf_trsm_upper_f64", Hw.data_ptr(), n, n, Bw.data_ptr(), p, p,
                                                             info.data_ptr(), None))) - copy_ms
out["potrf_trsm_ipm_fused"] = dict(entry(ms, n ** 3 / 3.0 + float(n) * n * p), what="ipm_potrf_trsm_upper_f64: one launch")
# ---- Lasso product
nl, K = 513, 4096
Q = torch.rand((nl, nl), dtype=torch.float64, device=dev, generator=g)
Z = torch.rand((nl, K), dtype=torch.float64, device=dev, generator=g)
X = torch.empty_like(Z)
ms = timed(lambda: torch.mm(Q, Z, out=X), reps=20)
out["lasso_product_cublas_dgemm"] = dict(entry(ms, 2.0 * nl * nl * K), what="Q (513x513) @ Z (513x4096): the GEMM alone, "
                                                                          "without prox / dual update / norms")
print(json.dumps(out))
