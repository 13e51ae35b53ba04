"""Per-kernel timing on the B200 (CUDA events, L2 flushed between iterations).  Usage:
    python tools/kernel_bench.py [n] [m]"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ipm_b200 import _abi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * n
_abi.require_device()
g = torch.Generator(device="cuda").manual_seed(0)
C_ = torch.rand((m, n), dtype=torch.float64, device="cuda", generator=g) * 4 - 2
w = torch.rand(m, dtype=torch.float64, device="cuda", generator=g) + 0.5
H = torch.zeros((n, n), dtype=torch.float64, device="cuda")
x = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
y = torch.zeros(m, dtype=torch.float64, device="cuda")
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
info = torch.zeros(1, dtype=torch.int32, device="cuda")
nws = _abi.lib().ipm_gemv_t_ws_doubles(m, n, 1)
ws = torch.empty(nws, dtype=torch.float64, device="cuda")
g_out = torch.zeros(n, dtype=torch.float64, device="cuda")


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


def syrk():
    _abi.call("ipm_gemm_tn_f64", C_.data_ptr(), n, C_.data_ptr(), n, w.data_ptr(), 1.0, 0.0, H.data_ptr(), n, n, n, m, 1, None)


def potrf():
    _abi.call("ipm_potrf_upper_f64", Hf.data_ptr(), n, n, info.data_ptr(), None)


res = {"n": n, "m": m}
t = timeit(syrk)
res["syrk_ms"] = t
res["syrk_tflops"] = m * n * (n + 1) / (t[1] * 1e-3) / 1e12
Hs = H.clone()
Hs.diagonal().add_(1.0)
Hf = Hs.clone()


def potrf_fresh():
    Hf.copy_(Hs)
    potrf()


tcopy = timeit(lambda: Hf.copy_(Hs))
t = timeit(potrf_fresh)
res["potrf_ms"] = (t[0] - tcopy[0], t[1] - tcopy[1])
res["potrf_tflops"] = n ** 3 / 3 / ((t[1] - tcopy[1]) * 1e-3) / 1e12
res["potrf_info"] = int(info.item())
b = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
tws = torch.zeros(n, dtype=torch.float64, device="cuda")
t = timeit(lambda: (_abi.call("ipm_trsv_upper_f64", Hf.data_ptr(), n, n, b.data_ptr(), 1, tws.data_ptr(), None),
                    _abi.call("ipm_trsv_upper_f64", Hf.data_ptr(), n, n, b.data_ptr(), 0, tws.data_ptr(), None)))
res["trsv2_ms"] = t
t = timeit(lambda: _abi.call("ipm_gemv_n_f64", C_.data_ptr(), n, m, n, x.data_ptr(), y.data_ptr(), 1.0, 0.0, None))
res["gemv_n_ms"] = t
res["gemv_n_gbs"] = 8.0 * m * n / (t[1] * 1e-3) / 1e9
t = timeit(lambda: _abi.call("ipm_gemv_t_f64", C_.data_ptr(), n, m, n, y.data_ptr(), 1, m, g_out.data_ptr(), n, 1.0, 0.0,
                             ws.data_ptr(), nws, None))
res["gemv_t_ms"] = t
res["gemv_t_gbs"] = 8.0 * m * n / (t[1] * 1e-3) / 1e9
print(json.dumps(res))
