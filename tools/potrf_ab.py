"""Stream-ordered vs single-launch tile-DAG Cholesky on one box: correctness at ragged sizes, then timing.
    python tools/potrf_ab.py [n ...]      (timing sizes, default 8192)"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ipm_b200 import _abi  # noqa: E402

_abi.require_device()
import ctypes as C  # noqa: E402

NAMES = ["ipm_potrf_upper_f64", "ipm_potrf_upper_dag_f64"]
for extra in ("ipm_internal_potrf_stream_f64", "ipm_internal_potrf_dag1_f64"):  # library-internal A/B entry points
    if hasattr(_abi.lib(), extra):
        fn = getattr(_abi.lib(), extra)
        fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        NAMES.append(extra)


def spd(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    C_ = torch.rand((2 * n, n), dtype=torch.float64, device="cuda", generator=g) * 4 - 2
    w = 10.0 ** (torch.rand(2 * n, dtype=torch.float64, device="cuda", generator=g) * 8 - 4)
    H = (C_.T * w) @ C_
    H.diagonal().add_(1e-3)
    return H


def call(name, *args):
    _abi.check(getattr(_abi.lib(), name)(*args), name)


def factor(name, H, n, ld):
    Hd = torch.full((n, ld), float("nan"), dtype=torch.float64, device="cuda")
    Hd[:, :n] = torch.triu(H)  # strict lower triangle = 0 here; it must stay untouched
    info = torch.full((1,), -7, dtype=torch.int32, device="cuda")
    call(name, Hd.data_ptr(), ld, n, info.data_ptr(), None)
    torch.cuda.synchronize()
    assert _abi.lib().ipm_device_fault() == 0, ("watchdog fired", name, n, _abi.lib().ipm_device_fault())
    return Hd, int(info.item())


for n in (385, 1024, 2500, 4097):
    H = spd(n, n)
    ld = (n + 15) // 16 * 16
    out = {}
    for name in NAMES:
        Hd, info = factor(name, H, n, ld)
        U = torch.triu(Hd[:, :n])
        err = float(((U.T @ U - H).abs().max() / H.abs().max()).item())
        low = float(torch.tril(Hd[:, :n], -1).abs().max().item())
        out[name] = (info, err, low)
    print(f"n={n}: " + "  ".join(f"{k[4:-4]}: info={v[0]} err={v[1]:.1e} lower={v[2]:.0e}" for k, v in out.items()), flush=True)
    for v in out.values():
        assert v[0] == 0 and v[1] < 1e-13 and v[2] == 0.0, out

# first bad pivot
n = 700
H = spd(n, 7)
H[600, 600] = -1.0
for name in NAMES:
    _, info = factor(name, H, n, 704)
    print(name, "bad pivot info =", info, flush=True)
    assert info == 601

flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
for n in [int(a) for a in sys.argv[1:]] or [8192]:
    H = spd(n, 1)
    work = torch.empty_like(H)
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    for name in NAMES:
        ts = []
        for rep in range(6):
            work.copy_(H)
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            call(name, work.data_ptr(), n, n, info.data_ptr(), None)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = float(np.median(ts[1:]))
        print(f"n={n} {name}: median {t:.3f} ms ({n ** 3 / 3 / (t * 1e-3) / 1e12:.2f} TFLOP/s) min {min(ts):.3f}", flush=True)
