"""Where the CTAs of the tile-DAG Cholesky spend their time (needs a library built with -DIPM_DAG_TIMING, passed through
IPM_B200_LIB).  Prints, per counter, mean / min / max over the CTAs as a share of the kernel's duration."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ipm_b200 import _abi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
_abi.require_device()
L = _abi.lib()
g = torch.Generator(device="cuda").manual_seed(1)
C_ = torch.rand((2 * n, n), dtype=torch.float64, device="cuda", generator=g) * 4 - 2
H = C_.T @ C_
H.diagonal().add_(1e-3)
work = torch.empty_like(H)
info = torch.zeros(1, dtype=torch.int32, device="cuda")
buf = (C.c_longlong * (256 * 16))()
names = ["kernel", "contraction", "wait diag", "potf2", "tile solve", "publish", "tasks", "poll (warp 0)",
         "solve: load U", "solve: load P", "solve: subst", "solve: DMMA", "solve: store", "subtract_acc"]
for rep in range(3):
    work.copy_(H)
    torch.cuda.synchronize()
    L.ipm_internal_dag_timing(buf, 256 * 16)  # clear
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _abi.call("ipm_potrf_upper_dag_f64", work.data_ptr(), n, n, info.data_ptr(), None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    L.ipm_internal_dag_timing(buf, 256 * 16)
t = np.array(buf[:], dtype=np.float64).reshape(256, 16)
t = t[t[:, 0] > 0]
kern = t[:, 0].max()
print(f"n={n}: {ms:.3f} ms, {len(t)} CTAs, longest CTA {kern:.3e} cycles -> {kern / ms / 1e3:.0f} MHz")
for k, name in enumerate(names):
    col = t[:, k]
    if name == "tasks":
        print(f"  {name:14s} mean {col.mean():.1f} min {col.min():.0f} max {col.max():.0f}")
    else:
        print(f"  {name:14s} mean {100 * col.mean() / kern:5.1f}%  min {100 * col.min() / kern:5.1f}%  max {100 * col.max() / kern:5.1f}%"
              f"   ({col.mean() / kern * ms * 1e3:.0f} us per CTA)")
# the CTAs that own the last diagonal tasks show the chain
order = np.argsort(-t[:, 2])[:5]
print("  CTAs with the longest diag waits:", [(int(i), round(float(t[i, 2] / kern * ms), 3)) for i in order])
