"""Same-box A/B of the Hessian kernel (C' diag(w) C, n = 8192, m = 16384): checks the result against torch and prints
the median of 7 L2-flushed launches.  Run once per variant (IPM_GEMM_WARPS16=0/1, or IPM_B200_LIB=...)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ipm_b200 import _abi  # noqa: E402

n, m = 8192, 16384
_abi.require_device()
g = torch.Generator(device="cuda").manual_seed(0)
C_ = torch.rand((m, n), dtype=torch.float64, device="cuda", generator=g) * 4 - 2
w = torch.rand(m, dtype=torch.float64, device="cuda", generator=g) + 0.5
H = torch.zeros((n, n), dtype=torch.float64, device="cuda")
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")


def syrk():
    _abi.call("ipm_gemm_tn_f64", C_.data_ptr(), n, C_.data_ptr(), n, w.data_ptr(), 1.0, 0.0, H.data_ptr(), n, n, n, m, 1, None)


syrk()
torch.cuda.synchronize()
ref = (C_[:, :512].T * w) @ C_
err = float(((torch.triu(H[:512]) - torch.triu(ref)).abs().max() / ref.abs().max()).item())
ref2 = (C_[:, -512:].T * w) @ C_[:, -512:]
err2 = float(((torch.triu(H[-512:, -512:]) - torch.triu(ref2)).abs().max() / ref2.abs().max()).item())
ts = []
for _ in range(7):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    syrk()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = float(np.median(ts))
print(f"warps16={os.environ.get('IPM_GEMM_WARPS16', '0')} median {t:.3f} ms min {min(ts):.3f} ms "
      f"{m * n * (n + 1) / (t * 1e-3) / 1e12:.2f} TFLOP/s  rel.err {err:.1e} {err2:.1e}")
assert err < 1e-12 and err2 < 1e-12
