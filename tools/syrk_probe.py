"""Launch the Hessian kernel (H = C' diag(w) C, cfg-2 shape) a few times -- target of `ncu --set full`."""
import sys
import torch
sys.path.insert(0, ".")
from ipm_b200 import _abi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * n
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
_abi.require_device()
g = torch.Generator(device="cuda").manual_seed(0)
C_ = torch.rand((m, n), dtype=torch.float64, device="cuda", generator=g) * 4 - 2
w = torch.rand(m, dtype=torch.float64, device="cuda", generator=g) + 0.5
H = torch.zeros((n, n), dtype=torch.float64, device="cuda")
for _ in range(reps):
    _abi.call("ipm_gemm_tn_f64", C_.data_ptr(), n, C_.data_ptr(), n, w.data_ptr(), 1.0, 0.0, H.data_ptr(), n, n, n, m, 1, None)
torch.cuda.synchronize()
print("ok", float(H[0, 0]))
