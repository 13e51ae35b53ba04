"""Run one potrf + trsv pair at size n (for ncu launch lists)."""
import sys
import torch
sys.path.insert(0, ".")
from ipm_b200 import _abi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
_abi.require_device()
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.rand((n, n), dtype=torch.float64, device="cuda", generator=g)
H = A @ A.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
info = torch.zeros(1, dtype=torch.int32, device="cuda")
b = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
tws = torch.zeros(n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
_abi.call("ipm_potrf_upper_f64", H.data_ptr(), n, n, info.data_ptr(), None)
_abi.call("ipm_trsv_upper_f64", H.data_ptr(), n, n, b.data_ptr(), 1, tws.data_ptr(), None)
_abi.call("ipm_trsv_upper_f64", H.data_ptr(), n, n, b.data_ptr(), 0, tws.data_ptr(), None)
torch.cuda.synchronize()
print("info", int(info.item()))
