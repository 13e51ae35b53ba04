"""One cfg-4 SOCP solve (n = 16384, 256 cones of 64 rows, warm start) for launch-list captures.  Tools only.
    ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file out.csv python tools/socp_probe.py"""
import sys
import time

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import problems  # noqa: E402
from ipm_b200.SOCPSolver import SOCPSolver  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
prob = problems.socp_family(seed=4, n=n, M=256, k=64)
x0 = prob["x0"].copy()
s = SOCPSolver(**{k: v for k, v in prob.items() if k != "x0"}, x0=x0, check_cvxpy=False, suppress_print=True,
               **problems.SOCP_TEST_SETTINGS)
torch.cuda.synchronize()
t0 = time.perf_counter()
val = s.solve()
torch.cuda.synchronize()
print("value", val, "newton steps", sum(s.inner_iters), "seconds", time.perf_counter() - t0)
if len(sys.argv) > 2:  # a second solve of the same object: what of the first one was one-time cost?
    s.x_dev.copy_(torch.as_tensor(x0).cuda())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    val = s.solve()
    torch.cuda.synchronize()
    print("second solve: value", val, "newton steps", sum(s.inner_iters), "seconds", time.perf_counter() - t0)
