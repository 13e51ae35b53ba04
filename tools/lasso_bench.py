"""BASELINE cfg-5: LassoSolver ADMM batch, A 2048x512 (+bias), K problems (default 4096).  Prints JSON."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import problems  # noqa: E402
from ipm_b200.LassoSolver import LassoSolver  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n, rows = 512, 2048
rs = np.random.RandomState(5)
A = rs.rand(rows, n)
nnz = int(n * K / 4)
x_true = np.zeros((n, K))
x_true[np.unravel_index(rs.randint(0, n * K, nnz), (n, K))] = rs.uniform(0, 50, nnz)
reg = 0.05 + 0.01 * rs.randn(K)
b = A @ x_true + rs.randn(rows, K)
out = {"K": K, "n": n + 1, "m": rows}
for name, kw in (("eps1e-6", dict(eps_abs=1e-6, eps_rel=1e-6, max_iters=5000)), ("defaults", dict())):
    s = LassoSolver(A, b, reg, rho=0.4, check_stop=10, add_bias=True, check_cvxpy=False, **kw)
    s.solve()  # warm
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = s.L.kernel_launches()
    e0.record()
    X, sol, _, its = s.solve()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out[name] = {"iters": its, "ms": ms, "ms_per_iter": ms / its, "solves_per_s": K / (ms * 1e-3),
                 "tflops": 2.0 * (n + 1) ** 2 * K * its / (ms * 1e-3) / 1e12, "launches": s.L.kernel_launches() - l0,
                 "obj0": float(sol[0])}
print(json.dumps(out))
