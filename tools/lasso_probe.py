"""A few launches of the persistent multi-iteration Lasso kernel at the cfg-5 shape -- target of `ncu --set full`.
    python tools/lasso_probe.py [K] [launches]"""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import problems  # noqa: E402
from ipm_b200.LassoSolver import LassoSolver  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 3
A, b, reg = problems.lasso_cfg5(4096)
b, reg = b[:, ::4096 // K], reg[::4096 // K]
s = LassoSolver(A, b, reg, rho=0.4, check_stop=10, add_bias=True, check_cvxpy=False, eps_abs=1e-6, eps_rel=1e-6,
                max_iters=10 * launches)
_, sol, _, its = s.solve()
torch.cuda.synchronize()
print("iterations", its, "objective sum", float(sol.sum()))
