"""One centering step of the BASELINE cfg-3 QP (n = 8192, 2048 equalities, 20 inequalities) -- target of an ncu launch
list for the infeasible-start Newton path (TRSM with 2048 right-hand sides, Schur complement, two factorisations)."""
import sys
import time

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import problems  # noqa: E402
from ipm_b200.QPSolver import QPSolver  # noqa: E402


def gram(Pp):
    t = torch.as_tensor(Pp).to("cuda")
    return (t.T @ t).cpu().numpy()


prob = problems.qp_dense_family(seed=3, n=8192, p=2048, k=20, gram=gram, with_feasible_point=True)
x_feas = prob.pop("x_feas")
kw = dict(problems.QP_TEST_SETTINGS)
kw["max_outer_iters"] = int(sys.argv[1]) if len(sys.argv) > 1 else 1
s = QPSolver(**prob, x0=x_feas, check_cvxpy=False, suppress_print=True, **kw)
torch.cuda.synchronize()
t0 = time.perf_counter()
val = s.solve()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("value", val, "newton steps", s.inner_iters, "phase-I", s.phase1_solver.inner_iters if s.phase1_solver else None,
      "time %.3f s" % dt, "ms per Newton step %.2f" % (1e3 * dt / max(1, sum(s.inner_iters))))
