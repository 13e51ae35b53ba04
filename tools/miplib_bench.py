"""Sparse-aware vs dense LP path on the aflow40b-shaped stand-in (SURVEY.md 8(d) cfg 1 / 8(f)-1).  Prints JSON."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import problems  # noqa: E402
from ipm_b200 import miplib  # noqa: E402
from ipm_b200.LPSolver import LPSolver  # noqa: E402

prob = miplib.synthetic_network_lp()
out = {"workload": "synthetic stand-in for MIPLIB aflow40b (n=2728, 78 equalities, 1364 inequalities, 0.17% non-zeros), "
                   "test_LP_sparse settings (testSolver.py:338-356)"}
for name, sparse in (("sparse", "auto"), ("dense", False)):
    s = LPSolver(**prob, check_cvxpy=False, suppress_print=True, sparse=sparse, **problems.LP_TEST_SETTINGS)
    s.solve()  # warm-up (allocator, attribute calls); phase-I is skipped on the re-solve as in the reference (Q7)
    s2 = LPSolver(**prob, check_cvxpy=False, suppress_print=True, sparse=sparse, **problems.LP_TEST_SETTINGS)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    val = s2.solve()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    steps = sum(s2.inner_iters) + sum(s2.phase1_solver.inner_iters)
    out[name] = {"time_to_solve_s": dt, "value": val, "newton_steps": steps, "ms_per_newton_step": 1e3 * dt / steps,
                 "hessian_entries": s2.data.sparse.nout if s2.data.sparse is not None else None}
print(json.dumps(out))
