"""BASELINE cfg-5: LassoSolver ADMM batch, A 2048x512 (+bias), K problems -- persistent multi-iteration launches
(csrc/lasso_multi.cu) against one launch per iteration (csrc/lasso.cu), same box, same inputs.

    python tools/lasso_bench.py [K ...]        (default 4096 2048 1024 512: the per-GPU shards of a 1/2/4/8-way split)
Prints one JSON line per K."""
import json
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import problems  # noqa: E402
from ipm_b200.LassoSolver import LassoSolver  # noqa: E402

PEAK = 37.1
n = 512
for K in [int(a) for a in sys.argv[1:]] or [4096, 2048, 1024, 512]:
    A, b, reg = problems.lasso_cfg5(4096)
    b, reg = b[:, ::4096 // K], reg[::4096 // K]  # strided columns, like the per-rank split of bench.py
    out = {"K": K, "n": n + 1}
    kw = dict(rho=0.4, check_stop=10, add_bias=True, check_cvxpy=False, eps_abs=1e-6, eps_rel=1e-6, max_iters=5000)
    ref = None
    for name, multi in (("multi_iteration", True), ("one_launch_per_iteration", False)):
        s = LassoSolver(A, b, reg, **kw)
        s.multi_iteration = multi
        s.solve()  # warm
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = s.L.kernel_launches()
            e0.record()
            X, sol, _, its = s.solve()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[1]
        tf = 2.0 * (n + 1) ** 2 * K * its / (ms * 1e-3) / 1e12
        out[name] = {"iters": its, "ms": ms, "us_per_iter": 1e3 * ms / its, "solves_per_s": K / (ms * 1e-3), "tflops": tf,
                     "frac_fp64_peak": tf / PEAK, "launches": s.L.kernel_launches() - l0, "obj_sum": float(sol.sum())}
        if ref is None:
            ref = (its, float(sol.sum()))
        else:
            out["same_iterations"] = its == ref[0]
            out["obj_sum_rel_diff"] = abs(float(sol.sum()) - ref[1]) / abs(ref[1])
    print(json.dumps(out), flush=True)
