#!/usr/bin/env python
"""Step-by-step comparison of a device solve with the CPU oracle on one golden case (GPU box; development aid).

For every Newton iteration both sides record (step size, Newton decrement / residual norm); the tool prints, per
solver phase, how many leading iterations agree exactly in step size and to 1e-6 in the decrement, and the first
iteration where they part.  A divergence after many identical iterations at a point where the Armijo test compares
barrier values at rounding level is noise; a divergence in the first iterations is a bug.

    python tools/debug_parity.py option_cases.json lp_dense_n64_warm__update_slacks_every_3 [more names ...]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import problems  # noqa: E402
from oracle import OracleLP, OracleQP, OracleSOCP  # noqa: E402


def compare(name, dev, ref):
    n = min(len(dev), len(ref))
    same = 0
    for k in range(n):
        a, b = dev[k], ref[k]
        if a[0] != b[0]:
            break
        if a[1] is not None and b[1] is not None and abs(a[1] - b[1]) > 1e-6 * max(1.0, abs(b[1])):
            break
        same += 1
    print(f"  {name}: device {len(dev)} Newton steps, oracle {len(ref)}; first {same} agree")
    for k in range(max(0, same - 2), min(n, same + 4)):
        print(f"    it {k:3d}  device step {dev[k][0]:.6g} dec {dev[k][1]}   |  oracle step {ref[k][0]:.6g} dec {ref[k][1]}")


def main():
    fname, names = sys.argv[1], sys.argv[2:]
    with open(os.path.join(ROOT, "tests", "golden", fname)) as f:
        cases = {c["name"]: c for c in json.load(f)}
    from ipm_b200.LPSolver import LPSolver
    from ipm_b200.QPSolver import QPSolver
    from ipm_b200.SOCPSolver import SOCPSolver

    dev_cls = {"LPSolver": LPSolver, "QPSolver": QPSolver, "SOCPSolver": SOCPSolver}
    ora_cls = {"LPSolver": OracleLP, "QPSolver": OracleQP, "SOCPSolver": OracleSOCP}
    for nm in names:
        case = cases[nm]
        prob = getattr(problems, case["generator"])(**case["generator_kwargs"])
        if isinstance(prob, list):
            prob = prob[case.get("index") or 0]
        print(nm, case["settings"])
        np.random.seed(0)
        s = dev_cls[case["solver"]](**prob, check_cvxpy=False, suppress_print=True, **case["settings"])
        s.ns.trace = []
        if s.phase1_solver is not None:
            s.phase1_solver.ns.trace = []
        val = s.solve()
        trace = {}
        prob = getattr(problems, case["generator"])(**case["generator_kwargs"])
        if isinstance(prob, list):
            prob = prob[case.get("index") or 0]
        np.random.seed(0)
        o = ora_cls[case["solver"]](**prob, **case["settings"], trace=trace)
        oval = o.solve()
        print(f"  value device {val!r} oracle {oval!r} golden {case['value']!r}")
        print(f"  inner_iters device {s.inner_iters}\n              oracle {o.inner_iters}")
        if s.phase1_solver is not None and s.phase1_solver.inner_iters:
            print(f"  phase-I device {s.phase1_solver.inner_iters} oracle {o.phase1.inner_iters}")
            compare("phase-I", s.phase1_solver.ns.trace, trace.get("phase1", []))
        compare("main", s.ns.trace, trace.get("main", []))
        if "lam_star" in case:
            s2 = dev_cls[case["solver"]](**prob, check_cvxpy=False, suppress_print=True, get_dual_variables=True,
                                         track_loss=True, **case["settings"])
            s2.solve()
            lam, ref = np.asarray(s2.lam_star).ravel(), np.array(case["lam_star"])
            print("  lam rel diff", np.linalg.norm(lam - ref) / np.linalg.norm(ref), "t_final", s2.t_final,
                  "objective_vals max rel diff",
                  np.max(np.abs(np.asarray(s2.objective_vals) - case["objective_vals"]) / np.abs(case["objective_vals"])))


if __name__ == "__main__":
    main()
