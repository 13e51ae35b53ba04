"""How stable are the reference's OWN outputs on the ``update_slacks_every`` and dual-variable goldens?

The three device-side comparisons that failed in round 1 (``lp_dense_n64_warm__update_slacks_every_3``,
``lp_dense_n64_cold__update_slacks_every_2``, ``lp_dense_n64_warm_duals``) compare quantities that the reference itself does
not reproduce under a 1e-14 RELATIVE perturbation of its input matrix C (i.e. below one unit in the last place of most
entries): with ``update_slacks_every > 0`` the Armijo test mixes a refreshed barrier term with a lagged objective term and
accepts or rejects steps on differences at rounding level, and ``lam_star = 1 / (t s)`` at t ~ 1e13 divides by slacks of
1e-13 that are themselves differences of O(1) numbers.

This script runs the CPU oracle (which reproduces the real reference step for step, tests/test_oracle_golden.py) on the
unperturbed problem and on SEEDS perturbed copies and writes, per golden case, the per-centering-step min / max Newton
count and the spread of ``lam_star`` into ``tests/golden/sensitivity.json``.  ``tests/test_solvers_gpu.py`` uses the
envelope (widened by the usual +-2) as the bar for those cases: a device count outside it is a real failure.

    python tests/golden/sensitivity_options.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import problems  # noqa: E402
from oracle import OracleLP, OracleQP, OracleSOCP  # noqa: E402

SEEDS = 12
REL = 1e-14
CLS = {"LPSolver": OracleLP, "QPSolver": OracleQP, "SOCPSolver": OracleSOCP}


def perturbed(prob, seed):
    p = dict(prob)
    if seed is not None:
        rs = np.random.RandomState(1000 + seed)
        if isinstance(p.get("A"), list):  # SOCP: the cone matrices
            p["A"] = [Ai * (1 + REL * rs.randn(*Ai.shape)) for Ai in p["A"]]
        else:
            p["C"] = p["C"] * (1 + REL * rs.randn(*p["C"].shape))
    for k in ("x0",):
        if k in p:
            p[k] = p[k].copy()
    return p


def run(case, seed, duals=False):
    prob = getattr(problems, case["generator"])(**case["generator_kwargs"])
    if isinstance(prob, list):
        prob = prob[case.get("index") or 0]
    np.random.seed(0)  # the default x0 of an unbounded SOCP is np.random.rand (SOCPSolver.py:166)
    s = CLS[case["solver"]](**perturbed(prob, seed), **case["settings"])
    val = s.solve()
    out = dict(value=float(val), inner_iters=list(s.inner_iters),
               phase1_inner_iters=None if s.phase1 is None or not s.phase1.inner_iters else list(s.phase1.inner_iters))
    if duals:
        lam, nu = s.dual_variables()
        out["lam"], out["nu"] = lam, nu
    return out


def envelope(runs, key):
    rows = [r[key] for r in runs if r[key] is not None]
    if not rows:
        return None
    assert len({len(r) for r in rows}) == 1, rows
    a = np.array(rows)
    return dict(min=a.min(axis=0).tolist(), max=a.max(axis=0).tolist())


LARGE_QP = {"qp_dense_n1024": 8, "qp_dense_n2048": 4}  # case -> perturbed runs (minutes of CPU each at n = 2048)


def large_qp(out):
    """The equality-constrained large cases: their late centering steps stop on a residual norm that sits at the rounding
    noise of t * grad f0 + A'v (tests/test_solvers_gpu.py: noise_dominated_steps), and the reference's own counts move by
    up to 4 there.  `--large` recomputes these entries; without it the previous ones are kept."""
    with open(os.path.join(HERE, "large_cases.json")) as f:
        cases = {c["name"]: c for c in json.load(f)}
    for name, seeds in LARGE_QP.items():
        case = cases[name]
        runs = [run(case, None)] + [run(case, sd) for sd in range(seeds)]
        assert runs[0]["inner_iters"] == case["inner_iters"], "oracle no longer reproduces the golden"
        out["cases"][name] = dict(inner_iters=envelope(runs, "inner_iters"),
                                  phase1_inner_iters=envelope(runs, "phase1_inner_iters"), perturbed_runs=seeds,
                                  value_spread=float(max(abs(r["value"] - runs[0]["value"]) for r in runs)))
        print(name, out["cases"][name])


# barrier_cases.json entries with a centering step whose count the reference itself does not reproduce under the
# perturbation (qp_seed1_n100_0: the step after a 21-iteration centering moves over 9..11); the INT8 tensor-core Hessian
# (different rounding than the FP64 kernel) is held to this envelope there
SENSITIVE_BARRIER = ["qp_seed1_n100_0"]


def barrier_cases(out):
    with open(os.path.join(HERE, "barrier_cases.json")) as f:
        cases = {c["name"]: c for c in json.load(f)}
    for name in SENSITIVE_BARRIER:
        case = cases[name]
        runs = [run(case, None)] + [run(case, sd) for sd in range(SEEDS)]
        assert runs[0]["inner_iters"] == case["inner_iters"], "oracle no longer reproduces the golden"
        out["cases"][name] = dict(inner_iters=envelope(runs, "inner_iters"),
                                  phase1_inner_iters=envelope(runs, "phase1_inner_iters"),
                                  value_spread=float(max(abs(r["value"] - runs[0]["value"]) for r in runs)))
        print(name, out["cases"][name])


def main():
    out = {"relative_perturbation": REL, "perturbed_runs": SEEDS, "cases": {}}
    path = os.path.join(HERE, "sensitivity.json")
    if "--barrier-only" in sys.argv:  # keep everything else as it is
        with open(path) as f:
            out = json.load(f)
        barrier_cases(out)
        with open(path, "w") as f:
            json.dump(out, f, indent=1)
        return
    if "--large" in sys.argv:
        large_qp(out)
    elif os.path.exists(path):
        with open(path) as f:
            old = json.load(f)["cases"]
        out["cases"].update({k: v for k, v in old.items() if k in LARGE_QP})
    with open(os.path.join(HERE, "option_cases.json")) as f:
        options = [c for c in json.load(f) if "update_slacks_every" in c["settings"]]
    with open(os.path.join(HERE, "cg_cases.json")) as f:  # NewtonSolverCG: 50 CG steps on Hessians with cond > 1e12
        options += json.load(f)
    for case in options:
        runs = [run(case, None)] + [run(case, sd) for sd in range(SEEDS)]
        assert runs[0]["inner_iters"] == case["inner_iters"], "oracle no longer reproduces the golden"
        out["cases"][case["name"]] = dict(inner_iters=envelope(runs, "inner_iters"),
                                          phase1_inner_iters=envelope(runs, "phase1_inner_iters"),
                                          value_spread=float(max(abs(r["value"] - runs[0]["value"]) for r in runs)))
        print(case["name"], out["cases"][case["name"]])
    with open(os.path.join(HERE, "dual_cases.json")) as f:
        duals = json.load(f)
    for case in duals:
        runs = [run(case, None, True)] + [run(case, sd, True) for sd in range(SEEDS)]
        lam0 = runs[0]["lam"]
        np.testing.assert_allclose(lam0, case["lam_star"], rtol=1e-9)
        spread = max(float(np.linalg.norm(r["lam"] - lam0) / np.linalg.norm(lam0)) for r in runs[1:])
        rec = dict(lam_rel_spread=spread, inner_iters=envelope(runs, "inner_iters"))
        if runs[0]["nu"] is not None:
            nu0 = runs[0]["nu"]
            rec["nu_rel_spread"] = max(float(np.linalg.norm(r["nu"] - nu0) / (1e-300 + np.linalg.norm(nu0)))
                                       for r in runs[1:])
        out["cases"][case["name"]] = rec
        print(case["name"], rec)
    barrier_cases(out)
    with open(os.path.join(HERE, "sensitivity.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
