"""Generate tests/golden/*.json by running the REAL reference (NumPy arm) from /root/reference.

Run once in the build container (the reference does not travel to the GPU box):

    python tests/golden/generate_golden.py

cvxpy / matplotlib are absent here and only used by the reference for its optional pre-check and
plots, so empty stub modules are injected; every solver is built with
``check_cvxpy=False``.  Each Newton class is instrumented (wrapping ``backtrack_search``) to record
the per-step sizes, which pins the line-search decision sequence (quirks Q1-Q4) and not only the
final optimum.  Inputs come from the seeded generators in ``tests/problems.py``; fixtures store the
generator call and the reference's outputs.
"""

import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("IPM_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "tests"))

for name in ("cvxpy", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, REF)

import problems  # noqa: E402
import NewtonSolver as _ns  # noqa: E402
import NewtonSolverInfeasibleStart as _nsi  # noqa: E402
from LPSolver import LPSolver  # noqa: E402
from QPSolver import QPSolver  # noqa: E402
from SOCPSolver import SOCPSolver  # noqa: E402
from LassoSolver import LassoSolver  # noqa: E402
import PhaseOne as _p1  # noqa: E402

STEPS = []


def _wrap(cls, infeasible):
    orig = cls.backtrack_search

    def wrapped(self, *a, **k):
        r = orig(self, *a, **k)
        step = r[0] if infeasible else r
        STEPS.append([bool(getattr(self, "phase1_flag", False)), float(step)])
        return r

    cls.backtrack_search = wrapped


_wrap(_ns.NewtonSolver, False)
_wrap(_nsi.NewtonSolverInfeasibleStart, True)


def _flt(x):
    return None if x is None else float(x)


def run_barrier(cls, gen, gen_kwargs, index, settings, name):
    prob = getattr(problems, gen)(**gen_kwargs)
    if isinstance(prob, list):
        prob = prob[index]
    STEPS.clear()
    np.random.seed(0)
    s = cls(**prob, check_cvxpy=False, suppress_print=True, **settings)
    ran_phase1 = hasattr(s, "phase1_solver") and s.phase1_solver.phase1_fm.s >= 1
    val = s.solve()
    rec = dict(
        name=name, solver=cls.__name__, generator=gen, generator_kwargs=gen_kwargs, index=index, settings=settings,
        value=float(val), inner_iters=[int(k) for k in s.inner_iters], outer_iters=int(s.outer_iters),
        phase1_inner_iters=[int(k) for k in s.phase1_solver.inner_iters] if ran_phase1 else None,
        xstar=[float(v) for v in np.asarray(s.xstar)], optimality_gap=float(s.optimality_gap),
        phase1_steps=[st for p, st in STEPS if p], main_steps=[st for p, st in STEPS if not p],
    )
    print(name, rec["value"], rec["inner_iters"], rec["phase1_inner_iters"])
    return rec


def group_lasso_socp():
    """demo.ipynb cells 26-28: SOCP form of the group lasso on example_data (27 vars, 8 cones)."""
    X = np.loadtxt(os.path.join(REF, "example_data/X_train.csv"), delimiter=",")
    X = np.hstack([np.ones(X.shape[0])[:, None], X])
    Y = np.loadtxt(os.path.join(REF, "example_data/Y_train.csv"), delimiter=",")
    groups = [[0], [1], [2], [3, 4, 5, 6, 7], [8, 9, 10, 11, 12, 13], [14, 15], [16], [17], [18]]
    w = np.sqrt([len(g) for g in groups])[1:]
    P = np.zeros((27, 27))
    P[:19, :19] = 1 / X.shape[0] * X.T @ X
    q = np.zeros(27)
    q[:19] = -1 / X.shape[0] * Y.T @ X
    q[19:] = 0.02 * w
    A, c = [], []
    for i in range(len(groups) - 1):
        Ai, ci = np.zeros((27, 27)), np.zeros(27)
        Ai[groups[i + 1], groups[i + 1]] = 1
        ci[i + 19] = 1
        A.append(Ai), c.append(ci)
    STEPS.clear()
    np.random.seed(0)
    s = SOCPSolver(P=P, q=q, A=[a.copy() for a in A], b=None, c=c, d=None, lower_bound=None, upper_bound=None,
                   check_cvxpy=False, suppress_print=True)
    x0 = np.array(s.x).copy()
    val = s.solve()
    rec = dict(name="socp_group_lasso", P=P.tolist(), q=q.tolist(), groups=groups, x0=x0.tolist(),
               value=float(val), offset=float(Y.T @ Y / (2 * X.shape[0])), fstar=49.9649387126726,
               inner_iters=[int(k) for k in s.inner_iters],
               phase1_inner_iters=[int(k) for k in s.phase1_solver.inner_iters],
               xstar=[float(v) for v in s.xstar], main_steps=[st for p, st in STEPS if not p],
               phase1_steps=[st for p, st in STEPS if p])
    print("socp_group_lasso", val, val + rec["offset"], rec["inner_iters"], rec["phase1_inner_iters"])
    return rec


def lasso_case(name, gen_kwargs, settings):
    prob = problems.lasso_testsolver(**gen_kwargs)
    s = LassoSolver(A=prob["A"].copy(), b=prob["b"], reg=prob["reg"], compute_loss=False, adaptive_rho=False,
                    use_gpu=False, check_cvxpy=False, **settings)
    X, sol, _, its = s.solve()
    rec = dict(name=name, generator="lasso_testsolver", generator_kwargs=gen_kwargs, settings=settings,
               solutions=[float(v) for v in sol], iterations=its if isinstance(its, list) else int(its),
               X_col0=[float(v) for v in X[:, 0]], X_frob=float(np.linalg.norm(X)),
               X_abs_sum=float(np.abs(X).sum()))
    print(name, rec["solutions"][:3], rec["iterations"])
    return rec


def standalone_phase_one():
    """Known-answer values of AutomatedTestsPhaseOne.py plus functional polytopes (:235-343)."""
    out = []
    # gradient / Hessian / objective KATs are restated analytically in tests/test_phase_one_kat.py;
    # here: functional runs of the real class
    cases = {
        "inside": (np.array([[1.0, 0], [0, 1], [-1, 0], [0, -1]]), np.array([1.0, 1, 1, 1]), np.array([0.0, 0.0])),
        "outside": (np.array([[1.0, 0], [0, 1], [-1, 0], [0, -1]]), np.array([1.0, 1, 1, 1]), np.array([5.0, 5.0])),
        "unbounded_set": (np.array([[1.0, 0], [0, 1]]), np.array([1.0, 1]), np.array([5.0, 5.0])),
        "empty_set": (np.array([[1.0, 0], [-1, 0]]), np.array([-1.0, -1]), np.array([0.0, 0.0])),
    }
    for nm, (G, h, x0) in cases.items():
        s = _p1.PhaseOneSolver(G, h, 15, x0=x0.copy())
        x, sv, warn = s.solve()
        out.append(dict(name=nm, G=G.tolist(), h=h.tolist(), x0=x0.tolist(), mu=15, x=[float(v) for v in x],
                        s=float(sv), warn=bool(warn)))
        print("phase_one", nm, x, sv, warn)
    np.random.seed(0)
    m, n = 200, 1000  # AutomatedTestsPhaseOne.py:325-343 shape; scaled down for fixture speed below
    m, n = 60, 40
    G = np.random.randn(m, n)
    xf = np.random.randn(n)
    h = G @ xf + np.random.rand(m)
    s = _p1.PhaseOneSolver(G, h, 15)
    x, sv, warn = s.solve()
    out.append(dict(name="random_60x40", seed=0, m=m, n=n, mu=15, s=float(sv), warn=bool(warn),
                    x=[float(v) for v in x], max_violation=float(np.max(G @ x - h))))
    print("phase_one random", sv, warn, np.max(G @ x - h))
    return out


def method_cases():
    """Newton-class dispatch (LPSolver.py:371-448, QPSolver.py:385-455): the reference's alternative
    ``linear_solve_method``s on the first test_LP / test_QP instance (both have equality constraints, so "kkt" is
    admissible).  The device engine maps all of them onto its Cholesky kernels; this pins that the outcome is the same."""
    out = []
    for method in ("np_solve", "np_lstsq", "direct", "kkt"):
        for cls, gen, settings, tag in ((LPSolver, "lp_testsolver", problems.LP_TEST_SETTINGS, "lp"),
                                        (QPSolver, "qp_testsolver", problems.QP_TEST_SETTINGS, "qp")):
            name = f"{tag}_seed1_n100_0_{method}"
            try:
                out.append(run_barrier(cls, gen, dict(seed=1, n=100, m=80, k=20, count=1), 0,
                                       dict(settings, linear_solve_method=method), name))
            except Exception as e:  # noqa: BLE001 -- some of the reference's own Newton classes crash on the NumPy arm
                print(name, "REFERENCE FAILED:", type(e).__name__, e)
                out.append(dict(name=name, solver=cls.__name__, reference_error=f"{type(e).__name__}: {e}"))
    with open(os.path.join(HERE, "method_cases.json"), "w") as f:
        json.dump(out, f, indent=1)


def dual_cases():
    """``get_dual_variables=True`` / ``track_loss=True`` outputs (LPSolver.py:608-609,641-646, QPSolver.py:593-594,
    626-631): lam_star = 1 / (t slacks(x*)) in the slack layout [C rows | upper bounds | lower bounds], v_star = v / t
    with the LAST t of the outer loop, objective_vals per accepted centering step."""
    out = []
    for cls, gen, gen_kwargs, settings, name in (
            (LPSolver, "lp_testsolver", dict(seed=1, n=100, m=80, k=20, count=1), problems.LP_TEST_SETTINGS,
             "lp_seed1_n100_0_duals"),
            (QPSolver, "qp_testsolver", dict(seed=1, n=100, m=80, k=20, count=1), problems.QP_TEST_SETTINGS,
             "qp_seed1_n100_0_duals"),
            (LPSolver, "lp_dense_family", dict(seed=0, n=64, warm=True), {}, "lp_dense_n64_warm_duals")):
        prob = getattr(problems, gen)(**gen_kwargs)
        if isinstance(prob, list):
            prob = prob[0]
        np.random.seed(0)
        s = cls(**prob, check_cvxpy=False, suppress_print=True, get_dual_variables=True, track_loss=True, **settings)
        val = s.solve()
        rec = dict(name=name, solver=cls.__name__, generator=gen, generator_kwargs=gen_kwargs, settings=settings,
                   value=float(val), objective_vals=[float(v) for v in s.objective_vals],
                   lam_star=[float(v) for v in np.asarray(s.lam_star).ravel()] if hasattr(s, "lam_star") else None,
                   v_star=[float(v) for v in np.asarray(s.v_star).ravel()] if hasattr(s, "v_star") else None)
        print(name, rec["value"], len(rec["objective_vals"]), None if rec["lam_star"] is None else len(rec["lam_star"]),
              None if rec["v_star"] is None else len(rec["v_star"]))
        out.append(rec)
    with open(os.path.join(HERE, "dual_cases.json"), "w") as f:
        json.dump(out, f, indent=1)


def behaviour_cases():
    """Control-flow behaviour of solve(): the phase-I failure on an empty feasible set (LPSolver.py:553-558) and a second
    solve() on the same object (quirk Q7: the iterate was updated in place, LPSolver.py:544, so the second run
    starts from the first one's end point)."""
    out = []
    prob = problems.lp_small_polytope(infeasible=True)
    try:
        LPSolver(**prob, check_cvxpy=False, suppress_print=True).solve()
        out.append(dict(name="lp_empty_set", error=None))
    except Exception as e:  # noqa: BLE001
        out.append(dict(name="lp_empty_set", error=type(e).__name__, message=str(e)))
    s = LPSolver(**problems.lp_small_polytope(), check_cvxpy=False, suppress_print=True)
    v1 = float(s.solve())
    it1 = [int(k) for k in s.inner_iters]
    v2 = float(s.solve())
    it2 = [int(k) for k in s.inner_iters]
    out.append(dict(name="lp_solve_twice", first=dict(value=v1, inner_iters=it1), second=dict(value=v2, inner_iters=it2)))
    print(out)
    with open(os.path.join(HERE, "behaviour_cases.json"), "w") as f:
        json.dump(out, f, indent=1)


def option_cases():
    """Constructor options outside the defaults: try_diag=False on a bounds-only LP (dense Cholesky of a diagonal Hessian,
    LPSolver.py:436-446), update_slacks_every > 0 (slacks refreshed inside the Armijo loop, NewtonSolver.py:196-202),
    use_psd_condition=True (SOCPSolver.py:20-54).  Same problems as barrier_cases.json."""
    with open(os.path.join(HERE, "barrier_cases.json")) as f:
        base = {c["name"]: c for c in json.load(f)}
    out = []
    for src, extra, cls in (("lp_bounds_only_n50", dict(try_diag=False), LPSolver),
                            ("lp_dense_n64_warm", dict(update_slacks_every=3), LPSolver),
                            ("lp_dense_n64_cold", dict(update_slacks_every=2), LPSolver),
                            ("socp_n48_warm", dict(use_psd_condition=True), SOCPSolver),
                            # update_slacks_every in the infeasible-start residual search
                            # (NewtonSolverInfeasibleStart.py:249-255) and with second-order cones
                            ("lp_seed1_n100_0", dict(update_slacks_every=2), LPSolver),
                            ("qp_seed1_n100_0", dict(update_slacks_every=3), QPSolver),
                            ("socp_n48_warm", dict(update_slacks_every=2), SOCPSolver),
                            ("socp_n48_cold", dict(update_slacks_every=3), SOCPSolver)):
        b = base[src]
        tag = "_".join(f"{k}_{v}" for k, v in extra.items())
        out.append(run_barrier(cls, b["generator"], b["generator_kwargs"], b["index"], dict(b["settings"], **extra),
                               f"{src}__{tag}"))
    with open(os.path.join(HERE, "option_cases.json"), "w") as f:
        json.dump(out, f, indent=1)


REF_POLYTOPES = {  # AutomatedTestsPhaseOne.py:253-318 (test_phase_one) and :364-367 (test_initialized_phase_one)
    "ref_inside": ([[1, 3], [1, 1], [-1, 0], [0, -1]], [9, 5, 0, 0], None),
    "ref_outside": ([[-1, -3], [-1, 1], [-1, 2], [1, 4]], [-6, 2, 2, 12], None),
    "ref_unbounded": ([[1, -2], [-3, 1]], [-2, 0], None),
    "ref_empty": ([[3, -1], [-1, 5], [-1, 0], [0, -1]], [-2, 1.5, 0, 0], None),
    "ref_initialized": ([[-1, -3], [-1, 1], [-1, 2], [1, 4]], [-6, 2, 2, 12], [-2, -3]),
}


def phase_one_kat():
    """The reference's OWN functional tests (AutomatedTestsPhaseOne.py:235-389) run through the real class: its four
    polytopes, the initialised start and the 200 x 1000 random problem (:325-343), with both linear solvers."""
    out = []
    for solver in ("solve", "cg"):
        for nm, (G, h, x0) in REF_POLYTOPES.items():
            G_, h_ = np.array(G), np.array(h)
            s = _p1.PhaseOneSolver(G_, h_, 15, x0=None if x0 is None else np.array(x0), linear_solver=solver)
            x, sv, warn = s.solve()
            out.append(dict(name=f"{nm}_{solver}", G=G, h=h, x0=x0, mu=15, linear_solver=solver,
                            x=[float(v) for v in x], s=float(sv), warn=bool(warn),
                            max_violation=float(np.max(G_ @ x - h_))))
            print("kat", nm, solver, x, sv, warn)
        np.random.seed(0)
        m, n = 200, 1000
        G = np.random.uniform(low=-10, high=10, size=(m, n))
        x = np.random.uniform(low=-5, high=5, size=(n))
        h = G @ x + 1
        s = _p1.PhaseOneSolver(G, h, 15, linear_solver=solver)
        x, sv, warn = s.solve()
        out.append(dict(name=f"ref_random_200x1000_{solver}", seed=0, m=m, n=n, mu=15, linear_solver=solver,
                        s=float(sv), warn=bool(warn), x=[float(v) for v in x], max_violation=float(np.max(G @ x - h))))
        print("kat random", solver, sv, warn, np.max(G @ x - h))
    with open(os.path.join(HERE, "phase_one_kat.json"), "w") as f:
        json.dump(out, f, indent=1)


def lasso_full_size():
    """BASELINE configs[4] at FULL size through the real reference: A 2048 x 512 (+ bias), K = 4096 problems, the settings of
    the reference's GPU arm (testSolver.py:1142-1159: eps 1e-6, max_iters 5000).  Same generator as bench.py
    (problems.lasso_cfg5).  Stores the iteration count and checksums of the solution."""
    A, b, reg = problems.lasso_cfg5(4096)
    kw = dict(rho=0.4, check_stop=10, add_bias=True, eps_abs=1e-6, eps_rel=1e-6, max_iters=5000)
    s = LassoSolver(A=A.copy(), b=b, reg=reg, compute_loss=False, adaptive_rho=False, use_gpu=False, check_cvxpy=False,
                    **kw)
    X, sol, _, its = s.solve()
    rec = dict(name="lasso_cfg5_full", K=4096, settings=kw, iterations=int(its), solutions_sum=float(np.sum(sol)),
               solutions_head=[float(v) for v in sol[:16]], X_frob=float(np.linalg.norm(X)),
               X_abs_sum=float(np.abs(X).sum()), X_nnz=int(np.count_nonzero(X)), X_col0=[float(v) for v in X[:, 0]])
    print("lasso_cfg5_full", its, rec["solutions_sum"], rec["X_frob"], rec["X_nnz"])
    with open(os.path.join(HERE, "lasso_full.json"), "w") as f:
        json.dump(rec, f, indent=1)


def cg_cases():
    """linear_solve_method="cg" (NewtonSolverCG, NewtonSolver.py:365-400: at most max_cg_iters = 50 conjugate-gradient
    steps per Newton system) on feasible-start LPs; phase-I stays on the Cholesky class (PhaseOneSolver.py:91-110)."""
    out = []
    for gen_kwargs, name in ((dict(seed=0, n=64, warm=True), "lp_dense_n64_warm_cg"),
                             (dict(seed=0, n=64), "lp_dense_n64_cold_cg"),
                             (dict(seed=0, n=256, warm=True), "lp_dense_n256_warm_cg")):
        out.append(run_barrier(LPSolver, "lp_dense_family", gen_kwargs, None, dict(linear_solve_method="cg"), name))
    with open(os.path.join(HERE, "cg_cases.json"), "w") as f:
        json.dump(out, f, indent=1)


def large_cases():
    """BASELINE cfg-2 family at n = 1024 and n = 2048 (cold and warm) and the cfg-3 family at n = 2048 (p = 512 equalities,
    k = 64 inequalities: the n = 8192, p = 2048, k = 20 shape of tests/test_fullsize_gpu.py scaled down 4x), constructor
    defaults -- the sizes at which the device engine switches to its multi-block kernels (look-ahead / tile-DAG Cholesky,
    persistent stream-K SYRK, multi-block triangular solves).  SURVEY Appendix A lists the n = 1024 values."""
    out = []
    for n in (1024, 2048):
        out.append(run_barrier(LPSolver, "lp_dense_family", dict(seed=0, n=n), None, {}, f"lp_dense_n{n}_cold"))
        out.append(run_barrier(LPSolver, "lp_dense_family", dict(seed=0, n=n, warm=True), None, {},
                               f"lp_dense_n{n}_warm"))
    out.append(run_barrier(QPSolver, "qp_dense_family", dict(seed=0, n=2048, p=512, k=64), None, {}, "qp_dense_n2048"))
    out.append(run_barrier(QPSolver, "qp_dense_family", dict(seed=1, n=1024, p=256, k=20), None, {}, "qp_dense_n1024"))
    with open(os.path.join(HERE, "large_cases.json"), "w") as f:
        json.dump(out, f, indent=1)


def main():
    if "--cg-only" in sys.argv:
        cg_cases()
        return
    if "--lasso-full-only" in sys.argv:
        lasso_full_size()
        return
    if "--kat-only" in sys.argv:
        phase_one_kat()
        return
    if "--large-only" in sys.argv:
        large_cases()
        return
    if "--options-only" in sys.argv:
        option_cases()
        return
    if "--behaviour-only" in sys.argv:
        behaviour_cases()
        return
    if "--methods-only" in sys.argv:
        method_cases()
        return
    if "--duals-only" in sys.argv:
        dual_cases()
        return
    barrier = []
    for i in range(3):
        barrier.append(run_barrier(LPSolver, "lp_testsolver", dict(seed=1, n=100, m=80, k=20, count=3), i,
                                   problems.LP_TEST_SETTINGS, f"lp_seed1_n100_{i}"))
    for i in range(2):
        barrier.append(run_barrier(QPSolver, "qp_testsolver", dict(seed=1, n=100, m=80, k=20, count=2), i,
                                   problems.QP_TEST_SETTINGS, f"qp_seed1_n100_{i}"))
    barrier.append(run_barrier(LPSolver, "lp_dense_family", dict(seed=0, n=64), None, {}, "lp_dense_n64_cold"))
    barrier.append(run_barrier(LPSolver, "lp_dense_family", dict(seed=0, n=64, warm=True), None, {},
                               "lp_dense_n64_warm"))
    barrier.append(run_barrier(LPSolver, "lp_dense_family", dict(seed=0, n=256), None, {}, "lp_dense_n256_cold"))
    barrier.append(run_barrier(LPSolver, "lp_dense_family", dict(seed=0, n=256, warm=True), None, {},
                               "lp_dense_n256_warm"))
    barrier.append(run_barrier(LPSolver, "lp_dense_family", dict(seed=7, n=97, m=150), None, {},
                               "lp_dense_n97_ragged"))
    barrier.append(run_barrier(QPSolver, "qp_dense_family", dict(seed=0, n=128, p=32, k=16), None, {},
                               "qp_dense_n128"))
    barrier.append(run_barrier(QPSolver, "qp_dense_family", dict(seed=0, n=512, p=128, k=64), None, {},
                               "qp_dense_n512"))
    barrier.append(run_barrier(LPSolver, "lp_bounds_only", dict(seed=3, n=50), None, {}, "lp_bounds_only_n50"))
    barrier.append(run_barrier(LPSolver, "lp_bounds_only", dict(seed=3, n=50, p=10), None, {},
                               "lp_bounds_eq_n50"))
    barrier.append(run_barrier(SOCPSolver, "socp_family", dict(seed=4, n=48, M=6, k=8), None,
                               problems.SOCP_TEST_SETTINGS, "socp_n48_warm"))
    barrier.append(run_barrier(SOCPSolver, "socp_family", dict(seed=4, n=48, M=6, k=8, p=6), None,
                               problems.SOCP_TEST_SETTINGS, "socp_n48_eq_warm"))
    barrier.append(run_barrier(SOCPSolver, "socp_family", dict(seed=4, n=48, M=6, k=8, warm=False), None,
                               problems.SOCP_TEST_SETTINGS, "socp_n48_cold"))
    barrier.append(run_barrier(SOCPSolver, "socp_family", dict(seed=11, n=96, M=12, k=16), None,
                               problems.SOCP_TEST_SETTINGS, "socp_n96_warm"))
    with open(os.path.join(HERE, "barrier_cases.json"), "w") as f:
        json.dump(barrier, f, indent=1)
    method_cases()
    with open(os.path.join(HERE, "socp_group_lasso.json"), "w") as f:
        json.dump(group_lasso_socp(), f, indent=1)
    lasso = [
        lasso_case("lasso_seed1_n100", dict(seed=1, n=100, m=80, num_problems=30), problems.LASSO_TEST_SETTINGS),
        lasso_case("lasso_seed1_n100_chunks3", dict(seed=1, n=100, m=80, num_problems=30),
                   dict(problems.LASSO_TEST_SETTINGS, num_chunks=3)),
        lasso_case("lasso_seed2_n64_defaults", dict(seed=2, n=64, m=40, num_problems=17),
                   dict(rho=0.4, max_iters=1000, check_stop=10, add_bias=True)),
        lasso_case("lasso_seed3_positive", dict(seed=3, n=32, m=30, num_problems=5),
                   dict(rho=0.4, max_iters=400, check_stop=10, add_bias=True, positive=True, normalize_A=True)),
    ]
    with open(os.path.join(HERE, "lasso_cases.json"), "w") as f:
        json.dump(lasso, f, indent=1)
    with open(os.path.join(HERE, "phase_one_standalone.json"), "w") as f:
        json.dump(standalone_phase_one(), f, indent=1)


if __name__ == "__main__":
    main()
