#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 interior-point engine (contract: see DESIGN.md, "Measurement").

Workload (BASELINE.json configs[1]): dense random LP, n = 8192 variables, m = 16384 inequality rows plus box
bounds, FP64, constructor defaults of the reference's ``LPSolver``; warm start (x0 strictly feasible, so the
timed region is the main barrier phase).  One bench "step" = one complete ``LPSolver.solve()``.

  metric `newton_steps_per_s` = Newton iterations performed / device time, summed over all ranks.
  value : problem data already resident in HBM (solver constructed before the timed region; what the reference
          times, testSolver.py:151-153).
  e2e   : through the public API from HOST NumPy buffers: constructor (H2D of C, d, c, bounds) + solve() + the
          device->host read of the solution, all inside the timed region.
  roofline : the Hessian kernel  H = C' diag(w) C  (ipm_gemm_tn_f64, FP64 DMMA): algorithmic m*n*(n+1) flop per
          launch / mean launch time from CUDA events recorded around every such launch inside the timed region.
  cpu_baseline : the CPU oracle (NumPy/SciPy restatement of the reference, oracle/) on this box's host cores on
          a bounded sample (a fixed number of full-size Newton steps).

N > 1 (torchrun): every rank solves its own copy of the same LP instance with no data-path collective ("weak"
scaling; BASELINE north_star: batches of independent LP instances split across GPUs with no communication).  The
copies are identical on purpose: instances drawn from different seeds need different numbers of Newton steps
(76 .. 130 at this size), and the max-over-ranks time would then measure that imbalance instead of the hardware.

`--impl reference` times the reference algorithm on the host CPU (oracle port; the reference itself is pure
Python/NumPy and does not travel to the GPU box) for the same metric.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FP64_TENSOR_PEAK_TFLOPS = 37.1  # measured DMMA.8x8x4 issue rate on this pool's B200 (profiles/fp64_peak_r01.json)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--m", type=int, default=None)
    ap.add_argument("--cpu-newton-steps", type=int, default=2, help="full-size Newton steps in the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--shard", default="instances", choices=["instances", "rows"],
                    help="N > 1: independent LP instance per GPU (weak, no collective) or ONE LP row-sharded with an "
                         "NCCL all-reduce of the partial Hessian (strong)")
    ap.add_argument("--lasso-k", type=int, default=4096, help="Lasso batch size (0 disables the Lasso section)")
    ap.add_argument("--workload", default="lp", choices=["lp", "socp"],
                    help="lp: BASELINE configs[1] (the headline); socp: configs[3] (n=16384, 256 cones of 64 rows, "
                         "test_SOCP settings) -- use with --shard rows for the cone-sharded Hessian scaling numbers")
    ap.add_argument("--cones", type=int, default=256)
    ap.add_argument("--factorisation", action="store_true",
                    help="N = 1: also time the n x n Cholesky alone, stream-ordered and single-launch tile-DAG, and add a "
                         "'factorisation' object (roofline against the FP64 tensor peak) to the JSON line")
    return ap.parse_args()


def factorisation_section(n, reps=5):
    """ipm_potrf_upper_f64 (stream-ordered) and ipm_potrf_upper_dag_f64 (one persistent launch) on an SPD matrix with
    the Hessian's structure (C' diag(w) C + I), L2 flushed before every factorisation; n^3 / 3 flop."""
    import torch

    from ipm_b200 import _abi

    g = torch.Generator(device="cuda").manual_seed(n)
    C_ = torch.rand((2 * n, n), dtype=torch.float64, device="cuda", generator=g) * 4 - 2
    w = torch.rand(2 * n, dtype=torch.float64, device="cuda", generator=g) + 0.5
    H = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    _abi.call("ipm_gemm_tn_f64", C_.data_ptr(), n, C_.data_ptr(), n, w.data_ptr(), 1.0, 0.0, H.data_ptr(), n, n, n,
              2 * n, 1, None)
    H.diagonal().add_(1.0)
    del C_
    work = torch.empty_like(H)
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    out = {"n": n, "flop": n ** 3 / 3.0, "l2": "256 MB written between repetitions"}
    import ctypes as C

    lib = _abi.lib()
    for nm in ("ipm_internal_potrf_stream_f64", "ipm_internal_potrf_dag1_f64"):  # library-internal A/B entry points
        getattr(lib, nm).restype, getattr(lib, nm).argtypes = C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                                                         C.c_void_p]
    for key, name in (("default", "ipm_potrf_upper_f64"), ("stream_ordered", "ipm_internal_potrf_stream_f64"),
                      ("tile_dag_per_tile_deps", "ipm_internal_potrf_dag1_f64"),
                      ("tile_dag_pipelined", "ipm_potrf_upper_dag_f64")):
        ts = []
        for _ in range(reps + 1):
            work.copy_(H)
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            _abi.check(getattr(lib, name)(work.data_ptr(), n, n, info.data_ptr(), None), name)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        assert int(info.item()) == 0
        ms = float(np.median(ts[1:]))
        tf = out["flop"] / (ms * 1e-3) / 1e12
        out[key] = {"ms": ms, "achieved": tf, "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s",
                    "frac": tf / FP64_TENSOR_PEAK_TFLOPS, "entry_point": name}
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self._stop = index, [], threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


def hessian_dram_traffic(n, m):
    """DRAM bytes (read + write) per launch of the Hessian kernel from the committed `ncu --set full` capture
    (profiles/syrk_hessian_ncu_r01d.csv, taken at the default cfg-2 shape); None for any other shape."""
    if (n, m) != (8192, 16384):
        return None
    try:
        import csv

        rows = list(csv.reader(open(os.path.join(ROOT, "profiles", "syrk_hessian_ncu_r01d.csv"))))
        h, units, first = rows[0], rows[1], rows[2]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        rd = float(first[h.index("dram__bytes_read.sum")]) * scale[units[h.index("dram__bytes_read.sum")]]
        wr = float(first[h.index("dram__bytes_write.sum")]) * scale[units[h.index("dram__bytes_write.sum")]]
        return rd + wr
    except Exception:
        return None


def workload(args, rank):
    import problems

    if args.workload == "socp":
        n = 16384 if args.n == 8192 else args.n  # --n default is the LP's
        args.n = n
        return problems.socp_family(seed=4, n=n, M=args.cones, k=64), args.cones * 66
    m = 2 * args.n if args.m is None else args.m
    prob = problems.lp_dense_family(seed=8192, n=args.n, m=m, warm=True)
    return prob, m


def cpu_sample(prob, newton_steps):
    """Bounded CPU sample: `newton_steps` full-size Newton iterations of the oracle (first centering step)."""
    from oracle import OracleLP

    p = dict(prob)
    p["x0"] = prob["x0"].copy()
    o = OracleLP(**p, max_outer_iters=1, max_inner_iters=newton_steps)
    t0 = time.perf_counter()
    o.solve()
    dt = time.perf_counter() - t0
    return sum(o.inner_iters), dt


def blas_info():
    """BLAS vendor / thread count behind NumPy on this host (report them with the CPU baseline)."""
    try:
        from threadpoolctl import threadpool_info

        libs = [f"{i.get('internal_api', '?')} {i.get('version', '')} x{i.get('num_threads', '?')}"
                for i in threadpool_info() if i.get("user_api") == "blas"]
        return "; ".join(libs) or "unknown BLAS"
    except Exception:
        return "unknown BLAS"


def blas_threads():
    """Threads the NumPy BLAS actually uses (the `cores` of the CPU baseline); os.cpu_count() if unknown."""
    try:
        from threadpoolctl import threadpool_info

        n = [i.get("num_threads") for i in threadpool_info() if i.get("user_api") == "blas"]
        return int(max(n)) if n else os.cpu_count()
    except Exception:
        return os.cpu_count()


def lasso_workload(K):
    """BASELINE configs[4]: A 2048x512 (+bias), K problems, generator of testSolver.py:1096-1104 (seed 5)."""
    import problems

    return problems.lasso_cfg5(K)


def lasso_section(K, rank, world):
    """Second headline metric (Lasso solves/s): the batch is split by strided columns across ranks (the
    reference's num_chunks semantics), no communication.  Settings of the reference's GPU arm
    (testSolver.py:1142-1159: eps 1e-6, max_iters 5000)."""
    import torch

    from ipm_b200 import dist as D
    from ipm_b200.LassoSolver import LassoSolver

    A, b, reg = lasso_workload(K)
    cols = D.strided_columns(K, rank, world)
    kw = dict(rho=0.4, check_stop=10, add_bias=True, check_cvxpy=False, eps_abs=1e-6, eps_rel=1e-6, max_iters=5000)
    s = LassoSolver(A, b[:, cols], reg[cols], **kw)
    s.solve()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = s.L.kernel_launches()
    e0.record()
    _, sol, _, its = s.solve()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = s.L.kernel_launches() - l0
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    s2 = LassoSolver(A, b[:, cols], reg[cols], **kw)
    X2, _, _, _ = s2.solve()
    f1.record()
    torch.cuda.synchronize()
    ms_e2e = f0.elapsed_time(f1)
    n1 = A.shape[1] + 1
    return dict(ms=ms, ms_e2e=ms_e2e, iters=int(its), k_local=len(cols), launches=launches,
                flop=2.0 * n1 * n1 * len(cols) * its, h2d=s2.h2d_bytes, d2h=int(X2.nbytes))


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank, which would time the CPU arm on ONE core.  The CPU arm runs on
    rank 0 alone (the other ranks exit), so it may -- and per the contract must -- use every core of the host."""
    try:
        import scipy.linalg  # noqa: F401  (SciPy ships its own OpenBLAS; it must be loaded before the limit is raised)
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=len(os.sched_getaffinity(0)), user_api="blas")
    except Exception:
        pass


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    prob, m = workload(args, 0)
    times, steps = [], 0
    for i in range(args.warmup + args.steps):
        k, dt = cpu_sample(prob, 1)
        if i >= args.warmup:
            times.append(dt)
            steps += k
    total = sum(times)
    val = steps / total
    cores = blas_threads()
    line = {
        "impl": "reference", "metric": "newton_steps_per_s", "value": val, "unit": "Newton steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"dense LP n={args.n} m={m} box+-3, warm start (BASELINE configs[1])",
                   "sample": "one full-size Newton iteration per step"},
        "cpu_baseline": {"value": val, "unit": "Newton steps/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} full-size Newton iterations of the oracle (NumPy/SciPy; BLAS: "
                                   f"{blas_info()})"},
        "e2e": {"value": val, "unit": "Newton steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from ipm_b200.LPSolver import LPSolver as _LP
    from ipm_b200.SOCPSolver import SOCPSolver as _SOCP
    import problems

    rows_mode = world > 1 and args.shard == "rows"
    prob, m = workload(args, 0 if rows_mode else rank)
    n = args.n
    socp = args.workload == "socp"
    if socp:
        args.lasso_k = 0
        args.no_cpu_baseline = True

        def LPSolver(**kw):  # same call shape as the LP arm
            return _SOCP(**kw, **problems.SOCP_TEST_SETTINGS)
        host = prob
    else:
        LPSolver = _LP
        host = {k: (torch.as_tensor(v).pin_memory().numpy() if isinstance(v, np.ndarray) else v)
                for k, v in prob.items()}
    x0 = prob["x0"].copy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ resident-data arm
    solver = LPSolver(**{k: v for k, v in host.items() if k != "x0"}, x0=x0.copy(), check_cvxpy=False,
                      suppress_print=True, shard_rows=rows_mode)
    x0_dev = solver.x_dev.clone()

    def one_solve():
        solver.x_dev.copy_(x0_dev)
        solver.solve()
        return sum(solver.inner_iters)

    for _ in range(args.warmup):
        one_solve()
    L = solver.launcher
    L.timed_ops = {"ipm_gemm_tn_f64": [], "ipm_syrk_scatter_f64": [], "range:hessian_formation": []}
    launches0 = L.kernel_launches()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    newton = 0
    with ClockSampler(local_rank) as clk:
        e0.record()
        for _ in range(args.steps):
            newton += one_solve()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    launches = L.kernel_launches() - launches0
    hess = [a.elapsed_time(b) for key in ("ipm_gemm_tn_f64", "ipm_syrk_scatter_f64")
            for a, b, tag in L.timed_ops[key] if tag == "hessian"]
    hform = [a.elapsed_time(b) for a, b, _ in L.timed_ops["range:hessian_formation"]]
    comm_bytes = getattr(solver.ns, "comm_bytes", 0)
    solver_ns = solver.ns
    L.timed_ops = None
    value_ref = solver.value
    m_local = solver.data.rows_w if socp else solver.data.m

    # ------------------------------------------------------------------ end-to-end arm (host buffers)
    e2e_ms, e2e_newton, h2d, d2h = None, 0, 0, 0
    if not args.no_e2e:
        del solver
        torch.cuda.empty_cache()

        def e2e_step():
            s = LPSolver(**{k: v for k, v in host.items() if k != "x0"}, x0=x0.copy(), check_cvxpy=False,
                         suppress_print=True, shard_rows=rows_mode)
            s.solve()
            xs = s.xstar  # device -> host read of the result
            return sum(s.inner_iters), s.data.h2d_bytes + 8 * n, xs.nbytes + 8

        e2e_step()  # warm-up (allocator)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            k, h2d, d2h = e2e_step()
            e2e_newton += k
        f1.record()
        barrier()
        e2e_ms = f0.elapsed_time(f1)

    # ------------------------------------------------------------------ Lasso batch (second headline metric)
    lasso = None
    if args.lasso_k > 0:
        torch.cuda.empty_cache()
        barrier()
        lasso = lasso_section(args.lasso_k, rank, world)
        lt = torch.tensor([lasso["ms"], lasso["ms_e2e"]], dtype=torch.float64, device="cuda")
        ls = torch.tensor([lasso["flop"], float(lasso["launches"])], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(lt, op=dist.ReduceOp.MAX)
            dist.all_reduce(ls, op=dist.ReduceOp.SUM)
        lasso.update(ms=float(lt[0]), ms_e2e=float(lt[1]), flop=float(ls[0]), launches=int(ls[1]))

    # ------------------------------------------------------------------ reduce over ranks (max time, sum work)
    stats = torch.tensor([ms, float(newton), e2e_ms or 0.0, float(e2e_newton), float(launches)], dtype=torch.float64,
                         device="cuda")
    if world > 1:
        tmax = stats.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = stats.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, e2e_ms = float(tmax[0]), float(tmax[2])
        launches = float(tsum[4])
        if not rows_mode:  # independent instances: work adds up; one sharded problem: every rank counted the same steps
            newton, e2e_newton = float(tsum[1]), float(tsum[3])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = newton / (ms * 1e-3)
    hess_ms = float(np.mean(hess)) if hess else None
    flops = float(m_local) * n * (n + 1)  # rows resident on this rank (all m rows unless row-sharded)
    achieved = flops / (hess_ms * 1e-3) / 1e12 if hess_ms else None
    line = {
        "metric": "newton_steps_per_s", "value": value, "unit": "Newton steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if rows_mode else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": (f"SOCP n={n}, {args.cones} cones of 64 rows, P=I, warm start, test_SOCP settings "
                                "(BASELINE configs[3]); one step = one SOCPSolver.solve()" if socp else
                                f"dense LP n={n} m={m} box+-3, warm start (BASELINE configs[1]); one step = one "
                                "LPSolver.solve()"), "l2": "inputs (%.2f GB) larger than L2" % (8e-9 * m * n),
                   "per_rank": ("ONE problem, constraint rows / whole cones sharded over the GPUs, NCCL all-reduce of "
                                "the partial Hessian" if rows_mode else "one copy of the instance per GPU (independent solves), no collective")
                   if world > 1 else "single GPU"},
        "time_to_solve_s": ms * 1e-3 / args.steps,
        "newton_steps_per_solve": newton / args.steps / (1 if rows_mode else world),
        "objective": value_ref, "gpu_launches": int(launches), "clocks": clk.summary(),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s",
                     "frac": achieved / FP64_TENSOR_PEAK_TFLOPS if achieved else None,
                     "traffic": hessian_dram_traffic(n, m), "algorithmic_bytes": 8.0 * (m * n + n * (n + 1) / 2),
                     "kernel": "gemm_tn_persistent_kernel<true> (Hessian C'diag(w)C / W'diag(w)W, upper tiles, "
                               "stream-K remainder)",
                     "flop_per_launch": flops, "ms_per_launch": hess_ms, "launches_timed": len(hess),
                     "peak_source": "FP64 DMMA issue-rate microbenchmark on this pool (tools/fp64_peak.cu, "
                                    "profiles/fp64_peak_r01.json); MEASURED_PEAKS.json has no FP64 entry"},
    }
    if rows_mode and hform:
        peer = getattr(solver_ns, "peer", None) is not None
        line["hessian_formation"] = {"ms": float(np.mean(hform)), "exchange": "peer-memory scatter / reduce / "
                                     "broadcast kernels (no NCCL)" if peer else "NCCL all-reduce of the n x ld buffer",
                                     "what": "local partial C_r' diag(w) C_r + exchange + diagonal terms, per Newton "
                                     "step (rank 0)",
                                     "allreduce_bytes_per_newton_step": comm_bytes / max(newton + args.warmup *
                                                                                        newton / args.steps, 1)}
    if e2e_ms:
        line["e2e"] = {"value": e2e_newton / (e2e_ms * 1e-3), "unit": "Newton steps/s", "h2d_bytes_per_step": int(h2d),
                       "d2h_bytes_per_step": int(d2h), "time_to_solve_s": e2e_ms * 1e-3 / args.steps}
    if lasso is not None:
        K = args.lasso_k
        line["lasso"] = {
            "metric": "lasso_solves_per_s", "workload": f"LassoSolver ADMM, A 2048x512 + bias, K={K} problems, eps 1e-6 "
            "(BASELINE configs[4]); strided column split across ranks, no collective",
            "value": K / (lasso["ms"] * 1e-3), "e2e_value": K / (lasso["ms_e2e"] * 1e-3), "unit": "solves/s",
            "admm_iterations": lasso["iters"], "ms_per_iteration": lasso["ms"] / lasso["iters"],
            "gpu_launches": lasso["launches"], "h2d_bytes": lasso["h2d"], "d2h_bytes": lasso["d2h"],
            "roofline": {"bound": "tensor", "achieved": lasso["flop"] / (lasso["ms"] * 1e-3) / 1e12 / world,
                         "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s per GPU",
                         "frac": lasso["flop"] / (lasso["ms"] * 1e-3) / 1e12 / world / FP64_TENSOR_PEAK_TFLOPS,
                         "note": "2 n^2 K flop per ADMM iteration over the whole solve() incl. stop checks"}}
    if args.factorisation and world == 1:
        line["factorisation"] = factorisation_section(n)
    if not args.no_cpu_baseline and world == 1:
        use_all_host_threads()
        k, dt = cpu_sample(prob, args.cpu_newton_steps)
        line["cpu_baseline"] = {"value": k / dt, "unit": "Newton steps/s", "cores": blas_threads(), "kind": "port",
                                "sample": f"{k} full-size Newton iterations (first centering step) of the oracle, "
                                          f"{dt:.1f} s; NumPy BLAS: {blas_info()}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
