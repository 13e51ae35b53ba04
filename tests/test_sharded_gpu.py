"""Row-sharded LP / QP and cone-sharded SOCP across 2 GPUs (NCCL + peer memory): same optimum / Newton counts as the
reference goldens, with and without equality constraints (infeasible-start method: the block elimination runs replicated,
the barrier pieces sharded), and the dual variables gathered from the shards.
Skipped on a single-GPU box (the CPU gloo test covers the partitioning logic there)."""

import os

import numpy as np
import pytest
import torch

import problems
from conftest import load_golden
from test_solvers_gpu import assert_iters_close, noise_dominated_steps

pytestmark = pytest.mark.gpu

CASES = {c["name"]: c for f in ("barrier_cases.json", "dual_cases.json", "large_cases.json") for c in load_golden(f)}


def _worker(rank, world, port, name, q, env=None):
    import torch.distributed as dist

    os.environ.update(env or {})
    # the distributed factorisation is off by default below n = 12288 (sharded_engine.PEER_POTRF_MIN_N): force it on so
    # that the n >= 512 cases here run ipm_potrf_upper_peer_f64
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), IPM_PEER_POTRF="1")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from ipm_b200.LPSolver import LPSolver
    from ipm_b200.QPSolver import QPSolver
    from ipm_b200.SOCPSolver import SOCPSolver

    case = CASES[name]
    prob = getattr(problems, case["generator"])(**case["generator_kwargs"])
    if isinstance(prob, list):
        prob = prob[case.get("index") or 0]
    cls = {"LPSolver": LPSolver, "QPSolver": QPSolver, "SOCPSolver": SOCPSolver}[case["solver"]]
    np.random.seed(0)
    duals = "lam_star" in case
    s = cls(**prob, check_cvxpy=False, suppress_print=True, shard_rows=True, get_dual_variables=duals, **case["settings"])
    assert s.sharded
    val = s.solve()
    p1 = s.phase1_solver.inner_iters if case.get("phase1_inner_iters") is not None else None
    if rank == 0:
        q.put((val, s.inner_iters, p1, np.asarray(s.xstar),
               (s.ns.peer is not None, bool(getattr(s.ns, "peer_potrf", False))), getattr(s.ns, "peer_error", None),
               np.asarray(s.lam_star) if duals else None, np.asarray(s.v_star) if duals and s.v_star is not None else None))
    dist.barrier()
    dist.destroy_process_group()


def _run(name, env=None):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q, env)) for r in range(2)]
    import queue
    import time

    for p in procs:
        p.start()
    out, t0 = None, time.time()
    while out is None:
        try:
            out = q.get(timeout=2)
        except queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            if dead or time.time() - t0 > 240:  # a crashed rank must not cost the whole timeout
                for p in procs:
                    p.kill()
                pytest.fail(f"worker exit codes {[p.exitcode for p in procs]} after {time.time() - t0:.0f} s")
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


@pytest.mark.parametrize("name", ["lp_dense_n256_cold", "lp_dense_n97_ragged", "socp_n48_warm", "socp_n48_cold",
                                  "socp_n96_warm", "lp_seed1_n100_0", "qp_seed1_n100_0", "qp_dense_n512",
                                  "socp_n48_eq_warm", "lp_dense_n1024_warm", "lp_dense_n1024_cold"])
def test_row_sharded_matches_reference(name):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    case = CASES[name]
    val, iters, p1, x, peer, peer_error, _, _ = _run(name)
    # the Hessian exchange ran over peer memory (fused SYRK + reduce-scatter + all-gather), not the NCCL fallback; the
    # n = 1024 cases also run the factorisation distributed over the two GPUs (ipm_potrf_upper_peer_f64)
    assert peer[0], peer_error
    assert peer[1] == (len(case["xstar"]) > 384), "distributed factorisation not exercised"
    assert val == pytest.approx(case["value"], rel=1e-6, abs=1e-9)
    prob = getattr(problems, case["generator"])(**case["generator_kwargs"])
    if isinstance(prob, list):
        prob = prob[case.get("index") or 0]
    cap = case["settings"].get("max_inner_iters", 50)
    assert_iters_close(iters, case["inner_iters"], cap=cap, noisy=noise_dominated_steps(case, prob, case["settings"]))
    if p1 is not None:
        assert_iters_close(p1, case["phase1_inner_iters"])
    assert np.linalg.norm(x - np.array(case["xstar"])) <= 1e-4 * (1 + np.linalg.norm(case["xstar"]))


@pytest.mark.parametrize("name", ["lp_dense_n1024_warm", "lp_dense_n97_ragged", "qp_dense_n512", "socp_n96_warm"])
def test_row_sharded_int8_hessian(name):
    """The same split with every rank's partial Hessian on the INT8 tensor pipe (ipm_hess_i8_scatter_f64: half tiles into the
    owners' inboxes) -- what `bench.py --gpus N` runs at cfg-2 size."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    case = CASES[name]
    val, iters, p1, x, peer, peer_error, _, _ = _run(name, {"IPM_HESSIAN_I8": "1"})
    assert peer[0], peer_error
    assert val == pytest.approx(case["value"], rel=1e-6, abs=1e-9)
    prob = getattr(problems, case["generator"])(**case["generator_kwargs"])
    if isinstance(prob, list):
        prob = prob[case.get("index") or 0]
    cap = case["settings"].get("max_inner_iters", 50)
    assert_iters_close(iters, case["inner_iters"], cap=cap, noisy=noise_dominated_steps(case, prob, case["settings"]))
    if p1 is not None:
        assert_iters_close(p1, case["phase1_inner_iters"])
    assert np.linalg.norm(x - np.array(case["xstar"])) <= 1e-4 * (1 + np.linalg.norm(case["xstar"]))


def test_row_sharded_dual_variables():
    """get_dual_variables=True with the rows sharded: lam_star gathered in the reference's slack layout, v_star replicated."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    case = CASES["lp_seed1_n100_0_duals"]
    val, _, _, _, _, _, lam, v = _run("lp_seed1_n100_0_duals")
    assert val == pytest.approx(case["value"], rel=1e-6, abs=1e-9)
    lam_ref, v_ref = np.array(case["lam_star"]), np.array(case["v_star"])
    assert lam.shape == lam_ref.shape and np.all(lam > 0)
    assert np.linalg.norm(lam - lam_ref) <= 1e-2 * np.linalg.norm(lam_ref)
    assert np.linalg.norm(v - v_ref) <= 1e-2 * (1e-12 + np.linalg.norm(v_ref))
