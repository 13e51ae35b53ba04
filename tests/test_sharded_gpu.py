"""Row-sharded LP and cone-sharded SOCP across 2 GPUs (NCCL): same optimum / Newton counts as the reference goldens.
Skipped on a single-GPU box (the CPU gloo test covers the partitioning logic there)."""

import os

import numpy as np
import pytest
import torch

import problems
from conftest import load_golden

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, name, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from ipm_b200.LPSolver import LPSolver
    from ipm_b200.SOCPSolver import SOCPSolver

    case = {c["name"]: c for c in load_golden("barrier_cases.json")}[name]
    prob = getattr(problems, case["generator"])(**case["generator_kwargs"])
    cls = {"LPSolver": LPSolver, "SOCPSolver": SOCPSolver}[case["solver"]]
    np.random.seed(0)
    s = cls(**prob, check_cvxpy=False, suppress_print=True, shard_rows=True, **case["settings"])
    assert s.sharded
    val = s.solve()
    p1 = s.phase1_solver.inner_iters if case["phase1_inner_iters"] is not None else None
    if rank == 0:
        q.put((val, s.inner_iters, p1, np.asarray(s.xstar), s.ns.peer is not None, getattr(s.ns, "peer_error", None)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["lp_dense_n256_cold", "lp_dense_n97_ragged", "socp_n48_warm", "socp_n48_cold",
                                  "socp_n96_warm"])
def test_row_sharded_matches_reference(name):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    case = {c["name"]: c for c in load_golden("barrier_cases.json")}[name]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    val, iters, p1, x, peer, peer_error = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # the Hessian exchange ran over peer memory (fused SYRK + reduce-scatter + all-gather), not the NCCL fallback
    assert peer, peer_error
    assert val == pytest.approx(case["value"], rel=1e-6, abs=1e-9)
    assert len(iters) == len(case["inner_iters"])
    assert all(abs(a - b) <= 2 or b >= 50 for a, b in zip(iters, case["inner_iters"])), (iters, case["inner_iters"])
    if p1 is not None:
        assert all(abs(a - b) <= 2 for a, b in zip(p1, case["phase1_inner_iters"])), (p1, case["phase1_inner_iters"])
    assert np.linalg.norm(x - np.array(case["xstar"])) <= 1e-4 * (1 + np.linalg.norm(case["xstar"]))
