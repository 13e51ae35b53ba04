"""world_size-2 gloo test (CPU) of the multi-GPU partitioning logic: strided column split == the reference's
num_chunks split, row-sharded partial Hessian / gradient / barrier sums + all-reduce == the single-process values."""

import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ipm_b200 import dist as D

    rs = np.random.RandomState(0)
    m, n, K = 37, 11, 10
    C = rs.uniform(-2, 2, (m, n))
    s = rs.uniform(0.1, 2.0, m)
    lo, hi = D.row_range(m, rank, world)
    w = 1.0 / s[lo:hi] ** 2
    H = torch.as_tensor(C[lo:hi].T @ (w[:, None] * C[lo:hi]))
    g = torch.as_tensor(C[lo:hi].T @ (1.0 / s[lo:hi]))
    red = torch.tensor([np.log(s[lo:hi]).sum()])
    mn = torch.tensor([s[lo:hi].min()])
    kmax = torch.tensor([3 + rank], dtype=torch.int32)
    D.allreduce_sum_(H), D.allreduce_sum_(g), D.allreduce_sum_(red), D.allreduce_min_(mn), D.allreduce_max_(kmax)
    cols = D.strided_columns(K, rank, world)
    X = D.gather_columns(np.full((2, len(cols)), float(rank)), cols, K)
    if rank == 0:
        q.put((H.numpy(), g.numpy(), float(red), float(mn), int(kmax), X, cols))
    dist.destroy_process_group()


def test_row_sharded_reduction_and_column_split():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    H, g, red, mn, kmax, X, cols0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rs = np.random.RandomState(0)
    m, n, K = 37, 11, 10
    C = rs.uniform(-2, 2, (m, n))
    s = rs.uniform(0.1, 2.0, m)
    np.testing.assert_allclose(H, C.T @ ((1 / s**2)[:, None] * C), rtol=1e-13)
    np.testing.assert_allclose(g, C.T @ (1 / s), rtol=1e-13)
    assert red == pytest.approx(np.log(s).sum(), rel=1e-13)
    assert mn == s.min() and kmax == 4
    # strided split == LassoSolver num_chunks semantics (LassoSolver.py:349-351)
    np.testing.assert_array_equal(cols0, np.arange(K)[0::2])
    np.testing.assert_array_equal(X[0], np.arange(K) % 2)


def test_row_range_covers_everything():
    from ipm_b200.dist import row_range

    for m in (1, 7, 16384, 16385):
        for w in (1, 2, 3, 8):
            spans = [row_range(m, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == m
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
