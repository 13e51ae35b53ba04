"""Host-side logic that needs no GPU: the reference's step-size sequence, default starting points, input
validation (same ValueErrors as LPSolver.py:226-318 / QPSolver / SOCPSolver) and method dispatch."""

import numpy as np
import pytest

from ipm_b200 import _solver_base as sb
from ipm_b200.engine import step_table


def test_step_table_is_the_reference_sequence():
    for beta in (0.5, 0.6, 0.7, 0.9):
        tab = step_table(beta)
        a, ref = 1, [1.0]
        while True:  # NewtonSolver.py:175-176
            a *= beta
            ref.append(a)
            if a < 1e-13:
                break
        assert tab == ref
        assert tab[-1] < 1e-13 <= tab[-2]
    with pytest.raises(ValueError):
        step_table(1.0)


def test_default_x0_matches_reference_rules():
    n = 4
    np.testing.assert_array_equal(sb.default_x0(n, np.array(-3.0), np.array(3.0)), np.zeros(n))
    np.testing.assert_allclose(sb.default_x0(n, np.array(0.0), None), 0.1 * np.ones(n))
    np.testing.assert_allclose(sb.default_x0(n, None, np.array(1e9)), (1e2 - 1e-1) * np.ones(n))
    np.random.seed(3)
    want = np.random.rand(n)
    np.random.seed(3)
    np.testing.assert_array_equal(sb.default_x0(n, None, None), want)


def test_lp_validation_errors_come_before_any_device_work():
    from ipm_b200.LPSolver import LPSolver

    c = np.ones(3)
    with pytest.raises(ValueError, match="Both A and b"):
        LPSolver(c=c, A=np.eye(3), check_cvxpy=False)
    with pytest.raises(ValueError, match="Both C and d"):
        LPSolver(c=c, d=np.ones(3), check_cvxpy=False)
    with pytest.raises(ValueError, match="agreeing dimensions"):
        LPSolver(c=c, C=np.eye(3), d=np.ones(2), check_cvxpy=False)
    with pytest.raises(ValueError, match="1-dimensional"):
        LPSolver(c=np.ones((3, 1)), check_cvxpy=False)
    with pytest.raises(ValueError, match="Lower bound must be lower"):
        LPSolver(c=c, C=np.eye(3), d=np.ones(3), lower_bound=2, upper_bound=1, check_cvxpy=False)
    with pytest.raises(ValueError, match="same number of entries"):
        LPSolver(c=c, C=np.ones((2, 4)), d=np.ones(2), check_cvxpy=False)


def test_qp_and_socp_validation():
    from ipm_b200.QPSolver import QPSolver
    from ipm_b200.SOCPSolver import SOCPSolver

    with pytest.raises(ValueError, match="just an LP"):
        QPSolver(P=None, q=np.ones(2))
    with pytest.raises(ValueError, match="square"):
        QPSolver(P=np.ones((2, 3)), q=np.ones(3))
    with pytest.raises(ValueError, match="No cone"):
        SOCPSolver(P=np.eye(2), q=np.ones(2), A=None)
    with pytest.raises(ValueError, match="equal number of A and b"):
        SOCPSolver(q=np.ones(2), A=[np.ones((3, 2)), np.ones((3, 2))], b=[np.ones(3)] * 3, c=[np.ones(2)] * 2,
                   d=[1.0, 1.0])
    with pytest.raises(ValueError, match="d must be a scalar"):
        SOCPSolver(q=np.ones(2), A=[np.ones((3, 2))], b=[np.ones(3)], c=[np.ones(2)], d=[np.ones(2)])


def test_method_dispatch_rules():
    chk = sb.BarrierSolverBase._check_method
    for m in ("cholesky", "np_solve", "np_lstsq", "direct"):
        chk(m, False)
        chk(m, True)
    chk("kkt", True)
    with pytest.raises(ValueError, match="No KKT"):  # LPSolver.py:423-430
        chk("kkt", False)
    with pytest.raises(ValueError, match="valid linear solve"):  # LPSolver.py:447-448
        chk("qr", False)
    chk("cg", False)  # NewtonSolverCG (NewtonSolver.py:365-400): implemented on the device (csrc/cg.cu)
    with pytest.raises(NotImplementedError):  # NewtonSolverInfeasibleStart.py:604
        chk("cg", True)


def test_host_array_answers_get():
    a = sb.HostArray(np.arange(3.0))
    np.testing.assert_array_equal(a.get(), np.arange(3.0))
    assert float(a.sum()) == 3.0


def test_sparse_rows_segment_map_on_cpu():
    """engine.SparseRows: the host-built CSR pair and the Hessian segment map, replayed in NumPy
    exactly as ipm_csr_gemv_f64 / ipm_sparse_syrk_f64 consume them, against the dense formulas."""
    import torch

    from ipm_b200.engine import SparseRows, _looks_sparse

    rs = np.random.RandomState(11)
    m, n = 300, 270
    Cm = np.where(rs.rand(m, n) < 0.012, rs.uniform(-2, 2, (m, n)), 0.0)
    Cm[7] = 0.0
    Cm[8, :] = 0.0
    Cm[8, 5] = 1.5  # a single-entry row
    assert _looks_sparse(Cm) and not _looks_sparse(rs.rand(200, 300))
    sp = SparseRows(Cm, n, torch.device("cpu"))
    rowptr, col, val = sp.rowptr.numpy(), sp.col.numpy(), sp.val.numpy()
    x, v, w = rs.randn(n), rs.randn(m), rs.uniform(0.1, 5.0, m)
    y = np.array([val[rowptr[r]:rowptr[r + 1]] @ x[col[rowptr[r]:rowptr[r + 1]]] for r in range(m)])
    np.testing.assert_allclose(y, Cm @ x, rtol=1e-13, atol=1e-13)
    trp, tc, tv = sp.t_rowptr.numpy(), sp.t_col.numpy(), sp.t_val.numpy()
    g = np.array([tv[trp[j]:trp[j + 1]] @ v[tc[trp[j]:trp[j + 1]]] for j in range(n)])
    np.testing.assert_allclose(g, Cm.T @ v, rtol=1e-13, atol=1e-13)
    segptr, seg_row, seg_prod = sp.segptr.numpy(), sp.seg_row.numpy(), sp.seg_prod.numpy()
    oi, oj = sp.out_i.numpy(), sp.out_j.numpy()
    H = np.zeros((n, n))
    for e in range(sp.nout):
        k0, k1 = segptr[e], segptr[e + 1]
        assert np.all(np.diff(seg_row[k0:k1]) > 0)       # ascending rows: fixed summation order on the device
        H[oi[e], oj[e]] += w[seg_row[k0:k1]] @ seg_prod[k0:k1]
    assert np.all(oi <= oj) and len(set(zip(oi, oj))) == sp.nout
    np.testing.assert_allclose(H, np.triu((Cm * w[:, None]).T @ Cm), rtol=1e-12, atol=1e-13)


def test_miplib_loader_validates_shapes(tmp_path):
    from ipm_b200 import miplib

    prob = miplib.synthetic_network_lp(seed=1, n=60, p=4, m=20, density=0.05)
    f = tmp_path / "ok.npy"
    miplib.save_lp(f, **prob)
    got = miplib.load_lp(f)
    assert set(got) == set(miplib.FIELDS) and got["A"].shape == (4, 60) and got["C"].shape == (20, 60)
    no_eq = dict(prob, A=None, b=None)
    miplib.save_lp(f, **no_eq)
    assert miplib.load_lp(f)["A"] is None and miplib.load_lp(f)["b"] is None
    bad = dict(prob, d=prob["d"][:-1])
    miplib.save_lp(f, **bad)
    with pytest.raises(ValueError):
        miplib.load_lp(f)


@pytest.mark.parametrize("T,G", [(3, 6), (4, 3), (20, 148), (33, 16), (64, 148)])
def test_tile_dag_schedule_cannot_deadlock(T, G):
    """Model of the tile-DAG Cholesky's schedule (csrc/chol.cu, namespace dag): tasks (i, j), i <= j, numbered row-major,
    CTA c runs tasks c, c + G, ... in order and blocks on ONE progress counter per block column.  Whatever the timing,
    the lowest-numbered unfinished task is always runnable, every column is published top-down, and all tasks finish."""
    tasks = [(i, j) for i in range(T) for j in range(i, T)]
    queues = [tasks[c::G] for c in range(G)]
    pos = [0] * G
    done = [0] * T  # published row blocks per block column
    rng = np.random.RandomState(T * 1000 + G)
    finished = 0
    while finished < len(tasks):
        runnable = []
        for c in range(G):
            if pos[c] == len(queues[c]):
                continue
            i, j = queues[c][pos[c]]
            # inputs: U(k, i), U(k, j) for k < i (counters >= i), and U(i, i) for an off-diagonal tile (done[i] >= i + 1)
            if done[i] >= i and done[j] >= i and (i == j or done[i] >= i + 1):
                runnable.append(c)
        assert runnable, "deadlock"
        lowest = min(tasks.index(queues[c][pos[c]]) for c in range(G) if pos[c] < len(queues[c]))
        assert lowest in [tasks.index(queues[c][pos[c]]) for c in runnable]
        c = runnable[rng.randint(len(runnable))]  # an arbitrary runnable CTA finishes next
        i, j = queues[c][pos[c]]
        assert done[j] == i  # top-down: the counter of column j goes i -> i + 1
        done[j] = i + 1
        pos[c] += 1
        finished += 1
    assert done == list(range(1, T + 1))


def test_flat_module_layout_imports_like_the_reference():
    """with interiorpoint-gpu_b200/ on sys.path the reference's flat module names work
    (`from LPSolver import LPSolver`, ...).  Fresh interpreter, so the package import of this test session cannot mask
    a missing flat-import fallback."""
    import os
    import subprocess
    import sys

    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "interiorpoint-gpu_b200")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from LPSolver import LPSolver\nfrom QPSolver import QPSolver\nfrom SOCPSolver import SOCPSolver\n"
            "from LassoSolver import LassoSolver\nfrom PhaseOneSolver import PhaseOneSolver\nimport PhaseOne, harness, miplib\n"
            "print(LPSolver.__name__, QPSolver.__name__, SOCPSolver.__name__, LassoSolver.__name__)\n" % pkg)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["LPSolver", "QPSolver", "SOCPSolver", "LassoSolver"]


def test_int8_hessian_policy(monkeypatch):
    """engine.hess_i8_slices: size-gated by default (the INT8 kernel only pays off on large operands), IPM_HESSIAN_I8
    forces it off / on / to a digit count, and shapes the kernel does not support fall back to the DMMA kernel."""
    from ipm_b200 import engine

    monkeypatch.delenv("IPM_HESSIAN_I8", raising=False)
    assert engine.hess_i8_slices(16384, 8192) == 8 and engine.hess_i8_slices(4096, 2048) == 8
    assert engine.hess_i8_slices(16384, 1024) == 0 and engine.hess_i8_slices(512, 8192) == 0
    assert engine.hess_i8_slices(2048, 8192) == 8      # an eighth of cfg 2's rows: the per-rank shard on 8 GPUs
    assert engine.hess_i8_slices(0, 8192) == 0
    monkeypatch.setenv("IPM_HESSIAN_I8", "0")
    assert engine.hess_i8_slices(16384, 8192) == 0
    monkeypatch.setenv("IPM_HESSIAN_I8", "1")
    assert engine.hess_i8_slices(100, 64) == 8
    assert engine.hess_i8_slices(70000, 64) == 0          # INT32 accumulators bound the contraction length
    monkeypatch.setenv("IPM_HESSIAN_I8", "6")
    assert engine.hess_i8_slices(100, 64) == 6


def test_int8_hessian_scheme_on_the_cpu():
    """The arithmetic of csrc/hess_i8.cu restated in NumPy (no GPU): digits by the same recurrence, exact integer slice-pair
    products, FP64 recombination.  (1) 8 digits reach the FP64 kernel's error class on weights spanning 20 decades;
    (2) every digit lies in [-64, 64]; (3) the INT32 accumulators cannot overflow for the longest contraction the library
    accepts (ipm_hess_i8_ws_bytes refuses m > 65408)."""
    rs = np.random.RandomState(0)
    m, n, s = 300, 40, 8
    Cm = rs.rand(m, n) * 4 - 2
    w = 10.0 ** (rs.rand(m) * 20 - 10)
    X = np.sqrt(w)[:, None] * Cm
    amax = np.abs(X).max(axis=0)
    e = np.frexp(amax)[1]                      # amax = f 2^e, f in [1/2, 1): sigma = 2^(e + 1) >= 2 amax
    sigma = np.ldexp(1.0, e + 1)
    r = X / sigma
    Q = []
    for _ in range(s):
        r = r * 128.0
        q = np.rint(r)
        assert np.abs(q).max() <= 64
        Q.append(q.astype(np.int64))
        r = r - q
        assert np.abs(r).max() <= 0.5
    H = np.zeros((n, n))
    for d in range(s - 1, -1, -1):             # Horner over the diagonals t + u = d, as the epilogue does
        acc = sum(Q[t].T @ Q[d - t] for t in range(d + 1))
        assert np.abs(acc).max() < 2 ** 31
        H = H * 2.0 ** -7 + acc
    H = H * 2.0 ** -14 * sigma[:, None] * sigma[None, :]
    ref = X.T @ X
    scale = np.abs(X).T @ np.abs(X)
    assert (np.abs(H - ref) / scale).max() < 1e-14
    # worst case: all digits +-64, 8 slice pairs on one diagonal, the longest padded contraction
    assert 8 * 65408 * 64 * 64 < 2 ** 31
    from ipm_b200 import _abi
    assert _abi.lib().ipm_hess_i8_ws_bytes(65408, 128, 8) > 0 and _abi.lib().ipm_hess_i8_ws_bytes(65409, 128, 8) == 0
