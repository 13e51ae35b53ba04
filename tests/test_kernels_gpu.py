"""Kernel-level parity (GPU): every C-ABI op against float64 NumPy/torch on the same inputs.

Tolerances are FP64 reorder-level: the DMMA contraction sums in a different order than BLAS, so results
agree to ~1e-13 relative to sum |a||b| (SURVEY 7.4-1 shows the solver needs <= 1e-13)."""

import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from ipm_b200 import _abi  # noqa: E402


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).cuda()


def padded(a, mult=16):
    """Row-major device copy with the leading dimension rounded up to `mult` doubles."""
    r, c = a.shape
    ld = (c + mult - 1) // mult * mult
    out = torch.zeros((r, ld), dtype=torch.float64, device="cuda")
    out[:, :c] = torch.as_tensor(a)
    return out, ld


@pytest.fixture(scope="module", autouse=True)
def _device():
    _abi.require_device()


@pytest.mark.parametrize("M,N,K,use_w,upper", [
    (128, 128, 64, False, 0), (256, 384, 200, True, 0), (100, 100, 37, True, 1), (513, 300, 513, False, 0),
    (1000, 1000, 777, True, 1), (8, 8, 4, True, 1), (130, 70, 1, False, 0), (257, 257, 1030, True, 1),
])
def test_gemm_tn(M, N, K, use_w, upper):
    rs = np.random.RandomState(M * 7 + N * 3 + K)
    A = rs.uniform(-2, 2, (K, M))
    B = A if upper else rs.uniform(-2, 2, (K, N))
    w = 10.0 ** rs.uniform(-6, 6, K) if use_w else None
    D0 = rs.uniform(-1, 1, (M, N))
    alpha, beta = -1.25, 0.5
    Ad, lda = padded(A)
    Bd, ldb = (Ad, lda) if upper else padded(B)
    Dd, ldd = padded(D0)
    wd = dev(w) if use_w else None
    _abi.call("ipm_gemm_tn_f64", Ad.data_ptr(), lda, Bd.data_ptr(), ldb, _abi.ptr(wd), alpha, beta, Dd.data_ptr(), ldd,
              M, N, K, upper, None)
    torch.cuda.synchronize()
    got = Dd[:, :N].cpu().numpy()
    Aw = A * (w[:, None] if use_w else 1.0)
    ref = beta * D0 + alpha * (Aw.T @ B)
    scale = np.abs(Aw).T @ np.abs(B) + np.abs(D0)
    if upper:
        iu = np.triu_indices(M)
        assert np.max(np.abs(got[iu] - ref[iu]) / scale[iu]) < 1e-14
        il = np.tril_indices(M, -1)
        np.testing.assert_array_equal(got[il], D0[il])  # strict lower triangle untouched
    else:
        assert np.max(np.abs(got - ref) / scale) < 1e-14
    assert torch.count_nonzero(Dd[:, N:]) == 0  # padding untouched


@pytest.mark.parametrize("M,N,K,use_w,upper", [
    (2304, 2304, 1030, True, 1),    # 171 upper tiles on 148 SMs: 1 full wave + 23 stream-K remainder tiles
    (1000, 2500, 1024, False, 0),   # 160 tiles: 12 remainder tiles, K a multiple of the k-tile
    (2400, 2400, 3000, True, 1),    # 190 tiles, ragged M / K
    (1280, 2048, 1200, True, 0),    # 160 tiles, rectangular, weighted
])
def test_gemm_tn_persistent_stream_k(M, N, K, use_w, upper):
    """Long-K contractions with more tiles than SMs take the persistent kernel (data-parallel waves + stream-K
    remainder with deterministic fix-up, gemm_tn_core.cuh); checked against float64 torch.matmul on the device and
    for run-to-run bit reproducibility."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.rand((K, M), dtype=torch.float64, device="cuda", generator=g) * 4 - 2
    B = A if upper else torch.rand((K, N), dtype=torch.float64, device="cuda", generator=g) * 4 - 2
    w = 10.0 ** (torch.rand(K, dtype=torch.float64, device="cuda", generator=g) * 8 - 4) if use_w else None
    D0 = torch.rand((M, N), dtype=torch.float64, device="cuda", generator=g)
    alpha, beta = 0.75, -0.5
    Ad, lda = padded(A)
    Bd, ldb = (Ad, lda) if upper else padded(B)
    outs = []
    for _ in range(2):
        Dd, ldd = padded(D0)
        _abi.call("ipm_gemm_tn_f64", Ad.data_ptr(), lda, Bd.data_ptr(), ldb, _abi.ptr(w), alpha, beta, Dd.data_ptr(),
                  ldd, M, N, K, upper, None)
        torch.cuda.synchronize()
        outs.append(Dd)
    assert torch.equal(outs[0], outs[1])  # deterministic fix-up order
    got = outs[0][:, :N]
    Aw = A * w[:, None] if use_w else A
    ref = beta * D0 + alpha * (Aw.T @ B)
    scale = Aw.abs().T @ B.abs() + D0.abs()
    err = (got - ref).abs() / scale
    if upper:
        assert float(torch.triu(err).max()) < 1e-14
        assert torch.equal(torch.tril(got, -1), torch.tril(D0, -1))
    else:
        assert float(err.max()) < 1e-14
    assert torch.count_nonzero(outs[0][:, N:]) == 0


def test_gemm_tn_beta_zero_ignores_nan_output():
    rs = np.random.RandomState(0)
    A = rs.randn(50, 40)
    Ad, lda = padded(A)
    Dd = torch.full((40, 48), float("nan"), dtype=torch.float64, device="cuda")
    _abi.call("ipm_gemm_tn_f64", Ad.data_ptr(), lda, Ad.data_ptr(), lda, None, 1.0, 0.0, Dd.data_ptr(), 48, 40, 40, 50,
              0, None)
    torch.cuda.synchronize()
    np.testing.assert_allclose(Dd[:, :40].cpu().numpy(), A.T @ A, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("rows,cols", [(1, 1), (7, 33), (300, 1000), (1000, 257), (2049, 4096)])
def test_gemv_n_and_t(rows, cols):
    rs = np.random.RandomState(rows + cols)
    Mx = rs.uniform(-2, 2, (rows, cols))
    x = rs.randn(cols)
    y0 = rs.randn(rows)
    Md, ld = padded(Mx)
    yd = dev(y0)
    _abi.call("ipm_gemv_n_f64", Md.data_ptr(), ld, rows, cols, dev(x).data_ptr(), yd.data_ptr(), 2.0, -1.0, None)
    torch.cuda.synchronize()
    np.testing.assert_allclose(yd.cpu().numpy(), 2.0 * (Mx @ x) - y0, rtol=1e-12, atol=1e-12)
    # transposed, two right-hand sides
    V = rs.randn(2, rows)
    Vd = dev(V)
    Yd = torch.zeros((2, cols), dtype=torch.float64, device="cuda")
    nws = _abi.lib().ipm_gemv_t_ws_doubles(rows, cols, 2)
    ws = torch.empty(nws, dtype=torch.float64, device="cuda")
    _abi.call("ipm_gemv_t_f64", Md.data_ptr(), ld, rows, cols, Vd.data_ptr(), 2, rows, Yd.data_ptr(), cols, 1.0, 0.0,
              ws.data_ptr(), nws, None)
    torch.cuda.synchronize()
    np.testing.assert_allclose(Yd.cpu().numpy(), V @ Mx, rtol=1e-12, atol=1e-11)


def spd(n, seed, cond_pow=6):
    rs = np.random.RandomState(seed)
    C_ = rs.uniform(-2, 2, (2 * n, n))
    w = 10.0 ** rs.uniform(-cond_pow, cond_pow, 2 * n)
    return (C_ * w[:, None]).T @ C_ + np.eye(n) * 1e-3


@pytest.mark.parametrize("n", [1, 5, 64, 128, 129, 300, 513, 1025])
def test_potrf_trsv(n):
    H = spd(n, n)
    Hd, ld = padded(H)
    Hd_low_before = torch.tril(Hd[:, :n], -1).clone()
    info = torch.full((1,), -7, dtype=torch.int32, device="cuda")
    _abi.call("ipm_potrf_upper_f64", Hd.data_ptr(), ld, n, info.data_ptr(), None)
    torch.cuda.synchronize()
    assert int(info.item()) == 0
    U = torch.triu(Hd[:, :n]).cpu().numpy()
    assert np.max(np.abs(U.T @ U - H)) / np.max(np.abs(H)) < 1e-13
    assert torch.equal(torch.tril(Hd[:, :n], -1), Hd_low_before)  # strict lower triangle untouched
    # solves
    rs = np.random.RandomState(n)
    b = rs.randn(n)
    bd = dev(b)
    tws = torch.zeros(n, dtype=torch.float64, device="cuda")
    _abi.call("ipm_trsv_upper_f64", Hd.data_ptr(), ld, n, bd.data_ptr(), 1, tws.data_ptr(), None)
    _abi.call("ipm_trsv_upper_f64", Hd.data_ptr(), ld, n, bd.data_ptr(), 0, tws.data_ptr(), None)
    torch.cuda.synchronize()
    x = bd.cpu().numpy()
    # backward-stable solve: residual small relative to |H||x| + |b|
    res = np.abs(H @ x - b) / (np.abs(H) @ np.abs(x) + np.abs(b))
    assert res.max() < 1e-12


@pytest.mark.parametrize("n,odd_ld", [(1, False), (37, False), (128, False), (200, True), (777, False), (2500, False),
                                      (1031, True)])
def test_trsv_both_directions(n, odd_ld):
    """Persistent left-looking trsv (csrc/trsv.cu) against scipy.linalg.solve_triangular, both directions, ragged
    sizes, and the scalar-load path (odd leading dimension)."""
    from scipy.linalg import solve_triangular

    rs = np.random.RandomState(n)
    U = np.triu(rs.uniform(-1, 1, (n, n))) / np.sqrt(n) + np.diag(rs.uniform(0.5, 2.0, n))
    b = rs.randn(n)
    if odd_ld:
        ld = n + 1 if (n + 1) % 2 else n + 2
        Ud = torch.full((n, ld), float("nan"), dtype=torch.float64, device="cuda")
        Ud[:, :n] = torch.as_tensor(U, device="cuda")
    else:
        Ud, ld = padded(U)
    for trans in (1, 0):
        bd = dev(b)
        _abi.call("ipm_trsv_upper_f64", Ud.data_ptr(), ld, n, bd.data_ptr(), trans, None, None)
        torch.cuda.synchronize()
        x = bd.cpu().numpy()
        ref = solve_triangular(U, b, trans=trans, lower=False)
        M = U.T if trans else U
        res = np.abs(M @ x - b) / (np.abs(M) @ np.abs(x) + np.abs(b))
        assert res.max() < 1e-13
        np.testing.assert_allclose(x, ref, rtol=1e-9, atol=1e-12)


def test_trsv_more_blocks_than_sms():
    """n / 128 > #SMs: every CTA owns several solution blocks (ascending order keeps the flag chain deadlock-free)."""
    n = 148 * 128 + 333
    g = torch.Generator(device="cuda").manual_seed(1)
    U = torch.rand((n, n), dtype=torch.float64, device="cuda", generator=g)
    U.mul_(2.0).sub_(1.0).div_(float(np.sqrt(n))).triu_()
    U.diagonal().copy_(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) + 0.5)
    b = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    for trans in (1, 0):
        x = b.clone()
        _abi.call("ipm_trsv_upper_f64", U.data_ptr(), n, n, x.data_ptr(), trans, None, None)
        torch.cuda.synchronize()
        M = U.T if trans else U
        res = (M @ x - b).abs() / (M.abs() @ x.abs() + b.abs())
        assert float(res.max()) < 1e-13


@pytest.mark.parametrize("n", [200, 385, 1024, 2500])
def test_potrf_tile_dag(n):
    """Single-launch tile-DAG factorisation (csrc/chol.cu, namespace dag): ragged last tile (385 = 3 * 128 + 1), more tasks
    than SMs (2500: 210 tiles), and a size it hands to the stream-ordered code (200: two tiles)."""
    H = spd(n, n + 7)
    Hd, ld = padded(H)
    low_before = torch.tril(Hd[:, :n], -1).clone()
    info = torch.full((1,), -7, dtype=torch.int32, device="cuda")
    for _ in range(2):  # second call: the progress counters carry the previous call's epoch
        Hd[:, :n] = torch.as_tensor(H, device="cuda")
        _abi.call("ipm_potrf_upper_dag_f64", Hd.data_ptr(), ld, n, info.data_ptr(), None)
        torch.cuda.synchronize()
        assert int(info.item()) == 0
        U = torch.triu(Hd[:, :n]).cpu().numpy()
        assert np.max(np.abs(U.T @ U - H)) / np.max(np.abs(H)) < 1e-13
        assert torch.equal(torch.tril(Hd[:, :n], -1), low_before)


def test_potrf_tile_dag_reports_first_bad_pivot():
    n = 700
    H = spd(n, 5, cond_pow=1)
    H[600, 600] = -1.0
    Hd, ld = padded(H)
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    _abi.call("ipm_potrf_upper_dag_f64", Hd.data_ptr(), ld, n, info.data_ptr(), None)
    torch.cuda.synchronize()
    assert int(info.item()) == 601


def test_watchdog_bounds_a_wait_that_never_ends():
    """common.cuh: spin_wait -- a device-side wait whose condition never holds gives up after the spin limit, records its
    code in the pinned fault word (ipm_device_fault) and the kernel returns; a later wait of the process returns at once."""
    L = _abi.lib()
    fn = L.ipm_internal_watchdog_selftest
    fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    gave_up = torch.full((1,), -1, dtype=torch.int32, device="cuda")
    L.ipm_clear_device_fault()
    old = L.ipm_set_spin_limit(8)  # 8 * 2^20 cycles ~ 4 ms
    try:
        assert fn(flag.data_ptr(), gave_up.data_ptr(), 42, None) == 0
        torch.cuda.synchronize()
        assert int(gave_up.item()) == 1 and L.ipm_device_fault() == 42
        with pytest.raises(_abi.IpmError, match="watchdog"):
            _abi.check_device_fault()
        assert L.ipm_device_fault() == 0  # check_device_fault clears it
    finally:
        L.ipm_set_spin_limit(old)
        L.ipm_clear_device_fault()


def test_potrf_reports_first_bad_pivot():
    n = 200
    H = spd(n, 3, cond_pow=1)
    H[150, 150] = -1.0
    Hd, ld = padded(H)
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    _abi.call("ipm_potrf_upper_f64", Hd.data_ptr(), ld, n, info.data_ptr(), None)
    torch.cuda.synchronize()
    assert int(info.item()) == 151  # LAPACK convention: 1-based order of the failing leading minor


@pytest.mark.parametrize("n,p", [(64, 10), (300, 129), (513, 64)])
def test_trsm_upper_t(n, p):
    H = spd(n, n + 1, cond_pow=2)
    U = np.linalg.cholesky(H).T
    rs = np.random.RandomState(p)
    B = rs.randn(n, p)
    Ud, ldu = padded(np.triu(U))
    Bd, ldb = padded(B)
    _abi.call("ipm_trsm_upper_t_f64", Ud.data_ptr(), ldu, n, Bd.data_ptr(), ldb, p, None)
    torch.cuda.synchronize()
    Y = Bd[:, :p].cpu().numpy()
    res = np.abs(U.T @ Y - B) / (np.abs(U.T) @ np.abs(Y) + np.abs(B))
    assert res.max() < 1e-12


def test_dots_and_axpy():
    rs = np.random.RandomState(5)
    a, b, c = rs.randn(1000), rs.randn(1000), rs.randn(77)
    ad, bd, cd = dev(a), dev(b), dev(c)
    out = torch.zeros(3, dtype=torch.float64, device="cuda")
    pa = (C.c_void_p * 3)(ad.data_ptr(), ad.data_ptr(), cd.data_ptr())
    pb = (C.c_void_p * 3)(bd.data_ptr(), ad.data_ptr(), cd.data_ptr())
    nn = (C.c_int * 3)(1000, 1000, 77)
    _abi.call("ipm_dots_f64", 3, pa, pb, nn, out.data_ptr(), None)
    step = dev(np.array([0.36]))
    _abi.call("ipm_axpy_dev_f64", 1000, step.data_ptr(), bd.data_ptr(), ad.data_ptr(), None)
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), [a @ b, a @ a, c @ c], rtol=1e-13)
    np.testing.assert_array_equal(ad.cpu().numpy(), a + 0.36 * b)  # bit-exact NumPy rounding


def test_sparse_rows_kernels():
    """CSR GEMV (both orientations) and the segment-map Hessian C' diag(w) C against dense NumPy."""
    from ipm_b200.engine import SparseRows

    rs = np.random.RandomState(3)
    m, n = 700, 333
    Cm = np.where(rs.rand(m, n) < 0.01, rs.uniform(-2, 2, (m, n)), 0.0)
    Cm[5] = 0.0  # an empty row
    sp = SparseRows(Cm, n, torch.device("cuda"))
    x, v, w = rs.randn(n), rs.randn(m), 10.0 ** rs.uniform(-3, 3, m)
    xd, vd, wd = dev(x), dev(v), dev(w)
    y, g = torch.zeros(m, dtype=torch.float64, device="cuda"), torch.zeros(n, dtype=torch.float64, device="cuda")
    _abi.call("ipm_csr_gemv_f64", sp.rowptr.data_ptr(), sp.col.data_ptr(), sp.val.data_ptr(), m, xd.data_ptr(),
              y.data_ptr(), 1.0, 0.0, None)
    _abi.call("ipm_csr_gemv_f64", sp.t_rowptr.data_ptr(), sp.t_col.data_ptr(), sp.t_val.data_ptr(), n, vd.data_ptr(),
              g.data_ptr(), 1.0, 0.0, None)
    ld = 336
    H0 = rs.randn(n, ld)
    Hd = dev(H0)
    _abi.call("ipm_sparse_syrk_f64", sp.nout, sp.segptr.data_ptr(), sp.seg_row.data_ptr(), sp.seg_prod.data_ptr(),
              sp.out_i.data_ptr(), sp.out_j.data_ptr(), wd.data_ptr(), Hd.data_ptr(), ld, None)
    torch.cuda.synchronize()
    np.testing.assert_allclose(y.cpu().numpy(), Cm @ x, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(g.cpu().numpy(), Cm.T @ v, rtol=1e-13, atol=1e-13)
    ref = H0.copy()
    ref[:, :n] += np.triu((Cm * w[:, None]).T @ Cm)
    got = Hd.cpu().numpy()
    scale = np.abs(Cm * w[:, None]).T @ np.abs(Cm) + np.abs(H0[:, :n])
    assert np.max(np.abs(got[:, :n] - ref[:, :n]) / (scale + 1e-300)) < 1e-14
    np.testing.assert_array_equal(got[:, n:], H0[:, n:])


@pytest.mark.parametrize("int8", [False, True], ids=["dmma", "int8"])
@pytest.mark.parametrize("n,m,R", [(300, 500, 2), (1100, 900, 3), (257, 64, 1)])
def test_peer_hessian_kernels_emulated_ranks(n, m, R, int8):
    """ipm_syrk_scatter_f64 (FP64 DMMA) or ipm_hess_i8_scatter_f64 (INT8 tensor pipe, two half tiles per 128 x 128 tile)
    + ipm_hess_reduce_bcast_f64 (row-sharded Hessian over peer memory) with R ranks emulated
    on ONE device: every rank has its own inbox / flags / H / counter and its own stream, the "peer" pointers are
    simply the other ranks' buffers.  All R Hessians must equal C' diag(w) C + tP * P and be bit-identical."""
    rs = np.random.RandomState(n + m)
    Cm = rs.uniform(-2, 2, (m, n))
    w = 10.0 ** rs.uniform(-3, 3, m)
    Pm = rs.uniform(-1, 1, (n, n))
    Pm = Pm + Pm.T
    tP = 0.75
    T = (n + 127) // 128
    tiles = T * (T + 1) // 2
    slots = (tiles + R - 1) // R
    ldh = (n + 15) // 16 * 16
    bounds = [m * r // R for r in range(R + 1)]
    Cd = [padded(Cm[bounds[r]:bounds[r + 1]]) for r in range(R)]
    wd = [dev(w[bounds[r]:bounds[r + 1]]) for r in range(R)]
    Pd, ldp = padded(Pm)
    inbox = [torch.zeros(R * slots * 128 * 128, dtype=torch.float64, device="cuda") for _ in range(R)]
    sig = [torch.zeros(slots * R + 64, dtype=torch.int32, device="cuda") for _ in range(R)]
    H = [torch.full((n, ldh), 7.0, dtype=torch.float64, device="cuda") for _ in range(R)]
    arr = lambda ts, off=0: (C.c_void_p * R)(*[t.data_ptr() + off for t in ts])  # noqa: E731
    p_inbox, p_flags, p_H, p_done = arr(inbox), arr(sig), arr(H), arr(sig, 4 * slots * R)
    # Emulation on one device and ONE stream: all ranks scatter first, then every rank reduces its owned tiles (all the
    # flags it waits for are already raised) with done_target = 0, i.e. without parking the stream; the completion
    # counters are checked on the host instead.  (Real ranks run these concurrently on their own devices.)
    i8ws = []
    if int8:
        for r in range(R):
            rows = bounds[r + 1] - bounds[r]
            buf = torch.empty(_abi.lib().ipm_hess_i8_ws_bytes(rows, n, 8), dtype=torch.uint8, device="cuda")
            _abi.call("ipm_hess_i8_prepare", buf.data_ptr(), rows, n, 8, None)
            i8ws.append(buf)
    for step in (1, 2):  # two consecutive "Newton steps": epochs and the running completion counters
        for r in range(R):
            Cr, ldc = Cd[r]
            if int8:
                _abi.call("ipm_hess_i8_scatter_f64", Cr.data_ptr(), ldc, bounds[r + 1] - bounds[r], n, wd[r].data_ptr(), 8,
                          i8ws[r].data_ptr(), p_inbox, p_flags, r, R, slots, step, None)
            else:
                _abi.call("ipm_syrk_scatter_f64", Cr.data_ptr(), ldc, wd[r].data_ptr(), n, bounds[r + 1] - bounds[r], 1.0,
                          None, 0, p_inbox, p_flags, r, R, slots, step, None)
        for r in range(R):
            if int8:  # the owners pull the partial tiles from the buffers of the ranks that computed them
                _abi.call("ipm_hess_reduce_bcast_pull_f64", p_inbox, sig[r].data_ptr(), p_H, p_done, ldh, n, r, R, slots,
                          step, 0, Pd.data_ptr(), ldp, tP, None)
            else:
                _abi.call("ipm_hess_reduce_bcast_f64", inbox[r].data_ptr(), sig[r].data_ptr(), p_H, p_done, ldh, n, r, R,
                          slots, step, 0, Pd.data_ptr(), ldp, tP, None)
        torch.cuda.synchronize()
        for r in range(R):
            assert int(sig[r][slots * R].item()) == step * tiles   # every tile of this step was delivered to rank r
    ref = np.triu((Cm * w[:, None]).T @ Cm + tP * Pm)
    scale = np.abs(Cm * w[:, None]).T @ np.abs(Cm) + np.abs(tP * Pm)
    got0 = H[0].cpu().numpy()
    iu = np.triu_indices(n)
    assert np.max(np.abs(got0[:, :n][iu] - ref[iu]) / scale[iu]) < 1e-13
    il = np.tril_indices(n, -1)
    assert np.all(got0[:, :n][il] == 7.0) and np.all(got0[:, n:] == 7.0)   # strict lower triangle / padding untouched
    for r in range(1, R):
        assert torch.equal(H[r], H[0])                                       # bit-identical on every "rank"


@pytest.mark.parametrize("n,p", [(385, 64), (513, 200), (1000, 129), (2500, 300)])
def test_potrf_trsm_fused_tile_dag(n, p):
    """ipm_potrf_trsm_upper_f64 / the forced tile-DAG entry: H = U'U and Y = U^{-T} B in one launch, ragged last row block
    and ragged last right-hand-side panel, twice in a row (epoch-stamped counters)."""
    import scipy.linalg

    fn = _abi.lib().ipm_internal_potrf_trsm_dag_f64
    fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    H = spd(n, n + 11, cond_pow=2)
    rs = np.random.RandomState(p)
    B = rs.randn(n, p)
    U = scipy.linalg.cholesky(H, lower=False)
    Y = scipy.linalg.solve_triangular(U, B, trans="T", lower=False)
    info = torch.full((1,), -7, dtype=torch.int32, device="cuda")
    for entry in ("forced", "public", "forced"):
        Hd, ld = padded(H)
        Bd, ldb = padded(B)
        guard = Bd.clone()
        if entry == "forced":
            _abi.check(fn(Hd.data_ptr(), ld, n, Bd.data_ptr(), ldb, p, info.data_ptr(), None), "potrf_trsm_dag")
        else:
            _abi.call("ipm_potrf_trsm_upper_f64", Hd.data_ptr(), ld, n, Bd.data_ptr(), ldb, p, info.data_ptr(), None)
        torch.cuda.synchronize()
        assert _abi.lib().ipm_device_fault() == 0 and int(info.item()) == 0
        Ud = torch.triu(Hd[:, :n]).cpu().numpy()
        assert np.max(np.abs(Ud.T @ Ud - H)) / np.max(np.abs(H)) < 1e-13
        Yd = Bd[:, :p].cpu().numpy()
        assert np.max(np.abs(Yd - Y)) <= 1e-10 * np.max(np.abs(Y))
        assert torch.equal(Bd[:, p:], guard[:, p:])  # padding columns untouched


@pytest.mark.parametrize("n,R", [(1000, 2), (1537, 3), (2500, 8)])
def test_potrf_peer_emulated_ranks(n, R):
    """The distributed tile-DAG factorisation (ipm_potrf_upper_peer_f64: block column j on rank j % R, finished rows pushed
    into every rank's copy, system-scope counters) with R emulated ranks on ONE device: a single cooperative grid whose
    CTAs [r * G, (r + 1) * G) act as rank r on rank r's copy of the matrix, counters and info word -- ranks as separate
    launches on one GPU are not guaranteed to be co-resident.  Every copy must end with the whole factor.  (The real
    multi-GPU run is tests/test_sharded_gpu.py.)"""
    L = _abi.lib()
    fn = L.ipm_internal_potrf_peer_emulated_f64
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_uint,
                   C.c_void_p]
    H = spd(n, n + 3, cond_pow=2)
    words = L.ipm_potrf_peer_prog_words()
    copies = [padded(H) for _ in range(R)]
    ld = copies[0][1]
    progs = [torch.zeros(words, dtype=torch.int64, device="cuda") for _ in range(R)]
    infos = [torch.full((4,), 0, dtype=torch.int32, device="cuda") for _ in range(R)]
    pH = (C.c_void_p * R)(*[c[0].data_ptr() for c in copies])
    pP = (C.c_void_p * R)(*[t.data_ptr() for t in progs])
    pI = (C.c_void_p * R)(*[t.data_ptr() for t in infos])
    L.ipm_clear_device_fault()
    for epoch in (1, 2):
        for c in copies:
            c[0][:, :n] = torch.as_tensor(H, device="cuda")
        _abi.check(fn(pH, ld, n, pI, pP, R, epoch, None), "potrf_peer_emulated")
        torch.cuda.synchronize()
        assert L.ipm_device_fault() == 0
        for r in range(R):
            assert int(infos[r][0]) == 0
            U = torch.triu(copies[r][0][:, :n]).cpu().numpy()
            assert np.max(np.abs(U.T @ U - H)) / np.max(np.abs(H)) < 1e-13, r
    # a non-positive pivot is reported to every rank
    Hbad = H.copy()
    Hbad[700, 700] = -1.0
    for c in copies:
        c[0][:, :n] = torch.as_tensor(Hbad, device="cuda")
    for t in infos:
        t.zero_()
    _abi.check(fn(pH, ld, n, pI, pP, R, 3, None), "potrf_peer_emulated")
    torch.cuda.synchronize()
    assert [int(t[0]) for t in infos] == [701] * R
    L.ipm_clear_device_fault()


def _admm_numpy(Qt, bA, eta, rho, alpha, u, add_bias, positive, iters):
    """LassoSolver.py:245-253 / 517-543 restated for the kernel's operands (Q~ = -m rho Q, z = u - alpha)."""
    r = d = a = un_ = None
    for _ in range(iters):
        x = bA + Qt @ (u - alpha)
        v = x + u
        an = np.maximum(v - eta, 0.0)
        if not positive:
            an -= np.maximum(-v - eta, 0.0)
        if add_bias:
            an[0] = v[0]
        un = u + x - an
        r, d, a, un_ = np.sum((x - an) ** 2), np.sum((rho * (an - alpha)) ** 2), np.sum(an ** 2), np.sum(un ** 2)
        alpha, u = an, un
    return alpha, u, np.array([r, d, a, un_])


@pytest.mark.parametrize("n,K,add_bias,positive", [(513, 300, 1, 0), (101, 30, 1, 0), (257, 3900, 1, 1), (64, 17, 0, 0),
                                                   (200, 1000, 0, 0), (1025, 2100, 1, 0)])
def test_lasso_admm_steps_multi_iteration(n, K, add_bias, positive):
    """ipm_lasso_admm_steps_f64 (persistent multi-iteration kernel, csrc/lasso_multi.cu) against NumPy and against the
    one-launch-per-iteration kernel: both tile shapes (1025 x 2100 takes the 128 x 128 tile, the others 64 x 32 -- 257 x 3900
    with two CTAs per SM, the small ones with one), ragged
    rows / columns / contraction tails, tail-row units (513 = 4 * 128 + 1, 257), two launches in a row (odd iteration
    parity), and the device-side stop test."""
    rs = np.random.RandomState(n + K)
    M = rs.randn(n, n)
    Qt = -(M @ M.T) / (n * 4.0)
    Qt = 0.5 * (Qt + Qt.T)
    bA = rs.randn(n, K)
    eta = np.abs(rs.randn(K)) * 0.3
    rho = 0.4
    ld = (K + 15) // 16 * 16
    Qd, ldq = padded(Qt)
    bAd, _ = padded(bA)
    eta_d = dev(eta)
    st = [torch.zeros((n, ld), dtype=torch.float64, device="cuda") for _ in range(4)]  # alpha, u, z0, z1
    L = _abi.lib()
    ws = torch.zeros(L.ipm_lasso_steps_ws_bytes(n, K), dtype=torch.uint8, device="cuda")
    state = ws[:8].view(torch.int32)
    off = L.ipm_lasso_steps_norms_offset(n, K)
    total = 0
    alpha_ref, u_ref = np.zeros((n, K)), np.zeros((n, K))
    for n_iters in (3, 1, 4):
        _abi.call("ipm_lasso_admm_steps_f64", Qd.data_ptr(), ldq, n, K, bAd.data_ptr(), eta_d.data_ptr(), rho,
                  st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr(), st[3].data_ptr(), ld, add_bias, positive, n_iters,
                  1, 0.0, 0.0, ws.data_ptr(), None)
        torch.cuda.synchronize()
        assert L.ipm_device_fault() == 0
        total += n_iters
        alpha_ref, u_ref, norms_ref = _admm_numpy(Qt, bA, eta[None, :], rho, alpha_ref, u_ref, add_bias, positive, n_iters)
        assert state[:2].tolist() == [0, total]
        scale = 1.0 + np.max(np.abs(u_ref))
        assert np.max(np.abs(st[0][:, :K].cpu().numpy() - alpha_ref)) < 1e-11 * scale
        assert np.max(np.abs(st[1][:, :K].cpu().numpy() - u_ref)) < 1e-11 * scale
        z = st[2 + (total & 1)][:, :K].cpu().numpy()
        assert np.max(np.abs(z - (u_ref - alpha_ref))) < 1e-11 * scale
        norms = ws[off:off + 32].view(torch.float64).cpu().numpy()
        np.testing.assert_allclose(norms, norms_ref, rtol=1e-9)
    # a stop test that holds trivially (huge absolute tolerance) raises the flag; the next launch must be a no-op
    _abi.call("ipm_lasso_admm_steps_f64", Qd.data_ptr(), ldq, n, K, bAd.data_ptr(), eta_d.data_ptr(), rho,
              st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr(), st[3].data_ptr(), ld, add_bias, positive, 2, 1, 1e300,
              0.0, ws.data_ptr(), None)
    torch.cuda.synchronize()
    assert state[:2].tolist() == [1, total + 2]
    snap = [t.clone() for t in st]
    _abi.call("ipm_lasso_admm_steps_f64", Qd.data_ptr(), ldq, n, K, bAd.data_ptr(), eta_d.data_ptr(), rho,
              st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr(), st[3].data_ptr(), ld, add_bias, positive, 5, 1, 1e300,
              0.0, ws.data_ptr(), None)
    torch.cuda.synchronize()
    assert state[:2].tolist() == [1, total + 2]
    assert all(torch.equal(a, b) for a, b in zip(st, snap))


def _hess_i8(Cm, ldc, m, n, w, beta, H, ldh, slices):
    nbytes = _abi.lib().ipm_hess_i8_ws_bytes(m, n, slices)
    assert nbytes > 0
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    _abi.call("ipm_hess_i8_prepare", ws.data_ptr(), m, n, slices, None)
    _abi.call("ipm_hess_i8_f64", Cm.data_ptr(), ldc, m, n, w.data_ptr(), beta, H.data_ptr(), ldh, slices, ws.data_ptr(), None)
    torch.cuda.synchronize()
    assert _abi.lib().ipm_device_fault() == 0
    return ws


@pytest.mark.parametrize("n,m,slices,beta,decades", [
    (64, 100, 8, 0.0, 6), (200, 300, 8, 0.0, 16), (1000, 2100, 8, 1.0, 16), (1024, 2048, 8, 0.0, 20),
    (385, 4000, 8, 1.0, 12), (513, 700, 6, 0.0, 8), (2048, 4096, 8, 0.0, 20),
])
def test_hess_i8_matches_fp64(n, m, slices, beta, decades):
    """C' diag(w) C through exact INT8 slice products (tcgen05.mma.kind::i8) against float64: same error class as the
    DMMA kernel -- relative to sum_k |x_ki x_kj| -- for 8 digits, weights spanning up to 20 decades (the late-solve
    regime), ragged n and m, and accumulation into a preloaded H (the QP path)."""
    rs = np.random.RandomState(n + m + slices)
    Cn = rs.rand(m, n) * 4 - 2
    Cn[rs.rand(m, n) < 0.05] = 0.0
    Cn[:, n // 3] = 0.0                                    # a column of zeros: sigma = 0
    wn = 10.0 ** (rs.rand(m) * decades - decades / 2)
    wn[:: 17] = 0.0                                        # rows that do not count
    Cm, ldc = padded(Cn)
    w = dev(wn)
    H0 = np.triu(rs.rand(n, n)) if beta else np.zeros((n, n))
    # only the upper triangle of H is defined on entry (the engine never writes the lower one): NaN below the diagonal
    H, ldh = padded(H0 + np.tril(np.full((n, n), np.nan), -1) if beta else np.full((n, n), np.nan))
    _hess_i8(Cm, ldc, m, n, w, beta, H, ldh, slices)
    X = np.sqrt(wn)[:, None] * Cn
    ref = beta * H0 + X.T @ X
    scale = np.abs(X).T @ np.abs(X) + 1e-300
    iu = np.triu_indices(n)
    err = (np.abs(H[:, :n].cpu().numpy() - ref) / scale)[iu].max()
    assert err <= (1e-14 if slices == 8 else 128.0 ** -(slices - 1)), err
    # and the DMMA kernel on the same operands sits in the same class
    H2, _ = padded(H0)
    _abi.call("ipm_gemm_tn_f64", Cm.data_ptr(), ldc, Cm.data_ptr(), ldc, w.data_ptr(), 1.0, beta, H2.data_ptr(), ldh, n, n, m,
              1, None)
    torch.cuda.synchronize()
    err2 = (np.abs(H2[:, :n].cpu().numpy() - ref) / scale)[iu].max()
    assert err2 <= 1e-14


def test_hess_i8_is_deterministic_and_reuses_its_workspace():
    """Same bits on every call, and a second call with other weights on the same workspace does not see the first."""
    n, m = 300, 1000
    rs = np.random.RandomState(3)
    Cm, ldc = padded(rs.randn(m, n))
    H1, ldh = padded(np.zeros((n, n)))
    H2, _ = padded(np.zeros((n, n)))
    H3, _ = padded(np.zeros((n, n)))
    w1, w2 = dev(rs.rand(m) + 0.1), dev(10.0 ** (rs.rand(m) * 10 - 5))
    ws = _hess_i8(Cm, ldc, m, n, w1, 0.0, H1, ldh, 8)
    for w, H in ((w2, H2), (w1, H3)):
        _abi.call("ipm_hess_i8_f64", Cm.data_ptr(), ldc, m, n, w.data_ptr(), 0.0, H.data_ptr(), ldh, 8, ws.data_ptr(), None)
    torch.cuda.synchronize()
    iu = torch.triu(torch.ones((n, n), dtype=torch.bool, device="cuda"))
    assert torch.equal(H1[:, :n][iu], H3[:, :n][iu])
    assert not torch.equal(H1[:, :n][iu], H2[:, :n][iu])


def test_hess_i8_rejects_bad_arguments():
    L = _abi.lib()
    assert L.ipm_hess_i8_ws_bytes(100, 64, 9) == 0 and L.ipm_hess_i8_ws_bytes(100, 64, 0) == 0
    assert L.ipm_hess_i8_ws_bytes(70000, 64, 8) == 0       # INT32 accumulators: m <= 65408
    assert L.ipm_hess_i8_ws_bytes(100, 64, 8) > 8 * 128 * 128


def test_hess_i8_propagates_nan_like_the_fp64_kernel():
    """A NaN (or negative) weight must not turn into finite garbage: the affected entries of H are NaN, so the
    factorisation reports the failure exactly as it does behind the FP64 kernel."""
    n, m = 200, 300
    rs = np.random.RandomState(5)
    Cm, ldc = padded(rs.randn(m, n))
    wn = rs.rand(m) + 0.1
    wn[7] = np.nan
    wn[11] = -1.0
    H, ldh = padded(np.zeros((n, n)))
    _hess_i8(Cm, ldc, m, n, dev(wn), 0.0, H, ldh, 8)
    assert torch.isnan(torch.triu(H[:, :n])[torch.triu(torch.ones((n, n), dtype=torch.bool, device="cuda"))]).all()
