"""NumPy emulation of the index arithmetic of dag::solve_pipelined and dag::potf2_pipelined (dag_pipelined.patch):
every shared-memory index expression and DMMA fragment mapping is transcribed literally, thread by thread, and the result
is checked against SciPy.  It validates the data flow of the kernels, not their synchronisation.
    python tools/experiments/emulate_solve_pipelined.py"""
import numpy as np
from scipy.linalg import cholesky, solve_triangular

NB, LD, THREADS = 128, 132, 256
rs = np.random.RandomState(0)


def dmma_m8n8k4(acc, a, b):
    """acc: per lane (c0, c1) with row = lane >> 2, cols 2 (lane & 3) + {0, 1}; a: per lane A[m = lane >> 2][k = lane & 3];
    b: per lane B[k = lane & 3][n = lane >> 2]."""
    A = np.zeros((8, 4))
    B = np.zeros((4, 8))
    for lane in range(32):
        A[lane >> 2, lane & 3] = a[lane]
        B[lane & 3, lane >> 2] = b[lane]
    C = A @ B
    for lane in range(32):
        acc[lane][0] += C[lane >> 2, 2 * (lane & 3)]
        acc[lane][1] += C[lane >> 2, 2 * (lane & 3) + 1]


def solve_pipelined(U, B):
    """U: 128 x 128 upper (the diagonal tile as published row by row), B: 128 x 128 -> X with U^T X = B."""
    Ps = np.zeros(NB * LD)
    for r in range(NB):
        Ps[r * LD:r * LD + NB] = B[r]
    Us = np.zeros(32 * LD)
    for b0 in range(0, NB, 32):
        # slab load
        Us[:] = np.nan
        for tid in range(THREADS):
            for idx in range(tid, 32 * 64, THREADS):
                r, cc, row = idx >> 6, (idx & 63) * 2, b0 + (idx >> 6)
                vx = vy = 0.0
                if cc + 1 >= row:
                    vx, vy = U[row, cc], U[row, cc + 1]
                Us[r * LD + cc] = vx if cc >= row else 0.0
                Us[r * LD + cc + 1] = vy
        rinv = np.array([1.0 / Us[t * LD + b0 + t] for t in range(32)])
        # substitution, one thread per column
        for c in range(NB):
            v = [Ps[(b0 + l) * LD + c] for l in range(32)]
            for l in range(32):
                x = v[l] * rinv[l]
                v[l] = x
                for r in range(l + 1, 32):
                    v[r] = v[r] - Us[l * LD + b0 + r] * x
            for l in range(32):
                Ps[(b0 + l) * LD + c] = v[l]
        # DMMA update of the rows below
        for wp in range(8):
            bf = [[[Ps[(b0 + 4 * kk + (lane & 3)) * LD + 16 * wp + 8 * cb + (lane >> 2)] for lane in range(32)]
                   for kk in range(8)] for cb in range(2)]
            for r0 in range(b0 + 32, NB, 32):
                for i in range(4):
                    af = [[-Us[(4 * kk + (lane & 3)) * LD + r0 + 8 * i + (lane >> 2)] for lane in range(32)]
                          for kk in range(8)]
                    for cb in range(2):
                        acc = [[Ps[(r0 + 8 * i + (lane >> 2)) * LD + 16 * wp + 8 * cb + 2 * (lane & 3)],
                                Ps[(r0 + 8 * i + (lane >> 2)) * LD + 16 * wp + 8 * cb + 2 * (lane & 3) + 1]]
                               for lane in range(32)]
                        for kk in range(8):
                            dmma_m8n8k4(acc, af[kk], bf[cb][kk])
                        for lane in range(32):
                            base = (r0 + 8 * i + (lane >> 2)) * LD + 16 * wp + 8 * cb + 2 * (lane & 3)
                            Ps[base], Ps[base + 1] = acc[lane]
    return np.array([Ps[r * LD:r * LD + NB] for r in range(NB)])


def stage_tile(A, acc_full, n, row0, col0, diag):
    """stage_tile: (A - acc) inside the matrix, identity padding on diagonal tiles, zeros elsewhere; acc_full is the
    128 x 128 accumulator tile in matrix layout (the fragment layout is row = wm 64 + i 8 + g8, col = wn 32 + jn 8 + 2 l4 + e)."""
    S = np.full((NB, LD), np.nan)
    for warp in range(8):
        wm, wn = warp // 4, warp % 4
        for lane in range(32):
            l4, g8 = lane & 3, lane >> 2
            for i in range(8):
                lr = wm * 64 + i * 8 + g8
                row = row0 + lr
                for jn in range(4):
                    lc = wn * 32 + jn * 8 + 2 * l4
                    col = col0 + lc
                    for e in range(2):
                        inside = row < n and col + e < n and not (diag and col + e < row)
                        if inside:
                            S[lr, lc + e] = A[row, col + e] - acc_full[lr, lc + e]
                        else:
                            S[lr, lc + e] = 1.0 if (diag and lr == lc + e and row >= n) else 0.0
    return S


if __name__ == "__main__":
    M = rs.uniform(-1, 1, (NB, 2 * NB))
    H = M @ M.T + NB * np.eye(NB)
    U = cholesky(H, lower=False)
    B = rs.uniform(-1, 1, (NB, NB))
    X = solve_pipelined(U, B)
    ref = solve_triangular(U, B, trans="T", lower=False)
    print("solve_pipelined   max rel err", np.max(np.abs(X - ref)) / np.max(np.abs(ref)))
    assert np.allclose(X, ref, rtol=1e-10, atol=1e-12)
    # staging: ragged last diagonal tile (n = 128 * 2 + 37), full off-diagonal tile, full diagonal tile
    n = 2 * NB + 37
    A = rs.uniform(-1, 1, (n, n))
    acc = rs.uniform(-1, 1, (NB, NB))
    S = stage_tile(A, acc, n, 2 * NB, 2 * NB, True)
    assert not np.isnan(S[:, :NB]).any()
    for lr in range(NB):
        for lc in range(NB):
            row, col = 2 * NB + lr, 2 * NB + lc
            want = (A[row, col] - acc[lr, lc]) if (row < n and col < n and col >= row) else (1.0 if (lr == lc and row >= n) else 0.0)
            assert S[lr, lc] == want, (lr, lc)
    S = stage_tile(A, acc, n, 0, NB, False)
    assert np.array_equal(S[:, :NB], A[:NB, NB:2 * NB] - acc)
    S = stage_tile(A, acc, n, NB, 2 * NB, False)  # ragged columns
    assert np.array_equal(S[:, :37], A[NB:2 * NB, 2 * NB:] - acc[:, :37]) and np.all(S[:, 37:NB] == 0)
    print("stage_tile        ok")


def store_rows(upper, r_begin, nrows, ncols, vec=True):
    """store_rows<UPPER>: returns the write count of every element of the 128 x 128 global tile."""
    cnt = np.zeros((NB, NB), dtype=int)
    for tid in range(THREADS):
        for idx in range(tid, 32 * 64, THREADS):
            r, c = r_begin + (idx >> 6), (idx & 63) * 2
            if r >= nrows:
                continue
            cmin = r if upper else 0
            if vec and c >= cmin and c + 1 < ncols:
                cnt[r, c] += 1
                cnt[r, c + 1] += 1
            else:
                if cmin <= c < ncols:
                    cnt[r, c] += 1
                if cmin <= c + 1 < ncols:
                    cnt[r, c + 1] += 1
    return cnt


def check_store_rows():
    for upper in (True, False):
        for nrows, ncols in ((128, 128), (37, 37) if upper else (128, 37), (1, 1) if upper else (128, 1)):
            total = np.zeros((NB, NB), dtype=int)
            for r_begin in range(0, NB, 32):
                total += store_rows(upper, r_begin, nrows, ncols)
            want = np.zeros((NB, NB), dtype=int)
            for r in range(min(nrows, NB)):
                for c in range(ncols):
                    if not upper or c >= r:
                        want[r, c] = 1
            assert np.array_equal(total, want), (upper, nrows, ncols)
    print("store_rows        ok")


if __name__ == "__main__":
    check_store_rows()
