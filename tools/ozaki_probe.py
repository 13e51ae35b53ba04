"""Would an Ozaki-style FP64 emulation on the INT8 tensor pipe (tcgen05 kind::i8) beat the FP64 DMMA Hessian kernel?

TOOLS ONLY -- nothing here is on the product path.  The tensor part uses torch._int_mm (cuBLASLt INT8 GEMM = a tuned
tcgen05 kernel): if even that is not clearly faster than the DMMA SYRK, a hand-written one cannot be.

Scheme (error-free splitting along the contraction index).  H = X'X with X = diag(sqrt w) C (K x n).  Column i of X is
scaled by a power of two sigma_i >= 2 max_k |X_ki| and cut into s signed-digit slices q_t in [-64, 64] (round to nearest):
    X_ki = sigma_i * sum_t q_t[k, i] 2^{-7(t+1)}     (truncation error < sigma_i 2^{-7s})
    H_ij = sigma_i sigma_j sum_{t+u <= s-1} 2^{-7(t+u+2)} (Q_t' Q_u)_ij
Every Q_t' Q_u is an exact INT8 x INT8 -> INT32 GEMM (|q q'| <= 2^12, K <= 2^17 products).  s(s+1)/2 slice-pair GEMMs.

  python tools/ozaki_probe.py time [n] [m]       cfg-2 shape: slicing + slice-pair GEMMs for s = 6..9, next to the DMMA SYRK
  python tools/ozaki_probe.py accuracy [n]       late-solve weights of a real LP solve: error of the emulated Hessian
                                                 relative to sum_k |x_ki||x_kj| for s = 5..10 (the kernel tests' bar: 1e-14)
  python tools/ozaki_probe.py solve [n]          the n-variable golden LP solved with the emulated Hessian substituted
                                                 for ipm_gemm_tn_f64: Newton counts against the golden for s = 7..10
Prints JSON lines."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import problems  # noqa: E402
from ipm_b200 import _abi  # noqa: E402

PEAK = 37.1


def slices_of(X, s):
    """X: K x n float64 (device).  Returns (Q [s, n, K] int8 -- transposed: both GEMM operands K-major --, sigma [n])."""
    amax = X.abs().amax(dim=0).clamp_min(1e-300)
    sigma = torch.exp2(torch.ceil(torch.log2(amax)) + 1)  # |x / sigma| <= 1/2
    r = (X / sigma).T.contiguous()                        # n x K, |r| <= 1/2
    Q = torch.empty((s,) + r.shape, dtype=torch.int8, device=X.device)
    for t in range(s):
        r = r * 128.0
        q = torch.round(r)                                # signed digit in [-64, 64]; the remainder stays in [-1/2, 1/2]
        Q[t] = q.to(torch.int8)
        r = r - q
    return Q, sigma


def emulated_syrk(X, s, out=None):
    """H = X'X through INT8 slice-pair GEMMs, FP64 accumulation of the (exact) INT32 results, diagonal by diagonal."""
    Q, sigma = slices_of(X, s)
    n = X.shape[1]
    H = torch.zeros((n, n), dtype=torch.float64, device=X.device) if out is None else out.zero_()
    for d in range(s):                                    # pairs with t + u = d share the scale 2^{-7(d+2)}
        acc = None
        for t in range(d + 1):
            g = torch._int_mm(Q[t], Q[d - t].T)            # (n x K) @ (K x n) -> int32, exact
            acc = g if acc is None else acc.add_(g)        # <= 8 * 2^26: no overflow
        H.add_(acc.to(torch.float64), alpha=2.0 ** (-7 * (d + 2)))
    H.mul_(sigma[:, None]).mul_(sigma[None, :])
    return H


def timed(fn, reps=3):
    ts = []
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[1:]))


def dmma_syrk(Cm, w, H):
    m, n = Cm.shape
    _abi.call("ipm_gemm_tn_f64", Cm.data_ptr(), n, Cm.data_ptr(), n, w.data_ptr(), 1.0, 0.0, H.data_ptr(), n, n, n, m, 1,
              None)


def cmd_time(n, m):
    g = torch.Generator(device="cuda").manual_seed(1)
    Cm = torch.rand((m, n), dtype=torch.float64, device="cuda", generator=g) * 4 - 2
    w = 10.0 ** (torch.rand(m, dtype=torch.float64, device="cuda", generator=g) * 16 - 8)
    H = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    out = {"mode": "time", "n": n, "m": m, "dmma_syrk_ms": timed(lambda: dmma_syrk(Cm, w, H))}
    X = Cm * torch.sqrt(w)[:, None]
    Q, _ = slices_of(X, 2)
    a, b = Q[0], Q[1].T
    pair_ms = timed(lambda: torch._int_mm(a, b))
    out["int8_pair_gemm_ms"] = pair_ms
    out["int8_pair_gemm_tops"] = 2.0 * n * n * m / (pair_ms * 1e-3) / 1e12
    del Q, a, b
    for s in (6, 7, 8, 9):
        pairs = s * (s + 1) // 2
        slice_ms = timed(lambda: slices_of(X, s), reps=1)
        out[f"s{s}"] = {
            "slice_pairs": pairs, "slicing_ms (torch ops, unfused)": slice_ms,
            "pair_gemms_full_ms": pairs * pair_ms,
            "pair_gemms_upper_tiles_only_ms": pairs * pair_ms * (n / 128 + 1) / (2 * n / 128),
            "speedup_vs_dmma_if_only_the_gemms_counted (upper tiles)":
                out["dmma_syrk_ms"] / (pairs * pair_ms * (n / 128 + 1) / (2 * n / 128))}
    print(json.dumps(out))


def late_weights(n):
    """Barrier weights 1/s^2 of the LAST Newton step of a real solve (they span the most orders of magnitude there)."""
    from ipm_b200.LPSolver import LPSolver

    prob = problems.lp_dense_family(seed=0, n=n, warm=True)
    s = LPSolver(**prob, check_cvxpy=False, suppress_print=True)
    s.solve()
    return prob, s.ns.ws.w[: s.data.m].clone(), s


def cmd_accuracy(n):
    prob, w, _ = late_weights(n)
    Cm = torch.as_tensor(prob["C"]).cuda()
    m = Cm.shape[0]
    X = Cm * torch.sqrt(w)[:, None]
    H = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    dmma_syrk(Cm, w, H)
    Href = torch.triu(H)
    # error scale: sum_k |x_ki||x_kj|; reference value: float64 GEMM of the explicitly scaled matrix in a different
    # summation order (cuBLAS), so that DMMA's own rounding shows up too
    absum = (X.abs().T @ X.abs())
    Hblas = torch.triu(X.T @ X)
    out = {"mode": "accuracy", "n": n, "m": m, "weight_range_log10": float(torch.log10(w.max() / w.min())),
           "dmma_vs_cublas_fp64": float(((Href - Hblas).abs() / absum).max())}
    for s in (5, 6, 7, 8, 9, 10):
        He = torch.triu(emulated_syrk(X, s))
        out[f"s{s}_vs_cublas_fp64"] = float(((He - Hblas).abs() / absum).max())
    print(json.dumps(out))


def cmd_solve(n):
    """Golden LP (tests/golden/large_cases.json or barrier_cases.json) with the Hessian contraction emulated."""
    from ipm_b200 import engine
    from ipm_b200.LPSolver import LPSolver

    name = f"lp_dense_n{n}_warm"
    cases = {}
    for f in ("barrier_cases.json", "large_cases.json"):
        cases.update({c["name"]: c for c in json.load(open(f"tests/golden/{f}"))})
    case = cases[name]
    orig = engine.LinearNewton._hessian
    out = {"mode": "solve", "case": name, "golden_inner_iters": case["inner_iters"], "golden_value": case["value"]}
    for s in (None, 7, 8, 9, 10):
        def hess(self, t, s=s):
            if s is None or self.phase1 or self.d.m == 0:
                return orig(self, t)
            d, ws = self.d, self.ws
            X = d.C[:, : d.n] * torch.sqrt(ws.w[: d.m])[:, None]
            ws.H[: d.n, : d.n] = emulated_syrk(X, s)
            self.L("ipm_hess_finish_f64", ws.H.data_ptr(), ws.ldh, d.n, ws.hdiag.data_ptr(), None, None, self.shift)
        engine.LinearNewton._hessian = hess
        prob = problems.lp_dense_family(**case["generator_kwargs"])
        sv = LPSolver(**prob, check_cvxpy=False, suppress_print=True)
        val = sv.solve()
        diff = [a - b for a, b in zip(sv.inner_iters, case["inner_iters"])]
        out["dmma" if s is None else f"s{s}"] = {"value_rel_err": abs(val - case["value"]) / abs(case["value"]),
                                                  "inner_iters": sv.inner_iters, "max_count_diff": max(map(abs, diff))}
    engine.LinearNewton._hessian = orig
    print(json.dumps(out))


if __name__ == "__main__":
    _abi.require_device()
    mode = sys.argv[1] if len(sys.argv) > 1 else "time"
    if mode == "time":
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
        cmd_time(n, int(sys.argv[3]) if len(sys.argv) > 3 else 2 * n)
    elif mode == "accuracy":
        cmd_accuracy(int(sys.argv[2]) if len(sys.argv) > 2 else 2048)
    else:
        cmd_solve(int(sys.argv[2]) if len(sys.argv) > 2 else 1024)
