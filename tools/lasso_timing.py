"""Where the CTAs of the persistent multi-iteration Lasso kernel spend their time (needs a library built with
-DIPM_LASSO_TIMING: `python interiorpoint-gpu_b200/build.py --variant lassotiming -DIPM_LASSO_TIMING`, selected with
IPM_B200_LIB=.../lib/variants/liblassotiming.so).  Prints, per K, the mean time per unit of every phase.

    IPM_B200_LIB=$PWD/interiorpoint-gpu_b200/lib/variants/liblassotiming.so python tools/lasso_timing.py [K ...]"""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import problems  # noqa: E402
from ipm_b200 import _abi  # noqa: E402
from ipm_b200.LassoSolver import LassoSolver  # noqa: E402

L = _abi.lib()
fn = L.ipm_internal_lasso_timing
fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_int]
buf = (C.c_longlong * (1024 * 16))()
names = ["kernel", "claim", "dependency wait", "TMA + DMMA loop", "rank-1 terms", "epilogue / tail unit", "publish",
         "tile units", "tail units"]
for K in [int(a) for a in sys.argv[1:]] or [4096, 512]:
    A, b, reg = problems.lasso_cfg5(4096)
    b, reg = b[:, ::4096 // K], reg[::4096 // K]
    s = LassoSolver(A, b, reg, rho=0.4, check_stop=10, add_bias=True, check_cvxpy=False, eps_abs=1e-6, eps_rel=1e-6,
                    max_iters=5000)
    s.solve()
    fn(buf, 1024 * 16)  # clear
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, _, _, its = s.solve()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    fn(buf, 1024 * 16)
    t = np.array(buf[:], dtype=np.float64).reshape(1024, 16)
    t = t[t[:, 0] > 0]
    mhz = 1965.0
    units = t[:, 7] + t[:, 8]
    print(f"K={K}: {its} iterations, {ms:.2f} ms = {1e3 * ms / its:.1f} us per iteration; {len(t)} CTAs, "
          f"{units.sum() / its:.0f} units per iteration ({t[:, 7].sum() / its:.0f} tiles + {t[:, 8].sum() / its:.0f} tail)")
    print(f"   kernel time per CTA (sum over launches): mean {t[:, 0].mean() / mhz / 1e3:.2f} ms")
    for k in range(1, 7):
        print(f"   {names[k]:22s} {t[:, k].sum() / units.sum() / mhz:8.2f} us per unit   "
              f"({100 * t[:, k].sum() / t[:, 0].sum():5.1f}% of CTA time)")
