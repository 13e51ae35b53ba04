"""Where does the INT8 tensor-core Hessian (ipm_hess_i8_f64, 8 digits) overtake the FP64 DMMA kernel (ipm_gemm_tn_f64)?
Times both through the C ABI over a ladder of operand sizes (m = 2n, the cfg-2 aspect) and prints one JSON line per size;
engine.HESS_I8_MIN_N is set from this table (profiles/hess_i8_sizes_r02.jsonl).  Tools only."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ipm_b200 import _abi  # noqa: E402


def timed(fn, reps=5):
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


def main():
    _abi.require_device()
    if len(sys.argv) > 2:
        sizes_arg = [(int(sys.argv[1]), int(sys.argv[2]))]
    else:
        sizes_arg = None
    sizes = sizes_arg or [(1024, 2048), (2048, 4096), (3072, 6144), (4096, 8192), (4096, 16384), (6144, 12288), (8192, 16384),
             (8192, 32768), (12288, 24576)]
    for n, m in sizes:
        g = torch.Generator(device="cuda").manual_seed(n)
        Cm = torch.rand((m, n), dtype=torch.float64, device="cuda", generator=g) * 4 - 2
        w = 10.0 ** (torch.rand(m, dtype=torch.float64, device="cuda", generator=g) * 16 - 8)
        H = torch.zeros((n, n), dtype=torch.float64, device="cuda")
        t_dmma = timed(lambda: _abi.call("ipm_gemm_tn_f64", Cm.data_ptr(), n, Cm.data_ptr(), n, w.data_ptr(), 1.0, 0.0,
                                         H.data_ptr(), n, n, n, m, 1, None))
        out = {"n": n, "m": m, "dmma_ms": t_dmma}
        for s in (8, 7):
            ws = torch.empty(_abi.lib().ipm_hess_i8_ws_bytes(m, n, s), dtype=torch.uint8, device="cuda")
            _abi.call("ipm_hess_i8_prepare", ws.data_ptr(), m, n, s, None)
            t = timed(lambda: _abi.call("ipm_hess_i8_f64", Cm.data_ptr(), n, m, n, w.data_ptr(), 0.0, H.data_ptr(), n, s,
                                        ws.data_ptr(), None))
            out[f"i8_s{s}_ms"] = t
            out[f"speedup_s{s}"] = t_dmma / t
            del ws
        print(json.dumps(out), flush=True)
        del Cm, w, H
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
