"""Runs tools/umma_probe.cu: INT8 UMMA issue rates (SS vs TS operands) and the TMEM layout check of a TS A operand.
Prints JSON lines (profiles/umma_probe_r02.jsonl).  Tools only."""
import ctypes as C
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = C.CDLL(os.path.join(ROOT, "tools", "_build", "libumma_probe.so"))

rs = np.random.RandomState(0)
A = rs.randint(-64, 65, size=(128, 32)).astype(np.int8)
B = rs.randint(-64, 65, size=(64, 32)).astype(np.int8)
ref = A.astype(np.int64) @ B.astype(np.int64).T
Ad, Bd = torch.as_tensor(A).cuda(), torch.as_tensor(B).cuda()
for use_ss in (1, 0):
    D = torch.full((128, 64), -7, dtype=torch.int32, device="cuda")
    rc = L.umma_probe_ts(C.c_void_p(Ad.data_ptr()), C.c_void_p(Bd.data_ptr()), C.c_void_p(D.data_ptr()), use_ss)
    got = D.cpu().numpy().astype(np.int64)
    print(json.dumps({"check": "SS reference path" if use_ss else "TS: A from TMEM (lane = row, column c = bytes 4c..4c+3)",
                      "rc": rc, "exact_fraction": float((got == ref).mean()),
                      "rows_exact": int((got == ref).all(axis=1).sum())}), flush=True)

out = torch.zeros(148, dtype=torch.float32, device="cuda")
for mode in (0, 1):
    for N in (64, 128, 192, 256):
        for ctas in (1, 148):
            rc = L.umma_probe_rate(mode, N, 8192, 4, C.c_void_p(out.data_ptr()), ctas)
            v = out[:ctas].cpu().numpy()
            ideal = N / 2.0
            print(json.dumps({"mode": "SS" if mode == 0 else "TS", "N": N, "ctas": ctas, "rc": rc,
                              "cycles_per_mma_mean": float(v.mean()), "max": float(v.max()), "ideal_cycles": ideal,
                              "pipe_fraction": ideal / float(v.mean()),
                              "smem_bytes_per_clk": (32.0 * N + (4096 if mode == 0 else 0)) / float(v.mean())}), flush=True)
