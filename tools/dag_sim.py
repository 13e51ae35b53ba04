"""Timeline model of the tile-DAG Cholesky (csrc/chol.cu, namespace dag): static schedule (task lin -> CTA lin % G, in
order), per-stage costs taken from the per-CTA clock64 breakdown in profiles/potrf_dag_r01f.txt.  Two dependency
granularities: "tile" (the kernel measured in round 1) and "rows32" (branch dag-pipelined: 32-row steps).  The model is
calibrated on the measured kernel (7.64 ms at 8192, 3.30 at 4096, 1.65 at 2048, 44.7 at 16384) and used to predict the
pipelined one, on one GPU and distributed over R GPUs.      python tools/dag_sim.py [n ...]"""
import sys

G = 148
T_KT = 4.53      # us per 32-row k-tile of one CTA (5.35 ms of DMMA per CTA / 1180 k-tiles at n = 8192)
T_FILL = 1.5     # first TMA round trip of a task
T_HOP = 1.5      # release store -> acquire load sees it (+ proxy fence)
T_SUB = 1.0      # A - acc hand-off (global RMW in the measured kernel, shared-memory staging in the pipelined one)
T_POTF2 = 42.0   # diagonal task after the contraction (measured inside the DAG kernel)
T_SOLVE = 25.0   # 128 x 128 tile solve (measured: 336 us / 13.6 tasks)
T_PUB = 0.7
# pipelined steps (us): potf2 step g = pivot block + row panel + store/publish, then the trailing update
P_AB = [3.15 + 1.2 + 1.0] * 3 + [3.15 + 1.0]
P_C = [2.4, 1.3, 0.5, 0.0]
S_STEP = [1.2 + 1.2 + 0.4 + d + 1.0 for d in (3.1, 2.0, 1.0, 0.0)]  # slab load + substitution + syncs + DMMA + publish
T_STAGE = 2.0


T_REMOTE = 3.0   # extra latency of a push over NVLink (remote stores + system-scope release) in the multi-GPU model


def simulate(n, mode, ranks=1):
    """ranks > 1: block column j belongs to rank j % ranks (its tasks run on that rank's G CTAs, dealt round-robin in
    task order); a finished step is pushed to every rank, so a consumer on another rank sees it T_REMOTE later."""
    T = (n + 127) // 128
    tasks = [(i, j) for i in range(T) for j in range(i, T)]
    g = min(G, len(tasks))
    cta_free = [[0.0] * g for _ in range(ranks)]
    count = [0] * ranks
    pub = {}  # (i, j) -> list of 4 publish times (32-row groups); tile mode: all equal
    busy = 0.0

    def seen(tile, q, rank):
        return pub[tile][q] + T_HOP + (T_REMOTE if tile[1] % ranks != rank else 0.0)

    for lin, (i, j) in enumerate(tasks):
        rk = j % ranks
        c = count[rk] % g
        count[rk] += 1
        cta_free_r = cta_free[rk]
        t = cta_free_r[c]
        if i > 0:
            t += T_FILL
            for kt in range(4 * i):
                kb, q = divmod(kt, 4)
                ready = max(seen((kb, i), q, rk), seen((kb, j), q, rk))
                t = max(t, ready) + T_KT
            busy += 4 * i * T_KT
        if mode == "tile":
            t += T_SUB
            if i == j:
                t += T_POTF2 + T_PUB
            else:
                t = max(t, seen((i, i), 3, rk)) + T_SOLVE + T_PUB
            pub[(i, j)] = [t] * 4
        else:
            t += T_STAGE
            times = []
            if i == j:
                for s in range(4):
                    t += P_AB[s]
                    times.append(t)
                    t += P_C[s]
            else:
                for s in range(4):
                    t = max(t, seen((i, i), s, rk)) + S_STEP[s]
                    times.append(t)
            pub[(i, j)] = times
        cta_free_r[c] = t
    total = max(max(r) for r in cta_free)
    return total, busy / g / ranks / total


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [2048, 4096, 8192, 16384]
    print("      n   tile-granular (model)   rows32 (model)    DMMA share tile / rows32")
    for n in sizes:
        a, ua = simulate(n, "tile")
        b, ub = simulate(n, "rows32")
        print(f"{n:7d}   {a / 1e3:10.3f} ms          {b / 1e3:8.3f} ms        {ua:.2f} / {ub:.2f}")
    print("rows32 on R GPUs (block columns dealt cyclically, push model):")
    for n in sizes:
        print(f"{n:7d}   " + "   ".join(f"R={r}: {simulate(n, 'rows32', r)[0] / 1e3:7.3f} ms" for r in (1, 2, 4, 8)))
