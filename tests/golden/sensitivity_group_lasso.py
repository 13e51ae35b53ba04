"""How stable are the reference's OWN per-centering Newton counts on the demo.ipynb group-lasso SOCP?

Runs the CPU oracle (which reproduces the real reference step for step, tests/test_oracle_golden.py) on the
unperturbed problem and on six copies whose P and q are perturbed by 1e-13 relative.  Output recorded on
2026-10-18 (this container, NumPy 2.3 / OpenBLAS):

    unperturbed   [50, 16, 22, 30, 34, 31, 30, 32, 5, 3, 12, 1]
    seed 0        [50, 16, 22, 30, 34, 31, 30, 32, 5, 3, 12, 1]
    seed 1        [50, 16, 22, 30, 34, 31, 30, 32, 5, 3, 19, 1]
    seed 2        [50, 16, 22, 30, 34, 31, 30, 32, 5, 3, 50, 9]
    seed 3        [50, 16, 22, 30, 34, 31, 30, 32, 5, 3, 10, 6]
    seed 4        [50, 16, 22, 30, 34, 31, 30, 32, 5, 3, 33, 1]
    seed 5        [50, 16, 22, 30, 34, 31, 30, 32, 5, 3, 11, 1]

Centering steps 10 and 11 (t = 5.8e10 and 8.6e11) are decided by rounding: the Armijo test compares barrier
objectives of magnitude t*|f0| ~ 1e13..1e14, whose rounding error (t*|f0|*2^-52 ~ 5e-3 .. 8e-2) is far above the
Newton decrement threshold 1e-5 that ends the step.  tests/test_solvers_gpu.py therefore only requires such steps
to terminate within the iteration cap (optimum, iterate and all earlier counts are still held to the full bar).
"""
import json, sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle import solvers as S
import inspect
print([n for n in dir(S) if 'Oracle' in n])
g=json.load(open('/root/repo/tests/golden/socp_group_lasso.json'))
P, q = np.array(g["P"]), np.array(g["q"])
def run(eps_rel, seed):
    rs=np.random.RandomState(seed)
    A, c = [], []
    for i, grp in enumerate(g["groups"][1:]):
        Ai, ci = np.zeros((27, 27)), np.zeros(27)
        Ai[grp, grp] = 1
        ci[i + 19] = 1
        A.append(Ai), c.append(ci)
    Pp = P*(1+eps_rel*rs.randn(*P.shape)); Pp=(Pp+Pp.T)/2
    qp = q*(1+eps_rel*rs.randn(*q.shape))
    s = S.OracleSOCP(P=Pp, q=qp, A=A, b=None, c=c, d=None, lower_bound=None, upper_bound=None, x0=np.array(g["x0"]))
    v = s.solve()
    return v, s.inner_iters
print(run(0,0))
for sd in range(6): print(run(1e-13, sd))
