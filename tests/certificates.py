"""Feasibility / objective of a returned point for the full-size BASELINE configurations, evaluated independently
of the product (torch / cuBLAS float64), plus Lasso duality gaps.

At n = 8192 / 16384 the CPU oracle needs minutes to hours per solve, so parity at full size is established through
convex duality: a strictly feasible point gives an UPPER bound on the optimum p*, a dual-feasible point a LOWER
bound.  The lower bound comes from an independent textbook barrier solve (tests/barrier_polish.py: `bracket`), not
from the product's own final iterate: the reference's last centering steps stop early (its Armijo test uses g'x, its
cone Hessian has +cc'), so barrier multipliers taken at ITS final point are valid but loose (LP
cfg-2: 1.6 %).
"""

import numpy as np
import torch

F64 = torch.float64


def _dev(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64)).to("cuda")


def lp_point(prob, x):
    """min c'x  s.t.  C x <= d,  lb <= x <= ub."""
    xd, c, C, d = _dev(x), _dev(prob["c"]), _dev(prob["C"]), _dev(prob["d"])
    s = d - C @ xd
    min_slack = float(torch.min(torch.cat([s, float(prob["upper_bound"]) - xd, xd - float(prob["lower_bound"])])))
    return dict(primal=float(c @ xd), min_slack=min_slack)


def qp_point(prob, x):
    """min 1/2 x'Px + q'x  s.t.  A x = b,  C x <= d,  lb <= x <= ub."""
    xd, P, q = _dev(x), _dev(prob["P"]), _dev(prob["q"])
    A, b, C, d = _dev(prob["A"]), _dev(prob["b"]), _dev(prob["C"]), _dev(prob["d"])
    s = d - C @ xd
    min_slack = float(torch.min(torch.cat([s, float(prob["upper_bound"]) - xd, xd - float(prob["lower_bound"])])))
    return dict(primal=float(0.5 * (xd @ (P @ xd)) + q @ xd), min_slack=min_slack,
                eq_residual=float(torch.linalg.norm(A @ xd - b)))


def socp_point(prob, x):
    """min 1/2 |x|^2 + q'x  s.t.  |A_i x + b_i| <= c_i'x + d_i."""
    xd, q = _dev(x), _dev(prob["q"])
    min_slack = np.inf
    for Ai, bi, ci, di in zip(prob["A"], prob["b"], prob["c"], prob["d"]):
        lhs = _dev(Ai) @ xd + _dev(bi)
        min_slack = min(min_slack, float(_dev(ci) @ xd) + di - float(torch.linalg.norm(lhs)))
    return dict(primal=float(0.5 * (xd @ xd) + q @ xd), min_slack=min_slack)


def lasso_gaps(A, b, reg, X, add_bias=True):
    """Duality gaps of  min_x 1/(2m) |A x - b_k|^2 + reg_k |x_pen|_1  (bias column unpenalised), one per column of
    b: the dual point is the scaled (and, with a bias, centred) residual.  Returns (primal[K], gap[K])."""
    Ad, bd, Xd, rd = _dev(A), _dev(b), _dev(X), _dev(reg)
    m = Ad.shape[0]
    if add_bias:
        Ad = torch.cat([torch.ones((m, 1), dtype=F64, device="cuda"), Ad], dim=1)
    R = bd - Ad @ Xd                                   # m x K
    pen = Xd[1:] if add_bias else Xd
    primal = (R * R).sum(0) / (2.0 * m) + rd * pen.abs().sum(0)
    rho = R / m
    if add_bias:
        rho = rho - rho.mean(0, keepdim=True)          # 1'theta = 0
    corr = (Ad[:, 1:] if add_bias else Ad).T @ rho     # n x K
    scale = torch.clamp(rd / corr.abs().amax(0), max=1.0)
    theta = rho * scale
    dual = -(m / 2.0) * (theta * theta).sum(0) + (theta * bd).sum(0)
    return primal.cpu().numpy(), (primal - dual).cpu().numpy()
