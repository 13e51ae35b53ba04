"""An INDEPENDENT textbook log-barrier centering (Boyd & Vandenberghe, alg. 10.1/11.1) in plain torch float64 --
test infrastructure only.  It shares no code, no kernels and none of the reference's quirks (exact Hessians,
Armijo on g'dx with fresh slacks, relative Newton-decrement stop) with the product path; torch.matmul /
torch.linalg (cuBLAS / cuSOLVER) do the arithmetic.

Used by tests/test_fullsize_gpu.py: follow the central path from the generator's strictly feasible point to a
moderate t; the barrier multipliers of a CENTRED point are dual feasible, so tests/certificates.py turns them into a
rigorous lower bound on the optimum that is (#constraints)/t tight.  The solver under test is not involved.

    minimise  1/2 x'Px + q'x   s.t.  C x <= d,  lo <= x <= hi,  |A_i x + b_i| <= c_i'x + d_i,  E x = e
"""

import torch

F64 = torch.float64


class Barrier:
    def __init__(self, n, P=None, q=None, C=None, d=None, lo=None, hi=None, cones=None, E=None, e=None, identity_P=False):
        self.n, self.P, self.q, self.C, self.d, self.lo, self.hi, self.E, self.e = n, P, q, C, d, lo, hi, E, e
        self.identity_P = identity_P
        if cones is not None:  # (A_stack [ktot x n], b_stack, Cc [M x n], dvec [M], k rows per cone)
            self.As, self.bs, self.Cc, self.dv, self.k = cones
            self.M = self.Cc.shape[0]
        else:
            self.M = 0

    def num_constraints(self):
        return (0 if self.C is None else self.C.shape[0]) + (0 if self.lo is None else 2 * self.n) + 2 * self.M

    def objective(self, x):
        f = x @ self.q if self.q is not None else x.new_zeros(())
        if self.identity_P:
            f = f + 0.5 * (x @ x)
        elif self.P is not None:
            f = f + 0.5 * (x @ (self.P @ x))
        return f

    def objective_grad(self, x):
        g = self.q.clone() if self.q is not None else torch.zeros_like(x)
        if self.identity_P:
            g = g + x
        elif self.P is not None:
            g = g + self.P @ x
        return g

    def slacks(self, x):
        out = {}
        if self.C is not None:
            out["s"] = self.d - self.C @ x
        if self.lo is not None:
            out["sl"], out["su"] = x - self.lo, self.hi - x
        if self.M:
            lhs = (self.As @ x + self.bs).view(self.M, self.k)
            rhs = self.Cc @ x + self.dv
            out["lhs"], out["rhs"], out["sc"] = lhs, rhs, rhs * rhs - (lhs * lhs).sum(1)
        return out

    @staticmethod
    def interior(sl):
        ok = True
        for key in ("s", "sl", "su", "sc", "rhs"):
            if key in sl:
                ok = ok and bool((sl[key] > 0).all())
        return ok

    def value(self, t, x, sl):
        f = t * self.objective(x)
        for key in ("s", "sl", "su", "sc"):
            if key in sl:
                f = f - torch.log(sl[key]).sum()
        return f

    def grad_hess(self, t, x, sl):
        n = self.n
        g = t * self.objective_grad(x)
        H = torch.zeros((n, n), dtype=F64, device=x.device)
        if self.identity_P:
            H.diagonal().add_(t)
        elif self.P is not None:
            H.add_(self.P, alpha=t)
        if "s" in sl:
            inv = 1.0 / sl["s"]
            g = g + self.C.T @ inv
            H.addmm_(self.C.T, self.C * (inv * inv).unsqueeze(1))
        if "sl" in sl:
            g = g + 1.0 / sl["su"] - 1.0 / sl["sl"]
            H.diagonal().add_(1.0 / sl["su"] ** 2 + 1.0 / sl["sl"] ** 2)
        if self.M:
            lhs, rhs, sc = sl["lhs"], sl["rhs"], sl["sc"]
            # grad of -log(rhs^2 - |lhs|^2):  G_i = 2 (A_i' lhs_i - c_i rhs_i) / s_i
            G = (2.0 / sc).unsqueeze(1) * (torch.einsum("mkn,mk->mn", self.As.view(self.M, self.k, n), lhs)
                                           - self.Cc * rhs.unsqueeze(1))
            g = g + G.sum(0)
            # exact Hessian:  G_i G_i' + (2 / s_i) (A_i'A_i - c_i c_i')
            w_rows = (2.0 / sc).repeat_interleave(self.k)
            H.addmm_(self.As.T, self.As * w_rows.unsqueeze(1))
            H.addmm_(self.Cc.T, self.Cc * (-2.0 / sc).unsqueeze(1))
            H.addmm_(G.T, G)
        return g, H


def bracket(bar, t, x, w=None, refine=False):
    """(upper, lower): objective at the strictly feasible x, and the Lagrange dual bound built from the barrier
    multipliers at (x, t):  lam = 1/(t s),  cone pairs (u_i, tau_i) = 2/(t s_i) (lhs_i, rhs_i),  nu = w/t.  Valid for
    ANY x in the interior (the multipliers are non-negative / in the dual cone by construction); (#constraints)/t
    tight when x is centred.
      * box present (LP / QP):  L(x') >= L(x) + h'(x' - x) by convexity, minimised over the box in closed form
        (refine=True: with the least-squares corrected multipliers of _refine_multipliers);
      * cones with P = I, no box:  the dual function in closed form."""
    sl = bar.slacks(x)
    upper = float(bar.objective(x))
    if bar.M:
        assert bar.identity_P and bar.lo is None and bar.C is None and bar.E is None
        lhs, rhs, sc = sl["lhs"], sl["rhs"], sl["sc"]
        u = (2.0 / (t * sc)).unsqueeze(1) * lhs                  # M x k
        tau = 2.0 * rhs / (t * sc)
        assert bool((torch.linalg.norm(u, dim=1) <= tau).all())
        wv = bar.q + torch.einsum("mkn,mk->n", bar.As.view(bar.M, bar.k, bar.n), u) - bar.Cc.T @ tau
        lower = float(-0.5 * (wv @ wv) + (u.reshape(-1) @ bar.bs) - tau @ bar.dv)
        return upper, lower
    lam = 1.0 / (t * sl["s"]) if "s" in sl else None
    nu = w / t if bar.E is not None else None
    if refine:
        lam, nu = _refine_multipliers(bar, x, sl, lam, nu)
    return upper, _box_bound(bar, x, sl, upper, lam, nu)


def _box_bound(bar, x, sl, fx, lam, nu):
    h = bar.objective_grad(x)
    L = fx
    if lam is not None:
        h = h + bar.C.T @ lam
        L -= float(lam @ sl["s"])
    if nu is not None:
        h = h + bar.E.T @ nu
        L += float(nu @ (bar.E @ x - bar.e))
    return L + float(torch.minimum(h * (bar.lo - x), h * (bar.hi - x)).sum())


def _refine_multipliers(bar, x, sl, lam, nu, active_tol=1e-5):
    """Least-squares correction of the multipliers on the active rows (and of nu) so that the reduced cost vanishes
    on the coordinates strictly inside the box.  lam = 1/(t s) carries the relative cancellation error of
    s = d - C x (about 1e-16 |d| / s, i.e. 1e-7 for the active rows at t ~ 1e8), which is what limits the bound;
    any non-negative lam is admissible, so the corrected one (clamped at 0) gives a valid, much tighter bound."""
    inside = (sl["sl"] > active_tol) & (sl["su"] > active_tol)
    cols, na = [], 0
    if lam is not None:
        act = sl["s"] < active_tol
        na = int(act.sum())
        if na:
            cols.append(bar.C[act][:, inside].T)
    if nu is not None:
        cols.append(bar.E[:, inside].T)
    if not cols:
        return lam, nu
    Mx = torch.cat(cols, dim=1).contiguous()
    if Mx.shape[0] < Mx.shape[1]:
        return lam, nu
    h = bar.objective_grad(x)
    if lam is not None:
        h = h + bar.C.T @ lam
    if nu is not None:
        h = h + bar.E.T @ nu
    delta = torch.linalg.lstsq(Mx, -h[inside].unsqueeze(1)).solution.squeeze(1)
    if not bool(torch.isfinite(delta).all()):
        return lam, nu
    if na:
        lam = lam.clone()
        lam[act] = torch.clamp(lam[act] + delta[:na], min=0.0)
    if nu is not None:
        nu = nu + delta[na:]
    return lam, nu


def centre(bar, t, x, max_iters=60, tol=1e-10, alpha=0.25, beta=0.5):
    """Damped Newton centering at parameter t from the strictly feasible x (which must satisfy E x = e when
    equalities are present).  Returns (x, w, decrement) with w the equality multipliers of the t-scaled problem."""
    w = None
    for _ in range(max_iters):
        sl = bar.slacks(x)
        assert bar.interior(sl), "polish started outside the domain"
        g, H = bar.grad_hess(t, x, sl)
        L = torch.linalg.cholesky(H)
        if bar.E is not None:
            # [H E'; E 0] [dx; w] = [-g; 0]  by block elimination
            Z = torch.cholesky_solve(bar.E.T.contiguous(), L)
            y = torch.cholesky_solve(g.unsqueeze(1), L).squeeze(1)
            S = bar.E @ Z
            w = torch.linalg.solve(S, -(bar.E @ y))
            dx = -(y + Z @ w)
            lam2 = float(-(g + bar.E.T @ w) @ dx)
        else:
            dx = -torch.cholesky_solve(g.unsqueeze(1), L).squeeze(1)
            lam2 = float(-(g @ dx))
        if lam2 / 2 <= tol:
            return x, w, lam2 / 2
        f0 = bar.value(t, x, sl)
        a = 1.0
        while True:
            xn = x + a * dx
            sn = bar.slacks(xn)
            # 1e-13 |f0|: rounding floor of the barrier value itself
            if bar.interior(sn) and float(bar.value(t, xn, sn)) <= float(f0) - alpha * a * lam2 + 1e-13 * abs(float(f0)):
                break
            a *= beta
            if a < 1e-10:  # at the rounding floor: take the last feasible point and let the caller judge the bracket
                break
        x = xn
    return x, w, lam2 / 2


def follow_path(bar, x, t0, t_end, mu=15.0, stage_tol=1e-2, final_tol=1e-6, verbose=False):
    """Textbook barrier method from the strictly feasible x: centre at t_end/mu^k, ..., t_end/mu, t_end (k chosen
    so that the first parameter is <= t0).  Returns (x, w, t)."""
    import math

    stages = max(0, math.ceil(math.log(t_end / t0) / math.log(mu)))
    t = t_end / mu ** stages  # land exactly on t_end
    for k in range(stages + 1):
        last = k == stages
        x, w, dec = centre(bar, t, x, tol=final_tol if last else stage_tol)
        if verbose:
            up, lo_ = bracket(bar, t, x, w, refine=last and bar.M == 0)
            print("  polish stage t=%.3g decrement %.2e bracket width %.3e (constraints/t = %.3e)" % (
                t, dec, up - lo_, bar.num_constraints() / t))
        if last:
            return x, w, t
        t *= mu
