"""Damped Newton centering with the reference's backtracking searches -- oracle restatement.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``FeasibleNewton``   follows ``NewtonSolver.solve`` / ``backtrack_search`` (NewtonSolver.py:80-206)
                     with the Cholesky (``NewtonSolver.py:250-341``) and diagonal
                     (``NewtonSolver.py:403-420``) linear solves.
``InfeasibleNewton`` follows ``NewtonSolverInfeasibleStart.solve`` / ``backtrack_search``
                     (NewtonSolverInfeasibleStart.py:72-273) with the block-elimination Cholesky
                     solve (``:356-538``) and its diagonal variant (``:757-809``).

Quirks Q1-Q4 of SURVEY.md section 8(a) are reproduced deliberately: the Armijo slope is g.x, the
barrier term is frozen during the Armijo loop, the evaluated point lags the step by one beta,
and the infeasible-start search re-uses stale reciprocal slacks.
"""

import numpy as np
import scipy.linalg
import scipy.sparse.linalg

STUCK = 1e-13  # NewtonSolver.py:176,189 ; NewtonSolverInfeasibleStart.py:186,242


class FeasibleNewton:
    def __init__(self, barrier, max_iters=50, epsilon=1e-5, alpha=0.2, beta=0.6, phase1_flag=False,
                 phase1_tol=0.1, use_psd_condition=False, update_slacks_every=0, diagonal=False,
                 trace=None, linear_solver="cholesky", max_cg_iters=50):
        self.fm = barrier
        self.max_iters, self.eps = max_iters, epsilon
        self.alpha, self.beta = alpha, beta
        self.phase1_flag, self.phase1_tol = phase1_flag, phase1_tol
        self.use_psd_condition = use_psd_condition
        self.update_slacks_every = update_slacks_every
        self.diagonal = diagonal
        self.use_backup = False  # sticky lstsq fallback, NewtonSolver.py:314-341
        self.trace = trace  # optional list collecting (step_size, decrement) per Newton step
        self.linear_solver, self.max_cg_iters = linear_solver, max_cg_iters  # "cg": NewtonSolverCG, NewtonSolver.py:365-400

    # -- linear solve ------------------------------------------------------------------------------
    def direction(self, g, x=None):
        if self.diagonal:  # NewtonSolver.py:415-420
            return -self.fm.inv_hessian_diag() * g
        H = self.fm.hessian()
        if self.linear_solver == "cg":  # NewtonSolver.py:374-400
            dc = np.dot(x, g)
            x0 = -dc * x / np.dot(x, np.dot(H, x)) if dc < 0 else np.zeros_like(x)
            return scipy.sparse.linalg.cg(-H, g, x0=x0, maxiter=self.max_cg_iters)[0]
        if not self.use_backup:
            try:
                if self.use_psd_condition:  # NewtonSolver.py:269-275
                    np.einsum("ii->i", H)[...] += 1e-9
                L, low = scipy.linalg.cho_factor(H, overwrite_a=True, check_finite=False)
                return scipy.linalg.cho_solve((L, low), -g, overwrite_b=True, check_finite=False)
            except np.linalg.LinAlgError:
                self.use_backup = True
        return np.linalg.lstsq(H, -g, rcond=None)[0]  # NewtonSolver.py:334-341

    # -- line search -------------------------------------------------------------------------------
    def backtrack(self, x, dx, g):
        """NewtonSolver.py:157-206."""
        fm = self.fm
        a = 1
        fx = fm.barrier_value()
        nxt = x + a * dx
        slope = g.dot(x)  # Q1: g.x, not g.dx
        fm.move(nxt)
        while (fm.slacks < 0).any():  # Q5: strict
            a *= self.beta
            if a < STUCK:
                return a
            nxt = x + a * dx
            fm.move(nxt)
        attempt = 0
        while fm.barrier_value() > fx + self.alpha * a * slope:
            attempt += 1
            nxt = x + a * dx  # Q3: built with the old step ...
            if a < STUCK:
                return a
            a *= self.beta  # ... tested next round against the new one
            refresh = False
            if self.update_slacks_every > 0:
                refresh = attempt % self.update_slacks_every == self.update_slacks_every - 1
            fm.move(nxt, refresh_slacks=refresh)  # Q2: stale barrier term
        fm.move(nxt)
        return a

    # -- centering loop ----------------------------------------------------------------------------
    def solve(self, x):
        """Returns (x, iters, decrement, success).  NewtonSolver.py:80-155."""
        fm = self.fm
        nd = None
        it = 0
        try:
            for it in range(self.max_iters):
                fm.move(x)
                g = fm.gradient()
                dx = self.direction(g, x)
                a = self.backtrack(x, dx, g)
                x += a * dx
                fm.move(x)
                nd = -g.dot(dx) / 2
                if self.trace is not None:
                    self.trace.append((float(a), float(nd)))
                if self.phase1_flag and x[-1] < -self.phase1_tol:  # NewtonSolver.py:105-107
                    return x, it + 1, None, True
                if a < STUCK:
                    return x, it + 1, nd, False
                elif nd < self.eps:
                    return x, it + 1, nd, True
            return x, it + 1, nd, False
        except np.linalg.LinAlgError:
            return x, it + 1, nd, False


class InfeasibleNewton:
    def __init__(self, barrier, A, b, max_iters=50, epsilon=1e-5, alpha=0.2, beta=0.6,
                 use_psd_condition=False, update_slacks_every=0, diagonal=False, trace=None):
        self.fm, self.A, self.b = barrier, A, b
        self.max_iters, self.eps = max_iters, epsilon
        self.alpha, self.beta = alpha, beta
        self.use_psd_condition = use_psd_condition
        self.update_slacks_every = update_slacks_every
        self.diagonal = diagonal
        self.use_backup = False
        self.trace = trace

    def direction(self, x, v, g):
        A = self.A
        if self.diagonal:  # NewtonSolverInfeasibleStart.py:774-809
            b2 = A @ x - self.b
            hinv = self.fm.inv_hessian_diag()
            L, low = scipy.linalg.cho_factor(A @ (hinv[:, None] * A.T), overwrite_a=False, check_finite=False)
            w = scipy.linalg.cho_solve((L, low), b2 - A @ (hinv * g), overwrite_b=False, check_finite=False)
            return -hinv * (g + A.T @ w), w - v
        H = self.fm.hessian()
        if H.ndim < 2:
            H = np.diag(H)
        b2 = np.matmul(A, x) - self.b
        if not self.use_backup:
            try:  # NewtonSolverInfeasibleStart.py:386-490
                if self.use_psd_condition:
                    np.einsum("ii->i", H)[...] += 1e-9
                L1 = scipy.linalg.cho_factor(H, overwrite_a=False, check_finite=False)
                Z = scipy.linalg.cho_solve(L1, A.T, overwrite_b=False, check_finite=False)
                y = scipy.linalg.cho_solve(L1, g, overwrite_b=False, check_finite=False)
                L, low = scipy.linalg.cho_factor(np.matmul(A, Z), overwrite_a=False, check_finite=False)
                w = scipy.linalg.cho_solve((L, low), b2 - np.matmul(A, y), overwrite_b=False, check_finite=False)
                dx = -scipy.linalg.cho_solve(L1, g + np.matmul(A.T, w), overwrite_b=False, check_finite=False)
                return dx, w - v
            except np.linalg.LinAlgError:
                self.use_backup = True
        # NewtonSolverInfeasibleStart.py:513-538
        Z = np.linalg.solve(H, A.T)
        y = np.linalg.solve(H, g)
        w = np.linalg.solve(np.matmul(A, Z), b2 - np.matmul(A, y))
        dx = -np.linalg.solve(H, g + np.matmul(A.T, w))
        return dx, w - v

    def backtrack(self, x, v, dx, dv, g):
        """NewtonSolverInfeasibleStart.py:170-273.  Returns (step, gradient at trial, residual norm)."""
        fm, A = self.fm, self.A
        a = 1
        nxt = x + a * dx
        fm.move(nxt)
        while (fm.slacks < 0).any():
            a *= self.beta
            if a < STUCK:
                return a, None, None
            nxt = x + a * dx
            fm.move(nxt)
        ATv = np.matmul(A.T, v)
        ATdv = np.matmul(A.T, dv)
        Axb = np.matmul(A, x) - self.b
        Adx = np.matmul(A, dx)
        r_norm = np.linalg.norm(np.append(g + ATv, Axb))
        g_next = fm.gradient()
        next_norm = np.linalg.norm(np.append(g_next + ATv + a * ATdv, Axb + a * Adx))
        attempt = 0
        while next_norm > (1 - self.alpha * a) * r_norm:
            attempt += 1
            a *= self.beta
            if a < STUCK:
                break
            nxt = x + a * dx
            refresh = False
            if self.update_slacks_every > 0:
                refresh = attempt % self.update_slacks_every == self.update_slacks_every - 1
            fm.move(nxt, refresh_slacks=refresh)  # Q4: stale reciprocal slacks
            g_next = fm.gradient()
            next_norm = np.linalg.norm(np.append(g_next + ATv + a * ATdv, Axb + a * Adx))
        fm.move(nxt)
        return a, g_next, next_norm

    def solve(self, x, v0=None):
        """Returns (x, v, iters, residual_norm, success).  NewtonSolverInfeasibleStart.py:72-168."""
        fm = self.fm
        v = np.zeros(self.A.shape[0]) if v0 is None else v0
        r_norm = None
        it = 0
        try:
            for it in range(self.max_iters):
                fm.move(x)
                g = fm.gradient()
                dx, dv = self.direction(x, v, g)
                a, _, r_norm = self.backtrack(x, v, dx, dv, g)
                x += a * dx
                v += a * dv
                fm.move(x)
                if self.trace is not None:
                    self.trace.append((float(a), None if r_norm is None else float(r_norm)))
                if a < STUCK:
                    return x, v, it + 1, r_norm, False
                elif r_norm < self.eps:
                    return x, v, it + 1, r_norm, True
            return x, v, it + 1, r_norm, False
        except np.linalg.LinAlgError:
            return x, v, it + 1, r_norm, False
