"""Batched ADMM Lasso -- oracle restatement.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Solves  min_x 1/(2m) ||A x - b_j||^2 + reg_j ||x[1:]||_1  for many columns ``b_j`` / ``reg_j``
sharing one ``A``.  Follows ``LassoSolver.__init__`` (LassoSolver.py:96-222), ``__run_admm``
(:240-337), ``__run_admm_chunks`` (:339-485) and ``prox`` (:517-543).  ``add_bias=False`` crashes
in the reference; here it simply skips the bias column.
"""

import numpy as np
import scipy.linalg


class OracleLasso:
    def __init__(self, A, b, reg, rho=0.4, max_iters=1000, check_stop=10, add_bias=False, normalize_A=False,
                 positive=False, eps_abs=1e-4, eps_rel=3e-2, num_chunks=0):
        self.num_chunks = max(1, num_chunks)  # LassoSolver.py:94
        b = b[:, None] if b.ndim < 2 else b
        self.b = np.array(b)
        self.reg = np.array(reg)
        self.rho, self.max_iters, self.check_stop = rho, max_iters, check_stop
        self.eps_abs, self.eps_rel, self.positive, self.add_bias = eps_abs, eps_rel, positive, add_bias
        self.K = max(self.b.shape[1], len(self.reg))
        A = np.array(A, dtype=float)
        self.m = A.shape[0]
        if normalize_A:
            A = A / A.std(axis=0)  # LassoSolver.py:120-121 (the reference divides in place)
        if add_bias:
            A = np.hstack((np.ones((self.m, 1)), A))  # LassoSolver.py:123-129
        self.A = A
        self.n = A.shape[1]
        AtA = np.matmul(A.T, A)
        L = scipy.linalg.cho_factor(np.diag(np.ones(self.n) * self.m * self.rho) + AtA, overwrite_a=False,
                                    check_finite=False)
        self.Qinv = scipy.linalg.cho_solve(L, np.eye(self.n), overwrite_b=False, check_finite=False)  # :176-188

    def prox(self, v, eta):
        """LassoSolver.py:517-543."""
        x = np.maximum(v - eta, 0)
        if not self.positive:
            x -= np.maximum(-v - eta, 0)
        if self.add_bias:
            x[0] = v[0]
        return x

    def _objective(self, alpha, b, reg):
        """LassoSolver.py:314-325."""
        f = 1 / (2 * self.m) * ((self.A @ alpha - b) ** 2).sum(axis=0)
        xa = alpha if self.positive else np.abs(alpha)
        f += reg * (xa[1:].sum(axis=0) if self.add_bias else xa.sum(axis=0))
        return f

    def _admm(self, b, reg, Qt, bA):
        K = b.shape[1]
        stop_mult = self.eps_abs * np.sqrt(self.n * K)  # LassoSolver.py:200,361
        eta = reg / self.rho
        x = np.zeros((self.n, K))
        alpha = np.zeros((self.n, K))
        u = np.zeros((self.n, K))
        it = 0
        for it in range(self.max_iters):
            x = bA + Qt @ (u - alpha)
            last = alpha
            alpha = self.prox(x + u, eta)
            u = u + x - alpha
            if it % self.check_stop == self.check_stop - 1:  # LassoSolver.py:273-298
                r_norm = np.linalg.norm(x - alpha)
                d_norm = np.linalg.norm(self.rho * (alpha - last))
                tol_p = stop_mult + self.eps_rel * np.linalg.norm(alpha)
                tol_d = stop_mult + self.eps_rel * self.rho * np.linalg.norm(u)
                if r_norm < tol_p and d_norm < tol_d:
                    break
        return alpha, it

    def solve(self):
        """Returns (X[n x K], objective per problem, iterations).  One chunk: ``iteration + 1``
        (LassoSolver.py:335); chunks: list of ``iteration`` per chunk (LassoSolver.py:479)."""
        if self.num_chunks == 1:
            bA = self.Qinv @ (self.A.T @ self.b)
            Qt = self.Qinv * (-self.m * self.rho)  # LassoSolver.py:218
            alpha, it = self._admm(self.b, self.reg, Qt, bA)
            self.X = alpha
            self.solutions = self._objective(alpha, self.b, self.reg)
            self.num_iterations = [it + 1]
            return self.X, self.solutions, it + 1
        X = np.zeros((self.n, self.b.shape[1]))
        sol = np.empty(self.K)
        its = []
        idx = np.arange(self.b.shape[1])
        for i in range(self.num_chunks):
            sel = idx[i :: self.num_chunks]  # LassoSolver.py:349-351
            bi = np.array(self.b[..., sel])
            ri = np.array(self.reg[sel]) if self.reg.ndim > 0 and len(self.reg) > 1 else self.reg
            bA = self.Qinv @ (self.A.T @ bi)
            Qt = self.Qinv * -self.m * self.rho  # LassoSolver.py:391
            alpha, it = self._admm(bi, ri, Qt, bA)
            X[:, sel] = alpha
            sol[sel] = self._objective(alpha, bi, ri)
            its.append(it)
        self.X, self.solutions, self.num_iterations = X, sol, its
        return X, sol, its
