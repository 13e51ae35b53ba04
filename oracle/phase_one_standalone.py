"""Stand-alone phase-I (``G x <= h``) -- oracle restatement of ``PhaseOne.PhaseOneSolver``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``PhaseOne.py:18-395``: minimise s s.t. Gx - h <= s with a textbook Armijo search
(slope g.d, alpha 0.2, beta 0.7, ``PhaseOne.py:187-218``), Hessian conditioned by 0.01 I
(``:123-127``) and ``numpy.linalg.solve`` for the direction (``linear_solver="solve"``).
"""

import numpy as np


class OracleStandalonePhaseOne:
    def __init__(self, G, h, mu, x0=None, eps=1e-8, max_iter_interior=200, max_iter_newton=200):
        self.G, self.h, self.mu, self.eps = G, h, mu, eps
        self.max_iter_interior, self.max_iter_newton = max_iter_interior, max_iter_newton
        m, n = G.shape
        self.x = np.ones(n) if x0 is None else x0  # PhaseOne.py:88-93
        self.s = np.max(G @ self.x - h) + 1  # PhaseOne.py:96
        self.warn = False
        self.newton_steps = 0

    def objective(self, x, s, t):
        return t * s - np.sum(np.log(s + self.h - self.G @ x))  # PhaseOne.py:164-185

    def gradient(self, t):
        """PhaseOne.py:240-272."""
        f = self.s + self.h - self.G @ self.x
        gx = np.sum(self.G / f[:, None], axis=0)
        return np.hstack([gx, t - np.sum(1 / f)])

    def hessian(self):
        """PhaseOne.py:274-328."""
        f = self.s + self.h - self.G @ self.x
        n = self.G.shape[1]
        sG = self.G / f[:, None]
        Hxx = sG.T @ sG
        hxs = np.reshape(np.sum(-self.G / f[:, None] ** 2, axis=0), (n, 1))
        hss = np.sum(1 / f**2)
        return np.block([[Hxx, hxs], [hxs.T, hss]])

    def linesearch(self, t, d, g, alpha=0.2, beta=0.7):
        """PhaseOne.py:187-218."""
        step = 1
        while not (np.max(self.G @ (self.x + step * d[:-1]) - self.h) < self.s + step * d[-1]):
            step *= beta
        while self.objective(self.x + step * d[:-1], self.s + step * d[-1], t) > self.objective(
            self.x, self.s, t
        ) + alpha * step * g @ d:
            step *= beta
        return step

    def newton(self, t):
        """PhaseOne.py:109-162."""
        n = self.G.shape[1]
        it = 0
        for it in range(self.max_iter_newton):
            H = self.hessian() + 0.01 * np.eye(n + 1)
            g = self.gradient(t)
            d = np.linalg.solve(H, -g)
            if (-g @ d) / 2 <= self.eps:
                break
            step = self.linesearch(t, d, g)
            self.x = self.x + step * d[:-1]
            self.s = self.s + step * d[-1]
            self.newton_steps += 1
            if self.s < 0:
                break
        return it == self.max_iter_newton - 1

    def solve(self):
        """PhaseOne.py:330-395.  Returns (x, s, warn)."""
        m = self.G.shape[0]
        if np.max(self.G @ self.x - self.h) <= 0:
            self.s = -1
            return self.x, self.s, self.warn
        t = 1
        for _ in range(self.max_iter_interior):
            if self.newton(t):
                self.warn = True
            if m / t <= self.eps:
                break
            if self.s < 0:
                break
            t *= self.mu
        return self.x, self.s, self.warn
