"""Barrier evaluation (objective / slacks / gradient / Hessian) -- oracle restatement.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference keeps a dirty-flag cache (``FunctionManager.py:11-195``); the observable semantics
are "every quantity is evaluated from the current ``(x, slacks, t)``, where ``slacks`` may be
*stale* if the last move asked not to refresh them" (``FunctionManager.py:199-206``).  The oracle
keeps exactly that state and recomputes on demand, with the reference's operation order so the
floating-point results agree bit for bit on the same BLAS.

Slack layout (``FunctionManager.py:50-62``): ``[C rows | upper-bound rows | lower-bound rows]``;
SOCP appends the cone right-hand sides ``c_i^T x + d_i`` for the ``< 0`` feasibility test only
(``FunctionManager.py:921,977-988``).
"""

import numpy as np

LOG_GUARD = 1e-15  # FunctionManager.py:223-227, 244-246
CONE_GUARD = 1e-12  # FunctionManager.py:1084, 1136, 1152


def _layout(m_ineq, n, has_ub, has_lb):
    """Slices into the slack vector (FunctionManager.py:50-62, 410-418, 910-921)."""
    start = m_ineq
    ineq = slice(0, m_ineq)
    ub_sl = lb_sl = None
    if has_ub:
        ub_sl = slice(start, start + n)
        start += n
    if has_lb:
        lb_sl = slice(start, start + n)
        start += n
    return ineq, ub_sl, lb_sl, start


class LinearBarrier:
    """LP (``P is None``) / QP barrier.

    Follows ``FunctionManagerLP`` (``FunctionManager.py:197-356``) and ``FunctionManagerQP``
    (``FunctionManager.py:619-831``).
    """

    def __init__(self, n, c=None, P=None, q=None, C=None, d=None, lb=None, ub=None, t=1.0, try_diag=True):
        self.n = n
        self.P, self.q = P, q
        self.is_qp = P is not None
        if not self.is_qp:
            self.c = np.ones(n) if c is None else c  # FunctionManager.py:76-92
        self.C, self.d, self.lb, self.ub = C, d, lb, ub
        self.t = t
        self.try_diag = try_diag
        self.bounded = lb is not None or ub is not None
        self.constrained = C is not None or self.bounded
        m = 0 if C is None else len(C)
        self.ineq, self.ub_sl, self.lb_sl, self.n_slacks = _layout(m, n, ub is not None, lb is not None)
        self.x = None
        self.slacks = None

    # -- state ---------------------------------------------------------------------------------
    def compute_slacks(self, x):
        """``update_slacks_fxn`` (FunctionManager.py:118-149)."""
        parts = []
        if self.d is not None:
            parts.append(self.d - np.matmul(self.C, x))
        if self.ub is not None:
            parts.append(self.ub - x)
        if self.lb is not None:
            parts.append(x - self.lb)
        if not parts:
            return None
        out = parts[0]
        for p in parts[1:]:
            out = np.append(out, p)
        return out.flatten() if out.ndim > 1 else out

    def move(self, x, refresh_slacks=True):
        """``update_x(x, update_slacks)`` (FunctionManager.py:199-206, 706-713)."""
        self.x = x
        if self.constrained and refresh_slacks:
            self.slacks = self.compute_slacks(x)

    def set_t(self, t):
        self.t = t  # FunctionManager.py:105-116

    # -- values --------------------------------------------------------------------------------
    def objective(self):
        if self.is_qp:  # FunctionManager.py:683-704
            obj = 0
            obj += 1 / 2 * self.x.dot(np.matmul(self.P, self.x))
            if self.q is not None:
                obj += self.q.dot(self.x)
            return obj
        return self.c.dot(self.x)  # FunctionManager.py:151-162

    def barrier_value(self):
        """``newton_objective`` (FunctionManager.py:208-230, 715-739)."""
        val = self.t * self.objective()
        if self.constrained:
            val -= np.log(self.slacks + LOG_GUARD).sum()
        return val

    def _inv_slacks(self):
        return 1 / (self.slacks + LOG_GUARD)

    def gradient(self):
        """FunctionManager.py:232-265 (LP), 741-781 (QP)."""
        inv = self._inv_slacks() if self.constrained else None
        if self.is_qp:
            g = np.matmul(self.P, self.x)
            if self.q is not None:
                g += self.q
            g *= self.t
        else:
            g = self.t * self.c
        if self.lb is not None:
            g -= inv[self.lb_sl]
        if self.ub is not None:
            g += inv[self.ub_sl]
        if self.C is not None:
            g += np.matmul(self.C.T, inv[self.ineq])
        return g

    def hessian(self):
        """FunctionManager.py:267-326 (LP), 783-827 (QP).  Returns a 1-D diagonal for the
        bounds-only LP with ``try_diag`` (FunctionManager.py:283-292)."""
        inv = self._inv_slacks() if self.constrained else None
        if self.is_qp:
            H = self.t * self.P
            if self.C is not None:
                H += np.matmul(self.C.T, (inv[self.ineq] ** 2)[:, None] * self.C)
        elif self.C is None:
            if self.try_diag and self.bounded:
                if self.lb is not None:
                    h = inv[self.lb_sl] ** 2
                    if self.ub is not None:
                        h += inv[self.ub_sl] ** 2
                else:
                    h = inv[self.ub_sl] ** 2
                return h
            H = np.zeros((self.n, self.n))
        else:
            H = np.matmul(self.C.T, (inv[self.ineq] ** 2)[:, None] * self.C)
        if self.bounded:
            diag = np.einsum("ii->i", H)
            if self.lb is not None:
                diag += 1 / (self.slacks[self.lb_sl]) ** 2  # no guard: FunctionManager.py:320-322
            if self.ub is not None:
                diag += 1 / (self.slacks[self.ub_sl]) ** 2
        return H

    def inv_hessian_diag(self):
        """FunctionManager.py:328-356."""
        if self.is_qp or self.C is not None or not self.try_diag:
            raise ValueError("Hessian is not diagonal, cannot use inv hessian function!")
        return 1 / self.hessian()


class PhaseOneLinearBarrier:
    """Phase-I barrier on z = (x, s): minimise s s.t. Cx - d <= s, bounds relaxed by s.

    Follows ``FunctionManagerPhase1`` (``FunctionManager.py:359-616``).
    """

    def __init__(self, C, d, x0, lb=None, ub=None, t=1.0):
        self.C, self.d, self.lb, self.ub = C, d, lb, ub
        self.n = len(x0)
        self.x = x0
        self.t = t
        self.ineq, self.ub_sl, self.lb_sl, self.n_slacks = _layout(len(C), self.n, ub is not None, lb is not None)
        # s0 = 1 - min slack (FunctionManager.py:390-393)
        self.s = 0
        self.slacks = self.compute_slacks(self.x, self.s)
        self.s = -self.slacks.min() + 1
        self.slacks = self.compute_slacks(self.x, self.s)

    def compute_slacks(self, x, s):
        """FunctionManager.py:429-449."""
        out = s + self.d - np.matmul(self.C, x)
        if self.ub is not None:
            out = np.append(out, s + self.ub - x)
        if self.lb is not None:
            out = np.append(out, s + x - self.lb)
        return out.flatten() if out.ndim > 1 else out

    def move(self, z, refresh_slacks=True):
        """FunctionManager.py:451-470."""
        if len(z) == self.n + 1:
            self.x = z[:-1]
            self.s = z[-1]
        elif len(z) == self.n:
            self.x = z
        else:
            raise ValueError("Provided x does not have the right dimensions!")
        if refresh_slacks:
            self.slacks = self.compute_slacks(self.x, self.s)

    def set_t(self, t):
        self.t = t

    def objective(self):
        return self.s  # FunctionManager.py:472-482

    def barrier_value(self):
        """FunctionManager.py:484-507."""
        val = self.t * self.objective()
        val -= np.log(self.slacks + LOG_GUARD).sum()
        return val

    def gradient(self):
        """FunctionManager.py:509-545."""
        inv = 1 / (self.slacks + LOG_GUARD)
        gx = np.matmul(self.C.T, inv[self.ineq])
        if self.lb is not None:
            gx -= inv[self.lb_sl]
        if self.ub is not None:
            gx += inv[self.ub_sl]
        gs = self.t - inv.sum()
        return np.append(gx, gs)

    def hessian(self):
        """Bordered (n+1) Hessian, FunctionManager.py:547-611 (returned as a plain ndarray; the
        reference's CPU arm wraps it in ``np.matrix`` via ``np.bmat``, which LAPACK ignores)."""
        inv2 = (1 / (self.slacks + LOG_GUARD)) ** 2
        Hxx = np.matmul(self.C.T, (inv2[self.ineq])[:, None] * self.C)
        hxs = -np.matmul(self.C.T, inv2[self.ineq])
        diag = np.einsum("ii->i", Hxx)
        if self.lb is not None:
            diag += inv2[self.lb_sl]
            hxs += inv2[self.lb_sl]
        if self.ub is not None:
            diag += inv2[self.ub_sl]
            hxs -= inv2[self.ub_sl]
        hss = inv2.sum()
        return np.block([[Hxx, hxs.reshape(-1, 1)], [hxs.reshape(1, -1), np.array(hss).reshape(1, 1)]])


class ConeBarrier:
    """SOCP barrier: 1/2 x'Px + q'x with cones ||A_i x + b_i|| <= c_i'x + d_i and box bounds.

    Follows ``FunctionManagerSOCP`` (``FunctionManager.py:834-1162``).  ``A[i]`` may be 1-D (a
    compressed diagonal, ``SOCPSolver.py:282-292``).  The reference caches ``A_i^T A_i`` per cone
    (``FunctionManager.py:869-894``); the oracle recomputes it, which is the same arithmetic.
    Quirks kept: ``+ c c^T`` in the Hessian (Q6), 1e-12 guard in gradient/Hessian but 1e-15 in
    the objective (Q5).
    """

    def __init__(self, n, P=None, q=None, A=None, b=None, c=None, d=None, lb=None, ub=None, t=1.0):
        self.n = n
        self.P, self.q, self.A, self.b, self.c, self.d = P, q, A, b, c, d
        self.lb, self.ub = lb, ub
        self.t = t
        self.bounded = lb is not None or ub is not None
        self.ineq, self.ub_sl, self.lb_sl, end = _layout(len(A), n, ub is not None, lb is not None)
        self.constraint_sl = slice(0, end)  # FunctionManager.py:921
        self.AtA = [np.matmul(Ai.T, Ai) if Ai.ndim > 1 else np.diag(Ai**2) for Ai in A]
        self.cct = [np.outer(ci, ci) for ci in c] if c is not None else None
        self.x = None
        self.slacks = self.lhs = self.rhs = None

    def compute_slacks(self, x):
        """FunctionManager.py:933-994.  Returns (slacks, lhs list, rhs list)."""
        lhs = [np.matmul(Ai, x) if Ai.ndim > 1 else Ai * x for Ai in self.A]
        if self.b is not None:
            for i in range(len(lhs)):
                lhs[i] += self.b[i]
        if self.c is not None:
            rhs = [ci.dot(x) for ci in self.c]
            if self.d is not None:
                for i in range(len(rhs)):
                    rhs[i] += self.d[i]
        elif self.d is not None:
            rhs = self.d
        else:
            rhs = 0
        s = np.array([r**2 - (l**2).sum() for r, l in zip(rhs, lhs)])
        if self.ub is not None:
            s = np.append(s, self.ub - x)
        if self.lb is not None:
            s = np.append(s, x - self.lb)
        s = np.append(s, rhs)
        return s, lhs, rhs

    def move(self, x, refresh_slacks=True):
        self.x = x
        if refresh_slacks:
            self.slacks, self.lhs, self.rhs = self.compute_slacks(x)

    def set_t(self, t):
        self.t = t

    def objective(self):
        """FunctionManager.py:996-1018."""
        obj = 0
        if self.P is not None:
            obj += 1 / 2 * self.x.dot(np.matmul(self.P, self.x))
        if self.q is not None:
            obj += self.q.dot(self.x)
        return obj

    def barrier_value(self):
        """FunctionManager.py:1029-1053."""
        val = self.t * self.objective()
        val -= np.log(self.slacks[self.constraint_sl] + LOG_GUARD).sum()
        return val

    def _cone_grad_terms(self, i):
        Ai = self.A[i]
        return np.matmul(Ai.T, self.lhs[i]) if Ai.ndim > 1 else Ai * self.lhs[i]

    def gradient(self):
        """FunctionManager.py:1055-1102."""
        g = 0
        if self.P is not None:
            g = np.matmul(self.P, self.x)
        if self.q is not None:
            g += self.q
        g *= self.t
        for i, (s, rhs, lhs) in enumerate(zip(self.slacks[self.ineq], self.rhs, self.lhs)):
            if self.c is not None:
                g -= 2 * self.c[i] * rhs / (s + CONE_GUARD)
            Ai = self.A[i]
            if Ai.ndim > 1:
                g += 2 * np.matmul(Ai.T, lhs) / (s + CONE_GUARD)
            else:
                g += 2 * Ai * lhs / (s + CONE_GUARD)
        if self.lb is not None:
            g -= 1 / (self.slacks[self.lb_sl] + LOG_GUARD)
        if self.ub is not None:
            g += 1 / (self.slacks[self.ub_sl] + LOG_GUARD)
        return g

    def hessian(self):
        """FunctionManager.py:1104-1158."""
        H = 0
        if self.P is not None:
            H += self.t * self.P
        for i, s in enumerate(self.slacks[self.ineq]):
            sh = 0
            gt = self._cone_grad_terms(i)
            sh += self.AtA[i]
            if self.c is not None:
                sh += self.cct[i]
                gt -= self.c[i] * self.rhs[i]
            sh *= 2 / (s + CONE_GUARD)
            gt *= 2 / (s + CONE_GUARD)
            sh += np.outer(gt, gt)
            H += sh
        if self.bounded:
            diag = np.einsum("ii->i", H)
            if self.lb is not None:
                diag += 1 / (self.slacks[self.lb_sl] + CONE_GUARD) ** 2
            if self.ub is not None:
                diag += 1 / (self.slacks[self.ub_sl] + CONE_GUARD) ** 2
        return H


class PhaseOneConeBarrier(ConeBarrier):
    """SOCP phase-I on z=(x,s): cone and bound slacks get ``+ s`` (FunctionManager.py:1165-1460)."""

    def __init__(self, A, b, c, d, x0, lb=None, ub=None, t=1.0):
        super().__init__(len(x0), None, None, A, b, c, d, lb, ub, t)
        self.x = x0
        self.s = 0  # FunctionManager.py:1232-1235
        self.slacks, self.lhs, self.rhs = self.compute_slacks(self.x)
        self.s = -self.slacks.min() + 1
        self.slacks, self.lhs, self.rhs = self.compute_slacks(self.x)

    def compute_slacks(self, x):
        s, lhs, rhs = super().compute_slacks(x)
        s[self.constraint_sl] += self.s  # FunctionManager.py:1258-1262
        return s, lhs, rhs

    def move(self, z, refresh_slacks=True):
        """FunctionManager.py:1264-1283."""
        if len(z) == self.n + 1:
            self.x = z[:-1]
            self.s = z[-1]
        elif len(z) == self.n:
            self.x = z
        else:
            raise ValueError("Provided x does not have the right dimensions!")
        if refresh_slacks:
            self.slacks, self.lhs, self.rhs = self.compute_slacks(self.x)

    def objective(self):
        return self.s

    def barrier_value(self):
        """FunctionManager.py:1298-1321."""
        val = self.t * self.objective()
        val -= np.log(self.slacks[self.constraint_sl] + LOG_GUARD).sum()
        return val

    def gradient(self):
        """FunctionManager.py:1323-1374."""
        inv = 1 / (self.slacks[self.constraint_sl] + LOG_GUARD)
        gx = 0
        for i, (invs, rhs, lhs) in enumerate(zip(inv[self.ineq], self.rhs, self.lhs)):
            if self.c is not None:
                gx -= 2 * self.c[i] * rhs * invs
            Ai = self.A[i]
            if Ai.ndim > 1:
                gx += 2 * np.matmul(Ai.T, lhs) * invs
            else:
                gx += 2 * Ai * lhs * invs
        if self.lb is not None:
            gx -= inv[self.lb_sl]
        if self.ub is not None:
            gx += inv[self.ub_sl]
        gs = self.t - inv.sum()
        return np.append(gx, gs)

    def hessian(self):
        """FunctionManager.py:1376-1455."""
        inv = 1 / (self.slacks[self.constraint_sl] + LOG_GUARD)
        inv2 = inv**2
        Hxx = 0
        hxs = 0
        for i, invs in enumerate(inv[self.ineq]):
            sh = 0
            gt = self._cone_grad_terms(i)
            sh += self.AtA[i]
            if self.c is not None:
                sh += self.cct[i]
                gt -= self.c[i] * self.rhs[i]
            sh *= 2 * invs
            gt *= 2 * invs
            hxs -= gt * invs
            sh += np.outer(gt, gt)
            Hxx += sh
        if self.bounded:
            diag = np.einsum("ii->i", Hxx)
            if self.lb is not None:
                diag += inv2[self.lb_sl]
                hxs += inv2[self.lb_sl]
            if self.ub is not None:
                diag += inv2[self.ub_sl]
                hxs -= inv2[self.ub_sl]
        hss = inv2.sum()
        return np.block([[Hxx, hxs.reshape(-1, 1)], [hxs.reshape(1, -1), np.array(hss).reshape(1, 1)]])
