"""Scalar forward-mode dual numbers for the plan compiler.

Every plan constant that depends on a model parameter (bond-price coefficients,
CIR++ shift, Cholesky entries, ...) is computed on the host once as
``value + tangents`` with respect to the flattened parameter list of the model
(reference parameter order: SURVEY Appendix E, src/models/*.py ``model_params``).
The kernels consume them as ``Dual<NT>`` constants, so first-order pathwise
sensitivities need no tape (replaces torch.autograd on the hot path,
reference: src/controller/controller.py:609-627).
"""
from __future__ import annotations

import math
import numpy as np


class D:
    """value + tangent vector (length n_params, possibly 0)."""

    __slots__ = ("v", "t")

    def __init__(self, v, t=None, n=0):
        self.v = float(v)
        self.t = np.zeros(n) if t is None else np.asarray(t, dtype=np.float64)

    # -- helpers -----------------------------------------------------------
    @staticmethod
    def const(v, n):
        return D(v, None, n)

    @staticmethod
    def var(v, idx, n):
        t = np.zeros(n)
        if n > 0:
            t[idx] = 1.0
        return D(v, t)

    @staticmethod
    def lift(x, n):
        return x if isinstance(x, D) else D(x, None, n)

    def _o(self, o):
        return o if isinstance(o, D) else D(o, None, self.t.shape[0])

    def pack(self):
        return np.concatenate(([self.v], self.t))

    # -- arithmetic --------------------------------------------------------
    def __add__(self, o):
        o = self._o(o)
        return D(self.v + o.v, self.t + o.t)

    __radd__ = __add__

    def __neg__(self):
        return D(-self.v, -self.t)

    def __sub__(self, o):
        o = self._o(o)
        return D(self.v - o.v, self.t - o.t)

    def __rsub__(self, o):
        o = self._o(o)
        return D(o.v - self.v, o.t - self.t)

    def __mul__(self, o):
        o = self._o(o)
        return D(self.v * o.v, self.t * o.v + self.v * o.t)

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = self._o(o)
        q = self.v / o.v
        return D(q, (self.t - q * o.t) / o.v)

    def __rtruediv__(self, o):
        return self._o(o) / self

    def __pow__(self, p):
        if isinstance(p, D):
            return dexp(p * dlog(self))
        return D(self.v ** p, p * self.v ** (p - 1) * self.t)

    def __float__(self):
        return self.v

    def __repr__(self):
        return f"D({self.v!r}, {self.t!r})"


def dexp(x):
    if not isinstance(x, D):
        return math.exp(x)
    e = math.exp(x.v)
    return D(e, e * x.t)


def dlog(x):
    if not isinstance(x, D):
        return math.log(x)
    return D(math.log(x.v), x.t / x.v)


def dsqrt(x):
    if not isinstance(x, D):
        return math.sqrt(x)
    s = math.sqrt(x.v)
    return D(s, x.t / (2.0 * s) if s > 0.0 else np.zeros_like(x.t))


def dval(x):
    return x.v if isinstance(x, D) else float(x)


def cholesky_dual(a):
    """Lower Cholesky factor of a symmetric matrix of ``D`` entries.

    Same recurrence the tangent of torch.linalg.cholesky satisfies
    (reference: src/models/model.py:50-73)."""
    n = len(a)
    L = [[None] * n for _ in range(n)]
    nt = a[0][0].t.shape[0]
    for i in range(n):
        for j in range(i + 1):
            s = a[i][j]
            for k in range(j):
                s = s - L[i][k] * L[j][k]
            if i == j:
                L[i][j] = dsqrt(s)
            else:
                L[i][j] = s / L[j][j]
        for j in range(i + 1, n):
            L[i][j] = D(0.0, None, nt)
    return L
