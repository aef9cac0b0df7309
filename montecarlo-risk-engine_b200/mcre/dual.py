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


_EMPTY = np.zeros(0)


class D:
    """value + tangent vector (length n_params, possibly 0).

    Plans without sensitivities (n = 0) are the common case and are lowered on every
    run_simulation() call, so every operator has a value-only fast path that never touches
    numpy (the empty tangent is one shared array)."""

    __slots__ = ("v", "t")

    def __init__(self, v, t=None, n=0):
        self.v = float(v)
        if t is None:
            self.t = _EMPTY if n == 0 else np.zeros(n)
        else:
            self.t = t if isinstance(t, np.ndarray) and t.dtype == np.float64 else np.asarray(t, dtype=np.float64)

    @staticmethod
    def _val(v):
        r = object.__new__(D)
        r.v = v
        r.t = _EMPTY
        return r

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
        if self.t is _EMPTY:
            return D._val(self.v + (o.v if isinstance(o, D) else o))
        o = self._o(o)
        return D(self.v + o.v, self.t + o.t)

    __radd__ = __add__

    def __neg__(self):
        if self.t is _EMPTY:
            return D._val(-self.v)
        return D(-self.v, -self.t)

    def __sub__(self, o):
        if self.t is _EMPTY:
            return D._val(self.v - (o.v if isinstance(o, D) else o))
        o = self._o(o)
        return D(self.v - o.v, self.t - o.t)

    def __rsub__(self, o):
        if self.t is _EMPTY:
            return D._val((o.v if isinstance(o, D) else o) - self.v)
        o = self._o(o)
        return D(o.v - self.v, o.t - self.t)

    def __mul__(self, o):
        if self.t is _EMPTY:
            return D._val(self.v * (o.v if isinstance(o, D) else o))
        o = self._o(o)
        return D(self.v * o.v, self.t * o.v + self.v * o.t)

    __rmul__ = __mul__

    def __truediv__(self, o):
        if self.t is _EMPTY:
            return D._val(self.v / (o.v if isinstance(o, D) else o))
        o = self._o(o)
        q = self.v / o.v
        return D(q, (self.t - q * o.t) / o.v)

    def __rtruediv__(self, o):
        if self.t is _EMPTY:
            return D._val((o.v if isinstance(o, D) else o) / self.v)
        return self._o(o) / self

    def __pow__(self, p):
        if isinstance(p, D):
            return dexp(p * dlog(self))
        if self.t is _EMPTY:
            return D._val(self.v ** p)
        return D(self.v ** p, p * self.v ** (p - 1) * self.t)

    def __float__(self):
        return self.v

    def __repr__(self):
        return f"D({self.v!r}, {self.t!r})"


def dexp(x):
    if not isinstance(x, D):
        return math.exp(x)
    e = math.exp(x.v)
    return D._val(e) if x.t is _EMPTY else D(e, e * x.t)


def dlog(x):
    if not isinstance(x, D):
        return math.log(x)
    return D._val(math.log(x.v)) if x.t is _EMPTY else D(math.log(x.v), x.t / x.v)


def dsqrt(x):
    if not isinstance(x, D):
        return math.sqrt(x)
    s = math.sqrt(x.v)
    if x.t is _EMPTY:
        return D._val(s)
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
