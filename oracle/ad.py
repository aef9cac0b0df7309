"""Forward-mode AD on numpy arrays for the oracle's Greeks.

A ``Dual`` carries a value array and one tangent array per model parameter
(shape (P,) + value.shape).  The oracle's formulas are written against the functions
below, so the same code yields values (plain ndarrays) or values + pathwise
sensitivities - the latter must agree with the reference's torch.autograd output
(src/controller/controller.py:609-627)."""
import numpy as np


def _outer(ta, tb):
    """t_i (x) t_j over the two leading parameter axes."""
    return ta[:, None] * tb[None, :]


class Dual:
    """value, tangents (P,) + shape and - for the reference's double backward (controller.py:631-648) - optionally the
    second derivatives h of shape (P, P) + shape.  Piecewise-linear functions (relu, clamp, where) mask h like t: zero
    curvature, as torch's double backward of those functions."""
    __array_priority__ = 1000

    def __init__(self, v, t, h=None):
        self.v = np.asarray(v, dtype=np.float64)
        self.t = np.asarray(t, dtype=np.float64)
        self.h = None if h is None else np.asarray(h, dtype=np.float64)

    @property
    def P(self):
        return self.t.shape[0]

    def _lift(self, o):
        if isinstance(o, Dual):
            return o
        o = np.asarray(o, dtype=np.float64)
        return Dual(o, np.zeros((self.P,) + o.shape))

    def _pair(self, o):
        """(other, my tangent, other tangent) with tangents reshaped so that the value
        dimensions broadcast from the right (the leading axis is the parameter axis)."""
        o = self._lift(o)
        nd = max(self.v.ndim, o.v.ndim)
        ta = self.t.reshape((self.P,) + (1,) * (nd - self.v.ndim) + self.v.shape)
        tb = o.t.reshape((o.P,) + (1,) * (nd - o.v.ndim) + o.v.shape)
        return o, ta, tb

    def _pair_h(self, o):
        """second derivatives of both operands, broadcast-ready (zeros for a first-order operand); None if neither has any"""
        if self.h is None and o.h is None:
            return None, None
        nd = max(self.v.ndim, o.v.ndim)
        P = self.P

        def shaped(x):
            if x.h is None:
                return np.zeros((P, P) + (1,) * nd)
            return x.h.reshape((P, P) + (1,) * (nd - x.v.ndim) + x.v.shape)
        return shaped(self), shaped(o)

    def __add__(self, o):
        o, ta, tb = self._pair(o)
        ha, hb = self._pair_h(o)
        return Dual(self.v + o.v, ta + tb, None if ha is None else ha + hb)

    __radd__ = __add__

    def __neg__(self):
        return Dual(-self.v, -self.t, None if self.h is None else -self.h)

    def __sub__(self, o):
        o, ta, tb = self._pair(o)
        ha, hb = self._pair_h(o)
        return Dual(self.v - o.v, ta - tb, None if ha is None else ha - hb)

    def __rsub__(self, o):
        o, ta, tb = self._pair(o)
        ha, hb = self._pair_h(o)
        return Dual(o.v - self.v, tb - ta, None if ha is None else hb - ha)

    def __mul__(self, o):
        o, ta, tb = self._pair(o)
        ha, hb = self._pair_h(o)
        h = None if ha is None else ha * o.v + self.v * hb + _outer(ta, tb) + _outer(tb, ta)
        return Dual(self.v * o.v, ta * o.v + self.v * tb, h)

    __rmul__ = __mul__

    def __truediv__(self, o):
        o, ta, tb = self._pair(o)
        ha, hb = self._pair_h(o)
        q = self.v / o.v
        h = None
        if ha is not None:
            h = (ha - (_outer(ta, tb) + _outer(tb, ta)) / o.v - q * hb + 2.0 * q * _outer(tb, tb) / o.v) / o.v
        return Dual(q, (ta - q * tb) / o.v, h)

    def __rtruediv__(self, o):
        return self._lift(o) / self

    def __pow__(self, p):
        if isinstance(p, Dual):
            return exp(p * log(self))
        h = None
        if self.h is not None:
            h = p * self.v ** (p - 1) * self.h + p * (p - 1) * self.v ** (p - 2) * _outer(self.t, self.t)
        return Dual(self.v ** p, p * self.v ** (p - 1) * self.t, h)

    def __getitem__(self, idx):
        idx = idx if isinstance(idx, tuple) else (idx,)
        return Dual(self.v[idx], self.t[(slice(None),) + idx], None if self.h is None else self.h[(slice(None), slice(None)) + idx])


def _chain(x, f, f1, f2):
    """f(x) with f' = f1 and f'' = f2 evaluated at x.v"""
    h = None if x.h is None else f1 * x.h + f2 * _outer(x.t, x.t)
    return Dual(f, f1 * x.t, h)


def _masked(x, on, off_value):
    return Dual(np.where(on, x.v, off_value), np.where(on, x.t, 0.0), None if x.h is None else np.where(on, x.h, 0.0))


def val(x):
    return x.v if isinstance(x, Dual) else np.asarray(x)


def tan(x, P):
    if isinstance(x, Dual):
        return x.t
    x = np.asarray(x)
    return np.zeros((P,) + x.shape)


def exp(x):
    if isinstance(x, Dual):
        e = np.exp(x.v)
        return _chain(x, e, e, e)
    return np.exp(x)


def log(x):
    if isinstance(x, Dual):
        if x.h is None:
            return Dual(np.log(x.v), x.t / x.v)
        return _chain(x, np.log(x.v), 1.0 / x.v, -1.0 / (x.v * x.v))
    return np.log(x)


def sqrt(x):
    if isinstance(x, Dual):
        s = np.sqrt(x.v)
        with np.errstate(divide="ignore", invalid="ignore"):
            g = np.where(s > 0, 0.5 / np.where(s > 0, s, 1.0), 0.0)
            g2 = np.where(s > 0, -0.5 * g / np.where(s > 0, x.v, 1.0), 0.0)
        return _chain(x, s, g, g2)
    return np.sqrt(x)


def relu(x):
    if isinstance(x, Dual):
        return _masked(x, x.v > 0, 0.0)
    return np.maximum(x, 0.0)


def clamp_min(x, c):
    """torch.clamp(x, min=c): gradient passes where x >= c."""
    if isinstance(x, Dual):
        return _masked(x, x.v >= c, c)
    return np.maximum(x, c)


def clamp(x, lo, hi):
    if isinstance(x, Dual):
        on = (x.v >= lo) & (x.v <= hi)
        return Dual(np.clip(x.v, lo, hi), np.where(on, x.t, 0.0), None if x.h is None else np.where(on, x.h, 0.0))
    return np.clip(x, lo, hi)


def where(cond, a, b):
    if isinstance(a, Dual) or isinstance(b, Dual):
        ref = a if isinstance(a, Dual) else b
        a = ref._lift(a)
        b, ta, tb = a._pair(b)
        ha, hb = a._pair_h(b)
        return Dual(np.where(cond, a.v, b.v), np.where(cond, ta, tb), None if ha is None else np.where(cond, ha, hb))
    return np.where(cond, a, b)


def fuzzy(x, is_fuzzy, eps):
    """compute_degree_of_truth (src/maths/maths.py:3-9)."""
    if not is_fuzzy:
        return (val(x) > 0).astype(np.float64)
    return clamp((x + eps) / (2 * eps), 0.0, 1.0)


def mean(x):
    if isinstance(x, Dual):
        h = None if x.h is None else x.h.reshape(x.P, x.P, -1).mean(axis=2)
        return Dual(x.v.mean(), x.t.reshape(x.P, -1).mean(axis=1), h)
    return np.mean(x)


def const_like(c, ref):
    """scalar constant broadcast like `ref` (Dual-aware)."""
    if isinstance(ref, Dual):
        return Dual(np.full(ref.v.shape, float(c)), np.zeros_like(ref.t), None if ref.h is None else np.zeros_like(ref.h))
    return np.full(np.shape(ref), float(c))


def params(values, differentiate, second_order=False):
    """List of scalar parameters; with differentiate=True each is a 0-d Dual seeded 1 (second_order: with a zero
    Hessian, so every result carries its second derivatives too)."""
    if not differentiate:
        return [np.float64(v) for v in values]
    P = len(values)
    out = []
    for i, v in enumerate(values):
        t = np.zeros(P)
        t[i] = 1.0
        out.append(Dual(np.float64(v), t, np.zeros((P, P)) if second_order else None))
    return out
