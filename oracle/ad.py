"""Forward-mode AD on numpy arrays for the oracle's Greeks.

A ``Dual`` carries a value array and one tangent array per model parameter
(shape (P,) + value.shape).  The oracle's formulas are written against the functions
below, so the same code yields values (plain ndarrays) or values + pathwise
sensitivities - the latter must agree with the reference's torch.autograd output
(src/controller/controller.py:609-627)."""
import numpy as np


class Dual:
    __array_priority__ = 1000

    def __init__(self, v, t):
        self.v = np.asarray(v, dtype=np.float64)
        self.t = np.asarray(t, dtype=np.float64)

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

    def __add__(self, o):
        o, ta, tb = self._pair(o)
        return Dual(self.v + o.v, ta + tb)

    __radd__ = __add__

    def __neg__(self):
        return Dual(-self.v, -self.t)

    def __sub__(self, o):
        o, ta, tb = self._pair(o)
        return Dual(self.v - o.v, ta - tb)

    def __rsub__(self, o):
        o, ta, tb = self._pair(o)
        return Dual(o.v - self.v, tb - ta)

    def __mul__(self, o):
        o, ta, tb = self._pair(o)
        return Dual(self.v * o.v, ta * o.v + self.v * tb)

    __rmul__ = __mul__

    def __truediv__(self, o):
        o, ta, tb = self._pair(o)
        q = self.v / o.v
        return Dual(q, (ta - q * tb) / o.v)

    def __rtruediv__(self, o):
        return self._lift(o) / self

    def __pow__(self, p):
        if isinstance(p, Dual):
            return exp(p * log(self))
        return Dual(self.v ** p, p * self.v ** (p - 1) * self.t)

    def __getitem__(self, idx):
        return Dual(self.v[idx], self.t[(slice(None),) + (idx if isinstance(idx, tuple) else (idx,))])


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
        return Dual(e, e * x.t)
    return np.exp(x)


def log(x):
    if isinstance(x, Dual):
        return Dual(np.log(x.v), x.t / x.v)
    return np.log(x)


def sqrt(x):
    if isinstance(x, Dual):
        s = np.sqrt(x.v)
        with np.errstate(divide="ignore", invalid="ignore"):
            g = np.where(s > 0, 0.5 / np.where(s > 0, s, 1.0), 0.0)
        return Dual(s, g * x.t)
    return np.sqrt(x)


def relu(x):
    if isinstance(x, Dual):
        on = x.v > 0
        return Dual(np.where(on, x.v, 0.0), np.where(on, x.t, 0.0))
    return np.maximum(x, 0.0)


def clamp_min(x, c):
    """torch.clamp(x, min=c): gradient passes where x >= c."""
    if isinstance(x, Dual):
        on = x.v >= c
        return Dual(np.where(on, x.v, c), np.where(on, x.t, 0.0))
    return np.maximum(x, c)


def clamp(x, lo, hi):
    if isinstance(x, Dual):
        on = (x.v >= lo) & (x.v <= hi)
        return Dual(np.clip(x.v, lo, hi), np.where(on, x.t, 0.0))
    return np.clip(x, lo, hi)


def where(cond, a, b):
    if isinstance(a, Dual) or isinstance(b, Dual):
        ref = a if isinstance(a, Dual) else b
        a = ref._lift(a)
        b, ta, tb = a._pair(b)
        return Dual(np.where(cond, a.v, b.v), np.where(cond, ta, tb))
    return np.where(cond, a, b)


def fuzzy(x, is_fuzzy, eps):
    """compute_degree_of_truth (src/maths/maths.py:3-9)."""
    if not is_fuzzy:
        return (val(x) > 0).astype(np.float64)
    return clamp((x + eps) / (2 * eps), 0.0, 1.0)


def mean(x):
    if isinstance(x, Dual):
        return Dual(x.v.mean(), x.t.reshape(x.P, -1).mean(axis=1))
    return np.mean(x)


def const_like(c, ref):
    """scalar constant broadcast like `ref` (Dual-aware)."""
    if isinstance(ref, Dual):
        return Dual(np.full(ref.v.shape, float(c)), np.zeros_like(ref.t))
    return np.full(np.shape(ref), float(c))


def params(values, differentiate):
    """List of scalar parameters; with differentiate=True each is a 0-d Dual seeded 1."""
    if not differentiate:
        return [np.float64(v) for v in values]
    P = len(values)
    out = []
    for i, v in enumerate(values):
        t = np.zeros(P)
        t[i] = 1.0
        out.append(Dual(np.float64(v), t))
    return out
