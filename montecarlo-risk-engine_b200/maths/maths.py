"""Fuzzy-logic indicator used by barrier/binary payoffs and Heston-QE
(reference: src/maths/maths.py:3-9).  Host-side torch version kept for API
compatibility; the kernels use the device restatement in csrc/dual.cuh."""
from common.packages import *


def symmetric_linear_smoothing(x, is_fuzzy, eps):
    if not is_fuzzy:
        return (x > 0).float()
    return torch.clamp((x + eps) / (2 * eps), min=0.0, max=1.0)


def compute_degree_of_truth(x, is_fuzzy, eps=0.05):
    return symmetric_linear_smoothing(x, is_fuzzy, eps)


def bisection_search(func, low: float = 1e-10, high: float = 5.0, tolerance: float = 1e-12, iters: int = 100):
    """Root of `func` on [low, high] (reference: src/maths/maths.py:14-33; the CDS bootstrap's solver).  The upper end is
    doubled up to 20 times until the end values differ in sign; None if they never do."""
    f_low, f_high = func(low), func(high)
    for _ in range(20):
        if f_low * f_high <= 0.0:
            break
        high *= 2.0
        f_high = func(high)
    else:
        if f_low * f_high > 0.0:
            return None
    mid = 0.5 * (low + high)
    for _ in range(iters):
        mid = 0.5 * (low + high)
        f_mid = func(mid)
        if abs(f_mid) < tolerance or high - low < 1e-12:
            return mid
        if f_low * f_mid <= 0.0:
            high = mid
        else:
            low, f_low = mid, f_mid
    return 0.5 * (low + high)
