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
