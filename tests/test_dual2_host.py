"""The kernels' forward-mode numbers (csrc/dual.cuh: value + tangents; csrc/dual2.cuh: + second derivatives), compiled
for the host (tests/host/dual2_host.cpp) and checked against torch autograd on a composite of every operation the
kernels use: the gradient against torch.autograd.grad, the Hessian against torch's double backward - the reference's
compute_higher_derivatives (controller.py:631-648).  Piecewise-linear functions must carry zero curvature like torch's."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "dual2_host.cpp")
LIB = os.path.join(HERE, "host", "_build", "libdual2_host.so")
CSRC = os.path.join(HERE, "..", "montecarlo-risk-engine_b200", "csrc")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    newest = max(os.path.getmtime(p) for p in (SRC, os.path.join(CSRC, "dual.cuh"), os.path.join(CSRC, "dual2.cuh")))
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas",
                               "-o", LIB, SRC])
    return C.CDLL(LIB)


def _torch_composite(p, z):
    s, v, r = p
    S = s * torch.exp((r - 0.5 * v * v) * 2.0 + v * (1.4142135623730951 * z))
    pay = torch.relu(S - 1.2) * torch.clamp((S - 1.0 + 0.5) / (2 * 0.5), min=0.0, max=1.0) / torch.sqrt(S)
    pay = pay + torch.log(S) * torch.log(S) - 1.0 / S
    pay = pay + torch.clamp(S, min=1.5) * (S / v) + torch.where(z > 0, S * S, torch.zeros_like(S)) - 2.0 / (S + 3.0) + (-S) * 0.25
    pay = pay + torch.sqrt(S + 1.0) * torch.exp(r * 0.01)
    return (pay * torch.exp(-(r * 2.0))).mean()


@pytest.mark.parametrize("x0", [(1.3, 0.4, 0.07), (0.9, 0.25, 0.02), (1.6, 0.6, -0.01)])
def test_second_order_numbers_match_torch_double_backward(lib, x0):
    z = np.array([0.3, -1.2, 0.8, 2.0, -0.1, 0.55, -2.3, 1.1])
    x = np.array(x0)
    out2, out1 = np.zeros(10), np.zeros(4)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    lib.dual2_host_second(ptr(x), ptr(z), C.c_int(z.size), ptr(out2))
    lib.dual2_host_first(ptr(x), ptr(z), C.c_int(z.size), ptr(out1))
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    zt = torch.tensor(z, dtype=torch.float64)
    val = _torch_composite(xt, zt)
    v0 = float(val.detach())
    grad = torch.autograd.grad(val, xt, create_graph=True)[0]
    hess = torch.stack([torch.autograd.grad(grad[i], xt, retain_graph=True)[0] for i in range(3)]).numpy()
    assert abs(out2[0] - v0) <= 1e-14 * abs(v0)
    assert abs(out1[0] - v0) <= 1e-14 * abs(v0)
    g = grad.detach().numpy()
    assert np.allclose(out1[1:4], g, rtol=1e-13, atol=1e-13)
    assert np.allclose(out2[1:4], g, rtol=1e-13, atol=1e-13)
    tri = [hess[i, j] for i in range(3) for j in range(i, 3)]
    assert np.allclose(out2[4:10], tri, rtol=1e-12, atol=1e-12), (out2[4:10], tri)
    assert np.allclose(hess, hess.T, rtol=1e-12, atol=1e-12)
