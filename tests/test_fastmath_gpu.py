"""Accuracy of the branch-free FP64 functions (csrc/fastmath.cuh) that replace libdevice
inside the fused kernels: <= 4 ulp against numpy on the argument ranges the Monte Carlo uses."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _probe(fn, x):
    from mcre import binding as B
    from mcre import runtime as RT
    dev = RT.compute_device()
    xd = torch.as_tensor(x, dtype=torch.float64, device=dev)
    yd = torch.empty_like(xd)
    B.check(B.lib().mcre_fastmath_probe(fn, xd.data_ptr(), yd.data_ptr(), xd.numel(), RT.stream_ptr()))
    return yd.cpu().numpy()


def _ulp_err(got, want):
    return np.abs(got - want) / np.spacing(np.abs(want))


def test_exp_log_sqrt_div():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-700, 700, 200000), rng.uniform(-2, 2, 200000), [0.0, -0.0, 1e-300, -1e-17]])
    assert _ulp_err(_probe(0, x), np.exp(x)).max() <= 4
    u = np.concatenate([rng.uniform(0, 1, 400000), 2.0 ** -rng.uniform(0, 54, 100000), [1.0, 0.5, 2.0 ** -54, 1 - 2.0 ** -53]])
    lg = _probe(1, u)
    want = np.log(u)
    assert np.all(np.abs(lg - want) <= 4 * np.spacing(np.abs(want)) + 1e-300)
    v = np.concatenate([rng.uniform(0, 100, 200000), 10.0 ** rng.uniform(-12, 6, 200000), [0.0, 1.0, 4.0]])
    assert _ulp_err(_probe(2, v)[v > 0], np.sqrt(v[v > 0])).max() <= 2
    assert _probe(2, np.array([0.0]))[0] == 0.0
    d = np.concatenate([rng.uniform(0.1, 10, 200000), -rng.uniform(0.1, 10, 1000)])
    assert _ulp_err(_probe(5, d), 1.0 / d).max() <= 2


def test_sincos_2pi():
    rng = np.random.default_rng(2)
    u = np.concatenate([rng.uniform(0, 1, 500000), [0.0, 0.25, 0.5, 0.75, 0.125, 1 - 2.0 ** -53, 2.0 ** -54]])
    ld = u.astype(np.longdouble)
    s = np.sin(2 * np.pi * ld).astype(np.float64) if np.finfo(np.longdouble).eps < 1e-18 else np.sin(2 * np.pi * u)
    c = np.cos(2 * np.pi * ld).astype(np.float64) if np.finfo(np.longdouble).eps < 1e-18 else np.cos(2 * np.pi * u)
    # absolute accuracy (the products rad * cos / rad * sin inherit it): a few 1e-16
    assert np.max(np.abs(_probe(3, u) - s)) < 1e-15
    assert np.max(np.abs(_probe(4, u) - c)) < 1e-15


def test_table_driven_variants():
    """Shared-memory-table versions used by the fused kernels (fastmath.cuh, MCRE_FAST_MATH 2)."""
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-700, 700, 300000), rng.uniform(-2, 2, 300000), rng.uniform(-60, 5, 200000),
                        [0.0, -0.0, 1e-300, -1e-17, 700.0, -700.0]])
    assert _ulp_err(_probe(10, x), np.exp(x)).max() <= 4
    u = np.concatenate([rng.uniform(0, 1, 600000), 2.0 ** -rng.uniform(0, 54, 200000), 1.0 - 2.0 ** -rng.uniform(1, 53, 200000),
                        [1.0, 0.5, 2.0 ** -54, 1 - 2.0 ** -53, 1 - 2.0 ** -8, np.nextafter(1 - 2.0 ** -8, 0), 1.4140625 / 2,
                         np.nextafter(1.4140625 / 2, 0)]])
    u = u[(u > 0) & (u <= 1.0)]
    lg, want = _probe(11, u), np.log(u)
    assert np.all(np.abs(lg - want) <= 4 * np.spacing(np.abs(want)) + 1e-300)
    v = np.concatenate([rng.uniform(0, 1, 600000), [0.0, 0.25, 0.5, 0.75, 0.125, 1 - 2.0 ** -53, 2.0 ** -54, 1 / 256, 3 / 256]])
    ld = v.astype(np.longdouble)
    two_pi = 2 * np.longdouble("3.14159265358979323846264338327950288")
    s = np.sin(two_pi * ld).astype(np.float64)
    c = np.cos(two_pi * ld).astype(np.float64)
    assert np.max(np.abs(_probe(13, v) - s)) < 5e-16
    assert np.max(np.abs(_probe(14, v) - c)) < 5e-16
