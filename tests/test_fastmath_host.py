"""Accuracy of the elementary functions of the fused kernels (csrc/fastmath.cuh), checked on the CPU: the same header
is compiled for the host with shims for the CUDA intrinsics (tests/host/fastmath_host.cpp; the MUFU seeds are
modelled pessimistically).  tests/test_fastmath_gpu.py repeats the measurement on the device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "fastmath_host.cpp")
LIB = os.path.join(HERE, "host", "_build", "libfm_host.so")


@pytest.fixture(scope="module")
def fm():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    hdr = os.path.join(HERE, "..", "montecarlo-risk-engine_b200", "csrc", "fastmath.cuh")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-ffp-contract=off", "-o", LIB, SRC])
    lib = C.CDLL(LIB)
    lib.fm_host_init()

    def ev(fn, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        lib.fm_host_eval(C.c_int(fn), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_longlong(x.size))
        return y
    return ev


def _ulp(got, want):
    return np.abs(got - want) / np.spacing(np.abs(want))


def _ld(x):
    return np.asarray(x, dtype=np.longdouble)


def test_exp_table(fm):
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-700, 700, 300000), rng.uniform(-2, 2, 300000), rng.uniform(-60, 5, 200000),
                        [0.0, -0.0, 1e-300, -1e-17, 700.0, -700.0]])
    assert _ulp(fm(10, x), np.exp(_ld(x)).astype(np.float64)).max() <= 2.0
    s = rng.uniform(-2.0 ** -6, 2.0 ** -6, 200000)
    assert _ulp(fm(16, s), np.exp(_ld(s)).astype(np.float64)).max() <= 1.5


def test_log_table(fm):
    rng = np.random.default_rng(4)
    u = np.concatenate([rng.uniform(0, 1, 600000), 2.0 ** -rng.uniform(0, 54, 200000), 1.0 - 2.0 ** -rng.uniform(1, 53, 200000),
                        rng.uniform(1, 2, 100000), 1.0 + 2.0 ** -rng.uniform(8, 50, 100000),
                        [1.0, 0.5, 2.0 ** -54, 1 - 2.0 ** -53, 1 - 2.0 ** -8, np.nextafter(1 - 2.0 ** -8, 0), 1.4140625 / 2,
                         np.nextafter(1.4140625 / 2, 0), 1.4140625, np.nextafter(1.4140625, 0), np.nextafter(2.0, 0)]])
    u = u[(u > 0) & (u < 2.0)]
    want = np.log(_ld(u))
    lg = fm(11, u)
    assert np.all(np.abs(lg - want.astype(np.float64)) <= 3 * np.spacing(np.abs(want.astype(np.float64))) + 1e-300)
    # the Box-Muller radius: -2 log u for u in (0, 1)
    uu = u[u < 1.0]
    w3 = (-2 * np.log(_ld(uu))).astype(np.float64)
    assert np.all(np.abs(fm(15, uu) - w3) <= 2 * np.spacing(w3))


def test_sqrt(fm):
    rng = np.random.default_rng(5)
    v = np.concatenate([rng.uniform(0, 100, 300000), 10.0 ** rng.uniform(-12, 6, 300000), [1.0, 4.0, 1e-12, 2.0]])
    assert _ulp(fm(12, v), np.sqrt(v)).max() <= 1.0
    assert _ulp(fm(2, v), np.sqrt(v)).max() <= 1.0
    assert fm(2, np.array([0.0]))[0] == 0.0


def test_sincos_and_box_muller_rotation(fm):
    rng = np.random.default_rng(6)
    v = np.concatenate([rng.uniform(0, 1, 600000), [0.0, 0.25, 0.5, 0.75, 0.125, 1 - 2.0 ** -52, 2.0 ** -52, 1 / 1024, 3 / 1024,
                                                    1 - 1 / 1024, np.nextafter(1 - 1 / 1024, 1)]])
    two_pi = 2 * np.longdouble("3.14159265358979323846264338327950288")
    s = np.sin(two_pi * _ld(v)).astype(np.float64)
    c = np.cos(two_pi * _ld(v)).astype(np.float64)
    assert np.max(np.abs(fm(13, v) - s)) < 4e-16
    assert np.max(np.abs(fm(14, v) - c)) < 4e-16
    # fm_polar_tv works on d = 1 + v, i.e. v on the 2^-52 grid
    vg = np.floor(v * 2.0 ** 52) * 2.0 ** -52
    sg = np.sin(two_pi * _ld(vg)).astype(np.float64)
    cg = np.cos(two_pi * _ld(vg)).astype(np.float64)
    assert np.max(np.abs(fm(17, vg) - cg)) < 4e-16
    assert np.max(np.abs(fm(18, vg) - sg)) < 4e-16


def test_reference_series_functions(fm):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-700, 700, 100000), rng.uniform(-2, 2, 100000)])
    assert _ulp(fm(0, x), np.exp(x)).max() <= 4
    u = np.concatenate([rng.uniform(0, 1, 200000), 2.0 ** -rng.uniform(0, 54, 50000)])
    want = np.log(u)
    assert np.all(np.abs(fm(1, u) - want) <= 4 * np.spacing(np.abs(want)) + 1e-300)
    d = np.concatenate([rng.uniform(0.1, 10, 100000), -rng.uniform(0.1, 10, 1000)])
    assert _ulp(fm(5, d), 1.0 / d).max() <= 2
