"""Host logic of the regressions (mcre/lsm.py), CPU only: the normal-equation solver against numpy's tall
least squares (minimum norm on rank-deficient dates, like LAPACK gelsy in controller.py:368-374) and the
differentiated normal equations (regression_tangents: what replaces torch.autograd through
torch.linalg.lstsq for sensitivities of exposure metrics) against central finite differences of the tall
least squares."""
import numpy as np
import pytest

import cases  # noqa: F401  (puts the package on sys.path through conftest)
from mcre.lsm import gram_pinv, regression_tangents, solve_normal_equations, solve_normal_equations_batch, to_raw_basis


def _design(u):
    return np.stack([np.ones_like(u), u, u * u], axis=1)


def _moments(u, y):
    A = _design(u)
    return A.T @ A, A.T @ y


def test_normal_equations_match_tall_least_squares():
    rng = np.random.default_rng(3)
    u = rng.standard_normal(5000)
    y = 0.3 - 1.2 * u + 0.7 * u * u + 0.05 * rng.standard_normal(5000)
    G, b = _moments(u, y)
    want = np.linalg.lstsq(_design(u), y, rcond=None)[0]
    assert np.allclose(solve_normal_equations(G, b), want, rtol=1e-10, atol=1e-12)
    assert np.allclose(solve_normal_equations_batch(G[None], b[None])[0], want, rtol=1e-10, atol=1e-12)
    # rank deficient: every path has the same explanatory value -> minimum-norm solution
    u0 = np.full(100, 0.0)
    y0 = rng.standard_normal(100)
    G0, b0 = _moments(u0, y0)
    c0 = solve_normal_equations(G0, b0)
    assert np.allclose(c0, [y0.mean(), 0.0, 0.0], atol=1e-12)
    assert np.allclose(gram_pinv(G0) @ G0 @ gram_pinv(G0), gram_pinv(G0), atol=1e-14)


def test_raw_basis_conversion_preserves_the_fitted_function():
    rng = np.random.default_rng(4)
    coef = rng.standard_normal((3, 3))
    basis = np.array([[0.03, 25.0], [0.05, 11.0], [-0.01, 3.0]])
    raw = to_raw_basis(coef, basis)
    x = rng.standard_normal(7) * 0.1
    for k in range(3):
        u = (x - basis[k, 0]) * basis[k, 1]
        assert np.allclose(coef[k, 0] + coef[k, 1] * u + coef[k, 2] * u * u,
                           raw[k, 0] + raw[k, 1] * x + raw[k, 2] * x * x, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("degenerate", [False, True])
def test_regression_tangents_match_finite_differences(degenerate):
    """u(theta), Y(theta) per path with known tangents: dc/dtheta from the nine moment sums per parameter
    (what mcre_irc_presim accumulates) vs central differences of the tall least squares.  The degenerate
    case is the t = 0 date (all paths share x): the fit there is the mean of Y and its derivative the mean of dY,
    whatever the rank-deficient coefficients do."""
    rng = np.random.default_rng(7)
    n, nt = 4000, 3
    z = rng.standard_normal(n)
    w = rng.standard_normal(n)

    def u_of(th):
        return (0.0 * z + th[0]) if degenerate else (th[0] + (1.0 + th[1]) * z + 0.1 * th[2] * z * z)

    def y_of(th):
        return np.exp(0.2 * th[1] * z) * (1.0 + th[0]) + th[2] * w + 0.5 * z * z

    th0 = np.array([0.1, 0.3, -0.2])
    eps = 1e-6
    u, y = u_of(th0), y_of(th0)
    du = np.stack([(u_of(th0 + eps * e) - u_of(th0 - eps * e)) / (2 * eps) for e in np.eye(nt)])
    dy = np.stack([(y_of(th0 + eps * e) - y_of(th0 - eps * e)) / (2 * eps) for e in np.eye(nt)])
    G, b = _moments(u, y)
    coef = solve_normal_equations(G, b)
    tm = np.zeros((nt, 9))
    for p in range(nt):
        tm[p, 0:4] = [np.sum(u ** m * du[p]) for m in range(4)]
        tm[p, 4:7] = [np.sum(u ** i * dy[p]) for i in range(3)]
        tm[p, 7], tm[p, 8] = np.sum(du[p] * y), np.sum(2.0 * u * du[p] * y)
    dc = regression_tangents(G, b, coef, tm)

    def fit_at(th, x):
        c = np.linalg.lstsq(_design(u_of(th)), y_of(th), rcond=None)[0]
        return c[0] + c[1] * x + c[2] * x * x

    if not degenerate:
        for p, e in enumerate(np.eye(nt)):
            fd = (np.linalg.lstsq(_design(u_of(th0 + eps * e)), y_of(th0 + eps * e), rcond=None)[0]
                  - np.linalg.lstsq(_design(u_of(th0 - eps * e)), y_of(th0 - eps * e), rcond=None)[0]) / (2 * eps)
            assert np.allclose(dc[p], fd, rtol=2e-6, atol=2e-7), (p, dc[p], fd)
    else:
        # derivative of the fitted value at the (moving) common point x = u(theta): d/dtheta [phi(x) . c]
        x0 = th0[0]
        phi, dphi = np.array([1.0, x0, x0 * x0]), np.array([0.0, 1.0, 2.0 * x0])
        for p, e in enumerate(np.eye(nt)):
            fd = (fit_at(th0 + eps * e, (th0 + eps * e)[0]) - fit_at(th0 - eps * e, (th0 - eps * e)[0])) / (2 * eps)
            got = phi @ dc[p] + (dphi @ coef) * e[0]
            assert got == pytest.approx(np.mean(dy[p]), rel=1e-9, abs=1e-9)
            assert got == pytest.approx(fd, rel=1e-5, abs=1e-6)
