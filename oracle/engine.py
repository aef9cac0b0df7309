"""Oracle restatement of MonteCarloEngine (src/engine/engine.py:8-123).  TEST ONLY."""
import numpy as np

from oracle import ad
from oracle import models as M
from oracle import philox


class InjectedDraws:
    """The reference's own stream: z[n_sub, N, d] (and u[n_sub, N] for QE)."""

    def __init__(self, z, u=None):
        self.z, self.u = np.asarray(z), (None if u is None else np.asarray(u))

    def normals(self, s):
        return [self.z[s, :, j] for j in range(self.z.shape[2])]

    def uniform(self, s):
        return self.u[s]


class PhiloxDraws:
    """Same counter-based stream as the CUDA kernels (csrc/philox.cuh)."""

    def __init__(self, seed, n_paths, n_sub, dim, stream=0, path_begin=0, with_uniforms=False, n_uniform=1):
        paths = np.arange(path_begin, path_begin + n_paths, dtype=np.uint64)
        self.z = philox.normals(paths, n_sub, dim, seed, stream)
        self.u = None
        if with_uniforms:
            # uniform #(a * n_sub + s) of a path belongs to asset a, sub-step s
            u = philox.uniforms(paths, n_sub * n_uniform, seed, stream)
            self.u = u if n_uniform == 1 else u.reshape(n_uniform, n_sub, n_paths).transpose(1, 2, 0).copy()

    def normals(self, s):
        return [self.z[s, :, j] for j in range(self.z.shape[2])]

    def uniform(self, s):
        return self.u[s]


def torch_reference_draws(seed, n_paths, n_sub, dim, qe=False):
    """Regenerate the reference's draws: torch.manual_seed(seed); per sub-step one
    torch.randn(N, d, float64) and, for QE, one torch.rand(N, 1) after it
    (engine.py:25, model.py:47, heston.py:191-192)."""
    import torch
    torch.manual_seed(seed)
    z = np.empty((n_sub, n_paths, dim))
    u = np.empty((n_sub, n_paths)) if qe else None
    for s in range(n_sub):
        z[s] = torch.randn(n_paths, dim, dtype=torch.float64).numpy()
        if qe:
            u[s] = torch.rand(n_paths, 1, dtype=torch.float64).numpy()[:, 0]
    return InjectedDraws(z, u)


def substeps(t0, timeline, num_steps):
    """(t1, t2, nominal dt, date index reached or -1) per sub-step + leading un-stepped dates
    (engine.py:48-60: t_prev accumulates, dt <= 0 intervals are skipped)."""
    out, t_prev = [], float(t0)
    for di, t_now in enumerate(timeline):
        dt = (float(t_now) - t_prev) / num_steps
        if dt > 0:
            for k in range(num_steps):
                out.append((t_prev, t_prev + dt, dt, di if k == num_steps - 1 else -1))
                t_prev = t_prev + dt
    return out


def count_substeps(t0, timeline, num_steps):
    return len(substeps(t0, timeline, num_steps))


def generate_paths(model, p, timeline, n_paths, num_steps, scheme, draws, smoothing=False):
    """-> list over simulation dates of the state (list of per-path columns)."""
    scheme = scheme if isinstance(scheme, str) else scheme.name
    t0 = M.t0_of(model)
    state = M.initial_state(model, p, n_paths)
    chol_cache = {}
    paths, s, t_prev = [], 0, t0
    for t_now in timeline:
        dt = (float(t_now) - t_prev) / num_steps
        if dt > 0:
            for _ in range(num_steps):
                if scheme == "ANALYTICAL":
                    if dt not in chol_cache:
                        chol_cache[dt] = M.cholesky(M.covariance(model, p, dt))
                    L = chol_cache[dt]
                else:
                    if None not in chol_cache:
                        chol_cache[None] = M.cholesky(M.correlation(model, p, scheme))
                    L = chol_cache[None]
                w = M.correlate(draws.normals(s), L)
                u = draws.uniform(s) if scheme == "QE" else None
                state = M.step(model, p, scheme, t_prev, t_prev + dt, state, w, u, smoothing)
                t_prev = t_prev + dt
                s += 1
        paths.append(list(state))
    return paths
