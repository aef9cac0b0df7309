"""Lowering + driver of the generic path generator (csrc/paths.cu): the compatibility
seam for MonteCarloEngine.generate_paths() and Model.simulate_time_step_*."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from common.enums import SimulationScheme
from mcre import binding as B
from mcre import runtime as RT
from mcre.dual import D, cholesky_dual
from mcre.timegrid import build_time_grid


class PathsDesc(C.Structure):
    _fields_ = [
        ("n_models", C.c_int32), ("model_kind", B.c_ip), ("model_nassets", B.c_ip), ("model_params", B.c_dp),
        ("model_flags", B.c_ip), ("scheme", C.c_int32), ("noise_dim", C.c_int32), ("state_dim", C.c_int32),
        ("n_sub", C.c_int32), ("n_dates", C.c_int32), ("n_pre_dates", C.c_int32),
        ("step_dt", B.c_dp), ("step_t1", B.c_dp), ("step_date", B.c_ip), ("chol", B.c_dp), ("step_chol", B.c_ip),
        ("step_aux", B.c_dp), ("init_state", B.c_dp),
    ]


_SCHEME = {SimulationScheme.EULER: B.SCHEME_EULER, SimulationScheme.ANALYTICAL: B.SCHEME_ANALYTICAL,
           SimulationScheme.QE: B.SCHEME_QE}


def flat_models(model):
    from models.model_config import ModelConfig
    return list(model.models) if isinstance(model, ModelConfig) else [model]


def joint_matrix(model, scheme, dt):
    """Correlation (EULER/QE) or step covariance (ANALYTICAL) of the joint noise as floats
    (reference: model.py:50-81, model_config.py:101-221)."""
    from models.black_scholes import BlackScholesModel
    from models.black_scholes_multi import BlackScholesMulti
    from models.model_config import ModelConfig
    from models.schwartz_two_factor import SchwartzTwoFactorModel
    from models.vasicek import VasicekModel
    subs = flat_models(model)
    ps = [m.dual_params() for m in subs]
    if scheme != SimulationScheme.ANALYTICAL:
        if isinstance(model, ModelConfig):
            return [[x.v for x in row] for row in model.joint_correlation(scheme, ps)]
        return [[x.v for x in row] for row in model.intra_correlation(scheme, ps[0])]

    def cov_of(m, p):
        if isinstance(m, BlackScholesModel):
            return [[p[1].v ** 2 * dt]]
        if isinstance(m, BlackScholesMulti):
            c = m.correlation_matrix.numpy()
            n = m.num_assets
            return [[p[n + i].v * c[i, j] * p[n + j].v * dt for j in range(n)] for i in range(n)]
        if isinstance(m, VasicekModel):
            _, nstd = m.exact_step_constants(p, dt)
            return [[nstd.v ** 2]]
        if isinstance(m, SchwartzTwoFactorModel):
            return [[x.v for x in row] for row in m.exact_covariance(p, dt)]
        n = m.simulation_dim
        return [[dt if i == j else 0.0 for j in range(n)] for i in range(n)]

    if not isinstance(model, ModelConfig):
        return cov_of(model, ps[0])
    n = model.num_assets
    cov = np.zeros((n, n))
    row, idx = 0, 0
    for i, m1 in enumerate(subs):
        n1 = m1.num_assets
        cov[row:row + n1, row:row + n1] = np.array(cov_of(m1, ps[i]))
        col = row + n1
        for j, m2 in enumerate(subs[i + 1:], start=i + 1):
            n2 = m2.num_assets
            if not (isinstance(m1, BlackScholesModel) and isinstance(m2, BlackScholesModel)):
                raise NotImplementedError("Inter covariance not implemented for the requested pair of models.")
            ic = model.inter_asset_correlation_matrix[idx].numpy()
            blk = ps[i][1].v * ps[j][1].v * np.broadcast_to(ic, (n1, n2)) * dt
            cov[row:row + n1, col:col + n2] = blk
            cov[col:col + n2, row:row + n1] = blk.T
            col += n2
            idx += 1
        row += n1
    return (0.5 * (cov + cov.T)).tolist()


def initial_state_values(model):
    from models.black_scholes import BlackScholesModel
    from models.black_scholes_multi import BlackScholesMulti
    from models.cirpp import CIRPPModel
    from models.heston import HestonModel
    from models.schwartz_two_factor import SchwartzTwoFactorModel
    from models.vasicek import VasicekModel
    out = []
    for m in flat_models(model):
        p = m.param_values()
        if isinstance(m, BlackScholesModel):
            out += [p[0]]
        elif isinstance(m, BlackScholesMulti):
            out += p[:m.num_assets]
        elif isinstance(m, HestonModel):
            out += [math.log(p[0]), p[6]]
        elif isinstance(m, VasicekModel):
            out += [p[0], 0.0]
        elif isinstance(m, CIRPPModel):
            out += [m.market_hazard(m.t0()) if m.deterministic else p[3], 0.0]
        elif isinstance(m, SchwartzTwoFactorModel):
            out += [math.log(m.curve_value(m.t0())), 0.0, 0.0]
        else:
            raise NotImplementedError(type(m).__name__)
    return out


def generate(model, timeline, n_paths, num_steps, scheme, seed, inject_z=None, inject_u=None,
             grid=None, stream_id=0, init_state=None, identity_chol=False, path_begin=0, n_total=None):
    """-> torch.Tensor [n_paths, n_dates, state_dim] on the compute device.
    `path_begin` / `n_total`: this process simulates global path ids [path_begin, path_begin + n_paths)
    of `n_total` (path sharding; injected draws are indexed by global id)."""
    from models.cirpp import CIRPPModel
    from models.schwartz_two_factor import SchwartzTwoFactorModel
    from models.vasicek import VasicekModel
    dev = RT.compute_device()
    L = B.lib()
    L.mcre_generate_paths.argtypes = [C.POINTER(PathsDesc), C.POINTER(B.Rng), C.POINTER(B.Shard), C.c_void_p, C.c_void_p]
    subs = flat_models(model)
    if grid is None:
        grid = build_time_grid(model.t0(), [float(t) for t in timeline], num_steps)
    n_sub, n_dates = grid.n_sub, len(grid.dates)
    noise_dim = sum(m.simulation_dim for m in subs)
    state_dim = sum(m.state_dim for m in subs)
    # Cholesky factors: one for the correlation, or one per distinct nominal dt (ANALYTICAL)
    chols, step_chol, key_to_idx = [], [], {}
    for s in range(n_sub):
        if identity_chol:      # caller supplies already-correlated noise
            step_chol.append(0)
            continue
        key = grid.dt_nominal[s] if scheme == SimulationScheme.ANALYTICAL else None
        if key not in key_to_idx:
            mat = joint_matrix(model, scheme, key)
            nt0 = [[D(x, None, 0) for x in row] for row in mat]
            Lm = cholesky_dual(nt0)
            chols.append(np.array([[x.v for x in row] for row in Lm]))
            key_to_idx[key] = len(chols) - 1
        step_chol.append(key_to_idx[key])
    if not chols:
        chols.append(np.eye(noise_dim))
    aux = np.zeros((max(n_sub, 1), len(subs), 4))
    for mi, m in enumerate(subs):
        p = m.dual_params()
        for s in range(n_sub):
            if isinstance(m, CIRPPModel):
                if m.deterministic:
                    aux[s, mi, 1], aux[s, mi, 2] = m.market_hazard(grid.t1[s]), m.market_hazard(grid.t2[s])
                else:
                    aux[s, mi, 0] = m.psi(p, grid.t1[s]).v
            elif isinstance(m, SchwartzTwoFactorModel):
                aux[s, mi, 0] = math.log(m.curve_value(grid.t2[s]))
            elif isinstance(m, VasicekModel):
                aux[s, mi, 3] = m.mean_level(p, grid.t1[s]).v
    keep = []

    def fp(a):
        arr, ptr = B.as_dp(a)
        keep.append(arr)
        return ptr

    def ip(a):
        arr, ptr = B.as_ip(a)
        keep.append(arr)
        return ptr

    d = PathsDesc()
    d.n_models = len(subs)
    d.model_kind = ip([m.KIND for m in subs])
    d.model_nassets = ip([m.num_assets for m in subs])
    d.model_params = fp([v for m in subs for v in m.param_values()])
    d.model_flags = ip([(1 if getattr(m, "deterministic", False) else 0) | (2 if m.perform_smoothing else 0) for m in subs])
    d.scheme, d.noise_dim, d.state_dim = _SCHEME[scheme], noise_dim, state_dim
    d.n_sub, d.n_dates, d.n_pre_dates = n_sub, n_dates, grid.n_pre_dates
    d.step_dt, d.step_t1 = fp(grid.dt if n_sub else [0.0]), fp(grid.t1 if n_sub else [0.0])
    d.step_date, d.step_chol = ip(grid.date_after if n_sub else [0]), ip(step_chol if n_sub else [0])
    d.chol = fp(np.stack(chols))
    d.step_aux = fp(aux)
    d.init_state = fp(initial_state_values(model) if init_state is None else init_state)
    rng = B.Rng()
    rng.seed, rng.stream, rng.n_paths_total = seed, stream_id, (n_paths if n_total is None else n_total)
    if inject_z is not None:
        z = torch.as_tensor(inject_z, dtype=torch.float64).to(dev).contiguous()
        keep.append(z)
        rng.mode, rng.d_z = B.RNG_INJECT, z.data_ptr()
        if inject_u is not None:
            u = torch.as_tensor(inject_u, dtype=torch.float64).to(dev).contiguous()
            keep.append(u)
            rng.d_u = u.data_ptr()
    else:
        rng.mode = B.RNG_PHILOX
    out = torch.empty((n_paths, n_dates, state_dim), dtype=torch.float64, device=dev)
    sh = B.Shard(path_begin, n_paths, 256)
    B.check(L.mcre_generate_paths(C.byref(d), C.byref(rng), C.byref(sh), out.data_ptr(), RT.stream_ptr()))
    for di in grid.zero_dt_dates:        # dates the running time had already reached: state unchanged (engine.py:48-60)
        out[:, di, :] = out[:, grid.alias[di], :]
    return out
