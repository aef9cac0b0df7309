"""Oracle restatement of the reference's models (TEST INFRASTRUCTURE ONLY).

Objects are read by duck typing (class name + the reference's attribute names), so both
the reference's own model objects and this repo's host-side descriptions work.  States
are lists of per-path column arrays (numpy or oracle.ad.Dual)."""
import math
from bisect import bisect_right

import numpy as np

from oracle import ad


def _f(x):
    try:
        return float(x)
    except (TypeError, ValueError):
        return float(np.asarray(x).reshape(-1)[0])


def kind(model):
    return type(model).__name__


def t0_of(model):
    return _f(model.calibration_date)


def param_values(model):
    return [_f(p) for p in model.model_params]


def submodels(model):
    return list(model.models) if kind(model) == "ModelConfig" else [model]


def state_dim(model):
    return {"BlackScholesModel": 1, "HestonModel": 2, "VasicekModel": 2, "HullWhiteModel": 2, "CIRPPModel": 2,
            "SchwartzTwoFactorModel": 3}.get(kind(model)) or (
        len(model.asset_ids) if kind(model) == "BlackScholesMulti" else sum(state_dim(m) for m in model.models))


def noise_dim(model):
    k = kind(model)
    if k in ("HestonModel", "SchwartzTwoFactorModel"):
        return 2
    if k == "BlackScholesMulti":
        return len(model.asset_ids)
    if k == "ModelConfig":
        return sum(noise_dim(m) for m in model.models)
    return 1


# ---------------------------------------------------------------------------- CIR++ helpers
def cir_market_hazard(model, t):
    """cirpp.py:66-75: right-closed buckets, flat extrapolation."""
    ten = [_f(x) for x in model.tenors]
    haz = [_f(x) for x in model.hazard_rates]
    for tt, h in zip(ten, haz):
        if t <= tt:
            return h
    return haz[-1]


def cir_market_survival(model, t):
    """helpers/cs_helper.py:80-107."""
    ten = [_f(x) for x in model.tenors]
    haz = [_f(x) for x in model.hazard_rates]
    surv, prev, idx = 1.0, 0.0, len(ten) - 1
    for i, mat in enumerate(ten):
        if mat <= t:
            surv *= math.exp(-haz[i] * (mat - prev))
            prev = mat
        else:
            idx = i
            break
    dt = t - prev
    if dt > 0:
        surv *= math.exp(-haz[idx] * dt)
    return surv


def _cir_h(p):
    return ad.sqrt(p[0] * p[0] + 2.0 * p[2] * p[2])


def cir_A(p, tau):
    """cirpp.py:92-104."""
    kappa, theta, sigma = p[0], p[1], p[2]
    h = _cir_h(p)
    num = 2.0 * h * ad.exp(0.5 * (kappa + h) * tau)
    den = 2.0 * h + (kappa + h) * (ad.exp(h * tau) - 1.0)
    return (num / den) ** ((2.0 * kappa * theta) / (sigma * sigma))


def cir_B(p, tau):
    """cirpp.py:106-113."""
    h = _cir_h(p)
    e = ad.exp(h * tau) - 1.0
    return (2.0 * e) / (2.0 * h + (p[0] + h) * e)


def cir_psi(model, p, t):
    """cirpp.py:119-142."""
    kappa, theta, sigma, y0 = p
    h = _cir_h(p)
    et = ad.exp(h * t)
    den = 2.0 * h + (kappa + h) * (et - 1.0)
    Dt = (2.0 * kappa * theta / (sigma * sigma)) * (0.5 * (kappa + h) - (h * (kappa + h) * et) / den)
    Et = (4.0 * h * h * et) / (den * den)
    return cir_market_hazard(model, t) + Dt - y0 * Et


def cir_conditional_survival(model, p, t, T, y):
    """cirpp.py:246-285."""
    if model.deterministic:
        return ad.const_like(cir_market_survival(model, T) / cir_market_survival(model, t), y)
    y0 = p[3]
    pref = ((cir_market_survival(model, T) / cir_market_survival(model, t)) * (cir_A(p, t) / cir_A(p, T))
            * ad.exp(cir_B(p, T) * y0 - cir_B(p, t) * y0))
    # lam_t - psi_t == y_t
    return pref * (cir_A(p, T - t) * ad.exp(-(cir_B(p, T - t) * y)))


# ---------------------------------------------------------------------------- Vasicek helpers
def vasicek_bond(p, t1, t2, r):
    """vasicek.py:114-128 (A = exp(alpha); price = A * exp(-B r))."""
    sigma, theta, a = p[1], p[2], p[3]
    tau = t2 - t1
    B = (1.0 - ad.exp(-a * tau)) / a
    alpha = (theta - (sigma * sigma / (2.0 * a * a))) * (B - tau) - (sigma * sigma / (4.0 * a)) * B * B
    return ad.exp(alpha) * ad.exp(-(B * r))


def schwartz_curve(model, t):
    ts = [float(x) for x in model.curve_times]
    vs = [_f(x) for x in model.curve_values]
    if t <= ts[0]:
        return vs[0]
    if t >= ts[-1]:
        return vs[-1]
    hi = bisect_right(ts, t)
    lo = hi - 1
    return vs[lo] + (vs[hi] - vs[lo]) * ((t - ts[lo]) / (ts[hi] - ts[lo]))


# ---------------------------------------------------------------------------- state / stepping
def initial_state(model, p, n):
    """get_state of each model (black_scholes.py:41-42, heston.py:92-97, vasicek.py:45-50,
    cirpp.py:145-153, schwartz_two_factor.py:114-119, model_config.py:80-91)."""
    k = kind(model)
    ones = np.ones(n)
    if k == "ModelConfig":
        cols, off = [], 0
        for m in model.models:
            np_ = len(m.model_params)
            cols += initial_state(m, p[off:off + np_], n)
            off += np_
        return cols
    if k == "BlackScholesModel":
        return [p[0] * ones]
    if k == "BlackScholesMulti":
        na = len(model.asset_ids)
        return [p[i] * ones for i in range(na)]
    if k == "HestonModel":
        return [ad.log(p[0]) * ones, p[6] * ones]
    if k in ("VasicekModel", "HullWhiteModel"):
        return [p[0] * ones, 0.0 * p[0] * ones]
    if k == "CIRPPModel":
        y0 = cir_market_hazard(model, t0_of(model)) + 0.0 * p[3] if model.deterministic else p[3]
        return [y0 * ones, 0.0 * p[3] * ones]
    if k == "SchwartzTwoFactorModel":
        z = 0.0 * p[0] * ones
        return [math.log(schwartz_curve(model, t0_of(model))) + z, z, z]
    raise NotImplementedError(k)


def correlation(model, p, scheme):
    """Matrix (list of lists, entries scalar / Dual) of the noise correlation."""
    k = kind(model)
    if k == "ModelConfig" and scheme == "QE" and all(kind(m) == "HestonModel" for m in model.models):
        # BUILD-DEFINED EXTENSION, parity unpinned by the reference (its ModelConfig cannot hold a
        # HestonModel, model_config.py:106-115): spot normals of the assets correlated by the
        # inter-asset matrix, variance normals independent (QE, heston.py:85-90).
        A = len(model.models)
        C = [[1.0 if i == j else 0.0 for j in range(2 * A)] for i in range(2 * A)]
        idx = 0
        for i in range(A):
            for j in range(i + 1, A):
                rho = float(np.asarray(model.inter_asset_correlation_matrix[idx], dtype=float).reshape(-1)[0])
                C[2 * i][2 * j] = C[2 * j][2 * i] = rho
                idx += 1
        return C
    if k == "ModelConfig":
        n = len(model.asset_ids)
        C = [[0.0 for _ in range(n)] for _ in range(n)]
        row, idx, off = 0, 0, 0
        offs = []
        for m in model.models:
            offs.append(off)
            off += len(m.model_params)
        for i, m1 in enumerate(model.models):
            n1 = len(m1.asset_ids)
            blk = correlation(m1, p[offs[i]:offs[i] + len(m1.model_params)], scheme)
            if len(blk) != n1:
                raise RuntimeError("sub-model noise dimension != number of assets (model_config.py:106-115)")
            for a in range(n1):
                for b in range(n1):
                    C[row + a][row + b] = blk[a][b]
            col = row + n1
            for m2 in model.models[i + 1:]:
                n2 = len(m2.asset_ids)
                ic = np.asarray(model.inter_asset_correlation_matrix[idx], dtype=float)
                up = np.broadcast_to(ic, (n1, n2))
                lo = np.broadcast_to(ic.T if ic.ndim >= 2 else ic, (n2, n1))
                for a in range(n1):
                    for b in range(n2):
                        C[row + a][col + b] = float(up[a, b])
                        C[col + b][row + a] = float(lo[b, a])
                col += n2
                idx += 1
            row += n1
        return [[0.5 * (C[i][j] + C[j][i]) for j in range(n)] for i in range(n)]
    if k == "BlackScholesMulti":
        c = np.asarray(model.correlation_matrix, dtype=float)
        return [[float(c[i, j]) for j in range(c.shape[1])] for i in range(c.shape[0])]
    if k == "HestonModel":
        if scheme == "QE":
            return [[1.0, 0.0], [0.0, 1.0]]
        # heston.py:53-58: the matrix is built in the constructor, before requires_grad() - a constant of the graph
        rho = float(ad.val(p[3]))
        return [[1.0, rho], [rho, 1.0]]
    if k == "SchwartzTwoFactorModel":
        rho = float(ad.val(p[5]))       # schwartz_two_factor.py:59-65: same (the exact covariance reads rho live)
        return [[1.0, rho], [rho, 1.0]]
    return [[1.0]]


def covariance(model, p, dt):
    """Step covariance for the ANALYTICAL scheme (model.py:79-81 and overrides)."""
    k = kind(model)
    if k == "BlackScholesModel":
        return [[p[1] * p[1] * dt]]
    if k == "BlackScholesMulti":
        na = len(model.asset_ids)
        c = np.asarray(model.correlation_matrix, dtype=float)
        return [[p[na + i] * float(c[i, j]) * p[na + j] * dt for j in range(na)] for i in range(na)]
    if k in ("VasicekModel", "HullWhiteModel"):
        sigma, a = p[1], p[3]
        decay = ad.exp(-a * dt)
        return [[(sigma * sigma / (2.0 * a)) * (1.0 - decay * decay)]]
    if k == "SchwartzTwoFactorModel":
        kappa, ss, sl, rho = p[1], p[2], p[4], p[5]
        if abs(float(ad.val(kappa))) <= 1e-12:
            vs = ss * ss * dt
        else:
            vs = ss * ss * (1.0 - ad.exp(-2.0 * kappa * dt)) / (2.0 * kappa)
        vl = sl * sl * dt
        cov = rho * ad.sqrt(vs * vl)
        return [[vs, cov], [cov, vl]]
    if k == "ModelConfig":
        # block covariance; only BS-BS cross blocks are defined (model_config.py:193-221)
        n = len(model.asset_ids)
        C = [[0.0 for _ in range(n)] for _ in range(n)]
        row, idx, off = 0, 0, 0
        offs = []
        for m in model.models:
            offs.append(off)
            off += len(m.model_params)
        for i, m1 in enumerate(model.models):
            n1 = len(m1.asset_ids)
            p1 = p[offs[i]:offs[i] + len(m1.model_params)]
            blk = covariance(m1, p1, dt)
            for a in range(n1):
                for b in range(n1):
                    C[row + a][row + b] = blk[a][b]
            col = row + n1
            for jj, m2 in enumerate(model.models[i + 1:]):
                j = i + 1 + jj
                n2 = len(m2.asset_ids)
                if kind(m1) != "BlackScholesModel" or kind(m2) != "BlackScholesModel":
                    raise NotImplementedError("Inter covariance not implemented for the requested pair of models.")
                p2 = p[offs[j]:offs[j] + len(m2.model_params)]
                ic = float(np.asarray(model.inter_asset_correlation_matrix[idx], dtype=float).reshape(-1)[0])
                C[row][col] = p1[1] * p2[1] * ic * dt
                C[col][row] = C[row][col]
                col += n2
                idx += 1
            row += n1
        return [[0.5 * (C[i][j] + C[j][i]) for j in range(n)] for i in range(n)]
    n = noise_dim(model)
    return [[dt if i == j else 0.0 for j in range(n)] for i in range(n)]


def cholesky(a):
    """Lower Cholesky factor with scalar / Dual entries (torch.linalg.cholesky, model.py:56-72)."""
    n = len(a)
    L = [[0.0] * n for _ in range(n)]
    for i in range(n):
        for j in range(i + 1):
            s = a[i][j]
            for k in range(j):
                s = s - L[i][k] * L[j][k]
            L[i][j] = ad.sqrt(s) if i == j else s / L[j][j]
    return L


def correlate(z, L):
    """z @ L.T for z = list of d noise columns (model.py:46-48)."""
    d = len(L)
    out = []
    for i in range(d):
        acc = None
        for j in range(i + 1):
            term = L[i][j] * z[j]
            acc = term if acc is None else acc + term
        out.append(acc)
    return out


def step(model, p, scheme, t1, t2, state, w, u=None, smoothing=False):
    """One sub-step of `model`; `w` are this model's correlated noise columns."""
    k = kind(model)
    dt = t2 - t1
    sq = math.sqrt(dt)
    if k == "ModelConfig":
        out, so, no, po = [], 0, 0, 0
        for mi, m in enumerate(model.models):
            sd, nd, npar = state_dim(m), noise_dim(m), len(m.model_params)
            um = u[:, mi] if (u is not None and np.ndim(u) == 2) else u   # one QE uniform per asset
            out += step(m, p[po:po + npar], scheme, t1, t2, state[so:so + sd], w[no:no + nd], um, smoothing)
            so, no, po = so + sd, no + nd, po + npar
        return out
    if k == "BlackScholesModel":
        S, (spot, sigma, rate) = state[0], p
        if scheme == "ANALYTICAL":      # black_scholes.py:50-67
            return [S * ad.exp(rate * dt + (w[0] - 0.5 * dt * sigma ** 2))]
        return [S + (rate * S * dt + sigma * S * sq * w[0])]  # :69-85
    if k == "BlackScholesMulti":
        na = len(model.asset_ids)
        rate = p[2 * na]
        out = []
        for i in range(na):
            sig = p[na + i]
            if scheme == "ANALYTICAL":  # black_scholes_multi.py:63-79
                out.append(state[i] * ad.exp((rate - 0.5 * sig * sig) * dt + w[i]))
            else:                        # :81-96
                out.append(state[i] + (rate * state[i] * dt + sig * state[i] * sq * w[i]))
        return out
    if k in ("VasicekModel", "HullWhiteModel"):
        r, logB = state
        r0, sigma, theta, a = p
        if k == "HullWhiteModel":
            for tt, lv in zip(getattr(model, "mean_times", []), getattr(model, "mean_levels", [])):
                if t1 <= tt:
                    theta = lv
                    break
        logB_next = logB + r * dt
        if scheme == "ANALYTICAL":      # vasicek.py:52-86
            r_next = (theta + (r - theta) * ad.exp(-a * dt)) + w[0]
        else:                            # vasicek.py:88-112
            r_next = r + a * (theta - r) * dt + sigma * sq * w[0]
        return [r_next, logB_next]
    if k == "CIRPPModel":
        y, logB = state
        if model.deterministic:         # cirpp.py:155-172
            return [0.0 * y + cir_market_hazard(model, t2), logB + cir_market_hazard(model, t1) * dt]
        kappa, theta, sigma, y0 = p      # cirpp.py:174-198
        y_next = y + kappa * (theta - y) * dt + sigma * ad.sqrt(ad.clamp_min(y, 0.0)) * sq * w[0]
        logB_next = logB + (y + cir_psi(model, p, t1)) * dt
        return [ad.clamp_min(y_next, 1e-12), logB_next]
    if k == "SchwartzTwoFactorModel":
        _, x, yl = state
        rate, kappa, ss, mu, sl, rho = p
        if scheme == "ANALYTICAL":      # schwartz_two_factor.py:147-171
            x_mean = x if abs(float(ad.val(kappa))) <= 1e-12 else x * ad.exp(-kappa * dt)
            x_next = x_mean + w[0]
            y_next = yl + mu * dt + w[1]
        else:                            # :173-196
            x_next = x - kappa * x * dt + ss * sq * w[0]
            y_next = yl + mu * dt + sl * sq * w[1]
        return [math.log(schwartz_curve(model, t2)) + x_next + y_next, x_next, y_next]
    if k == "HestonModel":
        logS, v = state
        spot, sigma, rate, rho, kappa, theta, v0 = p
        if scheme == "EULER":           # heston.py:99-121
            vp = ad.sqrt(ad.clamp_min(v, 0.0))
            logS_next = logS + (rate - 0.5 * v) * dt + vp * sq * w[0]
            v_next = ad.clamp_min(v + kappa * (theta - v) * dt + sigma * vp * sq * w[1], 0.0)
            return [logS_next, v_next]
        if scheme == "QE":              # heston.py:161-253
            eps = 1e-12
            e = ad.exp(-kappa * dt)
            m = theta + (v - theta) * e
            s2 = v * sigma ** 2 * e * (1 - e) / kappa + theta * sigma ** 2 * (1 - e) ** 2 / (2 * kappa)
            psi = s2 / (m * m + eps)
            zV = w[1]
            invpsi = 1.0 / (psi + eps)
            tq = ad.clamp_min(2.0 * invpsi - 1.0, 0.0)
            b2 = ad.clamp_min(2.0 * invpsi - 1.0 + ad.sqrt(2.0 * invpsi) * ad.sqrt(tq), 0.0)
            b = ad.sqrt(b2)
            a_ = m / (1.0 + b2)
            v1 = a_ * (b + zV) ** 2
            pp = ad.clamp((psi - 1.0) / (psi + 1.0), 0.0, 1.0 - 1e-6)
            beta = (1.0 - pp) / (m + eps)
            one_minus_u = np.maximum(1.0 - u, eps)
            v_tail = ad.log(ad.clamp_min(1.0 - pp, eps) / one_minus_u) / (beta + eps)
            v2 = ad.fuzzy(u - pp, smoothing, 0.3) * v_tail
            wq = ad.fuzzy(psi - 1.5, smoothing, 0.5)
            v_next = (1.0 - wq) * v1 + wq * v2
            K0 = -rho * kappa * theta / sigma * dt
            K1 = (kappa * rho / sigma - 0.5) * dt - rho / sigma
            K2 = rho / sigma
            K3 = (1.0 - rho * rho) * dt
            var_int = ad.clamp_min(K3 * v + 0.0 * v_next, 0.0)
            vol = ad.sqrt(ad.clamp_min(var_int, eps))
            return [logS + rate * dt + K0 + K1 * v + K2 * v_next + vol * w[0], v_next]
    raise NotImplementedError(f"{k} / {scheme}")


# ---------------------------------------------------------------------------- requests
def route(model, p, asset_id):
    """(sub-model, its params, its state offset) for an asset id / 'numeraire' (model_config.py:285-307)."""
    if kind(model) != "ModelConfig":
        return model, p, 0
    idx = model.id_to_model[asset_id]
    so = sum(state_dim(m) for m in model.models[:idx])
    po = sum(len(m.model_params) for m in model.models[:idx])
    m = model.models[idx]
    return m, p[po:po + len(m.model_params)], so


def _rate(model, p):
    k = kind(model)
    if k == "BlackScholesModel":
        return p[2]
    if k == "BlackScholesMulti":
        return p[2 * len(model.asset_ids)]
    if k == "HestonModel":
        return p[2]
    if k == "SchwartzTwoFactorModel":
        return p[0]
    raise NotImplementedError(k)


def volatility(model, p, asset_id):
    """model.get_volatility() as the Brownian-bridge barrier uses it (barrier_option.py:149): defined for a single
    Black-Scholes model only."""
    if kind(model) == "BlackScholesModel":
        return p[1]
    raise NotImplementedError("Brownian-bridge barrier monitoring: single Black-Scholes model only")


def spot(model, p, asset_id, state):
    m, mp, so = route(model, p, asset_id)
    k = kind(m)
    if k == "BlackScholesMulti":
        return state[so + m.asset_ids.index(asset_id)]
    if k in ("HestonModel", "SchwartzTwoFactorModel"):
        return ad.exp(state[so])
    return state[so]  # BS spot, Vasicek r, CIR++ y


def numeraire(model, p, t, state):
    m, mp, so = route(model, p, "numeraire")
    if kind(m) in ("VasicekModel", "HullWhiteModel"):
        return ad.exp(state[so + 1])
    return ad.exp(_rate(m, mp) * (t - t0_of(m))) + 0.0 * state[so]


def forward(model, p, asset_id, t1, t2, state):
    """FORWARD_RATE request: Vasicek zero-bond P(t1,t2;r); BS-type growth factor exp(r (t2-t1))."""
    m, mp, so = route(model, p, asset_id)
    if kind(m) in ("VasicekModel", "HullWhiteModel"):
        return vasicek_bond(mp, t1, t2, state[so])
    return ad.exp(_rate(m, mp) * (t2 - t1)) + 0.0 * state[so]


def libor(model, p, asset_id, t1, t2, state):
    m, mp, so = route(model, p, asset_id)
    if kind(m) in ("VasicekModel", "HullWhiteModel"):
        return (1.0 / vasicek_bond(mp, t1, t2, state[so]) - 1.0) / (t2 - t1)
    return (ad.exp(_rate(m, mp) * (t2 - t1)) - 1.0) / (t2 - t1) + 0.0 * state[so]


def survival(model, p, asset_id, state):
    m, mp, so = route(model, p, asset_id)
    return ad.exp(-state[so + 1])


def conditional_survival(model, p, asset_id, t, T, state):
    m, mp, so = route(model, p, asset_id)
    return cir_conditional_survival(m, mp, t, T, state[so])
