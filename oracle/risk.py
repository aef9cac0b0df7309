"""Oracle restatement of the reference's controller / products / netting / metrics.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows the reference's own data flow - full path tensor, per-date request resolution,
per-product cashflow rolls, tall least-squares - vectorised over paths in numpy.  It is
deliberately NOT organised like the CUDA implementation (no plan tables, no fused
passes), so that agreement between the two is meaningful.
"""
import math

import numpy as np

from oracle import ad
from oracle import engine as E
from oracle import models as M


def _f(x):
    return M._f(x)


def _fl(x):
    return [float(v) for v in np.asarray(x if not hasattr(x, "tolist") else x.tolist(), dtype=float).reshape(-1)]


def kind(obj):
    return type(obj).__name__


# ------------------------------------------------------------------------------------
# product descriptions
# ------------------------------------------------------------------------------------
def bond_schedule(startdate, maturity, tenor):
    """Payment dates and LIBOR periods by repeated accumulation (bond.py:37-68)."""
    dates, periods = [], []
    d = startdate + tenor
    while d < maturity:
        dates.append(d)
        periods.append((d - tenor, d))
        d += tenor
    dates.append(maturity)
    periods.append((d - tenor, maturity))
    return dates, periods


def asset_of(product):
    return product.asset_ids[0]


class Ctx:
    """Per-simulation context: simulated states per date + model access."""

    def __init__(self, model, p, timeline, paths, n, bridge=None):
        self.model, self.p, self.timeline, self.paths, self.n = model, p, timeline, paths, n
        self.idx = {float(t): i for i, t in enumerate(timeline)}
        #: bridge(product, n_intervals, second_barrier) -> [N, n_intervals] uniforms of the Brownian-bridge
        #: crossing draws of a barrier option (barrier_option.py:174, 200)
        self.bridge = bridge

    def state(self, t):
        return self.paths[self.idx[float(t)]]

    def numeraire(self, t):
        return M.numeraire(self.model, self.p, float(t), self.state(t))

    def spot(self, asset_id, t):
        return M.spot(self.model, self.p, asset_id, self.state(t))

    def libor(self, asset_id, t_obs, t1, t2):
        return M.libor(self.model, self.p, asset_id, t1, t2, self.state(t_obs))

    def forward(self, asset_id, t_obs, t1, t2):
        return M.forward(self.model, self.p, asset_id, t1, t2, self.state(t_obs))


def product_timeline(prod):
    return _fl(prod.product_timeline)


def modeling_timeline(prod):
    return _fl(prod.modeling_timeline)


def regression_timeline(prod):
    return _fl(prod.regression_timeline)


def num_states(prod):
    if kind(prod) == "FlexiCall":        # flexicall.py:50-54: states = rights left
        return int(prod.num_exercise_rights) + 1
    return 2 if kind(prod) in ("BermudanOption", "AmericanOption") else 1


def initial_state(prod):
    if kind(prod) == "FlexiCall":
        return int(prod.num_exercise_rights)
    return 1 if kind(prod) in ("BermudanOption", "AmericanOption") else 0


def bond_cashflow(bond, i, ctx):
    """Discounted coupon i of a bond (bond.py:165-214): no notional on coupons; LIBOR is read
    from the short rate at the payment date."""
    start, mat, tenor = _f(bond.startdate), _f(bond.maturity), _f(bond.tenor)
    dates, periods = bond_schedule(start, mat, tenor)
    prev = start if i == 0 else dates[i - 1]
    dt = dates[i] - prev
    numeraire = ctx.numeraire(dates[i])
    if bond.fixed_rate is not None:
        cf = _f(bond.fixed_rate) * dt + 0.0 * numeraire
    else:
        cf = ctx.libor(asset_of(bond), dates[i], periods[i][0], periods[i][1]) * dt
    if bond.pays_notional and i == len(dates) - 1:
        cf = cf + _f(bond.notional)
    return cf / numeraire


def underlying_value(und, t_obs, ctx):
    """Composite 'underlying value' request observed at t_obs (equity.py:32-40, bond.py:115-163,
    swap.py:129-140): the underlying is re-scheduled from the observation date and valued with
    forward zero-bonds P(t_obs, T_i; r(t_obs)), *with* notional."""
    k = kind(und)
    if k == "Equity":
        return ctx.spot(asset_of(und), t_obs)
    if k == "Bond":
        mat, tenor, notional = _f(und.maturity), _f(und.tenor), _f(und.notional)
        dates, periods = bond_schedule(t_obs, mat, tenor)
        a = asset_of(und)
        if und.fixed_rate is not None:
            total, prev = None, t_obs
            for d in dates:
                term = notional * _f(und.fixed_rate) * (d - prev) * ctx.forward(a, t_obs, t_obs, d)
                total = term if total is None else total + term
                prev = d
            if und.pays_notional:
                total = total + notional * ctx.forward(a, t_obs, t_obs, dates[-1])
            return total
        # floating: sum_i N (P(t, s_i) - P(t, s_{i+1})), s_i = accrual starts, last = maturity
        starts = [per[0] for per in periods] + [mat]
        total = None
        for i in range(len(dates)):
            term = notional * (ctx.forward(a, t_obs, t_obs, starts[i]) - ctx.forward(a, t_obs, t_obs, starts[i + 1]))
            total = term if total is None else total + term
        if und.pays_notional:
            total = total + notional * ctx.forward(a, t_obs, t_obs, mat)
        return total
    if k == "InterestRateSwap":
        fixed = underlying_value(_leg(und, True), t_obs, ctx)
        flt = underlying_value(_leg(und, False), t_obs, ctx)
        return flt - fixed if und.irs_type.name == "PAYER" else fixed - flt
    raise NotImplementedError(k)


class _LegView:
    """A swap leg seen as a bond description (swap.py:36-54)."""

    def __init__(self, swap, fixed):
        self.startdate, self.maturity, self.notional = swap.startdate, swap.enddate, swap.notional
        self.tenor = swap.tenor_fixed if fixed else swap.tenor_float
        self.fixed_rate = swap.fixed_rate if fixed else None
        self.pays_notional = False
        self.asset_ids = swap.asset_ids


_LegView.__name__ = "Bond"


def _leg(swap, fixed):
    return _LegView(swap, fixed)


def monitored_value(prod, t, ctx):
    """Spot monitored by a path-dependent option; with the `basket` extension (parity unpinned,
    the reference monitors one asset) the weighted arithmetic basket of several spots."""
    basket = getattr(prod, "basket", None)
    if basket is None:
        return ctx.spot(asset_of(prod), t)
    tot = None
    for a, w in zip(*basket):
        term = w * ctx.spot(a, t)
        tot = term if tot is None else tot + term
    return tot


def option_payoff(prod, s):
    k = _f(prod.strike)
    return ad.relu(s - k) if prod.option_type.name == "CALL" else ad.relu(k - s)


def cashflows(prod, i, ctx, state_matrix, regfn_degree, coeffs_of):
    """compute_normalized_cashflows(time_idx=i): (next_state [N,S] int, cashflows list of S columns)."""
    k = kind(prod)
    S = state_matrix.shape[1]
    if k == "Bond":
        return state_matrix, [bond_cashflow(prod, i, ctx)] * S
    if k == "InterestRateSwap":
        t = product_timeline(prod)[i]
        fixed, flt = _leg(prod, True), _leg(prod, False)
        fd, _ = bond_schedule(_f(fixed.startdate), _f(fixed.maturity), _f(fixed.tenor))
        ld, _ = bond_schedule(_f(flt.startdate), _f(flt.maturity), _f(flt.tenor))
        zero = 0.0 * ctx.numeraire(t)
        fcf = bond_cashflow(fixed, fd.index(t), ctx) if t in fd else zero
        lcf = bond_cashflow(flt, ld.index(t), ctx) if t in ld else zero
        cf = lcf - fcf if prod.irs_type.name == "PAYER" else fcf - lcf
        return state_matrix, [cf] * S
    if k == "EuropeanOption":          # european_option.py:45-68
        t = product_timeline(prod)[0]
        u = underlying_value(prod.underlying, t, ctx)
        return state_matrix, [option_payoff(prod, u) / ctx.numeraire(t)] * S
    if k == "BinaryOption":            # binary_option.py:37-42 (always fuzzy, eps = 1)
        t = product_timeline(prod)[0]
        ind = ad.fuzzy(ctx.spot(asset_of(prod), t) - _f(prod.strike), True, 1.0)
        pay = _f(prod.payment_amount) * (ind if prod.option_type.name == "CALL" else 1.0 - ind)
        return state_matrix, [pay / ctx.numeraire(t)] * S
    if k == "BasketOption":            # basket_option.py:55-78
        t = product_timeline(prod)[0]
        w = _fl(prod.weights)
        spots = [ctx.spot(a, t) for a in prod.asset_ids]

        def basket(geometric):
            if not geometric:
                tot = None
                for wi, s in zip(w, spots):
                    tot = wi * s if tot is None else tot + wi * s
                return tot
            tot = None
            for wi, s in zip(w, spots):
                term = wi * ad.log(s + 1e-10)
                tot = term if tot is None else tot + term
            return ad.exp(tot)

        geo = prod.basket_option_type.name == "GEOMETRIC"
        pay = option_payoff(prod, basket(geo))
        if prod.use_variation_reduction:
            pay = pay - option_payoff(prod, basket(True)) + closed_form_constant(prod, ctx.model, ctx.p)
        return state_matrix, [pay / ctx.numeraire(t)] * S
    if k == "AsianOption":             # asian_option.py:51-95
        obs = modeling_timeline(prod)
        spots = [monitored_value(prod, t, ctx) for t in obs]
        if prod.averaging_type.name == "GEOMETRIC":
            tot = None
            for s in spots:
                tot = ad.log(s + 1e-10) if tot is None else tot + ad.log(s + 1e-10)
            avg = ad.exp(tot / len(spots))
        else:
            tot = None
            for s in spots:
                tot = s if tot is None else tot + s
            avg = tot / len(spots)
        # numeraire request index len(product_timeline)-1 == 0 -> numeraire at the FIRST monitoring date
        return state_matrix, [option_payoff(prod, avg) / ctx.numeraire(obs[0])] * S
    if k == "BarrierOption":           # barrier_option.py:65-125, 300-314
        obs = modeling_timeline(prod)
        spots = [monitored_value(prod, t, ctx) for t in obs]
        mx, mn = spots[0], spots[0]
        for s in spots[1:]:
            mx = ad.where(ad.val(s) > ad.val(mx), s, mx)
            mn = ad.where(ad.val(s) < ad.val(mn), s, mn)
        pay = option_payoff(prod, spots[-1])

        def factor(barrier, btype):
            below = ad.fuzzy(barrier - mx, True, 0.05)
            above = ad.fuzzy(mn - barrier, True, 0.05)
            return {"UPANDOUT": below, "DOWNANDOUT": above, "UPANDIN": 1.0 - below, "DOWNANDIN": 1.0 - above}[btype.name]

        if getattr(prod, "use_brownian_bridge", False):
            # barrier_option.py:138-222: per interval the probability that the bridge between two monitored spots
            # crossed the barrier, exp(-2 ln(S_i / B) ln(S_i+1 / B) / (sigma^2 maturity / n_obs)), against one
            # uniform per (path, interval) - fuzzy (eps 0.05) like every indicator of this product
            # (all of it differentiable: the reference's autograd sees the spots and the volatility inside the crossing
            # probabilities and the band-limited indicators)
            sigma = M.volatility(ctx.model, ctx.p, asset_of(prod))
            dt_b = _f(prod.maturity) / len(obs)

            def nohit(barrier, second):
                u = ctx.bridge(prod, len(obs) - 1, second)
                keep = 1.0
                for i in range(len(obs) - 1):
                    prob = ad.exp(-2.0 * ad.log(spots[i] / barrier) * ad.log(spots[i + 1] / barrier) / (sigma ** 2 * dt_b))
                    keep = keep * (1.0 - ad.fuzzy(prob - u[:, i], True, 0.05))
                return keep

            def bfactor(barrier, btype, second):
                below = ad.fuzzy(barrier - mx, True, 0.05)
                above = ad.fuzzy(mn - barrier, True, 0.05)
                nh = nohit(barrier, second)
                return {"UPANDOUT": below * nh, "DOWNANDOUT": above * nh, "UPANDIN": (1.0 - below) * (1.0 - nh),
                        "DOWNANDIN": (1.0 - above) * (1.0 - nh)}[btype.name]

            pay = pay * bfactor(_f(prod.barrier1), prod.barrier_option_type1, False)
            if prod.barrier2 is not None and prod.barrier_option_type2 is not None:
                pay = pay * bfactor(_f(prod.barrier2), prod.barrier_option_type2, True)
            return state_matrix, [pay / ctx.numeraire(obs[0])] * S
        pay = pay * factor(_f(prod.barrier1), prod.barrier_option_type1)
        if prod.barrier2 is not None and prod.barrier_option_type2 is not None:
            pay = pay * factor(_f(prod.barrier2), prod.barrier_option_type2)
        return state_matrix, [pay / ctx.numeraire(obs[0])] * S
    if k in ("BermudanOption", "AmericanOption"):   # bermudan_option.py:93-188
        tl = product_timeline(prod)
        t = tl[i]
        u = underlying_value(prod.underlying, t, ctx)
        sign = 1.0 if prod.option_type.name == "CALL" else -1.0
        imm = ad.clamp_min(sign * (u - _f(prod.strike)), 0.0)
        x = ctx.spot(asset_of(prod), t)
        numeraire = ctx.numeraire(t)
        coeffs = None if i == len(tl) - 1 else coeffs_of(i)   # [S, degree]
        next_state = state_matrix.copy()
        cols = []
        for s in range(S):
            st = state_matrix[:, s]
            if coeffs is None:
                cont = 0.0 * ad.val(imm)
            else:
                cval = ad.val(coeffs) if isinstance(coeffs, ad.Dual) else coeffs
                basis = np.stack([ad.val(x) ** j for j in range(regfn_degree)], axis=1)
                cont = np.einsum("nj,nj->n", basis, cval[st])
            exercise = (ad.val(imm) > cont) & (st > 0)
            cols.append(imm * exercise.astype(np.float64) / numeraire)
            next_state[:, s] = np.where(exercise, np.where(st > 0, st - 1, st), st)
        return next_state, cols
    if k == "FlexiCall":               # flexicall.py:56-186
        tl = product_timeline(prod)
        t = tl[i]
        opt = prod.underlyings[i]
        u = underlying_value(opt.underlying, t, ctx)
        sign = 1.0 if prod.underlyings[0].option_type.name == "CALL" else -1.0
        imm = ad.clamp_min(sign * (u - _f(opt.strike)), 0.0)
        x = ctx.spot(asset_of(prod), t)
        numeraire = ctx.numeraire(t)
        coeffs = None if i == len(tl) - 1 else coeffs_of(i)   # [S, degree]
        next_state = state_matrix.copy()
        cols = []
        for s in range(S):
            st = state_matrix[:, s]
            after = np.where(st > 0, st - 1, st)               # state_after_exercise (flexicall.py:72-77)
            if coeffs is None:
                keep = exd = 0.0 * ad.val(imm)
            else:
                cval = ad.val(coeffs) if isinstance(coeffs, ad.Dual) else coeffs
                basis = np.stack([ad.val(x) ** j for j in range(regfn_degree)], axis=1)
                keep = np.einsum("nj,nj->n", basis, cval[st])
                exd = np.einsum("nj,nj->n", basis, cval[after])
            exercise = (ad.val(imm) + exd > keep) & (st > 0)    # flexicall.py:139-142
            cols.append(imm * exercise.astype(np.float64) / numeraire)
            next_state[:, s] = np.where(exercise, after, st)
        return next_state, cols
    raise NotImplementedError(k)


def closed_form_constant(prod, model, p):
    """The control variate's closed-form correction term (basket_option.py:72-78).  The reference
    evaluates it on the model's parameter tensors, so under AAD its parameter derivatives flow into
    the PV Greeks; here they come from torch.autograd on the same host formula."""
    if not any(isinstance(x, ad.Dual) for x in p):
        return _f(np.asarray(prod.compute_pv_analytically(model)))
    import torch
    saved = list(model.model_params)
    leaves = [q.detach().clone().requires_grad_(True) for q in saved]
    try:
        model.model_params = leaves
        val = prod.compute_pv_analytically(model).reshape(-1)[0]
        grads = torch.autograd.grad(val, leaves, allow_unused=True)
    finally:
        model.model_params = saved
    return ad.Dual(float(val.detach()), np.array([0.0 if g is None else float(g) for g in grads]))


def supports_analytic_exposure(prod, model):
    return kind(prod) == "EuropeanOption" and kind(model) in ("BlackScholesModel", "BlackScholesMulti")


def _norm_cdf(x):
    """Standard normal CDF, Dual-aware."""
    from math import sqrt as msqrt
    try:
        from scipy.special import erfc
    except Exception:  # pragma: no cover
        erfc = np.vectorize(math.erfc)
    if isinstance(x, ad.Dual):
        v = 0.5 * erfc(-x.v / msqrt(2.0))
        pdf = np.exp(-0.5 * x.v * x.v) / msqrt(2.0 * math.pi)
        return ad.Dual(v, pdf * x.t)
    return 0.5 * erfc(-np.asarray(x) / msqrt(2.0))


def bs_analytic_exposure(prod, model, p, t, spot, numeraire):
    """european_option.py:123-145 / 70-100."""
    ttm = _f(prod.exercise_date) - t
    if ttm <= 0.0:
        return 0.0 * spot
    if kind(model) == "BlackScholesMulti":
        na = len(model.asset_ids)
        i = model.asset_ids.index(asset_of(prod))
        sigma, rate = p[na + i], p[2 * na]
    else:
        sigma, rate = p[1], p[2]
    K = _f(prod.strike)
    vol_t = sigma * math.sqrt(ttm)
    d1 = (ad.log(spot / K) + (rate + 0.5 * sigma ** 2) * ttm) / vol_t
    d2 = d1 - vol_t
    disc = K * ad.exp(-rate * ttm)
    if prod.option_type.name == "CALL":
        price = spot * _norm_cdf(d1) - disc * _norm_cdf(d2)
    else:
        price = disc * _norm_cdf(-d2) - spot * _norm_cdf(-d1)
    return price / numeraire


# ------------------------------------------------------------------------------------
# regression
# ------------------------------------------------------------------------------------
def lstsq_dual(A, Y):
    """min-norm least squares c = pinv(A) Y with forward-mode tangents (pinv derivative for
    constant rank).  A: [N,p] (ndarray or Dual), Y: [N,S] -> c: [p,S].

    The values come from the raw basis [1, x, x^2] like the reference's torch.linalg.lstsq.  The TANGENTS are
    evaluated in the basis of the standardised variable u = (x - mean) / std (A' = A T with a constant T, c = T c'):
    the same function in exact arithmetic, but the residual term (A^T A)^-1 dA^T r loses ~7 digits to the 1e10
    condition number of the raw Gram matrix (x ~ 100), which is more than the 1e-6 the CUDA path is checked to."""
    Av, Yv = ad.val(A), ad.val(Y)
    c, *_ = np.linalg.lstsq(Av, Yv, rcond=None)
    if not isinstance(A, ad.Dual) and not isinstance(Y, ad.Dual):
        return c
    P = A.P if isinstance(A, ad.Dual) else Y.P
    dA, dY = ad.tan(A, P), ad.tan(Y, P)
    p = Av.shape[1]
    T = np.eye(p)
    if p >= 2 and all(np.allclose(Av[:, j], Av[:, 1] ** j, rtol=1e-12, atol=0.0) for j in range(p)):
        m, sd = float(np.mean(Av[:, 1])), float(np.std(Av[:, 1]))
        if sd > 1e-10 * max(abs(m), 1.0):      # (rank-deficient dates, all x equal: raw basis, min-norm like the reference)
            for j in range(p):
                for i in range(j + 1):
                    T[i, j] = math.comb(j, i) * (-m) ** (j - i) / sd ** j
    As = Av @ T                                # [N,p] standardised basis
    dAs = dA @ T
    cs = np.linalg.solve(T, c)                 # coefficients in the standardised basis
    Ap = np.linalg.pinv(As)                    # [p,N]
    resid = Yv - As @ cs
    proj = np.eye(p) - Ap @ As                 # I - A+ A
    ApT_c = Ap.T @ cs                          # [N,S]
    dc = np.empty((P,) + c.shape)
    for k in range(P):
        dc[k] = T @ (Ap @ (dY[k] - dAs[k] @ cs) + (Ap @ Ap.T) @ (dAs[k].T @ resid) + proj @ (dAs[k].T @ ApT_c))
    return ad.Dual(c, dc)


def stack_cols(cols):
    """list of S per-path columns -> [N,S] (Dual-aware)."""
    if any(isinstance(c, ad.Dual) for c in cols):
        P = next(c.P for c in cols if isinstance(c, ad.Dual))
        return ad.Dual(np.stack([ad.val(c) for c in cols], axis=1), np.stack([ad.tan(c, P) for c in cols], axis=2))
    return np.stack(cols, axis=1)


def to_f32(x):
    """Round values through float32 (tangents stay float64)."""
    if isinstance(x, ad.Dual):
        return ad.Dual(x.v.astype(np.float32).astype(np.float64), x.t)
    return np.asarray(x).astype(np.float32).astype(np.float64)


def gather_states(values, state_matrix):
    """values [N,S] gathered by state_matrix [N,S] (product.py:140-145)."""
    if isinstance(values, ad.Dual):
        v = np.take_along_axis(values.v, state_matrix, axis=1)
        t = np.take_along_axis(values.t, np.broadcast_to(state_matrix, values.t.shape), axis=2)
        return ad.Dual(v, t)
    return np.take_along_axis(values, state_matrix, axis=1)


def regress_product(prod, ctx, exposure_timeline, degree, store_product, store_exposure):
    """_perform_regression_for_product (controller.py:294-383) including the float32
    cashflow accumulators of lines 312-351."""
    reg_tl = sorted(set(regression_timeline(prod)) | set(exposure_timeline))
    ptl = product_timeline(prod)
    prod_reg = regression_timeline(prod)
    S, n = num_states(prod), ctx.n
    last = len(ptl)
    cache = {last: np.zeros((n, S))}
    for t_reg in reversed(reg_tl):
        pidx = int(np.searchsorted(np.asarray(ptl), t_reg, side="left"))
        if pidx >= len(ptl):
            continue
        t_next = pidx + 1 if ptl[pidx] == t_reg else pidx
        if t_next < last:
            sm = np.tile(np.arange(S), (n, 1))
            step_value = np.zeros((n, S))
            for i in range(t_next, last):
                sm, cols = cashflows(prod, i, ctx, sm, degree, lambda j: store_product["coeffs"][j])
                step_value = to_f32(step_value + stack_cols(cols))      # fp32 += fp64
            total = to_f32(step_value + gather_states(cache[last], sm))  # fp32 + fp32
            cache[t_next] = total
            last = t_next
        else:
            total = cache[t_next]
        numeraire = ctx.numeraire(t_reg)
        x = ctx.spot(asset_of(prod), t_reg)
        one = 0.0 * x + 1.0
        A = stack_cols([one if j == 0 else x ** j for j in range(degree)])
        num2 = stack_cols([numeraire] * S)
        coeffs = lstsq_dual(A, num2 * total)          # [degree, S]
        cT = ad.Dual(coeffs.v.T.copy(), np.transpose(coeffs.t, (0, 2, 1)).copy()) if isinstance(coeffs, ad.Dual) else coeffs.T.copy()
        if t_reg in prod_reg:
            store_product["coeffs"][prod_reg.index(t_reg)] = cT
        if t_reg in exposure_timeline:
            store_exposure[exposure_timeline.index(t_reg)] = cT


# ------------------------------------------------------------------------------------
# the run
# ------------------------------------------------------------------------------------
def threshold(x, h):
    """netting_set.py:48-72."""
    if h == 0.0:
        return x
    v = ad.val(x)
    return ad.where(v > h, x - h, ad.where(v < -h, x + h, 0.0 * x))


def mc_mean_and_error(x):
    """metric.py:26-35."""
    v = ad.val(x)
    n = v.shape[0]
    with np.errstate(invalid="ignore", divide="ignore"):
        err = np.std(v, ddof=1) / math.sqrt(n) if n > 1 else float("nan")
    return ad.mean(x), float(err)


def pfe_quantile_index(q, n):
    """pfe_metric.py:59: the product goes through a float32 tensor before the ceil."""
    return int(math.ceil(float(np.float32(q * n)))) - 1


def run(model, netting_sets, metrics, exposure_timeline, n_main, n_pre, num_steps, scheme,
        differentiate=False, draws_pre=None, draws_main=None, degree=3, storage_solver="gelsy", second_order=False):
    """Restatement of SimulationController.__init__ + run_simulation (controller.py:26-151, 663-709).
    Returns dict(results=[set][metric] -> [(value, err)], grads=[set][metric][eval] -> array[P] or None,
    coeffs=[product] -> [T_e] of [S,degree])."""
    scheme = scheme if isinstance(scheme, str) else scheme.name
    products = [p for ns in netting_sets for p in ns.products]
    prod_set = [i for i, ns in enumerate(netting_sets) for _ in ns.products]
    mtypes = [m.metric_type.name for m in metrics]
    need_expo = any(t != "PV" for t in mtypes)
    need_cfs = any(t == "PV" for t in mtypes)
    metric_tl = [float(t) for t in (exposure_timeline if exposure_timeline is not None else [])]
    expo = set(metric_tl)
    if need_expo:
        for ns in netting_sets:
            if ns.margin_period_of_risk is not None:
                expo |= {t - ns.margin_period_of_risk for t in np.asarray(metric_tl) if t - ns.margin_period_of_risk >= 0.0}
    expo_tl = sorted(float(t) for t in expo)
    expo_idx = {t: i for i, t in enumerate(expo_tl)}
    sim_tl = sorted({t for pr in products for t in modeling_timeline(pr)} | set(expo_tl))
    smoothing = bool(differentiate)
    pvals = M.param_values(model)
    p = ad.params(pvals, differentiate, second_order)   # second_order: compute_higher_derivatives (controller.py:253-255)
    P = len(pvals)

    def analytic_exposure_ok(pr):
        return all(t in ("PV", "EPE", "PFE") for t in mtypes) and supports_analytic_exposure(pr, model)

    def needs_regression(pr):
        if len(regression_timeline(pr)) > 0:
            return True
        return need_expo and not analytic_exposure_ok(pr)

    prod_coeffs = [dict(coeffs=[np.zeros((num_states(pr), degree)) for _ in regression_timeline(pr)]) for pr in products]
    expo_coeffs = [[np.zeros((num_states(pr), degree)) for _ in expo_tl] for pr in products]
    t0 = M.t0_of(model)
    n_sub = E.count_substeps(t0, sim_tl, num_steps)
    dim = M.noise_dim(model)
    qe = scheme == "QE"

    bridge_rngs = {}
    from oracle import storage as ST
    storage_coeffs = {}
    storage_expo_coeffs = {}
    n_storage = sum(kind(pr) == "Storage" for pr in products)
    if n_storage and differentiate and (n_storage < len(products) or need_expo):
        raise NotImplementedError("oracle: storages in mixed books / with exposure metrics are valued without sensitivities")

    def storage_market(pr, ctx, times=None):
        tl_ = product_timeline(pr) if times is None else times
        spots = [np.asarray(ad.val(ctx.spot(asset_of(pr), t)), dtype=float) for t in tl_]
        nums = [float(np.asarray(ad.val(ctx.numeraire(t))).reshape(-1)[0]) for t in tl_]
        return spots, nums

    def bridge_provider(draws, seed, n):
        """Uniforms of the Brownian-bridge draws: under Philox block pid * 4096 + interval of kind 2 (element 0 /
        1 = first / second barrier), like csrc/equity.cu; with the reference's draws each product's own
        numpy default_rng(12345), consumed call after call (barrier_option.py:49-50, 174, 200)."""
        philox_mode = isinstance(draws, E.PhiloxDraws) or draws is None

        def get(prod, n_int, second):
            pid = [id(q) for q in products].index(id(prod))
            if philox_mode:
                from oracle import philox
                paths_ = np.arange(n, dtype=np.uint64)
                cols = [philox._blocks(paths_, pid * 4096 + i, 2, seed, 0)[1 if second else 0] for i in range(n_int)]
                return np.stack(cols, axis=1)
            rng = bridge_rngs.setdefault(id(prod), np.random.default_rng(12345))
            return rng.uniform(0, 1, size=(n, n_int))
        return get

    if all(kind(pr) == "Storage" for pr in products) and not need_expo:
        return _run_storage(model, netting_sets, products, prod_set, mtypes, p, sim_tl, n_main, n_pre, num_steps, scheme,
                            draws_pre, draws_main, degree, n_sub, dim, storage_solver)

    if any(needs_regression(pr) for pr in products):
        if draws_pre is None:
            draws_pre = E.PhiloxDraws(42, n_pre, n_sub, dim, with_uniforms=qe)
        paths = E.generate_paths(model, p, sim_tl, n_pre, num_steps, scheme, draws_pre, smoothing)
        ctx = Ctx(model, p, sim_tl, paths, n_pre, bridge=bridge_provider(draws_pre, 42, n_pre))
        for k, pr in enumerate(products):
            if kind(pr) == "Storage":       # storages next to other products (pv_performance_large_netting_set.py)
                if need_expo:
                    expo = (expo_tl,) + storage_market(pr, ctx, expo_tl)
                    storage_coeffs[k], storage_expo_coeffs[k] = ST.regress(pr, *storage_market(pr, ctx), degree, expo=expo)
                else:
                    storage_coeffs[k] = ST.regress(pr, *storage_market(pr, ctx), degree)
            elif needs_regression(pr):
                regress_product(pr, ctx, expo_tl, degree, prod_coeffs[k], expo_coeffs[k])

    if draws_main is None:
        n_uni = len(M.submodels(model)) if qe else 1
        draws_main = E.PhiloxDraws(43, n_main, n_sub, dim, with_uniforms=qe, n_uniform=n_uni)
    paths = E.generate_paths(model, p, sim_tl, n_main, num_steps, scheme, draws_main, smoothing)
    ctx = Ctx(model, p, sim_tl, paths, n_main, bridge=bridge_provider(draws_main, 43, n_main))

    # ---- _evaluate_product (controller.py:385-471) -------------------------------------
    set_cfs = [None] * len(netting_sets)
    set_expo = [[None] * len(expo_tl) for _ in netting_sets]
    zero_path = 0.0 * ctx.numeraire(sim_tl[0])
    for k, pr in enumerate(products):
        si = prod_set[k]
        if kind(pr) == "Storage":
            if need_expo:
                expo = (expo_tl,) + storage_market(pr, ctx, expo_tl)
                cfs, expos = ST.evaluate_with_exposures(pr, *storage_market(pr, ctx), storage_coeffs[k], degree, expo,
                                                        storage_expo_coeffs[k])
                for i, e in enumerate(expos):
                    set_expo[si][i] = e if set_expo[si][i] is None else set_expo[si][i] + e
            else:
                cfs = ST.evaluate(pr, *storage_market(pr, ctx), storage_coeffs[k], degree)
            set_cfs[si] = cfs if set_cfs[si] is None else set_cfs[si] + cfs
            continue
        sm = np.full((n_main, 1), initial_state(pr), dtype=np.int64)
        ptl = product_timeline(pr)
        cfs, t_start = zero_path, 0
        expos = []
        coeffs_of = lambda j, k=k: prod_coeffs[k]["coeffs"][j]
        if not need_expo and need_cfs:
            while t_start < len(ptl):
                sm, cols = cashflows(pr, t_start, ctx, sm, degree, coeffs_of)
                cfs = cfs + cols[0]
                t_start += 1
        else:
            for i, t in enumerate(expo_tl):
                while t_start < len(ptl) and ptl[t_start] <= t:
                    sm, cols = cashflows(pr, t_start, ctx, sm, degree, coeffs_of)
                    cfs = cfs + cols[0]
                    t_start += 1
                numeraire = ctx.numeraire(t)
                x = ctx.spot(asset_of(pr), t)
                if analytic_exposure_ok(pr):
                    e = bs_analytic_exposure(pr, model, p, t, x, numeraire)
                else:
                    c = expo_coeffs[k][i]                       # [S, degree]
                    st = sm[:, 0]
                    if isinstance(c, ad.Dual):
                        csel = ad.Dual(c.v[st], c.t[:, st])      # [N,degree], [P,N,degree]
                    else:
                        csel = c[st]
                    cont = None
                    for j in range(degree):
                        col = csel[:, j] if not isinstance(csel, ad.Dual) else ad.Dual(csel.v[:, j], csel.t[:, :, j])
                        term = col * (x ** j if j > 0 else 1.0)
                        cont = term if cont is None else cont + term
                    e = cont / numeraire
                expos.append(e)
            if need_cfs:
                while t_start < len(ptl):
                    sm, cols = cashflows(pr, t_start, ctx, sm, degree, coeffs_of)
                    cfs = cfs + cols[0]
                    t_start += 1
        set_cfs[si] = cfs if set_cfs[si] is None else set_cfs[si] + cfs
        for i, e in enumerate(expos):
            set_expo[si][i] = e if set_expo[si][i] is None else set_expo[si][i] + e

    # ---- netting sets + metrics (controller.py:506-563, netting_set.py:156-184) ---------
    results, grads, hess = [], [], []
    for si, ns in enumerate(netting_sets):
        unsec = []
        if need_expo:
            for m, t in enumerate(metric_tl):
                e = set_expo[si][expo_idx[t]]
                if ns.margin_period_of_risk is None:
                    unsec.append(threshold(e, ns.threshold))
                else:
                    td = float(np.asarray(metric_tl)[m] - ns.margin_period_of_risk)
                    coll = 0.0 * e if td < 0.0 else threshold(set_expo[si][expo_idx[td]], ns.threshold)
                    unsec.append(e - coll)
        mres, mgrads, mhess = [], [], []
        for metric, mt in zip(metrics, mtypes):
            if (mt == "CVA" and ns.counterparty_id is not None
                    and getattr(metric, "counterparty_id", None) != ns.counterparty_id):
                vals = [(0.0, 0.0)]
            elif mt == "PV":
                vals = [mc_mean_and_error(set_cfs[si])]
            elif mt == "CE":
                vals = [mc_mean_and_error(ad.relu(unsec[0]))]
            elif mt == "EPE":
                vals = [mc_mean_and_error(ad.relu(e)) for e in unsec]
            elif mt == "ENE":
                vals = [mc_mean_and_error(-ad.relu(-e)) for e in unsec]
            elif mt == "EEPE":
                ee = [ad.mean(ad.relu(e)) for e in unsec]
                tot = None
                for v in ee:
                    tot = v if tot is None else tot + v
                ev = np.array([float(ad.val(v)) for v in ee])
                with np.errstate(invalid="ignore", divide="ignore"):
                    err = float(np.std(ev, ddof=1) / math.sqrt(len(ev))) if len(ev) > 1 else float("nan")
                vals = [(tot / len(ee), err)]
            elif mt == "PFE":
                q = metric.quantile
                qi = pfe_quantile_index(q, n_main)
                vals = []
                for e in unsec:
                    order = np.argsort(ad.val(e), kind="stable")
                    sv = ad.val(e)[order]
                    pfe = e[order[qi]] if isinstance(e, ad.Dual) else sv[qi]
                    if qi == 0 or qi == n_main - 1 or (sv[qi - 1] == sv[qi] and sv[qi + 1] == sv[qi]):
                        se = 0.0
                    else:
                        f = max((sv[qi + 1] - sv[qi - 1]) / 2.0, 1e-6)
                        se = math.sqrt(q * (1 - q) / (n_main * f * f))
                    vals.append((pfe, se))
            elif mt == "CVA":
                cp = metric.counterparty_id
                tot = zero_path
                for kk in range(len(metric_tl) - 1):
                    st = ctx.state(metric_tl[kk])
                    surv = M.survival(model, p, cp, st)
                    cond = M.conditional_survival(model, p, cp, metric_tl[kk], metric_tl[kk + 1], st)
                    tot = tot + ad.relu(unsec[kk]) * (surv * (1.0 - cond))
                vals = [mc_mean_and_error(tot * (1.0 - metric.recovery_rate))]
            else:
                raise NotImplementedError(mt)
            mres.append([(float(ad.val(v)), float(e)) for v, e in vals])
            mgrads.append([np.array(v.t, dtype=float).reshape(P) if isinstance(v, ad.Dual) else None for v, _ in vals])
            mhess.append([np.array(v.h, dtype=float).reshape(P, P) if isinstance(v, ad.Dual) and v.h is not None else None
                          for v, _ in vals])
        results.append(mres)
        grads.append(mgrads)
        hess.append(mhess)
    return dict(results=results, grads=grads, hess=hess, expo_coeffs=expo_coeffs, prod_coeffs=prod_coeffs,
                sim_timeline=sim_tl, exposure_timeline=expo_tl, n_sub=n_sub, noise_dim=dim)


def _run_storage(model, netting_sets, products, prod_set, mtypes, p, sim_tl, n_main, n_pre, num_steps, scheme,
                 draws_pre, draws_main, degree, n_sub, dim, solver="gelsy"):
    """Books of gas storages, PV only (oracle/storage.py; controller.py:294-383 for the regression pass,
    :399-410 for the valuation pass)."""
    from oracle import storage as ST
    if any(kind(pr) != "Storage" for pr in products) or any(t != "PV" for t in mtypes):
        raise NotImplementedError("oracle: storages are valued on their own, PV only")
    if draws_pre is None:
        draws_pre = E.PhiloxDraws(42, n_pre, n_sub, dim)
    if draws_main is None:
        draws_main = E.PhiloxDraws(43, n_main, n_sub, dim)
    ctx_pre = Ctx(model, p, sim_tl, E.generate_paths(model, p, sim_tl, n_pre, num_steps, scheme, draws_pre), n_pre)
    ctx_main = Ctx(model, p, sim_tl, E.generate_paths(model, p, sim_tl, n_main, num_steps, scheme, draws_main), n_main)
    set_cfs = [None] * len(netting_sets)
    set_tan = [np.zeros(len(M.param_values(model))) for _ in netting_sets]
    all_coeffs = []
    for k, pr in enumerate(products):
        tl = product_timeline(pr)
        asset = asset_of(pr)

        def market(ctx):
            spots = [np.asarray(ad.val(ctx.spot(asset, t)), dtype=float) for t in tl]
            nums = [float(np.asarray(ad.val(ctx.numeraire(t))).reshape(-1)[0]) for t in tl]
            return spots, nums
        std = None
        if solver == "moments":      # standardised basis (spot - F(t)) / (F(t) sd(log S(t))), mcre/storage.py
            std = []
            for t in tl:
                f = M.schwartz_curve(model, t)
                sd = f * ST.log_spot_std(model, t)
                std.append((f, 1.0 / sd if sd > 1e-300 else 0.0))
        coeffs = ST.regress(pr, *market(ctx_pre), degree, solver=ST.gelsy if solver == "gelsy" else ST.normal_equations,
                            std=std)
        P = len(M.param_values(model))
        if isinstance(p[0], ad.Dual):
            s_t = [np.broadcast_to(ad.tan(ctx_main.spot(asset, t), P), (P, n_main)) for t in tl]
            n_t = [np.asarray(ad.tan(ctx_main.numeraire(t), P), dtype=float).reshape(P, -1)[:, 0] for t in tl]
            cfs, tans = ST.evaluate(pr, *market(ctx_main), coeffs, degree, std=std, spot_tangents=s_t, numeraire_tangents=n_t)
            set_tan[prod_set[k]] = set_tan[prod_set[k]] + tans.mean(axis=1)
        else:
            cfs = ST.evaluate(pr, *market(ctx_main), coeffs, degree, std=std)
        all_coeffs.append(coeffs)
        si = prod_set[k]
        set_cfs[si] = cfs if set_cfs[si] is None else set_cfs[si] + cfs
    results = [[[tuple(float(v) for v in mc_mean_and_error(set_cfs[si]))] for _ in mtypes] for si in range(len(netting_sets))]
    grads = [[[set_tan[si] if isinstance(p[0], ad.Dual) else None] for _ in mtypes] for si in range(len(netting_sets))]
    return dict(results=results, grads=grads, prod_coeffs=all_coeffs, sim_timeline=sim_tl, n_sub=n_sub, noise_dim=dim)
