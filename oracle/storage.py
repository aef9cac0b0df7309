"""Oracle restatement of the gas-storage product and of the controller loops that value it.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows the reference's tensors one for one in numpy: [N, S] state matrices, one torch-style gather per
lookup, one tall least squares per action date (src/products/storage.py:16-308, src/controller/controller.py:294-383
and 385-471).  The contract data (rate curves, reachable-inventory envelope) is read from the product's
StorageConfig, which tests/test_storage_host.py pins against the reference separately.
"""
import numpy as np

DATE_TOL = 1e-12


def _knots(slice_):
    return np.array([k.point for k in slice_], dtype=float), np.array([k.rate for k in slice_], dtype=float)


def interpolate_rate(point, slice_):
    """storage_helpers.py:96-127 (torch.bucketize(right=False) == searchsorted side='left';
    torch.isclose defaults rtol 1e-5, atol 1e-8)."""
    xp, fp = _knots(slice_)
    if len(xp) == 1:
        return np.full_like(point, fp[0])
    idx = np.searchsorted(xp, point, side="left")
    left = np.clip(idx - 1, 0, len(xp) - 2)
    x0, x1, y0, y1 = xp[left], xp[left + 1], fp[left], fp[left + 1]
    close = np.abs(x0 - x1) <= 1e-8 + 1e-5 * np.abs(x1)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.where(close, 0.0, (point - x0) / (x1 - x0))
    out = y0 + w * (y1 - y0)
    out = np.where(point <= xp[0], fp[0], out)
    return np.where(point >= xp[-1], fp[-1], out)


def transition(prod, date, next_date, action, state):
    """_transition_volume + _state_from_volume (storage.py:114-190) -> (next_state, volume difference)."""
    cfg, S = prod.storage_config, prod.num_states
    nxt, now = cfg.get_volume_constraint(next_date), cfg.get_volume_constraint(date)
    vol = now.vmin + state * cfg.grid_step(now.vmin, now.vmax, S)
    period = max(next_date - date, 0.0)
    if action == "inj":
        rate = interpolate_rate(vol, cfg.get_injection_flexibility_slice(date))
        new = np.minimum(vol + rate * period, nxt.vmax)
    elif action == "wd":
        rate = interpolate_rate(vol, cfg.get_withdrawal_flexibility_slice(date))
        new = np.maximum(vol - rate * period, nxt.vmin)
    else:
        new = np.minimum(np.maximum(vol, nxt.vmin), nxt.vmax)      # torch.clamp(min, max)
    scale = cfg.state_scale(nxt.vmin, nxt.vmax, S)
    nstate = np.zeros_like(new) if scale == 0.0 else (new - nxt.vmin) * scale
    return nstate, new - vol


def lookup(values, state, S):
    """lookup_state_values (storage.py:200-213): linear interpolation between the two neighbouring grid states."""
    b = np.clip(state.astype(float), 0.0, S - 1.0)
    lo, hi = np.floor(b).astype(np.int64), np.ceil(b).astype(np.int64)
    w = b - lo
    vlo, vhi = np.take_along_axis(values, lo, axis=1), np.take_along_axis(values, hi, axis=1)
    return vlo + w * (vhi - vlo)


def basis(x, n_basis):
    return np.stack([x ** k for k in range(n_basis)], axis=1)


def cashflows(prod, i, spot, numeraire, state, coeffs_i, n_basis, std=(0.0, 1.0), want_delta=False):
    """compute_normalized_cashflows (storage.py:215-308): state [N, B] -> (next state [N, B], cashflows / numeraire).
    std = (centre, inverse scale) of the basis variable (the reference's raw basis: (0, 1))."""
    S = prod.num_states
    date, next_date = float(prod.product_timeline[i]), float(prod.next_action_dates[i])
    cfg = prod.storage_config
    inj_s, inj_d = transition(prod, date, next_date, "inj", state)
    wd_s, wd_d = transition(prod, date, next_date, "wd", state)
    no_s, no_d = transition(prod, date, next_date, "no", state)
    sp = spot[:, None]
    ci, cw = cfg.get_variable_injection_cost(date), cfg.get_variable_withdrawal_cost(date)
    inj_pay = -inj_d * (sp + ci)
    wd_pay = -wd_d * (sp - cw)
    no_pay = -no_d * np.where(no_d >= 0.0, sp + ci, sp - cw)
    if next_date >= prod.end_date - DATE_TOL:
        c_inj = c_no = c_wd = 0.0
    else:
        grid = basis((spot - std[0]) * std[1], n_basis) @ coeffs_i.T            # [N, S]
        c_inj, c_no, c_wd = lookup(grid, inj_s, S), lookup(grid, no_s, S), lookup(grid, wd_s, S)
    values = np.stack([inj_pay + c_inj, no_pay + c_no, wd_pay + c_wd], axis=2)
    best = np.argmax(values, axis=2)[..., None]                # first maximum, like torch.argmax on these sizes
    nstate = np.take_along_axis(np.stack([inj_s, no_s, wd_s], axis=2), best, axis=2)[..., 0]
    cf = np.take_along_axis(np.stack([inj_pay, no_pay, wd_pay], axis=2), best, axis=2)[..., 0]
    if want_delta:
        return nstate, cf / numeraire[:, None], np.take_along_axis(np.stack([inj_d, no_d, wd_d], axis=2), best, axis=2)[..., 0]
    return nstate, cf / numeraire[:, None]


def gelsy(A, Y):
    """torch.linalg.lstsq on CPU = LAPACK gelsy (pivoted QR, rank cut at rcond = eps * max(m, n))."""
    from scipy.linalg import lstsq
    sol, *_ = lstsq(A, Y, cond=np.finfo(float).eps * max(A.shape), lapack_driver="gelsy", check_finite=False)
    return sol


def normal_equations(A, Y):
    """The other solver of the CUDA backend (mcre/storage.py, "moments"): Gram matrix and right-hand sides of a
    standardised basis, minimum-norm solve of the small system.  Not the reference's routine: used for path counts at
    which reading the design matrix back is not an option, and checked against this function."""
    sol, *_ = np.linalg.lstsq(A.T @ A, A.T @ Y, rcond=1e-12)
    return sol


def log_spot_std(model, t):
    """Standard deviation of log S(t) of the Schwartz two-factor model (variances of schwartz_two_factor.py:124-145
    accumulated from the calibration date)."""
    import math
    _, kappa, ss, _, sl, rho = [float(np.asarray(v)) for v in model.model_params]
    tau = max(t - float(np.asarray(model.calibration_date).reshape(-1)[0]), 0.0)
    if abs(kappa) <= 1e-12:
        vs, cov = ss * ss * tau, rho * ss * sl * tau
    else:
        vs = ss * ss * (1.0 - math.exp(-2.0 * kappa * tau)) / (2.0 * kappa)
        cov = rho * ss * sl * (1.0 - math.exp(-kappa * tau)) / kappa
    return math.sqrt(max(vs + sl * sl * tau + 2.0 * cov, 0.0))


def regress(prod, spots, numeraires, n_basis, solver=gelsy, std=None, expo=None):
    """_perform_regression_for_product for a Storage (controller.py:294-383):
    spots [n_dates][N], numeraires [n_dates] -> coeffs [n_dates][S, n_basis].
    The regression dates are the action dates and, with expo = (times, spots [n_expo][N], numeraires [n_expo]), the
    exposure dates; walking them backwards, the window between two regression dates holds at most one action date, whose
    cashflows pass through a float32 accumulator (controller.py:331, 342) before the float64 tail is added (:345-352).
    With `expo` -> (coeffs, exposure coeffs [n_expo][S, n_basis])."""
    ptl = [float(t) for t in prod.product_timeline]
    n_dates, S, n = len(ptl), prod.num_states, spots[0].shape[0]
    std = std or [(0.0, 1.0)] * n_dates
    coeffs = [np.zeros((S, n_basis)) for _ in range(n_dates)]
    e_times, e_spots, e_nums = expo if expo is not None else ([], [], [])
    e_coeffs = [np.zeros((S, n_basis)) for _ in e_times]
    reg_tl = sorted(set(ptl) | set(float(t) for t in e_times))
    e_index = {float(t): i for i, t in enumerate(e_times)}
    cache = {n_dates: np.zeros((n, S))}
    last = n_dates
    for t_reg in reversed(reg_tl):
        pidx = int(np.searchsorted(np.asarray(ptl), t_reg, side="left"))
        if pidx >= n_dates:
            continue
        t_next = pidx + 1 if ptl[pidx] == t_reg else pidx
        if t_next < last:
            sm = np.tile(np.arange(S, dtype=float), (n, 1))
            step = np.zeros((n, S), dtype=np.float32)
            for i in range(t_next, last):
                sm, cf = cashflows(prod, i, spots[i], np.full(n, numeraires[i]), sm, coeffs[i], n_basis, std[i])
                step = (step.astype(np.float64) + cf).astype(np.float32)
            cache[t_next] = step.astype(np.float64) + lookup(cache[last], sm, S)
            last = t_next
        total = cache[t_next]
        if t_reg in ptl:
            j = ptl.index(t_reg)
            A = basis((spots[j] - std[j][0]) * std[j][1], n_basis)
            sol = solver(A, numeraires[j] * total).T.copy()
            coeffs[j] = sol
            if t_reg in e_index:
                e_coeffs[e_index[t_reg]] = sol
        else:
            e = e_index[t_reg]
            e_coeffs[e] = solver(basis(e_spots[e], n_basis), e_nums[e] * total).T.copy()
    return coeffs if expo is None else (coeffs, e_coeffs)


def evaluate_with_exposures(prod, spots, numeraires, coeffs, n_basis, expo, e_coeffs):
    """_evaluate_product, exposure branch (controller.py:412-458): per exposure date the actions of all action dates up
    to it are taken first, then exposure = continuation polynomials of that date interpolated at the realised inventory
    state, over the numeraire.  expo = (times, spots, numeraires) -> (cashflows [N], exposures [n_expo][N])."""
    ptl = [float(t) for t in prod.product_timeline]
    S, n = prod.num_states, spots[0].shape[0]
    e_times, e_spots, e_nums = expo
    sm = np.full((n, 1), float(prod.get_initial_state()))
    cfs, t_start, out = np.zeros(n), 0, []
    for e, t in enumerate(e_times):
        while t_start < len(ptl) and ptl[t_start] <= t:
            sm, cf = cashflows(prod, t_start, spots[t_start], np.full(n, numeraires[t_start]), sm, coeffs[t_start], n_basis)
            cfs = cfs + cf[:, 0]
            t_start += 1
        grid = basis(e_spots[e], n_basis) @ e_coeffs[e].T
        out.append(lookup(grid, sm, S)[:, 0] / e_nums[e])
    while t_start < len(ptl):
        sm, cf = cashflows(prod, t_start, spots[t_start], np.full(n, numeraires[t_start]), sm, coeffs[t_start], n_basis)
        cfs = cfs + cf[:, 0]
        t_start += 1
    return cfs, out


def evaluate(prod, spots, numeraires, coeffs, n_basis, std=None, spot_tangents=None, numeraire_tangents=None):
    """_evaluate_product, PV-only branch (controller.py:399-410): realised cashflows of the regression policy.
    With spot_tangents [n_dates][P, N] and numeraire_tangents [n_dates][P] also the pathwise sensitivities [P, N] the
    reference's autograd yields (controller.py:609-627): torch.argmax / gather and the inventory moves carry no gradient,
    a cashflow -dV (S +- cost) / N differentiates through S and N only."""
    n = spots[0].shape[0]
    std = std or [(0.0, 1.0)] * len(prod.product_timeline)
    sm = np.full((n, 1), float(prod.get_initial_state()))
    cfs = np.zeros(n)
    tans = None if spot_tangents is None else np.zeros((spot_tangents[0].shape[0], n))
    for i in range(len(prod.product_timeline)):
        sm, cf, dv = cashflows(prod, i, spots[i], np.full(n, numeraires[i]), sm, coeffs[i], n_basis, std[i], want_delta=True)
        cfs = cfs + cf[:, 0]
        if tans is not None:
            tans = tans + (-dv[:, 0] / numeraires[i]) * spot_tangents[i] \
                - cf[:, 0] * (numeraire_tangents[i] / numeraires[i])[:, None]
    return cfs if tans is None else (cfs, tans)
