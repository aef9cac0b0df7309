"""Host side of the gas storage (products/storage.py, products/storage_helpers.py, mcre/storage.py) against the
reference's own outputs (tests/golden/storage_envelope.json, made by tests/golden/make_storage_envelope.py with the
unmodified reference).  CPU only: contract logic and table lowering, no simulation."""
import json
import math
import os

import numpy as np
import pytest
import torch

import cases
import parity_helpers as helpers


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(helpers.GOLDEN_DIR, "storage_envelope.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("which", ["storage1", "storage2"])
def test_envelope_rates_and_costs_equal_the_reference_bit_for_bit(golden, which):
    g = golden[which]
    ns = cases.Namespace()
    _, sets, _, _ = cases.storage_s2f(ns, which=which)
    st = sets[0].products[0]
    cfg = st.storage_config
    env = [[w.start_date, w.end_date, w.vmin, w.vmax] for w in cfg.volume_constraints]
    assert env == g["envelope"]                      # the bisection of the envelope is reproduced operation by operation
    assert st.product_timeline.tolist() == g["action_dates"]
    assert st.next_action_dates.tolist() == g["next_dates"]
    levels = torch.tensor(g["levels"], dtype=torch.float64)
    for t, inj, wd, cost in zip(g["dates"], g["injection"], g["withdrawal"], g["costs"]):
        assert cfg.interpolate_rate_tensor(levels, cfg.get_injection_flexibility_slice(t)).tolist() == inj
        assert [cfg.get_withdrawal_flexibility_rate(t, float(v)) for v in g["levels"]] == wd
        assert [cfg.get_variable_injection_cost(t), cfg.get_variable_withdrawal_cost(t)] == cost
    assert [[cfg.grid_step(0.0, 90.0, 10), cfg.state_scale(0.0, 90.0, 10)],
            [cfg.grid_step(5.0, 5.0, 10), cfg.state_scale(5.0, 5.0, 10)]] == g["grid"]
    assert st.state_to_volume(200.0, torch.tensor([0.0, 2.5, 9.0])).tolist() == g["volume_of_state"]


def test_lowered_date_records_restate_the_contract():
    ns = cases.Namespace()
    _, sets, _, _ = cases.storage_s2f(ns, which="storage2")
    st = sets[0].products[0]
    cfg = st.storage_config
    rec = st.lower()
    acts, nxt = st.product_timeline.tolist(), st.next_action_dates.tolist()
    assert rec.shape == (len(acts), 48)
    for i in (0, 1, 180, 181, 272, 273, 400, len(acts) - 1):
        now, after = cfg.get_volume_constraint(acts[i]), cfg.get_volume_constraint(nxt[i])
        assert rec[i, 0] == now.vmin and rec[i, 1] == cfg.grid_step(now.vmin, now.vmax, 10)
        assert (rec[i, 2], rec[i, 3]) == (after.vmin, after.vmax)
        assert rec[i, 4] == cfg.state_scale(after.vmin, after.vmax, 10)
        assert rec[i, 5] == nxt[i] - acts[i] and (rec[i, 6], rec[i, 7]) == (0.35, 0.12)
        knots = cfg.get_injection_flexibility_slice(acts[i])
        assert rec[i, 8] == len(knots) and [rec[i, 16 + 2 * k] for k in range(len(knots))] == [k.point for k in knots]
        assert rec[i, 10] == (1.0 if i == len(acts) - 1 else 0.0)


def test_invalid_contracts_raise_like_the_reference():
    ns = cases.Namespace()
    cfg = ns.StorageConfig()
    with pytest.raises(ValueError):
        cfg.get_volume_constraint(0.0)
    cfg.add_volume_constraint(0.0, 10.0, 0.0, 10.0)
    cfg.add_injection_flexibility(0.0, 10.0, 0.0, 1.0)
    cfg.add_withdrawal_flexibility(0.0, 10.0, 0.0, 1.0)
    with pytest.raises(ValueError):
        ns.Storage("gas", 0.0, 5.0, 0.0, cfg, num_states=1)
    with pytest.raises(ValueError):
        ns.Storage("gas", 0.0, 5.0, 0.0, cfg, num_states=4, rollout_interval=0.0)
    # the initial inventory lies outside the contract's band: no feasible envelope (storage_helpers.py:418-431)
    with pytest.raises(ValueError):
        ns.Storage("gas", 0.0, 5.0, 50.0, cfg, num_states=4)


@pytest.mark.parametrize("scheme", ["ANALYTICAL", "EULER"])
def test_step_table_reproduces_the_oracle_model_step(scheme):
    """mcre/storage.py:step_table folded into csrc/storage.cu's two-factor step = the oracle's Schwartz step."""
    from mcre.storage import step_table
    from mcre.timegrid import build_time_grid
    from oracle import engine as E, models as M, ad
    ns = cases.Namespace()
    model, sets, _, _ = cases.storage_s2f(ns, which="storage2", end_day=30)
    tl = sets[0].products[0].product_timeline.tolist()
    grid = build_time_grid(0.0, tl, 2)
    tab, _ = step_table(model, grid, getattr(ns.SimulationScheme, scheme))
    rng = np.random.default_rng(5)
    z = rng.standard_normal((grid.n_sub, 64, 2))
    p = ad.params(M.param_values(model), False)
    paths = E.generate_paths(model, p, tl, 64, 2, scheme, E.InjectedDraws(z))
    x = np.zeros(64)
    y = np.zeros(64)
    for s in range(grid.n_sub):
        a, k, dt, m, cx, cy, lf = tab[s, :7]
        bx, by = tab[s, 8:10], tab[s, 16:18]
        x = (a * x - (k * x) * dt) + cx * (bx[0] * z[s, :, 0] + bx[1] * z[s, :, 1])
        y = (y + m) + cy * (by[0] * z[s, :, 0] + by[1] * z[s, :, 1])
        d = grid.date_after[s]
        if d >= 0:
            np.testing.assert_allclose(lf + x + y, np.asarray(paths[d][0]), rtol=0, atol=1e-13)


def test_unsupported_storage_runs_raise_before_any_device_work():
    ns = cases.Namespace()
    model, sets, metrics, _ = cases.storage_s2f(ns, which="storage1")
    from mcre.storage import StorageBackend

    def ctrl(**kw):
        args = dict(netting_sets=sets, model=model, risk_metrics=ns.RiskMetrics(metrics), num_paths_mainsim=256,
                    num_paths_presim=256, num_steps=1, simulation_scheme=ns.SimulationScheme.ANALYTICAL)
        args.update(kw)
        return ns.SimulationController(**args)
    assert StorageBackend(ctrl(regression_function=ns.PolyomialRegression(3))).mode == "lapack"
    assert StorageBackend(ctrl(num_paths_presim=1 << 17)).mode == "moments"
    assert StorageBackend(ctrl(differentiate=True)).nt == 6
    assert StorageBackend(ctrl(model=ns.BlackScholesModel(0.0, 100.0, 0.0, 0.2, asset_id="thegasprice"))).noise_dim == 1
    with pytest.raises(NotImplementedError):
        StorageBackend(ctrl(regression_function=ns.PolyomialRegression(7)))
    with pytest.raises(NotImplementedError):
        StorageBackend(ctrl(model=ns.HestonModel(0.0, 100.0, 0.03, 0.4, -0.7, 2.0, 0.04, 0.04)))
    assert StorageBackend(ctrl(risk_metrics=ns.RiskMetrics([ns.EPEMetric()], exposure_timeline=[0.0, 1.0]))).need_expo
    with pytest.raises(NotImplementedError):     # sensitivities of exposure metrics of a storage
        StorageBackend(ctrl(risk_metrics=ns.RiskMetrics([ns.EPEMetric()], exposure_timeline=[0.0, 1.0]), differentiate=True))


def test_step_tangent_table_equals_finite_differences_of_the_step_table():
    """d(A, B00, M, B10, B11, log F)/d(parameter) from the host duals against central differences of the value table."""
    from mcre.storage import step_table
    from mcre.timegrid import build_time_grid
    ns = cases.Namespace()
    for kind, scheme in (("s2f", "ANALYTICAL"), ("s2f", "EULER"), ("bs", "ANALYTICAL")):
        model, sets, _, _ = cases.storage_small(ns, model_kind=kind, end_day=5.0)
        grid = build_time_grid(0.0, sets[0].products[0].product_timeline.tolist(), 2)
        sch = getattr(ns.SimulationScheme, scheme)
        nt = len(model.model_params)
        _, tan = step_table(model, grid, sch, nt)

        def eff(m):
            t, _ = step_table(m, grid, sch)
            return np.stack([t[:, 0] - t[:, 1] * t[:, 2], t[:, 4] * t[:, 8], t[:, 3], t[:, 5] * t[:, 16], t[:, 5] * t[:, 17], t[:, 6]], axis=1)
        for k in range(nt):
            if k in model.unconnected_params(sch):
                assert np.all(tan[:, k, :] == 0.0)
                continue
            v0 = float(model.model_params[k])
            h = 1e-6 * max(abs(v0), 1e-3)
            model.model_params[k] = torch.tensor(v0 + h, dtype=torch.float64)
            up = eff(model)
            model.model_params[k] = torch.tensor(v0 - h, dtype=torch.float64)
            dn = eff(model)
            model.model_params[k] = torch.tensor(v0, dtype=torch.float64)
            np.testing.assert_allclose(tan[:grid.n_sub, k, :], (up - dn)[:grid.n_sub] / (2 * h), rtol=2e-6, atol=1e-9)


# ---- the reference's own unit tests of the inventory moves (tests/pytests/test_storage.py:19-113), restated ----------
def _window_storage(ns, shifting=False):
    cfg = ns.StorageConfig()
    if shifting:
        for a, b, lo, hi in ((0.0, 2.0, 0.0, 12.0), (2.0, 3.0, 0.0, 12.0), (3.0, 4.0, 3.0, 9.0)):
            cfg.add_volume_constraint(a, b, lo, hi, 0.0)
        cfg.add_injection_flexibility(0.0, 4.0, 0.0, 3.0)
        cfg.add_withdrawal_flexibility(0.0, 4.0, 0.0, 3.0)
        cfg.add_variable_injection_cost(0.0, 0.0)
        cfg.add_variable_withdrawal_cost(0.0, 0.0)
        return ns.Storage("thegasprice", 0.0, 4.0, 6.0, cfg, num_states=4)
    cfg.add_volume_constraint(0.0, 4.0, 0.0, 12.0, 0.0)
    for level, rate in ((0.0, 3.0), (6.0, 1.5)):
        cfg.add_injection_flexibility(0.0, 4.0, level, rate)
    for level, rate in ((0.0, 1.0), (6.0, 2.5)):
        cfg.add_withdrawal_flexibility(0.0, 4.0, level, rate)
    cfg.add_variable_injection_cost(0.0, 1.0)
    cfg.add_variable_withdrawal_cost(0.0, 1.0)
    return ns.Storage("thegasprice", 0.0, 4.0, 4.0, cfg, num_states=4)


def test_reference_unit_tests_of_the_inventory_moves():
    from products.storage import StorageAction
    ns = cases.Namespace()
    st = _window_storage(ns)
    states = torch.tensor([0.0, 1.0, 2.0, 3.0], dtype=torch.float64)
    now = st.state_to_volume(1.0, states)
    nxt_s = st.compute_next_state(1.0, 2.0, StorageAction.INJECTION)(states)
    nxt_v = st.state_to_volume(2.0, nxt_s)
    assert torch.all(nxt_s[1:] >= nxt_s[:-1]) and torch.all(nxt_v >= now)
    assert torch.allclose(nxt_v, torch.tensor([4.5, 5.5, 6.5, 7.5], dtype=torch.float64), atol=1e-10, rtol=0.0)
    assert nxt_v[-1].item() == st.storage_config.get_volume_constraint(2.0).vmax
    for action in StorageAction:
        s2 = st.compute_next_state(1.0, 2.0, action)(states)
        delta = st.compute_volume_difference(1.0, 2.0, action)(states)
        assert torch.allclose(delta, st.state_to_volume(2.0, s2) - now, atol=1e-10, rtol=0.0)
    sh = _window_storage(ns, shifting=True)
    held = sh.compute_next_state(2.0, 3.0, StorageAction.DO_NOTHING)(states)
    assert torch.allclose(sh.state_to_volume(3.0, held), torch.tensor([3.0, 4.0, 8.0, 9.0], dtype=torch.float64), atol=1e-10, rtol=0.0)
    assert held[1].item() == 0.5 and held[2].item() == 2.5
    grid = torch.tensor([[0.0, 10.0, 20.0, 30.0]], dtype=torch.float64)
    assert sh.lookup_state_values(grid, torch.tensor([[0.5, 2.5, 7.0]], dtype=torch.float64)).tolist() == [[5.0, 25.0, 30.0]]


def test_step_table_of_a_multi_asset_model_reproduces_the_oracle_model_step():
    """A storage on asset 2 of a three-asset Black-Scholes model: the asset's row of the Cholesky factor of the joint step
    covariance applied to the joint draw = the oracle's multi-asset step (black_scholes_multi.py:63-79)."""
    from mcre.storage import step_table
    from mcre.timegrid import build_time_grid
    from oracle import engine as E, models as M, ad
    ns = cases.Namespace()
    model, sets, _, _ = cases.storage_mixed_book(ns)
    st = sets[0].products[3]
    tl = st.product_timeline.tolist()[1:]
    grid = build_time_grid(0.0, tl, 2)
    tab, _ = step_table(model, grid, ns.SimulationScheme.ANALYTICAL, asset_id=st.get_asset_id())
    rng = np.random.default_rng(11)
    z = rng.standard_normal((grid.n_sub, 32, 3))
    p = ad.params(M.param_values(model), False)
    paths = E.generate_paths(model, p, tl, 32, 2, "ANALYTICAL", E.InjectedDraws(z))
    x = np.zeros(32)
    y = np.zeros(32)
    for s in range(grid.n_sub):
        a, k, dt, m, cx, cy, lf = tab[s, :7]
        w0 = sum(tab[s, 8 + j] * z[s, :, j] for j in range(3))
        x = (a * x - (k * x) * dt) + cx * w0
        y = (y + m) + cy * 0.0
        d = grid.date_after[s]
        if d >= 0:
            np.testing.assert_allclose(np.exp(lf + x + y), np.asarray(paths[d][2]), rtol=1e-13, atol=0)
