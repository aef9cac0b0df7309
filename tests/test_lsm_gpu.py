"""Parity of the Longstaff-Schwartz path (csrc/lsm.cu, Bermudan units of csrc/irc_main.cuh)
through the C ABI: Vasicek Bermudan payer swaption, EPE + PFE + PV (BASELINE config 4).

  1. the reference's torch.randn stream injected -> outputs of the unmodified reference
     (tests/golden/bermudan_swaption.json) at 1e-9 relative (the exercise indicator is hard, so
     agreement at this level means every path took the same exercise decision)
  2. native Philox vs the oracle on the same stream
"""
import numpy as np
import pytest

import cases
import parity_helpers as helpers

pytestmark = pytest.mark.gpu


def _compare(flat_a, flat_b, rtol, what, err_rtol=1e-7):
    for key, (va, ea) in flat_a.items():
        vb, eb = flat_b[key]
        scale = max(1.0, float(np.nanmax(np.abs(vb))) if len(vb) else 1.0)
        helpers.assert_close(va, vb, rtol, rtol * scale, f"{what} {key} value")
        helpers.assert_close(ea, eb, err_rtol, 1e-10 * scale, f"{what} {key} mc error")


@pytest.mark.parametrize("name", ["bermudan_swaption", "cfg4_bermudan_40"])
def test_bermudan_swaption_matches_reference_golden(name):
    """(cfg4_bermudan_40: BASELINE configs[3] at its exact shape - 40 quarterly exercise dates on an 11y swap,
    41 exposure dates - at 8192 paths.)"""
    gold = helpers.load_golden(name)
    res, sc = helpers.run_cuda(name, draws="torch")
    assert res.get_netting_set_names() == gold["sets"]
    assert res.get_metric_names() == gold["metrics"]
    assert [float(t) for t in sc.simulation_timeline] == gold["simulation_timeline"]
    flat = helpers.flatten_results(res)
    ref = {k: (np.array(v), np.array(gold["errors"][k])) for k, v in gold["values"].items()}
    # (the Monte Carlo error of PFE is a difference quotient of adjacent order statistics, pfe_metric.py:27-44: it
    # amplifies the 1e-10 agreement of the values by the ratio value / spacing)
    _compare(flat, ref, 1e-9, name, err_rtol=1e-6)


@pytest.mark.parametrize("kwargs", [dict(), dict(n_main=5000, n_pre=3000), dict(num_steps=2), dict(case="cfg4_bermudan_40")])
def test_bermudan_swaption_philox_matches_oracle(kwargs):
    kwargs = dict(kwargs)
    name = kwargs.pop("case", "bermudan_swaption")
    gold = helpers.load_golden(name)
    res, sc = helpers.run_cuda(name, draws="philox", **kwargs)
    out, _ = helpers.run_oracle(name, draws="philox", **kwargs)
    _compare(helpers.flatten_results(res), helpers.oracle_flat(out, gold["sets"], gold["metrics"]), 1e-8,
             name + " philox", err_rtol=1e-6)
    # regression coefficients exposed like the reference's controller.regression_coeffs (raw monomial basis)
    got = sc.regression_coeffs[0].numpy()          # [T_e, S, 3]
    want = np.stack([np.asarray(c) for c in out["expo_coeffs"][0]])
    fit_scale = np.abs(want).max(axis=(1, 2), keepdims=True) + 1e-30
    assert np.all(np.abs(got - want) <= 1e-6 * fit_scale), "exposure regression coefficients"


def test_bermudan_swaption_pv_greeks_match_oracle():
    """PV sensitivities of a Bermudan swaption (differentiate=True): the exercise policy is a hard indicator
    (bermudan_option.py:121, zero derivative), so the Greeks are the pathwise tangents of the exercised cashflow."""
    from oracle import risk
    ns = cases.Namespace()
    model, sets, metrics, tl = cases.bermudan_swaption(ns, n_ex=6)
    n = 3000
    sc = ns.SimulationController(sets, model, ns.RiskMetrics([ns.PVMetric()]), n, n, 1, ns.SimulationScheme.EULER, True)
    res = sc.run_simulation()
    out = risk.run(model, sets, [ns.PVMetric()], None, n, n, 1, "EULER", differentiate=True)
    helpers.assert_close(res.get_results("bermudan", "pv"), [out["results"][0][0][0][0]], 1e-8, 1e-12, "pv")
    got = np.array([0.0 if g is None else float(g) for g in res.get_derivatives("bermudan", "pv")[0]])
    want = np.asarray(out["grads"][0][0][0])
    helpers.assert_close(got, want, 1e-7, 1e-8 * max(1.0, float(np.abs(want).max())), "pv greeks")
    assert np.all(np.isfinite(got)) and np.any(got != 0.0)


def test_bermudan_swaption_exposure_greeks_match_oracle():
    """EPE / ENE / PV sensitivities of a Bermudan swaption (differentiate=True): the alive-state exposure proxies
    carry d(coefficients)/d(parameters) from the tangent Longstaff-Schwartz pass (irc_lsm_forward_tan_kernel,
    mcre_lsm_step_tangents, regression_tangents); exercise decisions stay hard.  Vs the oracle's forward-mode duals
    through the same regression chain."""
    from oracle import risk
    ns = cases.Namespace()
    model, sets, metrics, tl = cases.bermudan_swaption(ns, n_ex=6)
    metrics = [ns.PVMetric(), ns.EPEMetric(), ns.ENEMetric(), ns.PFEMetric(0.95)]
    n = 2500
    sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics, exposure_timeline=tl), n, n, 1, ns.SimulationScheme.EULER, True)
    res = sc.run_simulation()
    out = risk.run(model, sets, metrics, tl, n, n, 1, "EULER", differentiate=True)
    _compare(helpers.flatten_results(res), helpers.oracle_flat(out, ["bermudan"], res.get_metric_names()), 1e-8, "values",
             err_rtol=1e-6)
    for mi, m in enumerate(res.get_metric_names()):
        got = np.array([[0.0 if g is None else float(g) for g in row] for row in res.get_derivatives("bermudan", m)])
        want = np.array([np.zeros(got.shape[1]) if g is None else np.asarray(g) for g in out["grads"][0][mi]])
        scale = max(1.0, float(np.abs(want).max()))
        helpers.assert_close(got, want, 1e-6, 1e-7 * scale, f"bermudan {m} derivatives")
    assert np.any(np.array([[0.0 if g is None else float(g) for g in row] for row in res.get_derivatives("bermudan", "epe")]) != 0.0)


def test_bermudan_pv_only_and_mixed_book():
    """PV-only run (regression dates = exercise dates only) and a netting set mixing a swap
    with a Bermudan swaption, collateralised, all exposure metrics."""
    from oracle import risk
    ns = cases.Namespace()
    model, sets, metrics, tl = cases.bermudan_swaption(ns, n_ex=6)
    n = 4096
    sc = ns.SimulationController(sets, model, ns.RiskMetrics([ns.PVMetric()]), n, n, 1, ns.SimulationScheme.EULER)
    res = sc.run_simulation()
    out = risk.run(model, sets, [ns.PVMetric()], None, n, n, 1, "EULER")
    helpers.assert_close(res.get_results("bermudan", "pv"), [out["results"][0][0][0][0]], 1e-8, 1e-12, "pv only")
    helpers.assert_close(res.get_mc_error("bermudan", "pv"), [out["results"][0][0][0][1]], 1e-6, 1e-12, "pv only err")

    model, sets, _, tl = cases.bermudan_swaption(ns, n_ex=6)
    swap = ns.InterestRateSwap(0.0, 2.0, 1.0, 0.035, 0.25, 0.25, ns.IRSType.RECEIVER)
    mixed = [ns.NettingSet(name="mixed", products=[sets[0].products[0], swap], margin_period_of_risk=0.25, threshold=0.002),
             ns.NettingSet(name="swap_only", products=[ns.InterestRateSwap(0.0, 1.5, 1.0, 0.03, 0.25, 0.25, ns.IRSType.PAYER)])]
    metrics = [ns.PVMetric(), ns.CEMetric(), ns.EPEMetric(), ns.ENEMetric(), ns.EEPEMetric(), ns.PFEMetric(0.9)]
    sc = ns.SimulationController(mixed, model, ns.RiskMetrics(metrics, exposure_timeline=tl), n, n, 1, ns.SimulationScheme.EULER)
    res = sc.run_simulation()
    out = risk.run(model, mixed, metrics, tl, n, n, 1, "EULER")
    _compare(helpers.flatten_results(res), helpers.oracle_flat(out, res.get_netting_set_names(), res.get_metric_names()),
             1e-8, "mixed book", err_rtol=1e-6)


def _american(ns, n_ex=1000):
    model = ns.BlackScholesModel(0.0, 100, 0.05, 0.5)
    prod = ns.AmericanOption(ns.Equity("id"), 3.0, n_ex, 100.0, ns.OptionType.CALL)
    return model, [ns.NettingSet(name=prod.get_name(), products=[prod])], [ns.PVMetric()]


def test_american_option_reference_known_answer():
    """The reference's own known-answer test (tests/pytests/test_american_option.py:17-61): American call
    on Black-Scholes, 1000 exercise dates, 100k pre-simulation / 10k main paths, ANALYTICAL scheme,
    PV = 34.323036543142706 within the reference's threshold 1e-8 - with the reference's torch.randn
    stream injected, through the LSM kernels and the equity kernel's exercise events."""
    from oracle import engine
    ns = cases.Namespace()
    model, sets, metrics = _american(ns)
    sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics), 10000, 100000, 1, ns.SimulationScheme.ANALYTICAL)
    n_sub, dim = helpers.n_substeps(model, sets, None, metrics, 1)
    assert (n_sub, dim) == (999, 1)
    pre = engine.torch_reference_draws(42, 100000, n_sub, dim)
    main = engine.torch_reference_draws(43, 10000, n_sub, dim)
    sc.inject_normals(pre=pre.z, main=main.z)
    res = sc.run_simulation()
    pv = float(res.get_results(sets[0].get_name(), "pv")[0])
    assert abs(pv - 34.323036543142706) < 1e-8, pv


@pytest.mark.parametrize("which", ["bs_american", "heston_bermudan", "bs4_bermudan_greeks"])
def test_equity_exercise_philox_matches_oracle(which):
    from oracle import risk
    ns = cases.Namespace()
    S = ns.SimulationScheme
    if which == "bs_american":
        model, sets, metrics = _american(ns, n_ex=25)
        scheme, steps, diff = S.EULER, 2, False
    elif which == "heston_bermudan":
        model = ns.HestonModel(0.0, 100.0, 0.03, 0.4, -0.7, 2.0, 0.04, 0.04)
        put = ns.BermudanOption(ns.Equity(), [0.25, 0.5, 0.75, 1.0], 105.0, ns.OptionType.PUT)
        sets, metrics = [ns.NettingSet(name="put", products=[put, ns.EuropeanOption(ns.Equity(), 1.0, 100.0, ns.OptionType.CALL)])], [ns.PVMetric()]
        scheme, steps, diff = S.QE, 3, False
    else:
        ids = ["a", "b", "c", "d"]
        models = [ns.BlackScholesModel(0.0, 100.0 + 5 * i, 0.02, 0.3 + 0.05 * i, asset_id=a) for i, a in enumerate(ids)]
        model = ns.ModelConfig(models, inter_asset_correlation_matrix=np.array([[0.4] for _ in range(6)]))
        put = ns.BermudanOption(ns.Equity("c"), [0.5, 1.0, 1.5], 110.0, ns.OptionType.PUT, asset_id="c")
        sets, metrics = [ns.NettingSet(name="put_c", products=[put])], [ns.PVMetric()]
        scheme, steps, diff = S.ANALYTICAL, 2, True
    n = 5000
    sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics), n, n, steps, scheme, diff)
    res = sc.run_simulation()
    out = risk.run(model, sets, metrics, None, n, n, steps, scheme.name, differentiate=diff)
    name = sets[0].get_name()
    helpers.assert_close(res.get_results(name, "pv"), [out["results"][0][0][0][0]], 1e-8, 1e-10, which + " pv")
    helpers.assert_close(res.get_mc_error(name, "pv"), [out["results"][0][0][0][1]], 1e-6, 1e-10, which + " err")
    if diff:
        want = out["grads"][0][0][0]
        got = np.array([0.0 if g is None else float(g) for g in res.get_derivatives(name, "pv")[0]])
        helpers.assert_close(got, want, 1e-7, 1e-7 * max(1.0, float(np.max(np.abs(want)))), which + " greeks")


def test_hull_white_bermudan_swaption_extension_matches_oracle():
    """BASELINE config 4 names Hull-White: Vasicek dynamics with a piecewise-constant mean level theta(t)
    (build-defined extension - the reference's hull_white.py is dead code, SURVEY 8c: parity unpinned beyond
    the constant-theta case, which is the Vasicek golden above).  CUDA vs oracle on the same Philox stream."""
    from oracle import risk
    ns = cases.Namespace()
    model = ns.HullWhiteModel(0., 0.03, 0.05, 0.05, 0.02, mean_times=[0.5, 1.25], mean_levels=[0.02, 0.06])
    swap = ns.InterestRateSwap(0.0, 3.0, 1.0, 0.03, 0.25, 0.25, ns.IRSType.PAYER)
    opt = ns.BermudanOption(swap, [0.25 * (i + 1) for i in range(8)], 0.0, ns.OptionType.CALL)
    sets = [ns.NettingSet(name="bermudan", products=[opt])]
    metrics = [ns.EPEMetric(), ns.PFEMetric(0.95), ns.PVMetric()]
    tl = np.array([0.25 * i for i in range(9)])
    n = 4096
    sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics, exposure_timeline=tl), n, n, 2, ns.SimulationScheme.EULER)
    res = sc.run_simulation()
    out = risk.run(model, sets, metrics, tl, n, n, 2, "EULER")
    _compare(helpers.flatten_results(res), helpers.oracle_flat(out, res.get_netting_set_names(), res.get_metric_names()),
             1e-8, "hull-white bermudan", err_rtol=1e-6)


FLEXI_CASES = ["flexicall_pv", "flexicall_exposure", "flexicall_4_rights", "flexicall_6_rights", "mixed_book_pv", "mixed_book_exposure"]


@pytest.mark.parametrize("name", FLEXI_CASES)
def test_flexicall_and_mixed_book_match_reference_golden(name):
    """FlexiCall (src/products/flexicall.py: up to 6 exercise rights, state = rights left) alone and in the
    reference's mixed equity book (tests/pytests/test_netting_sets.py:375-528: Europeans, Americans, FlexiCall,
    barrier in one netting set), PV and EPE / PFE through state-dependent regression proxies.  The reference's
    own draws injected, outputs of the unmodified reference (tests/golden)."""
    gold = helpers.load_golden(name)
    res, sc = helpers.run_cuda(name, draws="torch")
    assert res.get_netting_set_names() == gold["sets"] and res.get_metric_names() == gold["metrics"]
    assert [float(t) for t in sc.simulation_timeline] == gold["simulation_timeline"]
    flat = helpers.flatten_results(res)
    ref = {k: (np.array(v), np.array(gold["errors"][k])) for k, v in gold["values"].items()}
    _compare(flat, ref, 1e-9, name)


@pytest.mark.parametrize("name", FLEXI_CASES)
def test_flexicall_and_mixed_book_philox_match_oracle(name):
    gold = helpers.load_golden(name)
    res, sc = helpers.run_cuda(name, draws="philox")
    out, _ = helpers.run_oracle(name, draws="philox")
    _compare(helpers.flatten_results(res), helpers.oracle_flat(out, gold["sets"], gold["metrics"]), 1e-8,
             name + " philox", err_rtol=1e-6)


@pytest.mark.parametrize("differentiate", [False, True])
def test_large_mixed_book_is_split_over_launches(differentiate):
    """A netting set with more path-dependent / exercise products than one launch tracks (4, or 2 with tangents):
    the book is evaluated in several launches that replay the same Philox streams and accumulate per-path
    cashflows (mcre_eq_set_pv_accumulator + mcre_sum_stats).  PV, MC error and pathwise Greeks vs the oracle,
    which evaluates the whole book at once like the reference (tests/pv_tests/pv_performance_large_netting_set.py)."""
    from oracle import risk
    ns = cases.Namespace()
    ids = ["asset_1", "asset_2"]
    model = ns.BlackScholesMulti(calibration_date=0.0, rate=0.03, asset_ids=ids, spots=[100.0, 105.0],
                                 volatilities=[0.20, 0.24], correlation_matrix=np.array([[1.0, 0.35], [0.35, 1.0]]))
    prods = []
    for k in range(3):
        a = ids[k % 2]
        prods.append(ns.AmericanOption(underlying=ns.Equity(a), maturity=1.0 + 0.25 * k, num_exercise_dates=5 + k,
                                       strike=95.0 + 5.0 * k, option_type=ns.OptionType.PUT if k % 2 else ns.OptionType.CALL,
                                       asset_id=a))
        prods.append(ns.BarrierOption(startdate=0.0, maturity=1.0 + 0.25 * k, strike=100.0, num_observation_timepoints=6,
                                      option_type=ns.OptionType.CALL, barrier1=130.0 + 5 * k,
                                      barrier_option_type1=ns.BarrierOptionType.UPANDOUT, asset_id=a))
        prods.append(ns.AsianOption(0.0, 1.0, 100.0 + 2 * k, 4 + k, ns.OptionType.CALL, asset_id=a))
        prods.append(ns.EuropeanOption(ns.Equity(a), 0.5 + 0.5 * k, 100.0, ns.OptionType.PUT, asset_id=a))
    prods.append(ns.FlexiCall(underlyings=[ns.EuropeanOption(ns.Equity("asset_1"), 0.5 * (j + 1), 95.0 + 5 * j, ns.OptionType.CALL,
                                                             asset_id="asset_1") for j in range(3)],
                              num_exercise_rights=2, asset_id="asset_1"))
    sets = [ns.NettingSet(name="big", products=prods),
            ns.NettingSet(name="small", products=[ns.EuropeanOption(ns.Equity("asset_2"), 1.0, 105.0, ns.OptionType.CALL, asset_id="asset_2")])]
    n = 3000
    sc = ns.SimulationController(sets, model, ns.RiskMetrics([ns.PVMetric()]), n, n, 1, ns.SimulationScheme.ANALYTICAL, differentiate)
    res = sc.run_simulation()
    out = risk.run(model, sets, [ns.PVMetric()], None, n, n, 1, "ANALYTICAL", differentiate=differentiate)
    for si, name in enumerate(["big", "small"]):
        want_v, want_e = out["results"][si][0][0]
        helpers.assert_close(res.get_results(name, "pv"), [want_v], 1e-8, 1e-10, f"{name} pv")
        helpers.assert_close(res.get_mc_error(name, "pv"), [want_e], 1e-6, 1e-10, f"{name} pv error")
        if differentiate:
            got = np.array([0.0 if g is None else float(g) for g in res.get_derivatives(name, "pv")[0]])
            want = np.asarray(out["grads"][si][0][0])
            helpers.assert_close(got, want, 1e-6, 1e-7 * max(1.0, float(np.abs(want).max())), f"{name} pv greeks")


def test_large_mixed_book_exposure_profiles_are_split_over_launches():
    """More path-dependent / exercise products than the 64 trackers of one value-only launch, with exposure metrics:
    every launch adds its products' netted exposures per (exposure date, path) to one accumulator
    (mcre_eq_set_exposure_accumulator), threshold / MPoR collateral are applied afterwards
    (mcre_eq_unsecured_exposures), EPE / ENE from mcre_sum_stats, PFE from the radix select.  Vs the oracle, which
    nets the whole book at once like the reference (tests/exposure_tests/ee_performance_large_netting_set.py)."""
    from oracle import risk
    ns = cases.Namespace()
    ids = ["asset_1", "asset_2"]
    model = ns.BlackScholesMulti(calibration_date=0.0, rate=0.03, asset_ids=ids, spots=[100.0, 105.0],
                                 volatilities=[0.20, 0.24], correlation_matrix=np.array([[1.0, 0.35], [0.35, 1.0]]))
    prods = []
    for k in range(66):
        a = ids[k % 2]
        if k % 3 == 0:
            prods.append(ns.AsianOption(0.0, 0.5 + 0.25 * (k % 3), 95.0 + (k % 7), 3 + k % 3, ns.OptionType.CALL if k % 2 else ns.OptionType.PUT,
                                        asset_id=a))
        else:
            prods.append(ns.BarrierOption(startdate=0.0, maturity=0.5 + 0.25 * (k % 4), strike=100.0, num_observation_timepoints=3 + k % 4,
                                          option_type=ns.OptionType.CALL, barrier1=125.0 + k % 11,
                                          barrier_option_type1=ns.BarrierOptionType.UPANDOUT, asset_id=a))
    prods.append(ns.AmericanOption(underlying=ns.Equity("asset_1"), maturity=1.0, num_exercise_dates=5, strike=100.0,
                                   option_type=ns.OptionType.PUT, asset_id="asset_1"))
    prods.append(ns.EuropeanOption(ns.Equity("asset_2"), 1.0, 100.0, ns.OptionType.CALL, asset_id="asset_2"))
    sets = [ns.NettingSet(name="big", products=prods, margin_period_of_risk=0.25, threshold=2.0)]
    metrics = [ns.PVMetric(), ns.EPEMetric(), ns.ENEMetric(), ns.PFEMetric(0.9)]
    tl = np.linspace(0.0, 1.25, 6)
    n = 2048
    sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics, exposure_timeline=tl), n, n, 1, ns.SimulationScheme.ANALYTICAL)
    res = sc.run_simulation()
    out = risk.run(model, sets, metrics, tl, n, n, 1, "ANALYTICAL")
    _compare(helpers.flatten_results(res), helpers.oracle_flat(out, ["big"], res.get_metric_names()), 1e-7, "big book exposures",
             err_rtol=1e-5)


def test_large_mixed_book_cva_is_split_over_launches():
    """CVA of a book with more tracked products than one launch holds, facing a counterparty with a CIR++ intensity
    (the shape of tests/exposure_tests/cva_perfprmance_large_netting_set.py): the first launch carries the credit
    factor and spills the default weights (mcre_eq_set_cva_weight_spill), every launch adds its exposures,
    mcre_eq_cva_paths + mcre_sum_stats finish.  Vs the oracle, which nets the whole book at once."""
    from oracle import risk
    ns = cases.Namespace()
    model, sets, metrics, tl = cases.big_cva_book(ns)
    n = 2048
    sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics, exposure_timeline=tl), n, n, 1, ns.SimulationScheme.EULER)
    res = sc.run_simulation()
    out = risk.run(model, sets, metrics, tl, n, n, 1, "EULER")
    _compare(helpers.flatten_results(res), helpers.oracle_flat(out, ["big"], res.get_metric_names()), 1e-7, "big book cva",
             err_rtol=1e-5)


@pytest.mark.parametrize("name", ["bermudan_swaption", "cfg4_bermudan_40"])
def test_device_side_backward_induction_equals_host_driven_one(name, monkeypatch):
    """The Longstaff-Schwartz backward induction of a rate Bermudan runs as a stream of kernels (fused step with the
    continuation coefficients read from device memory, 3x3 solve on the device: mcre_lsm_step_dev /
    mcre_lsm_solve_dev) instead of one device-to-host round trip per exercise date.  Same schedule, same arithmetic:
    every path must take the same exercise decisions as with the host-driven induction (numpy solve), so all results
    agree far inside the parity tolerance."""
    monkeypatch.setenv("MCRE_DEVICE_SOLVE", "1")
    res_d, sc_d = helpers.run_cuda(name, draws="philox")
    monkeypatch.setenv("MCRE_DEVICE_SOLVE", "0")
    res_h, sc_h = helpers.run_cuda(name, draws="philox")
    fd, fh = helpers.flatten_results(res_d), helpers.flatten_results(res_h)
    _compare(fd, fh, 1e-11, name + " device vs host induction", err_rtol=1e-6)
    for cd, ch in zip(sc_d.regression_coeffs, sc_h.regression_coeffs):
        cd, ch = cd.numpy(), ch.numpy()
        fit = np.abs(ch).max(axis=(1, 2), keepdims=True) + 1e-30
        assert np.all(np.abs(cd - ch) <= 1e-7 * fit)
