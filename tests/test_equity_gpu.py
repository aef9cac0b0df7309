"""Parity of the fused equity kernel (csrc/equity.cu) through the C ABI.

  1. reference goldens: the reference's own torch.randn / torch.rand stream injected, compared
     with the outputs of the unmodified reference (tests/golden/*.json): PV, MC error 1e-10,
     pathwise Greeks (kernel tangents vs the reference's autograd) 1e-8 relative
  2. native Philox vs the oracle run on the same Philox stream (1e-8) and vs the reference
     golden within 4 combined MC standard errors
"""
import numpy as np
import pytest

import cases
import parity_helpers as helpers

pytestmark = pytest.mark.gpu

EQ_CASES = ["bs_european", "bs_european_euler", "heston_european", "heston_european_greeks",
            "heston_path_dependent", "bs_basket", "bs_basket_euler", "bs_exposure_greeks", "bs_exposure_greeks_euler",
            # models the reference only runs standalone: Schwartz two-factor (both schemes), Heston under EULER
            "schwartz_analytical", "schwartz_euler", "heston_euler"]
RTOL = 1e-10


def _check_values(flat, ref, rtol, what, err_rtol=1e-7):
    for key, (va, ea) in flat.items():
        vb, eb = ref[key]
        scale = max(1.0, float(np.nanmax(np.abs(vb))))
        helpers.assert_close(va, vb, rtol, rtol * scale, f"{what} {key} value")
        helpers.assert_close(ea, eb, err_rtol, 1e-11 * scale, f"{what} {key} mc error")


@pytest.mark.parametrize("name", EQ_CASES)
def test_injected_draws_match_reference_golden(name):
    gold = helpers.load_golden(name)
    res, sc = helpers.run_cuda(name, draws="torch")
    assert res.get_netting_set_names() == gold["sets"]
    assert res.get_metric_names() == gold["metrics"]
    assert res.get_model_param_names() == gold["params"]
    flat = helpers.flatten_results(res)
    ref = {k: (np.array(v), np.array(gold["errors"][k])) for k, v in gold["values"].items()}
    _check_values(flat, ref, RTOL, name)
    if gold["run"]["differentiate"]:
        for s in gold["sets"]:
            for m in gold["metrics"]:
                want = np.array([[0.0 if g is None else g for g in row] for row in gold["derivatives"][f"{s}|{m}"]])
                got = np.array([[0.0 if g is None else float(g) for g in row] for row in res.get_derivatives(s, m)])
                helpers.assert_close(got, want, 1e-8, 1e-8 * max(1.0, float(np.max(np.abs(want)))),
                                     f"{name} {s}|{m} derivatives")
                # parameters outside the reference's autograd graph come back as None there, and here
                assert [[g is None for g in row] for row in res.get_derivatives(s, m)] == \
                       [[g is None for g in row] for row in gold["derivatives"][f"{s}|{m}"]], f"{name} {s}|{m} None pattern"


@pytest.mark.parametrize("name", EQ_CASES)
def test_philox_matches_oracle_and_reference_statistically(name):
    gold = helpers.load_golden(name)
    res, sc = helpers.run_cuda(name, draws="philox")
    out, _ = helpers.run_oracle(name, draws="philox")
    flat = helpers.flatten_results(res)
    _check_values(flat, helpers.oracle_flat(out, gold["sets"], gold["metrics"]), 1e-8, name + " philox", err_rtol=1e-6)
    for key, (v, e) in flat.items():
        rv, re_ = np.array(gold["values"][key]), np.array(gold["errors"][key])
        se = np.sqrt(e ** 2 + re_ ** 2)
        assert np.all(np.abs(v - rv) <= 4.0 * se + 1e-12), f"{name} {key}: {v} vs {rv} (se {se})"
    if gold["run"]["differentiate"]:
        for si, s in enumerate(gold["sets"]):
            for mi, m in enumerate(gold["metrics"]):
                for ev, want in enumerate(out["grads"][si][mi]):      # every evaluation (metric date) of the metric
                    got = np.array([0.0 if g is None else float(g) for g in res.get_derivatives(s, m)[ev]])
                    helpers.assert_close(got, want, 1e-7, 1e-7 * max(1.0, float(np.max(np.abs(want)))),
                                         f"{name} {s}|{m}[{ev}] philox derivatives")


def _run(ns, model, sets, metrics, n, steps, scheme, differentiate=False):
    sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics), n, 0, steps, scheme, differentiate)
    return sc.run_simulation(), sc


def test_other_payoffs_and_models_match_oracle():
    """Payoffs / models without a reference golden of their own: binary, double barrier,
    geometric Asian, control-variate basket on BlackScholesMulti, Schwartz two-factor
    (ANALYTICAL and EULER), Heston Euler - CUDA vs oracle on the same Philox stream."""
    from oracle import risk
    ns = cases.Namespace()
    S = ns.SimulationScheme
    bsm = ns.BlackScholesMulti(0.0, 0.03, ["a", "b", "c"], [100.0, 90.0, 110.0], [0.2, 0.3, 0.25],
                               np.array([[1.0, 0.5, 0.2], [0.5, 1.0, 0.3], [0.2, 0.3, 1.0]]))
    sch = ns.SchwartzTwoFactorModel(0.0, [0.0, 0.5, 1.0, 2.0], [50.0, 52.0, 51.0, 55.0], 0.03, 1.2, 0.4, 0.02, 0.15, 0.3)
    hes = lambda: ns.HestonModel(0.0, 100.0, 0.03, 0.4, -0.7, 2.0, 0.04, 0.04)
    bs = lambda: ns.BlackScholesModel(0.0, 100.0, 0.05, 0.2)

    def book_single():
        return [
            ns.NettingSet("binary", [ns.BinaryOption(1.0, 100.0, 10.0, ns.OptionType.CALL)]),
            ns.NettingSet("dbl", [ns.BarrierOption(0.0, 1.0, 95.0, 7, ns.OptionType.PUT, 130.0, ns.BarrierOptionType.UPANDOUT,
                                                  80.0, ns.BarrierOptionType.DOWNANDIN)]),
            ns.NettingSet("geo_asian+put", [ns.AsianOption(0.25, 1.0, 100.0, 4, ns.OptionType.CALL, ns.AsianAveragingType.GEOMETRIC),
                                            ns.EuropeanOption(ns.Equity(), 0.5, 100.0, ns.OptionType.PUT)]),
        ]

    def book_multi():
        w = [0.3, 0.3, 0.4]
        return [
            ns.NettingSet("cv", [ns.BasketOption(1.0, ["a", "b", "c"], w, 100.0, ns.OptionType.CALL,
                                                 ns.BasketOptionType.ARITHMETIC, True)]),
            ns.NettingSet("geo_put", [ns.BasketOption(1.0, ["a", "b", "c"], w, 100.0, ns.OptionType.PUT,
                                                      ns.BasketOptionType.GEOMETRIC)]),
            ns.NettingSet("single", [ns.EuropeanOption(ns.Equity("b"), 0.5, 90.0, ns.OptionType.CALL),
                                     ns.BinaryOption(1.0, 100.0, 5.0, ns.OptionType.PUT, asset_id="c")]),
        ]

    runs = [
        ("bs_analytical", bs(), book_single(), S.ANALYTICAL, 3, True),
        ("bs_euler", bs(), book_single(), S.EULER, 3, False),
        ("heston_euler", hes(), book_single(), S.EULER, 4, True),
        ("heston_qe", hes(), book_single(), S.QE, 4, False),
        ("schwartz_analytical", sch, book_single(), S.ANALYTICAL, 2, True),
        ("schwartz_euler", sch, book_single(), S.EULER, 2, True),
        ("bsm_analytical", bsm, book_multi(), S.ANALYTICAL, 2, False),
        ("bsm_euler", bsm, book_multi(), S.EULER, 2, True),
    ]
    n = 3000   # ragged: not a multiple of the chunk size
    for name, model, sets, scheme, steps, diff in runs:
        metrics = [ns.PVMetric()]
        res, sc = _run(ns, model, sets, metrics, n, steps, scheme, diff)
        out = risk.run(model, sets, metrics, None, n, 0, steps, scheme.name, differentiate=diff)
        for si, s in enumerate(res.get_netting_set_names()):
            want_v, want_e = out["results"][si][0][0]
            helpers.assert_close(res.get_results(s, "pv"), [want_v], 1e-8, 1e-9, f"{name} {s} pv")
            helpers.assert_close(res.get_mc_error(s, "pv"), [want_e], 1e-6, 1e-10, f"{name} {s} err")
            if diff:
                want = out["grads"][si][0][0]
                got = np.array([0.0 if g is None else float(g) for g in res.get_derivatives(s, "pv")[0]])
                helpers.assert_close(got, want, 1e-7, 1e-7 * max(1.0, float(np.max(np.abs(want)))), f"{name} {s} greeks")


def test_analytic_pv_shortcut_and_edge_cases():
    ns = cases.Namespace()
    model = ns.BlackScholesModel(0, 120.0, 0.05, 0.2)
    opt = ns.EuropeanOption(ns.Equity(), 2.0, 100.0, ns.OptionType.CALL)
    # ANALYTICAL evaluation type: no Monte Carlo at all (controller.py:506-533)
    from metrics.metric import Metric
    res, _ = _run(ns, model, [ns.NettingSet("call", [opt])], [ns.PVMetric(Metric.EvaluationType.ANALYTICAL)], 16, 1,
                  ns.SimulationScheme.ANALYTICAL)
    assert abs(res.get_results("call", "pv")[0] - 31.96484725190807) < 1e-9 and res.get_mc_error("call", "pv")[0] == 0.0
    # one path: NaN error like the reference; 1 sub-step
    res, _ = _run(ns, model, [ns.NettingSet("call", [ns.EuropeanOption(ns.Equity(), 2.0, 100.0, ns.OptionType.CALL)])],
                  [ns.PVMetric()], 1, 1, ns.SimulationScheme.ANALYTICAL)
    assert np.isfinite(res.get_results("call", "pv")[0]) and np.isnan(res.get_mc_error("call", "pv")[0])


def test_heston_basket_extension_matches_oracle():
    """BASELINE config 5 at reduced size: 5 correlated Heston QE assets, barrier + Asian on the
    basket, first-order pathwise Greeks w.r.t. all 35 model parameters.  Build-defined extension
    (parity unpinned by the reference beyond the single-asset case): checked against the
    oracle's composition of the pinned single-asset step on the same Philox stream."""
    from oracle import risk
    ns = cases.Namespace()
    for diff, steps, n in ((False, 3, 3000), (True, 2, 2048)):
        model, sets, metrics, _ = cases.heston_basket5(ns)
        res, sc = _run(ns, model, sets, metrics, n, steps, ns.SimulationScheme.QE, diff)
        out = risk.run(model, sets, metrics, None, n, 0, steps, "QE", differentiate=diff)
        for si, s in enumerate(res.get_netting_set_names()):
            want_v, want_e = out["results"][si][0][0]
            helpers.assert_close(res.get_results(s, "pv"), [want_v], 1e-8, 1e-9, f"basket5 {s} pv")
            helpers.assert_close(res.get_mc_error(s, "pv"), [want_e], 1e-6, 1e-10, f"basket5 {s} err")
            if diff:
                want = out["grads"][si][0][0]
                got = np.array([0.0 if g is None else float(g) for g in res.get_derivatives(s, "pv")[0]])
                assert got.shape == (35,)
                helpers.assert_close(got, want, 1e-7, 1e-7 * max(1.0, float(np.max(np.abs(want)))), f"basket5 {s} greeks")


def test_bs_european_analytic_exposure_profiles_match_oracle():
    """EPE / PFE profiles of European options under Black-Scholes use the analytic exposure
    BS(S_t, T - t) / N(t) and need no regression (european_option.py:123-145, controller.py:204-229;
    the reference's tests/pytests/test_netting_sets.py:101-165): CUDA vs oracle on the same Philox stream,
    with a threshold and an MPoR-collateralised netting set, and the martingale property the
    reference itself tests (discounted EPE = initial PV before maturity)."""
    from oracle import risk
    ns = cases.Namespace()
    bsm = ns.BlackScholesMulti(0.0, 0.03, ["eq1", "eq2"], [100.0, 110.0], [0.2, 0.25], np.array([[1.0, 0.2], [0.2, 1.0]]))
    call = lambda a, T, K: ns.EuropeanOption(ns.Equity(a), T, K, ns.OptionType.CALL, asset_id=a)
    put = lambda a, T, K: ns.EuropeanOption(ns.Equity(a), T, K, ns.OptionType.PUT, asset_id=a)
    sets = [ns.NettingSet(name="call", products=[call("eq1", 1.0, 100.0)]),
            ns.NettingSet(name="book", products=[call("eq2", 0.75, 105.0), put("eq1", 1.0, 95.0)], threshold=2.0),
            ns.NettingSet(name="collateralised", products=[call("eq1", 1.0, 100.0), put("eq2", 0.5, 110.0)], margin_period_of_risk=0.25)]
    metrics = [ns.EPEMetric(), ns.PFEMetric(0.95), ns.PVMetric()]
    tl = np.array([0.0, 0.25, 0.5, 0.75, 1.0])
    n = 6000
    for scheme, steps in ((ns.SimulationScheme.ANALYTICAL, 1), (ns.SimulationScheme.EULER, 3)):
        sc = ns.SimulationController(sets, bsm, ns.RiskMetrics(metrics, exposure_timeline=tl), n, n, steps, scheme)
        assert sc.requires_regression is False and sc._product_requires_regression(sets[0].products[0]) is False
        res = sc.run_simulation()
        out = risk.run(bsm, sets, metrics, tl, n, n, steps, scheme.name)
        got, want = helpers.flatten_results(res), helpers.oracle_flat(out, res.get_netting_set_names(), res.get_metric_names())
        for key in got:
            helpers.assert_close(got[key][0], want[key][0], 1e-8, 1e-9, f"{scheme.name} {key} value")
            helpers.assert_close(got[key][1], want[key][1], 1e-6, 1e-9, f"{scheme.name} {key} mc error")
    # martingale: the discounted expected exposure of the single call equals its PV at every date before maturity
    big = ns.SimulationController(sets[:1], bsm, ns.RiskMetrics(metrics, exposure_timeline=tl), 1 << 20, 0, 1,
                                  ns.SimulationScheme.ANALYTICAL).run_simulation()
    pv = float(sets[0].products[0].compute_pv_analytically(bsm).reshape(-1)[0])
    epe, err = np.array(big.get_results("call", "epe")), np.array(big.get_mc_error("call", "epe"))
    assert abs(epe[0] - pv) < 1e-10 and err[0] == 0.0
    assert np.all(np.abs(epe[1:4] - pv) <= 4.0 * err[1:4]) and epe[4] == 0.0


@pytest.mark.parametrize("which", ["heston_qe", "bsm_full_metrics", "bs_euler_mixed"])
def test_regression_proxy_exposure_profiles_match_oracle(which):
    """Exposure profiles of equity products through the regression proxy (controller.py:294-471): the
    pre-simulation spills spots and float32 discounted cashflows (mcre_eq_presim), per-date quadratic fits in
    the spot of the product's first asset, then netting / threshold / MPoR collateral and CE / EPE / ENE /
    EEPE / PFE in the fused kernel.  CUDA vs oracle on the same Philox streams."""
    from oracle import risk
    ns = cases.Namespace()
    S = ns.SimulationScheme
    full = [ns.PVMetric(), ns.CEMetric(), ns.EPEMetric(), ns.ENEMetric(), ns.EEPEMetric(), ns.PFEMetric(0.9)]
    if which == "heston_qe":
        model = ns.HestonModel(0.0, 100.0, 0.03, 0.4, -0.7, 2.0, 0.04, 0.04)
        sets = [ns.NettingSet(name="call", products=[ns.EuropeanOption(ns.Equity(), 1.0, 100.0, ns.OptionType.CALL)]),
                ns.NettingSet(name="exotics", products=[ns.BarrierOption(0.0, 1.0, 100.0, 5, ns.OptionType.CALL, 140.0, ns.BarrierOptionType.UPANDOUT),
                                                        ns.AsianOption(0.0, 0.75, 100.0, 4, ns.OptionType.PUT)], threshold=1.0),
                ns.NettingSet(name="collateralised", products=[ns.BinaryOption(1.0, 100.0, 10.0, ns.OptionType.CALL),
                                                               ns.EuropeanOption(ns.Equity(), 0.5, 95.0, ns.OptionType.PUT)],
                              margin_period_of_risk=0.25)]
        metrics, scheme, steps = full, S.QE, 3
    elif which == "bsm_full_metrics":
        model = ns.BlackScholesMulti(0.0, 0.03, ["a", "b", "c"], [100.0, 90.0, 110.0], [0.2, 0.3, 0.25],
                                     np.array([[1.0, 0.5, 0.2], [0.5, 1.0, 0.3], [0.2, 0.3, 1.0]]))
        w = [0.3, 0.3, 0.4]
        sets = [ns.NettingSet(name="basket", products=[ns.BasketOption(1.0, ["a", "b", "c"], w, 100.0, ns.OptionType.CALL)]),
                ns.NettingSet(name="single", products=[ns.EuropeanOption(ns.Equity("b"), 0.75, 90.0, ns.OptionType.CALL, asset_id="b"),
                                                       ns.BinaryOption(1.0, 100.0, 5.0, ns.OptionType.PUT, asset_id="c")],
                              margin_period_of_risk=0.5, threshold=0.5)]
        metrics, scheme, steps = full, S.ANALYTICAL, 1     # ENE / CE / EEPE force the regression proxy even for Europeans
    else:
        model = ns.BlackScholesModel(0.0, 100.0, 0.05, 0.2)
        sets = [ns.NettingSet(name="mixed", products=[ns.EuropeanOption(ns.Equity(), 1.0, 100.0, ns.OptionType.CALL),
                                                      ns.AsianOption(0.25, 1.0, 100.0, 4, ns.OptionType.CALL, ns.AsianAveragingType.GEOMETRIC)])]
        metrics, scheme, steps = [ns.PVMetric(), ns.EPEMetric(), ns.PFEMetric(0.95)], S.EULER, 2   # analytic + regression in one set
    tl = np.array([0.0, 0.25, 0.5, 0.75, 1.0])
    n = 6000
    sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics, exposure_timeline=tl), n, n, steps, scheme)
    assert sc.requires_regression
    res = sc.run_simulation()
    out = risk.run(model, sets, metrics, tl, n, n, steps, scheme.name)
    got, want = helpers.flatten_results(res), helpers.oracle_flat(out, res.get_netting_set_names(), res.get_metric_names())
    for key in got:
        scale = max(1.0, float(np.nanmax(np.abs(want[key][0]))))
        helpers.assert_close(got[key][0], want[key][0], 1e-7, 1e-8 * scale, f"{which} {key} value")
        helpers.assert_close(got[key][1], want[key][1], 1e-5, 1e-8 * scale, f"{which} {key} mc error")
    # regression coefficients exposed like the reference's controller.regression_coeffs (raw basis)
    for k, p in enumerate(sc.products):
        if sc._product_requires_regression(p):
            gotc = sc.regression_coeffs[k][:, 0, :].numpy()
            wantc = np.stack([np.asarray(cc)[0] for cc in out["expo_coeffs"][k]])
            x0 = 100.0
            fit = lambda cfs: cfs[:, 0] + cfs[:, 1] * x0 + cfs[:, 2] * x0 * x0     # compare fitted values near the money
            helpers.assert_close(fit(gotc), fit(wantc), 1e-6, 1e-7, f"{which} product {k} fitted continuation")


def test_brownian_bridge_barrier_matches_reference_golden_and_oracle():
    """Barrier options with the Brownian-bridge correction between monitoring dates (barrier_option.py:138-222):
    up-and-out, down-and-in, double knock-out, up-and-in.  (1) RNG compatibility mode - the reference's torch.randn
    normals AND its numpy default_rng(12345) bridge uniforms injected - against the outputs of the unmodified
    reference; (2) native Philox (bridge uniforms from stream kind 2) against the oracle on the same streams."""
    name = "bs_bridge_barrier"
    gold = helpers.load_golden(name)
    ns, model, sets, metrics, tl, rkw = helpers.build(name)
    sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics), rkw["n_main"], rkw["n_pre"], rkw["num_steps"],
                                 getattr(ns.SimulationScheme, rkw["scheme"]), rkw["differentiate"])
    sc.rng_compat = "torch"
    res = sc.run_simulation()
    ref = {k: (np.array(v), np.array(gold["errors"][k])) for k, v in gold["values"].items()}
    _check_values(helpers.flatten_results(res), ref, RTOL, name)
    res, _ = helpers.run_cuda(name, draws="philox")
    out, _ = helpers.run_oracle(name, draws="philox")
    _check_values(helpers.flatten_results(res), helpers.oracle_flat(out, gold["sets"], gold["metrics"]), 1e-8, name + " philox",
                  err_rtol=1e-6)
    # the correction removes knock-out paths that crossed between the dates; the reference's knock-in variant
    # multiplies the discrete knock-in indicator by the bridge-hit indicator, so it can only stay or shrink
    plain_sets = cases.bs_bridge_barrier(ns)[1]
    for s in plain_sets:
        s.products[0].use_brownian_bridge = False
    plain = ns.SimulationController(plain_sets, model, ns.RiskMetrics(metrics), rkw["n_main"], 0, rkw["num_steps"],
                                    ns.SimulationScheme.EULER, False).run_simulation()
    assert float(res.get_results("up_out", "pv")[0]) < float(plain.get_results("up_out", "pv")[0])
    assert float(res.get_results("up_in", "pv")[0]) <= float(plain.get_results("up_in", "pv")[0]) + 1e-12


def test_brownian_bridge_barrier_greeks_match_reference_autograd_and_oracle():
    """The same four options with differentiate=True (tests/pv_tests/pv_barrier_option.py): the crossing probabilities and
    band-limited hit indicators carry tangents through both monitored spots and the volatility.  RNG compatibility mode
    against the reference's autograd; native Philox against the oracle's duals."""
    name = "bs_bridge_barrier_greeks"
    gold = helpers.load_golden(name)
    ns, model, sets, metrics, tl, rkw = helpers.build(name)
    sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics), rkw["n_main"], rkw["n_pre"], rkw["num_steps"],
                                 getattr(ns.SimulationScheme, rkw["scheme"]), rkw["differentiate"])
    sc.rng_compat = "torch"
    res = sc.run_simulation()
    ref = {k: (np.array(v), np.array(gold["errors"][k])) for k, v in gold["values"].items()}
    _check_values(helpers.flatten_results(res), ref, RTOL, name)
    for s_ in gold["sets"]:
        want = np.array(gold["derivatives"][f"{s_}|pv"][0])
        got = np.array([float(g) for g in res.get_derivatives(s_, "pv")[0]])
        helpers.assert_close(got, want, 1e-8, 1e-8 * float(np.max(np.abs(want))), f"{name} {s_} greeks")
    res, _ = helpers.run_cuda(name, draws="philox")
    out, _ = helpers.run_oracle(name, draws="philox")
    _check_values(helpers.flatten_results(res), helpers.oracle_flat(out, gold["sets"], gold["metrics"]), 1e-8, name + " philox",
                  err_rtol=1e-6)
    for si, s_ in enumerate(gold["sets"]):
        want = out["grads"][si][0][0]
        got = np.array([float(g) for g in res.get_derivatives(s_, "pv")[0]])
        helpers.assert_close(got, want, 1e-7, 1e-7 * float(np.max(np.abs(want))), f"{name} {s_} philox greeks")


PROXY_GREEK_CASES = ["bs_eepe_greeks", "bs_proxy_greeks_mixed"]


@pytest.mark.parametrize("name", PROXY_GREEK_CASES)
def test_regression_proxy_exposure_greeks_match_reference_and_oracle(name):
    """Sensitivities of EPE / ENE / CE / EEPE of equity books through the regression proxy: tangent pre-simulation
    spill (mcre_eq_presim_tangents), differentiated normal equations (mcre_lsm_step_tangents + regression_tangents),
    dual-valued quadratic in the fused kernel.  Reference golden with its own draws injected: values 1e-9, Greeks of
    exposure metrics 2e-5 (the reference's autograd runs through its float32 cashflow accumulators, SURVEY A-19),
    PV Greeks 1e-8.  Native Philox vs the oracle's forward-mode duals: 1e-6."""
    gold = helpers.load_golden(name)
    res, sc = helpers.run_cuda(name, draws="torch")
    assert res.get_netting_set_names() == gold["sets"] and res.get_metric_names() == gold["metrics"]
    flat = helpers.flatten_results(res)
    ref = {k: (np.array(v), np.array(gold["errors"][k])) for k, v in gold["values"].items()}
    _check_values(flat, ref, 1e-9, name)
    for s in gold["sets"]:
        for m in gold["metrics"]:
            want = np.array([[0.0 if g is None else g for g in row] for row in gold["derivatives"][f"{s}|{m}"]])
            got = np.array([[0.0 if g is None else float(g) for g in row] for row in res.get_derivatives(s, m)])
            rtol = 1e-8 if m.startswith("pv") else 2e-5
            helpers.assert_close(got, want, rtol, rtol * max(1.0, float(np.max(np.abs(want)))), f"{name} {s}|{m} derivatives")
    res, sc = helpers.run_cuda(name, draws="philox")
    out, _ = helpers.run_oracle(name, draws="philox")
    _check_values(helpers.flatten_results(res), helpers.oracle_flat(out, gold["sets"], gold["metrics"]), 1e-8,
                  name + " philox", err_rtol=1e-6)
    for si, s in enumerate(gold["sets"]):
        for mi, m in enumerate(gold["metrics"]):
            for ev, want in enumerate(out["grads"][si][mi]):
                got = np.array([0.0 if g is None else float(g) for g in res.get_derivatives(s, m)[ev]])
                helpers.assert_close(got, want, 1e-6, 1e-6 * max(1.0, float(np.max(np.abs(want)))),
                                     f"{name} {s}|{m}[{ev}] philox derivatives")


HYBRID_CVA_CASES = ["equity_cva", "equity_cva_single_det", "equity_cva_exercise"]


@pytest.mark.parametrize("name", HYBRID_CVA_CASES)
def test_equity_book_cva_matches_reference_and_oracle(name):
    """CVA of equity books (tests/exposure_tests/cva_perfprmance_large_netting_set.py, reduced): ModelConfig of a
    Black-Scholes market model and the counterparty's CIR++ intensity, stepped in the fused equity kernel on the last
    column of the joint draw (mcre_eq_set_credit); regression-proxy exposures, collateralised and thresholded sets,
    stochastic (correlated) and deterministic credit.  Reference golden with its own draws injected 1e-9; native
    Philox vs the oracle 1e-8."""
    gold = helpers.load_golden(name)
    res, sc = helpers.run_cuda(name, draws="torch")
    assert res.get_netting_set_names() == gold["sets"] and res.get_metric_names() == gold["metrics"]
    assert res.get_model_param_names() == gold["params"]
    flat = helpers.flatten_results(res)
    ref = {k: (np.array(v), np.array(gold["errors"][k])) for k, v in gold["values"].items()}
    _check_values(flat, ref, 1e-9, name)
    res, sc = helpers.run_cuda(name, draws="philox")
    out, _ = helpers.run_oracle(name, draws="philox")
    _check_values(helpers.flatten_results(res), helpers.oracle_flat(out, gold["sets"], gold["metrics"]), 1e-8,
                  name + " philox", err_rtol=1e-6)


@pytest.mark.parametrize("name", ["equity_cva_det_greeks", "equity_cva_single_det_greeks", "equity_cva_greeks"])
def test_equity_book_cva_sensitivities_match_reference_autograd(name):
    """differentiate=True on the equity CVA books (mcre/hybrid.py:EquityCreditGreeks): CVA, EPE and PV gradients with
    respect to the market model's parameters against the reference's autograd (exposure metrics through its float32
    regression chain: 2e-5; PV pathwise: 1e-8).  Deterministic intensity: the credit model's parameters are outside the
    reference's graph (None).  Stochastic intensity correlated with the assets (`equity_cva_greeks`): per-path default
    weights, and the credit model's own parameters through the weights' tangents (mcre_eq_credit_weight_tangents).
    Native Philox against the oracle's duals."""
    gold = helpers.load_golden(name)
    res, sc = helpers.run_cuda(name, draws="torch")
    flat = helpers.flatten_results(res)
    ref = {k: (np.array(v), np.array(gold["errors"][k])) for k, v in gold["values"].items()}
    _check_values(flat, ref, 1e-9, name)

    def check(res, rows_of, rtol_expo, what):
        for s_ in gold["sets"]:
            for m in gold["metrics"]:
                rows, got = rows_of(s_, m), res.get_derivatives(s_, m)
                rtol = 1e-8 if m == "pv" else rtol_expo
                for ev, row in enumerate(rows):
                    if row is None:
                        continue
                    scale = max(1.0, max(abs(w) for w in row if w is not None))
                    for pname, g, w in zip(gold["params"], got[ev], row):
                        if w is None:
                            assert g is None, f"{what} {s_}|{m}[{ev}] d/d{pname}: expected None"
                        else:
                            assert g is not None and abs(float(g) - w) <= rtol * scale, f"{what} {s_}|{m}[{ev}] d/d{pname}: {g} vs {w}"
    check(res, lambda s_, m: gold["derivatives"][f"{s_}|{m}"], 2e-5, name)
    res, _ = helpers.run_cuda(name, draws="philox")
    out, _ = helpers.run_oracle(name, draws="philox")
    none_of = {(s_, m): [[w is None for w in row] for row in gold["derivatives"][f"{s_}|{m}"]] for s_ in gold["sets"] for m in gold["metrics"]}

    def oracle_rows(s_, m):
        si, mi = gold["sets"].index(s_), gold["metrics"].index(m)
        rows = []
        for ev, g in enumerate(out["grads"][si][mi]):
            rows.append(None if g is None else [None if none_of[(s_, m)][ev][k] else float(g[k]) for k in range(len(g))])
        return rows
    check(res, oracle_rows, 1e-6, name + " philox")


def test_pfe_sensitivities_of_an_equity_book_match_reference_autograd():
    """Gradient of PFE order statistics (pfe_metric.py:59-71 under autograd = the pathwise gradient of the selected path),
    next to EPE and PV, for thresholded and MPoR-collateralised sets of European options (analytic exposures): against the
    reference's autograd with its draws injected, against the oracle under native Philox."""
    name = "bs_pfe_greeks"
    gold = helpers.load_golden(name)
    res, sc = helpers.run_cuda(name, draws="torch")
    ref = {k: (np.array(v), np.array(gold["errors"][k])) for k, v in gold["values"].items()}
    flat = helpers.flatten_results(res)
    for key, (vb, eb) in ref.items():
        helpers.assert_close(flat[key][0], vb, 1e-9, 1e-9, f"{name} {key}")
    helpers.assert_gradients(res, gold["derivatives"], gold["params"], gold["sets"], gold["metrics"], lambda m: 1e-8, name)
    res, _ = helpers.run_cuda(name, draws="philox")
    out, _ = helpers.run_oracle(name, draws="philox")
    like = {f"{s_}|{m}": [None if g is None else [float(x) for x in g] for g in out["grads"][si][mi]]
            for si, s_ in enumerate(gold["sets"]) for mi, m in enumerate(gold["metrics"])}
    helpers.assert_gradients(res, like, gold["params"], gold["sets"], gold["metrics"], lambda m: 1e-7, name + " philox")


@pytest.mark.parametrize("name", ["bs_hessian", "bs_hessian_euler", "bs_hessian_multi"])
def test_pathwise_hessians_match_reference_double_backward(name):
    """compute_higher_derivatives() on Monte Carlo present values (controller.py:253-255, 631-648): the kernel carries
    Dual2<3> numbers per lane (csrc/dual2.cuh).  Against the reference's autograd-of-autograd with its draws injected,
    and against the oracle's second-order forward mode under native Philox."""
    gold = helpers.load_golden(name)
    res, sc = helpers.run_cuda(name, draws="torch")
    flat = helpers.flatten_results(res)
    for key, vb in gold["values"].items():
        helpers.assert_close(flat[key][0], np.array(vb), 1e-9, 1e-9, f"{name} {key}")
    helpers.assert_gradients(res, gold["derivatives"], gold["params"], gold["sets"], gold["metrics"], lambda m: 1e-9, name)
    helpers.assert_hessians(lambda si, mi, ev: res.second_derivatives[si][mi][ev], gold["second_derivatives"], gold["sets"],
                            gold["metrics"], 1e-9, name)
    # the reference's accessor shapes (simulation_results.py:261-334)
    s0, p = gold["sets"][0], gold["params"]
    named = res.get_second_derivatives(s0, "pv", evaluation_idx=0)
    assert abs(float(named[p[1]][p[-1]]) - gold["second_derivatives"][f"{s0}|pv"][0][1][len(p) - 1]) <= 1e-8 * 300
    res, _ = helpers.run_cuda(name, draws="philox")
    out, _ = helpers.run_oracle(name, draws="philox")
    like = {f"{s_}|{m}": [out["hess"][si][mi][ev].tolist() for ev in range(len(out["hess"][si][mi]))]
            for si, s_ in enumerate(gold["sets"]) for mi, m in enumerate(gold["metrics"])}
    helpers.assert_hessians(lambda si, mi, ev: res.second_derivatives[si][mi][ev], like, gold["sets"], gold["metrics"], 1e-8,
                            name + " philox")


def test_second_order_requests_outside_the_supported_books_raise():
    ns = cases.Namespace()
    bs = ns.BlackScholesModel(0.0, 100.0, 0.05, 0.2)
    am = ns.AmericanOption(ns.Equity("id"), 1.0, 4, 100.0, ns.OptionType.PUT)
    sc = ns.SimulationController([ns.NettingSet(name="a", products=[am])], bs, ns.RiskMetrics([ns.PVMetric()]), 1024, 1024, 1,
                                 ns.SimulationScheme.ANALYTICAL, True)
    sc.compute_higher_derivatives()
    with pytest.raises(NotImplementedError):
        sc.run_simulation()
    heston = ns.HestonModel(0.0, 100.0, 0.03, 0.4, -0.7, 2.0, 0.04, 0.04)
    eu = ns.EuropeanOption(ns.Equity("id"), 1.0, 100.0, ns.OptionType.CALL)
    sc = ns.SimulationController([ns.NettingSet(name="e", products=[eu])], heston, ns.RiskMetrics([ns.PVMetric()]), 1024, 0, 2,
                                 ns.SimulationScheme.EULER, True)
    sc.compute_higher_derivatives()
    with pytest.raises(NotImplementedError):
        sc.run_simulation()


def test_exposure_sensitivities_of_a_book_split_over_launches():
    """Six products, four of them path-dependent, per netting set: more than one launch with tangents tracks, so the
    values and the per-path duals of several launches are netted (mcre/hybrid.py:EquityCreditGreeks) - thresholded and
    MPoR-collateralised, against the reference's autograd."""
    name = "bs_split_book_greeks"
    gold = helpers.load_golden(name)
    res, sc = helpers.run_cuda(name, draws="torch")
    flat = helpers.flatten_results(res)
    for key, vb in gold["values"].items():
        helpers.assert_close(flat[key][0], np.array(vb), 1e-9, 1e-9, f"{name} {key}")
    helpers.assert_gradients(res, gold["derivatives"], gold["params"], gold["sets"], gold["metrics"],
                             lambda m: 1e-8 if m.startswith("pv") else 2e-5, name)


@pytest.mark.parametrize("name", ["heston_exposure_greeks_qe", "heston_exposure_greeks_euler"])
def test_exposure_sensitivities_under_heston_match_reference_autograd(name):
    """EEPE / EPE / ENE / PV gradients w.r.t. the seven Heston parameters through the regression proxy (tangent
    pre-simulation + differentiated normal equations, csrc/equity.cu exposure tangents on the Heston build): thresholded
    and MPoR-collateralised sets, QE and Euler, against the reference's autograd with its draws injected and against the
    oracle's duals under native Philox."""
    gold = helpers.load_golden(name)
    res, sc = helpers.run_cuda(name, draws="torch")
    flat = helpers.flatten_results(res)
    for key, vb in gold["values"].items():
        helpers.assert_close(flat[key][0], np.array(vb), 1e-9, 1e-9, f"{name} {key}")
    # (the reference regresses in float32: 2e-5 for the Black-Scholes books; the QE scheme's fuzzy branch weights
    # amplify that noise a little - the worst entry here differs by 2.1e-5 of its row's scale)
    helpers.assert_gradients(res, gold["derivatives"], gold["params"], gold["sets"], gold["metrics"],
                             lambda m: 1e-8 if m.startswith("pv") else 5e-5, name)
    res, _ = helpers.run_cuda(name, draws="philox")
    out, _ = helpers.run_oracle(name, draws="philox")
    # (parameters outside the reference's graph - rho under EULER - are None in the golden and here, 0.0 in the oracle)
    like = {}
    for si, s_ in enumerate(gold["sets"]):
        for mi, m in enumerate(gold["metrics"]):
            rows = []
            for ev, g in enumerate(out["grads"][si][mi]):
                ref_row = gold["derivatives"][f"{s_}|{m}"][ev]
                rows.append(None if g is None else [None if r is None else float(x) for x, r in zip(g, ref_row)])
            like[f"{s_}|{m}"] = rows
    helpers.assert_gradients(res, like, gold["params"], gold["sets"], gold["metrics"], lambda m: 1e-6, name + " philox")
