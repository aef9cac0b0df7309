"""Analytic PV shortcut of the controller (controller.py:204-229, 609-648): products with a closed form under
the model skip the Monte Carlo; with differentiate=True first and (compute_higher_derivatives) second derivatives
come from host autograd on the scalar closed form, like the reference.  No GPU needed: nothing is simulated.
Mirrors tests/pytests/test_european_option_hessian.py:65-108 and test_simulation_results_named_access.py:17-95."""
import math

import numpy as np
import torch
import pytest

import cases


def _bs_hessian(S, r, sig, T, K):
    """Closed-form second derivatives of the Black-Scholes call in (spot, volatility, rate)."""
    sq = math.sqrt(T)
    d1 = (math.log(S / K) + (r + 0.5 * sig * sig) * T) / (sig * sq)
    d2 = d1 - sig * sq
    pdf = math.exp(-0.5 * d1 * d1) / math.sqrt(2 * math.pi)
    pdf2 = math.exp(-0.5 * d2 * d2) / math.sqrt(2 * math.pi)
    disc = K * math.exp(-r * T)
    gamma = pdf / (S * sig * sq)
    vega = S * pdf * sq
    return {("spot", "spot"): gamma, ("volatility", "volatility"): vega * d1 * d2 / sig,
            ("spot", "volatility"): -pdf * d2 / sig, ("spot", "rate"): pdf * sq / sig,
            ("volatility", "rate"): -T * disc * pdf2 * d1 / sig,
            ("rate", "rate"): T * disc * (pdf2 * sq / sig - T * 0.5 * math.erfc(-d2 / math.sqrt(2)))}


def _controller(ns, products_by_set, differentiate=True):
    model = ns.BlackScholesModel(0.0, 100.0, 0.05, 0.3)
    from metrics.metric import Metric
    rm = ns.RiskMetrics(metrics=[ns.PVMetric(evaluation_type=Metric.EvaluationType.ANALYTICAL)])
    sets = [ns.NettingSet(name=ps[0].get_name(), products=ps) for ps in products_by_set]
    return model, ns.SimulationController(sets, model, rm, 1, 0, 1, ns.SimulationScheme.ANALYTICAL, differentiate)


def test_analytic_pv_second_derivatives_match_closed_form():
    ns = cases.Namespace()
    prod = ns.EuropeanOption(ns.Equity("id"), 2.0, 100.0, ns.OptionType.CALL)
    model, sc = _controller(ns, [[prod]])
    sc.compute_higher_derivatives()
    res = sc.run_simulation()
    assert res.get_product_names() == ["EuropeanOption"] and res.get_metric_names() == ["pv"]
    assert res.get_model_param_names() == ["spot", "volatility", "rate"]
    hess = res.get_second_derivatives("EuropeanOption", "pv", evaluation_idx=0)
    for (a, b), want in _bs_hessian(100.0, 0.05, 0.3, 2.0, 100.0).items():
        assert float(hess[a][b]) == pytest.approx(want, rel=1e-9, abs=1e-9), (a, b)
        assert float(hess[b][a]) == pytest.approx(want, rel=1e-9, abs=1e-9), (b, a)
    assert float(hess["spot"]["spot"]) == pytest.approx(float(prod.compute_dDeltadSpot_analytically(model)), rel=1e-9)
    assert float(hess["volatility"]["volatility"]) == pytest.approx(float(prod.compute_dVegadSigma_analytically(model)), rel=1e-9)


def test_named_access_of_analytic_results_and_first_derivatives():
    ns = cases.Namespace()
    p1 = ns.EuropeanOption(ns.Equity("id_1"), 2.0, 100.0, ns.OptionType.CALL)
    p2 = ns.EuropeanOption(ns.Equity("id_2"), 2.0, 120.0, ns.OptionType.CALL)
    model, sc = _controller(ns, [[p1], [p2]])
    res = sc.run_simulation()
    assert res.get_product_names() == ["EuropeanOption", "EuropeanOption#2"]
    pv1 = float(res.get_results("EuropeanOption", "pv", evaluation_idx=0))
    pv2 = float(res.get_results("EuropeanOption#2", "pv", evaluation_idx=0))
    assert pv1 == pytest.approx(float(p1.compute_pv_analytically(model)), rel=1e-14) and pv1 != pv2
    # legacy keyword aliases (simulation_results.py)
    assert float(res.get_results(prod_idx="EuropeanOption", metric_idx="pv", evaluation_index=0)) == pv1
    vega = float(res.get_derivatives("EuropeanOption", "pv", "volatility", evaluation_idx=0))
    sq, d1 = math.sqrt(2.0), (0.05 + 0.045) * 2.0 / (0.3 * math.sqrt(2.0))
    assert vega == pytest.approx(100.0 * math.exp(-0.5 * d1 * d1) / math.sqrt(2 * math.pi) * sq, rel=1e-10)
    grads = res.get_derivatives("EuropeanOption#2", "pv", evaluation_idx=0)
    assert set(grads) == {"spot", "volatility", "rate"} and all(np.isfinite(float(v)) for v in grads.values())


def test_netting_set_analytic_pv_sums_products_and_gradients():
    ns = cases.Namespace()
    a = ns.EuropeanOption(ns.Equity(), 1.0, 100.0, ns.OptionType.CALL)
    b = ns.EuropeanOption(ns.Equity(), 2.0, 90.0, ns.OptionType.PUT)
    model, sc = _controller(ns, [[a, b]])
    res = sc.run_simulation()
    name = res.get_product_names()[0]
    want = float(a.compute_pv_analytically(model)) + float(b.compute_pv_analytically(model))
    assert float(res.get_results(name, "pv", evaluation_idx=0)) == pytest.approx(want, rel=1e-14)
    _, one = _controller(ns, [[ns.EuropeanOption(ns.Equity(), 1.0, 100.0, ns.OptionType.CALL)]])
    _, two = _controller(ns, [[ns.EuropeanOption(ns.Equity(), 2.0, 90.0, ns.OptionType.PUT)]])
    d1 = float(one.run_simulation().get_derivatives(0, "pv", "spot", evaluation_idx=0))
    d2 = float(two.run_simulation().get_derivatives(0, "pv", "spot", evaluation_idx=0))
    assert float(res.get_derivatives(name, "pv", "spot", evaluation_idx=0)) == pytest.approx(d1 + d2, rel=1e-12)


def test_heston_semi_analytic_price_limits():
    """Heston with vanishing vol-of-vol and v0 = theta is Black-Scholes with sigma = sqrt(theta)."""
    ns = cases.Namespace()
    model = ns.HestonModel(0.0, 100.0, 0.03, 1e-4, 0.0, 1.5, 0.04, 0.04)
    call = ns.EuropeanOption(ns.Equity(), 1.0, 105.0, ns.OptionType.CALL)
    put = ns.EuropeanOption(ns.Equity(), 1.0, 105.0, ns.OptionType.PUT)
    bs = ns.BlackScholesModel(0.0, 100.0, 0.03, 0.2)
    assert float(call.compute_pv_analytically_heston(model)) == pytest.approx(float(call.compute_pv_analytically(bs)), abs=1e-6)
    assert float(put.compute_pv_analytically_heston(model)) == pytest.approx(float(put.compute_pv_analytically(bs)), abs=1e-6)


def test_netting_set_collateral_profile_uses_exact_delayed_indices():
    """tests/pytests/test_netting_sets.py:209-264: the host-side tensor helpers of NettingSet."""
    import torch
    ns = cases.Namespace()
    prod = ns.EuropeanOption(ns.Equity("eq"), 1.0, 100.0, ns.OptionType.CALL)
    nset = ns.NettingSet(name="collateral_ns", products=[prod], margin_period_of_risk=0.5)
    tl = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0], dtype=torch.float64)
    netted = torch.tensor([[0.0, 0.0], [5.0, 10.0], [10.0, 20.0], [15.0, 30.0], [20.0, 40.0]], dtype=torch.float64)
    metric_idx, delayed_idx = torch.tensor([0, 2, 4]), torch.tensor([-1, 1, 3])
    coll = nset.compute_collateral_profile(netted, tl, metric_idx, delayed_idx)
    unsec = nset.compute_unsecured_exposure_profiles(netted, tl, metric_idx, delayed_idx)
    assert torch.equal(coll, torch.tensor([[0.0, 0.0], [5.0, 10.0], [15.0, 30.0]], dtype=torch.float64))
    assert torch.equal(unsec, torch.tensor([[0.0, 0.0], [5.0, 10.0], [5.0, 10.0]], dtype=torch.float64))
    banded = ns.NettingSet(name="t", products=[ns.EuropeanOption(ns.Equity("eq"), 1.0, 100.0, ns.OptionType.CALL)], threshold=7.0)
    assert torch.equal(banded.apply_threshold(torch.tensor([-10.0, -3.0, 0.0, 7.0, 12.0], dtype=torch.float64)),
                       torch.tensor([-3.0, 0.0, 0.0, 0.0, 5.0], dtype=torch.float64))


def test_constructor_error_behaviour_matches_the_reference():
    """controller.py:40-48, 89-97 and netting_set.py:23-34: what the boundary rejects, with the reference's
    exception types."""
    ns = cases.Namespace()
    model = ns.BlackScholesModel(0.0, 100.0, 0.05, 0.3)
    opt = lambda: ns.EuropeanOption(ns.Equity(), 1.0, 100.0, ns.OptionType.CALL)   # noqa: E731
    rm = ns.RiskMetrics([ns.PVMetric()])
    S = ns.SimulationScheme.ANALYTICAL
    with pytest.raises(ValueError):
        ns.SimulationController([], model, rm, 10, 0, 1, S)
    shared = opt()
    with pytest.raises(ValueError):
        ns.SimulationController([ns.NettingSet(name="a", products=[shared]), ns.NettingSet(name="b", products=[shared])],
                                model, rm, 10, 0, 1, S)
    with pytest.raises(ValueError):
        ns.NettingSet(name="empty", products=[])
    with pytest.raises(ValueError):
        ns.NettingSet(name="neg", products=[opt()], threshold=-1.0)
    with pytest.raises(ValueError):
        ns.NettingSet(name="neg", products=[opt()], margin_period_of_risk=-0.1)
    # CVA needs a ModelConfig that contains the counterparty's credit model
    cva = ns.RiskMetrics([ns.CVAMetric("cp", 0.4)], exposure_timeline=np.linspace(0.0, 1.0, 3))
    with pytest.raises(Exception):
        ns.SimulationController([ns.NettingSet(name="a", products=[opt()], counterparty_id="cp")], model, cva, 10, 10, 1, S)
    with pytest.raises(AssertionError):
        ns.FlexiCall(underlyings=[opt()], num_exercise_rights=2)
    # gas storage (storage.py:28-31): at least two inventory states, a positive roll-out interval
    from products.storage import Storage
    from products.storage_helpers import StorageConfig
    cfg = StorageConfig()
    cfg.add_volume_constraint(0.0, 1.0, 0.0, 1.0)
    with pytest.raises(ValueError):
        Storage(asset_id="gas", start_date=0.0, end_date=1.0, initial_amount=0.0, storage_config=cfg, num_states=1)


def test_unsupported_combinations_raise_instead_of_falling_back():
    """what the CUDA backends do not cover raises NotImplementedError when the backend is chosen (before any
    device work, so this runs without a GPU); nothing falls back to a CPU path."""
    ns = cases.Namespace()
    S = ns.SimulationScheme
    tl = np.linspace(0.0, 1.5, 4)

    # hybrid equity + credit: sensitivities of the CVA only under a deterministic intensity, EULER only (the reference's
    # ModelConfig has the same scheme restriction)
    def hybrid(differentiate, scheme):
        model, sets, metrics, tl_ = cases.equity_cva(ns)
        return ns.SimulationController(sets, model, ns.RiskMetrics(metrics, exposure_timeline=tl_), 64, 64, 1, scheme, differentiate)
    for differentiate, scheme in ((False, S.ANALYTICAL), (True, S.ANALYTICAL)):
        with pytest.raises(NotImplementedError):
            hybrid(differentiate, scheme).run_simulation()

    # sensitivities of exposure profiles of equity books: Black-Scholes and Heston, not the Schwartz two-factor model
    builder, kwargs, _ = cases.GOLDEN_CASES["schwartz_euler"]
    model, sets, _, _ = builder(ns, **kwargs)
    sc = ns.SimulationController(sets, model, ns.RiskMetrics([ns.EPEMetric()], exposure_timeline=tl), 64, 64, 2, S.EULER, True)
    with pytest.raises(NotImplementedError):
        sc.run_simulation()


def test_regression_basis_other_than_quadratic_raises():
    """The CUDA regression kernels implement the reference's default basis PolyomialRegression(2) ([1, x, x^2]).
    The reference builds the design matrix of ANY regression_function (controller.py:361-374); a different basis must
    raise here rather than be evaluated silently as a quadratic (round-1 review)."""
    from maths.regression import PolyomialRegression
    ns = cases.Namespace()
    model, sets, metrics, tl = cases.vasicek_irs_collateral(ns, n_dates=5, maturity=1.0)
    rm = ns.RiskMetrics(metrics, exposure_timeline=tl)
    sc = ns.SimulationController(sets, model, rm, 64, 64, 1, ns.SimulationScheme.EULER,
                                 regression_function=PolyomialRegression(3))
    with pytest.raises(NotImplementedError, match="degree=2"):
        sc.run_simulation()
    # a product that needs no regression (PV of a European option) is unaffected: the error is raised only when the
    # basis would be used - checked here on the condition itself, without a device
    bs = ns.BlackScholesModel(0.0, 100.0, 0.05, 0.2)
    opt = ns.EuropeanOption(ns.Equity(), 1.0, 100.0, ns.OptionType.CALL)
    sc = ns.SimulationController([ns.NettingSet(name="c", products=[opt])], bs, ns.RiskMetrics([ns.PVMetric()]), 64, 0, 1,
                                 ns.SimulationScheme.ANALYTICAL, regression_function=PolyomialRegression(3))
    assert not any(sc._product_requires_regression(p) for p in sc.products)


def test_host_helpers_of_rate_and_credit_models_match_the_reference():
    """Jamshidian's bond-option closed form (european_option.py:264-288) and the CIR++ intensity lambda(t) = y + psi(t)
    (cirpp.py:240-244): values computed with the unmodified reference."""
    ns = cases.Namespace()
    m = ns.VasicekModel(calibration_date=0., rate=0.03, mean=0.05, mean_reversion_speed=0.02, volatility=0.02)
    b = ns.Bond(startdate=0.0, maturity=2.0, notional=1.0, tenor=2.0, pays_notional=True, fixed_rate=0.0)
    for ot, want in ((ns.OptionType.CALL, 0.03921194857833765), (ns.OptionType.PUT, 9.768465960734857e-05)):
        o = ns.EuropeanOption(underlying=b, exercise_date=1.0, strike=0.93, option_type=ot)
        assert abs(float(o.compute_pv_bond_option_analytically(m)) - want) < 1e-13
    c = ns.CIRPPModel(calibration_date=0., asset_id="cp", hazard_rates=cases.HAZARDS, kappa=0.1, theta=0.01, volatility=0.02,
                      y0=0.0001)
    got = c.lambda_t(torch.tensor(1.7), torch.tensor([0.002, 0.01])).tolist()
    assert np.allclose(got, [0.0100823457359637, 0.01808234541745138], rtol=1e-12, atol=0)
    det = ns.CIRPPModel(calibration_date=0., asset_id="cp", hazard_rates=cases.HAZARDS, kappa=0.1, theta=0.01,
                        volatility=0.02, y0=0.0001, deterministic=True)
    assert det.lambda_t(1.7, torch.tensor([0.3], dtype=torch.float64)).tolist() == [0.3]


def test_barrier_closed_forms_match_the_reference():
    """Reiner-Rubinstein up-and-out / down-and-out calls (barrier_option.py:245-301): values of the unmodified reference."""
    ns = cases.Namespace()
    B = ns.BarrierOptionType
    for S0, bar, K, bt, want in ((100.0, 120.0, 100.0, B.UPANDOUT, 0.4093649419852321), (110.0, 130.0, 95.0, B.UPANDOUT, 1.773254975855977),
                                 (100.0, 85.0, 100.0, B.DOWNANDOUT, 9.679474963397013), (125.0, 120.0, 100.0, B.UPANDOUT, 0.0)):
        m = ns.BlackScholesModel(0, S0, 0.03, 0.25)
        p = ns.BarrierOption(startdate=0.0, maturity=1.5, strike=K, num_observation_timepoints=10, option_type=ns.OptionType.CALL,
                             barrier1=bar, barrier_option_type1=bt)
        assert abs(float(p.compute_pv_analytically(m)) - want) < 1e-12


def test_cds_bootstrap_matches_the_reference_fixture():
    """helpers/cs_helper.py (reference: src/helpers/cs_helper.py:9-107): hazards, legs and default probabilities against
    tests/golden/cs_helper.json, written by the unmodified reference (tests/golden/make_cs_helper.py)."""
    import json
    import os
    import importlib.util
    from maths.maths import bisection_search
    # (the package's `helpers` has the reference's name, which the test suite's own helpers.py shadows: load by path)
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "montecarlo-risk-engine_b200")
    spec = importlib.util.spec_from_file_location("mcre_cs_helper", os.path.join(pkg, "helpers", "cs_helper.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    CSHelper = mod.CSHelper
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cs_helper.json")))
    pay = np.arange(0.25, 20.0 + 1e-7, 0.25)
    df = np.exp(-g["rate"] * pay)
    h = CSHelper()
    haz = h.bootstrap_hazards(credit_spreads=g["spreads"], maturities=g["tenors"], payment_days=pay,
                              discount_factors_payment_days=df, recovery_rate=g["recovery"])
    assert np.allclose(haz, g["hazards"], rtol=0, atol=1e-11)
    legs = [h._compute_cds_legs(g["tenors"][:i + 1], pay, df, g["recovery"], haz[:i + 1]) for i in range(len(haz))]
    assert np.allclose(legs, g["legs"], rtol=1e-12, atol=1e-13)
    # the bootstrapped curve reprices every quote
    assert np.allclose([b / a for a, b in legs], g["spreads"], rtol=0, atol=1e-10)
    t64 = lambda x: torch.tensor(x, dtype=torch.float64)
    pd = [float(h.probability_of_default(t64(haz), t64(g["tenors"]), t64(t))) for t in g["dates"]]
    assert np.allclose(pd, g["default_probability"], rtol=0, atol=1e-14)
    assert bisection_search(lambda x: x * x + 1.0) is None
    assert abs(bisection_search(lambda x: x * x - 2.0) - 2.0 ** 0.5) < 1e-11
