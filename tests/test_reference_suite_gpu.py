"""The reference's own pinned tests (tests/pytests/*.py), restated against this package's
public API: same constructors, same sizes, same thresholds.  The seeded ones pass in the reference
because of its torch.randn stream (seeds 42 / 43), so they run here in RNG compatibility mode
(`SimulationController.rng_compat = "torch"`, also MCRE_RNG=torch): the reference's normals are
regenerated on the host and fed to the kernels, every other step runs on the GPU.  The statistical ones
additionally run under native Philox with a tolerance in standard errors."""
import math

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def _compat(sc):
    sc.rng_compat = "torch"
    return sc


def test_american_option_known_answer_through_compat_mode():
    """tests/pytests/test_american_option.py:16-61: 34.323036543142706 at 1e-8."""
    ns = cases.Namespace()
    model = ns.BlackScholesModel(0.0, 100, 0.05, 0.5)
    prod = ns.AmericanOption(ns.Equity("id"), 3.0, 1000, 100.0, ns.OptionType.CALL)
    sc = _compat(ns.SimulationController([ns.NettingSet(name=prod.get_name(), products=[prod])], model,
                                         ns.RiskMetrics([ns.PVMetric()]), 10000, 100000, 1,
                                         ns.SimulationScheme.ANALYTICAL, False))
    res = sc.run_simulation()
    assert abs(float(res.get_results(prod.get_name(), "pv", evaluation_idx=0)) - 34.323036543142706) < 1e-8


def _credit_hybrid(ns, hazards, speed, rho, asset):
    vas = ns.VasicekModel(calibration_date=0., rate=0.03, mean=0.05, mean_reversion_speed=speed, volatility=0.2,
                          asset_id=asset)
    cp = "General Motors Co"
    cir = ns.CIRPPModel(calibration_date=0., y0=0.0001, theta=0.01, kappa=0.1, volatility=0.02, hazard_rates=hazards,
                        asset_id=cp)
    return vas, cir, cp, ns.ModelConfig(models=[vas, cir], inter_asset_correlation_matrix=np.array([rho]))


@pytest.mark.parametrize("rng", ["torch", "philox"])
def test_cva_corporate_bond_matches_expected_loss(hazards, rng):
    """tests/pytests/test_cva.py:35-111: zero-coupon bond, uncorrelated credit: CVA = (1-R)(1-S(0,T))P(0,T)
    within 2e-6 (reference stream); under Philox within 4 standard errors + the reference's tolerance."""
    ns = cases.Namespace()
    vas, cir, cp, model = _credit_hybrid(ns, hazards, 1, 0.0, "bond")
    T = 2.0
    bond = ns.Bond(startdate=0.0, maturity=T, notional=1, tenor=T, pays_notional=True, fixed_rate=0.0, asset_id="bond")
    metric = ns.CVAMetric(counterparty_id=cp, recovery_rate=0.4)
    sc = ns.SimulationController([ns.NettingSet(name=bond.get_name(), products=[bond], counterparty_id=cp)], model,
                                 ns.RiskMetrics(metrics=[metric], exposure_timeline=np.linspace(0, T, 100)),
                                 100000, 100000, 10, ns.SimulationScheme.EULER, False)
    sc.rng_compat = rng
    res = sc.run_simulation()
    cva = float(res.get_results(bond.get_name(), metric.get_name(), evaluation_idx=0))
    err = float(res.get_mc_error(bond.get_name(), metric.get_name(), evaluation_idx=0))
    expected = 0.6 * (1 - float(cir.survival_probability(0.0, T, 0.0001))) * float(vas.compute_bond_price(0.0, T, 0.03))
    assert abs(cva - expected) < (2e-6 if rng == "torch" else 2e-6 + 4.0 * err)


def test_cva_wwr_payer_swap_exceeds_uncorrelated_known_answer(hazards):
    """tests/pytests/test_cva.py:113-193: rho = 0.99999 CVA exceeds the uncorrelated known answer
    1.114576156484541 +- 0.0024446898428056294 by 3 sigma; and rho = 0 reproduces that known answer itself."""
    ns = cases.Namespace()
    out = {}
    for rho in (0.99999, 0.0):
        vas, cir, cp, model = _credit_hybrid(ns, hazards, 0.02, rho, "irs")
        irs = ns.InterestRateSwap(startdate=0.0, enddate=10.0, notional=1.0, fixed_rate=0.03, tenor_fixed=0.25,
                                  tenor_float=0.25, irs_type=ns.IRSType.PAYER, asset_id="irs")
        metric = ns.CVAMetric(counterparty_id=cp, recovery_rate=0.4)
        sc = _compat(ns.SimulationController(
            [ns.NettingSet(name=irs.get_name(), products=[irs], counterparty_id=cp)], model,
            ns.RiskMetrics(metrics=[metric], exposure_timeline=np.linspace(0, 10.0, 100)), 100000, 100000, 10,
            ns.SimulationScheme.EULER, False))
        res = sc.run_simulation()
        out[rho] = (float(res.get_results(irs.get_name(), metric.get_name(), evaluation_idx=0)),
                    float(res.get_mc_error(irs.get_name(), metric.get_name(), evaluation_idx=0)))
    ref, ref_err = 1.114576156484541, 0.0024446898428056294
    assert abs(out[0.0][0] - ref) < 1e-9 and abs(out[0.0][1] - ref_err) < 1e-10
    cva, err = out[0.99999]
    assert cva - ref > 3.0 * math.hypot(err, ref_err)


def _basket_book(ns, ids, control_variate):
    w = [0.25] * 4
    arith = ns.BasketOption(1.0, ids, w, 100, ns.OptionType.CALL, ns.BasketOptionType.ARITHMETIC, control_variate)
    arith.name = "basket_arithmetic"
    geo = ns.BasketOption(1.0, ids, w, 100, ns.OptionType.CALL, ns.BasketOptionType.GEOMETRIC)
    geo.name = "basket_geometric"
    return arith, geo


@pytest.mark.parametrize("which", ["config_analytical", "config_euler", "multi_control_variate"])
def test_basket_option_known_values(which):
    """tests/pytests/test_model_config.py:18-126 and test_pv_basket_option.py:16-71: 4 correlated Black-Scholes
    assets, arithmetic 12.60 / geometric 10.9551100513373 within 0.02 at 1e6 paths."""
    ns = cases.Namespace()
    ids = ["asset1", "asset2", "asset3", "asset4"]
    if which == "multi_control_variate":
        corr = np.full((4, 4), 0.5) + 0.5 * np.eye(4)
        model = ns.BlackScholesMulti(0.0, 0.0, ids, [100.0] * 4, [0.4] * 4, corr)
    else:
        model = ns.ModelConfig(models=[ns.BlackScholesModel(calibration_date=0.0, asset_id=a, spot=100.0, rate=0.0, sigma=0.4)
                                       for a in ids],
                               inter_asset_correlation_matrix=np.array([[0.5] for _ in range(6)]))
    arith, geo = _basket_book(ns, ids, which == "multi_control_variate")
    steps, scheme = (50, ns.SimulationScheme.EULER) if which == "config_euler" else (1, ns.SimulationScheme.ANALYTICAL)
    sc = _compat(ns.SimulationController([ns.NettingSet(name=arith.get_name(), products=[arith]),
                                          ns.NettingSet(name=geo.get_name(), products=[geo])], model,
                                         ns.RiskMetrics(metrics=[ns.PVMetric()]), 1000000, 0, steps, scheme, False))
    res = sc.run_simulation()
    pa = float(res.get_results(arith.get_name(), "pv", evaluation_idx=0))
    pg = float(res.get_results(geo.get_name(), "pv", evaluation_idx=0))
    assert abs(pa - 12.60) < 0.02
    target = float(geo.compute_pv_analytically(model)) if which == "multi_control_variate" else 10.9551100513373
    assert abs(pg - target) < 0.02


@pytest.mark.parametrize("rng", ["torch", "philox"])
def test_heston_qe_european_matches_semi_analytic_price(rng):
    """tests/pytests/test_pv_european_option_heston.py:76-106: QE, 50 steps, 1e6 paths vs the semi-analytic
    Heston price within 1e-3 relative."""
    ns = cases.Namespace()
    kappa, theta, sigma, rho, v0 = 0.01713417, 2.0, 0.45545583, -0.78975708, 0.0286834
    model = ns.HestonModel(0, 800, 0.04, sigma, kappa=kappa, theta=theta, v0=v0, rho=rho)
    prod = ns.EuropeanOption(underlying=ns.Equity(), exercise_date=1.0, strike=720, option_type=ns.OptionType.CALL)
    exact = float(prod.compute_pv_analytically_heston(model))
    sc = ns.SimulationController([ns.NettingSet(name=prod.get_name(), products=[prod])], model,
                                 ns.RiskMetrics(metrics=[ns.PVMetric()]), 1000000, 0, 50, ns.SimulationScheme.QE, False)
    sc.rng_compat = rng
    res = sc.run_simulation()
    price = float(res.get_results(prod.get_name(), "pv")[0])
    err = float(res.get_mc_error(prod.get_name(), "pv")[0])
    if rng == "torch":
        assert 2 * abs(price - exact) / (abs(price) + abs(exact)) < 1e-3     # the reference's own threshold
    else:
        # the reference's 1e-3 is about 1.3 standard errors at this size (it passes on its seed):
        # under Philox the bound is 4 standard errors
        assert abs(price - exact) < 4.0 * err, (price, exact, err)


@pytest.mark.parametrize("rng", ["torch", "philox"])
def test_bs_european_call_with_aad_matches_black_scholes(rng):
    """tests/pytests/test_pv_european_option.py:87-116: 1e6 paths, AAD on, price and Greeks vs closed form."""
    ns = cases.Namespace()
    model = ns.BlackScholesModel(0, 100.0, 0.05, 0.2)
    prod = ns.EuropeanOption(ns.Equity(), 1.0, 100.0, ns.OptionType.CALL)
    sc = ns.SimulationController([ns.NettingSet(name=prod.get_name(), products=[prod])], model,
                                 ns.RiskMetrics(metrics=[ns.PVMetric()]), 1000000, 0, 1, ns.SimulationScheme.ANALYTICAL, True)
    sc.rng_compat = rng
    res = sc.run_simulation()
    price = float(res.get_results(prod.get_name(), "pv")[0])
    exact = float(prod.compute_pv_analytically(model))
    if rng == "torch":
        assert abs(price - exact) / exact < 1e-3            # the reference's own threshold, on the reference's draws
    else:
        # 1e-3 relative is 0.7 standard errors at this size (the reference passes on its seed): under Philox
        # the bound is 4 standard errors
        err = float(res.get_mc_error(prod.get_name(), "pv")[0])
        assert abs(price - exact) < 4.0 * err, (price, exact, err)
    d1 = (math.log(100.0 / 100.0) + (0.05 + 0.02) * 1.0) / 0.2
    delta, vega = 0.5 * math.erfc(-d1 / math.sqrt(2)), 100.0 * math.exp(-0.5 * d1 * d1) / math.sqrt(2 * math.pi)
    g = res.get_derivatives(prod.get_name(), "pv", evaluation_idx=0)
    assert abs(float(g["spot"]) - delta) < 3e-3 and abs(float(g["volatility"]) - vega) < 0.2


@pytest.mark.parametrize("which,expected", [("storage1", 1055.330006881181), ("storage2", 3769746.378205333)])
def test_storage_s2f_pv_known_answers(which, expected):
    """tests/pytests/test_storage_s2f_pv.py:23-52: gas storage on the Schwartz two-factor model, 2000 / 4000 paths,
    one ANALYTICAL step per day, cubic regression: the pinned PVs at the reference's own tolerance (abs 1e-6)."""
    ns = cases.Namespace()
    model, sets, metrics, _ = cases.storage_s2f(ns, which=which)
    product = sets[0].products[0]
    pv_metric = metrics[0]
    sc = _compat(ns.SimulationController(netting_sets=sets, model=model, risk_metrics=ns.RiskMetrics(metrics=[pv_metric]),
                                         num_paths_mainsim=2000, num_paths_presim=4000, num_steps=1,
                                         simulation_scheme=ns.SimulationScheme.ANALYTICAL, differentiate=False,
                                         regression_function=ns.PolyomialRegression(degree=3)))
    pv = sc.run_simulation().get_results(product.get_name(), pv_metric.get_name(), evaluation_idx=0)
    assert float(pv) == pytest.approx(expected, abs=1e-6)
