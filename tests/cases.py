"""Shared scenario builders for the tests: each returns this repo's host objects for one
of the reference's test / benchmark configurations."""
import math

import numpy as np

HAZARDS = {0.5: 0.006402303360855854, 1.0: 0.01553038972325307, 2.0: 0.009729741230773657,
           3.0: 0.015552544648116201, 4.0: 0.021196186202801115, 5.0: 0.02284319986706472,
           7.0: 0.010111423894480876, 10.0: 0.00613267811172937, 15.0: 0.0036969930706003337,
           20.0: 0.003791311459217732}


def wwr_cva(ns_module, rho=0.3, maturity=10.0, n_expo=41, extra_metrics=True, vol=0.2, speed=0.02,
            deterministic=False):
    """Vasicek + CIR++ payer swap, CVA (+PV, EPE) - tests/pytests/test_cva.py:113-182."""
    m = ns_module
    vas = m.VasicekModel(0., 0.03, 0.05, speed, vol, asset_id="irs")
    cir = m.CIRPPModel(0., "GM", HAZARDS, 0.1, 0.01, 0.02, 0.0001, deterministic=deterministic)
    model = m.ModelConfig([vas, cir], inter_asset_correlation_matrix=np.array([rho]))
    irs = m.InterestRateSwap(0.0, maturity, 1.0, 0.03, 0.25, 0.25, m.IRSType.PAYER, asset_id="irs")
    sets = [m.NettingSet(name="irs", products=[irs], counterparty_id="GM")]
    metrics = [m.CVAMetric("GM", 0.4)]
    if extra_metrics:
        metrics += [m.PVMetric(), m.EPEMetric()]
    return model, sets, metrics, np.linspace(0, maturity, n_expo)


def vasicek_irs_collateral(ns_module, mpor=0.25, threshold=0.0, n_dates=21, maturity=5.0):
    """Vasicek payer IRS, uncollateralised + MPoR-collateralised netting sets, full metric
    set - tests/exposure_tests/ee_pfe_swap_collateralized.py:58-107 (config 2)."""
    m = ns_module
    model = m.VasicekModel(0., 0.03, 0.05, 0.02, 0.02)
    a = m.InterestRateSwap(0.0, maturity, 1.0, 0.03, 0.25, 0.25, m.IRSType.PAYER)
    b = m.InterestRateSwap(0.0, maturity, 1.0, 0.03, 0.25, 0.25, m.IRSType.PAYER)
    sets = [m.NettingSet(name="irs_uncollateralized", products=[a], threshold=threshold),
            m.NettingSet(name="irs_collateralized", products=[b], margin_period_of_risk=mpor, threshold=threshold)]
    metrics = [m.PVMetric(), m.CEMetric(), m.EPEMetric(), m.ENEMetric(), m.EEPEMetric(), m.PFEMetric(0.95)]
    return model, sets, metrics, np.linspace(0.0, maturity, n_dates)


def bs_european(ns_module, spot=120.0, sigma=0.2, rate=0.05, strike=100.0, T=2.0):
    """Config 1: BS European call PV + pathwise delta/vega/rho (tests/pytests/test_pv_european_option.py:36-85)."""
    m = ns_module
    model = m.BlackScholesModel(0, spot, rate, sigma)
    opt = m.EuropeanOption(m.Equity(), T, strike, m.OptionType.CALL)
    return model, [m.NettingSet(name="call", products=[opt])], [m.PVMetric()], None


def bermudan_swaption(ns_module, n_ex=8, a=0.002, vol=0.2):
    """Config 4 (reduced): Vasicek Bermudan payer swaption, EPE + PFE
    (tests/exposure_tests/ee_pfe_bermudan_swaption.py:21-68)."""
    m = ns_module
    model = m.VasicekModel(0., 0.03, 0.05, a, vol)
    last = 0.25 * n_ex
    swap = m.InterestRateSwap(0.0, last + 1.0, 1.0, 0.03, 0.25, 0.25, m.IRSType.PAYER)
    ex = [0.25 * (i + 1) for i in range(n_ex)]
    opt = m.BermudanOption(swap, ex, 0.0, m.OptionType.CALL)
    tl = np.array([0.25 * i for i in range(n_ex + 1)])
    return model, [m.NettingSet(name="bermudan", products=[opt])], [m.EPEMetric(), m.PFEMetric(0.95), m.PVMetric()], tl


def heston_european(ns_module):
    """Heston QE European call (tests/pytests/test_pv_european_option_heston.py:44-74)."""
    m = ns_module
    model = m.HestonModel(0.0, 100.0, 0.03, 0.4, -0.7, 2.0, 0.04, 0.04)
    opt = m.EuropeanOption(m.Equity(), 1.0, 100.0, m.OptionType.CALL)
    return model, [m.NettingSet(name="call", products=[opt])], [m.PVMetric()], None


def heston_path_dependent(ns_module):
    """Config 5 (single-asset part the reference pins): Heston QE up-and-out barrier call and
    arithmetic Asian call, 13 monthly monitoring dates."""
    m = ns_module
    model = m.HestonModel(0.0, 100.0, 0.03, 0.4, -0.7, 2.0, 0.04, 0.04)
    bar = m.BarrierOption(0.0, 1.0, 100.0, 13, m.OptionType.CALL, 140.0, m.BarrierOptionType.UPANDOUT)
    asian = m.AsianOption(0.0, 1.0, 100.0, 13, m.OptionType.CALL)
    return model, [m.NettingSet(name="barrier", products=[bar]), m.NettingSet(name="asian", products=[asian])], [m.PVMetric()], None


def heston_basket5(ns_module, n_assets=5, n_obs=13, rho_spot=0.5):
    """Config 5: up-and-out barrier call and arithmetic Asian call on an equally weighted basket of
    correlated Heston assets (QE).  BUILD-DEFINED EXTENSION: the reference's ModelConfig cannot hold
    Heston models and its barrier / Asian options monitor one asset - parity unpinned
    beyond the single-asset case, the oracle composes the pinned single-asset pieces."""
    m = ns_module
    ids = [f"h{i}" for i in range(n_assets)]
    models = [m.HestonModel(0.0, 100.0, 0.03, 0.4, -0.7, 2.0, 0.04, 0.04, asset_id=a) for a in ids]
    n_pairs = n_assets * (n_assets - 1) // 2
    model = m.ModelConfig(models, inter_asset_correlation_matrix=np.array([[rho_spot] for _ in range(n_pairs)]))
    w = [1.0 / n_assets] * n_assets
    bar = m.BarrierOption(0.0, 1.0, 100.0, n_obs, m.OptionType.CALL, 140.0, m.BarrierOptionType.UPANDOUT, basket=(ids, w))
    asian = m.AsianOption(0.0, 1.0, 100.0, n_obs, m.OptionType.CALL, basket=(ids, w))
    return model, [m.NettingSet(name="barrier", products=[bar]), m.NettingSet(name="asian", products=[asian])], [m.PVMetric()], None


def bs_basket(ns_module, euler=False):
    """4 x BS in a ModelConfig, arithmetic + geometric basket (tests/pytests/test_model_config.py:18-126)."""
    m = ns_module
    ids = ["asset1", "asset2", "asset3", "asset4"]
    models = [m.BlackScholesModel(0.0, 100.0, 0.0, 0.4, asset_id=a) for a in ids]
    model = m.ModelConfig(models, inter_asset_correlation_matrix=np.array([[0.5] for _ in range(6)]))
    w = [0.25] * 4
    b1 = m.BasketOption(1.0, ids, w, 100, m.OptionType.CALL, m.BasketOptionType.ARITHMETIC, False)
    b1.name = "basket_arithmetic"
    b2 = m.BasketOption(1.0, ids, w, 100, m.OptionType.CALL, m.BasketOptionType.GEOMETRIC)
    b2.name = "basket_geometric"
    return model, [m.NettingSet(name=b1.get_name(), products=[b1]), m.NettingSet(name=b2.get_name(), products=[b2])], [m.PVMetric()], None


#: name -> (builder, builder kwargs, run kwargs)
def flexicall_bs(ns_module, exposure=False, rights=2):
    """FlexiCall on Black-Scholes: 3 European calls, 2 exercise rights
    (tests/pytests/test_single_product_executor_parity.py "flexicall" case); rights > 2: 8 calls of alternating strikes, of
    which `rights` may be exercised (tests/exposure_tests/ee_pfe_flexicall.py: 4 rights)."""
    m = ns_module
    model = m.BlackScholesModel(0.0, 100.0, 0.03, 0.2, asset_id="asset")
    if rights > 2:
        unders = [m.EuropeanOption(m.Equity("asset"), 0.25 * (i + 1), 96.0 + 3.0 * (i % 4), m.OptionType.CALL, asset_id="asset")
                  for i in range(8)]
        flexi = m.FlexiCall(underlyings=unders, num_exercise_rights=rights, asset_id="asset")
        sets = [m.NettingSet(name="flexicall", products=[flexi])]
        return model, sets, [m.PVMetric(), m.EPEMetric(), m.PFEMetric(0.9)], np.linspace(0.0, 2.0, 9)
    flexi = m.FlexiCall(underlyings=[m.EuropeanOption(m.Equity("asset"), 0.5, 95.0, m.OptionType.CALL, asset_id="asset"),
                                     m.EuropeanOption(m.Equity("asset"), 1.0, 100.0, m.OptionType.CALL, asset_id="asset"),
                                     m.EuropeanOption(m.Equity("asset"), 1.5, 105.0, m.OptionType.CALL, asset_id="asset")],
                        num_exercise_rights=2, asset_id="asset")
    sets = [m.NettingSet(name="flexicall", products=[flexi])]
    if exposure:
        return model, sets, [m.PVMetric(), m.EPEMetric(), m.PFEMetric(0.9)], np.linspace(0.0, 1.5, 7)
    return model, sets, [m.PVMetric()], None


def bs_bridge_barrier(ns_module):
    """Barrier options with the Brownian-bridge correction between monitoring dates (barrier_option.py:138-222):
    single and double barrier, knock-out and knock-in, on one Black-Scholes asset."""
    m = ns_module
    model = m.BlackScholesModel(0.0, 100.0, 0.03, 0.25)
    B = m.BarrierOptionType
    specs = [("up_out", m.OptionType.CALL, 125.0, B.UPANDOUT, None, None), ("down_in", m.OptionType.PUT, 85.0, B.DOWNANDIN, None, None),
             ("double", m.OptionType.CALL, 130.0, B.UPANDOUT, 80.0, B.DOWNANDOUT), ("up_in", m.OptionType.CALL, 115.0, B.UPANDIN, None, None)]
    sets = []
    for name, ot, b1, t1, b2, t2 in specs:
        opt = m.BarrierOption(0.0, 1.0, 100.0, 7, ot, b1, t1, b2, t2)
        opt.set_use_brownian_bridge()
        sets.append(m.NettingSet(name=name, products=[opt]))
    return model, sets, [m.PVMetric()], None


def mixed_book(ns_module, exposure=True):
    """Mixed equity book on a 2-asset BlackScholesMulti: Europeans, Americans, a FlexiCall and a barrier option in
    one netting set (tests/pytests/test_netting_sets.py:375-528)."""
    m = ns_module
    ids = ["asset_1", "asset_2"]
    model = m.BlackScholesMulti(calibration_date=0.0, rate=0.03, asset_ids=ids, spots=[100.0, 105.0],
                                volatilities=[0.20, 0.24], correlation_matrix=np.array([[1.0, 0.35], [0.35, 1.0]]))
    products = [
        m.EuropeanOption(m.Equity("asset_1"), 1.0, 95.0, m.OptionType.CALL, asset_id="asset_1"),
        m.EuropeanOption(m.Equity("asset_2"), 1.5, 110.0, m.OptionType.PUT, asset_id="asset_2"),
        m.AmericanOption(underlying=m.Equity("asset_1"), maturity=1.0, num_exercise_dates=8, strike=100.0,
                         option_type=m.OptionType.PUT, asset_id="asset_1"),
        m.AmericanOption(underlying=m.Equity("asset_2"), maturity=1.5, num_exercise_dates=12, strike=102.5,
                         option_type=m.OptionType.CALL, asset_id="asset_2"),
        m.FlexiCall(underlyings=[m.EuropeanOption(m.Equity("asset_1"), 0.5, 95.0, m.OptionType.CALL, asset_id="asset_1"),
                                 m.EuropeanOption(m.Equity("asset_1"), 1.0, 100.0, m.OptionType.CALL, asset_id="asset_1"),
                                 m.EuropeanOption(m.Equity("asset_1"), 1.5, 105.0, m.OptionType.CALL, asset_id="asset_1")],
                    num_exercise_rights=2, asset_id="asset_1"),
        m.BarrierOption(startdate=0.0, maturity=1.25, strike=100.0, num_observation_timepoints=12,
                        option_type=m.OptionType.CALL, barrier1=130.0,
                        barrier_option_type1=m.BarrierOptionType.UPANDOUT, asset_id="asset_2"),
    ]
    sets = [m.NettingSet(name="mixed_ns", products=products)]
    if exposure:
        return model, sets, [m.EPEMetric(), m.PFEMetric(0.95)], np.linspace(0.0, 1.5, 6)
    return model, sets, [m.PVMetric()], None


def bs_exposure_greeks(ns_module, multi=True, pfe=False):
    """EPE / PV sensitivities of European options through the analytic Black-Scholes exposure
    (european_option.py:123-145, controller.py:609-627; the shape of tests/exposure_tests/eepe_simulation.py with the
    metric set that keeps the analytic branch): thresholded and MPoR-collateralised netting sets."""
    m = ns_module
    if multi:
        ids = ["asset_1", "asset_2"]
        model = m.BlackScholesMulti(calibration_date=0.0, rate=0.03, asset_ids=ids, spots=[100.0, 105.0],
                                    volatilities=[0.20, 0.24], correlation_matrix=np.array([[1.0, 0.35], [0.35, 1.0]]))
    else:
        ids = ["asset", "asset"]
        model = m.BlackScholesModel(0.0, 100.0, 0.05, 0.2, asset_id="asset")

    def book():
        return [m.EuropeanOption(m.Equity(ids[0]), 1.0, 95.0, m.OptionType.CALL, asset_id=ids[0]),
                m.EuropeanOption(m.Equity(ids[1]), 1.5, 110.0, m.OptionType.PUT, asset_id=ids[1])]
    sets = [m.NettingSet(name="thresholded", products=book(), threshold=12.0),
            m.NettingSet(name="collateralised", products=book(), margin_period_of_risk=0.25, threshold=1.0)]
    return model, sets, [m.PVMetric(), m.EPEMetric()] + ([m.PFEMetric(0.9)] if pfe else []), np.linspace(0.0, 1.5, 7)


def bs_hessian(ns_module, multi=False):
    """Pathwise second-order sensitivities of Monte Carlo present values (compute_higher_derivatives, controller.py:253-255,
    631-648: autograd of every first derivative): products that pay once, with the payoff smoothing differentiate=True
    switches on (binary / barrier indicators become fuzzy), so that the Hessian has curvature to carry."""
    m = ns_module
    if multi:
        ids = ["asset_1", "asset_2"]
        model = m.BlackScholesMulti(calibration_date=0.0, rate=0.03, asset_ids=ids, spots=[100.0, 105.0],
                                    volatilities=[0.20, 0.24], correlation_matrix=np.array([[1.0, 0.35], [0.35, 1.0]]))
    else:
        ids = ["asset", "asset"]
        model = m.BlackScholesModel(0.0, 100.0, 0.05, 0.2, asset_id="asset")
    vanilla = [m.EuropeanOption(m.Equity(ids[0]), 1.0, 95.0, m.OptionType.CALL, asset_id=ids[0]),
               m.BinaryOption(1.5, 100.0, 10.0, m.OptionType.PUT, asset_id=ids[1])]
    paths = [m.BarrierOption(0.0, 1.0, 100.0, 5, m.OptionType.CALL, 135.0, m.BarrierOptionType.UPANDOUT, asset_id=ids[1]),
             m.AsianOption(0.0, 1.0, 100.0, 5, m.OptionType.CALL, asset_id=ids[0])]
    sets = [m.NettingSet(name="vanilla", products=vanilla), m.NettingSet(name="path_dependent", products=paths)]
    return model, sets, [m.PVMetric()], None


def heston_exposure_greeks(ns_module):
    """Sensitivities of exposure metrics under Heston (no closed-form exposure: every product goes through the regression
    proxy on the spot, controller.py:294-383, 438-447, 609-627): a European put, a binary call and an Asian call in a
    thresholded and an MPoR-collateralised netting set."""
    m = ns_module
    model = m.HestonModel(0.0, 100.0, 0.03, 0.4, -0.7, 2.0, 0.04, 0.04)

    def book():
        return [m.EuropeanOption(m.Equity("id"), 1.0, 105.0, m.OptionType.PUT),
                m.BinaryOption(1.5, 100.0, 10.0, m.OptionType.CALL),
                m.AsianOption(0.0, 1.0, 100.0, 5, m.OptionType.CALL)]
    sets = [m.NettingSet(name="thresholded", products=book(), threshold=3.0),
            m.NettingSet(name="collateralised", products=book(), margin_period_of_risk=0.25, threshold=1.0)]
    return model, sets, [m.EEPEMetric(), m.EPEMetric(), m.ENEMetric(), m.PVMetric()], np.linspace(0.0, 1.5, 7)


def bs_split_book_greeks(ns_module):
    """Exposure sensitivities of a netting set with more path-dependent products than one launch with tangents tracks
    (two), so that the book is split over launches: the reference nets whatever the set holds (controller.py:438-447)."""
    m = ns_module
    model = m.BlackScholesModel(0.0, 100.0, 0.04, 0.25, asset_id="asset")

    def book():
        return [m.EuropeanOption(m.Equity("asset"), 1.0, 95.0, m.OptionType.PUT, asset_id="asset"),
                m.BarrierOption(0.0, 1.0, 100.0, 5, m.OptionType.CALL, 140.0, m.BarrierOptionType.UPANDOUT, asset_id="asset"),
                m.BarrierOption(0.0, 1.0, 105.0, 5, m.OptionType.PUT, 75.0, m.BarrierOptionType.DOWNANDOUT, asset_id="asset"),
                m.AsianOption(0.0, 1.0, 100.0, 5, m.OptionType.CALL, asset_id="asset"),
                m.AsianOption(0.0, 0.75, 98.0, 4, m.OptionType.PUT, asset_id="asset"),
                m.BinaryOption(0.5, 100.0, 10.0, m.OptionType.CALL, asset_id="asset")]
    sets = [m.NettingSet(name="split", products=book(), threshold=4.0),
            m.NettingSet(name="split_collateralised", products=book(), margin_period_of_risk=0.25, threshold=1.0)]
    return model, sets, [m.PVMetric(), m.EPEMetric(), m.ENEMetric()], np.linspace(0.0, 1.0, 5)


def bs_eepe_greeks(ns_module, book="european"):
    """Sensitivities of exposure metrics through the regression proxy (controller.py:294-383, 438-447, 609-627).
    book "european": tests/exposure_tests/eepe_simulation.py (EEPE of a Black-Scholes call; EEPE switches the analytic
    exposure branch off).  book "mixed": single-asset products that pay once on a 2-asset BlackScholesMulti in a
    thresholded and an MPoR-collateralised netting set."""
    m = ns_module
    if book == "european":
        model = m.BlackScholesModel(0, 100.0, 0.05, 0.2)
        opt = m.EuropeanOption(underlying=m.Equity("id"), exercise_date=2.0, strike=100, option_type=m.OptionType.CALL)
        return model, [m.NettingSet(name="eepe_option_ns", products=[opt])], [m.EEPEMetric(), m.EPEMetric(), m.PVMetric()], np.linspace(0.0, 2.0, 10)
    ids = ["asset_1", "asset_2"]
    model = m.BlackScholesMulti(calibration_date=0.0, rate=0.03, asset_ids=ids, spots=[100.0, 105.0],
                                volatilities=[0.20, 0.24], correlation_matrix=np.array([[1.0, 0.35], [0.35, 1.0]]))

    def products():
        return [m.EuropeanOption(m.Equity("asset_1"), 1.0, 95.0, m.OptionType.CALL, asset_id="asset_1"),
                m.BinaryOption(1.5, 100.0, 10.0, m.OptionType.PUT, asset_id="asset_2"),
                m.AsianOption(0.0, 1.0, 100.0, 5, m.OptionType.CALL, asset_id="asset_1")]
    sets = [m.NettingSet(name="thresholded", products=products(), threshold=6.0),
            m.NettingSet(name="collateralised", products=products(), margin_period_of_risk=0.25, threshold=1.0)]
    return model, sets, [m.PVMetric(), m.CEMetric(), m.EPEMetric(), m.ENEMetric(), m.EEPEMetric()], np.linspace(0.0, 1.5, 7)


def equity_cva(ns_module, rho=0.2, deterministic=False, single=False):
    """CVA of an equity book: Black-Scholes market model + CIR++ credit model of the counterparty in one ModelConfig
    (tests/exposure_tests/cva_perfprmance_large_netting_set.py:69-193, reduced), regression-proxy exposures,
    MPoR-collateralised and uncollateralised netting sets."""
    m = ns_module
    if single:
        ids = ["asset_0", "asset_0"]
        market = m.BlackScholesModel(0.0, 100.0, 0.03, 0.2, asset_id="asset_0")
        inter = [np.array([rho])]
    else:
        ids = ["asset_0", "asset_1"]
        market = m.BlackScholesMulti(calibration_date=0.0, rate=0.03, asset_ids=ids, spots=[95.0, 102.5],
                                     volatilities=[0.18, 0.21], correlation_matrix=np.array([[1.0, 0.35], [0.35, 1.0]]))
        inter = [np.full((2, 1), rho, dtype=float)]
    credit = m.CIRPPModel(calibration_date=0.0, asset_id="cp", hazard_rates=HAZARDS, kappa=0.10, theta=0.01,
                          volatility=0.02, y0=0.0001, deterministic=deterministic)
    model = m.ModelConfig(models=[market, credit], inter_asset_correlation_matrix=inter)

    def products():
        return [m.EuropeanOption(m.Equity(ids[0]), 1.0, 95.0, m.OptionType.CALL, asset_id=ids[0]),
                m.EuropeanOption(m.Equity(ids[1]), 1.5, 110.0, m.OptionType.PUT, asset_id=ids[1]),
                m.BinaryOption(1.25, 100.0, 10.0, m.OptionType.CALL, asset_id=ids[1]),
                m.AsianOption(0.0, 1.0, 100.0, 5, m.OptionType.CALL, asset_id=ids[0])]
    sets = [m.NettingSet(name="collateralised", products=products(), counterparty_id="cp", margin_period_of_risk=0.25),
            m.NettingSet(name="open", products=products(), counterparty_id="cp", threshold=2.0)]
    return model, sets, [m.CVAMetric("cp", 0.4), m.EPEMetric(), m.PVMetric()], np.linspace(0.0, 1.5, 7)


def equity_cva_exercise(ns_module):
    """The mixed equity book of tests/pytests/test_netting_sets.py:375-528 (Europeans, Americans, a FlexiCall, a
    barrier option) facing a counterparty with a CIR++ intensity: CVA + EPE of an MPoR-collateralised netting set, the
    product mix of tests/exposure_tests/cva_perfprmance_large_netting_set.py."""
    m = ns_module
    model0, sets0, _, _ = mixed_book(ns_module, exposure=True)
    credit = m.CIRPPModel(calibration_date=0.0, asset_id="cp", hazard_rates=HAZARDS, kappa=0.10, theta=0.01,
                          volatility=0.02, y0=0.0001)
    model = m.ModelConfig(models=[model0, credit], inter_asset_correlation_matrix=[np.full((2, 1), 0.2, dtype=float)])
    sets = [m.NettingSet(name="mixed_cva", products=sets0[0].products, counterparty_id="cp", margin_period_of_risk=0.25)]
    return model, sets, [m.CVAMetric("cp", 0.4), m.EPEMetric()], np.linspace(0.0, 1.5, 7)


def big_cva_book(ns_module):
    """More tracked products (Asians, barriers, an American) than one value-only launch holds, facing a counterparty
    with a CIR++ intensity: CVA + EPE + PV of an MPoR-collateralised, thresholded netting set (book splitting)."""
    m = ns_module
    ids = ["asset_1", "asset_2"]
    market = m.BlackScholesMulti(calibration_date=0.0, rate=0.03, asset_ids=ids, spots=[100.0, 105.0],
                                 volatilities=[0.20, 0.24], correlation_matrix=np.array([[1.0, 0.35], [0.35, 1.0]]))
    credit = m.CIRPPModel(calibration_date=0.0, asset_id="cp", hazard_rates=HAZARDS, kappa=0.10, theta=0.01,
                          volatility=0.02, y0=0.0001)
    model = m.ModelConfig(models=[market, credit], inter_asset_correlation_matrix=[np.full((2, 1), 0.2, dtype=float)])
    prods = []
    for k in range(66):
        a = ids[k % 2]
        if k % 3 == 0:
            prods.append(m.AsianOption(0.0, 0.5 + 0.25 * (k % 3), 95.0 + (k % 7), 3 + k % 3, m.OptionType.CALL if k % 2 else m.OptionType.PUT,
                                       asset_id=a))
        else:
            prods.append(m.BarrierOption(startdate=0.0, maturity=0.5 + 0.25 * (k % 4), strike=100.0, num_observation_timepoints=3 + k % 4,
                                         option_type=m.OptionType.CALL, barrier1=125.0 + k % 11,
                                         barrier_option_type1=m.BarrierOptionType.UPANDOUT, asset_id=a))
    prods.append(m.AmericanOption(underlying=m.Equity("asset_1"), maturity=1.0, num_exercise_dates=5, strike=100.0,
                                  option_type=m.OptionType.PUT, asset_id="asset_1"))
    prods.append(m.EuropeanOption(m.Equity("asset_2"), 1.0, 100.0, m.OptionType.CALL, asset_id="asset_2"))
    sets = [m.NettingSet(name="big", products=prods, counterparty_id="cp", margin_period_of_risk=0.25, threshold=2.0)]
    return model, sets, [m.CVAMetric("cp", 0.4), m.EPEMetric(), m.PVMetric()], np.linspace(0.0, 1.25, 6)


def cfg3_wwr(ns_module, rho=0.5, vol=0.2):
    """BASELINE configs[2] at its exact shape (what bench.py runs): Vasicek + CIR++ payer swap 0 -> 10y quarterly,
    exposure grid np.arange(241) / 24 (240 sub-steps, coupon dates on the grid), CVA the only metric
    (tests/pytests/test_cva.py:113-182 construction)."""
    m = ns_module
    vas = m.VasicekModel(0., 0.03, 0.05, 0.02, vol, asset_id="irs")
    cir = m.CIRPPModel(0., "GM", HAZARDS, 0.1, 0.01, 0.02, 0.0001)
    model = m.ModelConfig([vas, cir], inter_asset_correlation_matrix=np.array([rho]))
    irs = m.InterestRateSwap(0.0, 10.0, 1.0, 0.03, 0.25, 0.25, m.IRSType.PAYER, asset_id="irs")
    sets = [m.NettingSet(name="irs", products=[irs], counterparty_id="GM")]
    return model, sets, [m.CVAMetric("GM", 0.4)], np.arange(241) / 24.0


def cfg2_irs(ns_module, mpor=0.25):
    """BASELINE configs[1] at its exact shape: Vasicek payer IRS 0 -> 30y quarterly, exposure grid 0.25 * (0..120),
    uncollateralised + MPoR-collateralised netting sets (on-grid 0.25 or off-grid 10/252, which doubles the internal
    exposure grid), PV / EPE / ENE / EEPE / PFE (tests/exposure_tests/ee_pfe_swap_collateralized.py:58-107)."""
    m = ns_module
    model = m.VasicekModel(0., 0.03, 0.05, 0.02, 0.02)
    a = m.InterestRateSwap(0.0, 30.0, 1.0, 0.03, 0.25, 0.25, m.IRSType.PAYER)
    b = m.InterestRateSwap(0.0, 30.0, 1.0, 0.03, 0.25, 0.25, m.IRSType.PAYER)
    sets = [m.NettingSet(name="irs_uncollateralized", products=[a]),
            m.NettingSet(name="irs_collateralized", products=[b], margin_period_of_risk=mpor)]
    metrics = [m.PVMetric(), m.EPEMetric(), m.ENEMetric(), m.EEPEMetric(), m.PFEMetric(0.95)]
    return model, sets, metrics, np.arange(121) * 0.25


def schwartz_book(ns_module):
    """Schwartz two-factor standalone (schwartz_two_factor.py:124-196): the model cannot enter a ModelConfig and its
    only reference consumer is the storage product; European / binary / Asian payoffs on it pin the model step."""
    m = ns_module
    model = m.SchwartzTwoFactorModel(0.0, [0.0, 0.5, 1.0, 2.0], [50.0, 52.0, 51.0, 55.0], 0.03, 1.2, 0.4, 0.02, 0.15, 0.3)
    sets = [m.NettingSet("call", [m.EuropeanOption(m.Equity(), 1.0, 50.0, m.OptionType.CALL)]),
            m.NettingSet("binary", [m.BinaryOption(1.5, 52.0, 10.0, m.OptionType.PUT)]),
            m.NettingSet("asian", [m.AsianOption(0.25, 1.0, 51.0, 4, m.OptionType.CALL)])]
    return model, sets, [m.PVMetric()], None


def heston_euler_book(ns_module):
    """Heston under the EULER scheme (heston.py:99-121; the QE scheme has its own goldens)."""
    m = ns_module
    model = m.HestonModel(0.0, 100.0, 0.03, 0.4, -0.7, 2.0, 0.04, 0.04)
    sets = [m.NettingSet("call", [m.EuropeanOption(m.Equity(), 1.0, 100.0, m.OptionType.CALL)]),
            m.NettingSet("asian", [m.AsianOption(0.0, 1.0, 100.0, 5, m.OptionType.PUT)])]
    return model, sets, [m.PVMetric()], None


def hybrid_cva(ns_module, n_euro=8, n_bonds=4, n_swaps=40, spot=100.0, rate_level=0.03, deterministic=True,
               rho=(0.0, 0.0, 0.0), horizon=4.0, n_expo=30, extra_metrics=False, collateral=False, pfe=True):
    """Equity options + bonds + swaps in ONE netting set against a three-model ModelConfig: Black-Scholes (numeraire),
    Vasicek, CIR++ credit (tests/exposure_tests/cva_large_netting_set_derivatives.py:58-175, the book of
    tests/pytests/test_cva_large_netting_set_aad_vs_fd.py:26-57).  rho = inter-model correlations in the reference's
    order (equity-rates, equity-credit, rates-credit)."""
    m = ns_module
    cp = "large_counterparty"
    prods = []
    mats, strikes = np.linspace(0.5, 3.0, 8), np.linspace(0.85, 1.15, 10)
    for i in range(n_euro):
        o = m.EuropeanOption(m.Equity("equity"), float(mats[i % 8]), 100.0 * float(strikes[i % 10]), m.OptionType.CALL, asset_id="equity")
        o.name = f"large_european_call_{i}"
        prods.append(o)
    bmats, coupons = np.linspace(2.0, 6.0, 8), np.linspace(0.018, 0.030, 5)
    for i in range(n_bonds):
        b = m.Bond(startdate=0.0, maturity=float(bmats[i % 8]), notional=2.0, tenor=0.5, pays_notional=True,
                   fixed_rate=float(coupons[i % 5]), asset_id="rates")
        b.name = f"large_bond_{i}"
        prods.append(b)
    fixed = np.linspace(0.019, 0.031, 6)
    for i in range(n_swaps):
        sw = m.InterestRateSwap(startdate=0.0, enddate=float(bmats[i % 8]), notional=25.0, fixed_rate=float(fixed[i % 6]),
                                tenor_fixed=0.5, tenor_float=0.25, irs_type=m.IRSType.PAYER, asset_id="rates")
        sw.name = f"large_swap_{i}"
        prods.append(sw)
    eq = m.BlackScholesModel(calibration_date=0.0, spot=spot, rate=rate_level, sigma=0.22, asset_id="equity")
    rates = m.VasicekModel(calibration_date=0.0, rate=rate_level, mean=0.03, mean_reversion_speed=1.0, volatility=0.01,
                           asset_id="rates")
    credit = m.CIRPPModel(calibration_date=0.0, asset_id=cp, hazard_rates=HAZARDS, kappa=0.10, theta=0.01, volatility=0.02,
                          y0=0.0001, deterministic=deterministic)
    model = m.ModelConfig(models=[eq, rates, credit], inter_asset_correlation_matrix=[np.array([r]) for r in rho])
    metrics = [m.CVAMetric(counterparty_id=cp, recovery_rate=0.4)]
    if extra_metrics:
        metrics += [m.EPEMetric(), m.PVMetric()]
    if collateral:
        # the same book twice: thresholded, and MPoR-collateralised with a threshold; more metrics
        metrics += [m.ENEMetric()] + ([m.PFEMetric(0.9)] if pfe else [])
        tl = np.linspace(0.0, horizon, n_expo)
        _, twin, _, _ = hybrid_cva(ns_module, n_euro, n_bonds, n_swaps, spot, rate_level, deterministic, rho, horizon, n_expo)
        sets = [m.NettingSet(name="open", products=prods, counterparty_id=cp, threshold=1.5),
                m.NettingSet(name="margined", products=twin[0].products, counterparty_id=cp, threshold=0.5,
                             margin_period_of_risk=float(tl[1] - tl[0]))]
        return model, sets, metrics, tl
    return model, [m.NettingSet(name="large_cva_ns", products=prods, counterparty_id=cp)], metrics, np.linspace(0.0, horizon, n_expo)


def rate_european(ns_module, kind="bond_option"):
    """European options on rate underlyings under Vasicek: "bond_option" = a call on a zero-coupon bond, PV with pathwise
    Greeks (tests/pv_tests/pv_european_bond_option.py:41-63); "swaption" = a European receiver swaption with EPE / PFE / PV
    (tests/exposure_tests/ee_pfe_swaption.py:23-60), netted with a payer swap in a second, collateralised set."""
    m = ns_module
    if kind == "bond_option":
        model = m.VasicekModel(calibration_date=0., rate=0.03, mean=0.05, mean_reversion_speed=0.02, volatility=0.02)
        bond = m.Bond(startdate=0.0, maturity=2.0, notional=1.0, tenor=2.0, pays_notional=True, fixed_rate=0.0)
        opt = m.EuropeanOption(underlying=bond, exercise_date=1.0, strike=0.93, option_type=m.OptionType.CALL)
        return model, [m.NettingSet(name=opt.get_name(), products=[opt])], [m.PVMetric()], None
    model = m.VasicekModel(calibration_date=0., rate=0.03, mean=0.05, mean_reversion_speed=0.02, volatility=0.02)

    def swaption():
        und = m.InterestRateSwap(startdate=0.0, enddate=2.0, notional=1.0, fixed_rate=0.03, tenor_fixed=0.25, tenor_float=0.25,
                                 irs_type=m.IRSType.RECEIVER)
        return m.EuropeanOption(underlying=und, exercise_date=1.5, strike=0.0, option_type=m.OptionType.CALL)
    swap = m.InterestRateSwap(0.0, 2.0, 1.0, 0.03, 0.25, 0.25, m.IRSType.PAYER)
    sets = [m.NettingSet(name="swaption_ns", products=[swaption()]),
            m.NettingSet(name="hedged", products=[swaption(), swap], margin_period_of_risk=0.25, threshold=0.001)]
    return model, sets, [m.EPEMetric(), m.PFEMetric(0.9), m.PVMetric()], np.linspace(0.0, 3.0, 25)


def storage_exposure(ns_module, mixed=False):
    """Exposure profiles of a storage (tests/exposure_tests/ee_pfe_storage.py: EPE + PFE on an exposure grid that does not
    coincide with the daily decisions), here with ENE and PV, an MPoR-collateralised twin set, and - `mixed` - the storages
    netted with equity options on a multi-asset Black-Scholes model (tests/exposure_tests/ee_performance_large_netting_set.py)."""
    m = ns_module
    if mixed:
        model, sets, _, _ = storage_mixed_book(ns_module)
        _, twin, _, _ = storage_mixed_book(ns_module)
        sets = [m.NettingSet(name="open", products=sets[0].products, threshold=5.0),
                m.NettingSet(name="margined", products=twin[0].products, margin_period_of_risk=0.125)]
        tl = np.linspace(0.0, 1.5, 13)
    else:
        model, sets, _, _ = storage_s2f(ns_module, which="storage2", end_day=40, num_states=6, vols=(0.5, 0.2))
        _, twin, _, _ = storage_s2f(ns_module, which="storage2", end_day=40, num_states=6, vols=(0.5, 0.2))
        sets = [m.NettingSet(name="open", products=sets[0].products),
                m.NettingSet(name="margined", products=twin[0].products, margin_period_of_risk=4.0, threshold=1000.0)]
        tl = np.linspace(0.0, 44.0, 12)      # 4-day grid against daily decisions; the last dates lie beyond the contract
    return model, sets, [m.EPEMetric(), m.ENEMetric(), m.PFEMetric(0.95), m.PVMetric()], tl


def storage_cva_mixed(ns_module):
    """CVA + EPE of the mixed book (storages + equity options) against a counterparty with a CIR++ intensity correlated
    with the assets, MPoR-collateralised, EULER scheme: the shape of tests/exposure_tests/cva_perfprmance_large_netting_set.py."""
    m = ns_module
    market, sets, _, _ = storage_mixed_book(ns_module)
    credit = m.CIRPPModel(calibration_date=0.0, asset_id="cp", hazard_rates=HAZARDS, kappa=0.10, theta=0.01, volatility=0.02,
                          y0=0.0001)
    model = m.ModelConfig(models=[market, credit], inter_asset_correlation_matrix=[np.full((3, 1), 0.2, dtype=float)])
    nset = m.NettingSet(name="mixed_cva", products=sets[0].products, counterparty_id="cp", margin_period_of_risk=10 / 252)
    return model, [nset], [m.CVAMetric("cp", 0.4), m.EPEMetric()], np.linspace(0.0, 1.5, 7)


def storage_small(ns_module, model_kind="bs", num_states=4, end_day=2.0):
    """The storage of tests/pytests/test_single_product_executor_parity.py:43-60 / 162-168 (Black-Scholes "gas" price,
    PV with pathwise sensitivities), and the same contract over 12 days on a Schwartz two-factor curve."""
    m = ns_module
    cfg = m.StorageConfig()
    cfg.add_volume_constraint(0.0, end_day, 0.0, 6.0, 0.0)
    cfg.add_injection_flexibility(0.0, end_day, 0.0, 2.0)
    cfg.add_withdrawal_flexibility(0.0, end_day, 0.0, 2.0)
    cfg.add_variable_injection_cost(0.0, 0.1)
    cfg.add_variable_withdrawal_cost(0.0, 0.1)
    st = m.Storage(asset_id="gas", start_date=0.0, end_date=end_day, initial_amount=1.0, storage_config=cfg, num_states=num_states)
    st.name = "storage"
    if model_kind == "bs":
        model = m.BlackScholesModel(0.0, 100.0, 0.03, 0.2, asset_id="gas")
    else:
        model = m.SchwartzTwoFactorModel(0.0, [0.0, 4.0, 9.0, 12.0], [20.0, 23.0, 19.0, 22.0], 0.002, 0.35, 0.08, 0.001, 0.03, 0.3,
                                         asset_id="gas")
    return model, [m.NettingSet(name="storage", products=[st])], [m.PVMetric()], None


def book_storage(m, i, asset, maturity, capacity, rollout, states, initial, inj_cost, wd_cost):
    """One storage of the reference's large mixed book (tests/pv_tests/pv_performance_large_netting_set.py:43-90): three
    inventory bands, two-knot rate curves that change at the band edges, costs stepping up by 10 %."""
    cfg = m.StorageConfig()
    up, flat = 0.35 * maturity, 0.70 * maturity
    cfg.add_volume_constraint(0.0, up, 0.0, 0.55 * capacity, 0.0)
    cfg.add_volume_constraint(up, flat, 0.10 * capacity, 0.85 * capacity, 0.0)
    cfg.add_volume_constraint(flat, maturity, 0.0, capacity, 0.0)
    for a, b, lo, hi in ((0.0, up, 0.30, 0.18), (up, maturity, 0.22, 0.12)):
        cfg.add_injection_flexibility(a, b, 0.0, lo * capacity)
        cfg.add_injection_flexibility(a, b, 0.60 * capacity, hi * capacity)
    for a, b, lo, hi in ((0.0, flat, 0.16, 0.24), (flat, maturity, 0.24, 0.32)):
        cfg.add_withdrawal_flexibility(a, b, 0.0, lo * capacity)
        cfg.add_withdrawal_flexibility(a, b, 0.60 * capacity, hi * capacity)
    cfg.add_variable_injection_cost(0.0, inj_cost)
    cfg.add_variable_injection_cost(flat, inj_cost * 1.10)
    cfg.add_variable_withdrawal_cost(0.0, wd_cost)
    cfg.add_variable_withdrawal_cost(flat, wd_cost * 1.10)
    st = m.Storage(asset_id=asset, start_date=0.0, end_date=maturity, initial_amount=initial, storage_config=cfg,
                   num_states=states, rollout_interval=rollout)
    st.name = f"storage_{i}"
    return st


def storage_mixed_book(ns_module):
    """Storages next to European / American / Asian options in ONE netting set on a multi-asset Black-Scholes model: the
    shape of tests/pv_tests/pv_performance_large_netting_set.py:43-90, 228-251 (three inventory bands, two-knot rate
    curves, costs stepping up, roll-out intervals of 0.05 - 0.125 years) at a size the reference runs in seconds."""
    m = ns_module
    ids = ["asset_0", "asset_1", "asset_2"]
    model = m.BlackScholesMulti(calibration_date=0.0, rate=0.03, asset_ids=ids, spots=[95.0, 102.5, 110.0],
                                volatilities=[0.18, 0.21, 0.24],
                                correlation_matrix=np.array([[1.0, 0.35, 0.35], [0.35, 1.0, 0.35], [0.35, 0.35, 1.0]]))

    def storage(i, asset, maturity, capacity, rollout, states):
        return book_storage(m, i, asset, maturity, capacity, rollout, states, 2.0 + 0.5 * i, 0.10 + 0.02 * i, 0.08 + 0.015 * i)
    prods = [m.EuropeanOption(m.Equity(ids[0]), 1.0, 95.0, m.OptionType.CALL, asset_id=ids[0]),
             storage(0, ids[1], 1.0, 18.0, 0.05, 6),
             m.AmericanOption(underlying=m.Equity(ids[2]), maturity=1.0, num_exercise_dates=6, strike=108.0,
                              option_type=m.OptionType.PUT, asset_id=ids[2]),
             storage(1, ids[2], 1.5, 26.0, 0.125, 7),
             m.AsianOption(0.0, 0.75, 100.0, 6, m.OptionType.CALL, asset_id=ids[1]),
             storage(2, ids[0], 1.0, 34.0, 0.10, 8)]
    return model, [m.NettingSet(name="mixed_state_dependent_book", products=prods)], [m.PVMetric()], None


def _days(a, b):
    import datetime
    return float((datetime.date(*b) - datetime.date(*a)).days)


def storage_s2f(ns_module, which="storage1", end_day=None, num_states=10, vols=None):
    """Gas storage on a Schwartz two-factor curve, daily decisions (tests/storage_s2f_cases.py:139-214 and
    tests/pytests/test_storage_s2f_pv.py:23-52): "storage1" = two months, one flat 90-unit band, nearly
    deterministic prices; "storage2" = 15 months, five inventory bands, level-dependent rate curves.
    `end_day` cuts the contract short (smaller cases); `vols` = (short-term, long-term) annual volatilities."""
    m = ns_module
    if which == "storage1":
        start, end = (2024, 12, 1), (2025, 1, 31)
        n_days = _days(start, end)
        bands = [((2024, 12, 1), (2025, 2, 1), 0.0, 90.0)]
        inj = [((2024, 12, 1), (2025, 2, 1), 0.0, 90.0)]
        wd = [((2024, 12, 1), (2025, 2, 1), 0.0, 90.0)]
        costs = (0.2, 0.0)
        curve = [(0.0, 100.0), (float(int(round(n_days * 0.25))), 100.0), (float(int(round(n_days * 0.55))), 110.0),
                 (n_days, 112.0)]
        kappa, sig_s, sig_l = 8.0, 0.00001, 0.00005
    else:
        start, end = (2025, 1, 1), (2026, 3, 31)
        n_days = _days(start, end)
        bands = [((2025, 1, 1), (2025, 7, 1), 0.0, 200000.0), ((2025, 7, 1), (2025, 10, 1), 50000.0, 260000.0),
                 ((2025, 10, 1), (2026, 1, 1), 180000.0, 280000.0), ((2026, 1, 1), (2026, 3, 1), 40000.0, 260000.0),
                 ((2026, 3, 1), (2026, 4, 1), 0.0, 260000.0)]
        levels = (0.0, 60000.0, 150000.0, 225000.0)
        first, second = ((2025, 1, 1), (2025, 10, 1)), ((2025, 10, 1), (2026, 4, 1))
        inj = [(*first, lv, r) for lv, r in zip(levels, (3400.0, 2920.0, 2200.0, 1480.0))] + \
              [(*second, lv, r) for lv, r in zip(levels, (5800.0, 4840.0, 3400.0, 1960.0))]
        wd = [(*first, lv, r) for lv, r in zip(levels, (1720.0, 2800.0, 3880.0, 4600.0))] + \
             [(*second, lv, r) for lv, r in zip(levels, (2200.0, 4000.0, 5800.0, 7000.0))]
        costs = (0.35, 0.12)
        curve = [(0.0, 90.0), (_days(start, (2025, 4, 1)), 94.0), (_days(start, (2025, 7, 1)), 88.0),
                 (_days(start, (2025, 10, 1)), 96.0), (_days(start, (2026, 1, 1)), 104.0), (n_days, 98.0)]
        kappa, sig_s, sig_l = 1.5, 0.18, 0.08
    if vols is not None:
        sig_s, sig_l = vols
    cfg = m.StorageConfig()
    for a, b, lo, hi in bands:
        cfg.add_volume_constraint(_days(start, a), _days(start, b), lo, hi, 0.0)
    for a, b, level, rate in inj:
        cfg.add_injection_flexibility(_days(start, a), _days(start, b), level, rate)
    for a, b, level, rate in wd:
        cfg.add_withdrawal_flexibility(_days(start, a), _days(start, b), level, rate)
    cfg.add_variable_injection_cost(0.0, costs[0])
    cfg.add_variable_withdrawal_cost(0.0, costs[1])
    storage = m.Storage(asset_id="thegasprice", start_date=0.0, end_date=n_days if end_day is None else float(end_day),
                        initial_amount=0.0, storage_config=cfg, num_states=num_states, rollout_interval=1.0)
    model = m.SchwartzTwoFactorModel(calibration_date=0.0, curve_times=[t for t, _ in curve],
                                     curve_values=[v for _, v in curve], rate=0.0 / 365.0,
                                     short_term_mean_reversion=kappa / 365.0, short_term_vol=sig_s / math.sqrt(365.0),
                                     long_term_drift=0.0 / 365.0, long_term_vol=sig_l / math.sqrt(365.0), rho=0.2,
                                     asset_id="thegasprice")
    return model, [m.NettingSet(name=storage.get_name(), products=[storage])], [m.PVMetric()], None


GOLDEN_CASES = {
    "wwr_cva": (wwr_cva, dict(rho=0.3), dict(n_main=4096, n_pre=4096, num_steps=2, scheme="EULER", differentiate=False)),
    "wwr_cva_neg": (wwr_cva, dict(rho=-0.9, extra_metrics=False), dict(n_main=2048, n_pre=2048, num_steps=1, scheme="EULER", differentiate=False)),
    "wwr_cva_greeks": (wwr_cva, dict(rho=0.5, n_expo=11, maturity=2.5), dict(n_main=1024, n_pre=1024, num_steps=1, scheme="EULER", differentiate=True)),
    "irs_collateral_greeks": (vasicek_irs_collateral, dict(mpor=0.25, threshold=0.002, n_dates=9, maturity=2.0), dict(n_main=1024, n_pre=1024, num_steps=1, scheme="EULER", differentiate=True)),
    "cva_deterministic": (wwr_cva, dict(rho=0.0, deterministic=True, n_expo=21, maturity=5.0), dict(n_main=2048, n_pre=2048, num_steps=1, scheme="EULER", differentiate=False)),
    "irs_collateral": (vasicek_irs_collateral, dict(mpor=0.25), dict(n_main=4096, n_pre=4096, num_steps=1, scheme="EULER", differentiate=False)),
    "irs_collateral_offgrid": (vasicek_irs_collateral, dict(mpor=10 / 252, threshold=0.005), dict(n_main=4096, n_pre=4096, num_steps=1, scheme="EULER", differentiate=False)),
    "irs_analytical": (vasicek_irs_collateral, dict(mpor=0.5, n_dates=11, maturity=2.5), dict(n_main=2048, n_pre=2048, num_steps=3, scheme="ANALYTICAL", differentiate=False)),
    "bs_european": (bs_european, dict(), dict(n_main=100000, n_pre=0, num_steps=1, scheme="ANALYTICAL", differentiate=True)),
    "bs_european_euler": (bs_european, dict(T=1.0), dict(n_main=20000, n_pre=0, num_steps=8, scheme="EULER", differentiate=True)),
    "bermudan_swaption": (bermudan_swaption, dict(), dict(n_main=4096, n_pre=4096, num_steps=1, scheme="EULER", differentiate=False)),
    "heston_european": (heston_european, dict(), dict(n_main=8192, n_pre=0, num_steps=20, scheme="QE", differentiate=False)),
    "heston_european_greeks": (heston_european, dict(), dict(n_main=4096, n_pre=0, num_steps=10, scheme="QE", differentiate=True)),
    "heston_path_dependent": (heston_path_dependent, dict(), dict(n_main=4096, n_pre=0, num_steps=4, scheme="QE", differentiate=True)),
    "bs_basket": (bs_basket, dict(), dict(n_main=8192, n_pre=0, num_steps=1, scheme="ANALYTICAL", differentiate=False)),
    "bs_bridge_barrier": (bs_bridge_barrier, dict(), dict(n_main=4096, n_pre=0, num_steps=2, scheme="EULER", differentiate=False)),
    # ... with pathwise Greeks: the crossing probabilities and fuzzy hit indicators are differentiable (tests/pv_tests/pv_barrier_option.py)
    "bs_bridge_barrier_greeks": (bs_bridge_barrier, dict(), dict(n_main=4096, n_pre=0, num_steps=1, scheme="ANALYTICAL", differentiate=True)),
    "flexicall_pv": (flexicall_bs, dict(), dict(n_main=4096, n_pre=4096, num_steps=1, scheme="ANALYTICAL", differentiate=False)),
    "flexicall_exposure": (flexicall_bs, dict(exposure=True), dict(n_main=2048, n_pre=2048, num_steps=2, scheme="EULER", differentiate=False)),
    "flexicall_4_rights": (flexicall_bs, dict(rights=4), dict(n_main=2048, n_pre=2048, num_steps=1, scheme="ANALYTICAL", differentiate=False)),
    "flexicall_6_rights": (flexicall_bs, dict(rights=6), dict(n_main=1024, n_pre=2048, num_steps=2, scheme="EULER", differentiate=False)),
    "mixed_book_pv": (mixed_book, dict(exposure=False), dict(n_main=1024, n_pre=1024, num_steps=1, scheme="ANALYTICAL", differentiate=False)),
    "mixed_book_exposure": (mixed_book, dict(exposure=True), dict(n_main=512, n_pre=512, num_steps=1, scheme="ANALYTICAL", differentiate=False)),
    "bs_exposure_greeks": (bs_exposure_greeks, dict(), dict(n_main=2048, n_pre=0, num_steps=1, scheme="ANALYTICAL", differentiate=True)),
    "bs_exposure_greeks_euler": (bs_exposure_greeks, dict(multi=False), dict(n_main=2048, n_pre=0, num_steps=3, scheme="EULER", differentiate=True)),
    # sensitivities of PFE order statistics (pathwise gradient of the selected path): equity book, hybrid book
    "bs_pfe_greeks": (bs_exposure_greeks, dict(pfe=True), dict(n_main=2048, n_pre=0, num_steps=1, scheme="ANALYTICAL", differentiate=True)),
    "hybrid_pfe_greeks": (hybrid_cva, dict(n_euro=2, n_bonds=1, n_swaps=3, rho=(0.25, 0.0, 0.0), horizon=2.0, n_expo=9, extra_metrics=True, collateral=True, pfe=True),
                          dict(n_main=1024, n_pre=1024, num_steps=2, scheme="EULER", differentiate=True)),
    # pathwise Hessians of Monte Carlo present values (compute_higher_derivatives): exact and Euler steps, one and two assets
    "bs_hessian": (bs_hessian, dict(), dict(n_main=2048, n_pre=0, num_steps=1, scheme="ANALYTICAL", differentiate=True, second_order=True)),
    "bs_hessian_euler": (bs_hessian, dict(), dict(n_main=2048, n_pre=0, num_steps=3, scheme="EULER", differentiate=True, second_order=True)),
    "bs_hessian_multi": (bs_hessian, dict(multi=True), dict(n_main=2048, n_pre=0, num_steps=2, scheme="EULER", differentiate=True, second_order=True)),
    "heston_exposure_greeks_qe": (heston_exposure_greeks, dict(), dict(n_main=2048, n_pre=2048, num_steps=3, scheme="QE", differentiate=True)),
    "heston_exposure_greeks_euler": (heston_exposure_greeks, dict(), dict(n_main=2048, n_pre=2048, num_steps=3, scheme="EULER", differentiate=True)),
    "bs_split_book_greeks": (bs_split_book_greeks, dict(), dict(n_main=2048, n_pre=2048, num_steps=2, scheme="EULER", differentiate=True)),
    "bs_eepe_greeks": (bs_eepe_greeks, dict(), dict(n_main=4096, n_pre=1000, num_steps=1, scheme="ANALYTICAL", differentiate=True)),
    "bs_proxy_greeks_mixed": (bs_eepe_greeks, dict(book="mixed"), dict(n_main=2048, n_pre=2048, num_steps=2, scheme="EULER", differentiate=True)),
    "equity_cva": (equity_cva, dict(), dict(n_main=2048, n_pre=2048, num_steps=2, scheme="EULER", differentiate=False)),
    "equity_cva_single_det": (equity_cva, dict(rho=0.0, deterministic=True, single=True), dict(n_main=1024, n_pre=1024, num_steps=1, scheme="EULER", differentiate=False)),
    # sensitivities of CVA / EPE / PV of equity books against a counterparty with a deterministic intensity
    # stochastic intensity correlated with the assets: the credit model's own parameters enter through the default weights
    "equity_cva_greeks": (equity_cva, dict(rho=0.2), dict(n_main=1024, n_pre=1024, num_steps=2, scheme="EULER", differentiate=True)),
    "equity_cva_det_greeks": (equity_cva, dict(rho=0.0, deterministic=True), dict(n_main=1024, n_pre=1024, num_steps=2, scheme="EULER", differentiate=True)),
    "equity_cva_single_det_greeks": (equity_cva, dict(rho=0.0, deterministic=True, single=True), dict(n_main=1024, n_pre=1024, num_steps=1, scheme="EULER", differentiate=True)),
    "equity_cva_exercise": (equity_cva_exercise, dict(), dict(n_main=512, n_pre=512, num_steps=1, scheme="EULER", differentiate=False)),
    "bs_basket_euler": (bs_basket, dict(), dict(n_main=4096, n_pre=0, num_steps=5, scheme="EULER", differentiate=True)),
    # the BASELINE.json configs at their exact shapes (reduced path counts)
    "cfg3_wwr_cva": (cfg3_wwr, dict(rho=0.5), dict(n_main=32768, n_pre=32768, num_steps=1, scheme="EULER", differentiate=False)),
    # same shape with a short-rate volatility of 2 % instead of 20 %: the CVA integrand loses its heavy tail, so Monte
    # Carlo standard errors are reliable and a 3-sigma comparison of native Philox with the reference is meaningful
    "cfg3_wwr_cva_lowvol": (cfg3_wwr, dict(rho=0.5, vol=0.02), dict(n_main=16384, n_pre=16384, num_steps=1, scheme="EULER", differentiate=False)),
    "cfg2_irs_ongrid": (cfg2_irs, dict(mpor=0.25), dict(n_main=8192, n_pre=8192, num_steps=1, scheme="EULER", differentiate=False)),
    "cfg2_irs_offgrid": (cfg2_irs, dict(mpor=10 / 252), dict(n_main=8192, n_pre=8192, num_steps=1, scheme="EULER", differentiate=False)),
    "cfg4_bermudan_40": (bermudan_swaption, dict(n_ex=40), dict(n_main=8192, n_pre=8192, num_steps=1, scheme="EULER", differentiate=False)),
    # models the reference only runs standalone
    "schwartz_analytical": (schwartz_book, dict(), dict(n_main=4096, n_pre=0, num_steps=2, scheme="ANALYTICAL", differentiate=True)),
    "schwartz_euler": (schwartz_book, dict(), dict(n_main=4096, n_pre=0, num_steps=3, scheme="EULER", differentiate=True)),
    # three-model hybrid: equity + rates products netted in one set (test_cva_large_netting_set_aad_vs_fd.py at its
    # own sizes, value-only; a stochastic-credit, correlated twin with more metrics)
    "hybrid_cva": (hybrid_cva, dict(), dict(n_main=1024, n_pre=1024, num_steps=4, scheme="EULER", differentiate=False)),
    "hybrid_cva_corr": (hybrid_cva, dict(n_euro=3, n_bonds=2, n_swaps=5, deterministic=False, rho=(0.3, -0.2, 0.4), horizon=2.0, n_expo=9, extra_metrics=True),
                        dict(n_main=2048, n_pre=2048, num_steps=2, scheme="EULER", differentiate=False)),
    # first-order sensitivities of a hybrid book: the AAD leg of test_cva_large_netting_set_aad_vs_fd.py at its own
    # sizes, and a smaller book with threshold / MPoR sets and more metrics (deterministic credit)
    "hybrid_cva_greeks": (hybrid_cva, dict(), dict(n_main=1024, n_pre=1024, num_steps=4, scheme="EULER", differentiate=True)),
    # stochastic intensity correlated with both market factors: sensitivities to all three models' parameters, PFE included
    "hybrid_stochastic_greeks": (hybrid_cva, dict(n_euro=2, n_bonds=1, n_swaps=3, deterministic=False, rho=(0.25, 0.1, -0.3), horizon=2.0, n_expo=9, extra_metrics=True, collateral=True, pfe=True),
                                 dict(n_main=1024, n_pre=1024, num_steps=2, scheme="EULER", differentiate=True)),
    "hybrid_collateral_greeks": (hybrid_cva, dict(n_euro=2, n_bonds=1, n_swaps=3, rho=(0.25, 0.0, 0.0), horizon=2.0, n_expo=9, extra_metrics=True, collateral=True, pfe=False),
                                 dict(n_main=1024, n_pre=1024, num_steps=2, scheme="EULER", differentiate=True)),
    "hybrid_collateral": (hybrid_cva, dict(n_euro=2, n_bonds=1, n_swaps=3, deterministic=False, rho=(0.25, 0.1, -0.3), horizon=2.0, n_expo=9, extra_metrics=True, collateral=True),
                          dict(n_main=1024, n_pre=1024, num_steps=2, scheme="EULER", differentiate=False)),
    # gas storage (the reference's tests/pytests/test_storage_s2f_pv.py at its own sizes, and cut-down twins);
    # "degree" = polynomial degree of the regression basis (PolyomialRegression(degree), default 2)
    "storage1": (storage_s2f, dict(which="storage1"), dict(n_main=2000, n_pre=4000, num_steps=1, scheme="ANALYTICAL", differentiate=False, degree=3)),
    "storage2": (storage_s2f, dict(which="storage2"), dict(n_main=2000, n_pre=4000, num_steps=1, scheme="ANALYTICAL", differentiate=False, degree=3)),
    "storage2_short_euler": (storage_s2f, dict(which="storage2", end_day=100, num_states=6), dict(n_main=1024, n_pre=2048, num_steps=2, scheme="EULER", differentiate=False, degree=2)),
    # (cubic, not quartic: on the first dates a raw-basis quartic of spots within a few percent of 100 puts the last
    # pivot of LAPACK's rank-revealing QR within a factor 3 of its cut-off, and the fit - the reference's as much as this
    # package's - then depends on the host CPU's BLAS kernels: a quartic golden made here failed on one GPU box)
    "storage1_vol": (storage_s2f, dict(which="storage1", vols=(0.9, 0.3), num_states=5), dict(n_main=3000, n_pre=3000, num_steps=1, scheme="ANALYTICAL", differentiate=False, degree=3)),
    # pathwise PV sensitivities of storages (realised cashflows of the regression policy; decisions carry no gradient)
    "storage_bs_greeks": (storage_small, dict(), dict(n_main=256, n_pre=256, num_steps=1, scheme="ANALYTICAL", differentiate=True)),
    "storage_s2f_greeks": (storage_small, dict(model_kind="s2f", num_states=5, end_day=12.0), dict(n_main=512, n_pre=512, num_steps=2, scheme="ANALYTICAL", differentiate=True, degree=3)),
    "storage_s2f_greeks_euler": (storage_small, dict(model_kind="s2f", num_states=5, end_day=12.0), dict(n_main=512, n_pre=512, num_steps=2, scheme="EULER", differentiate=True)),
    "storage_exposure": (storage_exposure, dict(), dict(n_main=2048, n_pre=2048, num_steps=1, scheme="ANALYTICAL", differentiate=False, degree=3)),
    "storage_exposure_mixed": (storage_exposure, dict(mixed=True), dict(n_main=1000, n_pre=1000, num_steps=1, scheme="ANALYTICAL", differentiate=False)),
    "bond_option_european": (rate_european, dict(), dict(n_main=8192, n_pre=0, num_steps=10, scheme="EULER", differentiate=True)),
    "swaption_european": (rate_european, dict(kind="swaption"), dict(n_main=4096, n_pre=4096, num_steps=2, scheme="EULER", differentiate=False)),
    "storage_cva_mixed": (storage_cva_mixed, dict(), dict(n_main=1000, n_pre=1000, num_steps=2, scheme="EULER", differentiate=False)),
    "storage_mixed_book": (storage_mixed_book, dict(), dict(n_main=1000, n_pre=1000, num_steps=1, scheme="ANALYTICAL", differentiate=False)),
    "heston_euler": (heston_euler_book, dict(), dict(n_main=4096, n_pre=0, num_steps=8, scheme="EULER", differentiate=True)),
}


class Namespace:
    """Bundle of the API classes, taken either from this repo or from the reference."""

    def __init__(self):
        from common.enums import SimulationScheme
        from controller.controller import SimulationController
        from metrics.ce_metric import CEMetric
        from metrics.cva_metric import CVAMetric
        from metrics.eepe_metric import EEPEMetric
        from metrics.ene_metric import ENEMetric
        from metrics.epe_metric import EPEMetric
        from metrics.pfe_metric import PFEMetric
        from metrics.pv_metric import PVMetric
        from metrics.risk_metrics import RiskMetrics
        from models.black_scholes import BlackScholesModel
        from models.black_scholes_multi import BlackScholesMulti
        from models.cirpp import CIRPPModel
        from models.heston import HestonModel
        from models.model_config import ModelConfig
        from models.vasicek import VasicekModel
        from products.asian_option import AsianAveragingType, AsianOption
        from products.barrier_option import BarrierOption, BarrierOptionType
        from products.basket_option import BasketOption, BasketOptionType
        from products.bermudan_option import AmericanOption, BermudanOption
        from products.binary_option import BinaryOption
        from products.bond import Bond
        from products.equity import Equity
        from products.european_option import EuropeanOption
        from products.flexicall import FlexiCall
        from products.netting_set import NettingSet
        from products.product import OptionType
        from products.swap import InterestRateSwap, IRSType
        self.__dict__.update({k: v for k, v in locals().items() if k != "self"})
        try:   # dead code in the reference; an extension in this package
            from models.hull_white import HullWhiteModel
            self.HullWhiteModel = HullWhiteModel
        except Exception:  # pragma: no cover
            pass
        try:
            from maths.regression import PolyomialRegression
            from products.storage import Storage
            from products.storage_helpers import StorageConfig
            self.PolyomialRegression, self.Storage, self.StorageConfig = PolyomialRegression, Storage, StorageConfig
        except Exception:  # pragma: no cover
            pass
        try:
            from models.schwartz_two_factor import SchwartzTwoFactorModel
            self.SchwartzTwoFactorModel = SchwartzTwoFactorModel
        except Exception:  # pragma: no cover
            pass
