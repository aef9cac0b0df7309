"""Netting sets that mix equity and interest-rate products under a three-model ModelConfig (Black-Scholes numeraire +
Vasicek + CIR++ credit; mcre/hybrid.py) against goldens of the unmodified reference and against the oracle."""
import numpy as np
import pytest

import cases
import parity_helpers as helpers

pytestmark = pytest.mark.gpu

HYBRID_CASES = ["hybrid_cva", "hybrid_cva_corr", "hybrid_collateral"]


@pytest.mark.parametrize("name", HYBRID_CASES)
def test_hybrid_book_matches_reference_golden(name):
    """The reference's torch.randn stream injected.  hybrid_cva is the book of
    tests/pytests/test_cva_large_netting_set_aad_vs_fd.py:26-57 (8 calls, 4 bonds, 40 swaps, 30 exposure dates)."""
    gold = helpers.load_golden(name)
    res, sc = helpers.run_cuda(name, draws="torch")
    flat = helpers.flatten_results(res)
    for key, ref in gold["values"].items():
        vals, errs = flat[key]
        scale = max(1.0, float(np.max(np.abs(ref))))
        # (exposure regressions pass through the reference's float32 cashflow accumulators: 1e-9 like the other goldens)
        helpers.assert_close(vals, ref, 1e-9, 1e-10 * scale, f"{name} {key}")
        if not key.split("|")[1].startswith("pfe"):
            helpers.assert_close(errs, gold["errors"][key], 1e-6, 1e-10 * scale, f"{name} {key} mc error")


@pytest.mark.parametrize("name", ["hybrid_cva_corr", "hybrid_collateral"])
def test_hybrid_book_native_philox_matches_oracle_philox(name):
    res, sc = helpers.run_cuda(name, draws="philox")
    out, (ns, model, sets, metrics, tl, rkw) = helpers.run_oracle(name, draws="philox")
    flat = helpers.flatten_results(res)
    want = helpers.oracle_flat(out, res.get_netting_set_names(), res.get_metric_names())
    for key, (vals, errs) in want.items():
        scale = max(1.0, float(np.max(np.abs(vals))))
        helpers.assert_close(flat[key][0], vals, 1e-8, 1e-9 * scale, f"{name} {key} philox")


def test_hybrid_cva_is_positive_and_moves_with_spot_and_rate():
    """Value half of tests/pytests/test_cva_large_netting_set_surface.py:26-44 and of the finite-difference leg of
    test_cva_large_netting_set_aad_vs_fd.py (bumps of spot by 1 and of both initial rates by 25 bp on common random
    numbers)."""
    ns = cases.Namespace()

    def cva(spot, rate):
        model, sets, metrics, tl = cases.hybrid_cva(ns, n_euro=10, n_bonds=5, n_swaps=50, spot=spot, rate_level=rate)
        sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics, exposure_timeline=tl), 1024, 1024, 4,
                                     ns.SimulationScheme.EULER, False)
        r = sc.run_simulation()
        return float(r.get_results("large_cva_ns", metrics[0].get_name(), evaluation_idx=0))
    base = cva(100.0, 0.03)
    assert base > 0.0
    d_spot = cva(101.0, 0.03) - base
    d_rate = (cva(100.0, 0.0325) - base) / 0.0025
    assert np.isfinite(d_spot) and np.isfinite(d_rate) and d_spot > 0.0      # calls gain with the spot


@pytest.mark.parametrize("name", ["hybrid_cva_greeks", "hybrid_collateral_greeks", "hybrid_pfe_greeks", "hybrid_stochastic_greeks"])
def test_hybrid_sensitivities_match_reference_autograd(name):
    """differentiate=True on a hybrid book: every metric's gradient with respect to the 11 parameters of the three models
    against torch.autograd of the unmodified reference.  Exposure metrics pass through the float32 regression chain on
    the reference's side (2e-5 like the other exposure Greeks), PV is pathwise (1e-8); the credit model's parameters are
    outside the reference's graph in deterministic mode (None)."""
    gold = helpers.load_golden(name)
    res, sc = helpers.run_cuda(name, draws="torch")
    flat = helpers.flatten_results(res)
    for key, ref in gold["values"].items():
        scale = max(1.0, float(np.max(np.abs(ref))))
        helpers.assert_close(flat[key][0], ref, 1e-9, 1e-10 * scale, f"{name} {key}")
    for key, rows in gold["derivatives"].items():
        s, m = key.split("|")
        got = res.get_derivatives(s, m)
        rtol = 1e-8 if m == "pv" else 2e-5
        for ev, row in enumerate(rows):
            scale = max(1.0, max(abs(w) for w in row if w is not None))
            for pname, g, w in zip(gold["params"], got[ev], row):
                if w is None:
                    assert g is None, f"{name} {key}[{ev}] d/d{pname}: expected None, got {g}"
                else:
                    assert g is not None, f"{name} {key}[{ev}] d/d{pname} is None"
                    assert abs(float(g) - w) <= rtol * scale, f"{name} {key}[{ev}] d/d{pname}: {float(g)} vs {w}"


def test_large_netting_set_cva_aad_matches_finite_differences():
    """tests/pytests/test_cva_large_netting_set_aad_vs_fd.py:26-57 restated: the AAD sensitivities of the CVA to the spot
    and to both initial rates against bump-and-revalue on common random numbers, the reference's tolerances."""
    ns = cases.Namespace()

    def run(spot, rate, diff):
        model, sets, metrics, tl = cases.hybrid_cva(ns, spot=spot, rate_level=rate)
        sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics, exposure_timeline=tl), 1024, 1024, 4,
                                     ns.SimulationScheme.EULER, diff)
        r = sc.run_simulation()
        name = metrics[0].get_name()
        cva = float(r.get_results("large_cva_ns", name, evaluation_idx=0))
        if not diff:
            return cva
        d = r.get_derivatives("large_cva_ns", name, evaluation_idx=0)
        return cva, float(d["equity.spot"]), float(d["equity.rate"]) + float(d["rates.rate"])
    cva, d_spot, d_rate = run(100.0, 0.03, True)
    base = run(100.0, 0.03, False)
    assert abs(cva - base) <= 1e-12 * abs(base)
    fd_spot = (run(101.0, 0.03, False) - base) / 1.0
    fd_rate = (run(100.0, 0.0325, False) - base) / 0.0025
    assert abs(d_spot - fd_spot) < 2e-3 and abs(d_rate - fd_rate) < 0.1


def test_unsupported_hybrid_sensitivities_raise():
    """Sensitivities of hybrid books are lowered for one Black-Scholes market model; two of them must raise, not fall back."""
    ns = cases.Namespace()
    eq1 = ns.BlackScholesModel(calibration_date=0.0, spot=100.0, rate=0.03, sigma=0.22, asset_id="equity")
    eq2 = ns.BlackScholesModel(calibration_date=0.0, spot=90.0, rate=0.03, sigma=0.3, asset_id="equity_2")
    rates = ns.VasicekModel(calibration_date=0.0, rate=0.03, mean=0.03, mean_reversion_speed=1.0, volatility=0.01, asset_id="rates")
    model = ns.ModelConfig(models=[eq1, eq2, rates], inter_asset_correlation_matrix=[np.array([0.0])] * 3)
    prods = [ns.EuropeanOption(ns.Equity("equity"), 1.0, 100.0, ns.OptionType.CALL, asset_id="equity"),
             ns.EuropeanOption(ns.Equity("equity_2"), 1.0, 90.0, ns.OptionType.PUT, asset_id="equity_2"),
             ns.Bond(startdate=0.0, maturity=2.0, notional=2.0, tenor=0.5, pays_notional=True, fixed_rate=0.02, asset_id="rates")]
    sc = ns.SimulationController([ns.NettingSet(name="s", products=prods)], model,
                                 ns.RiskMetrics([ns.EPEMetric(), ns.PVMetric()], exposure_timeline=np.linspace(0.0, 2.0, 5)),
                                 256, 256, 1, ns.SimulationScheme.EULER, True)
    with pytest.raises(NotImplementedError):
        sc.run_simulation()
