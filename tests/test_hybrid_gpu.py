"""Netting sets that mix equity and interest-rate products under a three-model ModelConfig (Black-Scholes numeraire +
Vasicek + CIR++ credit; mcre/hybrid.py) against goldens of the unmodified reference and against the oracle."""
import numpy as np
import pytest

import cases
import helpers

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


def test_hybrid_sensitivities_raise():
    ns = cases.Namespace()
    model, sets, metrics, tl = cases.hybrid_cva(ns, n_euro=1, n_bonds=1, n_swaps=1)
    sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics, exposure_timeline=tl), 256, 256, 1,
                                 ns.SimulationScheme.EULER, True)
    with pytest.raises(NotImplementedError):
        sc.run_simulation()
