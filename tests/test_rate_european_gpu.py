"""European options on bonds and swaps under Vasicek (the reference's pv_european_bond_option.py and ee_pfe_swaption.py):
run as one-date exercise units of the interest-rate kernels (mcre/irc.py:with_single_exercise_proxies)."""
import numpy as np
import pytest

import cases
import parity_helpers as helpers

pytestmark = pytest.mark.gpu


def test_bond_option_pv_and_greeks_match_reference_golden_and_closed_form():
    gold = helpers.load_golden("bond_option_european")
    res, sc = helpers.run_cuda("bond_option_european", draws="torch")
    pv = helpers.flatten_results(res)["EuropeanOption|pv"]
    helpers.assert_close(pv[0], gold["values"]["EuropeanOption|pv"], 1e-10, 1e-12, "bond option pv")
    helpers.assert_close(pv[1], gold["errors"]["EuropeanOption|pv"], 1e-8, 1e-12, "bond option mc error")
    got = [float(g) for g in res.get_derivatives("EuropeanOption", "pv")[0]]
    helpers.assert_close(got, gold["derivatives"]["EuropeanOption|pv"][0], 1e-8, 1e-10, "bond option greeks")
    # the Monte Carlo value sits within 4 standard errors of Jamshidian's closed form (pv_european_bond_option.py:52, 78)
    opt = sc.products[0]
    exact = float(opt.compute_pv_bond_option_analytically(sc.model))
    assert abs(float(pv[0][0]) - exact) < 4.0 * float(pv[1][0]) + 2e-4      # (+ the Euler bias of 10 steps per year)


@pytest.mark.parametrize("draws", ["torch", "philox"])
def test_european_swaption_exposures(draws):
    name = "swaption_european"
    res, sc = helpers.run_cuda(name, draws=draws)
    flat = helpers.flatten_results(res)
    if draws == "torch":
        gold = helpers.load_golden(name)
        want = {k: (np.array(v), np.array(gold["errors"][k])) for k, v in gold["values"].items()}
    else:
        out, _ = helpers.run_oracle(name, draws="philox")
        want = helpers.oracle_flat(out, res.get_netting_set_names(), res.get_metric_names())
    for key, (vals, errs) in want.items():
        helpers.assert_close(flat[key][0], vals, 1e-9, 1e-11, f"{name} {draws} {key}")
        if not key.split("|")[1].startswith("pfe"):
            helpers.assert_close(flat[key][1], errs, 1e-6, 1e-11, f"{name} {draws} {key} mc error")
    # the European product keeps its one-state coefficient tensor: the alive state of its proxy
    opt = sc.netting_sets[0].products[0]
    coeffs = sc.regression_coeffs[opt.product_id]
    assert tuple(coeffs.shape) == (len(sc.exposure_timeline), 1, 3) and float(coeffs.abs().sum()) > 0.0
