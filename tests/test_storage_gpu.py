"""Gas storage on the GPU (csrc/storage.cu through mcre/storage.py) against goldens of the unmodified reference,
against the oracle on the same draws, and through size-independent properties at large path counts."""
import numpy as np
import pytest

import cases
import parity_helpers as helpers

pytestmark = pytest.mark.gpu

STORAGE_CASES = ["storage1", "storage2", "storage2_short_euler", "storage1_vol"]


@pytest.mark.parametrize("name", STORAGE_CASES)
def test_storage_pv_matches_reference_golden(name):
    """The reference's torch.randn stream injected: PV and its standard error to 1e-9 / 1e-7 relative.  storage1 and
    storage2 are tests/pytests/test_storage_s2f_pv.py at its own sizes (1055.330006881181 / 3769746.378205333)."""
    gold = helpers.load_golden(name)
    res, sc = helpers.run_cuda(name, draws="torch")
    flat = helpers.flatten_results(res)
    for key, ref in gold["values"].items():
        vals, errs = flat[key]
        helpers.assert_close(vals, ref, 1e-9, 1e-9, f"{name} {key}")
        helpers.assert_close(errs, gold["errors"][key], 1e-7, 1e-9, f"{name} {key} mc error")
    prod = sc.products[0]
    assert tuple(prod.regression_coeffs.shape) == (len(prod.product_timeline), prod.num_states, sc.regression_function.get_degree())


@pytest.mark.parametrize("name", ["storage2_short_euler", "storage1_vol"])
def test_storage_native_philox_matches_oracle_philox(name):
    res, sc = helpers.run_cuda(name, draws="philox")
    out, _ = helpers.run_oracle(name, draws="philox")
    v, e = out["results"][0][0][0]
    got = helpers.flatten_results(res)["Storage|pv"]
    helpers.assert_close(got[0], [v], 1e-9, 1e-9, f"{name} philox value")
    helpers.assert_close(got[1], [e], 1e-7, 1e-9, f"{name} philox error")
    if name == "storage1_vol":
        # (raw-basis cubic of spots within a few percent of 100 on the first dates: scipy's gelsy (oracle, OpenBLAS) and
        # torch's (product, MKL - the reference's) need not agree on the coefficients to 1e-6 there; PVs agree all the
        # same.  The regression itself is compared on the other case.)
        return
    # regression of the product: the LAPACK solve sees the device's spots and value grid; fitted continuation values
    # at the forward curve of each date
    coeffs = np.stack(out["prod_coeffs"][0])
    mine = sc.products[0].regression_coeffs.numpy()
    model = sc.model
    x = np.array([model.curve_value(t) for t in sc.products[0].product_timeline.tolist()])
    powers = x[:, None, None] ** np.arange(coeffs.shape[2])[None, None, :]
    fit_o, fit_m = (coeffs * powers).sum(axis=2), (mine * powers).sum(axis=2)
    assert np.max(np.abs(fit_m - fit_o)) < 1e-6 * np.max(np.abs(fit_o))


@pytest.mark.parametrize("name", ["storage2_short_euler", "storage1_vol", "storage2"])
def test_storage_device_moments_solver_matches_oracle_normal_equations(name):
    """storage_regression = "moments": Gram moments of the standardised basis accumulated on the device."""
    res, sc = helpers.run_cuda(name, draws="torch", storage_solver="moments")
    out, _ = helpers.run_oracle(name, draws="torch", storage_solver="moments")
    v, e = out["results"][0][0][0]
    got = helpers.flatten_results(res)["Storage|pv"]
    helpers.assert_close(got[0], [v], 1e-8, 1e-8, f"{name} moments value")
    helpers.assert_close(got[1], [e], 1e-6, 1e-8, f"{name} moments error")


def test_two_storages_in_two_netting_sets_and_one_set_of_two():
    """Per-set accumulation: PV(set of two storages) = PV(a) + PV(b) on the same paths; each product has its own
    action dates inside the common simulation grid."""
    ns = cases.Namespace()
    model, sets_a, metrics, _ = cases.storage_s2f(ns, which="storage2", end_day=50, num_states=6)
    _, sets_b, _, _ = cases.storage_s2f(ns, which="storage2", end_day=35, num_states=8)
    a, b = sets_a[0].products[0], sets_b[0].products[0]

    def run(sets):
        sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics), 4096, 4096, 1, ns.SimulationScheme.ANALYTICAL,
                                     False, regression_function=ns.PolyomialRegression(degree=3))
        return sc.run_simulation()
    r2 = run([ns.NettingSet(name="a", products=[a]), ns.NettingSet(name="b", products=[b])])
    pa, pb = float(r2.get_results("a", "pv")[0]), float(r2.get_results("b", "pv")[0])
    _, sets_a2, _, _ = cases.storage_s2f(ns, which="storage2", end_day=50, num_states=6)
    _, sets_b2, _, _ = cases.storage_s2f(ns, which="storage2", end_day=35, num_states=8)
    r1 = run([ns.NettingSet(name="ab", products=[sets_a2[0].products[0], sets_b2[0].products[0]])])
    assert abs(float(r1.get_results("ab", "pv")[0]) - (pa + pb)) <= 1e-9 * abs(pa + pb)


def test_storage_large_run_is_deterministic_and_consistent_with_the_small_one():
    """2^18 paths in both passes, device moments solver, native Philox: bit-identical when repeated, and within
    4 standard errors of the 4096-path LAPACK run (policy bias is far below that at these sizes)."""
    ns = cases.Namespace()

    def run(n, mode):
        model, sets, metrics, _ = cases.storage_s2f(ns, which="storage2", end_day=90, num_states=8)
        sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics), n, n, 1, ns.SimulationScheme.ANALYTICAL,
                                     False, regression_function=ns.PolyomialRegression(degree=3))
        sc.storage_regression = mode
        r = sc.run_simulation()
        return float(r.get_results("Storage", "pv")[0]), float(r.get_mc_error("Storage", "pv")[0])
    big1, big2 = run(1 << 18, "moments"), run(1 << 18, "moments")
    assert big1 == big2
    small = run(4096, "lapack")
    assert abs(big1[0] - small[0]) < 4.0 * np.hypot(big1[1], small[1])
    assert big1[1] < small[1] / 6.0            # standard error shrinks like 1 / sqrt(64)


def test_storage_pv_withdraws_initial_inventory():
    """tests/pytests/test_storage.py:116-166: one unit in store, price pinned at 10, default quadratic basis: PV = 10."""
    ns = cases.Namespace()
    cfg = ns.StorageConfig()
    cfg.add_volume_constraint(0.0, 2.0, 0.0, 2.0, 0.0)
    cfg.add_injection_flexibility(0.0, 2.0, 0.0, 1.0)
    cfg.add_withdrawal_flexibility(0.0, 2.0, 0.0, 1.0)
    cfg.add_variable_injection_cost(0.0, 0.0)
    cfg.add_variable_withdrawal_cost(0.0, 0.0)
    product = ns.Storage("thegasprice", 0.0, 2.0, 1.0, cfg, num_states=3)
    model = ns.SchwartzTwoFactorModel(0.0, [0.0, 2.0], [10.0, 10.0], 0.0, 1.0, 1e-8, 0.0, 1e-8, 0.0, asset_id="thegasprice")
    for compat in ("torch", "philox"):
        sc = ns.SimulationController([ns.NettingSet(name=product.get_name(), products=[product])], model,
                                     ns.RiskMetrics([ns.PVMetric()]), 2000, 2000, 1, ns.SimulationScheme.ANALYTICAL, False)
        sc.rng_compat = compat
        pv = sc.run_simulation().get_results(product.get_name(), "pv", evaluation_idx=0)
        assert abs(float(pv) - 10.0) < 1e-3


GREEK_CASES = ["storage_bs_greeks", "storage_s2f_greeks", "storage_s2f_greeks_euler"]


@pytest.mark.parametrize("name", GREEK_CASES)
def test_storage_pathwise_greeks_match_reference_autograd(name):
    """differentiate=True: PV and its sensitivities to every model parameter against torch.autograd of the unmodified
    reference (storage_bs_greeks is the storage case of tests/pytests/test_single_product_executor_parity.py:162-168);
    parameters outside the reference's graph (rho under EULER) come back as None."""
    gold = helpers.load_golden(name)
    res, sc = helpers.run_cuda(name, draws="torch")
    flat = helpers.flatten_results(res)
    helpers.assert_close(flat["storage|pv"][0], gold["values"]["storage|pv"], 1e-9, 1e-9, f"{name} pv")
    helpers.assert_close(flat["storage|pv"][1], gold["errors"]["storage|pv"], 1e-7, 1e-9, f"{name} pv error")
    got = res.get_derivatives("storage", "pv")[0]
    want = gold["derivatives"]["storage|pv"][0]
    for pname, g, w in zip(gold["params"], got, want):
        if w is None:
            assert g is None, f"{name} d/d{pname}: expected None"
        else:
            assert g is not None and abs(float(g) - w) <= 1e-8 * max(1.0, abs(w)), f"{name} d/d{pname}: {float(g)} vs {w}"


@pytest.mark.parametrize("name", ["storage_bs_greeks", "storage_s2f_greeks"])
def test_storage_greeks_native_philox_match_oracle(name):
    res, sc = helpers.run_cuda(name, draws="philox")
    out, _ = helpers.run_oracle(name, draws="philox")
    v, e = out["results"][0][0][0]
    got = helpers.flatten_results(res)["storage|pv"]
    helpers.assert_close(got[0], [v], 1e-9, 1e-9, f"{name} philox value")
    grads = np.array([float(g) for g in res.get_derivatives("storage", "pv")[0]])
    helpers.assert_close(grads, out["grads"][0][0][0], 1e-8, 1e-8, f"{name} philox greeks")


@pytest.mark.parametrize("draws", ["torch", "philox"])
def test_storages_next_to_equity_products_in_one_netting_set(draws):
    """Three storages on different assets of a multi-asset Black-Scholes model netted with a European, an American and
    an Asian option (the book shape of tests/pv_tests/pv_performance_large_netting_set.py): against the reference's golden
    with its injected draws, against the oracle under native Philox."""
    name = "storage_mixed_book"
    res, sc = helpers.run_cuda(name, draws=draws)
    got = helpers.flatten_results(res)["mixed_state_dependent_book|pv"]
    if draws == "torch":
        gold = helpers.load_golden(name)
        v, e = gold["values"]["mixed_state_dependent_book|pv"], gold["errors"]["mixed_state_dependent_book|pv"]
    else:
        out, _ = helpers.run_oracle(name, draws="philox")
        v, e = [out["results"][0][0][0][0]], [out["results"][0][0][0][1]]
    helpers.assert_close(got[0], v, 1e-9, 1e-9, f"{name} {draws} value")
    helpers.assert_close(got[1], e, 1e-7, 1e-9, f"{name} {draws} mc error")


@pytest.mark.parametrize("name", ["storage_exposure", "storage_exposure_mixed", "storage_cva_mixed"])
@pytest.mark.parametrize("draws", ["torch", "philox"])
def test_storage_exposure_profiles(name, draws):
    """EPE / ENE / PFE / PV of storages (tests/exposure_tests/ee_pfe_storage.py) on an exposure grid that does not coincide
    with the decisions, open and MPoR-collateralised sets; alone on the Schwartz model and netted with equity options on a
    multi-asset Black-Scholes model; storage_cva_mixed: CVA + EPE of that mixed book against a CIR++ counterparty under EULER
    (tests/exposure_tests/cva_perfprmance_large_netting_set.py): the reference's golden with its injected draws, the oracle
    under native Philox."""
    res, sc = helpers.run_cuda(name, draws=draws)
    flat = helpers.flatten_results(res)
    if draws == "torch":
        gold = helpers.load_golden(name)
        want = {k: (np.array(v), np.array(gold["errors"][k])) for k, v in gold["values"].items()}
    else:
        out, _ = helpers.run_oracle(name, draws="philox")
        want = helpers.oracle_flat(out, res.get_netting_set_names(), res.get_metric_names())
    for key, (vals, errs) in want.items():
        scale = max(1.0, float(np.max(np.abs(vals))))
        # (regression proxies of the equity products pass through the reference's float32 chain: 1e-9 like the other goldens)
        helpers.assert_close(flat[key][0], vals, 1e-9, 1e-9 * scale, f"{name} {draws} {key}")
        if not key.split("|")[1].startswith("pfe"):
            helpers.assert_close(flat[key][1], errs, 1e-6, 1e-9 * scale, f"{name} {draws} {key} mc error")
