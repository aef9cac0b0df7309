"""Pins the oracle (oracle/, numpy restatement) against outputs of the reference itself.

tests/golden/*.json were produced by tests/golden/make_golden.py, which runs the
UNMODIFIED reference on each scenario.  The oracle is run here on the same scenario
with the reference's own torch.randn stream regenerated from seeds 42 / 43, and must
agree to 1e-9 relative on every metric value and MC error (Greeks through the FP32
regression chain: 1e-5).  CPU only."""
import numpy as np
import pytest

import cases
import parity_helpers as helpers

VALUE_RTOL = 1e-9


@pytest.mark.parametrize("name", sorted(cases.GOLDEN_CASES))
def test_oracle_matches_reference_golden(name):
    gold = helpers.load_golden(name)
    out, (ns, model, sets, metrics, tl, rkw) = helpers.run_oracle(name, draws="torch")
    flat = helpers.oracle_flat(out, gold["sets"], gold["metrics"])
    for key, ref_vals in gold["values"].items():
        vals, errs = flat[key]
        scale = max(1.0, float(np.max(np.abs(ref_vals))))
        helpers.assert_close(vals, ref_vals, VALUE_RTOL, 1e-12 * scale, f"{name} {key} value")
        helpers.assert_close(errs, gold["errors"][key], 1e-7, 1e-11 * scale, f"{name} {key} mc error")
    if rkw["differentiate"]:
        for si, s in enumerate(gold["sets"]):
            for mi, m in enumerate(gold["metrics"]):
                ref = gold["derivatives"][f"{s}|{m}"]
                for ev, row in enumerate(ref):
                    got = out["grads"][si][mi][ev]
                    want = np.array([0.0 if g is None else g for g in row])
                    if got is None:
                        got = np.zeros_like(want)
                    exposure_metric = not m.startswith("pv")
                    rtol = 2e-5 if exposure_metric else 1e-8
                    helpers.assert_close(got, want, rtol, rtol * max(1.0, float(np.max(np.abs(want)))),
                                         f"{name} {s}|{m}[{ev}] derivatives")
    if rkw.get("second_order"):
        # the reference's double backward (controller.py:631-648) against the oracle's second-order forward mode
        helpers.assert_hessians(lambda si, mi, ev: out["hess"][si][mi][ev], gold["second_derivatives"], gold["sets"],
                                gold["metrics"], 1e-10, name)


def test_reference_known_answer_uncorrelated_cva():
    """The reference's own hard-coded constant (tests/pytests/test_cva.py:188-189): CVA of the
    uncorrelated Vasicek + CIR++ payer swap = 1.114576156484541 +- 0.0024446898428056294 at
    100k paths, 10 sub-steps, 100 exposure dates.  Reproducing it needs the full seeded
    pipeline (Euler stepping, FP32 regression chain, CVA integrand) to match."""
    ns = cases.Namespace()
    model, sets, metrics, tl = cases.wwr_cva(ns, rho=0.0, n_expo=100, extra_metrics=False)
    n = 100000
    from oracle import engine, risk
    n_sub, dim = helpers.n_substeps(model, sets, tl, metrics, 10)
    pre = engine.torch_reference_draws(42, n, n_sub, dim)
    out_pre_only = None
    main = engine.torch_reference_draws(43, n, n_sub, dim)
    out = risk.run(model, sets, metrics, tl, n, n, 10, "EULER", draws_pre=pre, draws_main=main)
    cva, err = out["results"][0][0][0]
    assert abs(cva - 1.114576156484541) < 1e-11
    assert abs(err - 0.0024446898428056294) < 1e-12


def test_reference_known_answer_american_option():
    """The reference's other hard-coded constant (tests/pytests/test_american_option.py:61): American call
    on Black-Scholes(100, r 5%, sigma 50%), 1000 exercise dates over 3y, 100k pre-simulation and 10k main
    paths, ANALYTICAL scheme: PV = 34.323036543142706 (reference threshold 1e-8).  Pins the oracle's
    Longstaff-Schwartz schedule (float32 cashflow roll, per-date lstsq, hard exercise indicator)."""
    from oracle import engine, risk
    ns = cases.Namespace()
    model = ns.BlackScholesModel(0.0, 100, 0.05, 0.5)
    prod = ns.AmericanOption(ns.Equity("id"), 3.0, 1000, 100.0, ns.OptionType.CALL)
    sets, metrics = [ns.NettingSet(name=prod.get_name(), products=[prod])], [ns.PVMetric()]
    n_sub, dim = helpers.n_substeps(model, sets, None, metrics, 1)
    pre = engine.torch_reference_draws(42, 100000, n_sub, dim)
    main = engine.torch_reference_draws(43, 10000, n_sub, dim)
    out = risk.run(model, sets, metrics, None, 10000, 100000, 1, "ANALYTICAL", draws_pre=pre, draws_main=main)
    assert abs(out["results"][0][0][0][0] - 34.323036543142706) < 1e-8
