"""Randomised parity sweep of the equity family's exposure features against the oracle (same Philox streams): random
Black-Scholes models (1-3 assets), books of European / binary / Asian / barrier options, thresholded and
MPoR-collateralised netting sets, metric mixes, with / without a CIR++ counterparty (CVA), with / without Greeks.
    python tests/fuzz_equity.py [n_cases] [seed]
Prints one line per case and exits non-zero on a mismatch; tests/test_fuzz_equity_gpu.py runs a fixed set of seeds."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
importlib.import_module("montecarlo-risk-engine_b200")
import numpy as np  # noqa: E402

import cases  # noqa: E402
import parity_helpers as helpers  # noqa: E402
from oracle import risk  # noqa: E402


def build(rng, ns):
    A = int(rng.integers(1, 4))
    ids = [f"a{i}" for i in range(A)]
    hybrid = bool(rng.integers(0, 2))
    differentiate = (not hybrid) and bool(rng.integers(0, 2))
    if A == 1:
        market = ns.BlackScholesModel(0.0, float(rng.uniform(80, 120)), float(rng.uniform(0.0, 0.06)), float(rng.uniform(0.1, 0.4)),
                                      asset_id=ids[0])
    else:
        c = np.full((A, A), float(rng.uniform(-0.2, 0.6))) + 0.0
        np.fill_diagonal(c, 1.0)
        market = ns.BlackScholesMulti(calibration_date=0.0, rate=float(rng.uniform(0.0, 0.06)), asset_ids=ids,
                                      spots=[float(x) for x in rng.uniform(80, 120, A)],
                                      volatilities=[float(x) for x in rng.uniform(0.1, 0.4, A)], correlation_matrix=c)
    model = market
    if hybrid:
        credit = ns.CIRPPModel(calibration_date=0.0, asset_id="cp", hazard_rates=cases.HAZARDS, kappa=0.10, theta=0.01,
                               volatility=0.02, y0=0.0001, deterministic=bool(rng.integers(0, 2)))
        rho = float(rng.uniform(-0.5, 0.5))
        inter = [np.array([rho])] if A == 1 else [np.full((A, 1), rho, dtype=float)]
        model = ns.ModelConfig(models=[market, credit], inter_asset_correlation_matrix=inter)
    step = 0.25
    n_dates = int(rng.integers(3, 8))
    tl = np.arange(n_dates) * step

    def product():
        a = ids[int(rng.integers(0, A))]
        T = step * int(rng.integers(1, n_dates + 2))
        K = float(rng.uniform(85, 115))
        ot = ns.OptionType.CALL if rng.integers(0, 2) else ns.OptionType.PUT
        kind = int(rng.integers(0, 4))
        if kind == 0:
            return ns.EuropeanOption(ns.Equity(a), T, K, ot, asset_id=a)
        if kind == 1:
            return ns.BinaryOption(T, K, float(rng.uniform(5, 15)), ot, asset_id=a)
        if kind == 2:
            return ns.AsianOption(0.0, T, K, int(rng.integers(2, 6)), ot, asset_id=a)
        return ns.BarrierOption(startdate=0.0, maturity=T, strike=K, num_observation_timepoints=int(rng.integers(2, 6)),
                                option_type=ot, barrier1=float(rng.uniform(120, 150)),
                                barrier_option_type1=ns.BarrierOptionType.UPANDOUT, asset_id=a)
    sets = []
    for s in range(int(rng.integers(1, 3))):
        n_prod = int(rng.integers(1, 3 if differentiate else 5))      # tangent builds track 2 path-dependent products
        kw = dict(name=f"set{s}", products=[product() for _ in range(n_prod)])
        if hybrid:
            kw["counterparty_id"] = "cp"
        if rng.integers(0, 2):
            kw["threshold"] = float(rng.uniform(0.0, 5.0))
        if rng.integers(0, 2):
            kw["margin_period_of_risk"] = step * int(rng.integers(1, 3))
        sets.append(ns.NettingSet(**kw))
    pool = [ns.PVMetric(), ns.EPEMetric(), ns.ENEMetric(), ns.CEMetric(), ns.EEPEMetric()]
    metrics = [pool[i] for i in sorted(set(int(x) for x in rng.integers(0, len(pool), 3)))]
    if not differentiate and rng.integers(0, 2):
        metrics.append(ns.PFEMetric(0.9))
    if hybrid:
        metrics.append(ns.CVAMetric("cp", 0.4))
    if all(m.metric_type.name == "PV" for m in metrics):
        metrics.append(ns.EPEMetric())
    scheme = "EULER" if (hybrid or rng.integers(0, 2)) else "ANALYTICAL"
    return model, sets, metrics, tl, scheme, differentiate, int(rng.integers(1, 3))


def run_cases(n_cases, seed, n=512, log=print):
    """-> number of mismatching cases (values 1e-7 relative, Greeks 1e-5)."""
    rng = np.random.default_rng(seed)
    ns = cases.Namespace()
    bad = 0
    for case in range(n_cases):
        model, sets, metrics, tl, scheme, differentiate, num_steps = build(rng, ns)
        desc = (f"seed {seed} case {case}: {type(model).__name__} sets={[(len(s.products), s.threshold, s.margin_period_of_risk) for s in sets]} "
                f"metrics={[m.get_name() for m in metrics]} {scheme} steps={num_steps} greeks={differentiate}")
        try:
            sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics, exposure_timeline=tl), n, n, num_steps,
                                         getattr(ns.SimulationScheme, scheme), differentiate)
            res = sc.run_simulation()
        except NotImplementedError as e:
            log(desc + " -> not supported: " + str(e)[:80])
            continue
        out = risk.run(model, sets, metrics, tl, n, n, num_steps, scheme, differentiate=differentiate)
        worst = 0.0
        for si, s in enumerate(res.get_netting_set_names()):
            for mi, m in enumerate(res.get_metric_names()):
                got = np.asarray(res.get_results(s, m), dtype=float)
                want = np.array([v for v, _ in out["results"][si][mi]])
                scale = max(1.0, float(np.max(np.abs(want))))
                worst = max(worst, float(np.max(np.abs(got - want))) / scale)
                if differentiate:
                    for ev, w in enumerate(out["grads"][si][mi]):
                        if w is None:
                            continue
                        g = np.array([0.0 if x is None else float(x) for x in res.get_derivatives(s, m)[ev]])
                        worst = max(worst, 0.01 * float(np.max(np.abs(g - w))) / max(1.0, float(np.max(np.abs(w)))))
        ok = worst < 1e-7
        bad += not ok
        log(desc + f" -> worst rel diff {worst:.2e} " + ("OK" if ok else "MISMATCH"))
    return bad


if __name__ == "__main__":
    sys.exit(1 if run_cases(int(sys.argv[1]) if len(sys.argv) > 1 else 20, int(sys.argv[2]) if len(sys.argv) > 2 else 1) else 0)
