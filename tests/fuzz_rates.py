"""Randomised parity sweep of the rates / credit family against the oracle (same Philox streams): random Vasicek models
(+ a correlated CIR++ counterparty), books of payer / receiver swaps, bonds and a Bermudan swaption, thresholded and
MPoR-collateralised netting sets, metric mixes (PV, CE, EPE, ENE, EEPE, PFE, CVA), EULER / ANALYTICAL, with / without Greeks.
    python tests/fuzz_rates.py [n_cases] [seed]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
importlib.import_module("montecarlo-risk-engine_b200")
import numpy as np  # noqa: E402

import cases  # noqa: E402
from oracle import risk  # noqa: E402


def build(rng, ns):
    hybrid = bool(rng.integers(0, 2))
    differentiate = bool(rng.integers(0, 3) == 0)
    vas = ns.VasicekModel(0., float(rng.uniform(0.01, 0.05)), float(rng.uniform(0.02, 0.06)), float(rng.uniform(0.01, 0.5)),
                          float(rng.uniform(0.005, 0.05)), asset_id="ir")
    model = vas
    if hybrid:
        cir = ns.CIRPPModel(0., "cp", cases.HAZARDS, 0.1, 0.01, 0.02, 0.0001, deterministic=bool(rng.integers(0, 4) == 0))
        model = ns.ModelConfig([vas, cir], inter_asset_correlation_matrix=np.array([float(rng.uniform(-0.8, 0.8))]))
    step = 0.25
    n_dates = int(rng.integers(3, 10))
    tl = np.arange(n_dates) * step
    horizon = step * (n_dates - 1)
    berm_used = [False]

    def product():
        kind = int(rng.integers(0, 5))
        T = step * int(rng.integers(2, n_dates + 3))
        if kind <= 2:
            return ns.InterestRateSwap(0.0, T, float(rng.uniform(0.5, 2.0)), float(rng.uniform(0.01, 0.05)),
                                       0.25 * int(rng.integers(1, 3)), 0.25, ns.IRSType.PAYER if rng.integers(0, 2) else ns.IRSType.RECEIVER,
                                       asset_id="ir")
        if kind == 3 or berm_used[0] or differentiate and hybrid:
            return ns.Bond(startdate=0.0, maturity=T, notional=float(rng.uniform(0.5, 2.0)), tenor=0.25 * int(rng.integers(1, 3)),
                           pays_notional=bool(rng.integers(0, 2)), fixed_rate=float(rng.uniform(0.0, 0.05)), asset_id="ir")
        berm_used[0] = True
        n_ex = int(rng.integers(2, 5))
        swap = ns.InterestRateSwap(0.0, step * (n_ex + 3), 1.0, 0.03, 0.25, 0.25, ns.IRSType.PAYER, asset_id="ir")
        return ns.BermudanOption(swap, [step * (i + 1) for i in range(n_ex)], 0.0, ns.OptionType.CALL, asset_id="ir")
    sets = []
    for s in range(int(rng.integers(1, 3))):
        kw = dict(name=f"set{s}", products=[product() for _ in range(int(rng.integers(1, 4)))])
        if hybrid:
            kw["counterparty_id"] = "cp"
        if rng.integers(0, 2):
            kw["threshold"] = float(rng.uniform(0.0, 0.02))
        if rng.integers(0, 2):
            kw["margin_period_of_risk"] = step * int(rng.integers(1, 3))
        sets.append(ns.NettingSet(**kw))
    pool = [ns.PVMetric(), ns.EPEMetric(), ns.ENEMetric(), ns.CEMetric(), ns.EEPEMetric(), ns.PFEMetric(0.9)]
    metrics = [pool[i] for i in sorted(set(int(x) for x in rng.integers(0, len(pool), 3)))]
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
        desc = (f"seed {seed} case {case}: {type(model).__name__} "
                f"sets={[([type(p).__name__[:4] for p in s.products], round(s.threshold, 4), s.margin_period_of_risk) for s in sets]} "
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
