"""Debug probe: per (set, metric, date) difference between the CUDA exposure Greeks and the oracle's duals."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib
importlib.import_module("montecarlo-risk-engine_b200")
import numpy as np
import helpers

name = sys.argv[1] if len(sys.argv) > 1 else "bs_proxy_greeks_mixed"
for n_main in (2048, 4096):
    gold = helpers.load_golden(name)
    res, sc = helpers.run_cuda(name, draws="philox", n_main=n_main)
    out, _ = helpers.run_oracle(name, draws="philox", n_main=n_main)
    for si, s in enumerate(gold["sets"]):
        for mi, m in enumerate(gold["metrics"]):
            for ev, want in enumerate(out["grads"][si][mi]):
                got = np.array([0.0 if g is None else float(g) for g in res.get_derivatives(s, m)[ev]])
                d = np.abs(got - want)
                if d.max() > 1e-8 * max(1.0, np.abs(want).max()):
                    print(n_main, s, m, ev, "diff", d, "want", want, "val diff", res.get_results(s, m)[ev] - out["results"][si][mi][ev][0])
