"""Debug probe: tangents of the regression-proxy coefficients, CUDA (standardised basis -> raw) vs the oracle's duals."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib
importlib.import_module("montecarlo-risk-engine_b200")
import numpy as np
import helpers
from mcre import equity

name = sys.argv[1] if len(sys.argv) > 1 else "bs_proxy_greeks_mixed"
cap = {}
orig = equity.EquityBackend.presim_regression
def wrapped(self, products, dev):
    cap["be"], cap["products"] = self, products
    return orig(self, products, dev)
equity.EquityBackend.presim_regression = wrapped
res, sc = helpers.run_cuda(name, draws="philox")
out, _ = helpers.run_oracle(name, draws="philox")
be = cap["be"]
np.set_printoptions(linewidth=200, precision=6)
for k, p in enumerate(sc.products):
    if id(p) not in be.expo_dcoef:
        continue
    coef, basis = be.expo_coef[id(p)]
    dcoef = be.expo_dcoef[id(p)]          # [n_expo, nt, 3]
    a = be._asset_index(p.asset_ids[0])
    gmap = be.assets[a].gmap
    for e in range(coef.shape[0]):
        oc = out["expo_coeffs"][k][e]
        if not hasattr(oc, "t"):
            continue
        m, s = basis[e]
        def raw(c):
            return np.array([c[0] - c[1] * m * s + c[2] * m * m * s * s, c[1] * s - 2 * c[2] * m * s * s, c[2] * s * s])
        print(type(p).__name__, k, "date", e, "raw coef cuda", raw(coef[e]), "oracle", oc.v[1] if oc.v.shape[0] > 1 else oc.v[0])
        for j, g in enumerate(gmap):
            want = oc.t[g][-1]
            got = raw(dcoef[e, j])
            print("   param", g, "cuda", got, "oracle", want, "diff", got - want)
