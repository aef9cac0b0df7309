"""Generates tests/golden/*.json by running the UNMODIFIED reference
(/root/reference/src on sys.path, PyTorch CPU) on the scenarios of tests/cases.py.

Run in the build container only (the reference is not available on the GPU box):
    python tests/golden/make_golden.py [case ...]
The reference draws its normals from torch.manual_seed(42|43) + torch.randn per sub-step;
the tests regenerate the same stream (oracle.engine.torch_reference_draws), so only the
outputs are stored here.
"""
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))  # tests/ for cases.py
sys.path.insert(0, "/root/reference/src")  # first: the reference has its own `helpers` package

# the reference's `helpers` is a namespace package (no __init__.py) and would lose against tests/helpers.py
_helpers = types.ModuleType("helpers")
_helpers.__path__ = ["/root/reference/src/helpers"]
sys.modules["helpers"] = _helpers

# matplotlib is imported by some reference modules but absent here
mpl = types.ModuleType("matplotlib")
mpl.pyplot = types.ModuleType("matplotlib.pyplot")
sys.modules.setdefault("matplotlib", mpl)
sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import cases  # noqa: E402


def _num(x):
    if x is None:
        return None
    return float(np.asarray(x))


def run_case(name):
    builder, bkw, rkw = cases.GOLDEN_CASES[name]
    ns = cases.Namespace()
    model, sets, metrics, tl = builder(ns, **bkw)
    rm = ns.RiskMetrics(metrics, exposure_timeline=tl) if tl is not None else ns.RiskMetrics(metrics)
    extra = dict(regression_function=ns.PolyomialRegression(degree=rkw["degree"])) if "degree" in rkw else {}
    sc = ns.SimulationController(sets, model, rm, rkw["n_main"], rkw["n_pre"], rkw["num_steps"],
                                 getattr(ns.SimulationScheme, rkw["scheme"]), rkw["differentiate"], **extra)
    if rkw.get("second_order"):
        sc.compute_higher_derivatives()
    res = sc.run_simulation()
    out = dict(case=name, builder=builder.__name__, builder_kwargs=bkw, run=rkw,
               torch=torch.__version__, sets=res.get_netting_set_names(), metrics=res.get_metric_names(),
               params=res.get_model_param_names(), simulation_timeline=[float(t) for t in sc.simulation_timeline],
               exposure_timeline=[float(t) for t in sc.exposure_timeline], values={}, errors={}, derivatives={})
    for s in out["sets"]:
        for m in out["metrics"]:
            key = f"{s}|{m}"
            out["values"][key] = [float(v) for v in res.get_results(s, m)]
            out["errors"][key] = [float(v) for v in res.get_mc_error(s, m)]
            if rkw["differentiate"]:
                d = res.get_derivatives(s, m)
                out["derivatives"][key] = [[_num(g) for g in ev] for ev in d]
            if rkw.get("second_order"):
                # [evaluation][parameter i][parameter j]; None = not connected in the reference's autograd graph
                h = res.second_derivatives[out["sets"].index(s)][out["metrics"].index(m)]
                out.setdefault("second_derivatives", {})[key] = [[[_num(x) for x in row] for row in ev] for ev in h]
    return out


if __name__ == "__main__":
    names = sys.argv[1:] or list(cases.GOLDEN_CASES)
    for name in names:
        data = run_case(name)
        with open(os.path.join(HERE, f"{name}.json"), "w") as f:
            json.dump(data, f, indent=1)
        print("wrote", name)
