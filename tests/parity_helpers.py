"""Helpers shared by the tests: build a scenario with this repo's host objects, run the
oracle on it, run the CUDA controller on it, compare nested results."""
import json
import math
import os

import numpy as np

import cases

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN_DIR, f"{name}.json")) as f:
        return json.load(f)


def build(name, **override):
    builder, bkw, rkw = cases.GOLDEN_CASES[name]
    ns = cases.Namespace()
    model, sets, metrics, tl = builder(ns, **bkw)
    rkw = dict(rkw, **override)
    return ns, model, sets, metrics, tl, rkw


def n_substeps(model, sets, tl, metrics, num_steps):
    """Number of sub-steps / noise dimension of a run, from the oracle's own timeline logic."""
    from oracle import engine, models, risk
    products = [p for s in sets for p in s.products]
    expo = set(float(t) for t in (tl if tl is not None else []))
    if any(m.metric_type.name != "PV" for m in metrics):
        for s in sets:
            if s.margin_period_of_risk is not None:
                expo |= {float(t - s.margin_period_of_risk) for t in np.asarray(tl) if t - s.margin_period_of_risk >= 0}
    sim = sorted({t for p in products for t in risk.modeling_timeline(p)} | expo)
    return engine.count_substeps(models.t0_of(model), sim, num_steps), models.noise_dim(model)


def reference_draws(model, sets, tl, metrics, rkw):
    """The reference's torch.randn stream for this run (seed 42 pre, 43 main)."""
    from oracle import engine
    n_sub, dim = n_substeps(model, sets, tl, metrics, rkw["num_steps"])
    qe = rkw["scheme"] == "QE"
    pre = engine.torch_reference_draws(42, rkw["n_pre"], n_sub, dim, qe) if rkw["n_pre"] > 0 else None
    main = engine.torch_reference_draws(43, rkw["n_main"], n_sub, dim, qe)
    return pre, main


def run_oracle(name, draws="torch", **override):
    from oracle import risk
    ns, model, sets, metrics, tl, rkw = build(name, **override)
    pre = main = None
    if draws == "torch":
        pre, main = reference_draws(model, sets, tl, metrics, rkw)
    out = risk.run(model, sets, metrics, tl, rkw["n_main"], rkw["n_pre"], rkw["num_steps"], rkw["scheme"],
                   differentiate=rkw["differentiate"], draws_pre=pre, draws_main=main,
                   degree=rkw.get("degree", 2) + 1, storage_solver=rkw.get("storage_solver", "gelsy"),
                   second_order=rkw.get("second_order", False))
    return out, (ns, model, sets, metrics, tl, rkw)


def run_cuda(name, draws="torch", **override):
    """Run the scenario through this repo's SimulationController (CUDA kernels)."""
    ns, model, sets, metrics, tl, rkw = build(name, **override)
    rm = ns.RiskMetrics(metrics, exposure_timeline=tl) if tl is not None else ns.RiskMetrics(metrics)
    extra = dict(regression_function=ns.PolyomialRegression(degree=rkw["degree"])) if "degree" in rkw else {}
    sc = ns.SimulationController(sets, model, rm, rkw["n_main"], rkw["n_pre"], rkw["num_steps"],
                                 getattr(ns.SimulationScheme, rkw["scheme"]), rkw["differentiate"], **extra)
    if "storage_solver" in rkw:
        sc.storage_regression = "lapack" if rkw["storage_solver"] == "gelsy" else rkw["storage_solver"]
    if rkw.get("second_order"):
        sc.compute_higher_derivatives()
    if draws == "torch":
        pre, main = reference_draws(model, sets, tl, metrics, rkw)
        sc.inject_normals(pre=None if pre is None else pre.z, main=main.z)
        if main.u is not None:
            sc.injected_uniforms = {"pre": None if pre is None else pre.u, "main": main.u}
    res = sc.run_simulation()
    return res, sc


def flatten_results(res):
    """SimulationResults -> {"set|metric": (values, errors)}."""
    out = {}
    for s in res.get_netting_set_names():
        for m in res.get_metric_names():
            out[f"{s}|{m}"] = (np.asarray(res.get_results(s, m), dtype=float),
                               np.asarray(res.get_mc_error(s, m), dtype=float))
    return out


def oracle_flat(out, set_names, metric_names):
    flat = {}
    for si, s in enumerate(set_names):
        for mi, m in enumerate(metric_names):
            vals = out["results"][si][mi]
            flat[f"{s}|{m}"] = (np.array([v for v, _ in vals]), np.array([e for _, e in vals]))
    return flat


def assert_close(a, b, rtol, atol, what):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    both_nan = np.isnan(a) & np.isnan(b)
    err = np.abs(a - b)
    tol = atol + rtol * np.maximum(np.abs(a), np.abs(b))
    bad = ~(both_nan | (err <= tol))
    assert not bad.any(), f"{what}: max abs diff {np.nanmax(err):.3e} (rtol {rtol}, atol {atol}); a={a[bad][:3]} b={b[bad][:3]}"


def assert_gradients(res, gold_like, params, sets, metrics, rtol_of, what):
    """Every metric's gradient rows against `gold_like[f"{set}|{metric}"]` (rows per evaluation, None = not connected)."""
    for s_ in sets:
        for m in metrics:
            rows, got = gold_like[f"{s_}|{m}"], res.get_derivatives(s_, m)
            for ev, row in enumerate(rows):
                if row is None:
                    continue
                scale = max([1.0] + [abs(w) for w in row if w is not None])
                for pname, g, w in zip(params, got[ev], row):
                    if w is None:
                        assert g is None, f"{what} {s_}|{m}[{ev}] d/d{pname}: expected None, got {g}"
                    else:
                        assert g is not None, f"{what} {s_}|{m}[{ev}] d/d{pname} is None"
                        assert abs(float(g) - w) <= rtol_of(m) * scale, f"{what} {s_}|{m}[{ev}] d/d{pname}: {float(g)} vs {w}"


def assert_hessians(hess_of, gold_like, sets, metrics, rtol, what):
    """`hess_of(set index, metric index, evaluation)` -> [P][P] (entries may be None) against
    `gold_like[f"{set}|{metric}"][evaluation][i][j]`.  The reference returns None or 0.0 for structurally zero entries
    depending on its autograd graph's connectivity; both compare as zero."""
    for si, s_ in enumerate(sets):
        for mi, m in enumerate(metrics):
            for ev, want in enumerate(gold_like[f"{s_}|{m}"]):
                got = hess_of(si, mi, ev)
                scale = max([1.0] + [abs(x) for row in want for x in row if x is not None])
                for i, row in enumerate(want):
                    for j, w in enumerate(row):
                        g = got[i][j]
                        g, w = (0.0 if g is None else float(g)), (0.0 if w is None else float(w))
                        assert abs(g - w) <= rtol * scale, f"{what} {s_}|{m}[{ev}] d2/d{i}d{j}: {g} vs {w}"
