#!/usr/bin/env python
"""Runs the BASELINE.json configurations at full size through the public API
(SimulationController.run_simulation) on one B200 and prints one JSON line per config:
wall time of the call, path-steps/s and a few result values.  Not the headline bench
(bench.py is); used to size and sanity-check the other configs.

    python tools/run_configs.py [1 2 2o 3 3g 4 5 5g] [--scale-log2 -2]   # --scale shrinks the path counts
"""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["1", "2", "2o", "4", "5", "5g"])
    ap.add_argument("--scale-log2", type=int, default=0)
    ap.add_argument("--repeats", type=int, default=2)
    args = ap.parse_args()
    importlib.import_module("montecarlo-risk-engine_b200")
    import torch
    import cases
    from mcre import binding as B
    ns = cases.Namespace()
    S = ns.SimulationScheme
    sc_ = args.scale_log2

    def n(log2):
        return 1 << max(log2 + sc_, 8)

    def cfg(name):
        if name == "1":
            model, sets, metrics, tl = cases.bs_european(ns, spot=100.0, strike=100.0, T=1.0)
            return dict(model=model, sets=sets, metrics=metrics, tl=None, n_main=100000, n_pre=0, steps=1,
                        scheme=S.ANALYTICAL, diff=True, sub=1, show=[("call", "pv")])
        if name in ("2", "2o"):
            mpor = 0.25 if name == "2" else 10 / 252
            model, sets, metrics, tl = cases.vasicek_irs_collateral(ns, mpor=mpor, n_dates=121, maturity=30.0)
            return dict(model=model, sets=sets, metrics=metrics, tl=tl, n_main=n(22), n_pre=n(22), steps=1, scheme=S.EULER,
                        diff=False, sub=120 if name == "2" else 240, show=[("irs_collateralized", "eepe"), ("irs_uncollateralized", "eepe")])
        if name in ("3", "3g"):
            # config 3 through the public API: value-only at full size; with all 8 first-order sensitivities of the CVA
            # (through the regression: tangent pre-simulation + differentiated normal equations) at 2^22 paths
            import bench
            model, sets, metrics, tl = bench.build_case(ns, 0.3)
            g = name == "3g"
            return dict(model=model, sets=sets, metrics=metrics, tl=tl, n_main=n(22 if g else 24), n_pre=n(18 if g else 20), steps=1,
                        scheme=S.EULER, diff=g, sub=240, show=[("irs", "cva[GM]")])
        if name == "4":
            model, sets, metrics, tl = cases.bermudan_swaption(ns, n_ex=40)
            return dict(model=model, sets=sets, metrics=metrics, tl=tl, n_main=n(22), n_pre=n(22), steps=1, scheme=S.EULER,
                        diff=False, sub=40, show=[("bermudan", "pv")])
        if name in ("5", "5g"):
            model, sets, metrics, tl = cases.heston_basket5(ns)
            diff = name == "5g"
            return dict(model=model, sets=sets, metrics=metrics, tl=None, n_main=n(24), n_pre=0, steps=21, scheme=S.QE,
                        diff=diff, sub=252, show=[("barrier", "pv"), ("asian", "pv")])
        if name in ("st", "stL"):
            # gas storage on the Schwartz two-factor model (tests/pytests/test_storage_s2f_pv.py: "storage2", 454 daily
            # decisions, 10 inventory states, cubic regression): "st" at the reference's own 4000 / 2000 paths (LAPACK
            # solve of the regressions), "stL" at 2^20 paths in both passes (device moments)
            model, sets, metrics, tl = cases.storage_s2f(ns, which="storage2")
            big = name == "stL"
            return dict(model=model, sets=sets, metrics=metrics, tl=None, n_main=n(20) if big else 2000,
                        n_pre=n(20) if big else 4000, steps=1, scheme=S.ANALYTICAL, diff=False, sub=454,
                        show=[("Storage", "pv")], extra=dict(regression_function=ns.PolyomialRegression(degree=3)))
        if name in ("hy", "hyL"):
            # three-model hybrid book of tests/pytests/test_cva_large_netting_set_aad_vs_fd.py (8 calls, 4 bonds, 40 swaps,
            # 30 exposure dates x 4 sub-steps): at its own 1024 paths and at 2^20
            model, sets, metrics, tl = cases.hybrid_cva(ns)
            big = name == "hyL"
            return dict(model=model, sets=sets, metrics=metrics, tl=tl, n_main=n(20) if big else 1024, n_pre=n(20) if big else 1024,
                        steps=4, scheme=S.EULER, diff=False, sub=116, show=[("large_cva_ns", "cva[large_counterparty]")])
        raise SystemExit(f"unknown config {name}")

    # algorithmic FP64 flop per path-step, derived like SURVEY 8(d) does for config 3 (add/mul 1, FMA 2,
    # exp / log 20, sincos 30, sqrt / div 8):
    #  1: one exact BS step + payoff + 3 tangents                                  ~ 64 + 30 + 3 * 30
    #  2: Box-Muller pair 64 (one normal used) + Vasicek Euler 7 + date work 110 (numeraire exp, LIBOR exp,
    #     2 coupons, 2 sets x (poly, threshold / MPoR, relu, sums))                = 181
    #  2o: as 2 but only every second sub-step is a metric date                     ~ 64 + 7 + 60
    #  4: 64 + 7 + exercise date: ~22 zero bonds x (exp 20 + 2) + payoff / decision 10 + exposure 30 + numeraire 20 = 615
    #  5: 5 assets x (2 normals 64 + uniform 4 + ~90 algebraic + 2 exp + 1 log + 4 sqrt + 4 div = 60 + 64) ~ 5 x 282 + basket 20
    #  st / stL: Box-Muller pair 64 + two-factor step 12 + exp 20 + decision (3 rate-curve interpolations ~30, 6 cubics 36,
    #     payoffs / arg-max / division ~20)                                          ~ 180
    #  hy / hyL: 3 normals 96 + correlation 9 + BS / Vasicek / credit steps ~25 + per exposure date (every 4th sub-step)
    #     numeraire exp 20 + 2 polynomials + netting 20                              ~ 150
    flop_model = {"st": 180.0, "stL": 180.0, "hy": 150.0, "hyL": 150.0, "1": 184.0, "2": 181.0, "2o": 131.0, "3": 225.0, "3g": 225.0, "4": 615.0, "5": 1430.0, "5g": 1430.0}
    import ctypes as C
    peak = C.c_double(0.0)
    from mcre import runtime as RT
    RT.compute_device()
    B.check(B.lib().mcre_dfma_peak(C.byref(peak), RT.stream_ptr()))
    for name in args.configs:
        best, res = None, None
        for rep in range(args.repeats):
            c = cfg(name)
            rm = ns.RiskMetrics(c["metrics"], exposure_timeline=c["tl"]) if c["tl"] is not None else ns.RiskMetrics(c["metrics"])
            torch.cuda.synchronize()
            l0 = B.launch_count()
            t0 = time.perf_counter()
            ctl = ns.SimulationController(c["sets"], c["model"], rm, c["n_main"], c["n_pre"], c["steps"], c["scheme"], c["diff"],
                                          **c.get("extra", {}))
            res = ctl.run_simulation()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            launches = B.launch_count() - l0
        out = {"config": name, "n_main": c["n_main"], "n_pre": c["n_pre"], "sub_steps": c["sub"], "differentiate": c["diff"],
               "seconds": best, "path_steps_per_s": c["n_main"] * c["sub"] / best, "launches": launches,
               "timings": {k: round(v, 4) for k, v in ctl.last_timings.items()},
               "max_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30,
               "flop_per_path_step_model": flop_model[name],
               "kernel_tflops_model": c["n_main"] * c["sub"] * flop_model[name] / max(ctl.last_timings["path_generation"], 1e-9) * 1e-12,
               "dfma_peak_tflops": peak.value}
        for s, m in c["show"]:
            out[f"{s}|{m}"] = [float(res.get_results(s, m)[0]), float(res.get_mc_error(s, m)[0])]
            if c["diff"]:
                out[f"{s}|{m}|d"] = [None if g is None else float(g) for g in res.get_derivatives(s, m)[0]][:7]
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
