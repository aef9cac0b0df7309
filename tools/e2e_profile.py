#!/usr/bin/env python
"""cProfile of one end-to-end SimulationController.run_simulation() call of the headline config
(host-side cost breakdown: lowering, plan upload, pre-simulation, solve, main pass, finishing)."""
import cProfile
import importlib
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
importlib.import_module("montecarlo-risk-engine_b200")
import torch  # noqa: E402
import bench  # noqa: E402
import cases  # noqa: E402

ns = cases.Namespace()


def run(i):
    model, sets, metrics, tl = bench.build_case(ns, float(bench.RHOS[i]))
    rm = ns.RiskMetrics(metrics, exposure_timeline=tl)
    sc = ns.SimulationController(sets, model, rm, 1 << 24, 1 << 20, 1, ns.SimulationScheme.EULER)
    res = sc.run_simulation()
    torch.cuda.synchronize()
    return sc


run(0); run(1)
t0 = time.perf_counter(); sc = run(2); print("e2e %.1f ms" % ((time.perf_counter() - t0) * 1e3), sc.last_timings)
cProfile.run("run(3)", "/tmp/e2e.prof")
pstats.Stats("/tmp/e2e.prof").sort_stats("cumtime").print_stats(28)
