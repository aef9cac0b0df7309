#!/usr/bin/env python
"""Cost of the sensitivity builds of the equity kernel on one book (tests/cases.py:bs_hessian, Euler, 2^20 paths): value
only, first order (Dual<3>), second order (Dual2<3>).  python tools/hessian_timing.py [--paths-log2 20]"""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
importlib.import_module("montecarlo-risk-engine_b200")
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--paths-log2", type=int, default=20)
    args = ap.parse_args()
    import torch
    import cases
    ns = cases.Namespace()
    n = 1 << args.paths_log2
    out = {"paths": n, "book": "bs_hessian (European, binary, barrier, Asian; 2 netting sets), EULER, 3 sub-steps per interval"}
    for label, diff, second in (("value", False, False), ("first_order", True, False), ("second_order", True, True)):
        best = None
        for _ in range(3):
            model, sets, metrics, _ = cases.bs_hessian(ns)
            sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics), n, 0, 3, ns.SimulationScheme.EULER, diff)
            if second:
                sc.compute_higher_derivatives()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = sc.run_simulation()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        steps = len(sc.simulation_timeline) * 3
        out[label] = {"seconds": best, "path_steps_per_s": n * steps / best}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
