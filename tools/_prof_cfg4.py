import cProfile, pstats, importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
importlib.import_module("montecarlo-risk-engine_b200")
import torch, cases
ns = cases.Namespace()
def run():
    model, sets, metrics, tl = cases.bermudan_swaption(ns, n_ex=40)
    rm = ns.RiskMetrics(metrics, exposure_timeline=tl)
    sc = ns.SimulationController(sets, model, rm, 1 << 22, 1 << 22, 1, ns.SimulationScheme.EULER)
    res = sc.run_simulation(); torch.cuda.synchronize(); return sc
run(); t0=time.perf_counter(); sc=run(); print("wall %.1f ms"%((time.perf_counter()-t0)*1e3), sc.last_timings)
cProfile.run("run()", "/tmp/c4.prof")
pstats.Stats("/tmp/c4.prof").sort_stats("cumtime").print_stats(30)
