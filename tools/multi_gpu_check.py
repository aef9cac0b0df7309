#!/usr/bin/env python
"""GPU-count independence check: runs a set of scenarios through the public API with the
paths sharded over WORLD_SIZE GPUs (torchrun) and writes every result to --out as JSON with
the doubles' exact bit patterns.  Run once with 1 process and once under torchrun; the two
files must be identical (disjoint Philox streams keyed by global path id + fixed-order chunk
tree reduction).

    python tools/multi_gpu_check.py --out gpurun_out/mg1.json
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py --out gpurun_out/mg2.json
    python tools/multi_gpu_check.py --compare gpurun_out/mg1.json gpurun_out/mg2.json
"""
import argparse
import importlib
import json
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def bits(x):
    return struct.pack("<d", float(x)).hex()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out")
    ap.add_argument("--compare", nargs=2)
    ap.add_argument("--paths-log2", type=int, default=17)
    args = ap.parse_args()
    if args.compare:
        a, b = (json.load(open(f)) for f in args.compare)
        bad = [k for k in a if a[k] != b.get(k)]
        print(json.dumps({"scenarios": len(a), "values": sum(len(v) for v in a.values()), "mismatching_scenarios": bad}))
        sys.exit(1 if bad or set(a) != set(b) else 0)
    import torch
    import torch.distributed as dist
    importlib.import_module("montecarlo-risk-engine_b200")
    import cases
    import parity_helpers as helpers
    from mcre import runtime as RT
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = RT.compute_device()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 1 << args.paths_log2
    out = {}
    for name in ["cfg3_wwr_cva", "cfg2_irs_offgrid", "wwr_cva", "wwr_cva_greeks", "irs_collateral", "irs_collateral_offgrid", "bermudan_swaption",
                 "heston_path_dependent", "bs_basket_euler", "flexicall_exposure", "mixed_book_exposure",
                 "bs_exposure_greeks", "bs_proxy_greeks_mixed", "equity_cva", "equity_cva_exercise",
                 # round 2: three-model hybrid books (values and sensitivities), gas storage with the device-moments
                 # solver (pre-simulation sharded) and with the LAPACK solver (pre-simulation replicated), storage Greeks
                 "hybrid_cva_corr", "hybrid_collateral", "hybrid_collateral_greeks", "storage2_short_euler",
                 "storage2_short_euler:gelsy", "storage_s2f_greeks",
                 # sensitivities from per-path duals: PFE order statistics, equity + credit, books split over launches;
                 # pathwise Hessians
                 "bs_pfe_greeks", "hybrid_pfe_greeks", "equity_cva_det_greeks", "bs_split_book_greeks", "bs_hessian_multi",
                 # stochastic intensity: per-path default weights and their tangents
                 "equity_cva_greeks", "hybrid_stochastic_greeks"]:
        extra = {}
        if ":" in name:
            name, solver = name.split(":")
            extra["storage_solver"] = solver
        key = name + ("".join(f":{v}" for v in extra.values()))
        res, sc = helpers.run_cuda(name, draws="philox", n_main=n, n_pre=(n if cases.GOLDEN_CASES[name][2]["n_pre"] else 0), **extra)
        vals = []
        for s in res.get_netting_set_names():
            for m in res.get_metric_names():
                vals += [bits(v) for v in res.get_results(s, m)] + [bits(v) for v in res.get_mc_error(s, m)]
                if cases.GOLDEN_CASES[name][2]["differentiate"]:
                    for row in res.get_derivatives(s, m):
                        vals += [bits(0.0 if g is None else g) for g in row]
                if cases.GOLDEN_CASES[name][2].get("second_order"):
                    si, mi = res.get_netting_set_names().index(s), res.get_metric_names().index(m)
                    vals += [bits(0.0 if x is None else x) for ev in res.second_derivatives[si][mi] for row in ev for x in row]
        out[key] = vals
    # fewer chunks than ranks: ranks with an empty shard must finish with the same numbers (the pilot shift of the
    # shifted sums runs on every rank)
    for name, n_small in [("bs_european", 32), ("bs_exposure_greeks", 32), ("heston_path_dependent", 32), ("irs_collateral", 512),
                          ("wwr_cva_greeks", 512), ("bs_pfe_greeks", 32)]:
        res, sc = helpers.run_cuda(name, draws="philox", n_main=n_small, n_pre=(n_small if cases.GOLDEN_CASES[name][2]["n_pre"] else 0))
        vals = []
        for s in res.get_netting_set_names():
            for m in res.get_metric_names():
                vals += [bits(v) for v in res.get_results(s, m)] + [bits(v) for v in res.get_mc_error(s, m)]
                if cases.GOLDEN_CASES[name][2]["differentiate"]:
                    for row in res.get_derivatives(s, m):
                        vals += [bits(0.0 if g is None else g) for g in row]
        out[f"{name}@{n_small}"] = vals
    ns = cases.Namespace()
    model, sets, metrics, _ = cases.heston_basket5(ns)
    sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics), n, 0, 2, ns.SimulationScheme.QE, True)
    res = sc.run_simulation()
    out["heston_basket5"] = [bits(res.get_results(s, "pv")[0]) for s in res.get_netting_set_names()] + \
        [bits(g) for s in res.get_netting_set_names() for g in res.get_derivatives(s, "pv")[0]]
    # CVA of a book split over launches (default weights spilled by the first launch)
    model, sets, metrics, tl = cases.big_cva_book(ns)
    sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics, exposure_timeline=tl), n, n, 1, ns.SimulationScheme.EULER)
    res = sc.run_simulation()
    out["big_cva_book"] = [bits(v) for m in res.get_metric_names() for v in list(res.get_results("big", m)) + list(res.get_mc_error("big", m))]
    # every rank must hold the same results (SPMD contract)
    import hashlib
    digest = hashlib.sha256(json.dumps(out, sort_keys=True).encode()).hexdigest()
    digests = [digest]
    if world > 1:
        digests = [None] * world
        dist.all_gather_object(digests, digest)
    out["_all_ranks_agree"] = [str(len(set(digests)) == 1)]
    if RT.dist_info()[0] == 0:
        with open(args.out, "w") as f:
            json.dump(out, f)
        print(f"world {world}: wrote {sum(len(v) for v in out.values())} values to {args.out}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
