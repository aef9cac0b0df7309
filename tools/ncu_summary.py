#!/usr/bin/env python
"""Summarises one kernel of an .ncu-rep (read with `ncu -i`): duration, registers, occupancy,
pipe utilisation, warp-state stalls and the SASS opcode mix per path-step.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep --path-steps $((2**22 * 240))
"""
import argparse
import collections
import csv
import io
import subprocess


def ncu(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--path-steps", type=float, default=None, help="path-steps of the profiled launch")
    ap.add_argument("--top", type=int, default=24)
    a = ap.parse_args()
    rows = list(csv.reader(io.StringIO(ncu(["-i", a.rep, "--page", "raw", "--csv"]))))
    d = dict(zip(rows[0], rows[-1]))
    want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__occupancy_limit_registers", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
            "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
            "smsp__issue_active.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]
    print("| metric | value |\n|---|---|")
    for k in want:
        if k in d:
            print(f"| {k} | {d[k]} |")
    for k in sorted(d):
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            v = float(d[k] or 0)
            if v >= 0.03:
                print(f"| {k.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', '')} | {v:.3f} |")
    if a.path_steps:
        inst = float(d["smsp__inst_executed.sum"])
        print(f"| warp instructions per 32 path-steps | {inst / a.path_steps * 32:.1f} |")
    src = list(csv.reader(io.StringIO(ncu(["-i", a.rep, "--page", "source", "--csv"]))))
    h0 = next(i for i, r in enumerate(src) if "Source" in r and "Address" in r)
    hdr = src[h0]
    ci, ce = hdr.index("Source"), hdr.index("Instructions Executed")
    ops = collections.Counter()
    for r in src[h0 + 1:]:
        try:
            n = float(r[ce])
        except (ValueError, IndexError):
            continue
        toks = r[ci].split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        ops[op.rstrip(";")] += n
    total = sum(ops.values())
    scale = (32.0 / a.path_steps) if a.path_steps else 1.0
    print("\n| opcode | warp instructions" + (" per 32 path-steps" if a.path_steps else "") + " |\n|---|---|")
    for op, n in ops.most_common(a.top):
        print(f"| {op} | {n * scale:.1f} |")
    fp64 = sum(n for op, n in ops.items() if op.startswith(("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")))
    print(f"| total | {total * scale:.1f} |\n| FP64 (DFMA+DMUL+DADD+DSETP) | {fp64 * scale:.1f} |")


if __name__ == "__main__":
    main()
