#!/usr/bin/env python
"""Builds tuning variants of the library (paths per thread x resident blocks per SM of the
IRC main kernel) into montecarlo-risk-engine_b200/variants/ so one GPU call can time them:

    python tools/tune_irc.py build 2x4 2x5 2x6 1x8 4x2           (here, no GPU)
    python tools/tune_irc.py run 2x4 2x5 ...                      (on the GPU box)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "montecarlo-risk-engine_b200")
sys.path.insert(0, PKG)
VAR = os.path.join(PKG, "variants")


def lib_of(v):
    return os.path.join(VAR, f"libmcre_b200_{v}.so")


def main():
    mode, variants = sys.argv[1], sys.argv[2:]
    os.makedirs(VAR, exist_ok=True)
    if mode == "build":
        from mcre import build
        for v in variants:
            pp, rest = v.split("x")
            prefetch = rest.endswith("p")
            rest = rest.rstrip("p")
            minb, _, unroll = rest.partition("u")
            flags = [f"-DMCRE_IRC_PP={pp}", f"-DMCRE_IRC_MINB={minb}"] + ([f"-DMCRE_IRC_UNROLL={unroll}"] if unroll else [])
            flags += ["-DMCRE_IRC_PREFETCH=1"] if prefetch else []
            build.build(force=True, extra_flags=flags, lib=lib_of(v), tag="_" + v)
            print("built", lib_of(v))
    else:
        for v in variants:
            env = dict(os.environ, MCRE_LIB_PATH=lib_of(v))
            out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "4", "--warmup", "3",
                                  "--no-cpu-baseline", "--no-e2e"], env=env, capture_output=True, text=True)
            try:
                d = json.loads(out.stdout.strip().splitlines()[-1])
                print(v, f"{d['value']:.4e} path-steps/s  {d['ms_per_step']:.2f} ms  frac {d['roofline']['frac']:.3f}", flush=True)
            except Exception:
                print(v, "FAILED", out.stdout[-300:], out.stderr[-600:], flush=True)


if __name__ == "__main__":
    main()
