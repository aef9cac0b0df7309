#!/usr/bin/env python
"""Builds tuning variants of the library (paths per thread x resident blocks per SM of the
CVA-only kernel, csrc/irc_cva.cu) into montecarlo-risk-engine_b200/variants/ so one GPU call can time them:

    python tools/tune_irc.py build 8x2 4x3 4x4 6x2           (here, no GPU)
    python tools/tune_irc.py run 8x2 4x3 ...                  (on the GPU box)

Only irc_cva.cu is recompiled per variant; the other objects come from the product build.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "montecarlo-risk-engine_b200")
sys.path.insert(0, PKG)
VAR = os.path.join(PKG, "variants")


def safe(v):
    return "".join(ch if ch.isalnum() else "_" for ch in v)


def lib_of(v):
    return os.path.join(VAR, f"libmcre_b200_{safe(v)}.so")


def main():
    mode, variants = sys.argv[1], sys.argv[2:]
    os.makedirs(VAR, exist_ok=True)
    if mode == "build":
        from mcre import build
        build.build()
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        objs = [os.path.join(PKG, "build", os.path.basename(s)[:-3] + ".o") for s in build.sources()]
        for v in variants:
            pp, _, rest = v.partition("x")
            minb, _, extra = rest.partition(",")          # e.g. 4x3,MCRE_CVA_PF=0
            flags = [f"-DMCRE_CVA_PP={pp}", f"-DMCRE_CVA_MINB={minb}"] + ([f"-D{e}" for e in extra.split(",") if e])
            obj = os.path.join(VAR, f"irc_cva_{safe(v)}.o")
            out = subprocess.run([nvcc, *build.NVCC_FLAGS, *flags, "-Xptxas", "-v", "-c", os.path.join(build.CSRC, "irc_cva.cu"), "-o", obj],
                                 capture_output=True, text=True)
            if out.returncode:
                sys.exit(out.stderr)
            regs = [l for l in out.stderr.splitlines() if "registers" in l]
            link = [o if not o.endswith("irc_cva.o") else obj for o in objs]
            subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_of(v), *link, "-lcudart"])
            print("built", lib_of(v), regs[-1].strip() if regs else "")
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
