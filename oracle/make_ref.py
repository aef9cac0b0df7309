#!/usr/bin/env python
"""Test / benchmark infrastructure: copies the UNMODIFIED reference sources (pure Python) from /root/reference/src
into oracle/_ref/src so that they travel to the GPU box, where /root/reference does not exist.  oracle/_ref/ is
git-ignored (reference sources never enter this repo's history) but not gpurun-ignored.

    python oracle/make_ref.py            # in the build container; __graft_entry__.build() calls it too

Used only by `bench.py --impl reference` and the `cpu_baseline` leg of `bench.py` (the reference's own
SimulationController timed on the host cores) - never by the product path.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src"
DST = os.path.join(HERE, "_ref", "src")


def make(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"{SRC} not present: keeping {DST} as it is", file=sys.stderr)
        return os.path.isdir(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    digest = hashlib.sha256()
    files = []
    for root, _, names in sorted(os.walk(DST)):
        for n in sorted(names):
            p = os.path.join(root, n)
            with open(p, "rb") as f:
                b = f.read()
            digest.update(b)
            files.append(os.path.relpath(p, DST))
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": files, "sha256": digest.hexdigest()}, f, indent=1)
    if verbose:
        print(f"copied {len(files)} files to {DST} (sha256 {digest.hexdigest()[:16]})")
    return True


def available():
    return os.path.isfile(os.path.join(DST, "controller", "controller.py"))


if __name__ == "__main__":
    make()
