"""The C-ABI shared library loads and exports every symbol include/mcre.h declares.
No compute calls: runs without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mcre.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcre_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    import __graft_entry__ as g
    g.build()
    from mcre import binding
    lib = ctypes.CDLL(binding.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 15
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in include/mcre.h but not exported: {missing}"
    assert sorted(binding.SYMBOLS) == declared
    lib.mcre_abi_version.restype = ctypes.c_int
    assert lib.mcre_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    """The product path must fail loudly, not fall back, when no CUDA device is present."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cases
    from mcre.binding import McreError
    ns = cases.Namespace()
    model, sets, metrics, tl = cases.wwr_cva(ns)
    sc = ns.SimulationController(sets, model, ns.RiskMetrics(metrics, exposure_timeline=tl), 256, 256, 1,
                                 ns.SimulationScheme.EULER)
    with pytest.raises(McreError):
        sc.run_simulation()


def test_product_code_does_not_import_oracle():
    pkg = os.path.join(ROOT, "montecarlo-risk-engine_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders
