"""Debug probe: cProfile of tools/large_book.py with every C-ABI entry point visible by name (ctypes calls are
otherwise charged to their Python caller) and kernels made synchronous (CUDA_LAUNCH_BLOCKING=1), so that GPU time
lands on the call that queued it."""
import cProfile
import os
import pstats
import sys
import types

os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import importlib
importlib.import_module("montecarlo-risk-engine_b200")
from mcre import binding as B

real = B.lib()


class Proxy:
    def __init__(self):
        self._cache = {}

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        if name not in self._cache:
            fn = getattr(real, name)

            def call(*a):
                # modules that set argtypes / restype lazily set them on this wrapper: forward them
                if "argtypes" in call.__dict__:
                    fn.argtypes = call.__dict__.pop("argtypes")
                if "restype" in call.__dict__:
                    fn.restype = call.__dict__.pop("restype")
                return fn(*a)
            call.__code__ = call.__code__.replace(co_name="ABI_" + name)
            self._cache[name] = call
        return self._cache[name]


B._lib = Proxy()
import large_book
sys.argv = ["large_book.py"] + sys.argv[1:]
pr = cProfile.Profile()
pr.enable()
large_book.main()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(40)
