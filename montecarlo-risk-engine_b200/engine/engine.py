"""MonteCarloEngine compatibility class (reference: src/engine/engine.py:8-123).

`generate_paths()` returns the materialised tensor [N, T, D] like the reference; it is a
debug / notebook seam.  SimulationController does NOT use it: the
fused kernels never build paths.  Draws come from Philox with key 42 (pre-simulation) or
43 (main simulation), mirroring the reference's torch.manual_seed(42|43)."""
from common.packages import *
from common.enums import SimulationScheme


class MonteCarloEngine:
    def __init__(self, simulation_timeline, simulation_type, model, num_paths, num_steps,
                 is_pre_simulation=False):
        self.simulation_type = simulation_type
        self.model = model
        self.num_paths = num_paths
        self.num_steps = num_steps
        self.simulation_timeline = simulation_timeline
        self.seed = 42 if is_pre_simulation else 43
        self.injected_normals = None
        self.injected_uniforms = None

    def generate_paths(self):
        from mcre.paths import generate
        return generate(self.model, [float(t) for t in self.simulation_timeline], self.num_paths,
                        self.num_steps, self.simulation_type, self.seed,
                        inject_z=self.injected_normals, inject_u=self.injected_uniforms)
