"""Model base class (reference: src/models/model.py:5-141).

Host-side models are *descriptions*: parameters (0-d FP64 tensors in the
reference's documented order), asset ids and closed-form scalar helpers.  Path
stepping happens in csrc/models.cuh; the scalar helpers below are written
against ``mcre.dual.D`` so the plan compiler gets parameter tangents for free.
"""
from common.packages import *
from common.enums import SimulationScheme
from mcre.dual import D


class Model:
    #: kernel model-kind tag, see include/mcre.h (MCRE_MODEL_*)
    KIND = -1

    def __init__(self, calibration_date, simulation_dim=1, state_dim=1, asset_ids=None):
        self.calibration_date = torch.tensor([float(calibration_date)], dtype=FLOAT, device=device)
        self.asset_ids = asset_ids if asset_ids else [""]
        self.model_params: list[torch.Tensor] = []
        self.num_assets = len(self.asset_ids)
        self.simulation_dim = simulation_dim
        self.state_dim = state_dim
        self.perform_smoothing = False
        self.differentiate = False

    # -- parameter access (reference: model.py:30-36, 83-90) -----------------
    def get_model_params(self):
        return self.model_params

    def get_model_param_names(self):
        return [f"param_{i}" for i in range(len(self.model_params))]

    def requires_grad(self):
        """Switch on pathwise sensitivities (and fuzzy smoothing, like the
        reference does when autograd is enabled)."""
        self.perform_smoothing = True
        self.differentiate = True

    def unconnected_params(self, scheme):
        """Parameter indices that the reference's autograd graph does not reach for this scheme: their
        sensitivities come back as None (torch.autograd.grad(..., allow_unused=True), controller.py:618-624)."""
        return set()

    def param_values(self):
        return [float(p) for p in self.model_params]

    def dual_params(self, offset=0, n_total=0):
        """Parameters as dual numbers seeded at ``offset`` in a tangent space of
        size ``n_total`` (0 = no tangents)."""
        vals = self.param_values()
        if n_total == 0:
            return [D(v, None, 0) for v in vals]
        return [D.var(v, offset + i, n_total) for i, v in enumerate(vals)]

    def t0(self):
        # (cached: plan compilers call this per product record - tensor -> float costs microseconds)
        v = self.__dict__.get("_t0_cache")
        if v is None:
            v = self.__dict__["_t0_cache"] = float(self.calibration_date[0])
        return v

    # -- correlation description (reference: model.py:75-81) ------------------
    def intra_correlation(self, scheme, p):
        """simulation_dim x simulation_dim correlation among this model's own
        noise sources as nested lists of D."""
        n = self.simulation_dim
        nt = p[0].t.shape[0] if p else 0
        return [[D(1.0 if i == j else 0.0, None, nt) for j in range(n)] for i in range(n)]

    # -- compat shims: single-step debugging API ------------------------------
    def get_state(self, num_paths):
        from mcre.compat import initial_state
        return initial_state(self, num_paths)

    def simulate_time_step_euler(self, time1, time2, state, corr_randn):
        from mcre.compat import single_step
        return single_step(self, SimulationScheme.EULER, time1, time2, state, corr_randn)

    def simulate_time_step_analytically(self, time1, time2, state, corr_randn):
        from mcre.compat import single_step
        return single_step(self, SimulationScheme.ANALYTICAL, time1, time2, state, corr_randn)

    def simulate_time_step_qe(self, time1, time2, state, corr_randn):
        from mcre.compat import single_step
        return single_step(self, SimulationScheme.QE, time1, time2, state, corr_randn)
