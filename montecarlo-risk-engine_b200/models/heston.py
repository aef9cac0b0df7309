"""Heston stochastic volatility, Euler and Andersen-QE with fuzzy branching
(reference: src/models/heston.py:20-280)."""
from models.model import *
from mcre.dual import D, dexp


class HestonModel(Model):
    KIND = 2  # MCRE_MODEL_HESTON

    def __init__(self, calibration_date, spot, rate, sigma, rho, kappa, theta, v0, asset_id=None):
        super().__init__(calibration_date=calibration_date, asset_ids=[asset_id] if asset_id else None,
                         simulation_dim=2, state_dim=2)
        # parameter order: [spot, sigma(vol-of-vol), rate, rho, kappa, theta, v0]
        self.model_params = [torch.tensor(float(v), dtype=FLOAT, device=device)
                             for v in (spot, sigma, rate, rho, kappa, theta, v0)]

    def get_spot(self):
        return torch.stack([self.model_params[0]])

    def get_volatility(self):
        return torch.stack([self.model_params[1]])

    def get_rate(self):
        return torch.stack([self.model_params[2]])

    def get_rho(self):
        return torch.stack([self.model_params[3]])

    def get_kappa(self):
        return torch.stack([self.model_params[4]])

    def get_theta(self):
        return torch.stack([self.model_params[5]])

    def get_initial_variance(self):
        return torch.stack([self.model_params[6]])

    def get_model_param_names(self):
        return ["spot", "volatility", "rate", "rho", "kappa", "theta", "initial_variance"]

    def intra_correlation(self, scheme, p):
        # QE draws independent normals; Euler correlates spot/variance noise by rho
        nt = p[0].t.shape[0]
        one, zero = D(1.0, None, nt), D(0.0, None, nt)
        if scheme == SimulationScheme.QE:
            return [[one, zero], [zero, one]]
        # the reference builds this matrix in the constructor, before requires_grad() (heston.py:53-58): a constant of
        # the autograd graph, so rho gets no sensitivity under EULER (its derivative comes back as None)
        rho = D(p[3].v, None, nt)
        return [[one, rho], [rho, one]]

    def unconnected_params(self, scheme):
        """Parameter indices the reference's autograd graph does not reach (returned as None)."""
        return set() if scheme == SimulationScheme.QE else {3}

    def rate_dual(self, p):
        return p[2]

    def numeraire(self, p, t):
        return dexp(p[2] * (t - self.t0()))

    def growth_factor(self, p, t1, t2):
        return dexp(p[2] * (t2 - t1))
