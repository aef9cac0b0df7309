"""Black-Scholes single asset (reference: src/models/black_scholes.py:4-111)."""
from models.model import *
from mcre.dual import D, dexp


class BlackScholesModel(Model):
    KIND = 0  # MCRE_MODEL_BS

    def __init__(self, calibration_date, spot, rate, sigma, asset_id=None):
        super().__init__(calibration_date=calibration_date, asset_ids=[asset_id] if asset_id else None)
        # parameter order fixed by the reference: [spot, sigma, rate]
        self.model_params = [torch.tensor(float(v), dtype=FLOAT, device=device) for v in (spot, sigma, rate)]

    def get_spot(self):
        return torch.stack([self.model_params[0]])

    def get_volatility(self):
        return torch.stack([self.model_params[1]])

    def get_rate(self):
        return torch.stack([self.model_params[2]])

    def get_model_param_names(self):
        return ["spot", "volatility", "rate"]

    # scalar request formulas (reference: black_scholes.py:87-111)
    def rate_dual(self, p):
        return p[2]

    def numeraire(self, p, t):
        return dexp(p[2] * (t - self.t0()))

    def growth_factor(self, p, t1, t2):
        return dexp(p[2] * (t2 - t1))
