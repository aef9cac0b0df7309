"""Black-Scholes, several correlated assets, one rate
(reference: src/models/black_scholes_multi.py:6-128)."""
import numpy as np
from models.model import *
from mcre.dual import D, dexp


class BlackScholesMulti(Model):
    KIND = 1  # MCRE_MODEL_BSM

    def __init__(self, calibration_date, rate, asset_ids, spots, volatilities, correlation_matrix):
        super().__init__(calibration_date=calibration_date, simulation_dim=len(asset_ids),
                         state_dim=len(spots), asset_ids=list(asset_ids))
        # parameter order: spots..., vols..., rate
        self.model_params = [torch.tensor(float(v), dtype=FLOAT, device=device)
                             for v in list(spots) + list(volatilities) + [rate]]
        self.correlation_matrix = torch.tensor(np.asarray(correlation_matrix, dtype=float), dtype=FLOAT, device=device)

    def get_spot(self):
        return torch.stack(self.model_params[:self.num_assets])

    def get_volatility(self):
        return torch.stack(self.model_params[self.num_assets:2 * self.num_assets])

    def get_rate(self):
        return self.model_params[2 * self.num_assets]

    def get_model_param_names(self):
        return [*(f"spot[{a}]" for a in self.asset_ids), *(f"volatility[{a}]" for a in self.asset_ids), "rate"]

    def intra_correlation(self, scheme, p):
        nt = p[0].t.shape[0]
        c = self.correlation_matrix.numpy()
        n = self.num_assets
        return [[D(float(c[i, j]), None, nt) for j in range(n)] for i in range(n)]

    def rate_dual(self, p):
        return p[2 * self.num_assets]

    def numeraire(self, p, t):
        return dexp(self.rate_dual(p) * (t - self.t0()))

    def growth_factor(self, p, t1, t2):
        return dexp(self.rate_dual(p) * (t2 - t1))
