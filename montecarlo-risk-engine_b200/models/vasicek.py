"""Vasicek short rate with money-market numeraire
(reference: src/models/vasicek.py:5-156)."""
import math

from models.model import *
from mcre.dual import D, dexp, dsqrt, dval


class VasicekModel(Model):
    KIND = 3  # MCRE_MODEL_VASICEK

    def __init__(self, calibration_date, rate, mean, mean_reversion_speed, volatility, asset_id=None):
        super().__init__(calibration_date=calibration_date, state_dim=2, asset_ids=[asset_id])
        # parameter order: [rate, volatility, mean, mean_reversion_speed]
        self.model_params = [torch.tensor(float(v), dtype=FLOAT, device=device)
                             for v in (rate, volatility, mean, mean_reversion_speed)]

    def get_rate(self):
        return torch.stack([self.model_params[0]])

    def get_volatility(self):
        return torch.stack([self.model_params[1]])

    def get_mean(self):
        return torch.stack([self.model_params[2]])

    def get_mean_reversion_speed(self):
        return torch.stack([self.model_params[3]])

    def get_model_param_names(self):
        return ["rate", "volatility", "mean", "mean_reversion_speed"]

    def mean_level(self, p, t):
        """Long-run mean at time t (constant here; HullWhiteModel overrides)."""
        return p[2]

    # -- closed forms (reference: vasicek.py:114-128) -------------------------
    def bond_coefficients(self, p, t1, t2):
        """(alpha, B) with P(t1,t2;r) = exp(alpha - B r)."""
        sigma, theta, a = p[1], p[2], p[3]
        tau = t2 - t1
        if sigma.t.shape[0] == 0:
            # value-only plans: the same expression on plain floats (plans of exercise products evaluate this for
            # every zero bond of every exercise date - thousands of calls)
            sg, th, av = sigma.v, theta.v, a.v
            Bv = (1.0 - math.exp(-av * tau)) / av
            al = (th - sg * sg / (2.0 * av * av)) * (Bv - tau) - (sg * sg / (4.0 * av)) * Bv * Bv
            return D._val(al), D._val(Bv)
        B = (1.0 - dexp(-a * tau)) / a
        alpha = (theta - sigma * sigma / (2.0 * a * a)) * (B - tau) - (sigma * sigma / (4.0 * a)) * B * B
        return alpha, B

    def exact_step_constants(self, p, dt):
        """(decay, noise std) of the exact OU transition (reference: vasicek.py:52-86)."""
        sigma, a = p[1], p[3]
        decay = dexp(-a * dt)
        var = (sigma * sigma / (2.0 * a)) * (1.0 - decay * decay)
        return decay, dsqrt(var)

    def compute_bond_price(self, time1, time2, rate):
        """Zero-coupon bond price P(time1, time2; rate) as a tensor (host helper)."""
        p = self.dual_params()
        alpha, B = self.bond_coefficients(p, float(torch.as_tensor(time1).reshape(-1)[0]),
                                          float(torch.as_tensor(time2).reshape(-1)[0]))
        r = torch.as_tensor(rate, dtype=FLOAT)
        return torch.exp(alpha.v - B.v * r).reshape(-1) if r.ndim else torch.exp(alpha.v - B.v * r).reshape(1)
