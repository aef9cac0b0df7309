"""Schwartz-Smith two-factor commodity model around a forward curve
(reference: src/models/schwartz_two_factor.py:9-216)."""
from bisect import bisect_right
import math
from models.model import *
from mcre.dual import D, dexp, dsqrt


class SchwartzTwoFactorModel(Model):
    KIND = 5  # MCRE_MODEL_SCHWARTZ2F

    def __init__(self, calibration_date, curve_times, curve_values, rate, short_term_mean_reversion,
                 short_term_vol, long_term_drift, long_term_vol, rho, asset_id=None):
        super().__init__(calibration_date=calibration_date, asset_ids=[asset_id] if asset_id else None,
                         simulation_dim=2, state_dim=3)
        if len(curve_times) != len(curve_values):
            raise ValueError("curve_times and curve_values must have identical lengths.")
        if len(curve_times) < 2:
            raise ValueError("At least two curve points are required.")
        if any(v <= 0.0 for v in curve_values):
            raise ValueError("Curve values must be strictly positive.")
        self.curve_times = [float(t) for t in curve_times]
        self.curve_values = torch.tensor(curve_values, dtype=FLOAT, device=device)
        # parameter order: [rate, kappa_s, sigma_s, mu_l, sigma_l, rho]
        self.model_params = [torch.tensor(float(v), dtype=FLOAT, device=device)
                             for v in (rate, short_term_mean_reversion, short_term_vol,
                                       long_term_drift, long_term_vol, rho)]

    def get_rate(self):
        return self.model_params[0]

    def get_model_param_names(self):
        return ["rate", "short_term_mean_reversion", "short_term_vol", "long_term_drift",
                "long_term_vol", "rho"]

    def curve_value(self, t):
        """Piecewise-linear forward curve, flat outside (reference: :95-112)."""
        ts, vs = self.curve_times, self.curve_values.tolist()
        if t <= ts[0]:
            return vs[0]
        if t >= ts[-1]:
            return vs[-1]
        hi = bisect_right(ts, t)
        lo = hi - 1
        w = (t - ts[lo]) / (ts[hi] - ts[lo])
        return vs[lo] + (vs[hi] - vs[lo]) * w

    def intra_correlation(self, scheme, p):
        nt = p[0].t.shape[0]
        one = D(1.0, None, nt)
        # built in the reference's constructor, before requires_grad() (schwartz_two_factor.py:59-65): a constant of
        # the autograd graph; the exact scheme's covariance (below) reads rho at step time and does carry it
        rho = D(p[5].v, None, nt)
        return [[one, rho], [rho, one]]

    def unconnected_params(self, scheme):
        """Parameter indices the reference's autograd graph does not reach (returned as None)."""
        return {5} if scheme == SimulationScheme.EULER else set()

    def exact_covariance(self, p, dt):
        """2x2 covariance of (short, long) factor increments (reference: :124-145)."""
        kappa, ss, sl, rho = p[1], p[2], p[4], p[5]
        if abs(kappa.v) <= 1e-12:
            var_s = ss * ss * dt
        else:
            var_s = ss * ss * (1.0 - dexp(-2.0 * kappa * dt)) / (2.0 * kappa)
        var_l = sl * sl * dt
        cov = rho * dsqrt(var_s * var_l)
        return [[var_s, cov], [cov, var_l]]

    def rate_dual(self, p):
        return p[0]

    def numeraire(self, p, t):
        return dexp(p[0] * (t - self.t0()))

    def growth_factor(self, p, t1, t2):
        return dexp(p[0] * (t2 - t1))
