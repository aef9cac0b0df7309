"""Hull-White one-factor short rate = Vasicek dynamics with a piecewise-constant,
time-dependent mean level theta(t).

The reference's hull_white.py is unimportable dead code, so
this is a build-defined extension: with a single theta bucket it coincides with
VasicekModel (which is the pinned oracle); with several buckets parity is
unpinned.  Only the EULER scheme supports a time-dependent theta."""
from models.vasicek import *


class HullWhiteModel(VasicekModel):
    def __init__(self, calibration_date, rate, mean, mean_reversion_speed, volatility,
                 asset_id=None, mean_times=None, mean_levels=None):
        super().__init__(calibration_date, rate, mean, mean_reversion_speed, volatility, asset_id)
        self.mean_times = [float(t) for t in (mean_times or [])]
        self.mean_levels = [float(v) for v in (mean_levels or [])]

    def mean_level(self, p, t):
        # theta(t) = mean_levels[i] on (mean_times[i-1], mean_times[i]]; base `mean` afterwards.
        for tt, lv in zip(self.mean_times, self.mean_levels):
            if t <= tt:
                return D(lv, None, p[2].t.shape[0])
        return p[2]
