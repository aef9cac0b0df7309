"""Sub-step grid shared by all backends.

Restates the loop of MonteCarloEngine (reference: src/engine/engine.py:36-123) as a
table: ``num_steps`` sub-steps per interval between consecutive simulation dates,
intervals with dt <= 0 contribute no step, and the running time is *accumulated*
(t_prev + dt) rather than reset to the date, so the (time1, time2) pairs handed to the
models carry the same rounding as the reference's."""
from __future__ import annotations

from dataclasses import dataclass, field


@dataclass
class TimeGrid:
    dates: list            # simulation dates (floats)
    n_pre_dates: int       # leading dates reached without stepping (dt <= 0)
    t1: list = field(default_factory=list)        # accumulated start time of each sub-step
    t2: list = field(default_factory=list)        # t1 + dt as the reference forms it
    dt_nominal: list = field(default_factory=list)  # dt_total / num_steps (Cholesky cache key, model.py:45-64)
    dt: list = field(default_factory=list)        # time2 - time1 as the models see it
    date_after: list = field(default_factory=list)  # date index completed by the sub-step, or -1
    zero_dt_dates: list = field(default_factory=list)  # non-leading dates with dt <= 0 (state unchanged)
    alias: dict = field(default_factory=dict)         # zero-dt date index -> index of the date whose state it shares

    @property
    def n_sub(self):
        return len(self.dt)

    def index_map(self):
        """time -> index of the simulation date whose state is the state at that time.  Dates the running time has
        already reached or passed (dt <= 0: two dates one rounding error apart, e.g. 0.15 from np.linspace and
        0.15000000000000002 from repeated addition of 0.05) are not stepped to by the reference (engine.py:48-60: the
        state is unchanged) - their events join the date before them."""
        return {t: self.alias.get(i, i) for i, t in enumerate(self.dates)}


def build_time_grid(calibration_date, dates, num_steps):
    dates = [float(t) for t in dates]
    g = TimeGrid(dates=dates, n_pre_dates=0)
    t_prev = float(calibration_date)
    stepped = False
    for di, t_now in enumerate(dates):
        dt = (t_now - t_prev) / num_steps
        if dt > 0:
            stepped = True
            for k in range(num_steps):
                t_next = t_prev + dt
                g.t1.append(t_prev)
                g.t2.append(t_next)
                g.dt_nominal.append(dt)
                g.dt.append(t_next - t_prev)
                g.date_after.append(di if k == num_steps - 1 else -1)
                t_prev = t_next
        elif not stepped:
            g.n_pre_dates = di + 1
        else:
            g.zero_dt_dates.append(di)
            g.alias[di] = g.alias.get(di - 1, di - 1)
    return g
