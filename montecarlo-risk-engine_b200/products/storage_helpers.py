"""Contract data of a gas storage: inventory limits, injection / withdrawal rate curves, variable costs
(reference: src/products/storage_helpers.py:48-437).

The reference keeps dataclass lists and torch helpers; here everything is plain floats in sorted tuples, and
the per-date view the kernels need is produced by `Storage.lower()` (products/storage.py) as flat tables.  The
arithmetic of the reachable-inventory envelope (`optimize_volume_constraints`) follows the reference operation by
operation, because the envelope bounds define the inventory grid of the dynamic programme
(storage_helpers.py:290-437): a bound that differs in the last bit would move every state of that date."""
from __future__ import annotations

import bisect
import math

DATE_TOL = 1e-12
VOLUME_TOL = 1e-12


def _same(a, b, tol):
    return math.isclose(a, b, abs_tol=tol)


class _Window:
    """Half-open date window [start, end) carrying inventory limits; a degenerate window matches its date only
    (storage_helpers.py:50-54)."""
    __slots__ = ("start_date", "end_date", "vmin", "vmax", "penalty")

    def __init__(self, start_date, end_date, vmin, vmax, penalty=0.0):
        self.start_date, self.end_date = start_date, end_date
        self.vmin, self.vmax, self.penalty = vmin, vmax, penalty

    def contains(self, date):
        return StorageConfig._date_in_window(self.start_date, self.end_date, date)

    def __repr__(self):
        return f"_Window({self.start_date}, {self.end_date}, vmin={self.vmin}, vmax={self.vmax})"


class _Curve:
    """Rate curve (inventory level -> daily rate) valid on a date window: sorted (level, rate) knots."""
    __slots__ = ("start_date", "end_date", "knots")

    def __init__(self, start_date, end_date):
        self.start_date, self.end_date, self.knots = start_date, end_date, []

    def contains(self, date):
        return StorageConfig._date_in_window(self.start_date, self.end_date, date)


class _Knot:
    __slots__ = ("point", "rate")

    def __init__(self, point, rate):
        self.point, self.rate = point, rate


def curve_rate(level, knots):
    """Piecewise-linear rate at a scalar inventory level, flat beyond the end knots
    (storage_helpers.py:68-94; the scalar twin of the tensor rule the kernels implement)."""
    if not knots:
        raise ValueError("Flexibility slice is empty.")
    if len(knots) == 1:
        return knots[0].rate
    xs = [k.point for k in knots]
    if level <= xs[0]:
        return knots[0].rate
    if level >= xs[-1]:
        return knots[-1].rate
    hi = bisect.bisect_right(xs, level)
    a, b = knots[hi - 1], knots[hi]
    if _same(a.point, b.point, VOLUME_TOL):
        return b.rate
    return a.rate + (level - a.point) / (b.point - a.point) * (b.rate - a.rate)


class StorageConfig:
    def __init__(self):
        self.initial_volume_constraints = []
        self.volume_constraints = []          # the envelope, filled by optimize_volume_constraints
        self.injection_flexibility = []
        self.withdrawal_flexibility = []
        self.injection_costs = []             # sorted (date, cost)
        self.withdrawal_costs = []

    # -- static helpers of the reference's API --------------------------------------------------------------------
    @staticmethod
    def _date_in_window(start_date, end_date, date):
        if _same(start_date, end_date, DATE_TOL):
            return _same(start_date, date, DATE_TOL)
        return start_date - DATE_TOL <= date < end_date - DATE_TOL

    @staticmethod
    def grid_step(vmin, vmax, num_states):
        """Inventory per state index (storage_helpers.py:56-60)."""
        if num_states <= 1 or _same(vmin, vmax, VOLUME_TOL):
            return 0.0
        return (vmax - vmin) / (num_states - 1.0)

    @staticmethod
    def state_scale(vmin, vmax, num_states):
        """State index per unit of inventory (storage_helpers.py:62-66)."""
        if num_states <= 1 or _same(vmin, vmax, VOLUME_TOL):
            return 0.0
        return (num_states - 1.0) / (vmax - vmin)

    @staticmethod
    def _interpolate_rate(point, rate_points):
        return curve_rate(point, rate_points)

    @staticmethod
    def interpolate_rate_tensor(point, rate_points):
        """Tensor form used by host-side checks (storage_helpers.py:96-127); the simulation evaluates the same
        rule in csrc/storage.cu:curve_rate_dev."""
        import torch
        if not rate_points:
            raise ValueError("Flexibility slice is empty.")
        if len(rate_points) == 1:
            return torch.full_like(point, rate_points[0].rate)
        xp = point.new_tensor([k.point for k in rate_points])
        fp = point.new_tensor([k.rate for k in rate_points])
        left = torch.clamp(torch.bucketize(point, xp) - 1, min=0, max=len(rate_points) - 2)
        x0, x1, y0, y1 = xp[left], xp[left + 1], fp[left], fp[left + 1]
        w = torch.where(torch.isclose(x0, x1), torch.zeros_like(point), (point - x0) / (x1 - x0))
        out = y0 + w * (y1 - y0)
        out = torch.where(point <= xp[0], fp[0], out)
        return torch.where(point >= xp[-1], fp[-1], out)

    # -- inventory limits -----------------------------------------------------------------------------------------
    def add_volume_constraint(self, start_date, end_date, vmin, vmax, penalty=0.0):
        self.initial_volume_constraints.append(_Window(start_date, end_date, vmin, vmax, penalty))
        self.initial_volume_constraints.sort(key=lambda w: w.start_date)

    @staticmethod
    def _window_at(date, windows):
        for w in windows:
            if w.contains(date):
                return w
        if not windows:
            raise ValueError("No volume constraints configured.")
        return windows[-1]

    def get_initial_volume_constraint(self, date):
        return self._window_at(date, self.initial_volume_constraints)

    def get_volume_constraint(self, date):
        return self._window_at(date, self.volume_constraints or self.initial_volume_constraints)

    # -- rate curves ----------------------------------------------------------------------------------------------
    @staticmethod
    def _add_knot(curves, start_date, end_date, point, rate):
        for c in curves:
            if _same(c.start_date, start_date, DATE_TOL) and _same(c.end_date, end_date, DATE_TOL):
                c.knots.append(_Knot(point, rate))
                c.knots.sort(key=lambda k: k.point)
                return
        c = _Curve(start_date, end_date)
        c.knots.append(_Knot(point, rate))
        curves.append(c)
        curves.sort(key=lambda k: k.start_date)

    @staticmethod
    def _curve_at(date, curves):
        for c in curves:
            if c.contains(date):
                return c.knots
        if not curves:
            raise ValueError("No flexibility slice configured.")
        return curves[-1].knots

    def add_injection_flexibility(self, start_date, end_date, point, rate):
        self._add_knot(self.injection_flexibility, start_date, end_date, point, rate)

    def add_withdrawal_flexibility(self, start_date, end_date, point, rate):
        self._add_knot(self.withdrawal_flexibility, start_date, end_date, point, rate)

    def get_injection_flexibility_slice(self, date):
        return self._curve_at(date, self.injection_flexibility)

    def get_withdrawal_flexibility_slice(self, date):
        return self._curve_at(date, self.withdrawal_flexibility)

    def get_injection_flexibility_rate(self, date, point):
        return curve_rate(point, self.get_injection_flexibility_slice(date))

    def get_withdrawal_flexibility_rate(self, date, point):
        return curve_rate(point, self.get_withdrawal_flexibility_slice(date))

    # -- variable costs: piecewise constant from their date on (storage_helpers.py:243-262) ---------------------------
    @staticmethod
    def _cost_at(date, costs):
        if not costs:
            raise ValueError("No variable costs configured.")
        i = bisect.bisect_left([d for d, _ in costs], date)
        if i == len(costs):
            return costs[-1][1]
        if i == 0 or _same(costs[i][0], date, DATE_TOL):
            return costs[i][1]
        return costs[i - 1][1]

    def add_variable_injection_cost(self, date, cost):
        self.injection_costs.append((date, cost))
        self.injection_costs.sort(key=lambda c: c[0])

    def add_variable_withdrawal_cost(self, date, cost):
        self.withdrawal_costs.append((date, cost))
        self.withdrawal_costs.sort(key=lambda c: c[0])

    def get_variable_injection_cost(self, date):
        return self._cost_at(date, self.injection_costs)

    def get_variable_withdrawal_cost(self, date):
        return self._cost_at(date, self.withdrawal_costs)

    # -- reachable-inventory envelope ---------------------------------------------------------------------------------
    def _bisect_upper(self, date, period, env, i):
        """Lower env[i].vmax until the fastest withdrawal from it reaches env[i+1].vmax (storage_helpers.py:279-301:
        interval halving down to 1/1000 of the starting bracket, keeping the feasible end)."""
        target = env[i + 1].vmax
        lo, hi = target, env[i].vmax
        stop = (hi - lo) / 1000.0
        width = float("inf")
        while width > stop:
            mid = lo + 0.5 * (hi - lo)
            if mid - self.get_withdrawal_flexibility_rate(date, mid) * period <= target:
                lo = mid
            else:
                hi = mid
            width = hi - lo
        env[i].vmax = lo

    def _bisect_lower(self, date, period, env, i):
        """Raise env[i].vmin until the fastest injection from it reaches env[i+1].vmin (storage_helpers.py:303-320)."""
        target = env[i + 1].vmin
        hi, lo = target, env[i].vmin
        stop = (hi - lo) / 1000.0
        width = float("inf")
        while width > stop:
            mid = hi - 0.5 * (hi - lo)
            if mid + self.get_injection_flexibility_rate(date, mid) * period <= target:
                lo = mid
            else:
                hi = mid
            width = hi - lo
        env[i].vmin = hi

    def optimize_volume_constraints(self, start_date, end_date, rollout_interval, initial_volume):
        """Per action date, the inventory band that is reachable from the initial inventory and from which the
        later limits stay reachable (storage_helpers.py:322-437).  Forward sweeps cap the next date's band by what
        one period of injection / withdrawal can do; where a LATER limit cannot be met from the current band the
        band of the current date is tightened by bisection and the sweep starts over."""
        dates, contract, env = [], [], []
        t = start_date
        while t <= end_date + DATE_TOL:
            t_next = min(t + rollout_interval, end_date)
            w = self.get_initial_volume_constraint(t)
            lo, hi = w.vmin, w.vmax
            if _same(t, start_date, DATE_TOL):
                lo = hi = initial_volume
            contract.append(w)
            env.append(_Window(t, t_next, lo, hi, w.penalty))
            dates.append(t)
            if t >= end_date - DATE_TOL:
                break
            t = t_next

        again = True
        while again:
            again = False
            for i in range(len(env) - 1):
                date = env[i].start_date
                period = dates[i + 1] - dates[i]
                hi_i, hi_n, lo_i, lo_n = env[i].vmax, env[i + 1].vmax, env[i].vmin, env[i + 1].vmin
                wd_hi = self.get_withdrawal_flexibility_rate(date, hi_i) * period
                wd_lo = self.get_withdrawal_flexibility_rate(date, lo_i) * period
                inj_hi = self.get_injection_flexibility_rate(date, hi_i) * period
                inj_lo = self.get_injection_flexibility_rate(date, lo_i) * period

                if hi_i < hi_n:
                    if hi_i + inj_hi < hi_n:
                        env[i + 1].vmax = hi_i + inj_hi
                elif hi_i - wd_hi > hi_n:
                    self._bisect_upper(date, period, env, i)
                    again = True

                if lo_i < lo_n:
                    if lo_i + inj_lo < lo_n:
                        self._bisect_lower(date, period, env, i)
                        again = True
                elif lo_i - wd_lo > lo_n:
                    env[i + 1].vmin = lo_i - wd_lo

                for j in (i, i + 1):
                    if env[j].vmin > contract[j].vmax or env[j].vmax < contract[j].vmin:
                        raise ValueError(f"Initial volume constraints cannot be satisfied at date {dates[j]}.")
                if again:
                    break
        self.volume_constraints = env
