"""Gas storage: a daily inject / hold / withdraw decision on a continuous inventory, valued by least-squares
Monte Carlo on an interpolated inventory grid (reference: src/products/storage.py:16-308).

The product is a contract description.  `lower()` flattens it into one record per action date for
csrc/storage.cu, which evaluates the reference's `compute_normalized_cashflows` (storage.py:215-308) for all
paths (and, in the backward induction, all grid states) in registers: inventory transition of the three actions,
interpolated continuation, arg-max, realised cashflow."""
from __future__ import annotations

from enum import Enum

import numpy as np

from products.product import *
from products.product import _ft
from products.storage_helpers import DATE_TOL, StorageConfig

#: knots per rate curve a date record can hold (csrc/storage.cu: STORAGE_MAX_KNOTS)
MAX_KNOTS = 8
#: doubles per date record: 16 header values + 2 curves x MAX_KNOTS x (level, rate)
RECORD = 16 + 4 * MAX_KNOTS


class StorageAction(Enum):
    INJECTION = 0
    WITHDRAWAL = 1
    DO_NOTHING = 2


class Storage(Product):
    def __init__(self, asset_id, start_date, end_date, initial_amount, storage_config: StorageConfig, num_states,
                 rollout_interval=1.0):
        super().__init__(asset_ids=[asset_id])
        if num_states < 2:
            raise ValueError("Storage requires at least two discrete states.")
        if rollout_interval <= 0.0:
            raise ValueError("Rollout interval must be positive.")
        self.start_date, self.end_date = float(start_date), float(end_date)
        self.initial_amount = float(initial_amount)
        self.storage_config = storage_config
        self.num_states = num_states
        self.rollout_interval = float(rollout_interval)
        storage_config.optimize_volume_constraints(start_date=self.start_date, end_date=self.end_date,
                                                   rollout_interval=self.rollout_interval,
                                                   initial_volume=self.initial_amount)
        # action dates by repeated accumulation, the last period cut at the end date (storage.py:47-55)
        acts, nexts = [], []
        t = self.start_date
        while t < self.end_date - DATE_TOL:
            t_next = min(t + self.rollout_interval, self.end_date)
            acts.append(t)
            nexts.append(t_next)
            t = t_next
        self.product_timeline = _ft(acts)
        self.modeling_timeline = self.product_timeline
        self.regression_timeline = self.product_timeline
        self.next_action_dates = _ft(nexts)

    def get_num_states(self):
        return self.num_states

    def get_state_dtype(self):
        return FLOAT

    def get_initial_state(self):
        return 0.0

    def state_to_volume(self, date, state):
        w = self.storage_config.get_volume_constraint(float(date))
        return w.vmin + torch.as_tensor(state, dtype=FLOAT) * StorageConfig.grid_step(w.vmin, w.vmax, self.num_states)

    # -- host-side views of the inventory moves (storage.py:114-213): the API the reference's unit tests exercise.
    #    The simulation does not call them; csrc/storage.cu:transitions evaluates the same rule per path and state.
    def _move(self, date, next_date, action, state):
        cfg = self.storage_config
        now, nxt = cfg.get_volume_constraint(date), cfg.get_volume_constraint(next_date)
        state = torch.as_tensor(state, dtype=FLOAT)
        vol = now.vmin + state * StorageConfig.grid_step(now.vmin, now.vmax, self.num_states)
        period = max(next_date - date, 0.0)
        if action == StorageAction.INJECTION:
            rate = cfg.interpolate_rate_tensor(vol, cfg.get_injection_flexibility_slice(date))
            new = torch.clamp(vol + rate * period, max=nxt.vmax)
        elif action == StorageAction.WITHDRAWAL:
            rate = cfg.interpolate_rate_tensor(vol, cfg.get_withdrawal_flexibility_slice(date))
            new = torch.clamp(vol - rate * period, min=nxt.vmin)
        else:
            new = torch.clamp(vol, min=nxt.vmin, max=nxt.vmax)
        return vol, new, nxt

    def compute_next_state(self, date, next_date, action_type):
        def mapping(previous_state):
            _, new, nxt = self._move(date, next_date, action_type, previous_state)
            scale = StorageConfig.state_scale(nxt.vmin, nxt.vmax, self.num_states)
            return torch.zeros_like(new) if scale == 0.0 else (new - nxt.vmin) * scale
        return mapping

    def compute_volume_difference(self, date, next_date, action_type):
        def mapping(previous_state):
            vol, new, _ = self._move(date, next_date, action_type, previous_state)
            return new - vol
        return mapping

    def lookup_state_values(self, values_by_state, state_matrix):
        b = torch.clamp(state_matrix.to(dtype=FLOAT), 0.0, self.num_states - 1.0)
        lo, hi = torch.floor(b).long(), torch.ceil(b).long()
        v_lo, v_hi = values_by_state.gather(dim=1, index=lo), values_by_state.gather(dim=1, index=hi)
        return v_lo + (b - lo.to(dtype=FLOAT)) * (v_hi - v_lo)

    def lower(self):
        """-> float64 array [n_dates, RECORD], one record per action date (layout shared with csrc/storage.cu):
             0 vmin of the date's band        1 inventory per state index (storage_helpers.py:56-60)
             2 vmin of the next date's band   3 vmax of the next date's band
             4 state index per unit inventory on the next date (storage_helpers.py:62-66; 0: degenerate band)
             5 period to the next action date 6 injection cost   7 withdrawal cost
             8 number of injection knots      9 number of withdrawal knots
            10 1.0 on the last action date (no continuation, storage.py:255-258)
            16.. injection knots (level, rate) x MAX_KNOTS, then withdrawal knots."""
        cfg, S = self.storage_config, self.num_states
        acts, nexts = self.product_timeline.tolist(), self.next_action_dates.tolist()
        rec = np.zeros((len(acts), RECORD))
        env = cfg.volume_constraints

        def band(i, t):
            # the envelope has one window per action date, in order: the linear search of get_volume_constraint
            # (454 x 455 window tests for a 15-month contract) finds exactly this one
            if i < len(env) and env[i].contains(t):
                return env[i]
            return cfg.get_volume_constraint(t)
        for i, (t, t_next) in enumerate(zip(acts, nexts)):
            now, nxt = band(i, t), band(i + 1, t_next)
            inj, wd = cfg.get_injection_flexibility_slice(t), cfg.get_withdrawal_flexibility_slice(t)
            if len(inj) > MAX_KNOTS or len(wd) > MAX_KNOTS:
                raise NotImplementedError(f"at most {MAX_KNOTS} knots per injection / withdrawal curve")
            r = rec[i]
            r[0], r[1] = now.vmin, StorageConfig.grid_step(now.vmin, now.vmax, S)
            r[2], r[3], r[4] = nxt.vmin, nxt.vmax, StorageConfig.state_scale(nxt.vmin, nxt.vmax, S)
            r[5] = max(t_next - t, 0.0)
            r[6], r[7] = cfg.get_variable_injection_cost(t), cfg.get_variable_withdrawal_cost(t)
            r[8], r[9] = len(inj), len(wd)
            r[10] = 1.0 if t_next >= self.end_date - DATE_TOL else 0.0
            for k, kn in enumerate(inj):
                r[16 + 2 * k], r[17 + 2 * k] = kn.point, kn.rate
            for k, kn in enumerate(wd):
                r[16 + 2 * MAX_KNOTS + 2 * k], r[17 + 2 * MAX_KNOTS + 2 * k] = kn.point, kn.rate
        return rec
