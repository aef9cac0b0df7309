"""Bermudan / American option with one exercise right
(reference: src/products/bermudan_option.py:6-193)."""
import numpy as np
from products.product import *
from products.product import _ft


class BermudanOption(Product):
    def __init__(self, underlying, exercise_dates, strike, option_type, asset_id=None):
        super().__init__(asset_ids=[asset_id], product_family=ProductFamily.BERMUDAN_EXERCISE)
        self.strike = _ft([strike])
        self.option_type = option_type
        self.product_timeline = torch.tensor(np.asarray(exercise_dates, dtype=float), dtype=FLOAT, device=device)
        self.modeling_timeline = self.product_timeline
        self.regression_timeline = self.product_timeline
        self.num_exercise_rights = 1
        self.underlying = underlying
        self.underlyings_at_exercise = [underlying.with_startdate(float(t)) for t in self.product_timeline]

    def get_num_states(self):
        return 2

    def get_initial_state(self):
        return 1


class AmericanOption(BermudanOption):
    def __init__(self, underlying, maturity, num_exercise_dates, strike, option_type, asset_id=None):
        dates = np.linspace(0.0, maturity, num_exercise_dates) if num_exercise_dates > 1 else [maturity]
        super().__init__(underlying, dates, strike, option_type, asset_id=asset_id)
