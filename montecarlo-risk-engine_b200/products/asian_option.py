"""Discretely monitored Asian option (reference: src/products/asian_option.py:11-95)."""
from products.product import *
from products.product import _ft


class AsianAveragingType(Enum):
    ARITHMETIC = 0
    GEOMETRIC = 1


class AsianOption(Product):
    def __init__(self, startdate, maturity, strike, num_observation_timepoints, option_type,
                 averaging_type=AsianAveragingType.ARITHMETIC, asset_id=None, basket=None):
        super().__init__(asset_ids=[asset_id], product_family=ProductFamily.ASIAN_PATH_TERMINAL)
        #: extension (not in the reference, which monitors one asset): (asset_ids, weights) of a
        #: weighted arithmetic basket monitored instead of a single spot (BASELINE config 5)
        self.basket = None if basket is None else (list(basket[0]), [float(w) for w in basket[1]])
        if self.basket is not None:
            self.asset_ids = list(self.basket[0])
        self.maturity = _ft([maturity])
        self.strike = _ft([strike])
        self.option_type = option_type
        self.averaging_type = averaging_type
        self.product_timeline = _ft([maturity])
        self.modeling_timeline = torch.linspace(startdate, maturity, num_observation_timepoints,
                                                dtype=FLOAT, device=device)
