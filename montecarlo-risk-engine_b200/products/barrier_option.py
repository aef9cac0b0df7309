"""Discretely monitored single / double barrier option with an always-fuzzy
barrier indicator (reference: src/products/barrier_option.py:15-314)."""
from products.product import *
from products.product import _ft


class BarrierOptionType(Enum):
    DOWNANDOUT = "Down-And-Out"
    UPANDOUT = "Up-And-Out"
    DOWNANDIN = "Down-And-In"
    UPANDIN = "Up-And-In"


class BarrierOption(Product):
    FUZZY_EPS = 0.05

    def __init__(self, startdate, maturity, strike, num_observation_timepoints, option_type, barrier1,
                 barrier_option_type1, barrier2=None, barrier_option_type2=None, asset_id=None, basket=None):
        super().__init__(asset_ids=[asset_id], product_family=ProductFamily.BARRIER_PATH_TERMINAL)
        #: extension (not in the reference, which monitors one asset): (asset_ids, weights) of a
        #: weighted arithmetic basket monitored instead of a single spot (BASELINE config 5)
        self.basket = None if basket is None else (list(basket[0]), [float(w) for w in basket[1]])
        if self.basket is not None:
            self.asset_ids = list(self.basket[0])
        self.strike = _ft([strike])
        self.maturity = _ft([maturity])
        self.product_timeline = _ft([maturity])
        self.modeling_timeline = torch.linspace(startdate, maturity, num_observation_timepoints,
                                                dtype=FLOAT, device=device)
        self.barrier1 = _ft([barrier1])
        self.barrier_option_type1 = barrier_option_type1
        self.barrier2 = None if barrier2 is None else _ft([barrier2])
        self.barrier_option_type2 = barrier_option_type2
        self.option_type = option_type
        self.use_brownian_bridge = False

    def set_use_brownian_bridge(self):
        """Brownian-bridge correction between monitoring dates (reference: barrier_option.py:62-63, 138-222):
        per interval the crossing probability exp(-2 ln(S_i/B) ln(S_i+1/B) / (sigma^2 maturity / n_obs)) is
        compared (fuzzy, eps 0.05) with one uniform per path and interval.  The kernel draws those uniforms from
        Philox (kind 2); in RNG compatibility mode the reference's numpy default_rng(12345) stream is injected.
        Single Black-Scholes model, value-only runs."""
        self.use_brownian_bridge = True
