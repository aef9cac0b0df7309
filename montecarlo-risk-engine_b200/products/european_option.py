"""European option on an equity, bond or swap underlying
(reference: src/products/european_option.py:15-145).

The semi-analytic Heston and Vasicek bond-option pricers of the reference are
host-side validation helpers and out of scope; the
Black-Scholes closed form is kept because the controller's analytic PV /
analytic exposure shortcuts depend on it."""
import math
from products.product import *
from products.product import _ft
from models.black_scholes import BlackScholesModel
from models.black_scholes_multi import BlackScholesMulti


def _norm_cdf(x):
    return 0.5 * torch.erfc(-x / math.sqrt(2.0))


class EuropeanOption(Product):
    def __init__(self, underlying, exercise_date, strike, option_type, asset_id=None):
        super().__init__(asset_ids=[asset_id], product_family=ProductFamily.VANILLA_TERMINAL_OPTION)
        self.exercise_date = _ft([exercise_date])
        self.strike = _ft([strike])
        self.option_type = option_type
        self.product_timeline = _ft([exercise_date])
        self.modeling_timeline = self.product_timeline
        self.underlying = underlying
        #: the underlying as observed from the exercise date (reference: :43)
        self.underlying_at_exercise = underlying.with_startdate(exercise_date)

    # -- Black-Scholes closed form (reference: european_option.py:70-145) ------
    def _bs_spot_vol(self, model):
        spot, sigma = model.get_spot(), model.get_volatility()
        if spot.numel() > 1 or sigma.numel() > 1:
            asset_id = self.get_asset_id()
            if asset_id not in model.asset_ids:
                raise ValueError(f"Asset id '{asset_id}' not found in model asset ids {model.asset_ids}.")
            i = model.asset_ids.index(asset_id)
            spot, sigma = spot.reshape(-1)[i:i + 1], sigma.reshape(-1)[i:i + 1]
        return spot, sigma

    def _bs_price(self, spot, rate, sigma, ttm):
        vol_t = sigma * torch.sqrt(ttm)
        d1 = (torch.log(spot / self.strike) + (rate + 0.5 * sigma ** 2) * ttm) / vol_t
        d2 = d1 - vol_t
        disc_k = self.strike * torch.exp(-rate * ttm)
        if self.option_type == OptionType.CALL:
            return spot * _norm_cdf(d1) - disc_k * _norm_cdf(d2)
        return disc_k * _norm_cdf(-d2) - spot * _norm_cdf(-d1)

    def compute_pv_analytically(self, model):
        spot, sigma = self._bs_spot_vol(model)
        return self._bs_price(spot, model.get_rate(), sigma, self.exercise_date)

    def supports_analytic_pv(self, model):
        return isinstance(model, (BlackScholesModel, BlackScholesMulti))

    def supports_analytic_exposure(self, model):
        return isinstance(model, (BlackScholesModel, BlackScholesMulti))
