"""Cash-or-nothing digital with an always-fuzzy indicator of half-width 1
(reference: src/products/binary_option.py:6-64)."""
from products.product import *
from products.product import _ft
from products.european_option import _norm_cdf


class BinaryOption(Product):
    FUZZY_EPS = 1.0

    def __init__(self, maturity, strike, payment_amount, option_type, asset_id=None):
        super().__init__(asset_ids=[asset_id], product_family=ProductFamily.BINARY_TERMINAL_PAYOFF)
        self.maturity = _ft([maturity])
        self.strike = _ft([strike])
        self.option_type = option_type
        self.payment_amount = _ft([payment_amount])
        self.product_timeline = _ft([maturity])
        self.modeling_timeline = self.product_timeline

    def compute_pv_analytically(self, model):
        spot, rate, sigma = model.get_spot(), model.get_rate(), model.get_volatility()
        d2 = (torch.log(spot / self.strike) + (rate - 0.5 * sigma ** 2) * self.maturity) / (
            sigma * torch.sqrt(self.maturity))
        df = torch.exp(-rate * self.maturity)
        sign = 1.0 if self.option_type == OptionType.CALL else -1.0
        return self.payment_amount * df * _norm_cdf(sign * d2)
