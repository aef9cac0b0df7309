"""Arithmetic / geometric basket option, optional geometric control variate
(reference: src/products/basket_option.py:10-142)."""
from products.product import *
from products.product import _ft
from products.european_option import _norm_cdf


class BasketOptionType(Enum):
    ARITHMETIC = 0
    GEOMETRIC = 1


class BasketOption(Product):
    def __init__(self, maturity, asset_ids, weights, strike, option_type,
                 basket_option_type=BasketOptionType.ARITHMETIC, use_variation_reduction=False):
        super().__init__(asset_ids=asset_ids, product_family=ProductFamily.BASKET_TERMINAL_PAYOFF)
        self.maturity = _ft([maturity])
        self.strike = _ft([strike])
        self.weights = _ft(weights)
        self.option_type = option_type
        self.product_timeline = _ft([maturity])
        self.modeling_timeline = self.product_timeline
        self.basket_option_type = basket_option_type
        self.use_variation_reduction = use_variation_reduction

    def compute_pv_analytically(self, model):
        """Geometric-basket closed form under BlackScholesMulti (reference: :103-141)."""
        S, r, sig = model.get_spot(), model.get_rate(), model.get_volatility()
        T, K, w, n = self.maturity, self.strike, self.weights, len(S)
        Sd = torch.diag(sig)
        cov = Sd @ model.correlation_matrix @ Sd * T
        var = torch.dot(w, torch.mv(cov, w))
        sigma = torch.sqrt(var)
        F = torch.exp(torch.log(S).mean()) * torch.exp((r - 0.5 * torch.sum(sig ** 2) / n + 0.5 * sigma ** 2) * T)
        sst = sigma * torch.sqrt(T)
        d1 = (torch.log(F / K) + 0.5 * sigma ** 2 * T) / sst
        d2 = d1 - sst
        if self.option_type == OptionType.CALL:
            return torch.exp(-r * T) * (F * _norm_cdf(d1) - K * _norm_cdf(d2))
        return torch.exp(-r * T) * (K * _norm_cdf(-d2) - F * _norm_cdf(-d1))
