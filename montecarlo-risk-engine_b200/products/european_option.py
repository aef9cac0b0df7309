"""European option on an equity, bond or swap underlying
(reference: src/products/european_option.py:15-145).

The Black-Scholes closed form drives the controller's analytic PV / analytic exposure
shortcuts.  The semi-analytic Heston price (a host-side validation helper of the
reference, european_option.py:146-242, used by tests/pytests/test_pv_european_option_heston.py)
is provided through one characteristic function of log S_T in the branch-cut-free form;
the Vasicek bond-option closed form is a host helper next to it."""
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

    # -- option on a zero-coupon bond under Vasicek (reference: european_option.py:264-288) ----------------------
    def compute_pv_bond_option_analytically(self, model):
        """Jamshidian's closed form: with P(0,T) the discount bond to the exercise date T, P(0,S) to the bond's maturity
        S and v = sigma sqrt((1 - exp(-2aT)) / (2a)) (1 - exp(-a(S - T))) / a the volatility of the forward bond price,
        call = P(0,S) N(d1) - K P(0,T) N(d2), d1 = ln(P(0,S) / (K P(0,T))) / v + v / 2, d2 = d1 - v.  Host helper."""
        from products.bond import Bond
        if not isinstance(self.underlying, Bond):
            raise TypeError("Expected self.underlying to be of type Bond")
        a, r0, sigma, t0 = model.get_mean_reversion_speed(), model.get_rate(), model.get_volatility(), model.calibration_date
        expiry, maturity = self.exercise_date, self.underlying.maturity
        p_expiry = model.compute_bond_price(t0, expiry, r0)
        p_maturity = model.compute_bond_price(t0, maturity, r0)
        v = sigma * torch.sqrt((1.0 - torch.exp(-2.0 * a * (expiry - t0))) / (2.0 * a)) * (1.0 - torch.exp(-a * (maturity - expiry))) / a
        d1 = torch.log(p_maturity / (p_expiry * self.strike)) / v + 0.5 * v
        d2 = d1 - v
        if self.option_type == OptionType.CALL:
            return p_maturity * _norm_cdf(d1) - self.strike * p_expiry * _norm_cdf(d2)
        return self.strike * p_expiry * _norm_cdf(-d2) - p_maturity * _norm_cdf(-d1)

    # -- semi-analytic Heston price (validation helper, host only) --------------
    def compute_pv_analytically_heston(self, model):
        """Call/put under Heston by Fourier inversion of the characteristic function of log S_T:
        C = S0 P1 - K exp(-rT) P2,  P2 = 1/2 + 1/pi int_0^inf Re[exp(-iu ln K) phi(u) / (iu)] du and
        P1 the same with phi(u - i) / phi(-i).  phi is written with the root d and ratio g that keep
        exp(-dT) decaying, so the complex logarithm never crosses its branch cut."""
        from models.heston import HestonModel
        if not isinstance(model, HestonModel):
            raise TypeError("Expected model to be of type HestonModel")
        import numpy as np
        from scipy.integrate import quad
        s0, sig, r, rho, kappa, theta, v0 = (float(q) for q in model.model_params)
        K, T = float(self.strike[0]), float(self.exercise_date[0])
        lnk, fwd = math.log(K), math.log(s0) + r * T

        def phi(u):
            iu = 1j * u
            beta = kappa - rho * sig * iu
            d = np.sqrt(beta * beta + sig * sig * (iu + u * u))
            if d.real < 0:
                d = -d
            g = (beta - d) / (beta + d)
            e = np.exp(-d * T)
            A = kappa * theta / (sig * sig) * ((beta - d) * T - 2.0 * np.log((1.0 - g * e) / (1.0 - g)))
            Bv = (beta - d) / (sig * sig) * (1.0 - e) / (1.0 - g * e)
            return np.exp(iu * fwd + A + Bv * v0)

        norm = s0 * math.exp(r * T)          # phi(-i): the forward
        p2 = 0.5 + quad(lambda u: (np.exp(-1j * u * lnk) * phi(u) / (1j * u)).real, 1e-12, 200.0, limit=400)[0] / math.pi
        p1 = 0.5 + quad(lambda u: (np.exp(-1j * u * lnk) * phi(u - 1j) / (1j * u * norm)).real, 1e-12, 200.0,
                        limit=400)[0] / math.pi
        call = s0 * p1 - K * math.exp(-r * T) * p2
        if self.option_type == OptionType.CALL:
            return call
        return call - s0 + K * math.exp(-r * T)      # put-call parity

    # -- closed-form second-order Greeks (validation helpers; reference: european_option.py:290-320) --
    def _bs_d1_d2(self, model):
        spot, sigma = self._bs_spot_vol(model)
        vol_t = sigma * torch.sqrt(self.exercise_date)
        d1 = (torch.log(spot / self.strike) + (model.get_rate() + 0.5 * sigma ** 2) * self.exercise_date) / vol_t
        return spot, sigma, vol_t, d1, d1 - vol_t

    def compute_dDeltadSpot_analytically(self, model):
        """Black-Scholes gamma  phi(d1) / (S sigma sqrt(T))."""
        spot, sigma, vol_t, d1, _ = self._bs_d1_d2(model)
        return torch.exp(-0.5 * d1 ** 2) / math.sqrt(2.0 * math.pi) / (spot * vol_t)

    def compute_dVegadSigma_analytically(self, model):
        """Black-Scholes vomma  vega d1 d2 / sigma,  vega = S phi(d1) sqrt(T)."""
        spot, sigma, vol_t, d1, d2 = self._bs_d1_d2(model)
        vega = spot * torch.exp(-0.5 * d1 ** 2) / math.sqrt(2.0 * math.pi) * torch.sqrt(self.exercise_date)
        return vega * d1 * d2 / sigma

    def compute_pv_analytically(self, model):
        spot, sigma = self._bs_spot_vol(model)
        return self._bs_price(spot, model.get_rate(), sigma, self.exercise_date)

    def supports_analytic_pv(self, model):
        return isinstance(model, (BlackScholesModel, BlackScholesMulti))

    def supports_analytic_exposure(self, model):
        return isinstance(model, (BlackScholesModel, BlackScholesMulti))
