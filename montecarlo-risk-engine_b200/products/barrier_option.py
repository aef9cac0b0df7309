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

    def compute_pv_analytically(self, model):
        """Continuously monitored single-barrier calls under Black-Scholes (Reiner-Rubinstein; reference:
        barrier_option.py:245-301, the benchmark of tests/pv_tests/pv_barrier_option.py): up-and-out with strike below
        the barrier, down-and-out with strike above it.  With x -> d(x) = (ln x + (r + sigma^2 / 2) T) / (sigma sqrt T),
        N+ = N(d(.)), N- = N(d(.) - sigma sqrt T) and the reflection image B^2 / (K S):
          up-and-out    S [N+(S/K) - N+(S/B)] - S (B/S)^(1 + 2r/s^2) [N+(B^2/(KS)) - N+(B/S)]
                        - K e^(-rT) {[N-(S/K) - N-(S/B)] - (S/B)^(1 - 2r/s^2) [N-(B^2/(KS)) - N-(B/S)]}      (S < B)
          down-and-out  vanilla call - S (B/S)^(2r/s^2) [(B/S) N+(B^2/(KS)) - (K/S) e^(-rT) N-(B^2/(KS))]     (S > B)
        Host helper; other barrier types return NotImplementedError like the reference."""
        import math
        S, r, sig = model.get_spot(), model.get_rate(), model.get_volatility()
        B, K, T = self.barrier1, self.strike, self.maturity
        vol = sig * torch.sqrt(T)
        ncdf = lambda x: 0.5 * torch.erfc(-x / math.sqrt(2.0))   # noqa: E731
        plus = lambda x: ncdf((torch.log(x) + (r + 0.5 * sig ** 2) * T) / vol)   # noqa: E731
        minus = lambda x: ncdf((torch.log(x) + (r + 0.5 * sig ** 2) * T) / vol - vol)   # noqa: E731
        disc_k = K * torch.exp(-r * T)
        image = B ** 2 / (K * S)
        if self.option_type != OptionType.CALL:
            return NotImplementedError(f"Analytical method for {self.barrier_option_type1} {self.option_type} not yet implemented")
        if self.barrier_option_type1 == BarrierOptionType.UPANDOUT:
            spot_leg = S * ((plus(S / K) - plus(S / B)) - (B / S) ** (1.0 + 2.0 * r / sig ** 2) * (plus(image) - plus(B / S)))
            strike_leg = disc_k * ((minus(S / K) - minus(S / B)) - (S / B) ** (1.0 - 2.0 * r / sig ** 2) * (minus(image) - minus(B / S)))
            return (S < B) * (spot_leg - strike_leg)
        if self.barrier_option_type1 == BarrierOptionType.DOWNANDOUT:
            vanilla = S * plus(S / K) - disc_k * minus(S / K)
            mirror = (B / S) * plus(image) - (K / S) * torch.exp(-r * T) * minus(image)
            return (S > B) * (vanilla - S * (B / S) ** (2.0 * r / sig ** 2) * mirror)
        return NotImplementedError(f"Analytical method for {self.barrier_option_type1} not yet implemented")
