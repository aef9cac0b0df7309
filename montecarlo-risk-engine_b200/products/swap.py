"""Plain-vanilla interest-rate swap = floating leg minus fixed leg for a payer
(reference: src/products/swap.py:8-172)."""
from products.product import *
from products.product import _ft
from products.bond import Bond


class IRSType(Enum):
    PAYER = 0
    RECEIVER = 1


class InterestRateSwap(Product):
    def __init__(self, startdate, enddate, notional, fixed_rate, tenor_fixed, tenor_float, irs_type,
                 asset_id=None):
        super().__init__(asset_ids=[asset_id])
        self.startdate, self.enddate, self.notional = startdate, enddate, notional
        self.fixed_rate, self.tenor_fixed, self.tenor_float = fixed_rate, tenor_fixed, tenor_float
        self.irs_type = irs_type
        self.fixed_leg = Bond(startdate, enddate, notional, tenor_fixed, pays_notional=False,
                              fixed_rate=fixed_rate, asset_id=asset_id)
        self.floating_leg = Bond(startdate, enddate, notional, tenor_float, pays_notional=False,
                                 asset_id=asset_id)
        times = sorted(set(self.fixed_leg.payment_dates.tolist()) | set(self.floating_leg.payment_dates.tolist()))
        self.product_timeline = _ft(times)
        self.modeling_timeline = self.product_timeline
        self.regression_timeline = _ft([])

    def with_startdate(self, observation_date):
        return InterestRateSwap(observation_date, self.enddate, self.notional, self.fixed_rate,
                                self.tenor_fixed, self.tenor_float, self.irs_type,
                                asset_id=self.get_asset_id())

    def __eq__(self, other):
        return (isinstance(other, InterestRateSwap) and self.startdate == other.startdate
                and self.enddate == other.enddate and self.notional == other.notional
                and self.fixed_rate == other.fixed_rate and self.tenor_fixed == other.tenor_fixed
                and self.tenor_float == other.tenor_float)

    def __hash__(self):
        return hash((self.startdate, self.enddate, self.notional, self.fixed_rate, self.tenor_fixed,
                     self.tenor_float))
