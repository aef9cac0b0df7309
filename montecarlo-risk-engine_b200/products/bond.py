"""Fixed / floating coupon bond (reference: src/products/bond.py:6-214)."""
from products.product import *
from products.product import _ft


class Bond(Product):
    def __init__(self, startdate, maturity, notional, tenor, pays_notional=True, fixed_rate=None,
                 asset_id=None):
        super().__init__(asset_ids=[asset_id])
        self.startdate = _ft([startdate])
        self.maturity = _ft([maturity])
        self.notional = _ft([notional])
        self.tenor = _ft([tenor])
        self.fixed_rate = None if fixed_rate is None else _ft([fixed_rate])
        self.pays_notional = pays_notional
        # Payment dates by repeated accumulation (NOT startdate + i*tenor): the
        # simulation grid is merged on float equality, so the rounding of the
        # accumulation is part of the contract (reference: bond.py:37-68).
        dates, d = [], startdate + tenor
        while d < maturity:
            dates.append(d)
            d += tenor
        dates.append(maturity)
        #: accrual start of the last coupon as the reference computes it (date - tenor)
        self.last_accrual_start = d - tenor
        self.payment_dates = _ft(dates)
        self.product_timeline = self.payment_dates
        self.modeling_timeline = self.payment_dates
        self.regression_timeline = _ft([])

    def is_fixed(self):
        return self.fixed_rate is not None

    def libor_periods(self):
        """(t1, t2) of the LIBOR request attached to each payment date of a floating
        leg.  The reference builds t1 as ``date - tenor`` (reference: bond.py:56,64)."""
        out, tenor = [], float(self.tenor)
        dates = self.payment_dates.tolist()
        for d in dates[:-1]:
            out.append((d - tenor, d))
        out.append((self.last_accrual_start, dates[-1]))
        return out

    def accrual_fractions(self):
        """dt used for the coupon amount: payment date minus previous payment date
        (startdate for the first) (reference: bond.py:174-178)."""
        dates = self.payment_dates.tolist()
        prev = [float(self.startdate)] + dates[:-1]
        return [d - p for d, p in zip(dates, prev)]

    def with_startdate(self, observation_date):
        """The same bond observed (re-scheduled) from ``observation_date``; used for
        option underlyings (reference: bond.py:102-113)."""
        return Bond(observation_date, float(self.maturity), float(self.notional), float(self.tenor),
                    self.pays_notional, None if self.fixed_rate is None else float(self.fixed_rate),
                    asset_id=self.get_asset_id())

    def __eq__(self, other):
        return (isinstance(other, Bond) and float(self.startdate) == float(other.startdate)
                and float(self.maturity) == float(other.maturity) and float(self.tenor) == float(other.tenor)
                and (None if self.fixed_rate is None else float(self.fixed_rate))
                == (None if other.fixed_rate is None else float(other.fixed_rate))
                and self.pays_notional == other.pays_notional)

    def __hash__(self):
        return hash((float(self.startdate), float(self.maturity), float(self.tenor),
                     None if self.fixed_rate is None else float(self.fixed_rate), self.pays_notional))
