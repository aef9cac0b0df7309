"""FlexiCall: a strip of European options of which at most `num_exercise_rights` may be exercised
(reference: src/products/flexicall.py:4-186).

Host description only.  States = rights left (initial state = num_exercise_rights, state 0 = nothing
left).  At exercise date i with strike K_i a path in state s > 0 exercises iff
    immediate_i + continuation_i(s - 1) > continuation_i(s)        (hard indicator, flexicall.py:139-142)
where continuation_i(s) is the regression proxy of state s at that date (zero on the last date and for
state 0); it then receives immediate_i / numeraire and moves to state s - 1.  The Longstaff-Schwartz
pre-simulation (mcre/lsm.py, csrc/lsm.cu) and the fused main kernel (csrc/equity.cu) carry the state."""
from products.european_option import *
from products.product import _ft


class FlexiCall(Product):
    def __init__(self, underlyings, num_exercise_rights, asset_id=None):
        super().__init__(asset_ids=[asset_id], product_family=ProductFamily.FLEXICALL_EXERCISE)
        assert num_exercise_rights <= len(underlyings), "Number of exercise rights cannot exceed number of underlyings"
        assert all(o.option_type == underlyings[0].option_type for o in underlyings), \
            "All underlyings must have the same option type"
        self.underlyings = sorted(underlyings, key=lambda o: float(o.exercise_date[0]))
        dates = [float(o.exercise_date[0]) for o in self.underlyings]
        assert all(a < b for a, b in zip(dates, dates[1:])), "Exercise dates must be distinct"
        self.product_timeline = _ft(dates)
        self.modeling_timeline = self.product_timeline
        self.regression_timeline = self.product_timeline
        self.num_exercise_rights = int(num_exercise_rights)
        self.option_type = self.underlyings[0].option_type
        #: the option strip's underlying (an Equity), as observed from each exercise date
        self.underlying = self.underlyings[0].underlying
        self.strikes = [float(o.strike[0]) for o in self.underlyings]

    def get_num_states(self):
        return self.num_exercise_rights + 1

    def get_initial_state(self):
        return self.num_exercise_rights
