"""CDS bootstrap of piecewise-constant hazard rates and the default probability they imply
(reference: src/helpers/cs_helper.py:9-107, used by tests/exposure_tests/cva_*.py to build the CIR++ market curve).

Host code, numpy per hazard bucket.  The reference's leg conventions are kept as they are, because the hazards a
caller bootstraps feed CIRPPModel and must not move:
  * inside the bucket (T_{i-1}, T_i] every coupon date t_k measures default against the survival probability at the
    START of the bucket, Q(T_{i-1}) - Q(t_k), for both the accrual-on-default and the protection term;
  * the coupon that falls on T_i closes bucket i and opens bucket i + 1 (where it adds its premium once more and no
    default term, as no time has passed in that bucket).
"""
from __future__ import annotations

import numpy as np
import torch

from maths.maths import bisection_search


class CSHelper:
    def _compute_cds_legs(self, maturities, payment_days, discount_factors_payment_days, recovery_rate, hazard_rates):
        """-> (premium leg per unit spread, protection leg) of a CDS maturing at maturities[-1]."""
        t = np.asarray(payment_days, dtype=float)
        df = np.asarray(discount_factors_payment_days, dtype=float)
        accr = np.diff(t, prepend=0.0)                       # coupon accrual periods (the first one starts today)
        last = np.searchsorted(t, np.asarray(maturities, dtype=float))
        premium = protection = 0.0
        q_start, t_start, first = 1.0, 0.0, 0
        for hazard, k_end, maturity in zip(hazard_rates, last, maturities):
            k = slice(first, int(k_end) + 1)
            q = q_start * np.exp(-hazard * (t[k] - t_start))
            defaulted = q_start - q
            premium += float(np.sum(accr[k] * df[k] * (q + 0.5 * defaulted)))
            protection += float((1.0 - recovery_rate) * np.sum(df[k] * defaulted))
            q_start, t_start, first = float(q[-1]), float(maturity), int(k_end)
        return premium, protection

    def bootstrap_hazards(self, credit_spreads, maturities, payment_days, discount_factors_payment_days, recovery_rate):
        """Hazard of bucket i = root of spread_i * premium(T_i) - protection(T_i) with the earlier buckets fixed.
        Every maturity must be a coupon date."""
        if len(payment_days) != len(discount_factors_payment_days):
            raise AssertionError("one discount factor per payment day")
        solved = []
        for i, spread in enumerate(credit_spreads):
            upto = maturities[:i + 1]

            def par_gap(candidate):
                prem, prot = self._compute_cds_legs(upto, payment_days, discount_factors_payment_days, recovery_rate,
                                                    solved + [candidate])
                return spread * prem - prot
            solved.append(bisection_search(par_gap))
        return solved

    def probability_of_default(self, hazards: torch.Tensor, tenors: torch.Tensor, date: torch.Tensor) -> torch.Tensor:
        """Cumulative default probability at `date`: hazards[i] on (tenors[i-1], tenors[i]], the last one extended flat."""
        h = torch.as_tensor(hazards, dtype=torch.float64).reshape(-1)
        ends = torch.as_tensor(tenors, dtype=torch.float64).reshape(-1)
        d = torch.as_tensor(date, dtype=torch.float64).reshape(())
        starts = torch.cat([torch.zeros(1, dtype=torch.float64), ends[:-1]])
        time_in_bucket = (torch.minimum(ends, d) - starts).clamp(min=0.0)
        time_in_bucket[-1] = time_in_bucket[-1] + (d - ends[-1]).clamp(min=0.0)
        return 1.0 - torch.exp(-(h * time_in_bucket).sum())
