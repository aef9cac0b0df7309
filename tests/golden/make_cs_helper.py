"""Fixture of the reference's CDS bootstrap (src/helpers/cs_helper.py) for tests/test_analytic_host.py: run in the build
container, where /root/reference exists:  python tests/golden/make_cs_helper.py"""
import json
import os
import sys

sys.path.insert(0, "/root/reference/src")
import numpy as np  # noqa: E402
import torch  # noqa: E402
from helpers.cs_helper import CSHelper  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
tenors = [0.5, 1.0, 2.0, 3.0, 4.0, 5.0, 7.0, 10.0, 15.0, 20.0]
spreads = [0.0021, 0.0028, 0.0041, 0.0060, 0.0078, 0.0095, 0.0118, 0.0131, 0.0139, 0.0142]
pay = np.arange(0.25, 20.0 + 1e-7, 0.25)
df = np.exp(-0.02 * pay)
h = CSHelper()
haz = h.bootstrap_hazards(credit_spreads=spreads, maturities=tenors, payment_days=pay, discount_factors_payment_days=df,
                          recovery_rate=0.4)
legs = [h._compute_cds_legs(tenors[:i + 1], pay, df, 0.4, haz[:i + 1]) for i in range(len(tenors))]
dates = [0.0, 0.1, 0.5, 0.75, 3.0, 6.2, 20.0, 23.5]
pd_ = [float(h.probability_of_default(torch.tensor(haz, dtype=torch.float64), torch.tensor(tenors, dtype=torch.float64),
                                      torch.tensor(t, dtype=torch.float64))) for t in dates]
json.dump(dict(tenors=tenors, spreads=spreads, rate=0.02, recovery=0.4, hazards=[float(x) for x in haz],
               legs=[[float(a), float(b)] for a, b in legs], dates=dates, default_probability=pd_),
          open(os.path.join(HERE, "cs_helper.json"), "w"), indent=1)
print("wrote cs_helper.json")
