"""Metric descriptions (reference: src/metrics/metric.py:6-60).

A metric here only says *what* to reduce; the reductions themselves (sum, sum of
squares, order statistics) are done by the kernels and finished on the host in
`mcre.finish`.  MC error = unbiased std / sqrt(N) (reference: metric.py:26-35)."""
from enum import Enum
from common.packages import *


class MetricType(Enum):
    PV = "Present Value"
    CE = "Current Exposure"
    EPE = "Expected Positive Exposure"
    ENE = "Expected Negative Exposure"
    PFE = "Potential Future Exposure"
    EEPE = "Effective Expected Positive Exposure"
    CVA = "Credit Valuation Adjustment"


class Metric:
    class EvaluationType(Enum):
        ANALYTICAL = "Analytical"
        NUMERICAL = "Numerical"

    def __init__(self, metric_type, evaluation_type):
        self.metric_type = metric_type
        self.evaluation_type = evaluation_type

    def set_requests(self, exposure_timeline):
        pass

    def get_counterparty_ids(self):
        return None

    def get_name(self):
        return self.metric_type.name.lower()
