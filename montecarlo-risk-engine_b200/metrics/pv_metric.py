"""Present value = mean of discounted cashflows (reference: src/metrics/pv_metric.py:3-18)."""
from metrics.metric import *


class PVMetric(Metric):
    def __init__(self, evaluation_type=Metric.EvaluationType.NUMERICAL):
        super().__init__(metric_type=MetricType.PV, evaluation_type=evaluation_type)
