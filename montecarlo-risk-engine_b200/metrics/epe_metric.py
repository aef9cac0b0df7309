"""Expected positive exposure profile, mean relu(E_k) (reference: src/metrics/epe_metric.py:3-16)."""
from metrics.metric import *


class EPEMetric(Metric):
    def __init__(self, evaluation_type=Metric.EvaluationType.NUMERICAL):
        super().__init__(metric_type=MetricType.EPE, evaluation_type=evaluation_type)
