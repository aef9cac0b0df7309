"""Current exposure = mean relu(E_0) (reference: src/metrics/ce_metric.py:3-13)."""
from metrics.metric import *


class CEMetric(Metric):
    def __init__(self, evaluation_type=Metric.EvaluationType.NUMERICAL):
        super().__init__(metric_type=MetricType.CE, evaluation_type=evaluation_type)
