"""Time average of the EPE profile; error = std over time / sqrt(T) (reference: src/metrics/eepe_metric.py:3-15)."""
from metrics.metric import *


class EEPEMetric(Metric):
    def __init__(self, evaluation_type=Metric.EvaluationType.NUMERICAL):
        super().__init__(metric_type=MetricType.EEPE, evaluation_type=evaluation_type)
