"""Expected negative exposure profile, mean -relu(-E_k) (reference: src/metrics/ene_metric.py:3-16)."""
from metrics.metric import *


class ENEMetric(Metric):
    def __init__(self, evaluation_type=Metric.EvaluationType.NUMERICAL):
        super().__init__(metric_type=MetricType.ENE, evaluation_type=evaluation_type)
