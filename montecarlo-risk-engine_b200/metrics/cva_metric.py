"""Unilateral CVA = (1-R) * sum_k relu(E_k) S(0,t_k) (1 - S(t_k,t_{k+1}|y_k))
(reference: src/metrics/cva_metric.py:7-100)."""
from metrics.metric import *


class CVAMetric(Metric):
    def __init__(self, counterparty_id, recovery_rate, evaluation_type=Metric.EvaluationType.NUMERICAL):
        super().__init__(metric_type=MetricType.CVA, evaluation_type=evaluation_type)
        self.counterparty_id = counterparty_id
        self.recovery_rate = recovery_rate

    def get_counterparty_ids(self):
        return [self.counterparty_id]

    def get_name(self):
        return f"cva[{self.counterparty_id}]"
