"""Potential future exposure: order statistic ceil(q N)-1 of the signed exposure
per date, with a finite-difference density standard error
(reference: src/metrics/pfe_metric.py:4-73)."""
import math
import numpy as np
from metrics.metric import *


class PFEMetric(Metric):
    def __init__(self, quantile=0.95, evaluation_type=Metric.EvaluationType.NUMERICAL):
        super().__init__(metric_type=MetricType.PFE, evaluation_type=evaluation_type)
        self.quantile = quantile

    def get_name(self):
        return f"pfe[{self.quantile:g}]"

    def quantile_index(self, num_paths):
        """The reference rounds q*N through a float32 tensor before the ceil
        (pfe_metric.py:59,65); reproduce that exactly."""
        return int(math.ceil(float(np.float32(self.quantile * num_paths)))) - 1
