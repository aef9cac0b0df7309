"""Collection of metrics + the metric exposure timeline
(reference: src/metrics/risk_metrics.py:9-69)."""
from enum import Enum
import numpy as np
from common.packages import *
from metrics.metric import Metric, MetricType


class PathwisePrimitive(Enum):
    DISCOUNTED_CASHFLOWS = "discounted_cashflows"
    EXPOSURE_PROFILES = "exposure_profiles"


class RiskMetrics:
    def __init__(self, metrics, exposure_timeline=None):
        self.metrics = metrics
        if exposure_timeline is None:
            exposure_timeline = []
        self.exposure_timeline = torch.tensor(np.asarray(exposure_timeline, dtype=float), dtype=FLOAT, device=device)
        self.any_pv = any(m.metric_type == MetricType.PV for m in metrics)
        self.any_xva = any(m.metric_type == MetricType.CVA for m in metrics)
        self.any_exposure = any(m.metric_type != MetricType.PV for m in metrics)
        prims = []
        if self.any_pv:
            prims.append(PathwisePrimitive.DISCOUNTED_CASHFLOWS)
        if self.any_exposure:
            prims.append(PathwisePrimitive.EXPOSURE_PROFILES)
        self._required_primitives = frozenset(prims)
        if self.any_exposure:
            assert len(exposure_timeline) > 0, \
                "For exposure simulation at least one exposure time point needs to be provided."
        for m in self.metrics:
            m.set_requests(exposure_timeline)
        self.counterparty_ids = [cp for m in self.metrics for cp in (m.get_counterparty_ids() or [])]

    def requires_discounted_cashflows(self):
        return PathwisePrimitive.DISCOUNTED_CASHFLOWS in self._required_primitives

    def requires_exposure_profiles(self):
        return PathwisePrimitive.EXPOSURE_PROFILES in self._required_primitives

    def required_pathwise_primitives(self):
        return self._required_primitives

    def requires_primitive(self, primitive):
        return primitive in self._required_primitives
