"""Netting set description (reference: src/products/netting_set.py:12-184).

Netting, the threshold dead-band and the MPoR-delayed collateral are applied per
path inside the fused main-simulation kernel (csrc/irc_kernels.cu,
`apply_collateral`); this class only validates and carries the terms."""
from __future__ import annotations
from dataclasses import dataclass
from typing import Sequence
import torch
from products.product import Product


@dataclass
class NettingSet:
    name: str
    products: Sequence[Product]
    threshold: float = 0.0
    margin_period_of_risk: float | None = None
    counterparty_id: str | None = None
    collateral_interpolation: str = "linear"

    def __post_init__(self):
        self.products = list(self.products)
        if len(self.products) == 0:
            raise ValueError("A netting set must contain at least one product.")
        if self.threshold < 0.0:
            raise ValueError("Netting set threshold must be non-negative.")
        if self.margin_period_of_risk is not None and self.margin_period_of_risk < 0.0:
            raise ValueError("Netting set margin period of risk must be non-negative.")
        if self.collateral_interpolation not in {"linear", "previous"}:
            raise ValueError("Collateral interpolation must be one of {'linear', 'previous'}.")

    def get_name(self):
        return self.name

    def is_collateralized(self):
        return self.margin_period_of_risk is not None

    def get_collateral_query_times(self, exposure_timeline: torch.Tensor) -> torch.Tensor:
        if not self.is_collateralized():
            return torch.zeros(0, dtype=exposure_timeline.dtype, device=exposure_timeline.device)
        delayed = exposure_timeline - self.margin_period_of_risk
        return delayed[delayed >= 0.0]

    # -- exposure-tensor helpers (reference: netting_set.py:48-184) ----------------------------------
    # The simulation applies these terms per path inside the fused kernels; the methods below give the
    # same arithmetic on caller-held [T_e, N] exposure tensors (scripts / unit checks).
    def apply_threshold(self, exposures: torch.Tensor) -> torch.Tensor:
        """Symmetric dead band: x - h above h, x + h below -h, 0 inside."""
        h = float(self.threshold)
        if h == 0.0:
            return exposures
        return torch.where(exposures > h, exposures - h,
                           torch.where(exposures < -h, exposures + h, torch.zeros_like(exposures)))

    def _metric_and_delayed_indices(self, exposure_timeline, metric_exposure_indices, delayed_exposure_indices):
        n = exposure_timeline.shape[0]
        dev = exposure_timeline.device
        if metric_exposure_indices is None:
            metric_exposure_indices = torch.arange(n, dtype=torch.long, device=dev)
        if delayed_exposure_indices is None:
            lookup = {float(t): i for i, t in enumerate(exposure_timeline.tolist())}
            mpor = float(self.margin_period_of_risk or 0.0)
            times = exposure_timeline.index_select(0, metric_exposure_indices).tolist()
            delayed_exposure_indices = torch.tensor([lookup.get(float(t - mpor), -1) if t - mpor >= 0.0 else -1
                                                     for t in times], dtype=torch.long, device=dev)
        return metric_exposure_indices, delayed_exposure_indices

    def compute_collateral_profile(self, netted_exposures, exposure_timeline, metric_exposure_indices=None,
                                   delayed_exposure_indices=None):
        """Collateral held at each metric date: the threshold-adjusted netted exposure observed one margin
        period earlier, looked up by exact index (-1: nothing posted yet)."""
        metric_idx, delayed_idx = self._metric_and_delayed_indices(exposure_timeline, metric_exposure_indices,
                                                                  delayed_exposure_indices)
        out = torch.zeros((metric_idx.shape[0],) + tuple(netted_exposures.shape[1:]), dtype=netted_exposures.dtype,
                          device=netted_exposures.device)
        if not self.is_collateralized() or netted_exposures.numel() == 0:
            return out
        have = delayed_idx >= 0
        if bool(have.any()):
            out[have] = self.apply_threshold(netted_exposures.index_select(0, delayed_idx[have]))
        return out

    def compute_unsecured_exposure_profiles(self, netted_exposures, exposure_timeline, metric_exposure_indices=None,
                                            delayed_exposure_indices=None):
        """Netted exposure at the metric dates after the threshold (uncollateralised sets) or after
        subtracting the collateral profile (collateralised sets)."""
        if netted_exposures.numel() == 0:
            return netted_exposures
        metric_idx, delayed_idx = self._metric_and_delayed_indices(exposure_timeline, metric_exposure_indices,
                                                                  delayed_exposure_indices)
        at_metric = netted_exposures.index_select(0, metric_idx)
        if not self.is_collateralized():
            return self.apply_threshold(at_metric)
        return at_metric - self.compute_collateral_profile(netted_exposures, exposure_timeline, metric_idx, delayed_idx)
