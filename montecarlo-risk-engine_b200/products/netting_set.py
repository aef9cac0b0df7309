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
