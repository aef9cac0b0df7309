"""Equity underlying: value = spot of its asset (reference: src/products/equity.py:7-40)."""
from products.product import *


class Equity(Product):
    def __init__(self, asset_id=None):
        super().__init__(asset_ids=[asset_id])

    def with_startdate(self, observation_date):
        return Equity(self.get_asset_id())

    def __eq__(self, other):
        return isinstance(other, Equity) and self.get_asset_id() == other.get_asset_id()

    def __hash__(self):
        return hash(self.get_asset_id())
