"""Request vocabulary (reference: src/request_interface/request_types.py:10-40).

The reference resolves requests into per-path vectors; here they are only a
host-side description that the plan compiler lowers into per-date scalar tables."""
from __future__ import annotations
from enum import Enum


class AtomicRequestType(Enum):
    SPOT = 1
    DISCOUNT_FACTOR = 2
    NUMERAIRE = 3
    FORWARD_RATE = 4
    LIBOR_RATE = 5
    SURVIVAL_PROBABILITY = 6
    CONDITIONAL_SURVIVAL_PROBABILITY = 7


class AtomicRequest:
    def __init__(self, request_type, time1=None, time2=None, id=None):
        self.request_type = request_type
        self.id = id
        self.time1 = time1
        self.time2 = time2
        self.handle = None

    def key(self):
        return (self.request_type, self.id, self.time1, self.time2)

    def __eq__(self, other):
        return isinstance(other, AtomicRequest) and self.key() == other.key()

    def __hash__(self):
        return hash(self.key())
