"""Simulation schemes (reference: src/common/enums.py:4-8)."""
from __future__ import annotations
from enum import Enum


class SimulationScheme(Enum):
    EULER = 0
    MILSTEIN = 1
    ANALYTICAL = 2
    QE = 3
