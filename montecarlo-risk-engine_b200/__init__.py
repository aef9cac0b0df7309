"""B200-native Monte Carlo risk engine (hot path of konstantineder/montecarlo-risk-engine).

This directory plays the role of the reference's ``src/`` directory: put it on
``sys.path`` and the reference's import statements (``from controller.controller
import SimulationController``, ``from models.vasicek import VasicekModel`` ...)
resolve to this implementation (reference: tests/pytests/context.py:1-4).

The directory name contains a hyphen, so import it with
``importlib.import_module("montecarlo-risk-engine_b200")``; importing it puts the
directory on ``sys.path``.
"""
import os as _os
import sys as _sys

PACKAGE_DIR = _os.path.dirname(_os.path.abspath(__file__))
if PACKAGE_DIR not in _sys.path:
    _sys.path.insert(0, PACKAGE_DIR)

__all__ = ["PACKAGE_DIR"]
