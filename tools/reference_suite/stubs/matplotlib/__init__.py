from unittest import mock
import sys
def use(*a, **k): pass
def __getattr__(name):
    return mock.MagicMock()
