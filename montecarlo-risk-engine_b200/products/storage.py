"""Gas storage (reference: src/products/storage.py:16-308) is OUT OF SCOPE of this build (SURVEY §8f item 3,
DESIGN.md §7): a continuous-inventory stochastic control problem with an interpolated continuation grid.
The names exist so that scripts importing them alongside supported products keep importing; constructing a
Storage raises."""
from enum import Enum


class StorageAction(Enum):
    WITHDRAW = -1
    HOLD = 0
    INJECT = 1


class Storage:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("gas storage is not implemented in this build (SURVEY §8f item 3)")
