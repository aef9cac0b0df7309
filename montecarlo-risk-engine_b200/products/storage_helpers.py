"""Gas-storage helpers of the reference (src/products/storage_helpers.py:48-437): out of scope, see
products/storage.py.  StorageConfig raises on construction."""


class StorageConfig:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("gas storage is not implemented in this build (SURVEY §8f item 3)")
