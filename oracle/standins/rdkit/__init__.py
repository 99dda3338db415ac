"""Import stub (rdkit is not installed here); only names touched at import time exist."""
from . import Chem  # noqa: F401


class RDLogger:  # noqa: D101
    @staticmethod
    def DisableLog(*_a, **_k):
        return None
