class _BondType:
    SINGLE, DOUBLE, TRIPLE, AROMATIC = "SINGLE", "DOUBLE", "TRIPLE", "AROMATIC"


class rdchem:  # noqa: N801
    BondType = _BondType


BondType = _BondType


class AllChem:  # noqa: D101
    pass


class RWMol:  # noqa: D101
    def __init__(self, *a, **k):
        raise RuntimeError("rdkit stand-in: chemistry is outside the oracle's scope")


def __getattr__(name):
    def _missing(*_a, **_k):
        raise RuntimeError(f"rdkit stand-in: Chem.{name} is outside the oracle's scope")
    return _missing
