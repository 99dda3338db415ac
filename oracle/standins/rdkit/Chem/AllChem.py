def __getattr__(name):
    def _missing(*_a, **_k):
        raise RuntimeError(f"rdkit stand-in: AllChem.{name} is outside the oracle's scope")
    return _missing
