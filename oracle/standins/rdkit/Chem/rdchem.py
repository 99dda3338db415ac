from . import _BondType as BondType  # noqa: F401
