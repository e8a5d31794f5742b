"""Mutation type codes (mut_types.py:4-12 of the reference)."""
from enum import Enum


class MutType(Enum):
    SN = 1
    IN = 2
    DE = 3
    DU = 4
    IV = 5
    TL = 6
    TLI = 7


# canonical device order = the reference's ARGS dict order (rmt.py:443-450, :91-94)
DEVICE_ORDER = (MutType.SN, MutType.IN, MutType.DE, MutType.IV, MutType.DU, MutType.TL, MutType.TLI)
DEVICE_CODE = {t: i for i, t in enumerate(DEVICE_ORDER)}
