"""Substance (src/substance.ts:1-11)."""


class Substance:
    __slots__ = ("refractive_index",)

    def __init__(self, index: float):
        self.refractive_index = float(index)


SUBSTANCE_AIR = Substance(1.0)
SUBSTANCE_WATER = Substance(1.333)
SUBSTANCE_GLASS = Substance(1.5)
