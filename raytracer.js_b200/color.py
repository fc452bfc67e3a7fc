"""Color (src/physics/color.ts:21-52)."""
from __future__ import annotations


class Color:
    __slots__ = ("r", "g", "b", "a")

    def __init__(self, r: float, g: float, b: float, a: float = 1.0):
        self.r, self.g, self.b, self.a = float(r), float(g), float(b), float(a)

    def __repr__(self):
        return f"Color({self.r}, {self.g}, {self.b}, {self.a})"


def color(r, g, b, a=1.0) -> Color:
    return Color(r, g, b, a)


def clone_color(c: Color) -> Color:
    return Color(c.r, c.g, c.b, c.a)


def mul_color(c1: Color, c2: Color) -> Color:
    return Color(c1.r * c2.r, c1.g * c2.g, c1.b * c2.b, c1.a)
