"""Axis-aligned space predicates (src/space.ts:24-103)."""
from __future__ import annotations

import enum

from .geometry import Vector


class RangeCoverage(enum.Enum):
    FULL = 0
    OPEN_CLOSE = 1
    CLOSE_OPEN = 2


class Space:
    __slots__ = ("pos", "size")

    def __init__(self, pos: Vector, size: Vector):
        self.pos, self.size = pos, size


class AABB:
    __slots__ = ("pos", "size")

    def __init__(self, pos: Vector, size: float):
        self.pos, self.size = pos, float(size)


def point_in_space(point: Vector, space: Space, coverage=RangeCoverage.CLOSE_OPEN) -> bool:  # :55-83
    for p, s, e in zip(point.v, space.pos.v, space.size.v):
        if coverage is RangeCoverage.CLOSE_OPEN:
            ok = p >= s and p < s + e
        elif coverage is RangeCoverage.OPEN_CLOSE:
            ok = p > s and p <= s + e
        else:
            ok = p >= s and p <= s + e
        if not ok:
            return False
    return True


def space_in_space(interior: Space, exterior: Space) -> bool:  # :85-97
    for ip, isz, ep, esz in zip(interior.pos.v, interior.size.v, exterior.pos.v, exterior.size.v):
        if not (ip >= ep and ip + isz <= ep + esz):
            return False
    return True


def aabb_in_space(aabb: AABB, space: Space) -> bool:  # :99-103
    return space_in_space(Space(aabb.pos, Vector([aabb.size] * 3)), space)
