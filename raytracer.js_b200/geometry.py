"""Vector / Point values of the host API (src/math/geometry.ts, src/math/vector.ts): `{v: number[]}`."""
from __future__ import annotations

import math
from typing import Iterable, List


class Vector:
    """An object with a `.v` list of floats, like the reference's `{v: number[]}`."""
    __slots__ = ("v",)

    def __init__(self, values: Iterable[float]):
        self.v: List[float] = [float(x) for x in values]

    def __repr__(self):
        return f"Vector({self.v})"

    def __iter__(self):
        return iter(self.v)

    def __len__(self):
        return len(self.v)


Point = Vector


def vector(*values: float) -> Vector:  # src/math/vector.ts:300-302
    return Vector(values)


def vector3(x: float, y: float, z: float) -> Vector:  # :308-310
    return Vector((x, y, z))


def point(*values: float) -> Vector:  # src/math/geometry.ts
    return Vector(values)


def as_vector(p) -> Vector:
    return p if isinstance(p, Vector) else Vector(p)


def clone(v: Vector) -> Vector:  # :49-51
    return Vector(v.v)


def dot(a: Vector, b: Vector) -> float:  # :76-84 (sum starts at 0, left to right)
    s = 0.0
    for x, y in zip(a.v, b.v):
        s += x * y
    return s


def add(a: Vector, b: Vector) -> Vector:
    return Vector(x + y for x, y in zip(a.v, b.v))


def sub(a: Vector, b: Vector) -> Vector:
    return Vector(x - y for x, y in zip(a.v, b.v))


def scale(a: Vector, s: float) -> Vector:
    return Vector(x * s for x in a.v)


def length(a: Vector) -> float:
    return math.sqrt(dot(a, a))


def normalize(a: Vector) -> Vector:
    return scale(a, 1.0 / length(a))


def cross(a: Vector, b: Vector) -> Vector:  # :86-92
    return Vector((a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2],
                   a.v[0] * b.v[1] - a.v[1] * b.v[0]))


def rotate_vectors(base_x: Vector, base_y: Vector, rot_vec: Vector):  # :318-323
    return (add(scale(base_x, rot_vec.v[0]), scale(base_y, rot_vec.v[1])),
            add(scale(base_x, -rot_vec.v[1]), scale(base_y, rot_vec.v[0])))


def js_int32(x: float) -> int:
    """ECMAScript ToInt32 (what `x << 0` does to a double)."""
    if x != x or x in (math.inf, -math.inf):
        return 0
    t = int(x)  # truncation toward zero
    t &= 0xFFFFFFFF
    return t - (1 << 32) if t & 0x80000000 else t
