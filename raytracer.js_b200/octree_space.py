"""Geometry on octrees (src/octree_space.ts:25-136).  The OctreeWalker of the reference (:159-408) is
what the CUDA kernel implements (csrc/rt_trace.cuh: walk_and_scan); there is no host walker here."""
from __future__ import annotations

from typing import Optional

from . import space as _space
from .geometry import Vector, add, js_int32, scale, sub, vector
from .octree import Octree, OctreePos


class OctreeDim:
    """{pos, size} of a node's cube (src/octree_space.ts:30-33)."""
    __slots__ = ("pos", "size")

    def __init__(self, pos: Vector, size: float):
        self.pos, self.size = pos, float(size)


def octant_adj_pos(octree: Octree, pos: Vector) -> int:  # :41-50
    dim = octree.id
    h = dim.size / 2
    px = int(pos.v[0] >= dim.pos.v[0] + h)
    py = int(pos.v[1] >= dim.pos.v[1] + h)
    pz = int(pos.v[2] >= dim.pos.v[2] + h)
    return (pz << 2) | (py << 1) | px


def node_at_pos(octree: Octree, pos: Optional[Vector], start_from_current=False,
                range_cover=_space.RangeCoverage.CLOSE_OPEN) -> Optional[OctreePos]:  # :61-93
    dim = octree.id
    if pos is None:
        return None
    cur_node = octree if start_from_current else octree.get_root()
    if not _space.point_in_space(pos, _space.Space(dim.pos, scale(vector(1, 1, 1), dim.size)), range_cover):
        return None
    cur_index = 0
    next_pos = list(dim.pos.v)
    next_size = dim.size
    next_node: Optional[Octree] = cur_node
    p = pos.v
    while isinstance(next_node, Octree):
        k = 2 / next_size
        i0 = js_int32((p[0] - next_pos[0]) * k)
        i1 = js_int32((p[1] - next_pos[1]) * k)
        i2 = js_int32((p[2] - next_pos[2]) * k)
        cur_node = next_node
        cur_index = (i2 << 2) + (i1 << 1) + (i0 << 0)
        next_node = cur_node.get(cur_index)
        next_size /= 2
        next_pos[0] += i0 * next_size
        next_pos[1] += i1 * next_size
        next_pos[2] += i2 * next_size
    return OctreePos(cur_node, cur_index)


def new_subtree(tree: Octree, n: int, allow_replace=False) -> Octree:  # :95-108
    if not allow_replace and tree.get(n) is not None:
        raise ValueError("Child already defined")
    half = tree.id.size / 2
    subdim = OctreeDim(add(tree.id.pos, scale(vector(n & 1, (n >> 1) & 1, (n >> 2) & 1), half)), half)
    subtree = Octree(subdim, tree)
    tree.set(n, subtree)
    return subtree


def index_within_parent(child: Octree) -> Optional[int]:  # :113-125
    if child.index_within_parent is not None:
        return child.index_within_parent
    parent = child.parent
    if parent is None:
        return None
    ind = scale(sub(child.id.pos, parent.id.pos), 2 / parent.id.size)
    return (js_int32(ind.v[2]) << 2) + (js_int32(ind.v[1]) << 1) + (js_int32(ind.v[0]) << 0)


def dim_relative_to_parent(parent: Octree, n: int) -> OctreeDim:  # :127-136
    ph = parent.id.size / 2
    return OctreeDim(add(parent.id.pos, scale(vector((n >> 0) & 1, (n >> 1) & 1, (n >> 2) & 1), ph)), ph)
