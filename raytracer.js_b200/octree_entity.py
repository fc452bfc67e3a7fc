"""Integration of entities with the octree (src/octree_entity.ts:32-202): the host-side scene build
that feeds the flattener.  Restated, not copied; the oracle holds an independent restatement and
tests/test_host_build.py checks that both produce identical trees."""
from __future__ import annotations

import math
from typing import Optional

from .entity import Entity
from .geometry import Vector, add, js_int32, scale, sub, vector
from .octree import Octree, OctreePos
from .octree_space import OctreeDim, node_at_pos
from .space import AABB, Space, aabb_in_space


class EntitySet:
    """Insertion-ordered set of entities (a JS Set; here a dict used as an ordered set)."""

    def __init__(self, octree_pos: Optional[OctreePos] = None):
        self._octree_pos = octree_pos
        self._set: dict = {}

    @property
    def set(self) -> dict:
        return self._set

    @property
    def octree_pos(self):
        return self._octree_pos


def new_entity_octree(dim: OctreeDim, parent: Optional[Octree], entity_set: Optional[EntitySet] = None) -> Octree:
    return Octree(dim, parent, entity_set if entity_set is not None else EntitySet())


def _cube(dim: OctreeDim) -> Space:
    return Space(dim.pos, scale(vector(1, 1, 1), dim.size))


def get_covering_node_for_entity(tree: Octree, entity: Entity) -> Optional[Octree]:  # :60-79
    apos, asize = entity.get_aabb()
    aabb = AABB(apos, asize)
    deepest = node_at_pos(tree, apos)
    if deepest is None:
        return None
    cur = deepest.tree
    while cur is not None:
        if aabb_in_space(aabb, _cube(cur.id)):
            break
        cur = cur.parent
    return cur


def _extend_tree_inside_to_fit_up_to_depth(root: Octree, node: Octree, aabb: AABB, max_depth: int) -> Octree:  # :92-114
    cur_depth = node.get_relative_level(root)
    cur = node
    while cur_depth < max_depth:
        q = scale(sub(aabb.pos, cur.id.pos), 2.0 / cur.id.size)
        x, y, z = (js_int32(c) for c in q.v)
        spos = add(cur.id.pos, scale(vector(x, y, z), cur.id.size / 2))
        ssize = cur.id.size / 2
        if not aabb_in_space(aabb, Space(spos, scale(vector(1, 1, 1), ssize))):
            break
        new_tree = new_entity_octree(OctreeDim(spos, ssize), cur)
        cur.set((z << 2) | (y << 1) | (x << 0), new_tree)
        cur = new_tree
        cur_depth += 1
    return cur


class TreeOutsideGrowError(Exception):  # :116-123
    def __init__(self, abs_root: Octree, msg: str):
        super().__init__(msg)
        self.abs_root = abs_root


def _extend_tree_outside_to_fit_up_to_depth(root: Octree, node: Octree, aabb: AABB, max_depth: int) -> Octree:  # :125-171
    if node.parent is not None:
        raise ValueError("'node' must be the absolute root (i.e no parent)")
    cur_depth = root.get_relative_level(node)
    cur = node
    fit = False
    while cur_depth < max_depth:
        al = scale(sub(aabb.pos, cur.id.pos), 1.0 / cur.id.size)
        al = vector(*(max(min(float(math.floor(c)), 0.0), -1.0) for c in al.v))
        parent_pos = add(cur.id.pos, scale(al, cur.id.size))
        parent_size = cur.id.size * 2
        idx = (js_int32(-al.v[2]) << 2) | (js_int32(-al.v[1]) << 1) | (js_int32(-al.v[0]) << 0)
        new_parent = new_entity_octree(OctreeDim(parent_pos, parent_size), None)
        new_parent.set(idx, cur)
        cur.parent = new_parent
        cur = new_parent
        if aabb_in_space(aabb, Space(parent_pos, scale(vector(1, 1, 1), parent_size))):
            fit = True
            break
        cur_depth += 1
    if not fit:
        raise TreeOutsideGrowError(cur, "The tree outside-depth limit exceeded")
    return cur


def add_entity_to_octree(tree: Octree, entity: Entity, config: dict) -> Octree:  # :174-188
    """config: {'max_in_depth': int, 'max_out_depth': int} (AddEntityToOctreeFlags)."""
    apos, asize = entity.get_aabb()
    aabb = AABB(apos, asize)
    fitting = get_covering_node_for_entity(tree, entity)
    if fitting is None:
        fitting = _extend_tree_outside_to_fit_up_to_depth(tree, tree.get_root(), aabb, config["max_out_depth"])
    fitting = _extend_tree_inside_to_fit_up_to_depth(tree, fitting, aabb, config["max_in_depth"])
    entity.set_octree(fitting)
    return fitting


def entity_at_pos(tree: Octree, p: Vector) -> Optional[Entity]:  # :191-202
    np_ = node_at_pos(tree, p)
    cur = np_.tree if np_ is not None else None
    while cur is not None:
        for entity in cur.value.set:
            if entity.is_within(p):
                return entity
        cur = cur.parent
    return None
