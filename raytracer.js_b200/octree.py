"""Generic dynamic octree node (src/octree.ts:25-126)."""
from __future__ import annotations

from typing import Any, Generic, List, Optional, TypeVar

T = TypeVar("T")
ID = TypeVar("ID")


class OctreeRootError(Exception):
    pass


class Octree(Generic[T, ID]):
    def __init__(self, id: ID, parent: "Optional[Octree]" = None, value: Optional[T] = None):
        self.id = id
        self.parent = parent
        self.index_within_parent: Optional[int] = None  # "used for optimization"; never assigned by the reference
        self.value = value
        self._invalidated = False
        self._nodes: List[Optional[Octree]] = [None] * 8

    @staticmethod
    def is_at_bounds(n: int) -> bool:
        return 0 <= n <= 7

    @staticmethod
    def check_bounds(n: int) -> None:
        if not Octree.is_at_bounds(n):
            raise IndexError("Node index out of range (0..7)")

    def get(self, n: int) -> "Optional[Octree]":
        Octree.check_bounds(n)
        return self._nodes[n]

    def subtree(self, n: int) -> "Octree":
        node = self.get(n)
        if not isinstance(node, Octree):
            raise TypeError("Node is not Octree")
        return node

    def set(self, n: int, value: "Optional[Octree]", keep_old_valid=False, invalidate_recurse=False):
        Octree.check_bounds(n)
        old = self._nodes[n]
        if old is value:
            return old
        self._nodes[n] = value
        if not keep_old_valid and isinstance(old, Octree):
            old.invalidate(bool(invalidate_recurse))
        return old

    def invalidate(self, recurse=False) -> None:
        self._invalidated = True
        if recurse:
            for c in self._nodes:
                if isinstance(c, Octree):
                    c.invalidate(recurse)

    def is_invalid(self) -> bool:
        return self._invalidated

    def get_root(self) -> "Octree":
        cur = self
        while cur.parent is not None:
            cur = cur.parent
        return cur

    def get_level(self) -> int:
        level, cur = 0, self
        while cur.parent is not None:
            cur = cur.parent
            level += 1
        return level

    def get_relative_level(self, root: "Octree") -> int:
        return self.get_level() - root.get_level()


class OctreePos:
    """Specific position within the octree (src/octree.ts:129-132)."""
    __slots__ = ("tree", "octant")

    def __init__(self, tree: Octree, octant: Optional[int]):
        self.tree = tree
        self.octant = octant

    def __eq__(self, other):
        return isinstance(other, OctreePos) and self.tree is other.tree and self.octant == other.octant

    def __repr__(self):
        return f"OctreePos(tree={id(self.tree):#x}, octant={self.octant})"
