"""Scene entities (src/entity.ts:38-101, src/entities/entity_basic.ts, entity_sphere.ts, entity_box.ts).

The host objects carry geometry and the material/texture/substance references and implement what the
octree builder needs (get_aabb, is_within).  collision_info / map_uv run inside the CUDA kernel
(csrc/rt_trace.cuh: candidate, exact_sphere, exact_box, uv_map_sphere)."""
from __future__ import annotations

from typing import Optional, Tuple

from . import space as _space
from .geometry import Vector, dot, point, scale, sub, vector3
from .material import Material
from .substance import Substance
from .texture import Texture


class Entity:
    def __init__(self, octree=None, substance: Optional[Substance] = None):
        self._octree = octree
        self.substance = substance

    def set_octree(self, tree, keep_in_current=False) -> None:  # src/entity.ts:50-56
        if not keep_in_current and self._octree is not None:
            self._octree.value.set.pop(self, None)
        self._octree = tree
        tree.value.set[self] = None

    @property
    def octree(self):
        return self._octree

    def get_substance(self) -> Optional[Substance]:
        return self.substance

    def set_substance(self, s: Optional[Substance]) -> Optional[Substance]:
        old, self.substance = self.substance, s
        return old

    def collision_info(self, ray):
        raise NotImplementedError("collision_info runs on the GPU (librt_b200); there is no CPU path")


class BasicEntity(Entity):
    def __init__(self, entity_otree, material: Material, texture: Texture, substance: Optional[Substance], pos: Vector):
        super().__init__(entity_otree, substance)
        self.material, self.texture, self.pos = material, texture, pos

    def get_pos(self) -> Vector:
        return self.pos

    def _set_pos(self, p: Vector) -> Vector:
        old, self.pos = self.pos, p
        return old

    def get_material(self) -> Material:
        return self.material

    def set_material(self, m: Material) -> Material:
        old, self.material = self.material, m
        return old

    def get_texture(self) -> Texture:
        return self.texture

    def set_texture(self, t: Texture) -> Texture:
        old, self.texture = self.texture, t
        return old


class SphereEntity(BasicEntity):
    def __init__(self, entity_otree, material, texture, substance, pos: Vector, diameter: float):
        super().__init__(entity_otree, material, texture, substance, pos)
        self.diameter = float(diameter)

    def get_diameter(self) -> float:
        return self.diameter

    def set_diameter(self, d: float) -> float:
        old, self.diameter = self.diameter, float(d)
        return old

    def is_within(self, p: Vector) -> bool:  # entity_sphere.ts:63-66
        dist = sub(p, self.get_pos())
        return dot(dist, dist) <= self.diameter * self.diameter / 4

    def get_aabb(self) -> Tuple[Vector, float]:  # :90-96
        d = self.diameter
        return sub(self.get_pos(), scale(point(d, d, d), 0.5)), d


class BoxEntity(BasicEntity):
    def __init__(self, entity_otree, material, texture, substance, pos: Vector, size: float):
        super().__init__(entity_otree, material, texture, substance, pos)
        self.size = float(size)

    def get_size(self) -> float:
        return self.size

    def set_size(self, size: float) -> float:
        old, self.size = self.size, float(size)
        return old

    def is_within(self, p: Vector) -> bool:  # entity_box.ts:47-52 (pos used as the min corner, as the reference does)
        return _space.point_in_space(p, _space.Space(self.get_pos(), vector3(self.size, self.size, self.size)))

    def get_aabb(self) -> Tuple[Vector, float]:  # :75-82
        h = self.size / 2
        return sub(self.get_pos(), point(h, h, h)), self.size
