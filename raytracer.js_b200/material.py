"""Materials (src/material.ts:22-103, src/materials/material_solid.ts:25-44).

Only the static description lives on the host; `alter_ray` (colour *= texture colour at the hit's uv)
runs inside the CUDA kernel (csrc/rt_trace.cuh: trace_path)."""
from __future__ import annotations

import enum


class ResponseType(enum.IntEnum):
    REFLECTION = 0
    TRANSMISSION = 1
    BOTH = 2


class Material:
    """Abstract base (src/material.ts:29-65)."""


class StaticMaterial(Material):
    def __init__(self, response: ResponseType, light_source: bool, mirror: bool, roughness: float):
        self.response = ResponseType(response)
        self.light_source = bool(light_source)
        self.mirror = bool(mirror)
        self.roughness_index = float(roughness)

    def response_type(self, _point=None) -> ResponseType:
        return self.response

    def is_mirror(self, _point=None) -> bool:
        return self.mirror

    def is_light_source(self) -> bool:
        return self.light_source


class SolidMaterial(StaticMaterial):
    """Solid colour material: ray.color *= texture.get_color(entity.map_uv(p))."""


SIMPLE_SMOOTH_MATERIAL = SolidMaterial(ResponseType.REFLECTION, False, True, 0)
SIMPLE_LIGHT_MATERIAL = SolidMaterial(ResponseType.REFLECTION, True, False, 0)
SIMPLE_ROUGH_MATERIAL = SolidMaterial(ResponseType.REFLECTION, False, True, 0.5)
SIMPLE_TRANSPARENT_MATERIAL = SolidMaterial(ResponseType.TRANSMISSION, False, False, 0)
