"""raytracer.js_b200 — B200-native drop-in for the per-pixel ray hot path of Dark565/raytracer.js.

The host side mirrors the reference's Space / Entity / Material / Raytracer API (same names and
argument meaning); `GpuRaytracer.trace_frame()` runs on the GPU through the C ABI of
include/rt_b200.h (librt_b200.so, hand-written sm_100a CUDA).  There is no CPU fallback."""
from . import _native
from .camera import Camera, CameraConfig
from .color import Color, color
from .entity import BasicEntity, BoxEntity, Entity, SphereEntity
from .exposure_buffer import ExposureBuffer
from .flatten import FlatScene, flatten_scene
from .geometry import Vector, point, vector, vector3
from .material import (SIMPLE_LIGHT_MATERIAL, SIMPLE_ROUGH_MATERIAL, SIMPLE_SMOOTH_MATERIAL,
                       SIMPLE_TRANSPARENT_MATERIAL, Material, ResponseType, SolidMaterial, StaticMaterial)
from .octree import Octree, OctreePos
from .octree_entity import (EntitySet, TreeOutsideGrowError, add_entity_to_octree, entity_at_pos,
                            get_covering_node_for_entity, new_entity_octree)
from .octree_space import OctreeDim, index_within_parent, new_subtree, node_at_pos, octant_adj_pos
from .raytracer import GpuRaytracer, RaytracerConfig, camera_desc
from .rng import PRNG, RNG, FpLcg
from .sky import Sky, SkySphere
from .substance import SUBSTANCE_AIR, SUBSTANCE_GLASS, SUBSTANCE_WATER, Substance
from .texture import ImageTexture, SolidTexture, Texture, TextureError
from .view import (GpuView, Screen, ToneMapper, ToneMapper_AbsDevAroundMean, ToneMapper_DRLimited, ToneMapper_Identity,
                   ToneMapper_StdDevAroundMean, View)

Raytracer = GpuRaytracer  # the drop-in name

__all__ = [n for n in dir() if not n.startswith("_")]
