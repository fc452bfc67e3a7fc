"""Synthetic scene recipes of the measurement plan: harness code, deterministic through
FpLcg so that the same scene can be rebuilt anywhere (tests feed the same entity list to the oracle)."""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .camera import Camera, CameraConfig
from .color import Color
from .entity import BoxEntity, Entity, SphereEntity
from .geometry import point
from .material import ResponseType, SolidMaterial
from .octree import Octree
from .octree_entity import add_entity_to_octree, new_entity_octree
from .octree_space import OctreeDim
from .rng import FpLcg
from .sky import SkySphere
from .substance import SUBSTANCE_AIR, SUBSTANCE_GLASS, SUBSTANCE_WATER, Substance
from .texture import ImageTexture, SolidTexture, Texture

BENCH_CAMERA_POS = (0.5013, 0.4987, 0.5021)  # strictly inside the root, off dyadic planes


@dataclass
class SceneBundle:
    tree: Octree
    entities: List[Entity]          # insertion order == entity id of the oracle
    sky: SkySphere
    default_substance: Substance
    refmax: int
    materials: List[SolidMaterial] = field(default_factory=list)
    description: str = ""


def bench_camera(width: int, height: int, pos=BENCH_CAMERA_POS, yaw_deg: float = 30.0, pitch: float = 0.0) -> Camera:
    """The demo pose (src/main.ts:353-366): 90 x 90 degrees, yaw 30 degrees."""
    conf = CameraConfig(fov_v=math.pi * 0.5, fov_h=math.pi * 0.5, screen_w=width, screen_h=height,
                        rot_v=math.pi / 30, rot_h=math.pi / 30, flags={"vertical_locked": True})
    return Camera(conf, point(*pos), pitch, math.pi / 180 * yaw_deg)


def _new_root() -> Octree:
    return new_entity_octree(OctreeDim(point(0, 0, 0), 1.0), None)


DIFFUSE = dict(response=ResponseType.REFLECTION, light=False, mirror=False, roughness=0.0)


def random_spheres(n: int, dmin: float = 0.002, dmax: float = 0.006, seed: float = 42.0, mix: str = "diffuse",
                   box_fraction: float = 0.0, textures: Optional[List[Texture]] = None,
                   max_in_depth: int = 16) -> SceneBundle:
    """Config 1/2/4 generator.  Per entity the draws are, in order: d, cx, cy, cz, [kind], material pick,
    texture colour r,g,b (or texture pick).  Sphere i: d = dmin + u*(dmax-dmin), centre = d/2 + u*(1-d)
    per axis (AABB inside the unit root, as max_out_depth: 0 requires).
    mix 'diffuse': every material is SolidMaterial(REFLECTION, light=False, mirror=False, 0) (config 1);
    mix 'mirrors': 70% smooth mirror, 15% diffuse, 10% rough mirror 0.5, 5% lights x5 (config 2/4)."""
    rng = FpLcg(seed)
    tree = _new_root()
    if mix == "diffuse":
        mats = [SolidMaterial(ResponseType.REFLECTION, False, False, 0.0)]
        cum = [1.0]
    elif mix == "mirrors":
        mats = [SolidMaterial(ResponseType.REFLECTION, False, True, 0.0),
                SolidMaterial(ResponseType.REFLECTION, False, False, 0.0),
                SolidMaterial(ResponseType.REFLECTION, False, True, 0.5),
                SolidMaterial(ResponseType.REFLECTION, True, False, 0.0)]
        cum = [0.70, 0.85, 0.95, 1.0]
    else:
        raise ValueError(mix)
    entities: List[Entity] = []
    cfg = {"max_in_depth": max_in_depth, "max_out_depth": 0}
    for _ in range(n):
        d = dmin + rng.next() * (dmax - dmin)
        c = [d / 2 + rng.next() * (1 - d) for _ in range(3)]
        is_box = box_fraction > 0 and rng.next() < box_fraction
        mi = 0
        if len(mats) > 1:
            u = rng.next()
            while mi < len(cum) - 1 and u > cum[mi]:
                mi += 1
        mat = mats[mi]
        if textures and not mat.light_source and not is_box:
            tex = textures[int(rng.next() * len(textures))]
        else:
            k = 5.0 if mat.light_source else 1.0
            tex = SolidTexture(Color(rng.next() * k, rng.next() * k, rng.next() * k, 1.0))
        cls = BoxEntity if is_box else SphereEntity
        e = cls(None, mat, tex, SUBSTANCE_AIR, point(*c), d)
        add_entity_to_octree(tree, e, cfg)
        entities.append(e)
    sky = SkySphere(SolidTexture(Color(0.2, 0.2, 0.7, 1.0)))  # the demo's fallback sky colour (main.ts:378)
    return SceneBundle(tree, entities, sky, SUBSTANCE_AIR, refmax=1 if mix == "diffuse" else 4, materials=mats,
                       description=f"{n} random spheres d in [{dmin},{dmax}], seed {seed}, mix {mix}")


def checker_texture(width: int, height: int, seed: int = 1) -> ImageTexture:
    """Array-backed image texture: random-colour checker with per-texel noise (no file IO)."""
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:height, 0:width]
    cells = ((xx // 16) + (yy // 16)) % 2
    a, b = rs.randint(32, 255, 3), rs.randint(32, 255, 3)
    img = np.where(cells[..., None] == 0, a, b).astype(np.int32) + rs.randint(-16, 16, (height, width, 3))
    return ImageTexture(np.clip(img, 0, 255).astype(np.uint8), Color(0, 0, 0, 1))


DEMO_CAMERA_POS = (0.5, 0.5, 0.5)  # reset_pos, src/main.ts:364


def _js_int32(x: float) -> int:
    """`x << 0` for the non-negative doubles this generator meets."""
    return int(x)


def demo_scene(seed: float = 0.0, n_entities: int = 16) -> SceneBundle:
    """BASELINE config 0: the reference's own demo scene (src/main.ts:97-147,389-408) as a headless run
    builds it: `generate_some_aligned_entities(otree, prng, 16, 0.0, [1,1,1,1], textures)` then the
    enclosing rough unit box added last to the root; image assets are not in the tree, so every texture is
    a SolidTexture and the sky is the ImageTexture's fallback colour (0.2,0.2,0.7) (:378).  The default
    seed of the page is 0 (`Number(null) ?? 42`, :149-152,350).  refmax 4 (:48)."""
    from .geometry import Vector, length, scale
    from .material import (SIMPLE_LIGHT_MATERIAL, SIMPLE_ROUGH_MATERIAL, SIMPLE_SMOOTH_MATERIAL,
                           SIMPLE_TRANSPARENT_MATERIAL)
    rng = FpLcg(seed)
    tree = _new_root()
    substances = [SUBSTANCE_AIR, SUBSTANCE_WATER, SUBSTANCE_GLASS]
    materials = [SIMPLE_LIGHT_MATERIAL, SIMPLE_ROUGH_MATERIAL, SIMPLE_SMOOTH_MATERIAL, SIMPLE_TRANSPARENT_MATERIAL]
    classes = [SphereEntity, BoxEntity]
    entities: List[Entity] = []
    existing = []
    for _ in range(n_entities):
        level = 1 + math.floor(rng.next() * 7)
        n_quant = 1 << level
        size = 1 / n_quant
        q = [_js_int32(rng.next() * n_quant) for _ in range(3)]
        x, y, z = (qq * size + size / 2 for qq in q)
        if q in existing:
            continue
        existing.append(q)
        cls = classes[_js_int32(rng.next() * 2)]
        substance = substances[_js_int32(rng.next() * 3)]
        # get_random_element_with_weights (:77-94) with weights [1,1,1,1]: the one-argument comparator of
        # :84 leaves the index array in order, so this is a cumulative pick in index order
        wn = 1.0 * (1 / (((0 + 1.0) + 1.0) + 1.0 + 1.0))
        rnd = rng.next()
        weight, mi = 0.0, 3
        for k in range(3):
            weight += wn
            if rnd <= weight:
                mi = k
                break
        material = materials[mi]
        intensity = 5.0 if material is SIMPLE_LIGHT_MATERIAL else 1.0
        rnd = rng.next()  # get_random_texture (:67-76) with probability 0: an image only for rnd == 0
        if rnd <= 0.0:
            rng.next()
            tex = SolidTexture(Color(0.0, 0.0, 0.0, 1.0))  # the unloaded ImageTexture's fallback (:56)
        else:
            c = Vector((rng.next(), rng.next(), rng.next()))
            c = scale(scale(c, 1.0 / length(c)), intensity)
            tex = SolidTexture(Color(c.v[0], c.v[1], c.v[2], 1.0))
        e = cls(None, material, tex, substance, point(x, y, z), size)
        add_entity_to_octree(tree, e, {"max_in_depth": 16, "max_out_depth": 0})
        entities.append(e)
    box = BoxEntity(None, SIMPLE_ROUGH_MATERIAL, SolidTexture(Color(1.0, 1.0, 1.0, 1.0)), SUBSTANCE_AIR,
                    point(0.5, 0.5, 0.5), 1.0)
    add_entity_to_octree(tree, box, {"max_in_depth": 1, "max_out_depth": 0})
    entities.append(box)
    sky = SkySphere(SolidTexture(Color(0.2, 0.2, 0.7, 1.0)))
    return SceneBundle(tree, entities, sky, SUBSTANCE_AIR, refmax=4, materials=materials,
                       description=f"demo scene of src/main.ts, FpLcg({seed}), {n_entities} attempts")


@dataclass
class FlatBundle:
    """A big scene that only exists as flat arrays (no Python entity objects, no pointer tree)."""
    flat: "FlatScene"
    sky_texture: int
    default_substance: int
    refmax: int
    n_entities: int
    description: str = ""


def fplcg_draws(seed: float, n: int) -> np.ndarray:
    """n consecutive FpLcg(seed).next() values (native restatement in librt_b200: rt_fplcg_fill)."""
    from . import _native as N
    out = np.zeros(int(n), np.float64)
    N.load().rt_fplcg_fill(float(seed), int(n), out.ctypes.data_as(N._dp))
    return out


def random_spheres_flat(n: int, dmin: float = 0.002, dmax: float = 0.006, seed: float = 42.0, mix: str = "diffuse",
                        box_fraction: float = 0.0, textures: Optional[List[Texture]] = None,
                        max_in_depth: int = 16) -> FlatBundle:
    """The same scenes as random_spheres() (same draws in the same order, so the same entities), built in
    bulk: draws from the native FpLcg, octree from the native restatement of add_entity_to_octree
    (flatten.flat_from_arrays).  This is how the 100 k - 1 M entity configs are built in seconds."""
    from .flatten import flat_from_arrays
    if mix == "diffuse":
        mats = [SolidMaterial(ResponseType.REFLECTION, False, False, 0.0)]
        cum = np.array([1.0])
    elif mix == "mirrors":
        mats = [SolidMaterial(ResponseType.REFLECTION, False, True, 0.0),
                SolidMaterial(ResponseType.REFLECTION, False, False, 0.0),
                SolidMaterial(ResponseType.REFLECTION, False, True, 0.5),
                SolidMaterial(ResponseType.REFLECTION, True, False, 0.0)]
        cum = np.array([0.70, 0.85, 0.95, 1.0])
    else:
        raise ValueError(mix)
    has_kind, has_pick = box_fraction > 0, len(mats) > 1
    sky = SolidTexture(Color(0.2, 0.2, 0.7, 1.0))
    if not textures:
        k = 4 + int(has_kind) + int(has_pick) + 3
        u = fplcg_draws(seed, n * k).reshape(n, k)
        d = dmin + u[:, 0] * (dmax - dmin)
        pos = d[:, None] / 2 + u[:, 1:4] * (1 - d[:, None])
        col = 4
        is_box = np.zeros(n, bool)
        if has_kind:
            is_box = u[:, col] < box_fraction
            col += 1
        mi = np.zeros(n, np.int64)
        if has_pick:
            mi = np.minimum(np.searchsorted(cum, u[:, col], side="left"), len(cum) - 1)  # first i with u <= cum[i]
            col += 1
        kcol = np.where(np.array([m.light_source for m in mats])[mi], 5.0, 1.0)
        rgb = u[:, col:col + 3] * kcol[:, None]
        tex_objs = [sky] + [SolidTexture(Color(float(r), float(g), float(b), 1.0)) for r, g, b in rgb]
        ent_tex = np.arange(1, n + 1)
    else:
        # variable number of draws per entity: sequential (medium-sized textured scenes only)
        rng = FpLcg(seed)
        d, pos, is_box, mi, ent_tex = np.zeros(n), np.zeros((n, 3)), np.zeros(n, bool), np.zeros(n, np.int64), np.zeros(n, np.int64)
        tex_objs = [sky] + list(textures)
        for i in range(n):
            d[i] = dmin + rng.next() * (dmax - dmin)
            pos[i] = [d[i] / 2 + rng.next() * (1 - d[i]) for _ in range(3)]
            is_box[i] = has_kind and rng.next() < box_fraction
            if has_pick:
                u = rng.next()
                while mi[i] < len(cum) - 1 and u > cum[mi[i]]:
                    mi[i] += 1
            if not mats[mi[i]].light_source and not is_box[i]:
                ent_tex[i] = 1 + int(rng.next() * len(textures))
            else:
                kk = 5.0 if mats[mi[i]].light_source else 1.0
                tex_objs.append(SolidTexture(Color(rng.next() * kk, rng.next() * kk, rng.next() * kk, 1.0)))
                ent_tex[i] = len(tex_objs) - 1
    flat = flat_from_arrays(is_box.astype(np.uint8), pos, d, mi, ent_tex, np.zeros(n, np.int32), mats, tex_objs,
                            [SUBSTANCE_AIR], max_in_depth=max_in_depth)
    return FlatBundle(flat, 0, 0, 1 if mix == "diffuse" else 4, n,
                      f"{n} random spheres d in [{dmin},{dmax}], seed {seed}, mix {mix} (bulk build)")


# The non-headline BASELINE.json configs: frame, samples per pixel and scene recipe.
#   c2: 1920x1080, 16 spp, 100 k spheres d in [0.002,0.006], 70% mirrors / 15% diffuse / 10% rough / 5% lights, refmax 4
#   c3: 3840x2160, 4 spp, 1 M entities (10% boxes) d in [0.0005,0.002], 4 image textures 1024x512, refmax 4
#   c4: 7680x4320, 64 spp, 1 M spheres d in [0.0005,0.002], config-2 material mix, refmax 4
BASELINE_CONFIGS = {
    "c2": dict(w=1920, h=1080, spp=16, n=100000, dmin=0.002, dmax=0.006, mix="mirrors", box=0.0, tex=False),
    "c3": dict(w=3840, h=2160, spp=4, n=1000000, dmin=0.0005, dmax=0.002, mix="mirrors", box=0.1, tex=True),
    "c4": dict(w=7680, h=4320, spp=64, n=1000000, dmin=0.0005, dmax=0.002, mix="mirrors", box=0.0, tex=False),
}


def build_config(cfg) -> FlatBundle:
    """The scene of one BASELINE_CONFIGS entry, bulk-built (random_spheres_flat).  Textured configs: the
    geometry and materials come from the bulk generator, the textured entities (spheres that are not lights)
    pick one of 4 shared 1024x512 image textures with a second stream of draws, FpLcg(43)."""
    fb = random_spheres_flat(cfg["n"], cfg["dmin"], cfg["dmax"], 42.0, cfg["mix"], cfg["box"])
    if cfg["tex"]:
        texs = [checker_texture(1024, 512, seed=s) for s in (1, 2, 3, 4)]
        a = fb.flat.arrays
        base = len(fb.flat.textures)
        for t in texs:
            fb.flat.texture_index(t)
        pick = (fplcg_draws(43.0, cfg["n"]) * 4).astype(np.int32)
        light = a["mat_light"][a["ent_material"]].astype(bool)
        textured = (~light) & (a["ent_type"] == 0)
        a["ent_texture"] = np.where(textured, base + pick, a["ent_texture"]).astype(np.int32)
        fb.flat.finalize_tables()
    return fb
