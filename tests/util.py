"""Shared helpers of the test-suite: feed ONE scene to the oracle (checker), to the test-only host
build of the kernel body (tests/hostsim) and to the CUDA library, and compare."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import oracle as orc
import raytracer_js_b200 as rt
from raytracer_js_b200 import _native as N
from raytracer_js_b200.texture import ImageTexture, SolidTexture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTSIM_DIR = os.path.join(ROOT, "tests", "hostsim")
HOSTSIM_LIB = os.path.join(HOSTSIM_DIR, "librt_hostsim.so")


def build_hostsim() -> str:
    src = os.path.join(HOSTSIM_DIR, "rt_hostsim.cpp")
    deps = [src] + [os.path.join(ROOT, "raytracer.js_b200", "csrc", f) for f in ("rt_trace.cuh", "rt_host.h", "rt_common.h")]
    if not os.path.exists(HOSTSIM_LIB) or any(os.path.getmtime(d) > os.path.getmtime(HOSTSIM_LIB) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-pthread", "-shared",
                               "-Wno-unknown-pragmas", "-Wno-comment", "-o", HOSTSIM_LIB, src])
    return HOSTSIM_LIB


_hostsim = None


def hostsim():
    global _hostsim
    if _hostsim is None:
        L = C.CDLL(build_hostsim())
        L.hostsim_render.restype = C.c_int
        L.hostsim_render.argtypes = [C.POINTER(N.SceneDesc), C.POINTER(N.CameraDesc), C.POINTER(N.Params), C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(N.Counters),
                                     C.c_char_p, C.c_int]
        _hostsim = L
    return _hostsim


def make_params(flat, bundle, n_frames=1, frame_first=0, rng_seed=1.0, refmax=None, attenuation=1.0) -> N.Params:
    p = N.Params()
    p.refmax = bundle.refmax if refmax is None else refmax
    p.sky_texture = flat.texture_index(bundle.sky.texture)
    p.default_substance = flat.substance_index(bundle.default_substance)
    p.distance_attenuation_factor = attenuation
    p.n_frames, p.frame_first = n_frames, frame_first
    p.rng_seed = rng_seed
    p.precision = N.RT_PRECISION_F32
    return p


def flat_of(bundle):
    return rt.flatten_scene(bundle.tree, extra_textures=[bundle.sky.texture],
                            extra_substances=[bundle.default_substance])


def insertion_ids(flat, bundle, flat_ids: np.ndarray) -> np.ndarray:
    """Map the flattener's entity ids (tree traversal order) to insertion indices (the oracle's ids)."""
    order = {id(e): i for i, e in enumerate(bundle.entities)}
    lut = np.array([order[id(e)] for e in flat.entities] + [-1], np.int32)
    return lut[flat_ids]


def hostsim_render(flat, camera, params, n_threads=8, reference_extents=False, rgb=None, tile_rank=0, tile_world=1,
                   pipeline=False, shard_frame_layout=False):
    """pipeline=False: ray by ray with work counters (rt_render_kernel<true>'s body); pipeline=True: the
    packet primary stage + bounce stage (the default GPU path), counters['confirms'] = queued pixels."""
    W, H = camera.conf.screen_w, camera.conf.screen_h
    if tile_world > 1 and not shard_frame_layout:
        from raytracer_js_b200.parallel import tiles_per_rank
        tpr = tiles_per_rank(W, H, tile_world)
        rgb = np.zeros((tpr * 256, 3), np.float32)
        ids = np.full(tpr * 256, -1, np.int32)
    else:
        if rgb is None:
            rgb = np.zeros((H, W, 3), np.float32)
        ids = np.full((H, W), -1, np.int32)
    cnt = N.Counters()
    err = C.create_string_buffer(512)
    d = flat.desc()
    cd = rt.camera_desc(camera, reference_extents)
    st = hostsim().hostsim_render(C.byref(d), C.byref(cd), C.byref(params), n_threads, tile_rank, tile_world,
                                  (1 if pipeline else 0) | (2 if shard_frame_layout else 0), rgb.ctypes.data, ids.ctypes.data, C.byref(cnt), err, 512)
    if st == N.RT_ERR_BOUNDS:
        raise IndexError(err.value.decode())
    if st != 0:
        raise RuntimeError(f"hostsim status {st}: {err.value.decode()}")
    return rgb, ids, cnt.as_dict()


def oracle_scene(flat, bundle, max_in_depth=16, max_out_depth=0):
    """Rebuild the scene inside the oracle from the entity list in insertion order.  The oracle builds
    its OWN octree (its restatement of add_entity_to_octree); table indices follow `flat`."""
    root = bundle.tree
    s = orc.Scene(root.id.pos.v, root.id.size)
    for m in flat.materials:
        s.add_material(int(m.response), m.light_source, m.mirror, m.roughness_index)
    for t in flat.textures:
        if isinstance(t, SolidTexture):
            s.add_texture_solid(t.color.r, t.color.g, t.color.b, t.color.a)
        else:
            fb = t.fallback_color
            s.add_texture_image(t.width, t.height, t.image_data, (fb.r, fb.g, fb.b, fb.a))
    for sub in flat.substances:
        s.add_substance(sub.refractive_index)
    ents = bundle.entities
    if ents:
        typ = [0 if isinstance(e, rt.SphereEntity) else 1 for e in ents]
        pos = [e.get_pos().v for e in ents]
        ext = [e.get_diameter() if isinstance(e, rt.SphereEntity) else e.get_size() for e in ents]
        s.add_entities(typ, pos, ext, [flat.material_index(e.get_material()) for e in ents],
                       [flat.texture_index(e.get_texture()) for e in ents],
                       [flat.substance_index(e.get_substance()) for e in ents], max_in_depth, max_out_depth)
    return s


def oracle_scene_flat(fb, max_in_depth=16):
    """The oracle's scene for a bulk-built FlatBundle (no Python entity objects): tables in the flat scene's
    order, entities in array order = insertion order = entity id on both sides.  The oracle builds its OWN
    octree from them with its restatement of add_entity_to_octree."""
    flat, a = fb.flat, fb.flat.arrays
    root_pos, root_size = a["node_pos"][0], float(a["node_size"][0])
    s = orc.Scene(root_pos, root_size)
    for m in flat.materials:
        s.add_material(int(m.response), m.light_source, m.mirror, m.roughness_index)
    for t in flat.textures:
        if isinstance(t, SolidTexture):
            s.add_texture_solid(t.color.r, t.color.g, t.color.b, t.color.a)
        else:
            fbk = t.fallback_color
            s.add_texture_image(t.width, t.height, t.image_data, (fbk.r, fbk.g, fbk.b, fbk.a))
    for sub in flat.substances:
        s.add_substance(sub.refractive_index)
    s.add_entities(a["ent_type"], a["ent_pos"], a["ent_extent"], a["ent_material"], a["ent_texture"],
                   a["ent_substance"], max_in_depth, 0)
    return s


def flat_params(fb, n_frames=1, frame_first=0, rng_seed=1.0) -> N.Params:
    p = N.Params()
    p.refmax, p.sky_texture, p.default_substance = fb.refmax, fb.sky_texture, fb.default_substance
    p.distance_attenuation_factor = 1.0
    p.n_frames, p.frame_first, p.rng_seed = n_frames, frame_first, rng_seed
    p.precision = N.RT_PRECISION_F32
    return p


def gpu_render_flat(fb, width, height, n_frames=1, pos=None, yaw=30.0, pitch=0.0):
    """A FlatBundle through the C ABI (rt_create / rt_scene_upload / rt_render with host buffers): the default
    pipeline.  Returns (rgb [H,W,3], first-hit entity ids [H,W], launches)."""
    from raytracer_js_b200 import scenes
    lib = N.load()
    ctx = C.c_void_p()
    N.check(None, lib.rt_create(-1, C.byref(ctx)))
    try:
        d = fb.flat.desc()
        N.check(ctx, lib.rt_scene_upload(ctx, C.byref(d)))
        cam = scenes.bench_camera(width, height, pos or scenes.BENCH_CAMERA_POS, yaw, pitch)
        cd = rt.camera_desc(cam)
        prm = flat_params(fb, n_frames)
        rgb = np.zeros((height, width, 3), np.float32)
        ids = np.full((height, width), -9, np.int32)
        N.check(ctx, lib.rt_render(ctx, C.byref(cd), C.byref(prm), 0, rgb.ctypes.data, ids.ctypes.data, None))
        launches = int(lib.rt_launch_count(ctx))
    finally:
        lib.rt_destroy(ctx)
    return rgb, ids, launches


def oracle_crop(oscene, fb, width, height, crop, n_frames=1, pos=None, yaw=30.0, pitch=0.0, n_threads=None):
    """The oracle on the crop (x, y, w, h) of the full frame: same camera, same per-pixel seeds (they depend on
    the pixel's place in the FULL frame).  Returns (rgb, ids) of the crop and the totals."""
    import math
    from raytracer_js_b200 import scenes
    ocam = orc.Camera(math.pi / 2, math.pi / 2, width, height, pos or scenes.BENCH_CAMERA_POS, pitch,
                      math.pi / 180 * yaw, vertical_locked=True)
    prm = flat_params(fb, n_frames)
    rgb, ids, _, tot = orc.render(oscene, ocam, refmax=prm.refmax, sky_texture=prm.sky_texture,
                                  default_substance=prm.default_substance, distance_attenuation_factor=1.0,
                                  fixed_extents=True, n_frames=n_frames, frame_first=0, rng_mode=1, seed=prm.rng_seed,
                                  n_threads=n_threads or min(32, os.cpu_count() or 8), crop=crop)
    x, y, w, h = crop
    return rgb[y:y + h, x:x + w], ids[y:y + h, x:x + w], tot


def oracle_render(oscene, ocam, flat, bundle, params: N.Params, fixed_extents=True, n_threads=8, rgb=None,
                  want_counters=False):
    return orc.render(oscene, ocam, refmax=params.refmax, sky_texture=params.sky_texture,
                      default_substance=params.default_substance,
                      distance_attenuation_factor=params.distance_attenuation_factor, fixed_extents=fixed_extents,
                      n_frames=params.n_frames, frame_first=params.frame_first, rng_mode=1, seed=params.rng_seed,
                      n_threads=n_threads, rgb=rgb, want_counters=want_counters)


def compare(rgb_a, ids_a, rgb_ref, ids_ref):
    """The parity gate of BASELINE.json: ids equal on >= 99.99 % of pixels; on id-equal pixels RGB within
    1/255 per channel (relative to max(1,|ref|) so that over-range light pixels are judged fairly)."""
    n = ids_ref.size
    same = ids_a == ids_ref
    frac = float(same.sum()) / n
    diff = np.abs(rgb_a.astype(np.float64) - rgb_ref.astype(np.float64))
    tol = (1.0 / 255.0) * np.maximum(1.0, np.abs(rgb_ref.astype(np.float64)))
    bad_rgb = (diff > tol).any(axis=-1) & same
    return dict(id_match=frac, id_mismatch=int(n - same.sum()), rgb_bad=int(bad_rgb.sum()),
                rgb_max_abs=float(diff[same].max()) if same.any() else 0.0)


def _is_dyadic(v, levels=24):
    m = v * (1 << levels)
    return m == int(m)


def classify_outliers(rgb_a, ids_a, rgb_ref, ids_ref, cam_pos=None, ocam=None, fixed_extents=True, image_textures=False,
                      offset=(0, 0)):
    """outliers are reported by kind, not budgeted silently.  Returns
    {"dyadic_tie": [...], "texel_edge": [...], "unexplained": [...]} with (x, y) pixels.
      dyadic_tie  id mismatch on a camera ray that runs INSIDE a cell-boundary plane of the octree: a direction
                  component is exactly 0 and the camera's coordinate on that axis is dyadic (the demo pose sits on
                  the root's centre planes, src/main.ts:364).  The reference's half-open cells put such a ray in the
                  upper cell only; entities of the lower cell that touch the plane are hit by the float64 formula
                  all the same (tangent / face contact), and the float32 search, conservative by design, finds them.
      texel_edge  colour mismatch on a pixel that reads an image texture (entity or sky) where the two libms'
                  atan2 differ in the last bit and (u*w)<<0 lands in the neighbouring texel; only possible when the
                  scene has image textures.
      unexplained anything else: a failure.
    ocam: the oracle's Camera of the FULL frame (directions come from its generator); offset: where the compared
    arrays sit in that frame (crops)."""
    same = ids_a == ids_ref
    diff = np.abs(rgb_a.astype(np.float64) - rgb_ref.astype(np.float64))
    tol = (1.0 / 255.0) * np.maximum(1.0, np.abs(rgb_ref.astype(np.float64)))
    bad_rgb = (diff > tol).any(axis=-1) & same
    out = {"dyadic_tie": [], "texel_edge": [], "unexplained": []}
    dirs = None
    if ocam is not None and cam_pos is not None and (~same).any():
        xy, d, _ = ocam.dirs(fixed_extents=fixed_extents)
        look = np.full((ocam.screen_h, ocam.screen_w), -1, np.int64)
        look[xy[:, 1], xy[:, 0]] = np.arange(len(xy))
        dirs = (look, d)
    for y, x in zip(*np.nonzero(~same)):
        kind = "unexplained"
        fx, fy = int(x) + offset[0], int(y) + offset[1]
        if dirs is not None and dirs[0][fy, fx] >= 0:
            d = dirs[1][dirs[0][fy, fx]]
            if any(d[k] == 0.0 and _is_dyadic(cam_pos[k]) for k in range(3)):
                kind = "dyadic_tie"
        out[kind].append((fx, fy))
    for y, x in zip(*np.nonzero(bad_rgb)):
        out["texel_edge" if image_textures else "unexplained"].append((int(x) + offset[0], int(y) + offset[1]))
    return out


def assert_parity(rgb_a, ids_a, rgb_ref, ids_ref, cam_pos=None, ocam=None, fixed_extents=True, image_textures=False,
                  exact_ids=False):
    """The gate, with nothing budgeted silently: ids equal on >= 99.99 % of the pixels (all of them with
    exact_ids), every id mismatch a classified dyadic tie, no colour outside 1/255 on id-equal pixels except
    classified texel-edge flips (scenes with image textures only)."""
    res = compare(rgb_a, ids_a, rgb_ref, ids_ref)
    kinds = classify_outliers(rgb_a, ids_a, rgb_ref, ids_ref, cam_pos=cam_pos, ocam=ocam, fixed_extents=fixed_extents,
                              image_textures=image_textures)
    assert res["id_match"] >= 0.9999 and not kinds["unexplained"], (res, kinds)
    assert res["rgb_bad"] == len(kinds["texel_edge"]), (res, kinds)
    if exact_ids:
        assert res["id_mismatch"] == 0, (res, kinds)
    return res, kinds
