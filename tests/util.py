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


def oracle_camera(camera):
    """An oracle Camera with the very same pose (basis copied through rotate calls is not possible, so
    the oracle camera is built from the same constructor arguments by the caller; this helper checks)."""
    raise NotImplementedError


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
