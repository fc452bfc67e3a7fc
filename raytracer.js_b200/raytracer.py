"""GpuRaytracer — the drop-in for `Raytracer` (src/raytracer.ts:281-339): same constructor arguments
and public members (config, set_camera, set_ebuffer, trace_frame, tree, rng), with the per-pixel loop
of trace_frame() executed by librt_b200 on the GPU.  There is no CPU fallback: constructing it
without the CUDA library or a CUDA device raises."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _native as N
from .camera import Camera
from .exposure_buffer import ExposureBuffer
from .flatten import FlatScene, flatten_scene
from .octree import Octree
from .rng import PRNG, RNG
from .sky import Sky, SkySphere
from .substance import Substance


class RaytracerConfig:
    """RaytracerConfig (src/raytracer.ts:33-43)."""

    def __init__(self, refmax: int, sky: Sky, default_substance: Substance, distance_attenuation_factor: float):
        self.refmax = int(refmax)
        self.sky = sky
        self.default_substance = default_substance
        self.distance_attenuation_factor = float(distance_attenuation_factor)

    def copy(self) -> "RaytracerConfig":
        return RaytracerConfig(self.refmax, self.sky, self.default_substance, self.distance_attenuation_factor)


def camera_desc(camera: Camera, reference_extents: bool = False) -> N.CameraDesc:
    cd = N.CameraDesc()
    cd.pos[:] = camera.get_pos().v
    cd.fr[:] = camera.norm_fr.v
    cd.lf[:] = camera.norm_lf.v
    cd.up[:] = camera.norm_up.v
    cd.fov_h, cd.fov_v = camera.conf.fov_h, camera.conf.fov_v
    cd.width, cd.height = camera.conf.screen_w, camera.conf.screen_h
    cd.flags = N.RT_CAM_REFERENCE_EXTENTS if reference_extents else 0
    return cd


class GpuRaytracer:
    def __init__(self, config: RaytracerConfig, otree: Octree, camera: Camera, ebuffer: ExposureBuffer, rng: RNG,
                 device: int = -1, reference_extents: bool = False, n_gpus: int = 1, devices=None,
                 precision: int = N.RT_PRECISION_F32, exact_ties: bool = False):
        """precision: RT_PRECISION_F32 (float search + float64 confirmation, the default) or RT_PRECISION_F64 (the
        reference's walker in float64, ray by ray: slower, but for octrees of any depth).
        exact_ties: RT_PARAM_EXACT_TIES - for scenes built on a dyadic lattice (entities that fill or touch their cells
        exactly): hits whose ray only touches the entity's cell are searched again by the float64 walker.  On by itself
        whenever the camera stands on a cell plane of the octree.
        n_gpus > 1 (or an explicit `devices` list): one process drives all of them behind the same calls
        (rt_create_multi): the scene is packed once and replicated device to device, trace_frame() shards the
        frame into interleaved tiles and every GPU stores its tiles straight into the ExposureBuffer."""
        self._lib = N.load()
        self._ctx = C.c_void_p()
        if devices is not None or n_gpus > 1:
            devs = list(devices) if devices is not None else list(range(n_gpus))
            arr = (C.c_int32 * len(devs))(*devs)
            N.check(None, self._lib.rt_create_multi(len(devs), arr, C.byref(self._ctx)))
        else:
            N.check(None, self._lib.rt_create(int(device), C.byref(self._ctx)))
        self.config = config.copy()
        self._otree = otree
        self._camera = camera
        self._ebuffer = ebuffer
        self._rng = rng
        self.reference_extents = bool(reference_extents)
        self.precision = int(precision)
        self.exact_ties = bool(exact_ties)
        self.last_counters: Optional[dict] = None
        self.last_first_ids: Optional[np.ndarray] = None
        self.flat: Optional[FlatScene] = None
        self.refresh_scene()

    # -- reference API ------------------------------------------------------------------------
    def set_camera(self, camera: Camera) -> None:
        self._camera = camera

    def set_ebuffer(self, ebuffer: ExposureBuffer) -> None:
        self._ebuffer = ebuffer

    @property
    def tree(self) -> Octree:
        return self._otree

    @property
    def rng(self) -> RNG:
        return self._rng

    def trace_frame(self, n_frames: int = 1, want_ids: bool = False, want_counters: bool = False) -> None:
        """One trace_frame() of the reference (or `n_frames` of them with next_frame() in between, the
        reference's way of taking several samples per pixel) into the ExposureBuffer."""
        eb, cam = self._ebuffer, self._camera
        if (eb.width, eb.height) != (cam.conf.screen_w, cam.conf.screen_h):
            raise IndexError("x or y out of bounds")  # ExposureBuffer.check_bounds
        if eb.max_exposure_frames >= 0:  # never blend more frames than the ExposureBuffer will count (next_frame)
            n_frames = max(1, min(n_frames, eb.max_exposure_frames - eb.current_frame))
        p = self.params(n_frames=n_frames, frame_first=eb.current_frame)
        cd = camera_desc(cam, self.reference_extents)
        ids = np.empty(eb.width * eb.height, np.int32) if want_ids else None
        cnt = N.Counters() if want_counters else None
        st = self._lib.rt_render(self._ctx, C.byref(cd), C.byref(p), 0, eb.pixels.ctypes.data,
                                 ids.ctypes.data if ids is not None else None,
                                 C.byref(cnt) if cnt is not None else None)
        N.check(self._ctx, st)
        for _ in range(n_frames - 1):
            eb.next_frame()
        self.last_first_ids = ids.reshape(eb.height, eb.width) if ids is not None else None
        self.last_counters = cnt.as_dict() if cnt is not None else None

    # -- beyond the reference API ---------------------------------------------------------------
    def refresh_scene(self) -> None:
        """Re-flatten the octree and upload it (call after entities were added or moved)."""
        sky = self.config.sky
        if not isinstance(sky, SkySphere):
            raise TypeError(f"unsupported Sky subclass {type(sky).__name__}")
        self.flat = flatten_scene(self._otree, extra_textures=[sky.texture],
                                  extra_substances=[self.config.default_substance])
        d = self.flat.desc()
        N.check(self._ctx, self._lib.rt_scene_upload(self._ctx, C.byref(d)))

    def move_entities(self, entities, positions, max_in_depth: int = 16) -> None:
        """Dynamic scenes: move entities the reference's way - `_set_pos(p)` then
        `add_entity_to_octree(tree, e, ...)` again, i.e. out of the node's Set and to the END of the Set of the node
        that now covers it - both in the host octree and, through rt_scene_update, in the library's copy of the flat
        scene: only the moved entities cross the boundary, the tree is not flattened again."""
        from .octree_entity import add_entity_to_octree
        index = {id(e): i for i, e in enumerate(self.flat.entities)}
        ids = np.array([index[id(e)] for e in entities], np.uint32)
        pos = np.ascontiguousarray([list(p.v) if hasattr(p, "v") else list(p) for p in positions], np.float64).reshape(-1, 3)
        st = self._lib.rt_scene_update(self._ctx, len(ids), ids.ctypes.data_as(N._up), pos.ctypes.data_as(N._dp), int(max_in_depth))
        N.check(self._ctx, st)
        from .geometry import point
        for e, p in zip(entities, pos):  # the host mirror follows (after the library accepted the moves)
            e._set_pos(point(*p))
            add_entity_to_octree(self._otree, e, {"max_in_depth": max_in_depth, "max_out_depth": 0})

    def params(self, n_frames: int = 1, frame_first: int = 0) -> N.Params:
        p = N.Params()
        p.refmax = self.config.refmax
        p.sky_texture = self.flat.texture_index(self.config.sky.texture)
        p.default_substance = self.flat.substance_index(self.config.default_substance)
        p.distance_attenuation_factor = self.config.distance_attenuation_factor
        p.n_frames, p.frame_first = int(n_frames), int(frame_first)
        # the harness RNG policy (rt_b200.h): per-pixel reseed through the public PRNG.seed()
        p.rng_seed = float(getattr(self._rng, "seed_value", 1.0)) if isinstance(self._rng, PRNG) else 1.0
        p.precision = self.precision
        p.flags = N.RT_PARAM_EXACT_TIES if self.exact_ties else 0
        return p

    @property
    def ctx(self) -> C.c_void_p:
        return self._ctx

    @property
    def lib(self):
        return self._lib

    def close(self) -> None:
        if getattr(self, "_ctx", None):
            self._lib.rt_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
