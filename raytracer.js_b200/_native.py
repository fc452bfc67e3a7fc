"""ctypes binding of librt_b200.so (include/rt_b200.h).  Fails loudly when the CUDA library is
missing or no CUDA device is present: there is no CPU path behind this module."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RT_B200_LIB") or os.path.join(HERE, "librt_b200.so")  # env: tuning builds only

RT_OK, RT_ERR_INVALID, RT_ERR_CUDA, RT_ERR_NO_SCENE, RT_ERR_UNSUPPORTED, RT_ERR_BOUNDS, RT_ERR_TEXTURE = range(7)
RT_ENTITY_SPHERE, RT_ENTITY_BOX = 0, 1
RT_TEXTURE_SOLID, RT_TEXTURE_IMAGE = 0, 1
RT_CAM_REFERENCE_EXTENTS = 1
RT_PRECISION_F32 = 0
RT_PARAM_EXACT_TIES = 4
RT_PRECISION_F64 = 1
RT_RENDER_COUNTERS = 1
RT_B200_ABI_VERSION = 1
RT_TONE_IDENTITY, RT_TONE_STDDEV, RT_TONE_ABSDEV = 0, 1, 2

_dp, _ip, _up, _bp, _qp = (C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint32),
                           C.POINTER(C.c_uint8), C.POINTER(C.c_uint64))


class SceneDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("n_nodes", C.c_uint32),
        ("node_pos", _dp), ("node_size", _dp), ("node_child", _ip), ("node_parent", _ip), ("node_octant", _ip),
        ("node_list_off", _up), ("n_list", C.c_uint32), ("list_entity", _up),
        ("n_entities", C.c_uint32), ("ent_type", _bp), ("ent_pos", _dp), ("ent_extent", _dp),
        ("ent_material", _ip), ("ent_texture", _ip), ("ent_substance", _ip),
        ("n_materials", C.c_uint32), ("mat_response", _bp), ("mat_light", _bp), ("mat_mirror", _bp),
        ("mat_roughness", _dp),
        ("n_textures", C.c_uint32), ("tex_kind", _bp), ("tex_color", _dp), ("tex_width", _ip), ("tex_height", _ip),
        ("tex_loaded", _bp), ("tex_texel_off", _qp), ("n_texels", C.c_uint64), ("texels", _bp),
        ("n_substances", C.c_uint32), ("sub_refractive_index", _dp),
    ]


class CameraDesc(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("fr", C.c_double * 3), ("lf", C.c_double * 3), ("up", C.c_double * 3),
                ("fov_h", C.c_double), ("fov_v", C.c_double), ("width", C.c_uint32), ("height", C.c_uint32),
                ("flags", C.c_uint32), ("_pad", C.c_uint32)]


class Params(C.Structure):
    _fields_ = [("refmax", C.c_int32), ("sky_texture", C.c_int32), ("default_substance", C.c_int32),
                ("_pad0", C.c_int32), ("distance_attenuation_factor", C.c_double), ("n_frames", C.c_uint32),
                ("frame_first", C.c_uint32), ("rng_seed", C.c_double), ("precision", C.c_uint32),
                ("flags", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("paths", "segments", "nodes", "tests", "shades", "confirms",
                                          "texture_errors", "acute_warnings")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class Tone(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("dynamic_range", C.c_uint32), ("min_dynamic", C.c_double), ("max_dynamic", C.c_double)]


class ExposureStats(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("mean", "variance", "absolute_dev", "drange_low", "drange_high")]

    def as_dict(self):
        return {n: float(getattr(self, n)) for n, _ in self._fields_}


EXPORTS = {
    "rt_abi_version": (C.c_uint32, []),
    "rt_create": (C.c_int, [C.c_int32, C.POINTER(C.c_void_p)]),
    "rt_create_multi": (C.c_int, [C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_void_p)]),
    "rt_group_size": (C.c_uint32, [C.c_void_p]),
    "rt_destroy": (None, [C.c_void_p]),
    "rt_last_error": (C.c_char_p, [C.c_void_p]),
    "rt_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rt_scene_upload": (C.c_int, [C.c_void_p, C.POINTER(SceneDesc)]),
    "rt_scene_update": (C.c_int, [C.c_void_p, C.c_uint32, _up, _dp, C.c_uint32]),
    "rt_render": (C.c_int, [C.c_void_p, C.POINTER(CameraDesc), C.POINTER(Params), C.c_uint32, C.c_void_p,
                            C.c_void_p, C.POINTER(Counters)]),
    "rt_render_begin": (C.c_int, [C.c_void_p, C.POINTER(CameraDesc), C.POINTER(Params), C.c_uint32, C.c_void_p, C.c_void_p]),
    "rt_render_end": (C.c_int, [C.c_void_p, C.POINTER(Counters)]),
    "rt_render_device": (C.c_int, [C.c_void_p, C.POINTER(CameraDesc), C.POINTER(Params), C.c_uint32, C.c_void_p,
                                   C.c_void_p]),
    "rt_get_counters": (C.c_int, [C.c_void_p, C.POINTER(Counters)]),
    "rt_camera_directions": (C.c_int, [C.c_void_p, C.POINTER(CameraDesc), C.c_void_p]),
    "rt_present_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(Tone), C.c_void_p]),
    "rt_present_stats": (C.c_int, [C.c_void_p, C.POINTER(ExposureStats)]),
    "rt_present": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(Tone), C.c_void_p,
                             C.POINTER(ExposureStats)]),
    "rt_render_present": (C.c_int, [C.c_void_p, C.POINTER(CameraDesc), C.POINTER(Params), C.c_uint32, C.POINTER(Tone),
                                    C.c_void_p, C.POINTER(ExposureStats), C.POINTER(Counters)]),
    "rt_exposure_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]),
    "rt_synchronize": (C.c_int, [C.c_void_p]),
    "rt_timer_start": (C.c_int, [C.c_void_p]),
    "rt_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "rt_launch_count": (C.c_uint64, [C.c_void_p]),
    "rt_set_profiling": (C.c_int, [C.c_void_p, C.c_int32]),
    "rt_stage_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "rt_tiles_per_rank": (C.c_uint32, [C.c_uint32, C.c_uint32, C.c_uint32]),
    "rt_render_tiles_device": (C.c_int, [C.c_void_p, C.POINTER(CameraDesc), C.POINTER(Params), C.c_uint32,
                                         C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "rt_untile_device": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "rt_render_shard_device": (C.c_int, [C.c_void_p, C.POINTER(CameraDesc), C.POINTER(Params), C.c_uint32,
                                         C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "rt_peer_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]),
    "rt_peer_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rt_peer_open": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "rt_peer_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rt_peer_barrier": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.c_uint32]),
    "rt_tree_build": (C.c_int, [_dp, C.c_double, C.c_uint32, _bp, _dp, _dp, C.c_uint32, C.POINTER(C.c_void_p)]),
    "rt_tree_build_gpu": (C.c_int, [C.c_void_p, _dp, C.c_double, C.c_uint32, _bp, _dp, _dp, C.c_uint32, C.POINTER(C.c_void_p)]),
    "rt_tree_node_count": (C.c_uint32, [C.c_void_p]),
    "rt_tree_export": (None, [C.c_void_p, _dp, _dp, _ip, _ip, _ip, _up, _up]),
    "rt_image_decode": (C.c_int, [C.c_char_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.POINTER(C.c_uint8))]),
    "rt_image_free": (None, [C.POINTER(C.c_uint8)]),
    "rt_tree_free": (None, [C.c_void_p]),
    "rt_fplcg_fill": (None, [C.c_double, C.c_uint64, _dp]),
    "rt_flush_l2": (C.c_int, [C.c_void_p]),
    "rt_host_register": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "rt_host_unregister": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rt_host_map": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
}

_lib = None


class RtError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


def load():
    """dlopen librt_b200.so.  Raises if it has not been built: the product has no other path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python raytracer.js_b200/build.py` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    if L.rt_abi_version() != RT_B200_ABI_VERSION:
        raise ImportError("librt_b200.so ABI version mismatch; rebuild it")
    _lib = L
    return L


def check(ctx, status: int) -> None:
    if status == RT_OK:
        return
    msg = load().rt_last_error(ctx)
    msg = msg.decode() if msg else f"rt status {status}"
    if status == RT_ERR_BOUNDS:
        raise IndexError(msg)  # ExposureBuffer.check_bounds: Error("x or y out of bounds")
    raise RtError(status, msg)
