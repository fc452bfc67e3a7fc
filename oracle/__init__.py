"""oracle — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes binding of oracle/oracle.cpp, the double-precision CPU restatement of the
reference's per-pixel ray hot path (see the header of oracle.cpp for what it
follows and how it is pinned).  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this package; the
product (raytracer.js_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile liboracle.so with g++ (oracle/Makefile).  Building the checker is not using it."""
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Config(C.Structure):
    _fields_ = [("refmax", C.c_int32), ("sky_texture", C.c_int32), ("default_substance", C.c_int32),
                ("distance_attenuation_factor", C.c_double)]


class Totals(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("nodes", "tests", "shades", "segments", "cell_steps", "would_throw",
                                          "within_tests", "texture_errors", "acute_warnings", "paths")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _up(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def _v3(p):
    a = np.ascontiguousarray(p, dtype=np.float64)
    assert a.shape == (3,)
    return a


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    vp, dp, ip, up = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)
    i32, f64 = C.c_int32, C.c_double
    sig = {
        "orc_scene_new": (vp, [dp, f64]),
        "orc_scene_free": (None, [vp]),
        "orc_scene_error": (C.c_char_p, [vp]),
        "orc_add_material": (i32, [vp, i32, i32, i32, f64]),
        "orc_add_substance": (i32, [vp, f64]),
        "orc_add_texture_solid": (i32, [vp, f64, f64, f64, f64]),
        "orc_add_texture_image": (i32, [vp, i32, i32, C.c_void_p, dp]),
        "orc_add_entity": (i32, [vp, i32, dp, f64, i32, i32, i32, i32, i32]),
        "orc_add_entities": (i32, [vp, i32, C.c_void_p, dp, dp, ip, ip, ip, i32, i32]),
        "orc_move_entity": (i32, [vp, i32, dp, i32, i32]),
        "orc_new_subtree": (i32, [vp, ip, i32, i32]),
        "orc_flat_counts": (None, [vp, up, up]),
        "orc_flat_export": (None, [vp, dp, dp, ip, ip, ip, up, up, ip]),
        "orc_entity_node": (i32, [vp, i32]),
        "orc_node_at_pos": (None, [vp, dp, ip, ip]),
        "orc_entity_at_pos": (i32, [vp, dp]),
        "orc_walk": (i32, [vp, dp, dp, i32, i32, ip, i32]),
        "orc_collision": (i32, [vp, i32, dp, dp, dp, dp]),
        "orc_uv_map_sphere": (None, [dp, dp]),
        "orc_fplcg": (None, [f64, i32, dp]),
        "orc_box_line": (i32, [dp, dp, dp, dp, dp, ip]),
        "orc_camera_new": (vp, [f64, f64, i32, i32, f64, f64, i32, dp, i32, f64, i32, f64]),
        "orc_camera_free": (None, [vp]),
        "orc_camera_rotate_h": (None, [vp, f64]),
        "orc_camera_rotate_v": (None, [vp, f64]),
        "orc_camera_rotate_h_step": (None, [vp, i32]),
        "orc_camera_rotate_v_step": (None, [vp, i32]),
        "orc_camera_set_pos": (None, [vp, dp]),
        "orc_camera_get_basis": (None, [vp, dp]),
        "orc_camera_dirs": (C.c_int64, [vp, i32, ip, dp, C.c_int64]),
        "orc_render": (i32, [vp, vp, C.POINTER(Config), i32, i32, i32, i32, f64, i32, i32, i32, i32, i32,
                             C.POINTER(C.c_float), ip, up, C.POINTER(Totals)]),
        "orc_build_demo_scene": (i32, [vp, f64, i32, ip]),
        "orc_entity_count": (i32, [vp]),
        "orc_entity_get": (None, [vp, i32, ip, dp, dp, ip, ip, ip]),
        "orc_material_count": (i32, [vp]),
        "orc_material_get": (None, [vp, i32, ip, ip, ip, dp]),
        "orc_texture_count": (i32, [vp]),
        "orc_texture_get": (None, [vp, i32, ip, dp, ip, ip, ip]),
        "orc_substance_count": (i32, [vp]),
        "orc_substance_get": (f64, [vp, i32]),
        "orc_exposure_stats": (None, [C.POINTER(C.c_float), C.c_int64, dp]),
        "orc_dynamic_range": (None, [i32, i32, f64, f64, dp, dp]),
        "orc_discretize": (None, [C.POINTER(C.c_float), C.c_int64, f64, f64, C.c_void_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


@dataclass
class FlatTree:
    node_pos: np.ndarray
    node_size: np.ndarray
    node_child: np.ndarray
    node_parent: np.ndarray
    node_octant: np.ndarray
    node_list_off: np.ndarray
    list_entity: np.ndarray
    root_index: int


class Scene:
    """The oracle's own pointer octree + entity/material/texture/substance tables."""

    def __init__(self, root_pos=(0.0, 0.0, 0.0), root_size=1.0):
        self._L = lib()
        self._h = self._L.orc_scene_new(_dp(_v3(root_pos)), float(root_size))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_scene_free(self._h)
            self._h = None

    # tables
    def add_material(self, response, light, mirror, roughness):
        return self._L.orc_add_material(self._h, int(response), int(bool(light)), int(bool(mirror)), float(roughness))

    def add_substance(self, n):
        return self._L.orc_add_substance(self._h, float(n))

    def add_texture_solid(self, r, g, b, a=1.0):
        return self._L.orc_add_texture_solid(self._h, r, g, b, a)

    def add_texture_image(self, width, height, rgb8, fallback=(0.0, 0.0, 0.0, 1.0)):
        fb = np.ascontiguousarray(fallback, dtype=np.float64)
        if rgb8 is None:
            return self._L.orc_add_texture_image(self._h, width, height, None, _dp(fb))
        px = np.ascontiguousarray(rgb8, dtype=np.uint8)
        assert px.size == width * height * 3
        return self._L.orc_add_texture_image(self._h, width, height, px.ctypes.data, _dp(fb))

    # add_entity_to_octree
    def add_entity(self, type_, pos, extent, material, texture, substance, max_in_depth=16, max_out_depth=0):
        r = self._L.orc_add_entity(self._h, int(type_), _dp(_v3(pos)), float(extent), int(material), int(texture),
                                   int(substance), int(max_in_depth), int(max_out_depth))
        if r < 0:
            raise RuntimeError(self._L.orc_scene_error(self._h).decode())
        return r

    def add_entities(self, type_, pos, extent, material, texture, substance, max_in_depth=16, max_out_depth=0):
        n = len(extent)
        t = np.ascontiguousarray(type_, dtype=np.uint8)
        p = np.ascontiguousarray(pos, dtype=np.float64).reshape(n, 3)
        e = np.ascontiguousarray(extent, dtype=np.float64)
        m = np.ascontiguousarray(material, dtype=np.int32)
        x = np.ascontiguousarray(texture, dtype=np.int32)
        s = np.ascontiguousarray(substance, dtype=np.int32)
        r = self._L.orc_add_entities(self._h, n, t.ctypes.data, _dp(p), _dp(e), _ip(m), _ip(x), _ip(s),
                                     int(max_in_depth), int(max_out_depth))
        if r < 0:
            raise RuntimeError(f"entity {-1 - r}: " + self._L.orc_scene_error(self._h).decode())
        return r

    def move_entity(self, eid, pos, max_in_depth=16, max_out_depth=0):
        """_set_pos + add_entity_to_octree again: out of its node's Set, to the END of the new node's."""
        r = self._L.orc_move_entity(self._h, int(eid), _dp(_v3(pos)), int(max_in_depth), int(max_out_depth))
        if r == -1:
            raise RuntimeError(self._L.orc_scene_error(self._h).decode())
        if r != 0:
            raise IndexError(f"entity {eid}")

    def new_subtree(self, path, n):
        p = np.ascontiguousarray(path, dtype=np.int32)
        return self._L.orc_new_subtree(self._h, _ip(p), len(p), int(n))

    def build_demo_scene(self, seed=0.0, n_entities=16):
        sky = C.c_int32(-1)
        n = self._L.orc_build_demo_scene(self._h, float(seed), int(n_entities), C.byref(sky))
        if n < 0:
            raise RuntimeError(self._L.orc_scene_error(self._h).decode())
        return n, sky.value

    # read-back
    def flat(self) -> FlatTree:
        nn, nl = C.c_uint32(), C.c_uint32()
        self._L.orc_flat_counts(self._h, C.byref(nn), C.byref(nl))
        n, l = nn.value, nl.value
        f = FlatTree(np.zeros((n, 3)), np.zeros(n), np.zeros((n, 8), np.int32), np.zeros(n, np.int32),
                     np.zeros(n, np.int32), np.zeros(n + 1, np.uint32), np.zeros(max(l, 1), np.uint32), 0)
        ri = C.c_int32()
        self._L.orc_flat_export(self._h, _dp(f.node_pos), _dp(f.node_size), _ip(f.node_child), _ip(f.node_parent),
                                _ip(f.node_octant), _up(f.node_list_off), _up(f.list_entity), C.byref(ri))
        f.list_entity = f.list_entity[:l]
        f.root_index = ri.value
        return f

    def entity_node(self, eid):
        return self._L.orc_entity_node(self._h, eid)

    def entities(self):
        n = self._L.orc_entity_count(self._h)
        out = []
        t, m, x, s = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        e = C.c_double()
        p = np.zeros(3)
        for i in range(n):
            self._L.orc_entity_get(self._h, i, C.byref(t), _dp(p), C.byref(e), C.byref(m), C.byref(x), C.byref(s))
            out.append(dict(type=t.value, pos=tuple(p.tolist()), extent=e.value, material=m.value,
                            texture=x.value, substance=s.value))
        return out

    def materials(self):
        out = []
        r, l, m = C.c_int32(), C.c_int32(), C.c_int32()
        ro = C.c_double()
        for i in range(self._L.orc_material_count(self._h)):
            self._L.orc_material_get(self._h, i, C.byref(r), C.byref(l), C.byref(m), C.byref(ro))
            out.append(dict(response=r.value, light=bool(l.value), mirror=bool(m.value), roughness=ro.value))
        return out

    def textures(self):
        out = []
        k, w, h, ld = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        col = np.zeros(4)
        for i in range(self._L.orc_texture_count(self._h)):
            self._L.orc_texture_get(self._h, i, C.byref(k), _dp(col), C.byref(w), C.byref(h), C.byref(ld))
            out.append(dict(kind=k.value, color=tuple(col.tolist()), width=w.value, height=h.value,
                            loaded=bool(ld.value)))
        return out

    def substances(self):
        return [self._L.orc_substance_get(self._h, i) for i in range(self._L.orc_substance_count(self._h))]

    # probes
    def node_at_pos(self, p):
        n, o = C.c_int32(), C.c_int32()
        self._L.orc_node_at_pos(self._h, _dp(_v3(p)), C.byref(n), C.byref(o))
        return n.value, o.value

    def entity_at_pos(self, p):
        return self._L.orc_entity_at_pos(self._h, _dp(_v3(p)))

    def walk(self, pos, direction, include_undefined=False, use_start_node=False, max_stops=4096):
        stops = np.zeros((max_stops, 3), np.int32)
        n = self._L.orc_walk(self._h, _dp(_v3(pos)), _dp(_v3(direction)), int(include_undefined),
                             int(use_start_node), _ip(stops), max_stops)
        assert n <= max_stops
        return stops[:n].copy()

    def collision(self, eid, pos, direction):
        pt, nm = np.zeros(3), np.zeros(3)
        hit = self._L.orc_collision(self._h, eid, _dp(_v3(pos)), _dp(_v3(direction)), _dp(pt), _dp(nm))
        return (pt, nm) if hit else None


class Camera:
    """src/view/camera.ts Camera(conf, init_pos, init_v_angle, init_h_angle)."""

    def __init__(self, fov_v, fov_h, screen_w, screen_h, pos, init_v_angle=None, init_h_angle=None,
                 rot_v=np.pi / 30, rot_h=np.pi / 30, vertical_locked=False):
        self._L = lib()
        self.screen_w, self.screen_h = int(screen_w), int(screen_h)
        self._h = self._L.orc_camera_new(float(fov_v), float(fov_h), int(screen_w), int(screen_h), float(rot_v),
                                         float(rot_h), int(vertical_locked), _dp(_v3(pos)),
                                         int(init_v_angle is not None), float(init_v_angle or 0.0),
                                         int(init_h_angle is not None), float(init_h_angle or 0.0))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_camera_free(self._h)
            self._h = None

    def rotate_h(self, a):
        self._L.orc_camera_rotate_h(self._h, float(a))

    def rotate_v(self, a):
        self._L.orc_camera_rotate_v(self._h, float(a))

    def rotate_h_step(self, n):
        self._L.orc_camera_rotate_h_step(self._h, int(n))

    def rotate_v_step(self, n):
        self._L.orc_camera_rotate_v_step(self._h, int(n))

    def set_pos(self, p):
        self._L.orc_camera_set_pos(self._h, _dp(_v3(p)))

    def basis(self):
        b = np.zeros(12)
        self._L.orc_camera_get_basis(self._h, _dp(b))
        return dict(pos=b[0:3].copy(), fr=b[3:6].copy(), lf=b[6:9].copy(), up=b[9:12].copy())

    def dirs(self, fixed_extents=False):
        n = self.screen_w * self.screen_h
        xy = np.zeros((n, 2), np.int32)
        d = np.zeros((n, 3))
        got = self._L.orc_camera_dirs(self._h, int(fixed_extents), _ip(xy), _dp(d), n)
        return xy[:min(got, n)], d[:min(got, n)], got


def render(scene: Scene, camera: Camera, *, refmax, sky_texture, default_substance, distance_attenuation_factor=1.0,
           fixed_extents=False, n_frames=1, frame_first=0, rng_mode=1, seed=1.0, n_threads=1, crop=None,
           rgb=None, want_counters=False):
    """Raytracer.trace_frame() x n_frames into an ExposureBuffer.  Returns (rgb f32 [H,W,3], ids i32 [H,W],
    counters u32 [H,W,4] or None, totals dict)."""
    L = lib()
    W, H = camera.screen_w, camera.screen_h
    if rgb is None:
        rgb = np.zeros((H, W, 3), np.float32)
    assert rgb.dtype == np.float32 and rgb.shape == (H, W, 3) and rgb.flags.c_contiguous
    ids = np.full((H, W), -1, np.int32)
    counters = np.zeros((H, W, 4), np.uint32) if want_counters else None
    cfg = Config(int(refmax), int(sky_texture), int(default_substance), float(distance_attenuation_factor))
    tot = Totals()
    cx, cy, cw, ch = crop if crop else (0, 0, 0, 0)
    rc = L.orc_render(scene._h, camera._h, C.byref(cfg), int(fixed_extents), int(n_frames), int(frame_first),
                      int(rng_mode), float(seed), int(n_threads), cx, cy, cw, ch,
                      rgb.ctypes.data_as(C.POINTER(C.c_float)), _ip(ids),
                      _up(counters) if counters is not None else None, C.byref(tot))
    if rc == -1:
        raise IndexError("x or y out of bounds")  # ExposureBuffer.check_bounds, exposure_buffer.ts:181-186
    if rc != 0:
        raise ValueError(f"orc_render: bad arguments ({rc})")
    return rgb, ids, counters, tot.as_dict()


def fplcg(seed, n):
    out = np.zeros(n)
    lib().orc_fplcg(float(seed), n, _dp(out))
    return out


def uv_map_sphere(d):
    uv = np.zeros(2)
    lib().orc_uv_map_sphere(_dp(_v3(d)), _dp(uv))
    return uv


def box_line(center, size, pos, direction):
    u = np.zeros(2)
    f = np.zeros(2, np.int32)
    n = lib().orc_box_line(_dp(_v3(center)), _dp(_v3(size)), _dp(_v3(pos)), _dp(_v3(direction)), _dp(u), _ip(f))
    return n, u, f


# ---- View.draw_ebuffer (src/view/view.ts:34-38): exposure statistics, tone mapper, 8-bit screen
TONE_IDENTITY, TONE_STDDEV, TONE_ABSDEV = 0, 1, 2


def exposure_stats(pixels):
    """ExposureBuffer.get_mean / get_variance / get_absolute_dev -> (mean, variance, absdev)."""
    px = np.ascontiguousarray(pixels, dtype=np.float32).reshape(-1)
    out = np.zeros(3)
    lib().orc_exposure_stats(px.ctypes.data_as(C.POINTER(C.c_float)), px.size // 3, _dp(out))
    return tuple(float(v) for v in out)


def dynamic_range(kind, dynamic_range_stops, min_dynamic, max_dynamic, stats):
    """ToneMapper.get_dynamic_range -> (drange_min, drange_max)."""
    st = np.ascontiguousarray(stats, dtype=np.float64)
    out = np.zeros(2)
    lib().orc_dynamic_range(int(kind), int(dynamic_range_stops), float(min_dynamic), float(max_dynamic), _dp(st), _dp(out))
    return float(out[0]), float(out[1])


def discretize(pixels, drange_low, drange_high):
    """ExposureBuffer.discretize_to_screen into a CanvasScreen image -> uint8 [n_pixels, 4] (RGBA)."""
    px = np.ascontiguousarray(pixels, dtype=np.float32).reshape(-1)
    out = np.zeros((px.size // 3, 4), np.uint8)
    lib().orc_discretize(px.ctypes.data_as(C.POINTER(C.c_float)), px.size // 3, float(drange_low), float(drange_high),
                         out.ctypes.data)
    return out
