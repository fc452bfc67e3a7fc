"""Camera (src/view/camera.ts:27-251): pose, rotation and movement API on the host.  Per-pixel ray
generation (get_dir_for_each_pixel) runs on the GPU from the pose; the generator is kept here for API
completeness and for small host-side checks."""
from __future__ import annotations

import math
from typing import Iterator, Optional, Tuple

from .geometry import Vector, add, clone, cross, normalize, rotate_vectors, scale, vector3


class CameraConfig:
    def __init__(self, fov_v: float, fov_h: float, screen_w: int, screen_h: int, rot_v: float = math.pi / 30,
                 rot_h: float = math.pi / 30, flags: Optional[dict] = None):
        self.fov_v, self.fov_h = float(fov_v), float(fov_h)
        self.screen_w, self.screen_h = int(screen_w), int(screen_h)
        self.rot_v, self.rot_h = float(rot_v), float(rot_h)
        self.flags = dict(flags or {})

    def copy(self) -> "CameraConfig":
        return CameraConfig(self.fov_v, self.fov_h, self.screen_w, self.screen_h, self.rot_v, self.rot_h, self.flags)


def _v2(x, y):
    return Vector((x, y))


class Camera:
    def __init__(self, conf: CameraConfig, init_pos: Vector, init_v_angle: Optional[float] = None,
                 init_h_angle: Optional[float] = None):
        self.conf = conf.copy()
        self.pos = clone(init_pos)
        self.norm_fr = vector3(1, 0, 0)
        self.norm_lf = vector3(0, 1, 0)
        self.norm_up = vector3(0, 0, 1)
        self._init_rot_vectors()
        if init_h_angle is not None:
            self.rotate_h(init_h_angle)
        if init_v_angle is not None:
            self.rotate_v(init_v_angle)

    def _init_rot_vectors(self) -> None:  # :77-87
        c = self.conf
        self.rot_h_v = _v2(math.cos(c.rot_h), math.sin(c.rot_h))
        self.rot_v_v = _v2(math.cos(c.rot_v), math.sin(c.rot_v))
        rad_h, rad_v = c.fov_h / c.screen_w, c.fov_v / c.screen_h
        self.rot_scan_h_v = _v2(math.cos(rad_h), math.sin(rad_h))
        self.rot_scan_v_v = _v2(math.cos(rad_v), math.sin(rad_v))

    def rotate_h(self, angle: float) -> None:
        self.rotate_h_v(_v2(math.cos(angle), math.sin(angle)))

    def rotate_v(self, angle: float) -> None:
        self.rotate_v_v(_v2(math.cos(angle), math.sin(angle)))

    def rotate_h_step(self, n: int) -> None:  # :102-107 (rot_v_v, as in the reference)
        r = self.rot_v_v if n >= 0 else _v2(self.rot_v_v.v[0], -self.rot_v_v.v[1])
        for _ in range(abs(n)):
            self.rotate_h_v(r)

    def rotate_v_step(self, n: int) -> None:  # :110-118 (rot_h_v, as in the reference)
        r = self.rot_h_v if n >= 0 else _v2(self.rot_h_v.v[0], -self.rot_h_v.v[1])
        for _ in range(abs(n)):
            if not self.rotate_v_v(r):
                break

    def rotate_h_v(self, v: Vector) -> bool:  # :123-133
        fr_xy = _v2(self.norm_fr.v[0], self.norm_fr.v[1])
        lf_xy = _v2(self.norm_lf.v[0], self.norm_lf.v[1])
        fr_xy = rotate_vectors(fr_xy, _v2(-fr_xy.v[1], fr_xy.v[0]), v)[0]
        lf_xy = rotate_vectors(lf_xy, _v2(-lf_xy.v[1], lf_xy.v[0]), v)[0]
        self.norm_fr = vector3(fr_xy.v[0], fr_xy.v[1], self.norm_fr.v[2])
        self.norm_lf = vector3(lf_xy.v[0], lf_xy.v[1], self.norm_lf.v[2])
        self.norm_up = cross(self.norm_fr, self.norm_lf)
        return True

    def rotate_v_v(self, v: Vector) -> bool:  # :137-149
        cmp_sign = 1 if v.v[1] < 0 else -1
        fr, up = rotate_vectors(self.norm_fr, self.norm_up, v)
        dz = fr.v[2] - self.norm_fr.v[2]
        sign = (dz > 0) - (dz < 0)
        if self.conf.flags.get("vertical_locked") and sign == cmp_sign:
            return False
        self.norm_fr, self.norm_up = fr, up
        return True

    def reset_angles(self) -> None:
        self.norm_fr, self.norm_lf, self.norm_up = vector3(1, 0, 0), vector3(0, 1, 0), vector3(0, 0, 1)

    def set_pos(self, p: Vector) -> None:
        self.pos = clone(p)

    def get_pos(self) -> Vector:
        return self.pos

    def move(self, move_vec: Vector) -> None:
        self.pos = add(self.pos, move_vec)

    def get_xy_front_vector(self, unnormalized=False) -> Vector:
        res = _v2(self.norm_fr.v[0], self.norm_fr.v[1])
        return res if unnormalized else normalize(res)

    def move_xy_forward(self, s=1.0) -> None:
        f = scale(self.get_xy_front_vector(), s)
        self.move(vector3(f.v[0], f.v[1], 0))

    def move_xy_backward(self, s=1.0) -> None:
        self.move_xy_forward(-s)

    def move_xy_left(self, s=1.0) -> None:
        f = self.get_xy_front_vector()
        self.move(vector3(-f.v[1] * -s, f.v[0] * -s, 0))

    def move_xy_right(self, s=1.0) -> None:
        self.move_xy_left(-s)

    def get_front_vector(self) -> Vector:
        return clone(self.norm_fr)

    def set_vertical_lock_(self, on: bool) -> None:
        self.conf.flags["vertical_locked"] = bool(on)

    def get_dir_for_each_pixel(self) -> Iterator[Tuple[int, int, Vector]]:  # :207-250, yields (x, y, dir)
        conf = self.conf
        h_ccw = _v2(self.rot_scan_h_v.v[0], -self.rot_scan_h_v.v[1])
        v_ccw = _v2(self.rot_scan_v_v.v[0], -self.rot_scan_v_v.v[1])

        def iter_h(from_x, to_x, y, rot, beg_fr, inc, first):
            fr, lf = beg_fr, self.norm_lf
            if first:
                fr, lf = rotate_vectors(fr, lf, rot)
            i = from_x
            while i != to_x:
                yield i, y, clone(fr)
                fr, lf = rotate_vectors(fr, lf, rot)
                i += inc

        def iter_v(from_y, to_y, rot, inc, first):
            fr, up = self.norm_fr, self.norm_up
            if first:
                fr, up = rotate_vectors(fr, up, rot)
            i = from_y
            while i != to_y:
                yield from iter_h(conf.screen_h >> 1, conf.screen_h, i, self.rot_scan_h_v, fr, 1, False)
                yield from iter_h((conf.screen_h >> 1) - 1, -1, i, h_ccw, fr, -1, True)
                fr, up = rotate_vectors(fr, up, rot)
                i += inc

        yield from iter_v(conf.screen_w >> 1, conf.screen_w, self.rot_scan_v_v, 1, False)
        yield from iter_v((conf.screen_w >> 1) - 1, -1, v_ccw, -1, True)
