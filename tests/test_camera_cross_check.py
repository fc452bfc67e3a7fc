"""Camera (src/view/camera.ts:27-250): the product's host mirror (raytracer.js_b200/camera.py, plain Python) and the
oracle's C++ restatement are two independent transliterations; they must agree bit for bit on the basis after any
sequence of rotations and on the per-pixel directions of get_dir_for_each_pixel (iterated rotations, centre-out
scan order), for several fields of view, sizes, poses, with and without the vertical lock.  The reference's own
test (test/view-camera.test.ts:17-49) only checks unit length."""
import math

import numpy as np
import pytest

import raytracer_js_b200 as rt


def pair(oracle, fov_v, fov_h, w, h, pos, v_angle, h_angle, locked, rot_v=math.pi / 30, rot_h=math.pi / 30):
    conf = rt.CameraConfig(fov_v=fov_v, fov_h=fov_h, screen_w=w, screen_h=h, rot_v=rot_v, rot_h=rot_h,
                           flags={"vertical_locked": locked})
    cam = rt.Camera(conf, rt.point(*pos), v_angle, h_angle)
    ocam = oracle.Camera(fov_v, fov_h, w, h, pos, v_angle, h_angle, rot_v=rot_v, rot_h=rot_h, vertical_locked=locked)
    return cam, ocam


def same_basis(cam, ocam):
    b = ocam.basis()
    assert list(b["fr"]) == cam.norm_fr.v and list(b["lf"]) == cam.norm_lf.v and list(b["up"]) == cam.norm_up.v
    assert list(b["pos"]) == cam.get_pos().v


def same_dirs(cam, ocam):
    xy, d, n = ocam.dirs(fixed_extents=False)
    mine = list(cam.get_dir_for_each_pixel())
    assert n == len(mine)
    assert [(x, y) for x, y, _ in mine] == [tuple(p) for p in xy.tolist()]  # the scan order
    np.testing.assert_array_equal(np.array([v.v for _, _, v in mine]), d)   # bit for bit


@pytest.mark.parametrize("fov_v,fov_h,size,pos,v_angle,h_angle,locked", [
    (math.pi / 2, math.pi / 2, 48, (0.5013, 0.4987, 0.5021), 0.0, math.pi / 6, True),
    (math.pi / 3, math.pi / 2, 33, (0.1, 0.9, 0.4), 0.4, -2.0, True),
    (math.pi, math.pi, 20, (0.0, 0.0, 0.0), None, None, False),
    (1.0, 0.7, 25, (3.0, -2.0, 8.5), -1.2, 4.0, False),
])
def test_basis_and_directions(oracle, fov_v, fov_h, size, pos, v_angle, h_angle, locked):
    cam, ocam = pair(oracle, fov_v, fov_h, size, size, pos, v_angle, h_angle, locked)
    same_basis(cam, ocam)
    same_dirs(cam, ocam)


def test_rotation_sequences(oracle):
    cam, ocam = pair(oracle, math.pi / 2, math.pi / 2, 16, 16, (0.3, 0.3, 0.3), 0.1, 0.2, True)
    rng = np.random.default_rng(4)
    for step in range(40):
        k = int(rng.integers(0, 4))
        if k == 0:
            a = float(rng.normal())
            cam.rotate_h(a); ocam.rotate_h(a)
        elif k == 1:
            a = float(rng.normal() * 0.5)
            cam.rotate_v(a); ocam.rotate_v(a)
        elif k == 2:
            n = int(rng.integers(-5, 6))
            cam.rotate_h_step(n); ocam.rotate_h_step(n)
        else:
            n = int(rng.integers(-5, 6))
            cam.rotate_v_step(n); ocam.rotate_v_step(n)
        same_basis(cam, ocam)
    same_dirs(cam, ocam)
    p = rt.point(0.7, 0.1, 0.6)
    cam.set_pos(p); ocam.set_pos(p.v)
    same_basis(cam, ocam)
