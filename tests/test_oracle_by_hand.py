"""Known answers worked out BY HAND from the reference's source lines, for the rows of the hot path that no jest
test of the reference pins: hit tests, normals, uv mapping, the RNG, the exposure blend, the
first-hit rule.  Each case is either an analytic value (a ray along an axis through a unit sphere ...) or an
independent transliteration of the cited TypeScript expression in plain Python floats (IEEE double, the same
arithmetic as a JS engine), so the C++ oracle is checked by something that shares no code with it.  The product's
own host mirror and native helpers are checked against the same values where they have the function."""
import ctypes as C
import math

import numpy as np
import pytest

import raytracer_js_b200 as rt
from raytracer_js_b200 import _native as N

EPS = 2.220446049250313e-16


def scene_with(oracle, *entities):
    """entities: (type, pos, extent); one mirror material, white texture, air."""
    s = oracle.Scene((0, 0, 0), 1.0) if all(0 <= p[k] - e / 2 and p[k] + e / 2 <= 1 for _, p, e in entities for k in range(3)) else oracle.Scene((-8, -8, -8), 16.0)
    m, t, sub = s.add_material(0, False, True, 0.0), s.add_texture_solid(1, 1, 1, 1), s.add_substance(1.0)
    ids = [s.add_entity(typ, p, e, m, t, sub, max_in_depth=8, max_out_depth=0) for typ, p, e in entities]
    return s, ids


# ---- FpLcg (src/math/rng/fp-lcg.ts:19-82)
def fplcg_by_the_book(seed, n):
    mul1, term1 = 3532205053565347.0 / 3768278866164713.0, 3773467585272041.0 / 4435662911655887.0
    mul2, term2 = 3632519696538149.0 / 4496133748415501.0, 3396159042346757.0 / 4429161683464229.0
    mul3, term3 = 4056279137291581.0 / 4272384783187219.0, 3685311960670787.0 / 3909517015383373.0
    s1, s2, s3 = seed, seed * mul3, seed * mul2  # seed(): :62-66
    out = []
    for _ in range(n):  # next(): :69-81, JS `%` on non-negative doubles == math.fmod
        a, b, c = math.fmod(s1 * mul1 + term1, 1.0), math.fmod(s2 * mul2 + term2, 1.0), math.fmod(s3 * mul3 + term3, 1.0)
        s1, s2, s3 = b + c, c, a + b
        out.append(math.fmod(a + b + c, 1.0))
    return out


@pytest.mark.parametrize("seed", [0.0, 1.0, 42.0, 12345.678])
def test_fplcg(oracle, seed):
    want = fplcg_by_the_book(seed, 64)
    assert oracle.fplcg(seed, 64).tolist() == want
    g = rt.FpLcg(seed)
    assert [g.next() for _ in range(64)] == want
    out = np.zeros(64)
    N.load().rt_fplcg_fill(seed, 64, out.ctypes.data_as(N._dp))  # the library's host helper (no GPU involved)
    assert out.tolist() == want
    assert all(0.0 <= v < 1.0 for v in want)
    if seed == 0.0:  # first value by hand: states are 0, so s_i = term_i % 1 = term_i and the result is their sum % 1
        t = [3773467585272041.0 / 4435662911655887.0, 3396159042346757.0 / 4429161683464229.0, 3685311960670787.0 / 3909517015383373.0]
        assert want[0] == math.fmod(t[0] + t[1] + t[2], 1.0)


# ---- Sphere.line_intersection + SphereEntity.collision_info (src/math/intersection.ts:109-128,207-220; src/entities/entity_sphere.ts:68-88)
def test_sphere_hits_by_hand(oracle):
    s, (e,) = scene_with(oracle, (0, (0.5, 0.5, 0.5), 0.5))  # centre (.5,.5,.5), diameter .5 -> radius .25
    # from outside along +z: roots t = 0.25 and 0.75 (|d| = 1), FORWARD picks the first; normal faces the ray
    pt, nm = s.collision(e, (0.5, 0.5, 0.0), (0, 0, 1))
    assert pt.tolist() == [0.5, 0.5, 0.25] and nm.tolist() == [0.0, 0.0, -1.0]
    # un-normalised direction: a = d.d = 4, roots are halved, the point is the same
    pt, nm = s.collision(e, (0.5, 0.5, 0.0), (0, 0, 2))
    assert pt.tolist() == [0.5, 0.5, 0.25] and nm.tolist() == [0.0, 0.0, -1.0]
    # from the centre: t1 = -0.25 < 0, t2 = 0.25: the far root; n = (p-c)*2/d = (0,0,1) is flipped towards the ray
    pt, nm = s.collision(e, (0.5, 0.5, 0.5), (0, 0, 1))
    assert pt.tolist() == [0.5, 0.5, 0.75] and nm.tolist() == [0.0, 0.0, -1.0]
    # behind the ray: both roots negative -> no collision; a clear miss: delta < 0
    assert s.collision(e, (0.5, 0.5, 1.0), (0, 0, 1)) is None
    assert s.collision(e, (0.0, 0.0, 0.0), (0, 0, 1)) is None
    # tangent ray (x = 0.75): delta == 0, one point; d.n == 0 -> Math.sign(0) = 0 -> the normal is the ZERO vector,
    # which Ray.trace's guard (src/raytracer.ts:200-203, `dot >= 0`) then treats as an acute hit
    pt, nm = s.collision(e, (0.75, 0.5, 0.0), (0, 0, 1))
    assert pt.tolist() == [0.75, 0.5, 0.5] and np.abs(nm).tolist() == [0.0, 0.0, 0.0]


def sphere_collision_by_the_book(c, diameter, o, d):
    """Transliteration of intersection.ts:109-128 + select_parameters FORWARD + entity_sphere.ts:68-88."""
    dot = lambda a, b: (0.0 + a[0] * b[0]) + a[1] * b[1] + a[2] * b[2]  # vector.dot: reduce from 0, left to right
    radius = diameter / 2
    dist = [o[k] - c[k] for k in range(3)]
    a = dot(d, d)
    b = dot(dist, d) * 2
    cc = dot(o, o) + dot(c, c) - dot(o, c) * 2 - radius * radius
    delta = b * b - a * cc * 4
    if delta < 0:
        return None
    sd = math.sqrt(delta)
    t1, t2 = -b / (a * 2) - sd / (a * 2), -b / (a * 2) + sd / (a * 2)
    t = t1 if t1 >= 0 else (t2 if t2 >= 0 else None)
    if t is None:
        return None
    p = [o[k] + d[k] * t for k in range(3)]
    n = [(p[k] - c[k]) * (2 / diameter) for k in range(3)]
    sg = dot(d, n)
    sg = -(1.0 if sg > 0 else (-1.0 if sg < 0 else sg))
    return p, [v * sg for v in n]


def test_sphere_hits_against_the_transliterated_formula(oracle):
    rng = np.random.default_rng(11)
    s, (e,) = scene_with(oracle, (0, (0.4, 0.55, 0.6), 0.3))
    hits = 0
    for _ in range(3000):
        o = rng.uniform(0, 1, 3).tolist()
        d = rng.normal(size=3).tolist()
        want = sphere_collision_by_the_book((0.4, 0.55, 0.6), 0.3, o, d)
        got = s.collision(e, o, d)
        assert (want is None) == (got is None)
        if want:
            hits += 1
            assert got[0].tolist() == want[0] and got[1].tolist() == want[1]  # bit for bit
    assert hits > 100


# ---- Box.line_intersection (src/math/intersection.ts:150-204) + BoxEntity.collision_info (src/entities/entity_box.ts:54-73)
def test_box_hits_by_hand(oracle):
    # unit box centred (0.5,0.5,0.5); ray from x = -2 along +x at y = z = 0.25: enters at u1 = 2 through face -x
    # (index 0), leaves at u2 = 3 through +x (index 1)
    n, u, f = oracle.box_line((0.5, 0.5, 0.5), (1, 1, 1), (-2, 0.25, 0.25), (1, 0, 0))
    assert n == 2 and u.tolist() == [2.0, 3.0] and f.tolist() == [0, 1]
    # along -y from above: faces +y (3) in, -y (2) out
    n, u, f = oracle.box_line((0.5, 0.5, 0.5), (1, 1, 1), (0.5, 3, 0.5), (0, -1, 0))
    assert n == 2 and u.tolist() == [2.0, 3.0] and f.tolist() == [3, 2]
    # a miss: u1 > u2 -> []
    n, _, _ = oracle.box_line((0.5, 0.5, 0.5), (1, 1, 1), (-2, 2, 0.25), (1, 0, 0))
    assert n == 0
    # exit ties go x, then y, then z (strict `<` in face order): the diagonal leaves through +x
    n, u, f = oracle.box_line((0.5, 0.5, 0.5), (1, 1, 1), (0.5, 0.5, 0.5), (1, 1, 1))
    assert n == 2 and u.tolist() == [-0.5, 0.5] and f.tolist() == [0, 1]
    # BoxEntity: first forward parameter, normal = the face's axis flipped towards the ray
    s, (e,) = scene_with(oracle, (1, (0.5, 0.5, 0.5), 0.5))
    pt, nm = s.collision(e, (0.0, 0.5, 0.5), (1, 0, 0))
    assert pt.tolist() == [0.25, 0.5, 0.5] and nm.tolist() == [-1.0, 0.0, 0.0]
    pt, nm = s.collision(e, (0.5, 0.5, 0.5), (0, 0, -1))  # from inside: u1 < 0, the exit face; normal against the ray
    assert pt.tolist() == [0.5, 0.5, 0.25] and nm.tolist() == [0.0, 0.0, 1.0]
    assert s.collision(e, (0.0, 0.5, 0.5), (-1, 0, 0)) is None


# ---- uv_map_sphere (src/math/uv_mapping.ts:19-25)
def test_uv_map_sphere_by_hand(oracle):
    def book(v):
        u = math.atan2(v[1], v[0]) / math.pi / 2 + 0.5 - EPS
        w = math.atan2(v[2], math.sqrt((0.0 + v[0] * v[0]) + v[1] * v[1])) / math.pi + 0.5 - EPS
        return [u, w]
    assert oracle.uv_map_sphere((1, 0, 0)).tolist() == [0.5 - EPS, 0.5 - EPS]
    assert oracle.uv_map_sphere((0, 0, 1)).tolist() == [0.5 - EPS, 1.0 - EPS]
    assert oracle.uv_map_sphere((0, 0, -1)).tolist() == [0.5 - EPS, 0.0 - EPS]  # below 0 by eps: get_color's `< 0 - eps` guard lets it through
    assert oracle.uv_map_sphere((-1, 0, 0)).tolist() == [1.0 - EPS, 0.5 - EPS]  # atan2(0,-1) = pi
    rng = np.random.default_rng(5)
    for _ in range(200):
        v = rng.normal(size=3).tolist()
        got = oracle.uv_map_sphere(v).tolist()
        want = book(v)
        assert got[0] == pytest.approx(want[0], abs=4e-16) and got[1] == pytest.approx(want[1], abs=4e-16)  # libm vs libm: <= 1 ulp


# ---- first hit in LIST ORDER, not nearest (src/raytracer.ts:186-195) and the exposure blend (src/view/exposure_buffer.ts:53-91)
def test_first_hit_rule_and_blend_by_hand(oracle):
    # two spheres in the same node (both straddle the root's centre planes), the FARTHER one inserted first:
    # the ray along +x (the camera's front vector at zero angles) meets the near one first in space, but the reference takes the first in the Set
    s = oracle.Scene((0, 0, 0), 1.0)
    m = s.add_material(0, False, False, 0.0)  # diffuse: the path ends at the hit with the texture colour
    red, green, sky = s.add_texture_solid(1, 0, 0, 1), s.add_texture_solid(0, 1, 0, 1), s.add_texture_solid(0.2, 0.2, 0.7, 1)
    sub = s.add_substance(1.0)
    far = s.add_entity(0, (0.7, 0.5, 0.5), 0.3, m, red, sub, max_in_depth=8, max_out_depth=0)
    near = s.add_entity(0, (0.45, 0.5, 0.5), 0.3, m, green, sub, max_in_depth=8, max_out_depth=0)
    assert s.entity_node(far) == 0 and s.entity_node(near) == 0
    cam = oracle.Camera(math.pi / 2, math.pi / 2, 3, 3, (0.05, 0.5, 0.5), 0.0, 0.0, vertical_locked=True)
    xy, d, n = cam.dirs()
    centre = [i for i in range(n) if tuple(xy[i]) == (1, 1)][0]
    front = d[centre]
    # the centre pixel looks along the camera's front vector; both spheres are on that line
    assert front.tolist() == [1.0, 0.0, 0.0]
    assert s.collision(near, (0.05, 0.5, 0.5), front) is not None and s.collision(far, (0.05, 0.5, 0.5), front) is not None
    rgb, ids, _, _ = oracle.render(s, cam, refmax=1, sky_texture=sky, default_substance=sub, n_frames=1, rng_mode=1, seed=1.0)
    assert ids[1, 1] == far and rgb[1, 1].tolist() == [1.0, 0.0, 0.0]  # the red, farther, first-inserted sphere wins
    # blend: frame k has weight 1/(1+k); a constant colour stays itself, float32 rounding per frame
    rgb3, _, _, _ = oracle.render(s, cam, refmax=1, sky_texture=sky, default_substance=sub, n_frames=3, rng_mode=1, seed=1.0)
    px = np.float32(0.0)
    for k in range(3):
        w = 1 / (1 + k)
        px = np.float32(0.7 * w + float(px) * (1 - w))
    corner = rgb3[0, 0]
    if ids[0, 0] < 0:  # a sky pixel: (0.2, 0.2, 0.7) three times
        assert corner[2] == px


# ---- Box.line_intersection, transliterated (src/math/intersection.ts:150-204; isNegative: src/math/mathutils.ts:45-47)
def box_line_by_the_book(centre, size, o, d):
    with np.errstate(divide="ignore", invalid="ignore"):
        tl = [centre[k] - size[k] * 0.5 for k in range(3)]
        p = [-d[0], d[0], -d[1], d[1], -d[2], d[2]]
        q = [o[0] - tl[0], tl[0] + size[0] - o[0], o[1] - tl[1], tl[1] + size[1] - o[1], o[2] - tl[2], tl[2] + size[2] - o[2]]
        u1, u2, i1, i2 = -math.inf, math.inf, None, None
        for i in range(6):
            u = float(np.float64(q[i]) / np.float64(p[i]))  # JS division: x/0 = +-Infinity, 0/0 = NaN
            if p[i] < 0 or (p[i] == 0 and math.copysign(1.0, p[i]) < 0):  # isNegative: includes -0
                if u > u1:
                    u1, i1 = u, i
            elif u < u2:
                u2, i2 = u, i
        return None if u1 > u2 else (u1, u2, i1, i2)


def test_box_line_against_the_transliterated_formula(oracle):
    rng = np.random.default_rng(23)
    hits = 0
    for trial in range(3000):
        c = rng.uniform(0.2, 0.8, 3).tolist()
        sz = [float(rng.uniform(0.05, 0.4))] * 3
        o = rng.uniform(0, 1, 3).tolist()
        d = rng.normal(size=3)
        if trial % 5 == 0:
            d[rng.integers(0, 3)] = 0.0  # axis-parallel rays: +-0 components, infinite parameters
        if trial % 10 == 0:
            d[rng.integers(0, 3)] = -0.0
        d = d.tolist()
        want = box_line_by_the_book(c, sz, o, d)
        n, u, f = oracle.box_line(c, sz, o, d)
        assert (want is None) == (n == 0), (c, sz, o, d)
        if want:
            hits += 1
            assert u.tolist() == [want[0], want[1]] and f.tolist() == [want[2], want[3]]
    assert hits > 200
