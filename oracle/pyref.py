"""A second, independent restatement of Ray.trace() (src/raytracer.ts:168-277) in plain Python, used to
cross-check the C++ oracle on small scenes: the path logic (guards, materials, reflect / scatter / refract,
refmax, sky, inverse-square law), the hit tests and the blend are transliterated here line by line from the
TypeScript, and the walker and node_at_pos come from oracle/pywalker.py, the same kind of transliteration of
src/octree_space.ts.  Nothing of the oracle's logic is used: only the tree it built (flat arrays), which
tests/test_host_build.py checks against the host API's own builder.  TEST INFRASTRUCTURE ONLY."""
import math

import numpy as np

from . import pywalker

EPS = 2.220446049250313e-16
BRANCHES = {}  # how often each branch of Ray.trace was taken (the tests assert that all of them are)


def _took(name):
    BRANCHES[name] = BRANCHES.get(name, 0) + 1


def dot(a, b):  # vector.dot (src/math/vector.ts:76-84)
    s = 0.0
    for i in range(3):
        s += a[i] * b[i]
    return s


def js_sign(x):
    return 1.0 if x > 0 else (-1.0 if x < 0 else x)


def is_negative(x):  # mathutils.isNegative
    return x < 0 or (x == 0 and math.copysign(1.0, x) < 0)


def jsdiv(a, b):
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.float64(a) / np.float64(b))


class FpLcg:  # src/math/rng/fp-lcg.ts:49-82
    M1, T1 = 3532205053565347.0 / 3768278866164713.0, 3773467585272041.0 / 4435662911655887.0
    M2, T2 = 3632519696538149.0 / 4496133748415501.0, 3396159042346757.0 / 4429161683464229.0
    M3, T3 = 4056279137291581.0 / 4272384783187219.0, 3685311960670787.0 / 3909517015383373.0

    def seed(self, s):
        self.s1, self.s2, self.s3 = s, s * self.M3, s * self.M2

    def next(self):
        a, b, c = math.fmod(self.s1 * self.M1 + self.T1, 1.0), math.fmod(self.s2 * self.M2 + self.T2, 1.0), math.fmod(self.s3 * self.M3 + self.T3, 1.0)
        self.s1, self.s2, self.s3 = b + c, c, a + b
        return math.fmod(a + b + c, 1.0)


class TextureError(Exception):
    pass


def uv_map_sphere(v):  # src/math/uv_mapping.ts:19-25
    u = math.atan2(v[1], v[0]) / math.pi / 2.0 + 0.5 - EPS
    w = math.atan2(v[2], math.sqrt((0.0 + v[0] * v[0]) + v[1] * v[1])) / math.pi + 0.5 - EPS
    return u, w


class Image:
    """ImageTexture (src/texture/texture_image.ts:28-63): image_data = byte / 255.0, nearest texel."""

    def __init__(self, pixels_u8):  # uint8 [height, width, 3]
        self.height, self.width = pixels_u8.shape[:2]
        self.data = [float(b) / 255.0 for b in pixels_u8.reshape(-1)]

    def get_color(self, u, v):
        if u < 0 - EPS or u > 1 - EPS or v < 0 - EPS or v > 1 - EPS:
            raise TextureError("Texture coordinates out of bounds")
        ui, vi = pywalker.to_int32(u * self.width), pywalker.to_int32(v * self.height)
        i = (vi * self.width + ui) * 3
        return self.data[i], self.data[i + 1], self.data[i + 2]


def texture_color(tex, u, v):
    return tex.get_color(u, v) if isinstance(tex, Image) else tex  # SolidTexture.get_color ignores u, v


class Ent:
    def __init__(self, kind, pos, extent, material, texture, substance):
        self.kind, self.pos, self.extent = kind, list(map(float, pos)), float(extent)
        self.material, self.texture, self.substance = material, texture, substance  # dict, (r,g,b), refractive index | None

    def collision_info(self, o, d):
        if self.kind == 0:  # SphereEntity.collision_info (src/entities/entity_sphere.ts:68-88)
            c, radius = self.pos, self.extent / 2
            dist = [o[k] - c[k] for k in range(3)]
            a = dot(d, d)
            b = dot(dist, d) * 2
            cc = dot(o, o) + dot(c, c) - dot(o, c) * 2 - radius * radius
            delta = b * b - a * cc * 4
            if delta < 0:
                return None
            sd = math.sqrt(delta)
            tmp1, tmp2 = jsdiv(-b, a * 2), jsdiv(sd, a * 2)
            t1, t2 = tmp1 - tmp2, tmp1 + tmp2
            ts = [t for t in (t1, t2) if t >= 0]
            if not ts:
                return None
            p = [o[k] + d[k] * ts[0] for k in range(3)]
            n = [(p[k] - c[k]) * (2 / self.extent) for k in range(3)]
            sg = -js_sign(dot(d, n))
            return p, [v * sg for v in n]
        # BoxEntity.collision_info (src/entities/entity_box.ts:54-73) over Box.line_intersection (intersection.ts:150-204)
        size = [self.extent] * 3
        tl = [self.pos[k] - size[k] * 0.5 for k in range(3)]
        pp = [-d[0], d[0], -d[1], d[1], -d[2], d[2]]
        q = [o[0] - tl[0], tl[0] + size[0] - o[0], o[1] - tl[1], tl[1] + size[1] - o[1], o[2] - tl[2], tl[2] + size[2] - o[2]]
        u1, u2, i1, i2 = -math.inf, math.inf, None, None
        for i in range(6):
            u = jsdiv(q[i], pp[i])
            if is_negative(pp[i]):
                if u > u1:
                    u1, i1 = u, i
            elif u < u2:
                u2, i2 = u, i
        if u1 > u2:
            return None
        normals = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]
        params = [(u, i) for u, i in ((u1, i1), (u2, i2)) if u >= 0]
        if not params:
            return None
        u, i = params[0]
        n = [float(v) for v in normals[i]]
        sg = -js_sign(dot(d, n))
        return [o[k] + d[k] * u for k in range(3)], [v * sg for v in n]

    def is_within(self, p):
        if self.kind == 0:  # entity_sphere.ts:63-66
            dist = [p[k] - self.pos[k] for k in range(3)]
            return dot(dist, dist) <= (self.extent / 2) * (self.extent / 2)
        # entity_box.ts:47-52: point_in_space(point, {pos: get_pos(), size}) - the CENTRE as the min corner, half-open
        return all(p[k] >= self.pos[k] and p[k] < self.pos[k] + self.extent for k in range(3))


def trace(tree, flat, ents, cfg, rng, start, direction, start_substance):
    """Ray.trace(): returns (r, g, b), first-hit entity index or -1."""
    refpoint, d = list(start), list(direction)
    walker = pywalker.Walker(tree)
    first_segment = True
    col = [1.0, 1.0, 1.0]
    refcount, path_distance, cur_substance = 0, 0.0, start_substance
    light_hit, first = False, -1
    while True:
        hit = None
        # walker.set_pos_and_dir(refpoint, dir[, startnode]) + next(): existing nodes in visit order
        stops = walker.stops(refpoint, d, use_start_node=first_segment)
        first_segment = False
        for _, _, node in stops:
            for e in flat.list_entity[flat.node_list_off[node]:flat.node_list_off[node + 1]]:
                ci = ents[e].collision_info(refpoint, d)
                if ci is not None:  # the FIRST entity of the list that is hit, not the nearest (:186-195)
                    hit = (int(e), ci)
                    break
            if hit:
                break
        if not hit:
            break
        e, (point, normal) = hit
        ent = ents[e]
        if first < 0:
            first = e
        if dot(d, normal) >= 0:  # :200-203
            _took("acute")
            return col, first
        refcount += 1
        # SolidMaterial.alter_ray (src/materials/material_solid.ts:30-36): entity.map_uv(p) -> texture.get_color -> mul_color
        u, v = uv_map_sphere([point[k] - ent.pos[k] for k in range(3)]) if ent.kind == 0 else (0, 0)
        tc = texture_color(ent.texture, u, v)
        col = [col[k] * tc[k] for k in range(3)]
        dd = [point[k] - refpoint[k] for k in range(3)]
        path_distance += math.sqrt(dot(dd, dd))
        refpoint = list(point)
        m = ent.material
        if m["light"]:
            _took("light")
            light_hit = True
            break
        if m["response"] == 0:  # REFLECTION
            if not m["mirror"]:
                _took("diffuse")
                return col, first
            _took("box mirror" if ent.kind == 1 else "mirror")
            k = -dot(d, normal) * 2  # vector.reflection (vector.ts:263-268)
            d = [d[i] + normal[i] * k for i in range(3)]
            if m["roughness"] > 0.0:  # scatter_ray :121-133
                _took("scatter")
                while True:  # isotropic_sphere_sample (vector_utils.ts:8-14)
                    rv = [rng.next() * 2 - 1, rng.next() * 2 - 1, rng.next() * 2 - 1]
                    if not dot(rv, rv) > 1:
                        break
                if dot(rv, normal) < 0:
                    rv = [-v for v in rv]
                rf = [d[i] * (1 - m["roughness"]) + rv[i] * m["roughness"] for i in range(3)]
                inv = 1.0 / math.sqrt(dot(rf, rf))
                d = [v * inv for v in rf]
            refpoint = [refpoint[i] + d[i] * 1e-3 for i in range(3)]  # move_slightly_forward
        elif m["response"] == 1:  # TRANSMISSION :238-249
            refpoint = [refpoint[i] + d[i] * 1e-3 for i in range(3)]
            rf_entity = entity_at_pos(tree, flat, ents, refpoint)
            substance = rf_entity.substance if rf_entity is not None else cfg["default_substance"]
            _took("undefined substance" if substance is None else ("transmission into an entity" if rf_entity is not None else "transmission into the default substance"))
            if substance is not None:  # refract_ray :135-150
                r_ratio = cur_substance / substance
                cosine = dot(d, normal)
                ref_sine_sq = (1 - cosine * cosine) * (r_ratio * r_ratio)
                _took("refraction" if ref_sine_sq <= 1 else "total internal reflection")
                if ref_sine_sq <= 1:
                    kk = math.sqrt(1 - ref_sine_sq) - cosine
                    d = [d[i] * r_ratio - normal[i] * kk for i in range(3)]
                else:
                    k = -dot(d, normal) * 2
                    d = [d[i] + normal[i] * k for i in range(3)]
                cur_substance = substance
        else:
            _took("both")
            return col, first
        if refcount >= cfg["refmax"]:
            _took("refmax")
            return [0.0, 0.0, 0.0], first
    if not light_hit:
        _took("sky")
        sky = texture_color(cfg["sky"], *uv_map_sphere(d))  # SkySphere.get_color (src/sky/sky_sphere.ts:22-27)
        return [col[k] * sky[k] for k in range(3)], first
    t = path_distance * cfg["attenuation"]
    isl = 1.0 / (EPS + t * t)
    return [c * isl for c in col], first


def entity_at_pos(tree, flat, ents, p):  # src/octree_entity.ts:191-202
    at = tree.node_at_pos(p)
    node = at[0] if at is not None else -1
    while node >= 0:
        for e in flat.list_entity[flat.node_list_off[node]:flat.node_list_off[node + 1]]:
            if ents[e].is_within(p):
                return ents[e]
        node = int(flat.node_parent[node])
    return None


def render(scene, ents, cam, cfg, n_frames=1, seed=1.0):
    """Raytracer.trace_frame() x n_frames with next_frame() between them, the harness RNG policy (reseed per pixel)."""
    flat = scene.flat()
    tree = pywalker.Tree(flat)
    xy, dirs, n = cam.dirs(fixed_extents=True)
    W, H = cam.screen_w, cam.screen_h
    rgb = np.zeros((H, W, 3), np.float32)
    ids = np.full((H, W), -1, np.int32)
    start = [float(v) for v in cam.basis()["pos"]]
    start_ent = entity_at_pos(tree, flat, ents, start)
    start_substance = start_ent.substance if start_ent is not None else cfg["default_substance"]
    rng = FpLcg()
    for f in range(n_frames):
        w = 1 / (1 + f)
        for i in range(n):
            x, y = int(xy[i][0]), int(xy[i][1])
            rng.seed(seed + float(y * W + x) + float(f) * float(W) * float(H))
            c, first = trace(tree, flat, ents, cfg, rng, start, [float(v) for v in dirs[i]], start_substance)
            for k in range(3):
                rgb[y, x, k] = np.float32(c[k] * w + float(rgb[y, x, k]) * (1 - w))
            ids[y, x] = first
    return rgb, ids
