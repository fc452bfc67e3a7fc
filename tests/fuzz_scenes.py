"""Random small scenes and poses for the differential fuzz tests (kernel body on the host / CUDA path vs the oracle):
every material response (reflection: mirror, rough mirror, diffuse, light; transmission: smooth and rough; BOTH), spheres
and boxes, all three substances, cameras inside and outside the cube, on dyadic planes, inside entities, axis-aligned
and arbitrary directions, solid and image textures, square and ragged frames, 1-5 exposure frames, refmax 1-9.  tools/fuzz_parity.py runs it for
as long as one likes (3 000 cases without a mismatch at the time of writing)."""
import math
import random

import oracle as orc
import raytracer_js_b200 as rt
from raytracer_js_b200 import scenes

MATERIAL_SETS = [["mirror", "diffuse", "light"], ["mirror", "rough", "light", "diffuse"], ["glass", "mirror", "both", "light"],
                 ["roughglass", "glass", "rough", "mirror"], ["diffuse"]]


def build_scene(seed, n, dmin, dmax, kinds, images=False):
    rng = rt.FpLcg(float(seed))
    texs = [scenes.checker_texture(64, 32, seed=s) for s in (1, 2)] if images else None
    tree = rt.new_entity_octree(rt.OctreeDim(rt.point(0, 0, 0), 1.0), None)
    mats = {"glass": rt.SolidMaterial(rt.ResponseType.TRANSMISSION, False, False, 0),
            "mirror": rt.SolidMaterial(rt.ResponseType.REFLECTION, False, True, 0),
            "rough": rt.SolidMaterial(rt.ResponseType.REFLECTION, False, True, 0.5),
            "diffuse": rt.SolidMaterial(rt.ResponseType.REFLECTION, False, False, 0),
            "light": rt.SolidMaterial(rt.ResponseType.REFLECTION, True, False, 0),
            "both": rt.SolidMaterial(rt.ResponseType.BOTH, False, False, 0),
            "roughglass": rt.SolidMaterial(rt.ResponseType.TRANSMISSION, False, False, 0.3)}
    subs = [rt.SUBSTANCE_AIR, rt.SUBSTANCE_WATER, rt.SUBSTANCE_GLASS]
    ents = []
    for _ in range(n):
        d = dmin + rng.next() * (dmax - dmin)
        c = [d / 2 + rng.next() * (1 - d) for _ in range(3)]
        m = mats[kinds[int(rng.next() * len(kinds))]]
        tex = rt.SolidTexture(rt.Color(0.2 + rng.next(), 0.2 + rng.next(), 0.2 + rng.next(), 1))
        if texs and rng.next() < 0.6:
            tex = texs[int(rng.next() * 2)]
        cls = rt.BoxEntity if rng.next() < 0.25 else rt.SphereEntity
        e = cls(None, m, tex, subs[int(rng.next() * 3)], rt.point(*c), d)
        rt.add_entity_to_octree(tree, e, {"max_in_depth": 16, "max_out_depth": 0})
        ents.append(e)
    sky = rt.SkySphere(scenes.checker_texture(128, 64, seed=4) if images else rt.SolidTexture(rt.Color(0.2, 0.3, 0.7, 1)))
    return scenes.SceneBundle(tree, ents, sky, rt.SUBSTANCE_AIR, 5)


def cases(seed, count, max_entities=4000):
    """Yields dicts: bundle, pos, yaw, pitch, w, h, n_frames, refmax (+ the generator's choices, for the report)."""
    R = random.Random(seed)
    for k in range(count):
        c = dict(case=k, seed=R.randint(1, 10 ** 6), n=min(max_entities, R.choice([40, 300, 1500, 4000])),
                 d=R.choice([(0.01, 0.05), (0.03, 0.2), (0.005, 0.02), (0.1, 0.45), (0.02, 0.08)]), kinds=R.choice(MATERIAL_SETS))
        c["pos"] = R.choice([(0.5, 0.5, 0.5), (R.random(), R.random(), R.random()), (R.random(), R.random(), R.random()),
                             (0.25, 0.75, 0.5), (0.013, 0.487, 0.021), (0.5, 0.5, 0.0009765625), (-0.2, 0.5, 0.4)])
        c["yaw"] = R.choice([0.0, 30.0, 45.0, 90.0, 180.0, R.uniform(-180, 180)])
        c["pitch"] = R.choice([0.0, 0.3, -0.7, R.uniform(-1.2, 1.2)])
        c["w"], c["h"] = R.choice([(40, 40), (64, 48), (33, 57), (96, 16), (72, 72)])
        c["n_frames"] = R.choice([1, 2, 5])
        c["refmax"] = R.choice([1, 4, 6, 9])
        c["images"] = R.random() < 0.25  # image textures (nearest texel, src/texture/texture_image.ts:40-63) and an image sky
        c["bundle"] = build_scene(c["seed"], c["n"], c["d"][0], c["d"][1], c["kinds"], c["images"])
        yield c


def cameras(c):
    cam = scenes.bench_camera(c["w"], c["h"], c["pos"], c["yaw"], c["pitch"])
    ocam = orc.Camera(math.pi / 2, math.pi / 2, c["w"], c["h"], c["pos"], c["pitch"], math.pi / 180 * c["yaw"], vertical_locked=True)
    return cam, ocam


def describe(c):
    return {k: v for k, v in c.items() if k != "bundle"}


# ---- lattice scenes: everything on a dyadic grid - entities that fill their cells exactly, touch each other and the
# cell planes, are centred on lattice planes; cameras on lattice points (the demo pose (0.5, 0.5, 0.5) is one) looking
# along the axes and the diagonals.  Exact ties everywhere: rays inside cell-boundary planes, through lattice corners,
# cameras standing on the corner of a box.  What the reference does there is decided by its half-open cells and the tie
# order of its walker, not by geometry (rt_trace.cuh: hit_only_touches_its_cell).
def lattice_scene(seed, cells, fill, kinds):
    R = random.Random(seed)
    tree = rt.new_entity_octree(rt.OctreeDim(rt.point(0, 0, 0), 1.0), None)
    mats = {"mirror": rt.SolidMaterial(rt.ResponseType.REFLECTION, False, True, 0),
            "diffuse": rt.SolidMaterial(rt.ResponseType.REFLECTION, False, False, 0),
            "light": rt.SolidMaterial(rt.ResponseType.REFLECTION, True, False, 0),
            "glass": rt.SolidMaterial(rt.ResponseType.TRANSMISSION, False, False, 0),
            "rough": rt.SolidMaterial(rt.ResponseType.REFLECTION, False, True, 0.5)}
    ents = []
    step = 1.0 / cells
    for i in range(cells):
        for j in range(cells):
            for k in range(cells):
                if R.random() > fill:
                    continue
                d = step * R.choice([1.0, 1.0, 0.5, 2.0])  # fills its cell / half of it / overlaps the neighbours
                c = [(i + 0.5) * step, (j + 0.5) * step, (k + 0.5) * step]
                if R.random() < 0.3:
                    c = [i * step, j * step, k * step]  # centred ON the lattice planes
                if any(x - d / 2 < 0 or x + d / 2 > 1 for x in c):
                    continue
                m = mats[R.choice(kinds)]
                tex = rt.SolidTexture(rt.Color(0.2 + R.random(), 0.2 + R.random(), 0.2 + R.random(), 1))
                cls = rt.BoxEntity if R.random() < 0.4 else rt.SphereEntity
                sub = rt.SUBSTANCE_GLASS if m.response == rt.ResponseType.TRANSMISSION else rt.SUBSTANCE_AIR
                e = cls(None, m, tex, sub, rt.point(*c), d)
                rt.add_entity_to_octree(tree, e, {"max_in_depth": 16, "max_out_depth": 0})
                ents.append(e)
    return scenes.SceneBundle(tree, ents, rt.SkySphere(rt.SolidTexture(rt.Color(0.2, 0.3, 0.7, 1))), rt.SUBSTANCE_AIR, 5)


def lattice_cases(seed, count):
    R = random.Random(seed)
    k = 0
    while k < count:
        cells = R.choice([4, 8, 8, 16])
        b = lattice_scene(R.randint(1, 10 ** 6), cells, R.choice([0.05, 0.2, 0.5]),
                          R.choice([["mirror", "diffuse", "light"], ["mirror", "rough"], ["glass", "mirror", "light"], ["diffuse"]]))
        g = 1.0 / R.choice([2, 4, 8, 16])
        pos = R.choice([(0.5, 0.5, 0.5), (R.randint(1, 7) / 8, R.randint(1, 7) / 8, R.randint(1, 7) / 8), (g, 0.5, 0.5 + g / 2),
                        (R.random(), R.random(), R.random()), (-0.25, 0.5, 0.5)])
        c = dict(case=k, cells=cells, n=len(b.entities), pos=pos, yaw=R.choice([0.0, 90.0, 180.0, -90.0, 45.0, R.uniform(-180, 180)]),
                 pitch=R.choice([0.0, 0.0, math.pi / 4, R.uniform(-1, 1)]), w=R.choice([33, 48, 64]), h=R.choice([33, 48]),
                 n_frames=R.choice([1, 2]), refmax=R.choice([1, 4, 7]), bundle=b, images=False)
        if not b.entities:
            continue
        k += 1
        yield c


def edge_cases(seed, count):
    """Cameras outside the cube on every side, on its faces and corners, a hair inside; entities from 1e-4 to 0.9 of the
    root; refmax 0 - 5; square frames with the reference's own extents (`reference_extents`), ragged ones without."""
    R = random.Random(seed)
    k = 0
    while k < count:
        lattice = R.random() < 0.4
        if lattice:
            b = lattice_scene(R.randint(1, 10 ** 6), R.choice([2, 4, 8]), R.choice([0.1, 0.4, 0.9]),
                              R.choice([["mirror", "diffuse", "light"], ["glass", "mirror", "light"], ["rough", "mirror"]]))
            if not b.entities:
                continue
        else:
            dm = R.choice([(1e-4, 1e-3), (0.3, 0.9), (0.0005, 0.3), (0.05, 0.06)])
            b = build_scene(R.randint(1, 10 ** 6), R.choice([3, 30, 600]), dm[0], dm[1], R.choice(MATERIAL_SETS), R.random() < 0.2)
        pos = R.choice([(-0.5, 0.5, 0.5), (1.5, 0.3, 0.7), (0.5, -2.0, 0.5), (0.5, 0.5, 3.0), (0.0, 0.5, 0.5), (1.0, 1.0, 1.0), (0.0, 0.0, 0.0),
                        (1e-9, 0.5, 0.5), (0.999999999, 0.2, 0.9), (R.uniform(-1, 2), R.uniform(-1, 2), R.uniform(-1, 2)), (0.5, 0.5, 0.5)])
        square = R.random() < 0.4
        w, h = (R.choice([32, 50, 64]),) * 2 if square else R.choice([(40, 24), (24, 40), (72, 8)])
        yield dict(case=k, lattice=lattice, n=len(b.entities), pos=pos, yaw=R.choice([0.0, 90.0, -90.0, 180.0, 45.0, 135.0, R.uniform(-180, 180)]),
                   pitch=R.choice([0.0, 0.5, -0.5, 1.4, -1.4, R.uniform(-1.5, 1.5)]), w=w, h=h, n_frames=R.choice([1, 3]),
                   refmax=R.choice([0, 1, 2, 5]), bundle=b, images=False, reference_extents=square and R.random() < 0.5)
        k += 1
