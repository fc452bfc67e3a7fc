"""The C++ oracle against oracle/pyref.py + oracle/pywalker.py, an independent plain-Python transliteration of
Ray.trace(), the hit tests, the OctreeWalker and node_at_pos (sharing no logic with the oracle): small frames of
scenes that exercise every branch of the path - mirrors, rough mirrors (RNG), lights (inverse-square law), glass
with defined and undefined substances, total internal reflection, boxes, refmax, the acute-normal guard - must
come out bit for bit the same.  This is the pin for the rows of SURVEY.md 8c that no reference test covers."""
import math

import numpy as np
import pytest

from oracle import pyref


def build(oracle, seed, n, kinds):
    """The same entities into the oracle's Scene and into pyref.Ent objects."""
    rng = pyref.FpLcg()
    rng.seed(seed)
    s = oracle.Scene((0, 0, 0), 1.0)
    mats = {"mirror": dict(response=0, light=False, mirror=True, roughness=0.0), "rough": dict(response=0, light=False, mirror=True, roughness=0.4),
            "diffuse": dict(response=0, light=False, mirror=False, roughness=0.0), "light": dict(response=0, light=True, mirror=False, roughness=0.0),
            "glass": dict(response=1, light=False, mirror=False, roughness=0.0), "both": dict(response=2, light=False, mirror=False, roughness=0.0)}
    mid = {k: s.add_material(m["response"], m["light"], m["mirror"], m["roughness"]) for k, m in mats.items()}
    subs = [1.0, 1.33, 1.5, None]
    sid = [s.add_substance(v) if v is not None else -1 for v in subs]
    sky = s.add_texture_solid(0.2, 0.2, 0.7, 1.0)
    ents = []
    for _ in range(n):
        d = 0.04 + rng.next() * 0.2
        c = [d / 2 + rng.next() * (1 - d) for _ in range(3)]
        kind = kinds[int(rng.next() * len(kinds))]
        k = 4.0 if kind == "light" else 1.0
        col = (0.2 + rng.next() * k, 0.2 + rng.next() * k, 0.2 + rng.next() * k)
        typ = 1 if rng.next() < 0.3 else 0
        si = int(rng.next() * 4)
        t = s.add_texture_solid(col[0], col[1], col[2], 1.0)
        assert s.add_entity(typ, c, d, mid[kind], t, sid[si], max_in_depth=16, max_out_depth=0) == len(ents)
        ents.append(pyref.Ent(typ, c, d, mats[kind], col, subs[si]))
    return s, ents, sky, sid[0]


@pytest.mark.parametrize("seed,n,kinds,refmax,pos", [
    (3.0, 60, ["mirror", "rough", "diffuse", "light"], 4, (0.5013, 0.4987, 0.5021)),
    (5.0, 50, ["glass", "glass", "mirror", "light", "both"], 6, (0.013, 0.487, 0.021)),
    (7.0, 80, ["mirror", "glass", "rough", "diffuse", "light", "both"], 5, (0.26, 0.37, 0.16)),
    (9.0, 40, ["glass"], 8, (0.5, 0.5, 0.5)),  # dyadic origin, camera possibly inside an entity: start substance
    (13.0, 60, ["glass", "glass", "light"], 8, (0.5, 0.5, 0.5)),  # total internal reflections
])
def test_python_restatement_equals_oracle(oracle, seed, n, kinds, refmax, pos):
    s, ents, sky, air = build(oracle, seed, n, kinds)
    W = H = 20
    cam = oracle.Camera(math.pi / 2, math.pi / 2, W, H, pos, 0.1, math.pi / 6 + 0.3, vertical_locked=True)
    cfg = dict(refmax=refmax, sky=(0.2, 0.2, 0.7), default_substance=1.0, attenuation=1.0)
    prgb, pids = pyref.render(s, ents, cam, cfg, n_frames=2, seed=1.0)
    orgb, oids, _, tot = oracle.render(s, cam, refmax=refmax, sky_texture=sky, default_substance=air, fixed_extents=True, n_frames=2,
                                       rng_mode=1, seed=1.0)
    assert np.array_equal(pids, oids)
    assert np.array_equal(prgb, orgb), np.abs(prgb - orgb).max()  # bit for bit: same double expressions in the same order
    assert tot["segments"] > W * H * 2  # paths really bounce
    assert (oids >= 0).mean() > 0.1


def test_every_branch_of_ray_trace_was_taken():
    """(runs after the parametrised cases above) the scenes together reach every branch of Ray.trace()."""
    # (the acute-normal guard needs an exactly tangent ray: covered at the collision level in test_oracle_by_hand.py)
    want = {"light", "diffuse", "mirror", "box mirror", "scatter", "undefined substance", "transmission into an entity",
            "transmission into the default substance", "refraction", "total internal reflection", "both", "refmax", "sky"}
    missing = want - set(pyref.BRANCHES)
    assert not missing, (missing, pyref.BRANCHES)


def test_image_textures_and_image_sky(oracle):
    """ImageTexture.get_color (nearest texel, value / 255) through uv_map_sphere for sphere hits and for the sky;
    boxes map to (0, 0) (BoxEntity.map_uv is a stub in the reference: texel 0 of the image)."""
    rng = np.random.default_rng(31)
    s = oracle.Scene((0, 0, 0), 1.0)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for w, h in ((64, 32), (17, 9), (128, 64))]
    tids = [s.add_texture_image(im.shape[1], im.shape[0], im) for im in imgs]
    pimgs = [pyref.Image(im) for im in imgs]
    mats = [dict(response=0, light=False, mirror=True, roughness=0.0), dict(response=0, light=False, mirror=False, roughness=0.0),
            dict(response=0, light=True, mirror=False, roughness=0.0)]
    mids = [s.add_material(m["response"], m["light"], m["mirror"], m["roughness"]) for m in mats]
    air = s.add_substance(1.0)
    ents = []
    for i in range(70):
        d = float(rng.uniform(0.05, 0.2))
        c = (d / 2 + rng.uniform(0, 1, 3) * (1 - d)).tolist()
        typ = int(rng.random() < 0.25)
        k, mi = int(rng.integers(0, 2)), int(rng.integers(0, 3))
        assert s.add_entity(typ, c, d, mids[mi], tids[k], air, max_in_depth=16, max_out_depth=0) == i
        ents.append(pyref.Ent(typ, c, d, mats[mi], pimgs[k], 1.0))
    W = H = 24
    cam = oracle.Camera(math.pi / 2, math.pi / 2, W, H, (0.5013, 0.4987, 0.5021), -0.2, 1.1, vertical_locked=True)
    cfg = dict(refmax=4, sky=pimgs[2], default_substance=1.0, attenuation=1.0)
    prgb, pids = pyref.render(s, ents, cam, cfg, n_frames=1)
    orgb, oids, _, tot = oracle.render(s, cam, refmax=4, sky_texture=tids[2], default_substance=air, fixed_extents=True, n_frames=1,
                                       rng_mode=1, seed=1.0)
    assert np.array_equal(pids, oids) and np.array_equal(prgb, orgb), np.abs(prgb - orgb).max()
    assert tot["texture_errors"] == 0 and (oids < 0).any() and (oids >= 0).any()
    assert len(np.unique(orgb.reshape(-1, 3), axis=0)) > 100  # really textured
