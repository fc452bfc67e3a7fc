"""Kernel-body logic against the oracle WITHOUT a GPU: tests/hostsim compiles csrc/rt_trace.cuh (the
RT_HD functions the CUDA kernel calls) for the host.  Test infrastructure only; the GPU parity tests
proper are tests/test_gpu_parity.py (-m gpu)."""
import math

import numpy as np
import pytest

import oracle as orc
import raytracer_js_b200 as rt
from raytracer_js_b200 import _native as N
from raytracer_js_b200 import scenes

from util import assert_parity, compare, flat_of, hostsim_render, insertion_ids, make_params, oracle_render, oracle_scene


def cameras(width, height, pos=scenes.BENCH_CAMERA_POS, yaw=30.0, pitch=0.0):
    cam = scenes.bench_camera(width, height, pos, yaw, pitch)
    ocam = orc.Camera(fov_v=math.pi / 2, fov_h=math.pi / 2, screen_w=width, screen_h=height, pos=pos,
                      init_v_angle=pitch, init_h_angle=math.pi / 180 * yaw, vertical_locked=True)
    b = ocam.basis()
    assert list(b["fr"]) == cam.norm_fr.v and list(b["lf"]) == cam.norm_lf.v and list(b["up"]) == cam.norm_up.v
    return cam, ocam


def run_both(bundle, width, height, n_frames=1, frame_first=0, pos=scenes.BENCH_CAMERA_POS, yaw=30.0, pitch=0.0,
             refmax=None):
    flat = flat_of(bundle)
    cam, ocam = cameras(width, height, pos, yaw, pitch)
    prm = make_params(flat, bundle, n_frames=n_frames, frame_first=frame_first, refmax=refmax)
    rgb, ids, cnt = hostsim_render(flat, cam, prm)
    # the default GPU path (packet primary stage + bounce stage over the continuation queue) must give the
    # very same pixels as the ray-by-ray path
    rgb_p, ids_p, _ = hostsim_render(flat, cam, prm, pipeline=True)
    np.testing.assert_array_equal(ids_p, ids)
    np.testing.assert_array_equal(rgb_p, rgb)
    ids = insertion_ids(flat, bundle, ids)
    os_ = oracle_scene(flat, bundle)
    orgb, oids, _, tot = oracle_render(os_, ocam, flat, bundle, prm, fixed_extents=True)
    return compare(rgb, ids, orgb, oids), cnt, tot, (rgb, ids, orgb, oids)


@pytest.mark.parametrize("W,H,pos,yaw,pitch", [
    (64, 64, scenes.BENCH_CAMERA_POS, 30.0, 0.0),
    (257, 131, (0.2, 0.7, 0.4), -75.0, 0.4),       # non-square, odd sizes (intent mapping)
    (1080, 1080, scenes.BENCH_CAMERA_POS, 30.0, 0.0),
    (3840, 16, scenes.DEMO_CAMERA_POS, 30.0, 0.0),  # rows as long as a 4K frame's: 1920 iterated rotations per half row
])
def test_ray_generation_is_the_generator_bit_for_bit(oracle, W, H, pos, yaw, pitch):
    """Camera.get_dir_for_each_pixel (src/view/camera.ts:207-250): the kernel body's ray generation (row table on
    the host, iterated scan along the rows: rt_host.h rt_build_camera_rows + rt_trace.cuh raygen_half_row) must
    give the oracle's generator's direction for every pixel, every bit: floating-point rotations iterated along a
    row have no closed form with the same bits, and a last-bit difference in a camera direction is a different
    path after three bounces off millimetre spheres (tests/test_gpu_full_size.py, configs[4])."""
    import ctypes as C
    from util import hostsim
    cam, ocam = cameras(W, H, pos, yaw, pitch)
    fixed = W != H
    xy, d, n = ocam.dirs(fixed_extents=fixed)
    assert n == W * H
    want = np.zeros((H, W, 3))
    want[xy[:, 1], xy[:, 0]] = d
    cd = rt.camera_desc(cam, reference_extents=not fixed)
    got = np.zeros((H, W, 3))
    L = hostsim()
    L.hostsim_raygen.restype = None
    L.hostsim_raygen(C.byref(cd), got.ctypes.data_as(C.c_void_p))
    np.testing.assert_array_equal(got, want)
    if W == 64:  # the host mirror's generator yields the same sequence too
        mine = list(cam.get_dir_for_each_pixel())
        assert [(x, y) for x, y, _ in mine] == [tuple(p) for p in xy.tolist()]
        np.testing.assert_array_equal(np.array([v.v for _, _, v in mine]), d)


def test_diffuse_spheres_primary(oracle):
    b = scenes.random_spheres(1500, 0.01, 0.04, seed=42.0, mix="diffuse")
    res, cnt, tot, _ = run_both(b, 96, 96)
    assert res["id_mismatch"] == 0 and res["rgb_bad"] == 0, res
    assert res["rgb_max_abs"] < 1e-6
    # the work counters the roofline is built from are the oracle's
    for k in ("segments", "nodes", "tests", "shades"):
        assert cnt[k] == tot[k], (k, cnt[k], tot[k])
    assert tot["would_throw"] == 0


def test_mirrors_lights_rough_boxes_multiframe(oracle):
    b = scenes.random_spheres(800, 0.02, 0.08, seed=11.0, mix="mirrors", box_fraction=0.25)
    res, cnt, tot, (rgb, ids, orgb, oids) = run_both(b, 80, 80, n_frames=3)
    assert res["id_match"] >= 0.9999 and res["rgb_bad"] == 0, res
    assert cnt["segments"] == tot["segments"] and cnt["shades"] == tot["shades"]
    assert (oids >= 0).mean() > 0.3  # the scene is actually hit


def test_camera_outside_root_sees_only_root_list(oracle):
    # origin outside the root cube => only the root's own list is tested
    b = scenes.random_spheres(400, 0.02, 0.2, seed=3.0, mix="diffuse")
    res, cnt, tot, (rgb, ids, orgb, oids) = run_both(b, 48, 48, pos=(-0.5, 0.5, 0.5), yaw=0.0)
    assert res["id_mismatch"] == 0 and res["rgb_bad"] == 0
    root_entities = {i for i, e in enumerate(b.entities) if e.octree is b.tree}
    assert set(np.unique(oids[oids >= 0]).tolist()) <= root_entities
    assert cnt["nodes"] == tot["nodes"]


def test_transmission_and_substances(oracle):
    """TRANSMISSION materials, entity_at_pos lookups, undefined substances (src/raytracer.ts:238-249)."""
    rng = rt.FpLcg(5.0)
    tree = rt.new_entity_octree(rt.OctreeDim(rt.point(0, 0, 0), 1.0), None)
    glass = rt.SolidMaterial(rt.ResponseType.TRANSMISSION, False, False, 0)
    mirror = rt.SolidMaterial(rt.ResponseType.REFLECTION, False, True, 0)
    light = rt.SolidMaterial(rt.ResponseType.REFLECTION, True, False, 0)
    both = rt.SolidMaterial(rt.ResponseType.BOTH, False, False, 0)
    subs = [rt.SUBSTANCE_AIR, rt.SUBSTANCE_WATER, rt.SUBSTANCE_GLASS, None]
    ents = []
    for i in range(300):
        d = 0.03 + rng.next() * 0.12
        c = [d / 2 + rng.next() * (1 - d) for _ in range(3)]
        m = [glass, glass, mirror, light, both][int(rng.next() * 5)]
        tex = rt.SolidTexture(rt.Color(0.3 + rng.next(), 0.3 + rng.next(), 0.3 + rng.next(), 1))
        cls = rt.BoxEntity if rng.next() < 0.3 else rt.SphereEntity
        e = cls(None, m, tex, subs[int(rng.next() * 4)], rt.point(*c), d)
        rt.add_entity_to_octree(tree, e, {"max_in_depth": 16, "max_out_depth": 0})
        ents.append(e)
    b = scenes.SceneBundle(tree, ents, rt.SkySphere(rt.SolidTexture(rt.Color(0.2, 0.2, 0.7, 1))), rt.SUBSTANCE_AIR, 6)
    res, cnt, tot, _ = run_both(b, 72, 72, pos=(0.013, 0.487, 0.021), yaw=10.0)
    assert res["id_match"] >= 0.9999 and res["rgb_bad"] == 0, res
    assert cnt["segments"] == tot["segments"]
    assert tot["within_tests"] > 0


def test_image_textures_and_sky(oracle):
    texs = [scenes.checker_texture(64, 32, seed=s) for s in (1, 2, 3)]
    b = scenes.random_spheres(600, 0.03, 0.1, seed=9.0, mix="mirrors", textures=texs)
    b.sky = rt.SkySphere(scenes.checker_texture(128, 64, seed=4))
    res, cnt, tot, _ = run_both(b, 80, 80)
    assert res["id_match"] >= 0.9999 and res["rgb_bad"] == 0, res
    assert tot["texture_errors"] == 0


def test_reference_extents_flag(oracle):
    b = scenes.random_spheres(50, 0.05, 0.2, seed=1.0)
    flat = flat_of(b)
    cam = scenes.bench_camera(64, 48)
    prm = make_params(flat, b)
    with pytest.raises(IndexError):  # ExposureBuffer.check_bounds, exposure_buffer.ts:181-186
        hostsim_render(flat, cam, prm, reference_extents=True)
    rgb, ids, _ = hostsim_render(flat, cam, prm)  # the intent mapping renders non-square frames
    os_ = oracle_scene(flat, b)
    ocam = orc.Camera(math.pi / 2, math.pi / 2, 64, 48, scenes.BENCH_CAMERA_POS, 0.0, math.pi / 6, vertical_locked=True)
    with pytest.raises(IndexError):
        oracle_render(os_, ocam, flat, b, prm, fixed_extents=False)
    orgb, oids, _, _ = oracle_render(os_, ocam, flat, b, prm, fixed_extents=True)
    res = compare(rgb, insertion_ids(flat, b, ids), orgb, oids)
    assert res["id_mismatch"] == 0 and res["rgb_bad"] == 0


def test_primary_acceleration_is_exact(oracle):
    """The camera-ray fast path (origin-relative records + lock-step pre-test of the origin chain) must
    give the same pixels, ids and reference-pattern counters as the generic per-segment path."""
    b = scenes.random_spheres(3000, 0.004, 0.03, seed=5.0, mix="mirrors", box_fraction=0.1)
    flat = flat_of(b)
    cam, _ = cameras(128, 128)
    prm = make_params(flat, b)
    rgb1, ids1, c1 = hostsim_render(flat, cam, prm)
    prm.flags = 1  # test-only switch of tests/hostsim: no prim_geom, no chain pre-test
    rgb2, ids2, c2 = hostsim_render(flat, cam, prm)
    np.testing.assert_array_equal(rgb1, rgb2)
    np.testing.assert_array_equal(ids1, ids2)
    assert {k: c1[k] for k in ("segments", "nodes", "tests", "shades")} == {k: c2[k] for k in ("segments", "nodes", "tests", "shades")}


@pytest.mark.parametrize("pos,yaw,pitch", [
    (scenes.BENCH_CAMERA_POS, 30.0, 0.0),
    (scenes.BENCH_CAMERA_POS, 0.0, 0.0),      # axis-aligned: a pixel row and a column with exactly-zero components
    ((0.5, 0.5, 0.5), 45.0, 0.3),             # on the root's centre planes (the demo pose, src/main.ts:364)
    ((0.031, 0.967, 0.5021), 20.0, -0.7),     # near a corner, looking down into the cube
    ((0.25, 0.75, 0.125), 135.0, 1.2),        # on dyadic planes of deeper levels
])
def test_packet_stage_matches_oracle_from_many_poses(oracle, pos, yaw, pitch):
    """Packet walk (one lock-step octree walk per 8x4 patch, per sign class) against the oracle: the
    reference's visit order must come out for every ray of the packet, whatever the direction signs."""
    b = scenes.random_spheres(4000, 0.004, 0.05, seed=21.0, mix="mirrors", box_fraction=0.15)
    flat = flat_of(b)
    W, H = 120, 88
    cam, ocam = cameras(W, H, pos, yaw, pitch)
    prm = make_params(flat, b, n_frames=2)
    rgb, ids, cnt = hostsim_render(flat, cam, prm, pipeline=True)
    assert 0 < cnt["confirms"] < W * H  # some paths end at the first hit, some continue through the queue
    orgb, oids, _, _ = oracle_render(oracle_scene(flat, b), ocam, flat, b, prm, fixed_extents=True)
    res = compare(rgb, insertion_ids(flat, b, ids), orgb, oids)
    assert res["id_match"] >= 0.9999 and res["rgb_bad"] == 0, res
    rgb1, ids1, _ = hostsim_render(flat, cam, prm, pipeline=False)
    assert (ids1 != ids).sum() <= 1 and np.abs(rgb1 - rgb).max() <= (0 if (ids1 == ids).all() else 10)


def test_packet_stage_tile_sharded(oracle):
    """The tile-major output of the pipeline for every rank of a 3-way split == the full frame."""
    from raytracer_js_b200.parallel import untile_numpy
    b = scenes.random_spheres(1500, 0.01, 0.06, seed=4.0, mix="mirrors")
    flat = flat_of(b)
    W, H, world = 100, 52, 3
    cam, _ = cameras(W, H)
    prm = make_params(flat, b)
    full, ids, _ = hostsim_render(flat, cam, prm, pipeline=True)
    parts = [hostsim_render(flat, cam, prm, pipeline=True, tile_rank=r, tile_world=world)[0] for r in range(world)]
    np.testing.assert_array_equal(untile_numpy(np.stack(parts), W, H, world), full)


def demo_pair(width, height, n_frames, seed=0.0):
    """BASELINE config 0 on both sides: the oracle builds the demo scene from ITS restatement of
    src/main.ts:97-147,389-396, the product's host mirror from its own (scenes.demo_scene); entity ids are
    insertion indices on both sides."""
    b = scenes.demo_scene(seed)
    flat = flat_of(b)
    os_ = orc.Scene((0.0, 0.0, 0.0), 1.0)
    n, sky = os_.build_demo_scene(seed, 16)
    assert n == len(b.entities)
    ents = os_.entities()
    for mine, theirs in zip(b.entities, ents):  # the two generators agree on every draw
        assert tuple(mine.get_pos().v) == theirs["pos"]
        assert (mine.get_diameter() if isinstance(mine, rt.SphereEntity) else mine.get_size()) == theirs["extent"]
        assert isinstance(mine, rt.BoxEntity) == (theirs["type"] == 1)
        assert os_.textures()[theirs["texture"]]["color"][:3] == (mine.get_texture().color.r, mine.get_texture().color.g, mine.get_texture().color.b)
    cam, ocam = cameras(width, height, scenes.DEMO_CAMERA_POS, 30.0, 0.0)
    prm = make_params(flat, b, n_frames=n_frames)
    orgb, oids, _, tot = orc.render(os_, ocam, refmax=4, sky_texture=sky, default_substance=0,
                                    distance_attenuation_factor=1.0, fixed_extents=False, n_frames=n_frames,
                                    frame_first=0, rng_mode=1, seed=prm.rng_seed, n_threads=8)
    return b, flat, cam, prm, orgb, oids, tot


def test_config0_demo_scene(oracle):
    """The reference's own demo scene (seed 0, camera on the root's centre planes, rough enclosing box,
    lights, glass, mirrors, refmax 4) at the demo page's 128x128 (dist/test.html:10), 2 exposure frames,
    with the reference's own (swapped, square-only) scan extents."""
    b, flat, cam, prm, orgb, oids, tot = demo_pair(128, 128, 2)
    for pipeline in (False, True):
        rgb, ids, cnt = hostsim_render(flat, cam, prm, reference_extents=True, pipeline=pipeline)
        assert_parity(rgb, insertion_ids(flat, b, ids), orgb, oids, scenes.DEMO_CAMERA_POS,
                      cameras(128, 128, scenes.DEMO_CAMERA_POS)[1], fixed_extents=False)
    assert (oids >= 0).all()  # the enclosing box: every ray hits something
    assert tot["segments"] > 2 * 128 * 128 * 1.5  # mirrors and glass: paths really bounce


def test_config0_demo_scene_dyadic_tie_is_classified(oracle):
    """240 x 240: the camera sits on the root's centre planes (src/main.ts:364) and the middle row's rays have
    dz == 0 exactly, i.e. they run INSIDE the plane z = 0.5.  One pixel of that row differs from the oracle (an
    entity of the lower cell touching the plane): it must be classified as a dyadic tie, nothing else may differ."""
    b, flat, cam, prm, orgb, oids, tot = demo_pair(240, 240, 1)
    rgb, ids, cnt = hostsim_render(flat, cam, prm, reference_extents=True, pipeline=True)
    res, kinds = assert_parity(rgb, insertion_ids(flat, b, ids), orgb, oids, scenes.DEMO_CAMERA_POS,
                               cameras(240, 240, scenes.DEMO_CAMERA_POS)[1], fixed_extents=False)
    assert res["rgb_bad"] == 0 and len(kinds["dyadic_tie"]) == res["id_mismatch"] <= 2
    assert all(y == 120 or x == 120 for x, y in kinds["dyadic_tie"])


@pytest.mark.parametrize("pipeline", [False, True])
def test_shards_write_one_frame(oracle, pipeline):
    """rt_render_shard_device's contract: every rank of a 3-way split renders its tiles into the SAME
    frame-layout buffer and touches no other pixel; together they give the full frame."""
    b = scenes.random_spheres(1500, 0.01, 0.06, seed=4.0, mix="mirrors")
    flat = flat_of(b)
    W, H, world = 100, 52, 3
    cam, _ = cameras(W, H)
    prm = make_params(flat, b)
    full, ids, _ = hostsim_render(flat, cam, prm, pipeline=pipeline)
    frame = np.full((H, W, 3), -7.0, np.float32)
    for r in range(world):
        before = frame.copy()
        hostsim_render(flat, cam, prm, pipeline=pipeline, tile_rank=r, tile_world=world, shard_frame_layout=True, rgb=frame)
        ty, tx = np.mgrid[0:H, 0:W]
        mine = ((ty // 16) * ((W + 15) // 16) + tx // 16) % world == r
        np.testing.assert_array_equal(frame[~mine], before[~mine])
    np.testing.assert_array_equal(frame, full)


def deep_scene(max_in_depth):
    """A mirror scene plus a few tiny mirror spheres close to the camera, small enough to sit `max_in_depth`
    levels down: the octree is deeper than the bounce stage's walk stack allows (RT_WALK_STACK), so the
    fallback walker does the work."""
    b = scenes.random_spheres(400, 0.02, 0.08, seed=13.0, mix="mirrors", box_fraction=0.1)
    mirror = rt.SolidMaterial(rt.ResponseType.REFLECTION, False, True, 0)
    cam, _ = cameras(64, 64)
    dirs = {(x, y): v.v for x, y, v in cam.get_dir_for_each_pixel()}
    p0 = np.array(scenes.BENCH_CAMERA_POS)
    tiny = 2.0 ** -(max_in_depth - 1)  # a sphere of this diameter fits cells down to about max_in_depth - 1
    for k, (x, y) in enumerate([(20, 20), (40, 24), (32, 40), (12, 50)]):
        d = np.array(dirs[(x, y)])
        c = p0 + d / np.linalg.norm(d) * tiny * 12  # ~1/12 rad wide: a few pixels at 64 px over 90 degrees
        e = rt.SphereEntity(None, mirror, rt.SolidTexture(rt.Color(0.9, 0.4 + 0.1 * k, 0.2, 1)), rt.SUBSTANCE_AIR, rt.point(*c), tiny)
        rt.add_entity_to_octree(b.tree, e, {"max_in_depth": max_in_depth, "max_out_depth": 0})
        b.entities.append(e)
    return b


def test_trees_below_float32_resolution_are_refused():
    """Cells smaller than a float32 ulp of the coordinates cannot be searched in float32: the scene is refused
    with a clear message instead of being rendered wrongly (the reference, float64 throughout, has no limit)."""
    b = deep_scene(30)
    flat = flat_of(b)
    cam, _ = cameras(64, 64)
    with pytest.raises(RuntimeError, match="float32 resolution"):
        hostsim_render(flat, cam, make_params(flat, b))


@pytest.mark.parametrize("max_in_depth", [30, 34])
def test_float64_search_renders_trees_of_any_depth(oracle, max_in_depth):
    """RT_PRECISION_F64: the walker in the reference's own float64 expressions (walk_and_scan64).  A tree 30 or 34 levels
    deep - refused by the float32 search - gives the oracle's frame, and the walk visits exactly the oracle's nodes
    (the counters are the reference's access pattern: equal, not close)."""
    b = deep_scene(max_in_depth)
    flat = flat_of(b)
    cam, ocam = cameras(64, 64)
    prm = make_params(flat, b, n_frames=2)
    prm.precision = N.RT_PRECISION_F64
    rgb, ids, cnt = hostsim_render(flat, cam, prm)
    orgb, oids, _, tot = oracle_render(oracle_scene(flat, b, max_in_depth=max_in_depth), ocam, flat, b, prm, fixed_extents=True,
                                       want_counters=True)
    ids = insertion_ids(flat, b, ids)
    res = compare(rgb, ids, orgb, oids)
    assert res["id_mismatch"] == 0 and res["rgb_bad"] == 0 and res["rgb_max_abs"] == 0.0, res
    assert len(set(ids.ravel().tolist()) & set(range(len(b.entities) - 4, len(b.entities)))) >= 2  # the tiny spheres are seen
    for key in ("paths", "segments", "nodes", "tests", "shades"):
        assert cnt[key] == tot[key], (key, cnt, tot)


def test_float64_search_equals_float32_search_on_an_ordinary_scene(oracle):
    b = scenes.random_spheres(3000, 0.01, 0.05, seed=8.0, mix="mirrors", box_fraction=0.1)
    flat = flat_of(b)
    cam, ocam = cameras(96, 96)
    prm = make_params(flat, b, n_frames=2)
    rgb32, ids32, cnt32 = hostsim_render(flat, cam, prm)
    prm.precision = N.RT_PRECISION_F64
    rgb64, ids64, cnt64 = hostsim_render(flat, cam, prm)
    np.testing.assert_array_equal(ids64, ids32)
    np.testing.assert_array_equal(rgb64, rgb32)
    _, _, _, tot = oracle_render(oracle_scene(flat, b), ocam, flat, b, prm, fixed_extents=True, want_counters=True)
    for key in ("paths", "segments", "nodes", "tests", "shades"):
        assert cnt64[key] == tot[key], (key, cnt64, tot)


@pytest.mark.parametrize("max_in_depth", [20, 23])
def test_trees_deeper_than_the_walk_stacks(oracle, max_in_depth):
    b = deep_scene(max_in_depth)
    flat = flat_of(b)
    depth = 0
    par = flat.arrays["node_parent"]
    for i in range(len(par)):  # deepest node
        d, j = 0, i
        while par[j] >= 0:
            j = par[j]
            d += 1
        depth = max(depth, d)
    assert depth >= max_in_depth - 3, depth
    cam, ocam = cameras(64, 64)
    prm = make_params(flat, b, n_frames=2)
    rgb, ids, cnt = hostsim_render(flat, cam, prm)
    rgb_p, ids_p, _ = hostsim_render(flat, cam, prm, pipeline=True)
    np.testing.assert_array_equal(ids_p, ids)
    np.testing.assert_array_equal(rgb_p, rgb)
    orgb, oids, _, tot = oracle_render(oracle_scene(flat, b, max_in_depth=max_in_depth), ocam, flat, b, prm, fixed_extents=True)
    ids = insertion_ids(flat, b, ids)
    res = compare(rgb, ids, orgb, oids)
    assert res["id_match"] >= 0.9999 and res["rgb_bad"] == 0, res
    assert len(set(ids.ravel().tolist()) & set(range(len(b.entities) - 4, len(b.entities)))) >= 2  # the tiny spheres are seen


def _scene_in_root(root_pos, root_size, n=500):
    rng = rt.FpLcg(5.0)
    tree = rt.new_entity_octree(rt.OctreeDim(rt.point(*root_pos), root_size), None)
    mats = [rt.SolidMaterial(rt.ResponseType.REFLECTION, False, True, 0.0), rt.SolidMaterial(rt.ResponseType.REFLECTION, False, False, 0.0),
            rt.SolidMaterial(rt.ResponseType.REFLECTION, False, True, 0.5), rt.SolidMaterial(rt.ResponseType.REFLECTION, True, False, 0.0)]
    ents = []
    for _ in range(n):
        d = 0.01 + rng.next() * 0.05
        c = [d / 2 + rng.next() * (1 - d) for _ in range(3)]
        m = mats[int(rng.next() * 4)]
        tex = rt.SolidTexture(rt.Color(rng.next(), rng.next(), rng.next(), 1))
        cls = rt.BoxEntity if rng.next() < 0.2 else rt.SphereEntity
        e = cls(None, m, tex, rt.SUBSTANCE_AIR, rt.point(*[root_pos[k] + root_size * c[k] for k in range(3)]), d * root_size)
        rt.add_entity_to_octree(tree, e, {"max_in_depth": 16, "max_out_depth": 0})
        ents.append(e)
    return scenes.SceneBundle(tree, ents, rt.SkySphere(rt.SolidTexture(rt.Color(0.2, 0.2, 0.7, 1))), rt.SUBSTANCE_AIR, 4)


@pytest.mark.parametrize("root_pos,root_size", [((-1.0, -1.0, -1.0), 2.0), ((4.0, -8.0, 12.0), 4.0), ((0.25, 0.5, -0.75), 0.25)])
def test_other_dyadic_roots(oracle, root_pos, root_size):
    """Roots other than the demo's unit cube at the origin (power-of-two size at a multiple of it)."""
    b = _scene_in_root(root_pos, root_size)
    pos = tuple(root_pos[k] + root_size * scenes.BENCH_CAMERA_POS[k] for k in range(3))
    res, cnt, tot, _ = run_both(b, 80, 80, n_frames=2, pos=pos)
    assert res["id_match"] >= 0.9999 and res["rgb_bad"] == 0, res
    assert cnt["segments"] == tot["segments"]


def test_roots_that_break_index_within_parent_are_refused():
    """index_within_parent() (src/octree_space.ts:110-125, its own FIXME) takes the child index from the cell
    positions; with a root like (-3.7, 2.1, 10.3) x 5.3 rounding makes it disagree with the slot a node sits in,
    and the reference's walker steps back into the wrong octant.  The tree is refused with that explanation."""
    b = _scene_in_root((-3.7, 2.1, 10.3), 5.3)
    flat = flat_of(b)
    cam, _ = cameras(32, 32)
    with pytest.raises(RuntimeError, match="index_within_parent"):
        hostsim_render(flat, cam, make_params(flat, b))


def test_published_entity_counts_small_frames(oracle):
    """The scenes of BASELINE configs[2] (100 k spheres, config-2 material mix) with their real entity count on a
    small frame, many exposure frames: kernel body on the host == oracle.  (The full frame sizes are the -m gpu
    tests of tests/test_gpu_full_size.py.)"""
    from util import flat_params, oracle_crop, oracle_scene_flat
    cfg = scenes.BASELINE_CONFIGS["c2"]
    fb = scenes.build_config(cfg)
    W = H = 144
    cam = scenes.bench_camera(W, H)
    rgb, ids, _ = hostsim_render(fb.flat, cam, flat_params(fb, 9), pipeline=True)
    orgb, oids, tot = oracle_crop(oracle_scene_flat(fb), fb, W, H, (0, 0, W, H), 9)
    res = compare(rgb, ids, orgb, oids)
    assert res["id_mismatch"] == 0 and res["rgb_bad"] == 0, res
    assert tot["segments"] > 1.2 * tot["paths"]


def test_walk_stack_overflow_is_reported_not_dropped(oracle, tmp_path):
    """The ordered walk's stack is sized at upload for the tree's depth (rt_ordered_walk_fits); if a ray ever needed
    more, the walk does not drop nodes silently: it stops, the frame's error flag is raised (RT_ERRFLAG_STACK) and the
    render call fails.  The kernel body is built here with a 16-entry stack, so that rays of a mirror scene overflow."""
    import ctypes as C
    import subprocess
    import util
    from raytracer_js_b200 import _native as N
    lib = tmp_path / "librt_hostsim_cap16.so"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-pthread", "-shared", "-DRT_TEST_WALK_CAP=16",
                           "-Wno-unknown-pragmas", "-Wno-comment", "-o", str(lib), util.os.path.join(util.HOSTSIM_DIR, "rt_hostsim.cpp")])
    L = C.CDLL(str(lib))
    L.hostsim_render.restype = C.c_int
    b = scenes.random_spheres(3000, 0.01, 0.05, seed=8.0, mix="mirrors", box_fraction=0.1)
    flat = flat_of(b)
    cam, ocam = cameras(96, 96)
    prm = make_params(flat, b, n_frames=2)
    rgb = np.zeros((96, 96, 3), np.float32)
    ids = np.full((96, 96), -1, np.int32)
    cnt = N.Counters()
    err = C.create_string_buffer(512)
    d, cd = flat.desc(), rt.camera_desc(cam)
    st = L.hostsim_render(C.byref(d), C.byref(cd), C.byref(prm), 8, 0, 1, 1, C.c_void_p(rgb.ctypes.data), C.c_void_p(ids.ctypes.data),
                          C.byref(cnt), err, 512)
    assert st == N.RT_ERR_UNSUPPORTED and b"stack" in err.value, (st, err.value)


@pytest.mark.parametrize("W,H", [(100, 100), (77, 45), (60, 52), (1, 1), (9, 40)])
def test_frames_whose_middle_column_is_not_tile_aligned(oracle, W, H):
    """The packet stage produces its pixels' directions cooperatively from the ray-generation checkpoints
    (packet_directions): frames whose middle column falls inside an 8-pixel sub-patch (both scan directions in one
    row of a sub-patch), ragged right and bottom edges, and frames narrower than a sub-patch."""
    b = scenes.random_spheres(600, 0.03, 0.12, seed=4.0, mix="mirrors", box_fraction=0.2)
    flat = flat_of(b)
    cam, ocam = cameras(W, H)
    prm = make_params(flat, b, n_frames=2)
    rgb, ids, _ = hostsim_render(flat, cam, prm, pipeline=True)
    orgb, oids, _, tot = oracle_render(oracle_scene(flat, b), ocam, flat, b, prm, fixed_extents=True)
    res = compare(rgb, insertion_ids(flat, b, ids), orgb, oids)
    assert res["id_mismatch"] == 0 and res["rgb_bad"] == 0 and res["rgb_max_abs"] == 0.0, res
