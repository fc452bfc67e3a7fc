"""GPU parity tests proper: the CUDA path, called through the C ABI (GpuRaytracer -> rt_render),
against the oracle on the same seeded scenes and cameras.  Gate (BASELINE.json): primary-hit entity
ids equal on >= 99.99 % of pixels, RGB within 1/255 per channel on id-equal pixels (relative to
max(1,|ref|) for over-range light pixels).  Parity is asserted on square frames (the reference
throws on non-square ones); non-square frames use the intent mapping on both sides."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle as orc
import raytracer_js_b200 as rt
from raytracer_js_b200 import _native as N
from raytracer_js_b200 import scenes

from util import assert_parity, compare, flat_of, insertion_ids, make_params, oracle_render, oracle_scene

pytestmark = pytest.mark.gpu


def gpu_render(bundle, width, height, n_frames=1, pos=scenes.BENCH_CAMERA_POS, yaw=30.0, pitch=0.0, refmax=None,
               reference_extents=False, frames_as_calls=False, precision=N.RT_PRECISION_F32):
    """Renders the frame twice through rt_render: (1) the default path, the two-stage pipeline (packet
    primary stage + bounce stage); (2) the counting variant, which walks every ray one by one and returns
    the reference-pattern work counters.  Both must give the same pixels; the pipeline's are returned."""
    cam = scenes.bench_camera(width, height, pos, yaw, pitch)
    cfg = rt.RaytracerConfig(bundle.refmax if refmax is None else refmax, bundle.sky, bundle.default_substance, 1.0)
    out = []
    for want_counters in (False, True):
        eb = rt.ExposureBuffer(width, height)
        tracer = rt.GpuRaytracer(cfg, bundle.tree, cam, eb, rt.FpLcg(1.0), reference_extents=reference_extents, precision=precision)
        assert tracer.lib.rt_launch_count(tracer.ctx) == 0
        if frames_as_calls:
            for f in range(n_frames):  # the reference's own loop: tick(); next_frame(); tick(); ...
                if f:
                    eb.next_frame()
                tracer.trace_frame(want_ids=True, want_counters=want_counters)
        else:
            tracer.trace_frame(n_frames=n_frames, want_ids=True, want_counters=want_counters)
        assert tracer.lib.rt_launch_count(tracer.ctx) >= 1  # our kernels ran, not a fallback
        out.append((eb.image().copy(), tracer.last_first_ids.copy(), tracer.last_counters, tracer))
    (rgb, ids, _, tracer), (rgb_c, ids_c, cnt, _) = out
    np.testing.assert_array_equal(ids, ids_c)
    np.testing.assert_array_equal(rgb, rgb_c)
    return rgb, insertion_ids(tracer.flat, bundle, ids), cnt, tracer


def ocam_for(width, height, pos=scenes.BENCH_CAMERA_POS, yaw=30.0, pitch=0.0):
    return orc.Camera(math.pi / 2, math.pi / 2, width, height, pos, pitch, math.pi / 180 * yaw, vertical_locked=True)


def oracle_for(tracer, bundle, width, height, n_frames=1, pos=scenes.BENCH_CAMERA_POS, yaw=30.0, pitch=0.0,
               refmax=None, want_counters=False):
    flat = tracer.flat
    ocam = ocam_for(width, height, pos, yaw, pitch)
    prm = make_params(flat, bundle, n_frames=n_frames, refmax=refmax)
    return oracle_render(oracle_scene(flat, bundle), ocam, flat, bundle, prm, fixed_extents=True,
                         want_counters=want_counters)


@pytest.mark.parametrize("W,H,pos,yaw,pitch", [
    (96, 96, scenes.BENCH_CAMERA_POS, 30.0, 0.0),
    (257, 131, (0.2, 0.7, 0.4), -75.0, 0.4),
    (1080, 1080, scenes.BENCH_CAMERA_POS, 30.0, 0.0),
    (1920, 1080, scenes.BENCH_CAMERA_POS, 30.0, 0.0),
    (7680, 64, scenes.DEMO_CAMERA_POS, 30.0, 0.0),  # rows of an 8K frame: 3840 iterated rotations per half row
])
def test_ray_generation_bit_for_bit(oracle, W, H, pos, yaw, pitch):
    """Camera.get_dir_for_each_pixel (src/view/camera.ts:207-250) on the device (rt_raygen_kernel through
    rt_camera_directions): every direction equals the oracle's generator's, bit for bit."""
    cam = scenes.bench_camera(W, H, pos, yaw, pitch)
    ocam = ocam_for(W, H, pos, yaw, pitch)
    fixed = W != H
    xy, d, n = ocam.dirs(fixed_extents=fixed)
    assert n == W * H
    want = np.zeros((H, W, 3))
    want[xy[:, 1], xy[:, 0]] = d
    lib = N.load()
    ctx = C.c_void_p()
    N.check(None, lib.rt_create(-1, C.byref(ctx)))
    got = np.zeros((H, W, 3))
    cd = rt.camera_desc(cam, reference_extents=not fixed)
    N.check(ctx, lib.rt_camera_directions(ctx, C.byref(cd), got.ctypes.data))
    assert lib.rt_launch_count(ctx) == 2  # the ray generation, and the expansion of its checkpoints to every pixel
    lib.rt_destroy(ctx)
    np.testing.assert_array_equal(got, want)


def test_config1_diffuse_spheres_square(oracle):
    """BASELINE config 1 scene (10 k spheres, diffuse + sky, refmax 1) on a 1080 x 1080 frame."""
    b = scenes.random_spheres(10000, 0.002, 0.006, seed=42.0, mix="diffuse")
    rgb, ids, cnt, tr = gpu_render(b, 1080, 1080)
    orgb, oids, _, tot = oracle_for(tr, b, 1080, 1080)
    res = compare(rgb, ids, orgb, oids)
    assert res["id_match"] >= 0.9999 and res["rgb_bad"] == 0, res
    for k in ("segments", "nodes", "tests", "shades"):
        assert abs(cnt[k] - tot[k]) <= 1e-4 * tot[k], (k, cnt[k], tot[k])
    assert (oids >= 0).sum() > 10000


def test_small_dense_scene_exact(oracle):
    b = scenes.random_spheres(1500, 0.01, 0.04, seed=42.0, mix="diffuse")
    rgb, ids, cnt, tr = gpu_render(b, 256, 256)
    orgb, oids, _, tot = oracle_for(tr, b, 256, 256)
    res = compare(rgb, ids, orgb, oids)
    assert res["id_mismatch"] == 0 and res["rgb_bad"] == 0, res
    assert {k: cnt[k] for k in ("segments", "nodes", "tests", "shades")} == {k: tot[k] for k in ("segments", "nodes", "tests", "shades")}


def test_config2_mirrors_lights_rough_multiframe(oracle):
    """Config-2 material mix (mirrors, rough mirrors with per-pixel-reseeded FpLcg, lights), 4 spp."""
    b = scenes.random_spheres(20000, 0.004, 0.012, seed=42.0, mix="mirrors", box_fraction=0.1)
    rgb, ids, cnt, tr = gpu_render(b, 400, 400, n_frames=4)
    orgb, oids, _, tot = oracle_for(tr, b, 400, 400, n_frames=4)
    assert_parity(rgb, ids, orgb, oids, scenes.BENCH_CAMERA_POS, ocam_for(400, 400))
    assert abs(cnt["segments"] - tot["segments"]) <= 1e-4 * tot["segments"]
    # n_frames in one call == the reference's call-per-frame loop with next_frame() in between
    rgb2, ids2, _, _ = gpu_render(b, 400, 400, n_frames=4, frames_as_calls=True)
    np.testing.assert_array_equal(rgb, rgb2)
    np.testing.assert_array_equal(ids, ids2)


def test_transmission_substances_boxes(oracle):
    rng = rt.FpLcg(5.0)
    tree = rt.new_entity_octree(rt.OctreeDim(rt.point(0, 0, 0), 1.0), None)
    glass = rt.SolidMaterial(rt.ResponseType.TRANSMISSION, False, False, 0)
    mirror = rt.SolidMaterial(rt.ResponseType.REFLECTION, False, True, 0)
    light = rt.SolidMaterial(rt.ResponseType.REFLECTION, True, False, 0)
    both = rt.SolidMaterial(rt.ResponseType.BOTH, False, False, 0)
    subs = [rt.SUBSTANCE_AIR, rt.SUBSTANCE_WATER, rt.SUBSTANCE_GLASS, None]
    ents = []
    for _ in range(300):
        d = 0.03 + rng.next() * 0.12
        c = [d / 2 + rng.next() * (1 - d) for _ in range(3)]
        m = [glass, glass, mirror, light, both][int(rng.next() * 5)]
        tex = rt.SolidTexture(rt.Color(0.3 + rng.next(), 0.3 + rng.next(), 0.3 + rng.next(), 1))
        cls = rt.BoxEntity if rng.next() < 0.3 else rt.SphereEntity
        e = cls(None, m, tex, subs[int(rng.next() * 4)], rt.point(*c), d)
        rt.add_entity_to_octree(tree, e, {"max_in_depth": 16, "max_out_depth": 0})
        ents.append(e)
    b = scenes.SceneBundle(tree, ents, rt.SkySphere(rt.SolidTexture(rt.Color(0.2, 0.2, 0.7, 1))), rt.SUBSTANCE_AIR, 6)
    pos = (0.013, 0.487, 0.021)
    rgb, ids, cnt, tr = gpu_render(b, 200, 200, pos=pos, yaw=10.0)
    orgb, oids, _, tot = oracle_for(tr, b, 200, 200, pos=pos, yaw=10.0)
    res = compare(rgb, ids, orgb, oids)
    assert res["id_match"] >= 0.9999 and res["rgb_bad"] == 0, res
    assert tot["within_tests"] > 0


def test_image_textures_and_sky(oracle):
    texs = [scenes.checker_texture(256, 128, seed=s) for s in (1, 2, 3)]
    b = scenes.random_spheres(3000, 0.02, 0.08, seed=9.0, mix="mirrors", textures=texs)
    b.sky = rt.SkySphere(scenes.checker_texture(512, 256, seed=4))
    rgb, ids, cnt, tr = gpu_render(b, 300, 300)
    orgb, oids, _, tot = oracle_for(tr, b, 300, 300)
    assert_parity(rgb, ids, orgb, oids, scenes.BENCH_CAMERA_POS, ocam_for(300, 300), image_textures=True)


def test_textures_decoded_from_files(oracle, tmp_path):
    """the texel pool filled from image FILES (ImageTexture.from_file, the host stand-in for the
    reference's browser-only load_image, src/texture/texture_image.ts:76-136) - a PNG with alpha, a
    BMP flipped both ways, and a file that does not decode (the reference then shades with fallback_color, :45-47) - rendered through
    the C ABI against the oracle."""
    from PIL import Image
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (64, 128, 4), dtype=np.uint8)
    Image.fromarray(a, "RGBA").save(tmp_path / "a.png")
    bimg = rng.integers(0, 256, (96, 48, 3), dtype=np.uint8)
    Image.fromarray(bimg, "RGB").save(tmp_path / "b.bmp")
    (tmp_path / "broken.png").write_bytes(b"not an image")
    texs = [rt.ImageTexture.from_file(str(tmp_path / "a.png"), rt.Color(1, 0, 1, 1)),
            rt.ImageTexture.from_file(str(tmp_path / "b.bmp"), rt.Color(0, 1, 1, 1), horizontal_flip=True, vertical_flip=True),
            rt.ImageTexture.from_file(str(tmp_path / "broken.png"), rt.Color(0.25, 0.5, 0.75, 1))]
    assert texs[0].get_size() == (128, 64) and texs[1].get_size() == (48, 96) and texs[2].get_size() is None
    b = scenes.random_spheres(2500, 0.02, 0.08, seed=11.0, mix="mirrors", textures=texs)
    b.sky = rt.SkySphere(texs[0])
    rgb, ids, cnt, tr = gpu_render(b, 256, 256, n_frames=2)
    orgb, oids, _, tot = oracle_for(tr, b, 256, 256, n_frames=2)
    assert_parity(rgb, ids, orgb, oids, scenes.BENCH_CAMERA_POS, ocam_for(256, 256), image_textures=True)
    assert (oids >= 0).mean() > 0.2


@pytest.mark.parametrize("pos,yaw,pitch", [
    (scenes.BENCH_CAMERA_POS, 0.0, 0.0),      # axis-aligned: a pixel row and a column with exactly-zero components
    ((0.5, 0.5, 0.5), 45.0, 0.3),             # on the root's centre planes (the demo pose, src/main.ts:364)
    ((0.031, 0.967, 0.5021), 20.0, -0.7),     # near a corner, looking down into the cube
    ((0.25, 0.75, 0.125), 135.0, 1.2),        # on dyadic planes of deeper levels
])
def test_packet_stage_from_many_poses(oracle, pos, yaw, pitch):
    """The packet walk must reproduce the reference's visit order for every direction-sign class."""
    b = scenes.random_spheres(6000, 0.004, 0.04, seed=21.0, mix="mirrors", box_fraction=0.15)
    rgb, ids, cnt, tr = gpu_render(b, 320, 320, n_frames=2, pos=pos, yaw=yaw, pitch=pitch)
    orgb, oids, _, tot = oracle_for(tr, b, 320, 320, n_frames=2, pos=pos, yaw=yaw, pitch=pitch)
    assert_parity(rgb, ids, orgb, oids, pos, ocam_for(320, 320, pos, yaw, pitch))
    assert (oids >= 0).mean() > 0.1


def test_camera_outside_root(oracle):
    b = scenes.random_spheres(400, 0.02, 0.2, seed=3.0, mix="diffuse")
    rgb, ids, cnt, tr = gpu_render(b, 96, 96, pos=(-0.5, 0.5, 0.5), yaw=0.0)
    orgb, oids, _, tot = oracle_for(tr, b, 96, 96, pos=(-0.5, 0.5, 0.5), yaw=0.0)
    res = compare(rgb, ids, orgb, oids)
    assert res["id_mismatch"] == 0 and res["rgb_bad"] == 0
    assert cnt["nodes"] == tot["nodes"]


def test_empty_scene_and_ragged_frame(oracle):
    """Empty octree (sky everywhere) and a frame that is not a multiple of the 16x16 tile."""
    b = scenes.random_spheres(0)
    rgb, ids, cnt, tr = gpu_render(b, 50, 50)
    assert (ids == -1).all()
    np.testing.assert_allclose(rgb, np.broadcast_to(np.float32([0.2, 0.2, 0.7]), rgb.shape))
    b = scenes.random_spheres(200, 0.05, 0.2, seed=2.0)
    rgb, ids, cnt, tr = gpu_render(b, 77, 45)  # non-square: intent mapping on both sides
    orgb, oids, _, _ = oracle_for(tr, b, 77, 45)
    res = compare(rgb, ids, orgb, oids)
    assert res["id_mismatch"] == 0 and res["rgb_bad"] == 0
    with pytest.raises(IndexError):  # the reference's behaviour on non-square frames
        gpu_render(b, 77, 45, reference_extents=True)


def test_full_size_properties():
    """1920x1080 (BASELINE config 1 frame): size-independent properties where the oracle is too slow to
    be run in full: determinism, exposure-blend linearity, every pixel written, ids within range."""
    b = scenes.random_spheres(10000, 0.002, 0.006, seed=42.0, mix="diffuse")
    rgb1, ids1, cnt1, tr = gpu_render(b, 1920, 1080)
    rgb2, ids2, cnt2, _ = gpu_render(b, 1920, 1080, n_frames=3)
    np.testing.assert_array_equal(ids1, ids2)
    np.testing.assert_allclose(rgb1, rgb2, atol=1e-6)  # no roughness: every frame is the same sample
    assert cnt1["paths"] == 1920 * 1080 and cnt2["paths"] == 3 * 1920 * 1080
    assert cnt1["segments"] == 1920 * 1080 and cnt2["tests"] == 3 * cnt1["tests"]
    assert ids1.min() >= -1 and ids1.max() < 10000 and np.isfinite(rgb1).all()
    sky = np.float32([0.2, 0.2, 0.7])
    assert (np.abs(rgb1[ids1 < 0] - sky) < 1e-6).all()  # misses are exactly the sky colour


def test_error_behaviour():
    lib = N.load()
    ctx = C.c_void_p()
    N.check(None, lib.rt_create(-1, C.byref(ctx)))
    cam = rt.camera_desc(scenes.bench_camera(16, 16))
    p = N.Params()
    p.n_frames = 1
    buf = np.zeros(16 * 16 * 3, np.float32)
    assert lib.rt_render(ctx, C.byref(cam), C.byref(p), 0, buf.ctypes.data, None, None) == N.RT_ERR_NO_SCENE
    assert b"no scene" in lib.rt_last_error(ctx)
    d = N.SceneDesc()
    assert lib.rt_scene_upload(ctx, C.byref(d)) == N.RT_ERR_INVALID  # struct_size mismatch
    flat = flat_of(scenes.random_spheres(20, 0.05, 0.1))
    good = flat.desc()
    N.check(ctx, lib.rt_scene_upload(ctx, C.byref(good)))
    p.sky_texture = 99
    assert lib.rt_render(ctx, C.byref(cam), C.byref(p), 0, buf.ctypes.data, None, None) == N.RT_ERR_INVALID
    flat.arrays["list_entity"][1] = flat.arrays["list_entity"][0]  # an entity listed twice
    bad = flat.desc()
    assert lib.rt_scene_upload(ctx, C.byref(bad)) == N.RT_ERR_INVALID
    assert b"more than one node" in lib.rt_last_error(ctx)
    lib.rt_destroy(ctx)


def test_tile_sharded_render_equals_full_frame():
    """rt_render_tiles_device for every rank of a 3-way split + rt_untile_device == rt_render_device
    (all on one GPU here; bench.py --gpus N runs one rank per GPU with an NCCL all-gather in between)."""
    import torch
    from raytracer_js_b200 import parallel
    W, H, world = 200, 120, 3
    b = scenes.random_spheres(3000, 0.01, 0.05, seed=8.0, mix="mirrors", box_fraction=0.1)
    cam = scenes.bench_camera(W, H)
    tracer = rt.GpuRaytracer(rt.RaytracerConfig(b.refmax, b.sky, b.default_substance, 1.0), b.tree, cam,
                             rt.ExposureBuffer(W, H), rt.FpLcg(1.0))
    lib, ctx = tracer.lib, tracer.ctx
    cd, prm = rt.camera_desc(cam), tracer.params(n_frames=2)
    dev = torch.device("cuda")
    full = torch.zeros(H * W * 3, dtype=torch.float32, device=dev)
    N.check(ctx, lib.rt_render_device(ctx, C.byref(cd), C.byref(prm), 0, C.c_void_p(full.data_ptr()), None))
    tpr = lib.rt_tiles_per_rank(W, H, world)
    assert tpr == parallel.tiles_per_rank(W, H, world)
    gathered = torch.zeros(world * tpr * 256 * 3, dtype=torch.float32, device=dev)
    for r in range(world):
        part = gathered[r * tpr * 256 * 3:]
        N.check(ctx, lib.rt_render_tiles_device(ctx, C.byref(cd), C.byref(prm), 0, r, world, C.c_void_p(part.data_ptr()), None))
    frame = torch.zeros_like(full)
    N.check(ctx, lib.rt_untile_device(ctx, W, H, world, C.c_void_p(gathered.data_ptr()), C.c_void_p(frame.data_ptr())))
    N.check(ctx, lib.rt_synchronize(ctx))
    torch.cuda.synchronize()
    assert torch.equal(frame, full)
    np.testing.assert_array_equal(parallel.untile_numpy(gathered.cpu().numpy(), W, H, world).reshape(-1), full.cpu().numpy())


@pytest.mark.parametrize("size,n_frames", [(128, 3), (240, 1)])
def test_config0_demo_scene(oracle, size, n_frames):
    """BASELINE config 0: the reference's own demo scene (src/main.ts:97-147,389-408, seed 0, refmax 4,
    camera on the root's centre planes) at the demo page's 128x128 (dist/test.html:10) and at 240x240,
    with the reference's own scan extents; the oracle builds the scene from ITS restatement of main.ts."""
    from test_hostsim_parity import demo_pair
    b, flat, cam, prm, orgb, oids, tot = demo_pair(size, size, n_frames)
    rgb, ids, cnt, tr = gpu_render(b, size, size, n_frames=n_frames, pos=scenes.DEMO_CAMERA_POS, reference_extents=True)
    # the demo pose sits on the root's centre planes (src/main.ts:364): the rays of the middle row run inside the
    # plane z = 0.5, the one kind of id mismatch that is allowed - and it must be classified as such
    assert_parity(rgb, ids, orgb, oids, scenes.DEMO_CAMERA_POS, ocam_for(size, size, scenes.DEMO_CAMERA_POS), fixed_extents=False)
    assert abs(cnt["segments"] - tot["segments"]) <= 1e-3 * tot["segments"]
    # the literal 320x240 of BASELINE.json is not renderable by the reference (F4): it throws, and so do we
    with pytest.raises(IndexError):
        gpu_render(b, 320, 240, pos=scenes.DEMO_CAMERA_POS, reference_extents=True)
    rgb2, ids2, _, _ = gpu_render(b, 320, 240, pos=scenes.DEMO_CAMERA_POS)  # intent mapping: renders
    assert np.isfinite(rgb2).all() and (ids2 >= 0).all()


@pytest.mark.parametrize("flags", [0, N.RT_RENDER_COUNTERS])
def test_shards_store_into_one_frame(flags):
    """rt_render_shard_device for every rank of a 3-way split, all into ONE frame-layout buffer (the
    multi-GPU path stores into rank 0's frame over NVLink; here the three shards run on one GPU) ==
    rt_render_device, and no shard touches another shard's pixels."""
    import torch
    W, H, world = 200, 120, 3
    b = scenes.random_spheres(3000, 0.01, 0.05, seed=8.0, mix="mirrors", box_fraction=0.1)
    cam = scenes.bench_camera(W, H)
    tracer = rt.GpuRaytracer(rt.RaytracerConfig(b.refmax, b.sky, b.default_substance, 1.0), b.tree, cam,
                             rt.ExposureBuffer(W, H), rt.FpLcg(1.0))
    lib, ctx = tracer.lib, tracer.ctx
    cd, prm = rt.camera_desc(cam), tracer.params(n_frames=2)
    dev = torch.device("cuda")
    full = torch.zeros(H * W * 3, dtype=torch.float32, device=dev)
    ids_full = torch.zeros(H * W, dtype=torch.int32, device=dev)
    N.check(ctx, lib.rt_render_device(ctx, C.byref(cd), C.byref(prm), 0, C.c_void_p(full.data_ptr()), C.c_void_p(ids_full.data_ptr())))
    frame = torch.full((H * W * 3,), -7.0, dtype=torch.float32, device=dev)
    ids = torch.full((H * W,), -9, dtype=torch.int32, device=dev)
    ty, tx = np.mgrid[0:H, 0:W]
    owner = torch.from_numpy(((ty // 16) * ((W + 15) // 16) + tx // 16) % world).to(dev).reshape(-1)
    for r in range(world):
        N.check(ctx, lib.rt_render_shard_device(ctx, C.byref(cd), C.byref(prm), flags, r, world,
                                                C.c_void_p(frame.data_ptr()), C.c_void_p(ids.data_ptr())))
        N.check(ctx, lib.rt_synchronize(ctx))
        done = owner <= r
        assert bool((ids[~done] == -9).all()) and bool((frame.view(-1, 3)[~done] == -7.0).all())
    assert torch.equal(frame, full) and torch.equal(ids, ids_full)


def test_repeated_device_renders_replay_a_graph():
    """Identical rt_render_device calls in a row are captured into a CUDA graph (2nd call) and replayed
    (3rd...); a different camera in between must drop the graph.  Every call must produce the frame an
    eager call produces, and the library's launch counter keeps counting the replayed kernels."""
    import torch
    W, H = 320, 200
    b = scenes.random_spheres(3000, 0.01, 0.05, seed=8.0, mix="mirrors", box_fraction=0.1)
    cam_a, cam_b = scenes.bench_camera(W, H), scenes.bench_camera(W, H, yaw_deg=75.0)
    tracer = rt.GpuRaytracer(rt.RaytracerConfig(b.refmax, b.sky, b.default_substance, 1.0), b.tree, cam_a,
                             rt.ExposureBuffer(W, H), rt.FpLcg(1.0))
    lib, ctx = tracer.lib, tracer.ctx
    prm = tracer.params(n_frames=2)
    dev = torch.device("cuda")
    out = torch.zeros(H * W * 3, dtype=torch.float32, device=dev)
    frames, launches = {}, []
    for name, cam in [("a", cam_a), ("a", cam_a), ("a", cam_a), ("a", cam_a), ("b", cam_b), ("a", cam_a), ("a", cam_a), ("a", cam_a)]:
        cd = rt.camera_desc(cam)
        out.fill_(-1.0)
        before = lib.rt_launch_count(ctx)
        N.check(ctx, lib.rt_render_device(ctx, C.byref(cd), C.byref(prm), 0, C.c_void_p(out.data_ptr()), None))
        N.check(ctx, lib.rt_synchronize(ctx))
        torch.cuda.synchronize()
        launches.append(lib.rt_launch_count(ctx) - before)
        if name in frames:
            assert torch.equal(frames[name], out), f"call with camera {name} differs from its first render"
        else:
            frames[name] = out.clone()
    assert not torch.equal(frames["a"], frames["b"])
    # every call is four launches: frame setup (control cells; per camera pose the origin-relative records, per
    # camera basis the ray generation - a replayed graph carries neither), primary stage, queue compaction (refmax > 1),
    # bounce stage
    assert launches == [4] * len(launches), launches


@pytest.mark.parametrize("max_in_depth", [20, 23])
def test_deep_trees_take_the_fallback_walker(oracle, max_in_depth):
    """An octree deeper than the bounce stage's walk stack (RT_WALK_STACK): secondary rays are searched by the
    reference-order walker instead of the lock-step one; same pixels as the oracle."""
    from test_hostsim_parity import deep_scene
    b = deep_scene(max_in_depth)
    rgb, ids, cnt, tracer = gpu_render(b, 64, 64, n_frames=2)
    flat = tracer.flat
    ocam = orc.Camera(math.pi / 2, math.pi / 2, 64, 64, scenes.BENCH_CAMERA_POS, 0.0, math.pi / 6, vertical_locked=True)
    prm = make_params(flat, b, n_frames=2)
    orgb, oids, _, tot = oracle_render(oracle_scene(flat, b, max_in_depth=max_in_depth), ocam, flat, b, prm, fixed_extents=True)
    res = compare(rgb, ids, orgb, oids)
    assert res["id_match"] >= 0.9999 and res["rgb_bad"] == 0, res


@pytest.mark.parametrize("max_in_depth", [30, 34])
def test_float64_search_renders_trees_of_any_depth(oracle, max_in_depth):
    """A tree deeper than float32 can resolve: the float32 search refuses the render call (RT_ERR_UNSUPPORTED, the
    upload is fine); RT_PRECISION_F64 - the reference's walker in float64, ray by ray - gives the oracle's frame, and
    its counting variant the oracle's counters, exactly."""
    from test_hostsim_parity import deep_scene
    deep = deep_scene(max_in_depth)
    cfg = rt.RaytracerConfig(deep.refmax, deep.sky, deep.default_substance, 1.0)
    refused = rt.GpuRaytracer(cfg, deep.tree, scenes.bench_camera(64, 64), rt.ExposureBuffer(64, 64), rt.FpLcg(1.0))
    with pytest.raises(N.RtError, match="float32 resolution") as e:
        refused.trace_frame()
    assert e.value.status == N.RT_ERR_UNSUPPORTED
    refused.close()
    rgb, ids, cnt, tracer = gpu_render(deep, 64, 64, n_frames=2, precision=N.RT_PRECISION_F64)
    flat = tracer.flat
    ocam = ocam_for(64, 64)
    prm = make_params(flat, deep, n_frames=2)
    orgb, oids, _, tot = oracle_render(oracle_scene(flat, deep, max_in_depth=max_in_depth), ocam, flat, deep, prm, fixed_extents=True,
                                       want_counters=True)
    res = compare(rgb, ids, orgb, oids)
    assert res["id_mismatch"] == 0 and res["rgb_bad"] == 0 and res["rgb_max_abs"] == 0.0, res
    for key in ("paths", "segments", "nodes", "tests", "shades"):
        assert cnt[key] == tot[key], (key, cnt, tot)


def test_float64_search_equals_float32_search(oracle):
    b = scenes.random_spheres(3000, 0.01, 0.05, seed=8.0, mix="mirrors", box_fraction=0.1)
    rgb32, ids32, _, _ = gpu_render(b, 200, 120, n_frames=2)
    rgb64, ids64, _, _ = gpu_render(b, 200, 120, n_frames=2, precision=N.RT_PRECISION_F64)
    np.testing.assert_array_equal(ids64, ids32)
    np.testing.assert_array_equal(rgb64, rgb32)


def test_resample_stage_equals_frames_in_a_row(oracle, monkeypatch):
    """From 4 exposure frames on, the frames of a rough pixel are traced as independent (pixel, frame) samples by
    the resample stage and blended in frame order: the same pixels, bit for bit, as one lane tracing them in a row
    (RT_B200_RESAMPLE=0), as the reference's call-per-frame loop, and - within the gate - as the oracle.  Also on a
    continued exposure (frame_first > 0) and for a frame count that is not a multiple of the pool geometry."""
    b = scenes.random_spheres(1200, 0.02, 0.07, seed=17.0, mix="mirrors", box_fraction=0.15)
    W = H = 160
    rgb, ids, cnt, tr = gpu_render(b, W, H, n_frames=11)
    orgb, oids, _, tot = oracle_for(tr, b, W, H, n_frames=11)
    res = compare(rgb, ids, orgb, oids)
    assert res["id_match"] >= 0.9999 and res["rgb_bad"] == 0, res
    launches = tr.lib.rt_launch_count(tr.ctx)
    monkeypatch.setenv("RT_B200_RESAMPLE", "0")
    rgb0, ids0, _, tr0 = gpu_render(b, W, H, n_frames=11)
    monkeypatch.delenv("RT_B200_RESAMPLE")
    np.testing.assert_array_equal(rgb0, rgb)
    np.testing.assert_array_equal(ids0, ids)
    rgbc, idsc, _, _ = gpu_render(b, W, H, n_frames=11, frames_as_calls=True)
    np.testing.assert_array_equal(rgbc, rgb)
    # a sample table that holds 248 pixels only: the resample queue takes several rounds (the first one reuses the
    # bounce stage's first-frame samples, the later ones trace all 11 frames)
    monkeypatch.setenv("RT_B200_SAMPLE_KIB", "64")
    rgbk, idsk, _, trk = gpu_render(b, W, H, n_frames=11)
    monkeypatch.delenv("RT_B200_SAMPLE_KIB")
    np.testing.assert_array_equal(rgbk, rgb)
    np.testing.assert_array_equal(idsk, ids)
    assert trk.lib.rt_launch_count(trk.ctx) > launches  # more rounds, more launches
    # a continued exposure: 3 frames, then 9 more through the resample stage == 12 frames in a row
    cam = scenes.bench_camera(W, H)
    cfg = rt.RaytracerConfig(b.refmax, b.sky, b.default_substance, 1.0)
    eb = rt.ExposureBuffer(W, H)
    t = rt.GpuRaytracer(cfg, b.tree, cam, eb, rt.FpLcg(1.0))
    t.trace_frame(n_frames=3)
    eb.next_frame()
    t.trace_frame(n_frames=9)
    eb2 = rt.ExposureBuffer(W, H)
    monkeypatch.setenv("RT_B200_RESAMPLE", "0")
    t2 = rt.GpuRaytracer(cfg, b.tree, cam, eb2, rt.FpLcg(1.0))
    t2.trace_frame(n_frames=12)
    np.testing.assert_array_equal(eb.pixels, eb2.pixels)


def test_banded_host_path_with_bouncing_paths(monkeypatch):
    """rt_render cuts the frame into bands only while no path bounces (refmax <= 1); RT_B200_BANDS forces it: every band
    then runs its own continuation queue, resample queue and rounds over the shared sample table - same frame."""
    b = scenes.random_spheres(1500, 0.02, 0.07, seed=19.0, mix="mirrors", box_fraction=0.1)
    W, H = 200, 150  # 10 tile rows: bands of 3 / 3 / 4
    rgb, ids, _, tr = gpu_render(b, W, H, n_frames=9)
    base = tr.lib.rt_launch_count(tr.ctx)
    monkeypatch.setenv("RT_B200_BANDS", "3")
    monkeypatch.setenv("RT_B200_SAMPLE_KIB", "32")
    rgb3, ids3, _, tr3 = gpu_render(b, W, H, n_frames=9)
    np.testing.assert_array_equal(rgb3, rgb)
    np.testing.assert_array_equal(ids3, ids)
    assert tr3.lib.rt_launch_count(tr3.ctx) > base


def test_zero_copy_host_frame(oracle):
    """rt_host_map: the kernels store the pixels straight into a mapped host buffer (the delivery every rank of a
    multi-GPU render uses for its own tiles); same frame as rt_render's copy path, also through the shard entry
    point with two ranks writing into one host frame."""
    b = scenes.random_spheres(2000, 0.01, 0.05, seed=8.0, mix="mirrors", box_fraction=0.1)
    W, H = 320, 208
    cam = scenes.bench_camera(W, H)
    eb = rt.ExposureBuffer(W, H)
    tracer = rt.GpuRaytracer(rt.RaytracerConfig(b.refmax, b.sky, b.default_substance, 1.0), b.tree, cam, eb, rt.FpLcg(1.0))
    tracer.trace_frame(n_frames=2)
    lib, ctx = tracer.lib, tracer.ctx
    cd, prm = rt.camera_desc(cam), tracer.params(n_frames=2)
    host = np.full(W * H * 3, -1.0, np.float32)
    devp = C.c_void_p()
    N.check(ctx, lib.rt_host_map(ctx, host.ctypes.data, host.nbytes, C.byref(devp)))
    try:
        assert devp.value
        N.check(ctx, lib.rt_render_device(ctx, C.byref(cd), C.byref(prm), 0, devp, None))
        N.check(ctx, lib.rt_synchronize(ctx))
        np.testing.assert_array_equal(host, eb.pixels)
        host[:] = -1.0
        for rank in (1, 0):
            N.check(ctx, lib.rt_render_shard_device(ctx, C.byref(cd), C.byref(prm), 0, rank, 2, devp, None))
        N.check(ctx, lib.rt_synchronize(ctx))
        np.testing.assert_array_equal(host, eb.pixels)
    finally:
        N.check(ctx, lib.rt_host_unregister(ctx, host.ctypes.data))


def test_pipelined_frames_equal_synchronous_frames():
    """rt_render_begin / rt_render_end (two frames in flight: the host copy of frame k overlaps the rendering of k + 1)
    deliver exactly the frames of rt_render, ids and counters included; a third begin without an end is refused."""
    b = scenes.random_spheres(3000, 0.01, 0.05, seed=8.0, mix="mirrors", box_fraction=0.1)
    W, H = 320, 200
    cams = [scenes.bench_camera(W, H, yaw_deg=30.0 + 7.0 * i) for i in range(5)]
    tracer = rt.GpuRaytracer(rt.RaytracerConfig(b.refmax, b.sky, b.default_substance, 1.0), b.tree, cams[0],
                             rt.ExposureBuffer(W, H), rt.FpLcg(1.0))
    lib, ctx = tracer.lib, tracer.ctx
    prm = tracer.params(n_frames=2)
    want = []
    for cam in cams:
        rgb, ids, cnt = np.zeros(W * H * 3, np.float32), np.zeros(W * H, np.int32), N.Counters()
        cd = rt.camera_desc(cam)
        N.check(ctx, lib.rt_render(ctx, C.byref(cd), C.byref(prm), 0, rgb.ctypes.data, ids.ctypes.data, C.byref(cnt)))
        want.append((rgb, ids, cnt.as_dict()))
    bufs = [(np.full(W * H * 3, -1.0, np.float32), np.full(W * H, -7, np.int32)) for _ in cams]
    got_cnt = []
    for i, cam in enumerate(cams):
        cd = rt.camera_desc(cam)
        N.check(ctx, lib.rt_render_begin(ctx, C.byref(cd), C.byref(prm), N.RT_RENDER_COUNTERS, bufs[i][0].ctypes.data, bufs[i][1].ctypes.data))
        if i == 1:  # two in flight: a third is refused, nothing is lost
            assert lib.rt_render_begin(ctx, C.byref(cd), C.byref(prm), 0, bufs[i][0].ctypes.data, None) == N.RT_ERR_INVALID
        if i >= 1:
            cnt = N.Counters()
            N.check(ctx, lib.rt_render_end(ctx, C.byref(cnt)))
            got_cnt.append(cnt.as_dict())
    cnt = N.Counters()
    N.check(ctx, lib.rt_render_end(ctx, C.byref(cnt)))
    got_cnt.append(cnt.as_dict())
    assert lib.rt_render_end(ctx, None) == N.RT_ERR_INVALID  # nothing in flight
    for i in range(len(cams)):
        np.testing.assert_array_equal(bufs[i][0], want[i][0])
        np.testing.assert_array_equal(bufs[i][1], want[i][1])
        assert got_cnt[i] == want[i][2]
