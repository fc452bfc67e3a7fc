"""Multi-GPU inside the library (rt_create_multi): ONE process, one ctx, the reference's call site
unchanged.  The sharded frame must be the single-GPU frame bit for bit (pixels are independent: interleaved 16x16
tiles, tile t -> member t % n), through the host-buffer entry point (every member stores its tiles straight into the
caller's ExposureBuffer, zero copy), the device-resident one (members store into the first GPU's frame over peer
memory) and the present path.  On a one-GPU box the group has several members on the same device, which exercises
the same sharding, worker threads, replication and event plumbing; with more GPUs present it spans all of them."""
import ctypes as C

import numpy as np
import pytest

import raytracer_js_b200 as rt
from raytracer_js_b200 import _native as N
from raytracer_js_b200 import scenes

pytestmark = pytest.mark.gpu


def device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0, 0, 0]]  # three members on one GPU
    if n >= 2:
        lists.append(list(range(n)))
    return lists


def tracer_for(b, W, H, **kw):
    cam = scenes.bench_camera(W, H)
    eb = rt.ExposureBuffer(W, H)
    cfg = rt.RaytracerConfig(b.refmax, b.sky, b.default_substance, 1.0)
    return rt.GpuRaytracer(cfg, b.tree, cam, eb, rt.FpLcg(1.0), **kw), eb, cam


@pytest.mark.parametrize("n_frames", [1, 9])
def test_group_frame_equals_single_gpu_frame(n_frames):
    b = scenes.random_spheres(3000, 0.01, 0.05, seed=8.0, mix="mirrors", box_fraction=0.1)
    W, H = 328, 200  # ragged: not a multiple of the tile
    single, eb1, _ = tracer_for(b, W, H)
    single.trace_frame(n_frames=n_frames, want_ids=True, want_counters=True)
    want_ids, want_counters = single.last_first_ids.copy(), dict(single.last_counters)
    assert single.lib.rt_group_size(single.ctx) == 1
    continued = rt.ExposureBuffer(W, H)  # n_frames frames, next_frame(), two more: on one GPU
    single.set_ebuffer(continued)
    single.trace_frame(n_frames=n_frames)
    continued.next_frame()
    single.trace_frame(n_frames=2)
    for devs in device_lists():
        group, eb, _ = tracer_for(b, W, H, devices=devs)
        assert group.lib.rt_group_size(group.ctx) == len(devs)
        group.trace_frame(n_frames=n_frames, want_ids=True, want_counters=True)  # counting variant, sharded
        np.testing.assert_array_equal(eb.pixels, eb1.pixels)
        np.testing.assert_array_equal(group.last_first_ids, want_ids)
        assert group.last_counters == want_counters
        eb2 = rt.ExposureBuffer(W, H)
        group.set_ebuffer(eb2)
        group.trace_frame(n_frames=n_frames, want_ids=True)  # the pipeline, sharded
        np.testing.assert_array_equal(eb2.pixels, eb1.pixels)
        np.testing.assert_array_equal(group.last_first_ids, want_ids)
        # a continued exposure reads the ExposureBuffer back through the mapping
        eb2.next_frame()
        group.trace_frame(n_frames=2)
        np.testing.assert_array_equal(eb2.pixels, continued.pixels)
        group.close()


def test_group_device_frame_and_present():
    """rt_render_device / rt_render_present of a group: the frame lives on the first GPU, the other members store
    their tiles into it (peer memory), the first GPU's stream waits for their events."""
    import torch
    b = scenes.random_spheres(2000, 0.01, 0.05, seed=5.0, mix="mirrors")
    W, H = 320, 208
    single, _, cam = tracer_for(b, W, H)
    cd, prm = rt.camera_desc(cam), single.params(n_frames=3)
    dev = torch.device("cuda", 0)
    want = torch.zeros(H * W * 3, dtype=torch.float32, device=dev)
    N.check(single.ctx, single.lib.rt_render_device(single.ctx, C.byref(cd), C.byref(prm), 0, C.c_void_p(want.data_ptr()), None))
    N.check(single.ctx, single.lib.rt_synchronize(single.ctx))
    tone = N.Tone(N.RT_TONE_STDDEV, 8, 1.0 / 256, 8.0)
    img1 = np.zeros(W * H * 4, np.uint8)
    N.check(single.ctx, single.lib.rt_render_present(single.ctx, C.byref(cd), C.byref(prm), 0, C.byref(tone), img1.ctypes.data, None, None))
    for devs in device_lists():
        group, _, _ = tracer_for(b, W, H, devices=devs)
        lib, ctx = group.lib, group.ctx
        got = torch.full((H * W * 3,), -1.0, dtype=torch.float32, device=dev)
        for _ in range(3):  # repeated calls: the members' graph caches and events are reused
            got.fill_(-1.0)
            torch.cuda.synchronize()
            N.check(ctx, lib.rt_render_device(ctx, C.byref(cd), C.byref(prm), 0, C.c_void_p(got.data_ptr()), None))
            N.check(ctx, lib.rt_synchronize(ctx))
            assert torch.equal(got, want)
        cnt = N.Counters()
        N.check(ctx, lib.rt_render_device(ctx, C.byref(cd), C.byref(prm), N.RT_RENDER_COUNTERS, C.c_void_p(got.data_ptr()), None))
        N.check(ctx, lib.rt_get_counters(ctx, C.byref(cnt)))
        assert cnt.paths == W * H * 3 and torch.equal(got, want)
        img = np.zeros(W * H * 4, np.uint8)
        N.check(ctx, lib.rt_render_present(ctx, C.byref(cd), C.byref(prm), 0, C.byref(tone), img.ctypes.data, None, None))
        np.testing.assert_array_equal(img, img1)
        group.close()


def test_group_errors():
    lib = N.load()
    ctx = C.c_void_p()
    assert lib.rt_create_multi(0, None, C.byref(ctx)) == N.RT_ERR_INVALID
    assert lib.rt_create_multi(1000, None, C.byref(ctx)) == N.RT_ERR_INVALID
    bad = (C.c_int32 * 2)(0, 999)
    assert lib.rt_create_multi(2, bad, C.byref(ctx)) == N.RT_ERR_INVALID and not ctx.value
    one = (C.c_int32 * 1)(0)
    N.check(None, lib.rt_create_multi(1, one, C.byref(ctx)))  # n_gpus == 1 is a plain ctx
    assert lib.rt_group_size(ctx) == 1
    lib.rt_destroy(ctx)
