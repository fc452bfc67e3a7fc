"""Single-GPU timing of ONE rank's share of the headline frame for world = 1, 2, 4, 8 (no barrier, local
frame): separates render granularity / fixed per-frame cost from the cost of the cross-GPU part."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
import raytracer_js_b200 as rt
from raytracer_js_b200 import _native as N
from raytracer_js_b200 import scenes

lib = N.load()
ctx = C.c_void_p()
N.check(None, lib.rt_create(0, C.byref(ctx)))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
N.check(ctx, lib.rt_set_stream(ctx, C.c_void_p(stream.cuda_stream)))
b = bench.build_bundle()
flat = rt.flatten_scene(b.tree, extra_textures=[b.sky.texture], extra_substances=[b.default_substance])
d = flat.desc()
N.check(ctx, lib.rt_scene_upload(ctx, C.byref(d)))
W, H = bench.WIDTH, bench.HEIGHT
cd = rt.camera_desc(scenes.bench_camera(W, H))
prm = N.Params()
prm.refmax, prm.sky_texture, prm.default_substance = 1, flat.texture_index(b.sky.texture), flat.substance_index(b.default_substance)
prm.distance_attenuation_factor, prm.n_frames, prm.frame_first, prm.rng_seed = 1.0, 1, 0, 1.0
frame = torch.zeros(W * H * 3, dtype=torch.float32, device="cuda")
for world in (1, 2, 4, 8):
    for rank in sorted({0, world - 1}):
        ms = []
        for i in range(13):
            N.check(ctx, lib.rt_flush_l2(ctx))
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            N.check(ctx, lib.rt_render_shard_device(ctx, C.byref(cd), C.byref(prm), 0, rank, world, C.c_void_p(frame.data_ptr()), None))
            e.record(stream)
            torch.cuda.synchronize()
            if i >= 3:
                ms.append(a.elapsed_time(e))
        ms.sort()
        print(f"world {world} rank {rank}: median {ms[len(ms)//2]*1e3:.1f} us  min {ms[0]*1e3:.1f} us", flush=True)
