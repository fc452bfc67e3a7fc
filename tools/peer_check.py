"""2+ GPU check of the peer-memory frame path (run under torchrun, one rank per GPU):
PeerFrame creation (CUDA IPC), the flag barrier, rt_render_shard_device into rank 0's frame == rt_render_device."""
import ctypes as C
import faulthandler
import os
import sys
import time

faulthandler.dump_traceback_later(60, exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import raytracer_js_b200 as rt
from raytracer_js_b200 import _native as N
from raytracer_js_b200 import parallel, scenes

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])


def say(*a):
    print(f"[rank {rank} {time.strftime('%X')}]", *a, flush=True)


torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = N.load()
ctx = C.c_void_p()
N.check(None, lib.rt_create(local, C.byref(ctx)))
say("ctx created")
W, H = 640, 360
pf = parallel.PeerFrame(lib, ctx, rank, world, W * H * 3, dst=0)
say("peer frame mapped", hex(pf.frame_ptr), [hex(p or 0) for p in pf.flag_ptrs])
for i in range(3):
    pf.barrier()
    N.check(ctx, lib.rt_synchronize(ctx))
    say("barrier", i, "done")
b = scenes.random_spheres(3000, 0.01, 0.05, seed=8.0, mix="mirrors", box_fraction=0.1)
flat = rt.flatten_scene(b.tree, extra_textures=[b.sky.texture], extra_substances=[b.default_substance])
d = flat.desc()
N.check(ctx, lib.rt_scene_upload(ctx, C.byref(d)))
cam = rt.camera_desc(scenes.bench_camera(W, H))
prm = N.Params()
prm.refmax, prm.sky_texture, prm.default_substance = 4, flat.texture_index(b.sky.texture), flat.substance_index(b.default_substance)
prm.distance_attenuation_factor, prm.n_frames, prm.frame_first, prm.rng_seed = 1.0, 2, 0, 1.0
for flags in (0, N.RT_RENDER_COUNTERS):
    N.check(ctx, lib.rt_render_shard_device(ctx, C.byref(cam), C.byref(prm), flags, rank, world, C.c_void_p(pf.frame_ptr), None))
    pf.barrier()
    N.check(ctx, lib.rt_synchronize(ctx))
    say("shard rendered, flags", flags)
    if rank == 0:
        ref = torch.zeros(W * H * 3, dtype=torch.float32, device=dev)
        N.check(ctx, lib.rt_render_device(ctx, C.byref(cam), C.byref(prm), 0, C.c_void_p(ref.data_ptr()), None))
        N.check(ctx, lib.rt_synchronize(ctx))
        got = pf.tensor()
        say("equal to the single-GPU frame:", bool(torch.equal(ref, got)), "nonzero:", int((got != 0).sum()))
        assert torch.equal(ref, got)
    pf.barrier()
    N.check(ctx, lib.rt_synchronize(ctx))
pf.close()
say("closed")
dist.destroy_process_group()
say("PEER CHECK OK")
