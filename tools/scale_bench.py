"""Multi-GPU timing of the non-headline BASELINE configs (tile-sharded, one process per GPU under torchrun).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/scale_bench.py c4 --spp 8
  python tools/scale_bench.py c2            (N = 1)

Same data path as bench.py at N > 1: the scene is built on rank 0 (bulk generator + native octree builder),
its flat buffers are broadcast with NCCL, every rank renders its interleaved 16x16 tiles and stores them
straight into rank 0's frame over NVLink (CUDA IPC peer mapping); a flag barrier closes the frame.  Timing:
per-step CUDA events on the launching stream, max over ranks.  Prints ONE JSON line on rank 0; the frame is
the same for every N ("scaling": "strong").  --spp overrides the config's samples per pixel (say so when
quoting the number).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import torch.distributed as dist

import raytracer_js_b200 as rt
from raytracer_js_b200 import _native as N
from raytracer_js_b200 import parallel, scenes
from config_bench import CONFIGS, build


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=sorted(CONFIGS))
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--emulate", type=int, default=0, help="N = 1 only: render just rank 0's share of an N-way shard (local frame, no barrier)")
    ap.add_argument("--emulate-rank", type=int, default=0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(CONFIGS[args.config])
    if args.spp:
        cfg["spp"] = args.spp
    lib = N.load()
    ctx = C.c_void_p()
    N.check(None, lib.rt_create(local_rank, C.byref(ctx)))
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    N.check(ctx, lib.rt_set_stream(ctx, C.c_void_p(stream.cuda_stream)))

    flat = extra = None
    t_build = 0.0
    if rank == 0:
        fb, t_build = build(cfg)
        flat = fb.flat
        extra = {"refmax": fb.refmax, "sky": fb.sky_texture, "sub": fb.default_substance}
    bcast_bytes, t_bcast = 0, 0.0
    if world > 1:
        t0 = time.perf_counter()
        flat, extra, bcast_bytes = parallel.broadcast_flat_scene(flat, 0, dev, extra)
        torch.cuda.synchronize()
        t_bcast = time.perf_counter() - t0
    d = flat.desc()
    t0 = time.perf_counter()
    N.check(ctx, lib.rt_scene_upload(ctx, C.byref(d)))
    t_upload = time.perf_counter() - t0

    W, H = cfg["w"], cfg["h"]
    cd = rt.camera_desc(scenes.bench_camera(W, H))
    prm = N.Params()
    prm.refmax, prm.sky_texture, prm.default_substance = extra["refmax"], extra["sky"], extra["sub"]
    prm.distance_attenuation_factor, prm.n_frames, prm.frame_first, prm.rng_seed = 1.0, 1, 0, 1.0
    npx = W * H
    peer = None
    if world > 1:
        peer = parallel.PeerFrame(lib, ctx, rank, world, npx * 3, dst=0)
        frame_ptr = peer.frame_ptr
        frame = peer.tensor() if rank == 0 else None
    else:
        frame = torch.zeros(npx * 3, dtype=torch.float32, device=dev)
        frame_ptr = frame.data_ptr()

    def step(flags=0):
        if world == 1 and args.emulate > 1:
            N.check(ctx, lib.rt_render_shard_device(ctx, C.byref(cd), C.byref(prm), flags, args.emulate_rank, args.emulate, C.c_void_p(frame_ptr), None))
        elif world == 1:
            N.check(ctx, lib.rt_render_device(ctx, C.byref(cd), C.byref(prm), flags, C.c_void_p(frame_ptr), None))
        else:
            N.check(ctx, lib.rt_render_shard_device(ctx, C.byref(cd), C.byref(prm), flags, rank, world, C.c_void_p(frame_ptr), None))
            peer.barrier()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # work counters of one 1-spp frame (counting kernel variant): segments per path
    step(N.RT_RENDER_COUNTERS)
    cnt = N.Counters()
    N.check(ctx, lib.rt_get_counters(ctx, C.byref(cnt)))
    c = torch.tensor([cnt.paths, cnt.segments, cnt.nodes, cnt.tests, cnt.shades], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(c)
    paths1, segments1, nodes1, tests1, shades1 = (int(x) for x in c.tolist())
    prm.n_frames = cfg["spp"]
    for _ in range(args.warmup):
        step()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in evs:
        N.check(ctx, lib.rt_flush_l2(ctx))
        a.record(stream)
        step()
        b.record(stream)
    barrier()
    tot = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ms = float(tot.item()) / args.steps
    finite = bool(torch.isfinite(frame).all()) if rank == 0 else True
    if peer is not None:
        peer.close()
    if rank == 0:
        paths = npx * cfg["spp"]
        seg_per_path = segments1 / paths1
        algo = (64 * nodes1 + 16 * tests1 + 16 * shades1 + 12 * paths1) / segments1
        print(json.dumps({
            "config": args.config, "frame": f"{W}x{H}x{cfg['spp']}spp", "entities": cfg["n"], "nodes": int(d.n_nodes), "n_gpus": world, "emulate": args.emulate,
            "frame_ms": ms, "Mpaths_per_s": paths / ms / 1e3, "Mrays_per_s": paths * seg_per_path / ms / 1e3,
            "segments_per_path": seg_per_path, "algorithmic_bytes_per_segment": algo, "scaling": "strong",
            "parallelism": f"interleaved 16x16 tiles over {world} GPU(s), scene replicated (NCCL broadcast), tiles stored into rank 0's frame over NVLink",
            "scene_build_s": t_build, "scene_broadcast_s": t_bcast, "scene_broadcast_bytes": bcast_bytes, "scene_upload_s": t_upload,
            "steps": args.steps, "finite": finite}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
