"""Times the non-headline BASELINE configs (parity-test cases, not bench lines) on one GPU and prints one
JSON line each.  usage: python tools/config_bench.py c2 [c3 c4 ...] [--spp N] [--steps K]

  c2: 1920x1080, 16 spp, 100 k spheres d in [0.002,0.006], 70% mirrors / 15% diffuse / 10% rough / 5% lights, refmax 4
  c3: 3840x2160, 4 spp, 1 M entities (10% boxes) d in [0.0005,0.002], 4 image textures 1024x512, refmax 4
  c4: 7680x4320, 64 spp, 1 M spheres d in [0.0005,0.002], config-2 material mix, refmax 4
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import raytracer_js_b200 as rt
from raytracer_js_b200 import _native as N
from raytracer_js_b200 import scenes

CONFIGS = scenes.BASELINE_CONFIGS


def build(cfg):
    t0 = time.perf_counter()
    fb = scenes.build_config(cfg)
    return fb, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+")
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--per-ray", action="store_true")
    args = ap.parse_args()
    lib = N.load()
    ctx = C.c_void_p()
    N.check(None, lib.rt_create(0, C.byref(ctx)))
    for name in args.configs:
        cfg = dict(CONFIGS[name])
        if args.spp:
            cfg["spp"] = args.spp
        fb, t_build = build(cfg)
        d = fb.flat.desc()
        t0 = time.perf_counter()
        N.check(ctx, lib.rt_scene_upload(ctx, C.byref(d)))
        t_upload = time.perf_counter() - t0
        W, H = cfg["w"], cfg["h"]
        cd = rt.camera_desc(scenes.bench_camera(W, H))
        prm = N.Params()
        prm.refmax, prm.sky_texture, prm.default_substance = fb.refmax, fb.sky_texture, fb.default_substance
        prm.distance_attenuation_factor, prm.n_frames, prm.frame_first, prm.rng_seed = 1.0, 1, 0, 1.0
        prm.flags = 2 if args.per_ray else 0
        frame = torch.zeros(W * H * 3, dtype=torch.float32, device="cuda")
        # work counters of one frame (counting variant, 1 spp): segments per path
        N.check(ctx, lib.rt_render_device(ctx, C.byref(cd), C.byref(prm), N.RT_RENDER_COUNTERS, C.c_void_p(frame.data_ptr()), None))
        cnt = N.Counters()
        N.check(ctx, lib.rt_get_counters(ctx, C.byref(cnt)))
        c1 = cnt.as_dict()
        prm.n_frames = cfg["spp"]
        ms = []
        for i in range(args.steps + 1):
            el = C.c_float()
            N.check(ctx, lib.rt_flush_l2(ctx))
            N.check(ctx, lib.rt_timer_start(ctx))
            N.check(ctx, lib.rt_render_device(ctx, C.byref(cd), C.byref(prm), 0, C.c_void_p(frame.data_ptr()), None))
            N.check(ctx, lib.rt_timer_stop(ctx, C.byref(el)))
            if i:
                ms.append(el.value)
        ms.sort()
        med = ms[len(ms) // 2]
        # live duration of every stage of one more frame (CUDA events between the kernels)
        N.check(ctx, lib.rt_set_profiling(ctx, 1))
        cd_moved = rt.camera_desc(scenes.bench_camera(W, H, yaw_deg=30.01))  # (a repeated identical call would replay a graph: no stage events)
        N.check(ctx, lib.rt_render_device(ctx, C.byref(cd_moved), C.byref(prm), 0, C.c_void_p(frame.data_ptr()), None))
        st = (C.c_float * 5)()
        N.check(ctx, lib.rt_stage_times(ctx, st))
        N.check(ctx, lib.rt_set_profiling(ctx, 0))
        stages = dict(zip(("setup", "primary", "queue", "bounce", "resample"), [round(float(v), 4) for v in st]))
        seg_per_path = c1["segments"] / c1["paths"]
        paths = W * H * cfg["spp"]
        algo = 64 * c1["nodes"] + 16 * c1["tests"] + 16 * c1["shades"] + 12 * c1["paths"]
        print(json.dumps({"config": name, "frame": f"{W}x{H}x{cfg['spp']}spp", "entities": cfg["n"], "nodes": int(d.n_nodes),
                          "frame_ms": med, "stage_ms": stages, "Mpaths_per_s": paths / med / 1e3, "Mrays_per_s": paths * seg_per_path / med / 1e3,
                          "segments_per_path": seg_per_path, "nodes_per_segment": c1["nodes"] / c1["segments"],
                          "tests_per_segment": c1["tests"] / c1["segments"], "algorithmic_bytes_per_segment": algo / c1["segments"],
                          "algorithmic_GBps": algo / c1["paths"] * paths / med / 1e6, "path": "per-ray" if args.per_ray else "pipeline",
                          "scene_build_s": t_build, "scene_upload_s": t_upload, "finite": bool(torch.isfinite(frame).all())}), flush=True)


if __name__ == "__main__":
    main()
