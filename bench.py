#!/usr/bin/env python
"""bench.py — Mrays/s and frame ms of the per-pixel ray hot path on BASELINE.json's headline config:
1920x1080, 1 spp, 10 000 random spheres (d in [0.002,0.006], FpLcg seed 42), octree from the host
builder, diffuse + sky, refmax 1.  A "step" is one trace_frame() of that frame.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU): the frame is cut into interleaved 16x16 tiles, every
rank renders its tiles from a replica of the scene (flat buffers broadcast from rank 0 with NCCL), the
tile buffers are all-gathered over NCCL/NVLink and rank 0 de-interleaves them ("scaling": "strong").
Prints ONE JSON line on rank 0.  See DESIGN.md §Measurement for every field.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 1920, 1080
N_SPHERES, DMIN, DMAX, SCENE_SEED = 10000, 0.002, 0.006, 42.0
WORKLOAD = "configs[1]: 1920x1080, 1 spp, 10k random spheres d in [0.002,0.006] (FpLcg seed 42), octree, diffuse + sky, refmax 1"
METRIC = "Mrays/s (ray segments traced per second), 1080p 10k-sphere octree scene"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_bundle():
    from raytracer_js_b200 import scenes
    return scenes.random_spheres(N_SPHERES, DMIN, DMAX, seed=SCENE_SEED, mix="diffuse")


def oracle_setup(bundle, flat, n_frames=1):
    """cpu_baseline / --impl reference only: the CPU restatement of the reference on the same scene."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as orc
    from raytracer_js_b200 import scenes
    from util import make_params, oracle_scene
    orc.build()
    oscene = oracle_scene(flat, bundle)
    ocam = orc.Camera(math.pi / 2, math.pi / 2, WIDTH, HEIGHT, scenes.BENCH_CAMERA_POS, 0.0, math.pi / 6,
                      vertical_locked=True)
    prm = make_params(flat, bundle, n_frames=n_frames)
    return orc, oscene, ocam, prm


def time_oracle(bundle, flat, threads: int, steps: int, warmup: int, keep_frame: bool = False):
    from util import oracle_render  # noqa: E402  (tests/ is on sys.path after oracle_setup)
    orc, oscene, ocam, prm = oracle_setup(bundle, flat)
    for _ in range(warmup):
        oracle_render(oscene, ocam, flat, bundle, prm, fixed_extents=True, n_threads=threads)
    t0 = time.perf_counter()
    seg = 0
    for _ in range(steps):
        rgb, ids, _, tot = oracle_render(oscene, ocam, flat, bundle, prm, fixed_extents=True, n_threads=threads)
        seg += tot["segments"]
    dt = time.perf_counter() - t0
    if keep_frame:
        return seg / dt / 1e6, dt / steps * 1e3, tot, rgb, ids
    return seg / dt / 1e6, dt / steps * 1e3, tot


def run_reference(args, rank: int):
    """The reference's own CPU implementation of the path (no JS runtime exists here or on the GPU box,
    so this is the double-precision C++ restatement, oracle/), all host threads, same config."""
    if rank != 0:
        return
    import raytracer_js_b200 as rt
    bundle = build_bundle()
    flat = rt.flatten_scene(bundle.tree, extra_textures=[bundle.sky.texture], extra_substances=[bundle.default_substance])
    cores = os.cpu_count() or 1
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    mrays, ms, _ = time_oracle(bundle, flat, cores, args.steps, args.warmup)
    sample = f"each step = the full {WIDTH}x{HEIGHT} frame ({WIDTH * HEIGHT} primary rays), {cores} threads over pixel chunks"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "restated C++ float64 port of the TypeScript reference (oracle/), not Node"},
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def source_fingerprint() -> str:
    """sha256 over the kernel sources the shipped library is built from: profile numbers (ncu captures under
    profiles/) are only quoted by bench.py when they were taken from exactly these sources."""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "raytracer.js_b200", "csrc")
    for f in sorted(os.listdir(csrc)) + [os.path.join("..", "..", "include", "rt_b200.h")]:
        with open(os.path.normpath(os.path.join(csrc, f)), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


def committed_profile(stage: str):
    """The profiles/*_summary.json that holds `stage` ("primary", "c2_bounce", ...) and was captured from the sources
    this run was built from; failing that the last one by name, flagged stale."""
    pdir = os.path.join(ROOT, "profiles")
    best = None
    for f in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if not f.endswith("_summary.json"):
            continue
        try:
            d = json.load(open(os.path.join(pdir, f)))
        except Exception:
            continue
        if stage in d and (best is None or best[1].get("fingerprint") != source_fingerprint()):
            best = (f, d)  # the capture of THESE sources when there is one, else the last by name
    if not best:
        return None
    f, d = best
    return {"file": f"profiles/{f}", "stale": d.get("fingerprint") != source_fingerprint(), "data": d[stage]}


def configs2_record(lib, with_parity: bool):
    """configs[2] (1920x1080, 100 k spheres, 70 % mirrors / 15 % diffuse / 10 % rough / 5 % lights, refmax 4) on one
    GPU: what the bounce and resample stages do.  Frame and per-stage times at 1 and 16 spp, the committed ncu figures
    of the bounce kernel when they were captured from these sources, and (with the cpu baseline leg) the parity of a
    256 x 128 crop of the 16-spp frame against the oracle."""
    import numpy as np
    import torch
    import raytracer_js_b200 as rt
    from raytracer_js_b200 import _native as N
    from raytracer_js_b200 import scenes
    cfg = scenes.BASELINE_CONFIGS["c2"]
    fb = scenes.build_config(cfg)
    W, H = cfg["w"], cfg["h"]
    ctx = C.c_void_p()
    N.check(None, lib.rt_create(0, C.byref(ctx)))
    out = {"workload": "configs[2]: 1920x1080, 100 k spheres d in [0.002,0.006], mirrors / diffuse / rough / lights 70/15/10/5 %, refmax 4"}
    try:
        d = fb.flat.desc()
        N.check(ctx, lib.rt_scene_upload(ctx, C.byref(d)))
        prm = N.Params()
        prm.refmax, prm.sky_texture, prm.default_substance = fb.refmax, fb.sky_texture, fb.default_substance
        prm.distance_attenuation_factor, prm.n_frames, prm.frame_first, prm.rng_seed = 1.0, 1, 0, 1.0
        frame = torch.zeros(W * H * 3, dtype=torch.float32, device="cuda:0")
        cams = [rt.camera_desc(scenes.bench_camera(W, H, yaw_deg=30.0 + 0.001 * i)) for i in range(6)]
        N.check(ctx, lib.rt_render_device(ctx, C.byref(cams[0]), C.byref(prm), N.RT_RENDER_COUNTERS, C.c_void_p(frame.data_ptr()), None))
        cnt = N.Counters()
        N.check(ctx, lib.rt_get_counters(ctx, C.byref(cnt)))
        seg_per_path = cnt.segments / max(cnt.paths, 1)
        el = C.c_float()
        N.check(ctx, lib.rt_set_profiling(ctx, 1))
        for spp in (1, cfg["spp"]):
            prm.n_frames = spp
            ms, stages = [], None
            for i in range(6):  # the camera moves every frame: no graph replay
                N.check(ctx, lib.rt_flush_l2(ctx))
                N.check(ctx, lib.rt_timer_start(ctx))
                N.check(ctx, lib.rt_render_device(ctx, C.byref(cams[i]), C.byref(prm), 0, C.c_void_p(frame.data_ptr()), None))
                N.check(ctx, lib.rt_timer_stop(ctx, C.byref(el)))
                if i:
                    ms.append(el.value)
            st = (C.c_float * 5)()
            N.check(ctx, lib.rt_stage_times(ctx, st))
            stages = dict(zip(("setup", "primary", "queue", "bounce", "resample"), [round(float(v), 4) for v in st]))
            ms.sort()
            med = ms[len(ms) // 2]
            out[f"spp{spp}"] = {"frame_ms": med, "Mrays_per_s": W * H * spp * seg_per_path / med / 1e3,
                                "Mpaths_per_s": W * H * spp / med / 1e3, "stage_ms_live": stages}
        out["segments_per_path"] = seg_per_path
        N.check(ctx, lib.rt_set_profiling(ctx, 0))
        prof = committed_profile("c2_bounce")
        if prof and not prof["stale"]:
            pk = prof["data"]
            out["bounce_kernel_ncu"] = {"file": prof["file"], "kernel_ms_under_ncu": pk["gpu__time_duration.sum"],
                                        "issue_slot_utilisation_pct": pk["smsp__issue_active.avg.pct_of_peak_sustained_active"],
                                        "active_lanes_per_instruction": pk["smsp__thread_inst_executed_per_inst_executed.ratio"],
                                        "issue_x_lanes": pk["smsp__issue_active.avg.pct_of_peak_sustained_active"] / 100
                                                         * pk["smsp__thread_inst_executed_per_inst_executed.ratio"] / 32,
                                        "registers_per_thread": pk["launch__registers_per_thread"],
                                        "dram_bytes": pk["dram__bytes_read.sum"] + pk["dram__bytes_write.sum"]}
        else:
            out["bounce_kernel_ncu"] = {"stale_profile": (prof or {}).get("file", "no committed capture")}
        if with_parity:  # the oracle on a crop of the 16-spp frame (same camera, same per-pixel seeds)
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from util import classify_outliers, compare, oracle_crop, oracle_scene_flat
            prm.n_frames = cfg["spp"]
            host = np.zeros((H, W, 3), np.float32)
            ids = np.zeros((H, W), np.int32)
            cam0 = rt.camera_desc(scenes.bench_camera(W, H))
            N.check(ctx, lib.rt_render(ctx, C.byref(cam0), C.byref(prm), 0, host.ctypes.data, ids.ctypes.data, None))
            crop = (1000, 500, 256, 128)
            x, y, w, h = crop
            orgb, oids, _ = oracle_crop(oracle_scene_flat(fb), fb, W, H, crop, cfg["spp"])
            par = compare(host[y:y + h, x:x + w], ids[y:y + h, x:x + w], orgb, oids)
            kinds = classify_outliers(host[y:y + h, x:x + w], ids[y:y + h, x:x + w], orgb, oids, cam_pos=scenes.BENCH_CAMERA_POS,
                                      ocam=None, image_textures=False, offset=(x, y))
            par.update({"crop": list(crop), "spp": cfg["spp"], "hit_fraction": float((oids >= 0).mean()),
                        "unexplained_outliers": len(kinds["unexplained"]),
                        "pass": bool(par["id_match"] >= 0.9999 and par["rgb_bad"] == 0 and not kinds["unexplained"])})
            out["parity"] = par
    finally:
        lib.rt_destroy(ctx)
    return out


def group_records(args, lib, n_gpus, flat, sky_tex, def_sub, cds, warm):
    """rt_create_multi(n_gpus) in THIS process: the headline frame and a configs[4]-class frame on n_gpus GPUs."""
    import numpy as np
    import raytracer_js_b200 as rt
    from raytracer_js_b200 import _native as N
    from raytracer_js_b200 import scenes
    out = {"n_gpus": n_gpus, "how": "one process, rt_create_multi: scene packed once and replicated device to device, one worker "
                                    "thread per GPU, interleaved 16x16 tiles; device frame on GPU 0 (peer stores) / host frame mapped into every GPU"}
    g = C.c_void_p()
    N.check(None, lib.rt_create_multi(n_gpus, None, C.byref(g)))
    try:
        t0 = time.perf_counter()
        d = flat.desc()
        N.check(g, lib.rt_scene_upload(g, C.byref(d)))
        out["scene_upload_s"] = time.perf_counter() - t0
        prm = N.Params()
        prm.refmax, prm.sky_texture, prm.default_substance = 1, sky_tex, def_sub
        prm.distance_attenuation_factor, prm.n_frames, prm.frame_first, prm.rng_seed = 1.0, 1, 0, 1.0
        npx = WIDTH * HEIGHT
        import torch
        frame = torch.zeros(npx * 3, dtype=torch.float32, device="cuda:0")
        el = C.c_float()

        def timed(cam):
            N.check(g, lib.rt_flush_l2(g))
            N.check(g, lib.rt_timer_start(g))
            N.check(g, lib.rt_render_device(g, C.byref(cam), C.byref(prm), 0, C.c_void_p(frame.data_ptr()), None))
            N.check(g, lib.rt_timer_stop(g, C.byref(el)))
            return el.value

        for i in range(warm):
            timed(cds[i])
        ms = sum(timed(cds[warm + i]) for i in range(args.steps)) / args.steps
        out["headline"] = {"ms_per_step": ms, "value": npx / (ms * 1e-3) / 1e6, "unit": "Mrays/s",
                           "timing": "CUDA events on GPU 0's stream around the whole group frame (it waits for the other GPUs' events)"}
        host = np.zeros(npx * 3, np.float32)
        if n_gpus == 1:  # a group of one is a plain ctx: page-lock the frame as the e2e leg does (a group maps it itself)
            N.check(g, lib.rt_host_register(g, host.ctypes.data, host.nbytes))
        for i in range(3):
            N.check(g, lib.rt_render(g, C.byref(cds[i]), C.byref(prm), 0, host.ctypes.data, None, None))
        t0 = time.perf_counter()
        for i in range(args.steps):
            N.check(g, lib.rt_render(g, C.byref(cds[warm + i]), C.byref(prm), 0, host.ctypes.data, None, None))
        dt = time.perf_counter() - t0
        out["headline"]["e2e"] = {"value": npx * args.steps / dt / 1e6, "unit": "Mrays/s", "frame_ms": dt / args.steps * 1e3,
                                  "d2h_bytes_per_step": npx * 12,
                                  "delivery": "every GPU stores its tiles into the mapped host frame over its own PCIe link" if n_gpus > 1
                                  else "one GPU: banded copies into the page-locked host frame"}
        if n_gpus == 1:
            lib.rt_host_unregister(g, host.ctypes.data)
        del frame
        # ---- configs[4] class
        cfg = dict(scenes.BASELINE_CONFIGS["c4"])
        cfg["spp"] = 8
        t0 = time.perf_counter()
        fb = scenes.build_config(cfg)
        t_build = time.perf_counter() - t0
        t0 = time.perf_counter()
        d4 = fb.flat.desc()
        N.check(g, lib.rt_scene_upload(g, C.byref(d4)))
        t_up = time.perf_counter() - t0
        W4, H4 = cfg["w"], cfg["h"]
        cd4 = rt.camera_desc(scenes.bench_camera(W4, H4))
        p4 = N.Params()
        p4.refmax, p4.sky_texture, p4.default_substance = fb.refmax, fb.sky_texture, fb.default_substance
        p4.distance_attenuation_factor, p4.n_frames, p4.frame_first, p4.rng_seed = 1.0, 1, 0, 1.0
        frame4 = torch.zeros(W4 * H4 * 3, dtype=torch.float32, device="cuda:0")
        N.check(g, lib.rt_render_device(g, C.byref(cd4), C.byref(p4), N.RT_RENDER_COUNTERS, C.c_void_p(frame4.data_ptr()), None))
        cnt = N.Counters()
        N.check(g, lib.rt_get_counters(g, C.byref(cnt)))
        seg_per_path = cnt.segments / max(cnt.paths, 1)
        p4.n_frames = cfg["spp"]
        times = []
        for i in range(3):
            N.check(g, lib.rt_flush_l2(g))
            N.check(g, lib.rt_timer_start(g))
            N.check(g, lib.rt_render_device(g, C.byref(cd4), C.byref(p4), 0, C.c_void_p(frame4.data_ptr()), None))
            N.check(g, lib.rt_timer_stop(g, C.byref(el)))
            if i:
                times.append(el.value)
        ms4 = sum(times) / len(times)
        paths = W4 * H4 * cfg["spp"]
        out["c4_class"] = {"workload": f"{W4}x{H4}, {cfg['spp']} spp (of configs[4]'s 64), 1 M spheres d in [0.0005,0.002], mirror mix, refmax 4",
                           "frame_ms": ms4, "Mrays_per_s": paths * seg_per_path / ms4 / 1e3, "Mpaths_per_s": paths / ms4 / 1e3,
                           "segments_per_path": seg_per_path, "scene_build_s": t_build, "scene_upload_s": t_up,
                           "finite": bool(torch.isfinite(frame4).all())}
    finally:
        lib.rt_destroy(g)
    return out


def run_ours(args, rank: int, world: int, local_rank: int):
    import numpy as np
    import torch
    import torch.distributed as dist

    import raytracer_js_b200 as rt
    from raytracer_js_b200 import _native as N
    from raytracer_js_b200 import scenes

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = N.load()
    ctx = C.c_void_p()
    N.check(None, lib.rt_create(local_rank, C.byref(ctx)))
    # one explicit (non-default) stream carries our kernels, the NCCL collectives and the timing events
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    N.check(ctx, lib.rt_set_stream(ctx, C.c_void_p(stream.cuda_stream)))

    # ---- scene: built on rank 0 by the host API, flat buffers broadcast with NCCL, one replica per GPU
    bundle = flat = None
    if rank == 0:
        bundle = build_bundle()
        flat = rt.flatten_scene(bundle.tree, extra_textures=[bundle.sky.texture],
                                extra_substances=[bundle.default_substance])
        sky_tex = flat.texture_index(bundle.sky.texture)
        def_sub = flat.substance_index(bundle.default_substance)
    scene_bcast_bytes = 0
    if world > 1:
        from raytracer_js_b200 import parallel
        extra = {"sky": sky_tex, "sub": def_sub} if rank == 0 else None
        flat, extra, scene_bcast_bytes = parallel.broadcast_flat_scene(flat, 0, dev, extra)  # NCCL over NVLink
        sky_tex, def_sub = extra["sky"], extra["sub"]
    desc = flat.desc()
    N.check(ctx, lib.rt_scene_upload(ctx, C.byref(desc)))

    # The camera MOVES between steps (yaw + 0.001 degree per step), as it does in the reference's interactive
    # loop (src/main.ts:288-330): every step recomputes the scan tables and the origin-relative records, and no
    # step is a CUDA-graph replay of the one before it.  Step 0 is the pose of SURVEY.md 8d (yaw 30 degrees).
    n_cams = max(args.warmup, 3) + args.steps + 1
    cds = [rt.camera_desc(scenes.bench_camera(WIDTH, HEIGHT, yaw_deg=30.0 + 1e-3 * i)) for i in range(n_cams)]
    cd = cds[0]
    prm = N.Params()
    prm.refmax, prm.sky_texture, prm.default_substance = 1, sky_tex, def_sub
    prm.distance_attenuation_factor, prm.n_frames, prm.frame_first, prm.rng_seed = 1.0, 1, 0, 1.0
    prm.precision = N.RT_PRECISION_F32
    prm.flags = 2 if args.path == "per-ray" else 0  # RT_PARAM_PER_RAY: the round-1 ray-by-ray kernel, for A/B runs

    npx = WIDTH * HEIGHT
    peer = None
    if world > 1 and args.gather == "peer":
        # ONE frame on rank 0; every rank's kernels store their tiles into it across NVLink (CUDA IPC mapping)
        from raytracer_js_b200 import parallel
        peer = parallel.PeerFrame(lib, ctx, rank, world, npx * 3, dst=0)
        frame = peer.tensor() if rank == 0 else None
    else:
        frame = torch.zeros(npx * 3, dtype=torch.float32, device=dev)
    tpr = lib.rt_tiles_per_rank(WIDTH, HEIGHT, world)
    nccl_gather = world > 1 and peer is None
    tiles = torch.zeros(tpr * 256 * 3, dtype=torch.float32, device=dev) if nccl_gather else None
    gathered = torch.zeros(world * tpr * 256 * 3, dtype=torch.float32, device=dev) if nccl_gather else None

    def step(flags=0, cam=None):
        """One frame with everything resident in HBM."""
        cam = cd if cam is None else cam
        if world == 1:
            N.check(ctx, lib.rt_render_device(ctx, C.byref(cam), C.byref(prm), flags, C.c_void_p(frame.data_ptr()), None))
        elif peer is not None:
            # fused render + gather: pixels go straight into rank 0's frame over NVLink; the barrier closes it
            N.check(ctx, lib.rt_render_shard_device(ctx, C.byref(cam), C.byref(prm), flags, rank, world,
                                                    C.c_void_p(peer.frame_ptr), None))
            peer.barrier()
        else:
            N.check(ctx, lib.rt_render_tiles_device(ctx, C.byref(cam), C.byref(prm), flags, rank, world,
                                                    C.c_void_p(tiles.data_ptr()), None))
            dist.all_gather_into_tensor(gathered, tiles)  # tile exchange over NCCL / NVLink
            if rank == 0:
                N.check(ctx, lib.rt_untile_device(ctx, WIDTH, HEIGHT, world, C.c_void_p(gathered.data_ptr()),
                                                  C.c_void_p(frame.data_ptr())))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- work counters of one frame (untimed, counting kernel variant): what the per-ray-bytes figure is built from
    step(N.RT_RENDER_COUNTERS)
    cnt = N.Counters()
    N.check(ctx, lib.rt_get_counters(ctx, C.byref(cnt)))
    c = torch.tensor([cnt.paths, cnt.segments, cnt.nodes, cnt.tests, cnt.shades], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(c)
    paths, segments, nodes, tests, shades = (int(x) for x in c.tolist())
    assert segments == npx  # refmax 1: one ray segment per pixel, whatever the camera looks at
    # ALGORITHMIC bytes: 64 B per node returned by the walker order, 16 B per entity hit
    # test, 16 B per shaded hit, 12 B per pixel store.  The counters count the REFERENCE's access pattern
    # (tests/test_gpu_parity.py asserts they equal the oracle's), not this kernel's own traffic.
    algo_bytes = 64 * nodes + 16 * tests + 16 * shades + 12 * paths
    bytes_per_segment = algo_bytes / segments

    warm = max(args.warmup, 3)
    for i in range(warm):
        step(cam=cds[i])
    barrier()

    # ---- timed region: K steps, per-step CUDA events on the launching stream, L2 flushed between steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = lib.rt_launch_count(ctx)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i, (a, b) in enumerate(evs):
        N.check(ctx, lib.rt_flush_l2(ctx))
        if peer is not None:
            # every rank has finished its (untimed) flush before any rank's timed step begins: without this a rank's
            # step time would include the flush of the slowest other rank, which it meets at the frame's own barrier
            peer.barrier()
        elif world > 1:
            dist.barrier()
        a.record(stream)
        step(cam=cds[warm + i])
        b.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = int(lib.rt_launch_count(ctx) - l0)
    step_ms = [a.elapsed_time(b) for a, b in evs]
    tot_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(tot_ms.item()) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    value = segments / (ms_per_step * 1e-3) / 1e6

    # ---- beside it: (a) the same K steps with a camera that stands still, which the library replays as a CUDA
    # graph (no scan-table copies, no origin-relative records: NOT what `value` is); (b) the K steps again with
    # CUDA events between the kernels (rt_set_profiling): the live duration of every stage
    replay_ms = stage_ms = None
    if world == 1:
        for _ in range(3):
            step()
        revs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in revs:
            N.check(ctx, lib.rt_flush_l2(ctx))
            a.record(stream)
            step()
            b.record(stream)
        torch.cuda.synchronize()
        replay_ms = sum(a.elapsed_time(b) for a, b in revs) / args.steps
        N.check(ctx, lib.rt_set_profiling(ctx, 1))
        acc = [0.0] * 5
        st = (C.c_float * 5)()
        for i in range(args.steps):
            N.check(ctx, lib.rt_flush_l2(ctx))
            step(cam=cds[warm + i])
            N.check(ctx, lib.rt_stage_times(ctx, st))
            for k in range(5):
                acc[k] += max(0.0, st[k])
        N.check(ctx, lib.rt_set_profiling(ctx, 0))
        stage_ms = dict(zip(("prepare", "primary", "shade", "bounce", "resample"), (v / args.steps for v in acc)))

    # ---- end to end through the C ABI with HOST buffers: camera tables H2D + kernel + full frame D2H
    e2e = None
    present = None
    gpu_ids = None
    if world == 1:
        host_rgb = np.zeros(npx * 3, np.float32)
        N.check(ctx, lib.rt_host_register(ctx, host_rgb.ctypes.data, host_rgb.nbytes))  # pinned, as the N-API shim does
        for i in range(3):
            N.check(ctx, lib.rt_render(ctx, C.byref(cds[i]), C.byref(prm), 0, host_rgb.ctypes.data, None, None))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(args.steps):
            N.check(ctx, lib.rt_render(ctx, C.byref(cds[warm + i]), C.byref(prm), 0, host_rgb.ctypes.data, None, None))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        # the frame of step 0's pose through the host path, with first-hit ids: compared with the device path here
        # and with the oracle's frame below (`parity`)
        gpu_ids = np.full(npx, -9, np.int32)
        N.check(ctx, lib.rt_render(ctx, C.byref(cd), C.byref(prm), 0, host_rgb.ctypes.data, gpu_ids.ctypes.data, None))
        step()
        torch.cuda.synchronize()
        assert np.array_equal(host_rgb, frame.cpu().numpy()), "host path and device path disagree"
        N.check(ctx, lib.rt_host_unregister(ctx, host_rgb.ctypes.data))
        e2e = {"value": segments * args.steps / dt / 1e6, "unit": "Mrays/s", "frame_ms": dt / args.steps * 1e3,
               "h2d_bytes_per_step": HEIGHT * 32, "d2h_bytes_per_step": npx * 12 + 72}
        # ---- the same frames PIPELINED (rt_render_begin / rt_render_end, two ExposureBuffers): the host copy of frame k
        # overlaps the rendering of frame k + 1; every frame still arrives complete in host memory
        host2 = [np.zeros(npx * 3, np.float32), np.zeros(npx * 3, np.float32)]
        for hb in host2:
            N.check(ctx, lib.rt_host_register(ctx, hb.ctypes.data, hb.nbytes))
        for rep in range(2):  # the first pass warms the second device frame up
            t0 = time.perf_counter()
            for i in range(args.steps):
                N.check(ctx, lib.rt_render_begin(ctx, C.byref(cds[warm + i]), C.byref(prm), 0, host2[i & 1].ctypes.data, None))
                if i:
                    N.check(ctx, lib.rt_render_end(ctx, None))
            N.check(ctx, lib.rt_render_end(ctx, None))
            dtp2 = time.perf_counter() - t0
        check = np.zeros(npx * 3, np.float32)
        N.check(ctx, lib.rt_render(ctx, C.byref(cds[warm + args.steps - 1]), C.byref(prm), 0, check.ctypes.data, None, None))
        assert np.array_equal(host2[(args.steps - 1) & 1], check), "pipelined frame differs from the synchronous one"
        del check
        for hb in host2:
            N.check(ctx, lib.rt_host_unregister(ctx, hb.ctypes.data))
        e2e["pipelined"] = {"value": segments * args.steps / dtp2 / 1e6, "unit": "Mrays/s", "frame_ms": dtp2 / args.steps * 1e3,
                            "note": "rt_render_begin / rt_render_end, two frames in flight, two host buffers: throughput of a frame loop; "
                                    "`value` above is the synchronous call's (latency of one frame)"}
        # ---- the step after the path: View.draw_ebuffer() on the device.  (a) the three
        # present kernels alone on the resident frame, (b) trace_frame + draw_ebuffer end to end with the
        # ExposureBuffer resident in HBM: only the RGBA8 screen travels to the host.
        tone = N.Tone(N.RT_TONE_STDDEV, 8, 1.0 / 256, 8.0)  # the demo's mapper (src/main.ts:369)
        rgba_dev = torch.zeros(npx * 4, dtype=torch.uint8, device=dev)
        pevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for _ in range(3):
            N.check(ctx, lib.rt_present_device(ctx, C.c_void_p(frame.data_ptr()), WIDTH, HEIGHT, C.byref(tone), C.c_void_p(rgba_dev.data_ptr())))
        for a, b in pevs:
            N.check(ctx, lib.rt_flush_l2(ctx))
            a.record(stream)
            N.check(ctx, lib.rt_present_device(ctx, C.c_void_p(frame.data_ptr()), WIDTH, HEIGHT, C.byref(tone), C.c_void_p(rgba_dev.data_ptr())))
            b.record(stream)
        torch.cuda.synchronize()
        present_ms = sum(a.elapsed_time(b) for a, b in pevs) / args.steps
        host_rgba = np.zeros(npx * 4, np.uint8)
        N.check(ctx, lib.rt_host_register(ctx, host_rgba.ctypes.data, host_rgba.nbytes))
        for _ in range(3):
            N.check(ctx, lib.rt_render_present(ctx, C.byref(cd), C.byref(prm), 0, C.byref(tone), host_rgba.ctypes.data, None, None))
        t0 = time.perf_counter()
        for i in range(args.steps):
            N.check(ctx, lib.rt_render_present(ctx, C.byref(cds[warm + i]), C.byref(prm), 0, C.byref(tone), host_rgba.ctypes.data, None, None))
        dtp = time.perf_counter() - t0
        N.check(ctx, lib.rt_render_present(ctx, C.byref(cd), C.byref(prm), 0, C.byref(tone), host_rgba.ctypes.data, None, None))
        assert np.array_equal(host_rgba, rgba_dev.cpu().numpy()), "resident present and device present disagree"
        N.check(ctx, lib.rt_host_unregister(ctx, host_rgba.ctypes.data))
        # algorithmic bytes of draw_ebuffer: get_mean, get_variance and discretize_to_screen each read the
        # 12 B pixel, the screen gets 4 B (src/view/exposure_buffer.ts:93-158)
        present_bytes = npx * (3 * 12 + 4)
        present = {"kernels_ms": present_ms, "algorithmic_bytes": present_bytes, "achieved_GBps": present_bytes / (present_ms * 1e-3) / 1e9,
                   "frac_of_hbm_peak": present_bytes / (present_ms * 1e-3) / 1e9 / peaks()[0], "tone_mapper": "ToneMapper_StdDevAroundMean(8, 1/256, 8)",
                   "e2e_render_present": {"value": segments * args.steps / dtp / 1e6, "unit": "Mrays/s", "frame_ms": dtp / args.steps * 1e3,
                                          "h2d_bytes_per_step": WIDTH * 16 + HEIGHT * 32 + 80, "d2h_bytes_per_step": npx * 4 + 68,
                                          "note": "trace_frame() + View.draw_ebuffer() with the ExposureBuffer resident on the device; RGBA8 out"}}
    else:
        # multi-GPU end to end: frame assembled on rank 0 and read back to pinned host memory every step
        host_t = torch.empty(npx * 3, dtype=torch.float32, pin_memory=True) if rank == 0 else None
        if rank == 0:
            # the sharded frame must be the single-GPU frame, bit for bit
            step()
            torch.cuda.synchronize()
            ref = torch.zeros(npx * 3, dtype=torch.float32, device=dev)
            N.check(ctx, lib.rt_render_device(ctx, C.byref(cd), C.byref(prm), 0, C.c_void_p(ref.data_ptr()), None))
            torch.cuda.synchronize()
            assert torch.equal(ref, frame), "sharded frame differs from the single-GPU frame"
        else:
            step()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step(cam=cds[warm + i])
            if rank == 0:
                host_t.copy_(frame, non_blocking=True)
            if peer is not None:
                peer.barrier()  # nobody stores into the next frame before rank 0 has read this one
        barrier()
        dt = time.perf_counter() - t0
        e2e = {"value": segments * args.steps / dt / 1e6, "unit": "Mrays/s", "frame_ms": dt / args.steps * 1e3,
               "h2d_bytes_per_step": (WIDTH * 16 + HEIGHT * 32 + 80) * world, "d2h_bytes_per_step": npx * 12,
               "delivery": "frame assembled on rank 0 over NVLink, one device-to-host copy from rank 0"}
        if peer is not None:
            # zero-copy delivery: the host frame is shared memory, mapped into every rank's device address space
            # (rt_host_map); each rank's kernels store its tiles straight into it over its OWN PCIe link.  The flag
            # barrier closes the frame (system-scope fences cover the host stores), a stream sync hands it to the host.
            from multiprocessing import shared_memory
            names = [None]
            shm = None
            if rank == 0:
                shm = shared_memory.SharedMemory(create=True, size=npx * 12)
                names = [shm.name]
            dist.broadcast_object_list(names, src=0)
            if rank != 0:
                shm = shared_memory.SharedMemory(name=names[0])
                try:  # the creator unlinks it; keep this process's resource tracker from trying again at exit
                    from multiprocessing import resource_tracker
                    resource_tracker.unregister(shm._name, "shared_memory")
                except Exception:
                    pass
            host = np.ndarray((npx * 3,), np.float32, buffer=shm.buf)
            devp = C.c_void_p()
            N.check(ctx, lib.rt_host_map(ctx, host.ctypes.data, host.nbytes, C.byref(devp)))

            def step_zc(cam):
                N.check(ctx, lib.rt_render_shard_device(ctx, C.byref(cam), C.byref(prm), 0, rank, world, devp, None))
                peer.barrier()
                N.check(ctx, lib.rt_synchronize(ctx))

            for _ in range(3):
                step_zc(cd)
            barrier()
            if rank == 0:
                assert np.array_equal(host, ref.cpu().numpy()), "zero-copy host frame differs from the single-GPU frame"
            barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):
                step_zc(cds[warm + i])
            barrier()
            dtz = time.perf_counter() - t0
            zc = {"value": segments * args.steps / dtz / 1e6, "unit": "Mrays/s", "frame_ms": dtz / args.steps * 1e3,
                  "h2d_bytes_per_step": (WIDTH * 16 + HEIGHT * 32 + 80) * world, "d2h_bytes_per_step": npx * 12,
                  "delivery": "every rank stores its tiles into the mapped shared host frame over its own PCIe link (rt_host_map)"}
            N.check(ctx, lib.rt_host_unregister(ctx, host.ctypes.data))
            del host
            shm.close()
            barrier()
            if rank == 0:
                shm.unlink()
            if zc["value"] > e2e["value"]:
                zc["gather_on_rank0"] = {k: e2e[k] for k in ("value", "frame_ms", "delivery")}
                e2e = zc
            else:
                e2e["zero_copy"] = {k: zc[k] for k in ("value", "frame_ms", "delivery")}

    if peer is not None:
        peer.close()
    # ---- the in-library group (rt_create_multi): ONE process - rank 0 - drives all N GPUs behind the plain entry
    # points, the other ranks wait at the barrier with their GPUs idle.  (a) the headline frame: device-resident on
    # GPU 0 and end to end into a host buffer; (b) a configs[4]-class frame (8K, 1 M spheres, mirror mix, 8 spp) on
    # the same N GPUs, for the scaling of a frame that is long enough to scale.
    group = None
    store = None
    if world > 1:
        barrier()
        try:  # the other ranks wait on the HOST (a key in the rendezvous store): an NCCL barrier would spin on their GPUs
            store = dist.distributed_c10d._get_default_store()
        except Exception:
            store = None
    configs2 = None
    if rank == 0 and not args.no_group:
        try:
            group = group_records(args, lib, world, flat, sky_tex, def_sub, cds, warm)
        except Exception as e:  # reported, never fatal for the contract line
            group = {"error": repr(e)[:300]}
        if world == 1:
            try:
                configs2 = configs2_record(lib, with_parity=not args.no_cpu_baseline)
            except Exception as e:
                configs2 = {"error": repr(e)[:300]}
    if world > 1:
        if store is not None:
            if rank == 0:
                store.set("rt_group_done", "1")
            else:
                import datetime
                store.wait(["rt_group_done"], datetime.timedelta(seconds=600))
        else:
            barrier()
    if rank != 0:
        return
    peak, peak_src = peaks()
    # ---- roofline.  The survey's per-ray-bytes figure (algorithmic bytes of the REFERENCE's access pattern over
    # the step time, against measured HBM bandwidth) is kept under `hbm`, but it is not a roofline fraction here:
    # the scene is L1/L2 resident and a packet shares every fetch among its rays, so it exceeds 1 by construction.
    # What bounds the dominant kernel is instruction issue: `achieved` = thread-instructions per second of the
    # primary stage (instruction count of the committed ncu capture of THESE sources x the kernel's LIVE duration
    # in this run), `peak` = SMs x 4 schedulers x 32 lanes x the SM clock sampled during this run.
    hbm_achieved = algo_bytes / world / (ms_per_step * 1e-3) / 1e9 if world == 1 else None
    roof = {"bound": "issue", "achieved": None, "peak": None, "unit": "Tthread-inst/s", "frac": None, "traffic": None,
            "kernel": "rt_primary_kernel", "kernel_ms_live": stage_ms["primary"] if stage_ms else None,
            "stage_ms_live": stage_ms, "source_fingerprint": source_fingerprint()}
    prof = committed_profile("primary")
    if prof and not prof["stale"] and stage_ms and stage_ms["primary"] > 0:
        pk = prof["data"]
        thread_inst = pk["smsp__inst_executed.sum"] * pk["smsp__thread_inst_executed_per_inst_executed.ratio"]
        sm_hz = (clocks or {}).get("sm_mhz") or 0.0
        roof["achieved"] = thread_inst / (stage_ms["primary"] * 1e-3) / 1e12
        roof["peak"] = 148 * 4 * 32 * sm_hz * 1e6 / 1e12 if sm_hz else None
        roof["frac"] = roof["achieved"] / roof["peak"] if roof["peak"] else None
        roof["traffic"] = pk["dram__bytes_read.sum"] + pk["dram__bytes_write.sum"]
        roof["ncu"] = {"file": prof["file"], "issue_slot_utilisation_pct": pk["smsp__issue_active.avg.pct_of_peak_sustained_active"],
                       "active_lanes_per_instruction": pk["smsp__thread_inst_executed_per_inst_executed.ratio"],
                       "issue_x_lanes": pk["smsp__issue_active.avg.pct_of_peak_sustained_active"] / 100 * pk["smsp__thread_inst_executed_per_inst_executed.ratio"] / 32,
                       "registers_per_thread": pk["launch__registers_per_thread"], "kernel_us_under_ncu": pk["gpu__time_duration.sum"],
                       "warp_instructions_per_launch": pk["smsp__inst_executed.sum"]}
    else:
        roof["stale_profile"] = (prof or {}).get("file", "no committed capture")
        roof["note_stale"] = "no ncu capture of these exact sources is committed: issue figures withheld"
    roof["hbm"] = {"achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                   "algorithmic_over_peak": (hbm_achieved / peak) if hbm_achieved else None,
                   "dram_traffic_over_peak": (roof["traffic"] / (stage_ms["primary"] * 1e-3) / 1e9 / peak) if roof["traffic"] and stage_ms else None,
                   "algorithmic_bytes_per_segment": bytes_per_segment,
                   "per_segment": {"nodes": nodes / segments, "tests": tests / segments, "shades": shades / segments},
                   "note": "algorithmic bytes follow the REFERENCE's access pattern (64 B/node + 16 B/test + 16 B/shade + 12 B/pixel) "
                           "over the whole step; > 1 x peak because the scene is cache resident and a packet shares each fetch: "
                           "not a roofline fraction, kept because SURVEY.md 8d defines it"}
    out = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "flushed (256 MiB memset) before every timed step, outside its event pair",
                   "timing": "per-step CUDA events on the launching stream, summed over K steps, max over ranks",
                   "camera": "moves every step (yaw + 0.001 degree): scan tables, origin-relative records and all kernels run "
                             "in every timed step; no CUDA-graph replay",
                   "parallelism": f"interleaved 16x16 tiles over {world} GPU(s), scene replicated"
                                  + ("" if world == 1 else ", tiles stored straight into rank 0's frame over NVLink peer memory + flag barrier"
                                     if peer is not None else ", NCCL all-gather of tile buffers + de-interleave"),
                   "segments_per_step": segments, "primary_paths_per_s": paths / (ms_per_step * 1e-3),
                   "frame_ms_kernel": ms_per_step, "wall_ms_per_step_incl_flush": t_wall / args.steps * 1e3,
                   "replay_ms_per_step_camera_standing_still": replay_ms,
                   "scene_broadcast_bytes": scene_bcast_bytes,
                   "precision": "float32 search + float64 confirmation/shading of the found hit",
                   "path": args.path},
        "clocks": clocks, "e2e": e2e, "present": present, "gpu_launches": launches, "roofline": roof,
        "single_process_group": group, "configs2": configs2,
    }
    if world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        mrays, ms, tot, orgb, oids = time_oracle(bundle, flat, 1, 1, 0, keep_frame=True)
        out["cpu_baseline"] = {"value": mrays, "unit": "Mrays/s", "cores": 1, "kind": "port", "frame_ms": ms,
                               "sample": f"one full {WIDTH}x{HEIGHT} frame ({tot['segments']} segments), 1 thread "
                                         "(the reference renderer is single-threaded)"}
        # the same counters from the oracle (float64 walk): equal up to rare float32 cell-boundary ties
        out["roofline"]["hbm"]["oracle_counters_rel_diff"] = {
            k: (v - tot[k]) / max(tot[k], 1) for k, v in (("segments", segments), ("nodes", nodes), ("tests", tests),
                                                          ("shades", shades))}
        # ---- parity of the very frame that was benchmarked (step 0's pose) against the oracle frame just rendered:
        # first-hit ids and float32 pixels of all 1920x1080 pixels (non-square: the intent mapping on both sides)
        from util import compare, insertion_ids
        mine = insertion_ids(flat, bundle, gpu_ids.reshape(HEIGHT, WIDTH))
        par = compare(host_rgb.reshape(HEIGHT, WIDTH, 3), mine, orgb, oids)
        par["pixels"] = npx
        par["hit_pixels"] = int((oids >= 0).sum())
        par["gate"] = "ids equal on >= 99.99 % of pixels, RGB within 1/255 on id-equal pixels"
        par["pass"] = bool(par["id_match"] >= 0.9999 and par["rgb_bad"] == 0)
        out["parity"] = par
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-group", action="store_true", help="skip the extra records (single_process_group, configs2)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N > 1: peer = kernels store into rank 0's frame over NVLink (CUDA IPC); nccl = all-gather + untile")
    ap.add_argument("--path", default="pipeline", choices=["pipeline", "per-ray"],
                    help="pipeline (default): packet primary stage + bounce stage; per-ray: every ray walked alone")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
