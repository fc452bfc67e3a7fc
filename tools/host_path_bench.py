"""Times rt_render (host buffers, banded copies) on a BASELINE config: the frame as the reference's call site gets it.
usage: python tools/host_path_bench.py c2 [--spp N] [--steps K]   (RT_B200_BANDS=n in the environment to compare)"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import raytracer_js_b200 as rt
from raytracer_js_b200 import _native as N
from raytracer_js_b200 import scenes
from util import flat_params

ap = argparse.ArgumentParser()
ap.add_argument("config")
ap.add_argument("--spp", type=int, default=0)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
cfg = dict(scenes.BASELINE_CONFIGS[a.config])
if a.spp:
    cfg["spp"] = a.spp
fb = scenes.build_config(cfg)
lib = N.load()
ctx = C.c_void_p()
N.check(None, lib.rt_create(0, C.byref(ctx)))
d = fb.flat.desc()
N.check(ctx, lib.rt_scene_upload(ctx, C.byref(d)))
W, H = cfg["w"], cfg["h"]
prm = flat_params(fb, cfg["spp"])
rgb = np.zeros(W * H * 3, np.float32)
N.check(ctx, lib.rt_host_register(ctx, rgb.ctypes.data, rgb.nbytes))
ms = []
for i in range(a.steps + 2):
    cd = rt.camera_desc(scenes.bench_camera(W, H, yaw_deg=30.0 + 0.01 * i))
    t0 = time.perf_counter()
    N.check(ctx, lib.rt_render(ctx, C.byref(cd), C.byref(prm), 0, rgb.ctypes.data, None, None))
    if i >= 2:
        ms.append((time.perf_counter() - t0) * 1e3)
ms.sort()
print(json.dumps({"config": a.config, "spp": cfg["spp"], "bands": os.environ.get("RT_B200_BANDS", "default"),
                  "resample_min": os.environ.get("RT_B200_RESAMPLE_MIN", "default"), "host_frame_ms": ms[len(ms) // 2]}))
lib.rt_host_unregister(ctx, rgb.ctypes.data)
lib.rt_destroy(ctx)
