"""Lane occupancy of the bounce stage's lock-step walk (a tools/ build of the library with -DRT_WALK_PROFILE, see
tools/walk_profile.sh): renders configs[2] once at 1 spp; the counters are printed by rt_destroy."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from raytracer_js_b200 import scenes  # noqa: E402
from util import gpu_render_flat  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
if name == "c1":  # the headline scene of bench.py (configs[1]): 10 k diffuse spheres, refmax 1
    cfg = dict(w=1920, h=1080)
    fb = scenes.random_spheres_flat(10000, 0.002, 0.006, seed=42.0, mix="diffuse")
else:
    cfg = scenes.BASELINE_CONFIGS[name]
    fb = scenes.build_config(cfg)
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 1
import ctypes as C  # noqa: E402

import torch  # noqa: E402

import raytracer_js_b200 as rt  # noqa: E402
from raytracer_js_b200 import _native as N  # noqa: E402
from util import flat_params  # noqa: E402

lib = N.load()
ctx = C.c_void_p()
N.check(None, lib.rt_create(0, C.byref(ctx)))
d = fb.flat.desc()
N.check(ctx, lib.rt_scene_upload(ctx, C.byref(d)))
W, H = cfg["w"], cfg["h"]
cd = rt.camera_desc(scenes.bench_camera(W, H))
prm = flat_params(fb, spp)
frame = torch.zeros(W * H * 3, dtype=torch.float32, device="cuda")
N.check(ctx, lib.rt_render_device(ctx, C.byref(cd), C.byref(prm), 0, C.c_void_p(frame.data_ptr()), None))  # one band, one launch per stage
N.check(ctx, lib.rt_synchronize(ctx))
print("launches", int(lib.rt_launch_count(ctx)))
lib.rt_destroy(ctx)
print("fields: iterations walking node_steps node_waiting pair_steps leaf_steps iters_with_leaf iters_with_node iters_with_pair "
      "passes begin idle end_top end")
