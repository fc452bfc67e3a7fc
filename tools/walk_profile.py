"""Lane occupancy of the bounce stage's lock-step walk (a tools/ build of the library with -DRT_WALK_PROFILE, see
tools/walk_profile.sh): renders configs[2] once at 1 spp; the counters are printed by rt_destroy."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from raytracer_js_b200 import scenes  # noqa: E402
from util import gpu_render_flat  # noqa: E402

cfg = scenes.BASELINE_CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
fb = scenes.build_config(cfg)
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rgb, ids, launches = gpu_render_flat(fb, cfg["w"], cfg["h"], spp)
print("launches", launches, "hit fraction", float((ids >= 0).mean()))
print("fields: iterations walking node_steps node_waiting pair_steps leaf_steps iters_with_leaf iters_with_node iters_with_pair "
      "passes begin idle end_top end")
