"""The differential fuzz of tests/test_fuzz_parity.py for as long as one likes, on the host build of the kernel body:
    python tools/fuzz_parity.py [seed=1] [cases=400]
Prints every mismatch with the generator's choices (enough to reproduce it), then a summary."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as orc  # noqa: E402

import fuzz_scenes  # noqa: E402
from util import classify_outliers, compare, flat_of, hostsim_render, insertion_ids, make_params, oracle_render, oracle_scene  # noqa: E402

orc.build()
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
count = int(sys.argv[2]) if len(sys.argv) > 2 else 400
bad, hits, t0 = 0, [], time.time()
for c in fuzz_scenes.cases(seed, count):
    flat = flat_of(c["bundle"])
    cam, ocam = fuzz_scenes.cameras(c)
    prm = make_params(flat, c["bundle"], n_frames=c["n_frames"], refmax=c["refmax"])
    rgb_p, ids_p, _ = hostsim_render(flat, cam, prm, pipeline=True)
    rgb, ids, _ = hostsim_render(flat, cam, prm)
    orgb, oids, _, tot = oracle_render(oracle_scene(flat, c["bundle"]), ocam, flat, c["bundle"], prm, fixed_extents=True)
    same = np.array_equal(rgb_p, rgb) and np.array_equal(ids_p, ids)
    ids_i = insertion_ids(flat, c["bundle"], ids)
    res = compare(rgb, ids_i, orgb, oids)
    kinds = classify_outliers(rgb, ids_i, orgb, oids, cam_pos=c["pos"], ocam=ocam, image_textures=c["images"])
    hits.append(float((oids >= 0).mean()))
    if c["case"] % 3 == 0:  # RT_PRECISION_F64: the float64 walker must give the oracle's frame AND its counters, exactly
        from raytracer_js_b200 import _native as N
        prm.precision = N.RT_PRECISION_F64
        rgb64, ids64, cnt64 = hostsim_render(flat, cam, prm)
        prm.precision = N.RT_PRECISION_F32
        _, _, _, totc = oracle_render(oracle_scene(flat, c["bundle"]), ocam, flat, c["bundle"], prm, fixed_extents=True, want_counters=True)
        res64 = compare(rgb64, insertion_ids(flat, c["bundle"], ids64), orgb, oids)
        if res64["id_mismatch"] or res64["rgb_bad"] or any(cnt64[k] != totc[k] for k in ("paths", "segments", "nodes", "tests", "shades")):
            bad += 1
            print("MISMATCH (float64 search)", fuzz_scenes.describe(c), res64, {k: (cnt64[k], totc[k]) for k in ("segments", "nodes", "tests", "shades")})
    if not (same and res["rgb_bad"] == len(kinds["texel_edge"]) and not kinds["unexplained"]):
        bad += 1
        print("MISMATCH", fuzz_scenes.describe(c), "pipeline == ray by ray:", same, res, {k: len(v) for k, v in kinds.items()})
print(f"cases {count}  mismatches {bad}  cases with hits {sum(h > 0.05 for h in hits)}  {time.time() - t0:.0f} s")
