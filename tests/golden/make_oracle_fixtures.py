"""Generates tests/golden/oracle_*.npz: small frames rendered by the CPU oracle (oracle/oracle.cpp), so that
the GPU tests can be checked against committed outputs as well as against the oracle built on the spot, and
so that a change of the oracle itself shows up as a diff.  These are outputs of the RESTATEMENT (the reference
cannot be executed here: no JavaScript runtime), not of the reference.

  python tests/golden/make_oracle_fixtures.py        (re-creates the files; they are deterministic)

Each file holds: rgb float32 [H,W,3], ids int32 [H,W] (insertion index of the first-hit entity, -1 none),
totals (segments, nodes, tests, shades of the reference's access pattern), and - for the present fixture - the
exposure statistics, the dynamic range and the RGBA8 screen of View.draw_ebuffer()."""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as orc
from raytracer_js_b200 import scenes
from util import flat_of, make_params, oracle_render, oracle_scene

# name -> (scene kwargs, frame size, exposure frames)
FIXTURES = {
    "diffuse": (dict(n=3000, dmin=0.004, dmax=0.02, seed=42.0, mix="diffuse"), 96, 1),
    "mirrors": (dict(n=1500, dmin=0.01, dmax=0.06, seed=11.0, mix="mirrors", box_fraction=0.2), 96, 3),
}


def bundle_of(kw):
    kw = dict(kw)
    return scenes.random_spheres(kw.pop("n"), kw.pop("dmin"), kw.pop("dmax"), **kw)


def render(name):
    kw, size, frames = FIXTURES[name]
    b = bundle_of(kw)
    flat = flat_of(b)
    ocam = orc.Camera(math.pi / 2, math.pi / 2, size, size, scenes.BENCH_CAMERA_POS, 0.0, math.pi / 6, vertical_locked=True)
    prm = make_params(flat, b, n_frames=frames)
    rgb, ids, _, tot = oracle_render(oracle_scene(flat, b), ocam, flat, b, prm, want_counters=True)
    return b, rgb, ids, tot


def main():
    orc.build()
    for name in FIXTURES:
        _, rgb, ids, tot = render(name)
        out = dict(rgb=rgb.astype(np.float32), ids=ids.astype(np.int32),
                   totals=np.array([tot["segments"], tot["nodes"], tot["tests"], tot["shades"]], np.int64))
        if name == "mirrors":
            st = orc.exposure_stats(rgb)
            lo, hi = orc.dynamic_range(orc.TONE_STDDEV, 8, 1 / 256, 8.0, st)
            out.update(stats=np.array(st), drange=np.array([lo, hi]), screen=orc.discretize(rgb, lo, hi))
        np.savez_compressed(os.path.join(HERE, f"oracle_{name}.npz"), **out)
        print(name, rgb.shape, {k: int(v) for k, v in tot.items() if k in ("segments", "nodes", "tests", "shades")})


if __name__ == "__main__":
    main()
