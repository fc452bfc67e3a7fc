"""Parity gate between a frame rendered by the REAL reference (render_headless.ts, see README.md) and the oracle.
usage: python compare_with_oracle.py ref.bin --n 3000 --dmin 0.004 --dmax 0.02 --seed 42 --mix diffuse [--boxes 0.2]"""
import argparse
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # tests/ref_node -> repo root
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as orc
from raytracer_js_b200 import scenes
from util import compare, flat_of, make_params, oracle_render, oracle_scene


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("file")
    ap.add_argument("--n", type=int, required=True)
    ap.add_argument("--dmin", type=float, required=True)
    ap.add_argument("--dmax", type=float, required=True)
    ap.add_argument("--seed", type=float, default=42.0)
    ap.add_argument("--mix", default="diffuse")
    ap.add_argument("--boxes", type=float, default=0.0)
    a = ap.parse_args()
    raw = open(a.file, "rb").read()
    W, H, frames = np.frombuffer(raw[:12], np.int32)
    rgb = np.frombuffer(raw[12:12 + W * H * 12], np.float32).reshape(H, W, 3)
    ids = np.frombuffer(raw[12 + W * H * 12:], np.int32).reshape(H, W)
    b = scenes.random_spheres(a.n, a.dmin, a.dmax, seed=a.seed, mix=a.mix, box_fraction=a.boxes)
    flat = flat_of(b)
    orc.build()
    ocam = orc.Camera(math.pi / 2, math.pi / 2, int(W), int(H), scenes.BENCH_CAMERA_POS, 0.0, math.pi / 6, vertical_locked=True)
    orgb, oids, _, _ = oracle_render(oracle_scene(flat, b), ocam, flat, b, make_params(flat, b, n_frames=int(frames)))
    res = compare(orgb, oids, rgb, ids)
    print(res)
    ok = res["id_match"] >= 0.9999 and res["rgb_bad"] == 0
    print("PARITY OK: the oracle reproduces the reference" if ok else "PARITY FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
