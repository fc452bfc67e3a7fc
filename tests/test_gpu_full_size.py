"""Parity at the sizes that are PUBLISHED (BASELINE.json configs[2]-[4]): the real entity counts (100 k / 1 M),
the real material mixes and textures, frames of the real height and the real samples per pixel, the CUDA path
through the C ABI (rt_render, host buffers) against the oracle on crops of the frame (`oracle.render(crop=)`:
same camera, same per-pixel seeds - they depend on the pixel's place in the FULL frame).  Parity is asserted on
square frames (the reference throws on non-square ones) plus one crop of the literal non-square
frame with the intent mapping on both sides.  Gate: ids equal on >= 99.99 % of the pixels, no colour outside
1/255 on id-equal pixels, and every outlier classified (tests/util.py: classify_outliers) - none unexplained."""
import math

import numpy as np
import pytest

import oracle as orc
from raytracer_js_b200 import scenes

from util import classify_outliers, compare, gpu_render_flat, oracle_crop, oracle_scene_flat

pytestmark = pytest.mark.gpu


def check_crops(fb, oscene, width, height, n_frames, crops, image_textures=False):
    rgb, ids, launches = gpu_render_flat(fb, width, height, n_frames)
    assert launches >= 3  # prepare + primary stage + bounce stage at least: the pipeline ran, not a fallback
    assert np.isfinite(rgb).all() and ids.min() >= -1 and ids.max() < fb.n_entities
    hit_fraction = []
    for crop in crops:
        x, y, w, h = crop
        orgb, oids, tot = oracle_crop(oscene, fb, width, height, crop, n_frames)
        a_rgb, a_ids = rgb[y:y + h, x:x + w], ids[y:y + h, x:x + w]
        res = compare(a_rgb, a_ids, orgb, oids)
        ocam = orc.Camera(math.pi / 2, math.pi / 2, width, height, scenes.BENCH_CAMERA_POS, 0.0, math.pi / 6,
                          vertical_locked=True) if res["id_mismatch"] and width * height <= 2160 * 2160 else None
        kinds = classify_outliers(a_rgb, a_ids, orgb, oids, cam_pos=scenes.BENCH_CAMERA_POS, ocam=ocam,
                                  image_textures=image_textures, offset=(x, y))
        assert res["id_match"] >= 0.9999 and not kinds["unexplained"], (crop, res, kinds)
        assert res["rgb_bad"] == len(kinds["texel_edge"]), (crop, res, kinds)
        assert tot["paths"] == w * h * n_frames
        hit_fraction.append(float((oids >= 0).mean()))
    return hit_fraction


def test_config2_100k_spheres_16spp(oracle):
    """configs[2]: 100 k spheres d in [0.002,0.006], 70 % mirrors / 15 % diffuse / 10 % rough / 5 % lights, refmax 4,
    16 exposure frames (the resample stage runs): a 1080 x 1080 frame, 256 x 256 crops at the centre and at a corner;
    then the literal 1920 x 1080 frame (intent mapping on both sides), one 256 x 128 crop."""
    cfg = scenes.BASELINE_CONFIGS["c2"]
    fb = scenes.build_config(cfg)
    oscene = oracle_scene_flat(fb)
    hits = check_crops(fb, oscene, 1080, 1080, cfg["spp"], [(412, 412, 256, 256), (0, 824, 256, 256)])
    assert max(hits) > 0.5  # the crops really look at the spheres
    check_crops(fb, oscene, cfg["w"], cfg["h"], cfg["spp"], [(1000, 500, 256, 128)])


def test_config3_1m_entities_textured_4spp(oracle):
    """configs[3] (the reference-pinned variant): 1 M entities, 10 % boxes, spheres textured from 4
    image textures of 1024 x 512, 4 exposure frames: a 2160 x 2160 frame, 256 x 256 crop at the centre and a
    256 x 128 crop of the upper edge."""
    cfg = scenes.BASELINE_CONFIGS["c3"]
    fb = scenes.build_config(cfg)
    oscene = oracle_scene_flat(fb)
    hits = check_crops(fb, oscene, 2160, 2160, cfg["spp"], [(952, 952, 256, 256), (1700, 0, 256, 128)],
                       image_textures=True)
    assert max(hits) > 0.5


def test_config4_1m_spheres_8k_height(oracle):
    """configs[4]: 1 M spheres d in [0.0005,0.002], config-2 material mix, on a frame of the 8K frame's height
    (4320 x 4320), 8 exposure frames (resample stage), 160 x 160 crops at the centre and off-centre."""
    cfg = scenes.BASELINE_CONFIGS["c4"]
    fb = scenes.build_config(cfg)
    oscene = oracle_scene_flat(fb)
    hits = check_crops(fb, oscene, 4320, 4320, 8, [(2080, 2080, 160, 160), (300, 3900, 160, 160)])
    assert max(hits) > 0.5
