"""Differential fuzz (tests/fuzz_scenes.py): random scenes, poses, frame shapes, exposure counts and refmax - the kernel
body compiled for the host (pipeline and ray by ray) and, with a device, the CUDA path through the C ABI, against the
oracle.  No pixel may differ in colour; every id mismatch must be a classified dyadic tie."""
import numpy as np
import pytest

import raytracer_js_b200 as rt
from raytracer_js_b200 import _native as N
from raytracer_js_b200 import scenes  # noqa: F401

import fuzz_scenes
from util import classify_outliers, compare, flat_of, hostsim_render, insertion_ids, make_params, oracle_render, oracle_scene


def check(c, rgb, ids, flat, prm):
    _, ocam = fuzz_scenes.cameras(c)
    orgb, oids, _, tot = oracle_render(oracle_scene(flat, c["bundle"]), ocam, flat, c["bundle"], prm, fixed_extents=True)
    ids = insertion_ids(flat, c["bundle"], ids)
    res = compare(rgb, ids, orgb, oids)
    kinds = classify_outliers(rgb, ids, orgb, oids, cam_pos=c["pos"], ocam=ocam, image_textures=c["images"])
    assert res["rgb_bad"] == len(kinds["texel_edge"]) and not kinds["unexplained"], (fuzz_scenes.describe(c), res, {k: len(v) for k, v in kinds.items()})
    return float((oids >= 0).mean())


def test_host_build_of_the_kernel_body_against_the_oracle(oracle):
    hits = []
    for c in fuzz_scenes.cases(seed=20261019, count=60, max_entities=1500):
        flat = flat_of(c["bundle"])
        cam, _ = fuzz_scenes.cameras(c)
        prm = make_params(flat, c["bundle"], n_frames=c["n_frames"], refmax=c["refmax"])
        rgb_p, ids_p, _ = hostsim_render(flat, cam, prm, pipeline=True)
        rgb, ids, _ = hostsim_render(flat, cam, prm)
        np.testing.assert_array_equal(ids_p, ids, err_msg=str(fuzz_scenes.describe(c)))
        np.testing.assert_array_equal(rgb_p, rgb, err_msg=str(fuzz_scenes.describe(c)))
        hits.append(check(c, rgb, ids, flat, prm))
    assert sum(h > 0.05 for h in hits) >= 25  # the cases do look at geometry


def test_lattice_scenes_on_the_host_build(oracle):
    """Exact ties everywhere (fuzz_scenes.lattice_cases): a hit whose ray only touches the entity's cell is searched again
    by the float64 walker, so the frame is the oracle's - no id mismatch at all, not even the dyadic ties the float32
    walker used to leave - on the pipeline and ray by ray."""
    for c in fuzz_scenes.lattice_cases(seed=4, count=80):
        flat = flat_of(c["bundle"])
        cam, ocam = fuzz_scenes.cameras(c)
        prm = make_params(flat, c["bundle"], n_frames=c["n_frames"], refmax=c["refmax"])
        # (no RT_PARAM_EXACT_TIES: rt_pack_scene sees entity faces on cell planes and switches the tie checks on itself)
        rgb_p, ids_p, _ = hostsim_render(flat, cam, prm, pipeline=True)
        rgb, ids, _ = hostsim_render(flat, cam, prm)
        np.testing.assert_array_equal(ids_p, ids, err_msg=str(fuzz_scenes.describe(c)))
        np.testing.assert_array_equal(rgb_p, rgb, err_msg=str(fuzz_scenes.describe(c)))
        orgb, oids, _, tot = oracle_render(oracle_scene(flat, c["bundle"]), ocam, flat, c["bundle"], prm, fixed_extents=True)
        res = compare(rgb, insertion_ids(flat, c["bundle"], ids), orgb, oids)
        assert res["id_mismatch"] == 0 and res["rgb_bad"] == 0, (fuzz_scenes.describe(c), res)


def test_edge_cameras_and_sizes_on_the_host_build(oracle):
    """fuzz_scenes.edge_cases: no id and no colour may differ from the oracle's - not even a classified tie."""
    for c in fuzz_scenes.edge_cases(seed=13, count=80):
        flat = flat_of(c["bundle"])
        cam, ocam = fuzz_scenes.cameras(c)
        prm = make_params(flat, c["bundle"], n_frames=c["n_frames"], refmax=c["refmax"])
        ref = c["reference_extents"]
        rgb_p, ids_p, _ = hostsim_render(flat, cam, prm, pipeline=True, reference_extents=ref)
        rgb, ids, _ = hostsim_render(flat, cam, prm, reference_extents=ref)
        np.testing.assert_array_equal(ids_p, ids, err_msg=str(fuzz_scenes.describe(c)))
        np.testing.assert_array_equal(rgb_p, rgb, err_msg=str(fuzz_scenes.describe(c)))
        orgb, oids, _, tot = oracle_render(oracle_scene(flat, c["bundle"]), ocam, flat, c["bundle"], prm, fixed_extents=not ref)
        res = compare(rgb, insertion_ids(flat, c["bundle"], ids), orgb, oids)
        assert res["id_mismatch"] == 0 and res["rgb_bad"] == 0, (fuzz_scenes.describe(c), res)


def test_moved_entities_on_the_host_build(oracle):
    """rt_scene_update's path (rt_host.h: rt_scene_move_entities on the library's copy of the flat scene) under random
    moves - far away, by a hair, not at all, onto lattice points, the same entity twice - against the oracle's
    orc_move_entity on its own pointer tree: 1 to 200 moves per scene."""
    import random
    from util import hostsim
    R = random.Random(5)
    L = hostsim()
    L.hostsim_set_moves.restype = None
    for c in fuzz_scenes.cases(seed=31, count=30, max_entities=1500):
        b = c["bundle"]
        flat0 = flat_of(b)
        order0 = {id(e): i for i, e in enumerate(flat0.entities)}
        os_ = oracle_scene(flat0, b)
        mv = []
        for _ in range(R.choice([1, 5, 40, 200])):
            i = R.randrange(len(b.entities)) if not mv or R.random() > 0.15 else mv[-1][0]
            e = b.entities[i]
            d = e.get_diameter() if isinstance(e, rt.SphereEntity) else e.get_size()
            kind = R.random()
            if kind < 0.15:
                p = [x + R.choice([1e-9, -1e-9, 0.0]) for x in e.get_pos().v]
            elif kind < 0.35:
                g = 1.0 / R.choice([2, 4, 8, 16, 64])
                p = [round(R.random() / g) * g for _ in range(3)]
            else:
                p = [d / 2 + R.random() * (1 - d) for _ in range(3)]
            mv.append((i, [min(max(x, d / 2), 1 - d / 2) for x in p]))
        cam, ocam = fuzz_scenes.cameras(c)
        prm = make_params(flat0, b, n_frames=c["n_frames"], refmax=c["refmax"])
        ids = np.array([order0[id(b.entities[i])] for i, _ in mv], np.uint32)
        pos = np.array([p for _, p in mv], np.float64)
        L.hostsim_set_moves(len(ids), ids.ctypes.data_as(N._up), pos.ctypes.data_as(N._dp), 16)
        try:
            rgb, idm, _ = hostsim_render(flat0, cam, prm, pipeline=True)
        finally:
            L.hostsim_set_moves(0, None, None, 16)
        for i, p in mv:
            os_.move_entity(i, p)
        orgb, oids, _, tot = oracle_render(os_, ocam, flat0, b, prm, fixed_extents=True)
        res = compare(rgb, insertion_ids(flat0, b, idm), orgb, oids)
        assert res["id_mismatch"] == 0 and res["rgb_bad"] == 0, (fuzz_scenes.describe(c), len(mv), res)


@pytest.mark.gpu
def test_lattice_scenes_on_the_cuda_path(oracle):
    for c in fuzz_scenes.lattice_cases(seed=9, count=60):
        b = c["bundle"]
        cam, ocam = fuzz_scenes.cameras(c)
        out = []
        for want_counters in (False, True):  # the pipeline, and the counting variant (the float64 walker, ray by ray)
            eb = rt.ExposureBuffer(c["w"], c["h"])
            tracer = rt.GpuRaytracer(rt.RaytracerConfig(c["refmax"], b.sky, b.default_substance, 1.0), b.tree, cam, eb, rt.FpLcg(1.0),
                                     exact_ties=True)
            tracer.trace_frame(n_frames=c["n_frames"], want_ids=True, want_counters=want_counters)
            out.append((eb.image().copy(), tracer.last_first_ids.copy(), tracer.last_counters))
            flat = tracer.flat
            tracer.close()
        prm = make_params(flat, b, n_frames=c["n_frames"], refmax=c["refmax"])
        orgb, oids, _, tot = oracle_render(oracle_scene(flat, b), ocam, flat, b, prm, fixed_extents=True, want_counters=True)
        for rgb, ids, cnt in out:
            res = compare(rgb, insertion_ids(flat, b, ids), orgb, oids)
            assert res["id_mismatch"] == 0 and res["rgb_bad"] == 0, (fuzz_scenes.describe(c), res)
        for key in ("paths", "segments", "nodes", "tests", "shades"):
            assert out[1][2][key] == tot[key], (key, fuzz_scenes.describe(c), out[1][2], tot)


@pytest.mark.gpu
def test_cuda_path_against_the_oracle(oracle):
    hits = []
    for c in fuzz_scenes.cases(seed=977, count=50):
        b = c["bundle"]
        cam, _ = fuzz_scenes.cameras(c)
        eb = rt.ExposureBuffer(c["w"], c["h"])
        tracer = rt.GpuRaytracer(rt.RaytracerConfig(c["refmax"], b.sky, b.default_substance, 1.0), b.tree, cam, eb, rt.FpLcg(1.0))
        tracer.trace_frame(n_frames=c["n_frames"], want_ids=True)
        prm = make_params(tracer.flat, b, n_frames=c["n_frames"], refmax=c["refmax"])
        hits.append(check(c, eb.image().copy(), tracer.last_first_ids.copy(), tracer.flat, prm))
        tracer.close()
    assert sum(h > 0.05 for h in hits) >= 20


@pytest.mark.gpu
def test_group_frames_equal_single_gpu_frames():
    """rt_create_multi (three members on one GPU, or every GPU of the box) on the fuzz cases: the sharded frame - ids,
    pixels - is the single-GPU frame, bit for bit, random and lattice scenes alike."""
    import itertools
    import torch
    devs = list(range(torch.cuda.device_count())) if torch.cuda.device_count() > 1 else [0, 0, 0]
    for c in itertools.chain(fuzz_scenes.cases(seed=71, count=20, max_entities=1500), fuzz_scenes.lattice_cases(seed=72, count=20)):
        b = c["bundle"]
        cam, _ = fuzz_scenes.cameras(c)
        out = []
        for kw in ({}, {"devices": devs}):
            eb = rt.ExposureBuffer(c["w"], c["h"])
            t = rt.GpuRaytracer(rt.RaytracerConfig(c["refmax"], b.sky, b.default_substance, 1.0), b.tree, cam, eb, rt.FpLcg(1.0),
                                exact_ties=True, **kw)
            t.trace_frame(n_frames=c["n_frames"], want_ids=True)
            out.append((eb.pixels.copy(), t.last_first_ids.copy()))
            t.close()
        np.testing.assert_array_equal(out[1][0], out[0][0], err_msg=str(fuzz_scenes.describe(c)))
        np.testing.assert_array_equal(out[1][1], out[0][1], err_msg=str(fuzz_scenes.describe(c)))
