"""The committed fixtures under tests/golden/:
 * reference_jest_vectors.json - the reference's own known answers (transcribed with their file:line), replayed
   against the oracle (the parametrised tests of test_oracle_golden.py spell the same vectors out in code;
   this file checks that the transcription and the code agree, so neither can drift);
 * oracle_*.npz - frames rendered by the oracle (make_oracle_fixtures.py): the oracle must still produce them
   bit for bit, the test-only host build of the kernel body and - on a GPU - the CUDA path through the C ABI
   must match them under the parity gate."""
import json
import math
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_oracle_fixtures as mk
import raytracer_js_b200 as rt
from raytracer_js_b200 import _native as N
from raytracer_js_b200 import scenes
from util import compare, flat_of, hostsim_render, insertion_ids, make_params

EPS = 2.220446049250313e-16
VEC = json.load(open(os.path.join(HERE, "golden", "reference_jest_vectors.json")))


def _num(v):
    if isinstance(v, str):
        return eval(v, {"sqrt": math.sqrt, "eps": EPS, "pi": math.pi})
    return v


def test_jest_walker_vectors(oracle):
    for case in VEC["walker_one_level"]["cases"]:
        s = oracle.Scene((0, 0, 0), 1.0)
        stops = s.walk([_num(x) for x in case["pos"]], [_num(x) for x in case["dir"]], include_undefined=True)
        got = [int(st[1]) for st in stops[:-1]]
        want = case["head_behaviour"] if "head_behaviour" in case else case["octants"]
        assert got == want, case
        assert tuple(stops[-1]) == (-1, -1, 0)
    two = VEC["walker_two_level"]
    s = oracle.Scene((0, 0, 0), 1.0)
    for o in two["subtrees_at_root_octants"]:
        assert s.new_subtree([], o) == 0
    f = s.flat()
    names = {"tree": 0, "s1": int(f.node_child[0, 0]), "s2": int(f.node_child[0, 3]), "s3": int(f.node_child[0, 7])}
    stops = s.walk(two["pos"], two["dir"], include_undefined=True)
    assert [(int(a), int(b)) for a, b, _ in stops[:-1]] == [(names[n], o) for n, o in two["stops"]]


def test_jest_placement_and_location_vectors(oracle):
    nap = VEC["node_at_pos"]
    s = oracle.Scene((0, 0, 0), 1.0)
    path = []
    for o in nap["subtree_path"]:
        assert s.new_subtree(path, o) == 0
        path.append(o)
    f = s.flat()
    inner = int(f.node_child[0, nap["subtree_path"][0]])
    assert s.node_at_pos(nap["pos"]) == (int(f.node_child[inner, nap["subtree_path"][1]]), nap["expected"][1])
    pl = VEC["entity_placement"]
    s = oracle.Scene((0, 0, 0), 1.0)
    m, t, sub = s.add_material(0, False, True, 0.0), s.add_texture_solid(1, 1, 1, 1), s.add_substance(1.0)
    e = [s.add_entity(0, x["pos"], x["diameter"], m, t, sub, max_in_depth=pl["max_in_depth"], max_out_depth=pl["max_out_depth"])
         for x in pl["entities"]]
    f = s.flat()
    assert s.entity_node(e[0]) == int(f.node_child[0, 0]) and s.entity_node(e[1]) == 0
    s = oracle.Scene((0, 0, 0), 1.0)
    for i in VEC["octree_get_bounds"]["throws_for_indices"]:
        assert s.new_subtree([], i) == -2


@pytest.mark.parametrize("name", sorted(mk.FIXTURES))
def test_oracle_reproduces_its_fixtures(oracle, name):
    fx = np.load(os.path.join(HERE, "golden", f"oracle_{name}.npz"))
    _, rgb, ids, tot = mk.render(name)
    assert np.array_equal(ids, fx["ids"]) and np.array_equal(rgb, fx["rgb"])
    assert [tot["segments"], tot["nodes"], tot["tests"], tot["shades"]] == fx["totals"].tolist()
    if "screen" in fx.files:
        st = oracle.exposure_stats(rgb)
        assert np.array_equal(np.array(st), fx["stats"])
        assert np.array_equal(oracle.discretize(rgb, *fx["drange"]), fx["screen"])


@pytest.mark.parametrize("name", sorted(mk.FIXTURES))
def test_kernel_body_on_the_host_matches_fixtures(name):
    """tests/hostsim (rt_trace.cuh compiled for the host) against the committed frames: no oracle in the loop."""
    kw, size, frames = mk.FIXTURES[name]
    fx = np.load(os.path.join(HERE, "golden", f"oracle_{name}.npz"))
    b = mk.bundle_of(kw)
    flat = flat_of(b)
    cam = scenes.bench_camera(size, size)
    for pipeline in (False, True):
        rgb, ids, cnt = hostsim_render(flat, cam, make_params(flat, b, n_frames=frames), pipeline=pipeline)
        res = compare(rgb, insertion_ids(flat, b, ids), fx["rgb"], fx["ids"])
        assert res["id_match"] >= 0.9999 and res["rgb_bad"] == 0, (pipeline, res)
        if not pipeline:  # the counting variant reproduces the reference's access pattern
            got = np.array([cnt["segments"], cnt["nodes"], cnt["tests"], cnt["shades"]], float)
            assert np.allclose(got, fx["totals"], rtol=2e-3), (got, fx["totals"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mk.FIXTURES))
def test_gpu_matches_fixtures(name):
    """The CUDA path through the C ABI against the committed frames (and the committed screen)."""
    import ctypes as C
    kw, size, frames = mk.FIXTURES[name]
    fx = np.load(os.path.join(HERE, "golden", f"oracle_{name}.npz"))
    b = mk.bundle_of(kw)
    eb = rt.ExposureBuffer(size, size)
    tracer = rt.GpuRaytracer(rt.RaytracerConfig(b.refmax, b.sky, b.default_substance, 1.0), b.tree, scenes.bench_camera(size, size),
                             eb, rt.FpLcg(1.0))
    tracer.trace_frame(n_frames=frames, want_ids=True)
    assert tracer.lib.rt_launch_count(tracer.ctx) >= 3
    res = compare(eb.image(), insertion_ids(tracer.flat, b, tracer.last_first_ids), fx["rgb"], fx["ids"])
    assert res["id_match"] >= 0.9999 and res["rgb_bad"] == 0, res
    if "screen" in fx.files:
        # present the FIXTURE's float frame: the screen must be the committed one up to one level where the
        # parallel sums move the range in its last bits
        screen = np.zeros((size * size, 4), np.uint8)
        tone, st = N.Tone(N.RT_TONE_STDDEV, 8, 1 / 256, 8.0), N.ExposureStats()
        px = np.ascontiguousarray(fx["rgb"]).reshape(-1)
        N.check(tracer.ctx, tracer.lib.rt_present(tracer.ctx, px.ctypes.data, size, size, C.byref(tone), screen.ctypes.data, C.byref(st)))
        assert np.allclose([st.mean, st.variance, st.absolute_dev], fx["stats"], rtol=1e-11)
        assert np.allclose([st.drange_low, st.drange_high], fx["drange"], rtol=1e-11)
        diff = np.abs(screen.astype(int) - fx["screen"].astype(int))
        assert diff.max() <= 1 and (diff > 0).mean() < 1e-3
