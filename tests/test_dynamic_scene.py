"""Dynamic scenes: moving entities the reference's way - BasicEntity._set_pos
(src/entities/entity_basic.ts:38-42) + add_entity_to_octree again (src/octree_entity.ts:174-188), whose
Entity.set_octree (src/entity.ts:50-56) deletes the entity from its node's Set and adds it to the END of the new node's
Set.  Three implementations must agree: the oracle's (orc_move_entity on its pointer tree), the host mirror's
(Python octree, re-flattened from scratch), and the library's incremental one (rt_scene_update /
rt_host.h: rt_scene_move_entities on its copy of the flat scene: only the moved entities cross the boundary)."""
import ctypes as C

import numpy as np
import pytest

import raytracer_js_b200 as rt
from raytracer_js_b200 import _native as N
from raytracer_js_b200 import scenes
from raytracer_js_b200.octree_entity import add_entity_to_octree

from util import compare, flat_of, hostsim, hostsim_render, insertion_ids, make_params, oracle_render, oracle_scene
from test_hostsim_parity import cameras


def moves_for(b, k=40, seed=3.0):
    """k moves: some entities twice, some into the cell of another entity, one far away, one by a hair."""
    rng = rt.FpLcg(seed)
    n = len(b.entities)
    out = []
    for m in range(k):
        i = int(rng.next() * n) if m % 7 else out[-1][0] if out else 0  # every 7th move: the same entity again
        e = b.entities[i]
        d = e.get_diameter() if isinstance(e, rt.SphereEntity) else e.get_size()
        if m % 5 == 4:
            p = [c + 1e-9 for c in e.get_pos().v]  # by a hair: same node, but it goes to the end of its list
        else:
            p = [d / 2 + rng.next() * (1 - d) for _ in range(3)]
        out.append((i, p))
    return out


def scene():
    return scenes.random_spheres(900, 0.02, 0.09, seed=6.0, mix="mirrors", box_fraction=0.2)


def test_incremental_moves_equal_the_oracle_and_a_fresh_flatten(oracle):
    b = scene()
    flat0 = flat_of(b)  # entity ids of the flattener, before any move
    order0 = {id(e): i for i, e in enumerate(flat0.entities)}
    os_ = oracle_scene(flat0, b)  # the oracle's ids are insertion indices
    mv = moves_for(b)
    W = H = 96
    cam, ocam = cameras(W, H)
    prm = make_params(flat0, b, n_frames=2)
    # (1) the library's incremental path, on the host build of the kernel body
    ids = np.array([order0[id(b.entities[i])] for i, _ in mv], np.uint32)
    pos = np.array([p for _, p in mv], np.float64)
    L = hostsim()
    L.hostsim_set_moves.restype = None
    L.hostsim_set_moves(len(ids), ids.ctypes.data_as(N._up), pos.ctypes.data_as(N._dp), 16)
    try:
        rgb1, ids1, _ = hostsim_render(flat0, cam, prm, pipeline=True)
    finally:
        L.hostsim_set_moves(0, None, None, 16)
    ids1 = insertion_ids(flat0, b, ids1)
    # (2) the oracle: the same moves on its own tree
    for i, p in mv:
        os_.move_entity(i, p)
    orgb, oids, _, tot = oracle_render(os_, ocam, flat0, b, prm, fixed_extents=True)
    res = compare(rgb1, ids1, orgb, oids)
    assert res["id_mismatch"] == 0 and res["rgb_bad"] == 0 and res["rgb_max_abs"] == 0.0, res
    # (3) the host mirror moved the reference's way and flattened from scratch
    for i, p in mv:
        e = b.entities[i]
        e._set_pos(rt.point(*p))
        add_entity_to_octree(b.tree, e, {"max_in_depth": 16, "max_out_depth": 0})
    flat2 = flat_of(b)
    rgb2, ids2, _ = hostsim_render(flat2, cam, make_params(flat2, b, n_frames=2), pipeline=True)
    np.testing.assert_array_equal(rgb2, rgb1)
    np.testing.assert_array_equal(insertion_ids(flat2, b, ids2), ids1)
    # the moves changed the picture
    rgb0, _, _ = hostsim_render(flat0, cam, prm, pipeline=True)
    assert not np.array_equal(rgb0, rgb1)
    # per-node insertion order after the moves: the oracle's tree and the host mirror's agree list by list
    f = os_.flat()
    lists_oracle = sorted(tuple(f.list_entity[f.node_list_off[i]:f.node_list_off[i + 1]].tolist()) for i in range(len(f.node_size))
                          if f.node_list_off[i + 1] > f.node_list_off[i])
    a = flat2.arrays
    ins = {id(e): i for i, e in enumerate(b.entities)}
    lut = [ins[id(e)] for e in flat2.entities]
    lists_mirror = sorted(tuple(lut[j] for j in a["list_entity"][a["node_list_off"][i]:a["node_list_off"][i + 1]])
                          for i in range(len(a["node_size"])) if a["node_list_off"][i + 1] > a["node_list_off"][i])
    assert lists_oracle == lists_mirror


def test_move_out_of_the_root_is_refused(oracle):
    b = scenes.random_spheres(50, 0.05, 0.1, seed=2.0)
    flat = flat_of(b)
    cam, _ = cameras(32, 32)
    ids = np.array([3], np.uint32)
    pos = np.array([[0.99, 0.5, 0.5]], np.float64)  # AABB sticks out of the unit root: TreeOutsideGrowError
    L = hostsim()
    L.hostsim_set_moves.restype = None
    L.hostsim_set_moves(1, ids.ctypes.data_as(N._up), pos.ctypes.data_as(N._dp), 16)
    try:
        with pytest.raises(RuntimeError, match="outside-depth"):
            hostsim_render(flat, cam, make_params(flat, b), pipeline=True)
    finally:
        L.hostsim_set_moves(0, None, None, 16)


@pytest.mark.gpu
def test_gpu_scene_update(oracle):
    """rt_scene_update through the C ABI (GpuRaytracer.move_entities): same frame as the oracle with the same moves and
    as a full re-flatten + re-upload (refresh_scene); also on a group (several members, one per GPU when there are)."""
    for kw in ({}, {"devices": [0, 0]}):
        b = scene()
        W = H = 128
        cam = scenes.bench_camera(W, H)
        eb = rt.ExposureBuffer(W, H)
        tracer = rt.GpuRaytracer(rt.RaytracerConfig(b.refmax, b.sky, b.default_substance, 1.0), b.tree, cam, eb, rt.FpLcg(1.0), **kw)
        os_ = oracle_scene(tracer.flat, b)
        tracer.trace_frame(n_frames=2, want_ids=True)
        before = eb.pixels.copy()
        mv = moves_for(b)
        launches = tracer.lib.rt_launch_count(tracer.ctx)
        tracer.move_entities([b.entities[i] for i, _ in mv], [p for _, p in mv])
        eb.reset_exposure()
        tracer.trace_frame(n_frames=2, want_ids=True)
        assert tracer.lib.rt_launch_count(tracer.ctx) > launches
        got_rgb, got_ids = eb.image().copy(), insertion_ids(tracer.flat, b, tracer.last_first_ids)
        for i, p in mv:
            os_.move_entity(i, p)
        import math
        import oracle as orc
        ocam = orc.Camera(math.pi / 2, math.pi / 2, W, H, scenes.BENCH_CAMERA_POS, 0.0, math.pi / 6, vertical_locked=True)
        orgb, oids, _, _ = oracle_render(os_, ocam, tracer.flat, b, make_params(tracer.flat, b, n_frames=2), fixed_extents=True)
        res = compare(got_rgb, got_ids, orgb, oids)
        assert res["id_mismatch"] == 0 and res["rgb_bad"] == 0, res
        assert not np.array_equal(before, eb.pixels)
        # a full re-flatten of the (already moved) host tree gives the same frame
        flat_before = tracer.flat
        tracer.refresh_scene()
        eb2 = rt.ExposureBuffer(W, H)
        tracer.set_ebuffer(eb2)
        tracer.trace_frame(n_frames=2, want_ids=True)
        np.testing.assert_array_equal(eb2.pixels, eb.pixels)
        np.testing.assert_array_equal(insertion_ids(tracer.flat, b, tracer.last_first_ids), got_ids)
        # errors: out of the root -> RT_ERR_UNSUPPORTED, scene unchanged
        with pytest.raises(N.RtError, match="outside-depth"):
            tracer.move_entities([b.entities[0]], [(0.999, 0.5, 0.5)])
        eb3 = rt.ExposureBuffer(W, H)
        tracer.set_ebuffer(eb3)
        tracer.trace_frame(n_frames=2)
        np.testing.assert_array_equal(eb3.pixels, eb2.pixels)
        tracer.close()
