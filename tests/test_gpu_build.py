"""rt_tree_build_gpu (device-side restatement of add_entity_to_octree, csrc/rt_build_gpu.cuh)
against the host builder rt_tree_build (itself checked against one-by-one insertion through the host API and
against the oracle): the same tree - node set, float64 positions and sizes, child tables, per-node entity
order - and the same rendered frame.  Also the reference's placement vectors (test/octree-entity.test.ts:52-64)
and the TreeOutsideGrowError path."""
import ctypes as C
import time

import numpy as np
import pytest

import raytracer_js_b200 as rt
from raytracer_js_b200 import _native as N
from raytracer_js_b200 import scenes
from raytracer_js_b200.flatten import flat_from_arrays

from test_host_build import canonical_tree

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    lib = N.load()
    c = C.c_void_p()
    N.check(None, lib.rt_create(0, C.byref(c)))
    yield c
    lib.rt_destroy(c)


def arrays(n, dmin, dmax, box_fraction, seed=7.0):
    fb = scenes.random_spheres_flat(n, dmin, dmax, seed=seed, mix="mirrors", box_fraction=box_fraction)
    return fb, fb.flat.arrays


def rebuild_on_gpu(ctx, fb, max_in_depth=16):
    a = fb.flat.arrays
    return flat_from_arrays(a["ent_type"], a["ent_pos"], a["ent_extent"], a["ent_material"], a["ent_texture"], a["ent_substance"],
                            fb.flat.materials, fb.flat.textures, fb.flat.substances, max_in_depth=max_in_depth, gpu_ctx=ctx)


def assert_same_tree(A, B):
    ta, tb = canonical_tree(A), canonical_tree(B)
    assert ta.keys() == tb.keys()  # same nodes: bit-identical float64 position and size
    for key in ta:
        ca, ba, ea = ta[key]
        cb, bb, eb = tb[key]
        assert ca == cb
        assert A["list_entity"][ba:ea].tolist() == B["list_entity"][bb:eb].tolist(), key  # insertion order kept


@pytest.mark.parametrize("n,dmin,dmax,boxes,depth", [(3000, 0.004, 0.05, 0.2, 16), (50000, 0.002, 0.006, 0.0, 16),
                                                     (2000, 0.01, 0.3, 0.5, 16), (5000, 0.001, 0.01, 0.1, 3), (1, 0.1, 0.2, 0.0, 16)])
def test_gpu_tree_equals_host_tree(ctx, n, dmin, dmax, boxes, depth):
    fb = scenes.random_spheres_flat(n, dmin, dmax, seed=7.0, mix="mirrors", box_fraction=boxes, max_in_depth=depth)
    g = rebuild_on_gpu(ctx, fb, max_in_depth=depth)
    assert_same_tree(fb.flat.arrays, g.arrays)
    # pre-order numbering: a parent precedes its children, children of a node ascend with the octant
    G = g.arrays
    assert G["node_parent"][0] == -1 and (G["node_parent"][1:] < np.arange(1, len(G["node_parent"]))).all()
    ch = G["node_child"]
    for row in ch[:200]:
        ex = row[row >= 0]
        assert (np.diff(ex) > 0).all()


@pytest.mark.parametrize("n,dmin,dmax,boxes", [(50000, 0.002, 0.006, 0.1), (4000, 0.004, 0.08, 0.3)])
def test_gpu_tree_equals_the_oracles_tree(ctx, oracle, n, dmin, dmax, boxes):
    """The device-built tree against the ORACLE's own restatement of add_entity_to_octree (src/octree_entity.ts:60-188),
    not against this repo's host builder: both number the nodes in depth-first pre-order with children 0..7, so the
    arrays must be equal element by element - float64 positions and sizes bit for bit, child tables, parents,
    index_within_parent, list offsets, and every node's entities in insertion order."""
    from util import oracle_scene_flat
    fb = scenes.random_spheres_flat(n, dmin, dmax, seed=11.0, mix="mirrors", box_fraction=boxes)
    G = rebuild_on_gpu(ctx, fb).arrays
    f = oracle_scene_flat(fb).flat()
    assert f.root_index == 0 and len(f.node_size) == len(G["node_size"])
    np.testing.assert_array_equal(G["node_pos"], f.node_pos)
    np.testing.assert_array_equal(G["node_size"], f.node_size)
    np.testing.assert_array_equal(G["node_child"], f.node_child)
    np.testing.assert_array_equal(G["node_parent"], f.node_parent)
    np.testing.assert_array_equal(G["node_octant"], f.node_octant)
    np.testing.assert_array_equal(G["node_list_off"], f.node_list_off)
    np.testing.assert_array_equal(G["list_entity"], f.list_entity)  # entity ids are insertion indices on both sides


def test_gpu_tree_reference_placement_and_errors(ctx):
    mat = rt.SolidMaterial(rt.ResponseType.REFLECTION, False, True, 0)
    tex = rt.SolidTexture(rt.Color(1, 1, 1, 1))
    # test/octree-entity.test.ts:52-64: (0.25,0.25,0.25) d 0.5 -> child 0 of the root; (0.5,0.25,0.5) d 0.25 -> the root
    g = flat_from_arrays([0, 0], [[0.25, 0.25, 0.25], [0.5, 0.25, 0.5]], [0.5, 0.25], [0, 0], [0, 0], [0, 0], [mat], [tex],
                         [rt.SUBSTANCE_AIR], max_in_depth=10, gpu_ctx=ctx).arrays
    assert len(g["node_size"]) == 2 and g["node_child"][0].tolist() == [1, -1, -1, -1, -1, -1, -1, -1]
    assert g["list_entity"][g["node_list_off"][0]:g["node_list_off"][1]].tolist() == [1]
    assert g["list_entity"][g["node_list_off"][1]:g["node_list_off"][2]].tolist() == [0]
    # an entity that does not fit the root: TreeOutsideGrowError, naming the first such entity
    with pytest.raises(rt.TreeOutsideGrowError, match="entity 1 "):
        flat_from_arrays([0, 0, 0], [[0.5, 0.5, 0.5], [0.99, 0.5, 0.5], [1.5, 0.5, 0.5]], [0.1, 0.1, 0.1], [0] * 3, [0] * 3, [0] * 3,
                         [mat], [tex], [rt.SUBSTANCE_AIR], gpu_ctx=ctx)
    # an empty scene still has its root
    g = flat_from_arrays(np.zeros(0, np.uint8), np.zeros((0, 3)), np.zeros(0), np.zeros(0, np.int32), np.zeros(0, np.int32),
                         np.zeros(0, np.int32), [mat], [tex], [rt.SUBSTANCE_AIR], gpu_ctx=ctx).arrays
    assert len(g["node_size"]) == 1 and g["node_list_off"].tolist() == [0, 0]
    # deeper than the 16 levels a path key holds: unsupported here, the host builder does it
    with pytest.raises(N.RtError):
        flat_from_arrays([0], [[0.5, 0.5, 0.5]], [0.1], [0], [0], [0], [mat], [tex], [rt.SUBSTANCE_AIR], max_in_depth=17, gpu_ctx=ctx)


def test_gpu_built_scene_renders_the_same_frame(ctx):
    """The two builders number the nodes differently; the uploaded scene (re-numbered breadth-first at upload)
    and the rendered frame must not care."""
    import torch
    lib = N.load()
    fb = scenes.random_spheres_flat(20000, 0.003, 0.02, seed=3.0, mix="mirrors", box_fraction=0.1)
    g = rebuild_on_gpu(ctx, fb)
    W = H = 256
    cd = rt.camera_desc(scenes.bench_camera(W, H))
    prm = N.Params()
    prm.refmax, prm.sky_texture, prm.default_substance = fb.refmax, fb.sky_texture, fb.default_substance
    prm.distance_attenuation_factor, prm.n_frames, prm.frame_first, prm.rng_seed = 1.0, 2, 0, 1.0
    frames = []
    for flat in (fb.flat, g):
        d = flat.desc()
        N.check(ctx, lib.rt_scene_upload(ctx, C.byref(d)))
        rgb, ids = np.zeros(W * H * 3, np.float32), np.zeros(W * H, np.int32)
        N.check(ctx, lib.rt_render(ctx, C.byref(cd), C.byref(prm), 0, rgb.ctypes.data, ids.ctypes.data, None))
        frames.append((rgb, ids))
    assert np.array_equal(frames[0][0], frames[1][0]) and np.array_equal(frames[0][1], frames[1][1])


def test_gpu_build_million_entities(ctx):
    """configs[3]/[4] size: 1 M entities; same tree as the host builder, and the time of both (printed)."""
    fb = scenes.random_spheres_flat(1000000, 0.0005, 0.002, seed=42.0, mix="mirrors", box_fraction=0.1)
    a = fb.flat.arrays
    lib = N.load()
    rp = np.zeros(3)
    times = {}
    for name in ("host", "gpu", "gpu"):
        tree = C.c_void_p()
        args = (rp.ctypes.data_as(N._dp), 1.0, len(a["ent_extent"]), a["ent_type"].ctypes.data_as(N._bp), a["ent_pos"].ctypes.data_as(N._dp),
                a["ent_extent"].ctypes.data_as(N._dp), 16, C.byref(tree))
        t0 = time.perf_counter()
        st = lib.rt_tree_build_gpu(ctx, *args) if name == "gpu" else lib.rt_tree_build(*args)
        times[name] = time.perf_counter() - t0
        N.check(ctx if name == "gpu" else None, st)
        assert lib.rt_tree_node_count(tree) == len(a["node_size"])
        lib.rt_tree_free(tree)
    print(f"\n1 M entities, {len(a['node_size'])} nodes: rt_tree_build {times['host']*1e3:.0f} ms, rt_tree_build_gpu {times['gpu']*1e3:.0f} ms (second call)")
    g = rebuild_on_gpu(ctx, fb)
    assert_same_tree(a, g.arrays)
