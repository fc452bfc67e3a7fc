"""The product's host-side octree build + flattener (raytracer.js_b200/octree_entity.py, flatten.py)
against the oracle's independent restatement of add_entity_to_octree: identical trees and lists."""
import numpy as np
import pytest

import raytracer_js_b200 as rt
from raytracer_js_b200 import scenes

from util import flat_of, oracle_scene


def assert_same_tree(flat, ofl, bundle):
    a = flat.arrays
    assert len(a["node_size"]) == len(ofl.node_size)
    np.testing.assert_array_equal(a["node_pos"], ofl.node_pos)
    np.testing.assert_array_equal(a["node_size"], ofl.node_size)
    np.testing.assert_array_equal(a["node_child"], ofl.node_child)
    np.testing.assert_array_equal(a["node_parent"], ofl.node_parent)
    np.testing.assert_array_equal(a["node_octant"], ofl.node_octant)
    np.testing.assert_array_equal(a["node_list_off"], ofl.node_list_off)
    order = {id(e): i for i, e in enumerate(bundle.entities)}
    mine = np.array([order[id(flat.entities[i])] for i in a["list_entity"]], np.uint32)
    np.testing.assert_array_equal(mine, ofl.list_entity)  # same nodes, same insertion order inside each node


@pytest.mark.parametrize("n,dmin,dmax,boxes", [(300, 0.01, 0.05, 0.0), (2000, 0.002, 0.006, 0.2), (500, 0.05, 0.3, 0.5)])
def test_tree_matches_oracle(oracle, n, dmin, dmax, boxes):
    b = scenes.random_spheres(n, dmin, dmax, seed=7.0, mix="mirrors", box_fraction=boxes)
    flat = flat_of(b)
    os_ = oracle_scene(flat, b)
    assert_same_tree(flat, os_.flat(), b)


def test_reference_placement_cases():
    # test/octree-entity.test.ts:52-64 through the product's host API
    tree = rt.new_entity_octree(rt.OctreeDim(rt.point(0, 0, 0), 1), None)
    tex = rt.SolidTexture(rt.Color(1, 1, 1, 1))

    def add(pos, size):
        e = rt.SphereEntity(None, rt.SIMPLE_SMOOTH_MATERIAL, tex, rt.SUBSTANCE_AIR, pos, size)
        rt.add_entity_to_octree(tree, e, {"max_in_depth": 10, "max_out_depth": 10})
        return e

    e = add(rt.point(0.25, 0.25, 0.25), 0.5)
    assert e in tree.get(0).value.set
    e = add(rt.point(0.5, 0.25, 0.5), 0.25)
    assert e in tree.value.set


def test_node_at_pos_reference_case():
    # test/octree-space.test.ts:36-46 through the product's host API
    tree = rt.Octree(rt.OctreeDim(rt.point(0, 0, 0), 1))
    inner = rt.new_subtree(tree, 3)
    rt.new_subtree(inner, 5)
    nd = rt.node_at_pos(tree, rt.point(0.75, 0.5, 0.25))
    assert nd == rt.OctreePos(tree.subtree(3).subtree(5), 0)


def test_octree_get_bounds():
    # test/octree.test.ts:3-7
    o = rt.Octree(0)
    with pytest.raises(IndexError):
        o.get(8)
    with pytest.raises(IndexError):
        o.get(-1)


def test_outside_growth_and_error(oracle):
    tree = rt.new_entity_octree(rt.OctreeDim(rt.point(0, 0, 0), 1), None)
    tex = rt.SolidTexture(rt.Color(1, 1, 1, 1))
    e = rt.SphereEntity(None, rt.SIMPLE_SMOOTH_MATERIAL, tex, rt.SUBSTANCE_AIR, rt.point(1.5, 0.5, 0.5), 0.25)
    with pytest.raises(rt.TreeOutsideGrowError):
        rt.add_entity_to_octree(tree, e, {"max_in_depth": 4, "max_out_depth": 0})
    node = rt.add_entity_to_octree(tree, e, {"max_in_depth": 4, "max_out_depth": 3})
    assert tree.parent is not None and node.get_root() is tree.get_root()
    # the oracle grows the same way
    s = oracle.Scene((0, 0, 0), 1.0)
    m, t, sub = s.add_material(0, False, True, 0), s.add_texture_solid(1, 1, 1), s.add_substance(1.0)
    with pytest.raises(RuntimeError):
        s.add_entity(0, (1.5, 0.5, 0.5), 0.25, m, t, sub, 4, 0)
    s.add_entity(0, (1.5, 0.5, 0.5), 0.25, m, t, sub, 4, 3)
    f = s.flat()
    assert f.root_index != 0 and f.node_size[0] == tree.get_root().id.size
    # a grown tree is not accepted by the flattener (the reference's own node_at_pos mixes up the
    # dimensions of the passed tree and of the absolute root in that case)
    with pytest.raises(ValueError):
        rt.flatten_scene(tree)


def canonical_tree(a):
    """Numbering-independent description of a flat tree: {(pos, size): (sorted child octants, entity list)}."""
    out = {}
    for i in range(len(a["node_size"])):
        key = (tuple(a["node_pos"][i].tolist()), float(a["node_size"][i]))
        beg, end = int(a["node_list_off"][i]), int(a["node_list_off"][i + 1])
        out[key] = (tuple(int(c >= 0) for c in a["node_child"][i]), beg, end)
    return out


@pytest.mark.parametrize("mix,box_fraction", [("diffuse", 0.0), ("mirrors", 0.2)])
def test_bulk_builder_equals_one_by_one_insertion(mix, box_fraction):
    """rt_tree_build (native bulk restatement of add_entity_to_octree) against the host API inserting the
    same entities one at a time: same nodes (float64 positions and sizes), same per-node entity order, same
    entity / material / texture tables."""
    from raytracer_js_b200 import scenes
    n = 3000
    b = scenes.random_spheres(n, 0.004, 0.05, seed=7.0, mix=mix, box_fraction=box_fraction)
    ref = rt.flatten_scene(b.tree, extra_textures=[b.sky.texture], extra_substances=[b.default_substance])
    fb = scenes.random_spheres_flat(n, 0.004, 0.05, seed=7.0, mix=mix, box_fraction=box_fraction)
    A, B = ref.arrays, fb.flat.arrays
    order = {id(e): i for i, e in enumerate(b.entities)}
    ins = np.array([order[id(e)] for e in ref.entities])  # flattener id -> insertion index
    np.testing.assert_array_equal(B["ent_pos"][ins], A["ent_pos"])
    np.testing.assert_array_equal(B["ent_extent"][ins], A["ent_extent"])
    np.testing.assert_array_equal(B["ent_type"][ins], A["ent_type"])
    np.testing.assert_array_equal(B["tex_color"][B["ent_texture"][ins]], A["tex_color"][A["ent_texture"]])
    np.testing.assert_array_equal(B["mat_mirror"][B["ent_material"][ins]], A["mat_mirror"][A["ent_material"]])
    np.testing.assert_array_equal(B["mat_light"][B["ent_material"][ins]], A["mat_light"][A["ent_material"]])
    ta, tb = canonical_tree(A), canonical_tree(B)
    assert ta.keys() == tb.keys()
    for key in ta:
        ca, ba, ea = ta[key]
        cb, bb, eb = tb[key]
        assert ca == cb
        assert ins[A["list_entity"][ba:ea]].tolist() == B["list_entity"][bb:eb].tolist(), key


def test_bulk_builder_rejects_entities_outside_the_root():
    from raytracer_js_b200.flatten import flat_from_arrays
    mat = rt.SolidMaterial(rt.ResponseType.REFLECTION, False, False, 0)
    tex = rt.SolidTexture(rt.Color(1, 1, 1, 1))
    with pytest.raises(rt.TreeOutsideGrowError):
        flat_from_arrays([0], [[0.99, 0.5, 0.5]], [0.1], [0], [0], [0], [mat], [tex], [rt.SUBSTANCE_AIR])


def test_image_texture_from_file(tmp_path):
    """ImageTexture.from_file (host stand-in for the browser-only load_image, src/texture/texture_image.ts:76-136):
    RGB8 texels in row order, flips as the reference's index arithmetic does them, fallback when undecodable."""
    from PIL import Image
    px = np.arange(2 * 3 * 3, dtype=np.uint8).reshape(2, 3, 3) * 10
    rgba = np.dstack([px, np.full((2, 3, 1), 128, np.uint8)])
    Image.fromarray(rgba, "RGBA").save(tmp_path / "t.png")
    fb = rt.Color(1, 0, 1, 1)
    t = rt.ImageTexture.from_file(str(tmp_path / "t.png"), fb)
    assert t.get_size() == (3, 2) and np.array_equal(t.image_data, px)  # alpha dropped, rows top to bottom
    t = rt.ImageTexture.from_file(str(tmp_path / "t.png"), fb, horizontal_flip=True, vertical_flip=True)
    assert np.array_equal(t.image_data, px[::-1, ::-1])
    missing = rt.ImageTexture.from_file(str(tmp_path / "nope.png"), fb)
    assert missing.get_size() is None and missing.image_data is None and missing.fallback_color is fb
