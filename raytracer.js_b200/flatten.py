"""Flattener: EntityOtree (pointer octree of the host API) -> rt_scene_desc (include/rt_b200.h).

Nodes are numbered in DFS pre-order with children visited 0..7; every node's entity list keeps the
EntitySet insertion order, which is what the reference's first-hit rule depends on
(src/raytracer.ts:186-195).  Unknown Entity / Material / Texture / Sky subclasses raise: there is
no CPU fallback to hand them to."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import numpy as np

from . import _native as N
from .entity import BoxEntity, Entity, SphereEntity
from .material import StaticMaterial
from .octree import Octree
from .octree_space import index_within_parent
from .substance import Substance
from .texture import ImageTexture, SolidTexture, Texture


class FlatScene:
    """Owns the numpy arrays an rt_scene_desc points into."""

    def __init__(self):
        self.arrays: Dict[str, np.ndarray] = {}
        self.entities: List[Entity] = []
        self.materials: List[StaticMaterial] = []
        self.textures: List[Texture] = []
        self.substances: List[Substance] = []
        self._tex_index: Dict[int, int] = {}
        self._mat_index: Dict[int, int] = {}
        self._sub_index: Dict[int, int] = {}

    # -- table interning ----------------------------------------------------------------------
    def texture_index(self, t: Texture) -> int:
        if not isinstance(t, (SolidTexture, ImageTexture)):
            raise TypeError(f"unsupported Texture subclass {type(t).__name__}")
        k = id(t)
        if k not in self._tex_index:
            self._tex_index[k] = len(self.textures)
            self.textures.append(t)
        return self._tex_index[k]

    def material_index(self, m) -> int:
        if not isinstance(m, StaticMaterial):
            raise TypeError(f"unsupported Material subclass {type(m).__name__}")
        k = id(m)
        if k not in self._mat_index:
            self._mat_index[k] = len(self.materials)
            self.materials.append(m)
        return self._mat_index[k]

    def substance_index(self, s) -> int:
        if s is None:
            return -1
        if not isinstance(s, Substance):
            raise TypeError(f"unsupported Substance {type(s).__name__}")
        k = id(s)
        if k not in self._sub_index:
            self._sub_index[k] = len(self.substances)
            self.substances.append(s)
        return self._sub_index[k]

    # -- the C struct -------------------------------------------------------------------------
    def desc(self) -> N.SceneDesc:
        a = self.arrays
        d = N.SceneDesc()
        d.struct_size = C.sizeof(N.SceneDesc)

        def ptr(name, typ):
            arr = a[name]
            return arr.ctypes.data_as(typ) if arr.size else C.cast(None, typ)

        d.n_nodes = len(a["node_size"])
        d.node_pos, d.node_size = ptr("node_pos", N._dp), ptr("node_size", N._dp)
        d.node_child, d.node_parent, d.node_octant = ptr("node_child", N._ip), ptr("node_parent", N._ip), ptr("node_octant", N._ip)
        d.node_list_off = ptr("node_list_off", N._up)
        d.n_list = len(a["list_entity"])
        d.list_entity = ptr("list_entity", N._up)
        d.n_entities = len(a["ent_extent"])
        d.ent_type, d.ent_pos, d.ent_extent = ptr("ent_type", N._bp), ptr("ent_pos", N._dp), ptr("ent_extent", N._dp)
        d.ent_material, d.ent_texture, d.ent_substance = ptr("ent_material", N._ip), ptr("ent_texture", N._ip), ptr("ent_substance", N._ip)
        d.n_materials = len(a["mat_roughness"])
        d.mat_response, d.mat_light, d.mat_mirror = ptr("mat_response", N._bp), ptr("mat_light", N._bp), ptr("mat_mirror", N._bp)
        d.mat_roughness = ptr("mat_roughness", N._dp)
        d.n_textures = len(a["tex_kind"])
        d.tex_kind, d.tex_color = ptr("tex_kind", N._bp), ptr("tex_color", N._dp)
        d.tex_width, d.tex_height = ptr("tex_width", N._ip), ptr("tex_height", N._ip)
        d.tex_loaded, d.tex_texel_off = ptr("tex_loaded", N._bp), ptr("tex_texel_off", N._qp)
        d.n_texels = a["texels"].size // 3
        d.texels = ptr("texels", N._bp)
        d.n_substances = len(a["sub_refractive_index"])
        d.sub_refractive_index = ptr("sub_refractive_index", N._dp)
        return d

    def finalize_tables(self) -> None:
        """(Re)build the material / texture / substance arrays from the interned objects."""
        a = self.arrays
        m = self.materials
        a["mat_response"] = np.array([int(x.response) for x in m], np.uint8)
        a["mat_light"] = np.array([int(x.light_source) for x in m], np.uint8)
        a["mat_mirror"] = np.array([int(x.mirror) for x in m], np.uint8)
        a["mat_roughness"] = np.array([x.roughness_index for x in m], np.float64)
        t = self.textures
        nt = len(t)
        a["tex_kind"] = np.zeros(nt, np.uint8)
        a["tex_color"] = np.zeros((nt, 4), np.float64)
        a["tex_width"] = np.zeros(nt, np.int32)
        a["tex_height"] = np.zeros(nt, np.int32)
        a["tex_loaded"] = np.zeros(nt, np.uint8)
        a["tex_texel_off"] = np.zeros(nt, np.uint64)
        pool, off = [], 0
        for i, x in enumerate(t):
            if isinstance(x, SolidTexture):
                c = x.color
            else:
                a["tex_kind"][i] = N.RT_TEXTURE_IMAGE
                c = x.fallback_color
                if x.image_data is not None:
                    a["tex_loaded"][i] = 1
                    a["tex_width"][i], a["tex_height"][i] = x.width, x.height
                    a["tex_texel_off"][i] = off
                    pool.append(x.image_data.reshape(-1))
                    off += x.width * x.height
            a["tex_color"][i] = (c.r, c.g, c.b, c.a)
        a["texels"] = np.concatenate(pool) if pool else np.zeros(0, np.uint8)
        a["sub_refractive_index"] = np.array([s.refractive_index for s in self.substances], np.float64)


def flatten_scene(tree: Octree, extra_textures=(), extra_substances=()) -> FlatScene:
    """Walk the octree handed to the Raytracer and produce the flat description."""
    if tree.parent is not None:
        raise ValueError("unsupported octree: the tree handed to the raytracer must be the absolute root "
                         "(tree.parent == undefined)")
    fs = FlatScene()
    for t in extra_textures:
        fs.texture_index(t)
    for s in extra_substances:
        fs.substance_index(s)
    nodes: List[Octree] = []
    parent: List[int] = []
    stack = [(tree, -1)]
    # iterative DFS pre-order, children 0..7
    while stack:
        node, par = stack.pop()
        idx = len(nodes)
        nodes.append(node)
        parent.append(par)
        for c in range(7, -1, -1):
            ch = node.get(c)
            if ch is not None:
                stack.append((ch, idx))
    index = {id(n): i for i, n in enumerate(nodes)}
    n = len(nodes)
    node_pos = np.zeros((n, 3))
    node_size = np.zeros(n)
    node_child = np.full((n, 8), -1, np.int32)
    node_octant = np.full(n, -1, np.int32)
    list_off = np.zeros(n + 1, np.uint32)
    list_entity: List[int] = []
    ent_type, ent_pos, ent_extent, ent_mat, ent_tex, ent_sub = [], [], [], [], [], []
    for i, node in enumerate(nodes):
        node_pos[i] = node.id.pos.v
        node_size[i] = node.id.size
        for c in range(8):
            ch = node.get(c)
            if ch is not None:
                node_child[i, c] = index[id(ch)]
        if i > 0:
            node_octant[i] = index_within_parent(node)
        list_off[i] = len(list_entity)
        for e in node.value.set:
            if isinstance(e, SphereEntity):
                ent_type.append(N.RT_ENTITY_SPHERE)
                ent_extent.append(e.get_diameter())
            elif isinstance(e, BoxEntity):
                ent_type.append(N.RT_ENTITY_BOX)
                ent_extent.append(e.get_size())
            else:
                raise TypeError(f"unsupported Entity subclass {type(e).__name__}")
            ent_pos.append(e.get_pos().v)
            ent_mat.append(fs.material_index(e.get_material()))
            ent_tex.append(fs.texture_index(e.get_texture()))
            ent_sub.append(fs.substance_index(e.get_substance()))
            list_entity.append(len(fs.entities))
            fs.entities.append(e)
    list_off[n] = len(list_entity)
    a = fs.arrays
    a["node_pos"], a["node_size"], a["node_child"] = node_pos, node_size, node_child
    a["node_parent"] = np.array(parent, np.int32)
    a["node_octant"], a["node_list_off"] = node_octant, list_off
    a["list_entity"] = np.array(list_entity, np.uint32)
    a["ent_type"] = np.array(ent_type, np.uint8)
    a["ent_pos"] = np.array(ent_pos, np.float64).reshape(-1, 3)
    a["ent_extent"] = np.array(ent_extent, np.float64)
    a["ent_material"] = np.array(ent_mat, np.int32)
    a["ent_texture"] = np.array(ent_tex, np.int32)
    a["ent_substance"] = np.array(ent_sub, np.int32)
    fs.finalize_tables()
    return fs


def flat_from_arrays(ent_type, ent_pos, ent_extent, ent_material, ent_texture, ent_substance, materials, textures,
                     substances, root_pos=(0.0, 0.0, 0.0), root_size=1.0, max_in_depth=16, gpu_ctx=None) -> FlatScene:
    """Bulk path for big scenes: the octree is built by the native restatement of add_entity_to_octree
    (rt_tree_build in librt_b200, csrc/rt_build.h) straight from entity arrays, without Python entity
    objects or a pointer tree.  Entity ids are the array indices (= insertion order).  `materials`,
    `textures`, `substances` are lists of the host API's objects, indexed by the ent_* index arrays.
    gpu_ctx: an rt_ctx (GpuRaytracer.ctx) -> the tree is built on that GPU (rt_tree_build_gpu, csrc/rt_build_gpu.cuh):
    the same tree, nodes numbered in depth-first pre-order."""
    lib = N.load()
    fs = FlatScene()
    for m in materials:
        fs.material_index(m)
    for t in textures:
        fs.texture_index(t)
    for s in substances:
        fs.substance_index(s)
    a = fs.arrays
    a["ent_type"] = np.ascontiguousarray(ent_type, np.uint8)
    a["ent_pos"] = np.ascontiguousarray(ent_pos, np.float64).reshape(-1, 3)
    a["ent_extent"] = np.ascontiguousarray(ent_extent, np.float64)
    a["ent_material"] = np.ascontiguousarray(ent_material, np.int32)
    a["ent_texture"] = np.ascontiguousarray(ent_texture, np.int32)
    a["ent_substance"] = np.ascontiguousarray(ent_substance, np.int32)
    n = len(a["ent_extent"])
    rp = np.ascontiguousarray(root_pos, np.float64)
    tree = C.c_void_p()
    args = (rp.ctypes.data_as(N._dp), float(root_size), n, a["ent_type"].ctypes.data_as(N._bp),
            a["ent_pos"].ctypes.data_as(N._dp), a["ent_extent"].ctypes.data_as(N._dp), int(max_in_depth), C.byref(tree))
    st = lib.rt_tree_build_gpu(gpu_ctx, *args) if gpu_ctx is not None else lib.rt_tree_build(*args)
    if st != N.RT_OK:
        from .octree_entity import TreeOutsideGrowError
        msg = lib.rt_last_error(gpu_ctx).decode()
        if st == N.RT_ERR_UNSUPPORTED and "outside-depth" in msg:
            raise TreeOutsideGrowError(None, msg)
        raise N.RtError(st, msg)
    try:
        nn = lib.rt_tree_node_count(tree)
        a["node_pos"] = np.zeros((nn, 3))
        a["node_size"] = np.zeros(nn)
        a["node_child"] = np.zeros((nn, 8), np.int32)
        a["node_parent"] = np.zeros(nn, np.int32)
        a["node_octant"] = np.zeros(nn, np.int32)
        a["node_list_off"] = np.zeros(nn + 1, np.uint32)
        a["list_entity"] = np.zeros(n, np.uint32)
        lib.rt_tree_export(tree, a["node_pos"].ctypes.data_as(N._dp), a["node_size"].ctypes.data_as(N._dp),
                           a["node_child"].ctypes.data_as(N._ip), a["node_parent"].ctypes.data_as(N._ip),
                           a["node_octant"].ctypes.data_as(N._ip), a["node_list_off"].ctypes.data_as(N._up),
                           a["list_entity"].ctypes.data_as(N._up))
    finally:
        lib.rt_tree_free(tree)
    fs.finalize_tables()
    return fs
