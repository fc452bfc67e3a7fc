"""rt_pack_scene (csrc/rt_host.h) validates and packs a scene on the host's cores (the flatten /
upload side of the path).  The packed arrays - breadth-first node records, slots, the per-list BVHs, the walk
records - must not depend on the number of packing threads, and a broken scene must be refused with the error a
single thread walking the arrays in order meets first."""
import ctypes as C

import numpy as np
import pytest

from raytracer_js_b200 import _native as N
from raytracer_js_b200 import scenes

from util import hostsim


def digest(flat, threads, monkeypatch):
    monkeypatch.setenv("RT_B200_PACK_THREADS", str(threads))
    L = hostsim()
    L.hostsim_pack_digest.restype = C.c_int
    d = flat.desc()
    out, err = C.c_uint64(0), C.create_string_buffer(512)
    st = L.hostsim_pack_digest(C.byref(d), C.byref(out), err, 512)
    return st, out.value, err.value.decode()


@pytest.mark.parametrize("n,dmin,dmax,boxes", [(70_000, 0.002, 0.006, 0.1), (90_000, 0.0005, 0.02, 0.0)])
def test_packed_scene_does_not_depend_on_the_thread_count(monkeypatch, n, dmin, dmax, boxes):
    fb = scenes.random_spheres_flat(n, dmin, dmax, seed=3.0, mix="mirrors", box_fraction=boxes)
    want = digest(fb.flat, 1, monkeypatch)
    assert want[0] == N.RT_OK, want
    for threads in (2, 5, 8):
        assert digest(fb.flat, threads, monkeypatch) == want


def test_first_error_is_the_single_thread_one(monkeypatch):
    """Several defects at once: every thread count reports the one with the lowest index."""
    fb = scenes.random_spheres_flat(80_000, 0.002, 0.006, seed=5.0, mix="mirrors")
    flat = fb.flat
    ext = flat.arrays["ent_extent"].copy()
    lists = flat.arrays["list_entity"]
    # (the slots follow the breadth-first node order; the list order of the desc is the flattener's: take the defects'
    # places from a single-thread run instead of assuming one)
    bad = [int(lists[len(lists) // 7]), int(lists[len(lists) // 2]), int(lists[-5])]
    for e in bad:
        ext[e] = -1.0
    keep = flat.arrays["ent_extent"]
    flat.arrays["ent_extent"] = ext
    try:
        want = digest(flat, 1, monkeypatch)
        assert want[0] == N.RT_ERR_INVALID and "bad extent" in want[2]
        for threads in (3, 8):
            assert digest(flat, threads, monkeypatch) == want
    finally:
        flat.arrays["ent_extent"] = keep


def test_entity_listed_twice_is_refused_with_any_thread_count(monkeypatch):
    fb = scenes.random_spheres_flat(70_000, 0.002, 0.006, seed=6.0, mix="mirrors")
    flat = fb.flat
    keep = flat.arrays["list_entity"]
    lists = keep.copy()
    lists[len(lists) // 3] = lists[len(lists) // 3 + 1000]
    flat.arrays["list_entity"] = lists
    try:
        for threads in (1, 8):
            st, _, msg = digest(flat, threads, monkeypatch)
            assert st == N.RT_ERR_INVALID and "listed in more than one node" in msg
    finally:
        flat.arrays["list_entity"] = keep
