"""The oracle's OctreeWalker and node_at_pos against oracle/pywalker.py, a plain-Python transliteration of
src/octree_space.ts that shares no code with it: identical stop sequences (tree, octant, node) on deep random
trees for random rays, axis-parallel rays (zero and negative-zero components), origins on dyadic planes and on the
root's faces, origins outside the root, with and without include_undefined.  The reference's own two-level jest
vector pins both; this widens the pin to trees of real depth."""
import math

import numpy as np
import pytest

from oracle import pywalker


def random_tree(oracle, seed, n, dmin, dmax):
    rng = np.random.default_rng(seed)
    s = oracle.Scene((0, 0, 0), 1.0)
    m, t, sub = s.add_material(0, False, True, 0.0), s.add_texture_solid(1, 1, 1, 1), s.add_substance(1.0)
    for _ in range(n):
        d = float(rng.uniform(dmin, dmax))
        c = (d / 2 + rng.uniform(0, 1, 3) * (1 - d)).tolist()
        s.add_entity(0, c, d, m, t, sub, max_in_depth=12, max_out_depth=0)
    return s


def rays(seed, count):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(count):
        o = rng.uniform(0, 1, 3)
        d = rng.normal(size=3)
        kind = i % 8
        if kind == 1:
            d[rng.integers(0, 3)] = 0.0
        elif kind == 2:
            d[rng.integers(0, 3)] = -0.0
        elif kind == 3:
            o = np.round(o * 8) / 8  # on dyadic planes
        elif kind == 4:
            o[rng.integers(0, 3)] = 0.0  # on a face of the root (inside: half-open)
        elif kind == 5:
            o = o * 3 - 1  # mostly outside the root
        elif kind == 6:
            d = np.array([0.0, 0.0, 0.0])
            d[rng.integers(0, 3)] = rng.choice([-1.0, 1.0])  # along an axis
        out.append((o.tolist(), d.tolist()))
    return out


@pytest.mark.parametrize("seed,n,dmin,dmax", [(1, 40, 0.05, 0.3), (2, 400, 0.01, 0.08), (3, 1500, 0.002, 0.02)])
def test_walker_and_node_at_pos_equal_the_transliteration(oracle, seed, n, dmin, dmax):
    s = random_tree(oracle, seed, n, dmin, dmax)
    tree = pywalker.Tree(s.flat())
    def level(i):
        d = 0
        while tree.parent[i] >= 0:
            i, d = tree.parent[i], d + 1
        return d
    depth = max(level(i) for i in range(len(tree.size)))
    assert depth >= (2 if n < 100 else 4)
    compared = thrown = 0
    for o, d in rays(seed + 100, 300):
        want_nap = tree.node_at_pos(o)
        got_nap = s.node_at_pos(o)
        assert got_nap == (want_nap if want_nap is not None else (-1, -1)), (o, want_nap, got_nap)
        for include_undefined in (False, True):
            for use_start in (False, True):
                try:
                    want = pywalker.Walker(tree, include_undefined).stops(o, d, use_start_node=use_start)
                except pywalker.WalkerWouldThrow:
                    thrown += 1
                    continue
                got = [tuple(int(v) for v in row) for row in s.walk(o, d, include_undefined=include_undefined, use_start_node=use_start, max_stops=200000)]
                assert got == want, (o, d, include_undefined, use_start, got[:12], want[:12])
                compared += 1
    print(f"tree depth {depth}, {len(tree.size)} nodes: {compared} walks compared, {thrown} where the reference itself would throw")
    assert compared > 1000 and thrown < 30, (compared, thrown)
