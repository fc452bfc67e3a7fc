"""OctreeWalker (src/octree_space.ts:159-408), node_at_pos (:61-93), octant_adj_pos (:41-50), index_within_parent
(:110-125), dim_relative_to_parent (:127-136) and Box.line_intersection (src/math/intersection.ts:150-204)
transliterated into plain Python over a flat tree (arrays node_pos / node_size / node_child / node_parent), to
check the C++ oracle's walker with code that shares nothing with it.  TEST INFRASTRUCTURE ONLY (small cases: pure-Python loops)."""
import math

import numpy as np

FACE_NORMALS = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]


def to_int32(x):  # the ToInt32 of `x << 0`
    if not math.isfinite(x):
        return 0
    v = int(math.trunc(x)) & 0xffffffff
    return v - (1 << 32) if v >= (1 << 31) else v


def jsdiv(a, b):
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.float64(a) / np.float64(b))


def box_line(centre, size, start, d):
    """Box.line_intersection -> [] or [(u1, normal1), (u2, normal2)]"""
    tl = [centre[k] - size[k] * 0.5 for k in range(3)]
    p = [-d[0], d[0], -d[1], d[1], -d[2], d[2]]
    q = [start[0] - tl[0], tl[0] + size[0] - start[0], start[1] - tl[1], tl[1] + size[1] - start[1], start[2] - tl[2], tl[2] + size[2] - start[2]]
    u1, u2, i1, i2 = -math.inf, math.inf, None, None
    for i in range(6):
        u = jsdiv(q[i], p[i])
        if p[i] < 0 or (p[i] == 0 and math.copysign(1.0, p[i]) < 0):
            if u > u1:
                u1, i1 = u, i
        elif u < u2:
            u2, i2 = u, i
    if u1 > u2:
        return []
    return [(u1, FACE_NORMALS[i1] if i1 is not None else None), (u2, FACE_NORMALS[i2] if i2 is not None else None)]


class Tree:
    def __init__(self, flat):
        self.pos = [list(map(float, p)) for p in flat.node_pos]
        self.size = [float(s) for s in flat.node_size]
        self.child = [[int(c) for c in row] for row in flat.node_child]
        self.parent = [int(p) for p in flat.node_parent]
        self.root = int(getattr(flat, "root_index", 0))

    def get(self, n, o):
        if o < 0 or o > 7:
            raise IndexError("Octree.get: index out of range")  # src/octree.ts:48-54
        c = self.child[n][o]
        return c if c >= 0 else None

    def octant_adj_pos(self, n, p):  # :41-50
        h = self.size[n] / 2
        return (int(p[2] >= self.pos[n][2] + h) << 2) | (int(p[1] >= self.pos[n][1] + h) << 1) | int(p[0] >= self.pos[n][0] + h)

    def index_within_parent(self, n):  # :110-125 (the cached field is never set anywhere in the reference)
        par = self.parent[n]
        if par < 0:
            return None
        k = 2 / self.size[par]
        ind = [(self.pos[n][i] - self.pos[par][i]) * k for i in range(3)]
        return (to_int32(ind[2]) << 2) + (to_int32(ind[1]) << 1) + (to_int32(ind[0]) << 0)

    def node_at_pos(self, p):  # :61-93, from the root, CLOSE_OPEN
        n = self.root
        dpos, dsize = list(self.pos[n]), self.size[n]
        if not all(p[i] >= dpos[i] and p[i] < dpos[i] + dsize for i in range(3)):
            return None
        cur, idx, nxt = n, 0, n
        while nxt is not None:
            k = 2 / dsize
            ind = [(p[i] - dpos[i]) * k for i in range(3)]
            cur = nxt
            idx = (to_int32(ind[2]) << 2) + (to_int32(ind[1]) << 1) + (to_int32(ind[0]) << 0)
            nxt = self.get(cur, idx)
            dsize /= 2
            for i in range(3):
                dpos[i] += to_int32(ind[i]) * dsize
        return cur, idx


class WalkerWouldThrow(Exception):
    pass


class Walker:
    def __init__(self, tree, include_undefined=False):
        self.t, self.include_undefined = tree, include_undefined

    def set_pos_and_dir(self, pos, direction, node=None):  # :196-205,236-239
        self.direction = list(direction)
        self.cur_node = node if node is not None else self.t.node_at_pos(pos)  # (tree, octant) | None
        self.pos = list(pos)
        return self.setup_cur_node()

    def reset_state(self):  # :251-257
        self.next_pos = [self.pos, None]
        self.cur_returned = self.stepped_in = self.next_pos_is_ahead = False
        self.depth = 0

    def setup_cur_node(self):  # :262-290
        self.reset_state()
        if self.cur_node is not None:
            return True
        r = self.t.root
        c = [self.t.pos[r][i] + 0.5 * self.t.size[r] for i in range(3)]
        inter = [x for x in box_line(c, [self.t.size[r]] * 3, self.pos, self.direction) if x[0] >= 0]
        if not inter:
            return False
        u, nrm = inter[0]
        ipoint = [self.pos[i] + self.direction[i] * u for i in range(3)]
        self.cur_node = (r, None)
        self.next_pos = [ipoint, [-v for v in nrm]]
        return True

    def step_back(self):  # :292-322
        self.stepped_in = True
        tree, octant = self.cur_node
        if octant is None:
            self.cur_node = None
            self.cur_returned = False
            return
        if self.depth > 0:
            self.depth -= 1
            self.cur_returned = True
        else:
            self.cur_returned = False
        g = self.t.index_within_parent(tree)
        self.cur_node = (self.t.parent[tree], g) if g is not None else (tree, None)

    def update_next_pos(self):  # :381-395
        tree, octant = self.cur_node
        ph = self.t.size[tree] / 2  # dim_relative_to_parent
        dpos = [self.t.pos[tree][i] + ((octant >> i) & 1) * ph for i in range(3)]
        c = [dpos[i] + 0.5 * ph for i in range(3)]
        inter = box_line(c, [ph] * 3, self.pos, self.direction)
        if not inter:
            raise WalkerWouldThrow("update_next_pos: inter_param.pop() is undefined")
        u, nrm = inter[-1]
        if nrm is None:
            raise WalkerWouldThrow("update_next_pos: no exit face")
        self.next_pos = [[self.pos[i] + self.direction[i] * u for i in range(3)], list(nrm)]

    def next(self):  # :330-373 -> (pos.tree | -1, pos.octant | -1, node | -1) or None
        while self.cur_node is not None:
            tree, octant = self.cur_node
            node = self.t.get(tree, octant) if octant is not None else tree
            if not self.cur_returned:
                if self.include_undefined or node is not None:
                    self.cur_returned = True
                    return (tree if octant is not None else -1, octant if octant is not None else -1, node if node is not None else -1)
            if octant is not None:
                if not self.next_pos_is_ahead:
                    if not self.stepped_in and node is not None:
                        self.depth += 1  # step_in
                        self.cur_node = (node, self.t.octant_adj_pos(node, self.next_pos[0]))
                        self.cur_returned = False
                        continue
                    self.update_next_pos()
                nxt = [((octant >> i) & 1) + self.next_pos[1][i] for i in range(3)]
                if not any(x < 0 or x > 1 for x in nxt):
                    self.cur_node = (tree, int(nxt[0]) | (int(nxt[1]) << 1) | (int(nxt[2]) << 2))
                    self.cur_returned = False
                    self.stepped_in = False
                    self.next_pos_is_ahead = False
                    continue
                self.next_pos_is_ahead = True
            self.step_back()
        return None

    def stops(self, pos, direction, use_start_node=False, limit=100000):
        self.set_pos_and_dir(pos, direction, self.t.node_at_pos(pos) if use_start_node else None)
        out = []
        while True:
            s = self.next()
            if s is None or len(out) > limit:
                return out
            out.append(s)
