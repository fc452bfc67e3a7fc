// oracle/oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A CPU, double-precision restatement of the per-pixel ray hot path of
// Dark565/raytracer.js, written to be read side by side with the reference
// (every function cites the reference file:line it follows; paths are relative
// to /root/reference).  It is the *checker* for the CUDA path in
// raytracer.js_b200/csrc: only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it.  The product never does.
//
// Pinning status ("restated, not executed reference"): no JavaScript runtime
// exists in the build container or on the GPU box, so the reference itself was
// never executed.  The restatement is pinned against every golden vector the
// reference's own jest suites hold for this path (tests/test_oracle_golden.py:
// test/octree-space-walker.test.ts:29-35,57-70, test/octree-space.test.ts:6-46,
// test/octree-entity.test.ts:52-64, test/view-camera.test.ts:17-49,
// test/octree.test.ts:3-7).  Everything those suites do not cover (hit tests,
// first-hit rule, shading, RNG values, blend) is pinned only by the source
// lines cited below: "parity unpinned by tests" for those rows.  They are cross-checked by
// tests/test_oracle_by_hand.py (hand-derived known answers, independent transliterations of
// single functions), tests/test_oracle_vs_python_restatement.py and
// tests/test_oracle_walker_vs_python.py (oracle/pyref.py + oracle/pywalker.py: a second restatement
// of Ray.trace, the hit tests, the OctreeWalker and node_at_pos in plain Python that shares no
// logic with this file and must agree bit for bit).
//
// Build: g++ -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile): IEEE
// double everywhere, no FMA contraction, JS `%` == fmod, `x<<0` == ToInt32.

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr double JS_EPSILON = 2.220446049250313e-16;  // Number.EPSILON
constexpr double JS_PI = 3.141592653589793;            // Math.PI

using V3 = std::array<double, 3>;
using V2 = std::array<double, 2>;

// ---------------------------------------------------------------- JS semantics
// ECMAScript ToInt32, used by `x << 0` / `x << n` on doubles.
int32_t js_int32(double x) {
	if (!std::isfinite(x)) return 0;
	double t = std::trunc(x);
	double m = std::fmod(t, 4294967296.0);
	if (m < 0) m += 4294967296.0;
	return (int32_t)(uint32_t)m;
}
double js_sign(double x) {  // Math.sign keeps ±0 and NaN
	if (x > 0) return 1.0;
	if (x < 0) return -1.0;
	return x;
}
// src/math/mathutils.ts:45-47
bool is_negative(double x) { return x < 0 || (x == 0 && std::signbit(x)); }

// ---------------------------------------------------------------- src/math/vector.ts
// dot :76-84 (sum starts at 0, left to right)
double dot(const V3& a, const V3& b) {
	double s = 0;
	for (int i = 0; i < 3; i++) s += a[i] * b[i];
	return s;
}
V3 add(const V3& a, const V3& b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }     // :104-112
V3 sub(const V3& a, const V3& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }     // :114-122
V3 scale(const V3& a, double s) { return {a[0] * s, a[1] * s, a[2] * s}; }               // :124-131
V3 negate(const V3& a) { return {-a[0], -a[1], -a[2]}; }                                 // :133-140
double length(const V3& a) { return std::sqrt(dot(a, a)); }                              // :142-148
V3 cross(const V3& a, const V3& b) {                                                     // :86-92
	return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
// rotate_vectors :318-323 on 3-vectors
void rotate_vectors(const V3& bx, const V3& by, const V2& r, V3& ox, V3& oy) {
	V3 x = add(scale(bx, r[0]), scale(by, r[1]));
	V3 y = add(scale(bx, -r[1]), scale(by, r[0]));
	ox = x;
	oy = y;
}
// rotate_vectors on 2-vectors, first result only (camera.ts:126-127)
V2 rotate2_first(const V2& bx, const V2& by, const V2& r) {
	return {bx[0] * r[0] + by[0] * r[1], bx[1] * r[0] + by[1] * r[1]};
}
// reflection :263-268
V3 reflection(const V3& v, const V3& n) {
	double norm_scale = -dot(v, n);
	return add(v, scale(n, norm_scale * 2));
}

// ---------------------------------------------------------------- src/math/rng/fp-lcg.ts
struct FpLcg {
	static constexpr double MUL1 = 3532205053565347.0 / 3768278866164713.0;
	static constexpr double TERM1 = 3773467585272041.0 / 4435662911655887.0;
	static constexpr double MUL2 = 3632519696538149.0 / 4496133748415501.0;
	static constexpr double TERM2 = 3396159042346757.0 / 4429161683464229.0;
	static constexpr double MUL3 = 4056279137291581.0 / 4272384783187219.0;
	static constexpr double TERM3 = 3685311960670787.0 / 3909517015383373.0;
	double s1 = 0, s2 = 0, s3 = 0;
	void seed(double seed) {  // :62-66
		s1 = seed;
		s2 = seed * MUL3;
		s3 = seed * MUL2;
	}
	double next() {  // :69-82
		const double a = std::fmod(s1 * MUL1 + TERM1, 1.0);
		const double b = std::fmod(s2 * MUL2 + TERM2, 1.0);
		const double c = std::fmod(s3 * MUL3 + TERM3, 1.0);
		s1 = b + c;
		s2 = c;
		s3 = a + b;
		return std::fmod(a + b + c, 1.0);
	}
};

// src/math/vector_utils.ts:8-14
V3 isotropic_sphere_sample(FpLcg& rng) {
	V3 v;
	do {
		double x = rng.next() * 2 - 1;
		double y = rng.next() * 2 - 1;
		double z = rng.next() * 2 - 1;
		v = {x, y, z};
	} while (dot(v, v) > 1);
	return v;
}

// ---------------------------------------------------------------- src/math/intersection.ts
struct BoxParams {
	int n;  // 0 or 2
	double u1, u2;
	int i1, i2;  // face index, -1 = undefined
};
const V3 FACE_NORMALS[6] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};  // :140-147

// Box.line_intersection :150-204 (pos = centre, size = edge vector)
BoxParams box_line_intersection(const V3& pos, const V3& size, const V3& start, const V3& dir) {
	const V3 tl = sub(pos, scale(size, 0.5));
	const double p[6] = {-dir[0], dir[0], -dir[1], dir[1], -dir[2], dir[2]};
	const double q[6] = {start[0] - tl[0], tl[0] + size[0] - start[0], start[1] - tl[1],
	                     tl[1] + size[1] - start[1], start[2] - tl[2], tl[2] + size[2] - start[2]};
	double u1 = -std::numeric_limits<double>::infinity();
	double u2 = std::numeric_limits<double>::infinity();
	int i1 = -1, i2 = -1;
	for (int i = 0; i < 6; ++i) {
		const double elem = p[i];
		const double u = q[i] / elem;
		if (is_negative(elem)) {
			if (u > u1) { u1 = u; i1 = i; }
		} else {
			if (u < u2) { u2 = u; i2 = i; }
		}
	}
	if (u1 > u2) return {0, u1, u2, i1, i2};
	return {2, u1, u2, i1, i2};
}

struct SphereParams {
	int n;  // 0 or 2
	double t1, t2;
};
// Sphere.line_intersection :109-128 with the cached terms of :88-91
SphereParams sphere_line_intersection(const V3& pos, double radius, const V3& start, const V3& dir) {
	const double dot_pp = dot(pos, pos);
	const double radius_sq = radius * radius;
	const V3 dist = sub(start, pos);
	const double a = dot(dir, dir);
	const double b = dot(dist, dir) * 2;
	const double c = dot(start, start) + dot_pp - dot(start, pos) * 2 - radius_sq;
	const double delta = b * b - a * c * 4;
	if (delta < 0) return {0, 0, 0};
	const double s_delta = std::sqrt(delta);
	const double tmp1 = -b / (a * 2);
	const double tmp2 = s_delta / (a * 2);
	return {2, tmp1 - tmp2, tmp1 + tmp2};
}

// src/math/uv_mapping.ts:19-25
void uv_map_sphere(const V3& d, double& u, double& v) {
	u = std::atan2(d[1], d[0]) / JS_PI / 2.0 + 0.5 - JS_EPSILON;
	double s = 0;  // vector.length(vector.reduce(dir,2))
	s += d[0] * d[0];
	s += d[1] * d[1];
	v = std::atan2(d[2], std::sqrt(s)) / JS_PI + 0.5 - JS_EPSILON;
}

// ---------------------------------------------------------------- scene model
struct Material {  // src/material.ts:67-103
	int response;   // 0 REFLECTION, 1 TRANSMISSION, 2 BOTH
	bool light_source, mirror;
	double roughness;
};
struct Texture {  // src/texture/texture_solid.ts, texture_image.ts
	int kind;     // 0 solid, 1 image
	double color[4];  // solid colour or image fallback colour
	int width = 0, height = 0;
	bool loaded = false;
	std::vector<double> data;  // image_data: u8/255.0, RGB
};
struct Node;
struct Entity {  // src/entity.ts, entities/entity_basic.ts, entity_sphere.ts, entity_box.ts
	int type;  // 0 sphere, 1 box
	V3 pos;
	double extent;  // diameter or size
	int material, texture;
	int substance;  // -1 = undefined
	Node* node = nullptr;
};
struct Node {  // src/octree.ts:25-126 with id = OctreeDim, value = EntitySet
	Node* parent = nullptr;
	Node* child[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	V3 pos;
	double size;
	std::vector<int> set;  // EntitySet: JS Set keeps insertion order (octree_entity.ts:32-49)
	int flat_index = -1;
};
struct NodePos {  // OctreePos, src/octree.ts:129-132
	Node* tree = nullptr;
	int octant = -1;  // -1 == undefined
	bool defined = false;
};

struct Scene {
	std::vector<std::unique_ptr<Node>> pool;
	Node* root = nullptr;       // the tree handed to Raytracer (walker/node_at_pos use tree.get_root())
	std::vector<Entity> entities;
	std::vector<Material> materials;
	std::vector<Texture> textures;
	std::vector<double> substances;  // refractive_index
	std::string error;
	Node* new_node(const V3& pos, double size, Node* parent) {
		pool.emplace_back(new Node());
		Node* n = pool.back().get();
		n->pos = pos;
		n->size = size;
		n->parent = parent;
		return n;
	}
	Node* abs_root() const {  // Octree.get_root :93-102
		Node* n = root;
		while (n->parent) n = n->parent;
		return n;
	}
};

// ---------------------------------------------------------------- src/space.ts
// point_in_space, CLOSE_OPEN :55-66 with a cubic space
bool point_in_cube_close_open(const V3& p, const V3& pos, double size) {
	for (int i = 0; i < 3; i++)
		if (!(p[i] >= pos[i] && p[i] < pos[i] + size)) return false;
	return true;
}
// aabb_in_space -> space_in_space :85-103
bool aabb_in_cube(const V3& apos, double asize, const V3& spos, double ssize) {
	for (int d = 0; d < 3; d++) {
		const double ext_end = spos[d] + ssize;
		const double int_end = apos[d] + asize;
		if (!(apos[d] >= spos[d] && int_end <= ext_end)) return false;
	}
	return true;
}

// ---------------------------------------------------------------- src/octree_space.ts
// octant_adj_pos :41-50
int octant_adj_pos(const Node* n, const V3& pos) {
	const double h = n->size / 2;
	const int px = pos[0] >= n->pos[0] + h;
	const int py = pos[1] >= n->pos[1] + h;
	const int pz = pos[2] >= n->pos[2] + h;
	return (pz << 2) | (py << 1) | px;
}
// node_at_pos :61-93 (start_from_current = false)
NodePos node_at_pos(Node* octree, const V3& pos) {
	const V3 dim_pos = octree->pos;
	const double dim_size = octree->size;
	Node* cur = octree;
	while (cur->parent) cur = cur->parent;  // get_root()
	if (!point_in_cube_close_open(pos, dim_pos, dim_size)) return NodePos{};
	int cur_index = 0;
	V3 next_pos = dim_pos;
	double next_size = dim_size;
	Node* next = cur;
	while (next != nullptr) {
		const V3 rel = sub(pos, next_pos);
		const V3 ind = scale(rel, 2 / next_size);
		cur = next;
		cur_index = (js_int32(ind[2]) << 2) + (js_int32(ind[1]) << 1) + (js_int32(ind[0]) << 0);
		if (cur_index < 0 || cur_index > 7) return NodePos{};  // Octree.get would throw (:47-50)
		next = cur->child[cur_index];
		next_size /= 2;
		for (int i = 0; i < 3; i++) next_pos[i] += js_int32(ind[i]) * next_size;
	}
	NodePos r;
	r.tree = cur;
	r.octant = cur_index;
	r.defined = true;
	return r;
}
// index_within_parent :113-125 (the cached field is never assigned in the reference)
int index_within_parent(const Node* child) {
	const Node* parent = child->parent;
	if (!parent) return -1;
	const V3 rel = sub(child->pos, parent->pos);
	const V3 ind = scale(rel, 2 / parent->size);
	return (js_int32(ind[2]) << 2) + (js_int32(ind[1]) << 1) + (js_int32(ind[0]) << 0);
}
// new_subtree :95-108
Node* new_subtree(Scene& s, Node* tree, int n) {
	const double half = tree->size / 2;
	const V3 bits = {(double)(n & 1), (double)((n >> 1) & 1), (double)((n >> 2) & 1)};
	Node* sub_ = s.new_node(add(tree->pos, scale(bits, half)), half, tree);
	tree->child[n] = sub_;
	return sub_;
}

struct WalkerCounters {
	uint64_t cell_steps = 0;   // update_next_pos calls (one slab test each)
	uint64_t would_throw = 0;  // update_next_pos with an empty parameter list (reference: TypeError)
};

// OctreeWalker :159-408
struct Walker {
	Node* tree;
	bool include_undefined = false;
	bool has_pos = false;
	V3 pos{}, direction{};
	NodePos cur_node;
	V3 next_point{};
	V3 next_normal{};
	bool cur_returned = false, stepped_in = false, next_pos_is_ahead = false;
	int depth = 0;
	WalkerCounters* ctr = nullptr;

	explicit Walker(Node* t) : tree(t) {}

	void reset_state() {  // :240-246
		next_point = pos;
		next_normal = {0, 0, 0};
		cur_returned = false;
		stepped_in = false;
		next_pos_is_ahead = false;
		depth = 0;
	}
	// set_pos_and_dir :223-226 -> set_direction, set_position :188-205
	bool set_pos_and_dir(const V3& p, const V3& d, const NodePos* node = nullptr) {
		direction = d;
		if (node != nullptr && node->defined)
			cur_node = *node;
		else
			cur_node = node_at_pos(tree, p);
		pos = p;
		has_pos = true;
		return setup_cur_node();
	}
	bool setup_cur_node() {  // :251-278
		reset_state();
		if (cur_node.defined) return true;
		const V3 c = add(tree->pos, scale(V3{0.5, 0.5, 0.5}, tree->size));
		const V3 sz = scale(V3{1, 1, 1}, tree->size);
		const BoxParams bp = box_line_intersection(c, sz, pos, direction);
		// select_parameters FORWARD :207-216
		double param;
		int face;
		if (bp.n == 0) return false;
		if (bp.u1 >= 0) { param = bp.u1; face = bp.i1; }
		else if (bp.u2 >= 0) { param = bp.u2; face = bp.i2; }
		else return false;
		cur_node.tree = tree;
		cur_node.octant = -1;
		cur_node.defined = true;
		next_point = add(pos, scale(direction, param));
		next_normal = face >= 0 ? negate(FACE_NORMALS[face]) : V3{0, 0, 0};
		return true;
	}
	void step_back() {  // :280-308
		stepped_in = true;
		if (cur_node.octant < 0) {
			cur_node = NodePos{};
			cur_returned = false;
			return;
		}
		if (depth > 0) {
			depth--;
			cur_returned = true;
		} else {
			cur_returned = false;
		}
		const int gi = index_within_parent(cur_node.tree);
		if (gi >= 0) {
			cur_node.octant = gi;
			cur_node.tree = cur_node.tree->parent;
		} else {
			cur_node.octant = -1;
		}
	}
	void update_next_pos() {  // :369-384 via dim_relative_to_parent :127-136
		const Node* t = cur_node.tree;
		const int n = cur_node.octant;
		const double ph = t->size / 2;
		const V3 bits = {(double)((n >> 0) & 1), (double)((n >> 1) & 1), (double)((n >> 2) & 1)};
		const V3 dpos = add(t->pos, scale(bits, ph));
		const V3 c = add(dpos, scale(V3{0.5, 0.5, 0.5}, ph));
		const V3 sz = scale(V3{1, 1, 1}, ph);
		const BoxParams bp = box_line_intersection(c, sz, pos, direction);
		if (ctr) {
			ctr->cell_steps++;
			if (bp.n == 0) ctr->would_throw++;  // inter_param.pop() === undefined in the reference
		}
		next_point = add(pos, scale(direction, bp.u2));
		next_normal = bp.i2 >= 0 ? FACE_NORMALS[bp.i2] : V3{0, 0, 0};
	}
	// next :316-361.  Returns false when exhausted; out_node may be null with include_undefined.
	bool next(Node*& out_node, NodePos& out_pos) {
		while (cur_node.defined) {
			const NodePos last = cur_node;
			Node* last_node = last.octant >= 0 ? last.tree->child[last.octant] : last.tree;
			if (!cur_returned) {
				if (include_undefined || last_node != nullptr) {
					cur_returned = true;
					out_node = last_node;
					out_pos = last;
					return true;
				}
			}
			if (last.octant >= 0) {
				if (!next_pos_is_ahead) {
					if (!stepped_in && last_node != nullptr) {
						const int n_octant = octant_adj_pos(last_node, next_point);
						depth++;  // step_in :310-314
						cur_node.tree = last_node;
						cur_node.octant = n_octant;
						cur_returned = false;
						continue;
					}
					update_next_pos();
				}
				const double nx = ((last.octant >> 0) & 1) + next_normal[0];
				const double ny = ((last.octant >> 1) & 1) + next_normal[1];
				const double nz = ((last.octant >> 2) & 1) + next_normal[2];
				const bool out = nx < 0 || nx > 1 || ny < 0 || ny > 1 || nz < 0 || nz > 1;
				if (!out) {
					cur_node.tree = last.tree;
					cur_node.octant = (int)nx | ((int)ny << 1) | ((int)nz << 2);
					cur_returned = false;
					stepped_in = false;
					next_pos_is_ahead = false;
					continue;
				} else {
					next_pos_is_ahead = true;
				}
			}
			step_back();
		}
		return false;
	}
};

// ---------------------------------------------------------------- src/octree_entity.ts
// Entity.get_aabb: entity_sphere.ts:90-96, entity_box.ts:75-82
void entity_aabb(const Entity& e, V3& pos, double& size) {
	if (e.type == 0) {
		const double d = e.extent;
		pos = sub(e.pos, scale(V3{d, d, d}, 0.5));
		size = d;
	} else {
		const double h = e.extent / 2;
		pos = sub(e.pos, V3{h, h, h});
		size = e.extent;
	}
}
int relative_level(const Node* n, const Node* root) {  // octree.ts:105-118
	auto level = [](const Node* x) { int l = 0; while ((x = x->parent)) l++; return l; };
	return level(n) - level(root);
}
// get_covering_node_for_entity :60-79
Node* get_covering_node(Scene& s, const V3& apos, double asize) {
	NodePos deepest = node_at_pos(s.root, apos);
	if (!deepest.defined) return nullptr;
	Node* cur = deepest.tree;
	do {
		if (aabb_in_cube(apos, asize, cur->pos, cur->size)) break;
		cur = cur->parent;
	} while (cur != nullptr);
	return cur;
}
// extend_tree_inside_to_fit_up_to_depth :92-114
Node* extend_inside(Scene& s, Node* root, Node* node, const V3& apos, double asize, int max_depth) {
	int cur_depth = relative_level(node, root);
	Node* cur = node;
	while (cur_depth < max_depth) {
		const V3 q = scale(sub(apos, cur->pos), 2.0 / cur->size);
		const int x = js_int32(q[0]), y = js_int32(q[1]), z = js_int32(q[2]);
		const V3 spos = add(cur->pos, scale(V3{(double)x, (double)y, (double)z}, cur->size / 2));
		const double ssize = cur->size / 2;
		if (!aabb_in_cube(apos, asize, spos, ssize)) break;
		const int idx = (z << 2) | (y << 1) | (x << 0);
		if (idx < 0 || idx > 7) break;  // Octree.set would throw
		Node* nt = s.new_node(spos, cur->size / 2, cur);
		cur->child[idx] = nt;
		cur = nt;
		cur_depth++;
	}
	return cur;
}
// extend_tree_outside_to_fit_up_to_depth :125-171.  Returns nullptr on TreeOutsideGrowError.
Node* extend_outside(Scene& s, Node* root, Node* node, const V3& apos, double asize, int max_depth) {
	int cur_depth = relative_level(root, node);
	Node* cur = node;
	bool fit = false;
	while (cur_depth < max_depth) {
		V3 al = scale(sub(apos, cur->pos), 1.0 / cur->size);
		for (int d = 0; d < 3; d++) al[d] = std::max(std::min(std::floor(al[d]), 0.0), -1.0);
		const V3 ppos = add(cur->pos, scale(al, cur->size));
		const double psize = cur->size * 2;
		const int idx = (js_int32(-al[2]) << 2) | (js_int32(-al[1]) << 1) | (js_int32(-al[0]) << 0);
		Node* np = s.new_node(ppos, psize, nullptr);
		np->child[idx] = cur;
		cur->parent = np;
		cur = np;
		if (aabb_in_cube(apos, asize, ppos, psize)) { fit = true; break; }
		cur_depth++;
	}
	if (!fit) return nullptr;
	return cur;
}
// add_entity_to_octree :174-188 + Entity.set_octree entity.ts:50-56
int add_entity(Scene& s, Entity e, int max_in_depth, int max_out_depth) {
	V3 apos;
	double asize;
	entity_aabb(e, apos, asize);
	Node* fitting = get_covering_node(s, apos, asize);
	if (fitting == nullptr) {
		fitting = extend_outside(s, s.root, s.abs_root(), apos, asize, max_out_depth);
		if (fitting == nullptr) {
			s.error = "TreeOutsideGrowError: The tree outside-depth limit exceeded";
			return -1;
		}
	}
	fitting = extend_inside(s, s.root, fitting, apos, asize, max_in_depth);
	const int id = (int)s.entities.size();
	e.node = fitting;
	s.entities.push_back(e);
	fitting->set.push_back(id);
	return id;
}

// Entity.is_within: entity_sphere.ts:63-66 (_radius_sq = d*d/4, :38), entity_box.ts:47-52
bool entity_is_within(const Entity& e, const V3& p) {
	if (e.type == 0) {
		const V3 dist = sub(p, e.pos);
		return dot(dist, dist) <= e.extent * e.extent / 4;
	}
	return point_in_cube_close_open(p, e.pos, e.extent);
}
// entity_at_pos :191-202
int entity_at_pos(const Scene& s, const V3& p, uint64_t* tests = nullptr) {
	NodePos np = node_at_pos(s.root, p);
	Node* cur = np.defined ? np.tree : nullptr;
	while (cur != nullptr) {
		for (int id : cur->set) {
			if (tests) (*tests)++;
			if (entity_is_within(s.entities[id], p)) return id;
		}
		cur = cur->parent;
	}
	return -1;
}

// ---------------------------------------------------------------- entities: collision_info
struct Collision {
	bool hit = false;
	V3 point{}, normal{};
};
// SphereEntity.collision_info entity_sphere.ts:68-88
Collision sphere_collision(const Entity& e, const V3& raypos, const V3& raydir) {
	Collision c;
	const SphereParams sp = sphere_line_intersection(e.pos, e.extent / 2, raypos, raydir);
	if (sp.n == 0) return c;
	double t;
	if (sp.t1 >= 0) t = sp.t1;
	else if (sp.t2 >= 0) t = sp.t2;
	else return c;
	c.hit = true;
	c.point = add(raypos, scale(raydir, t));
	c.normal = scale(sub(c.point, e.pos), 2 / e.extent);
	c.normal = scale(c.normal, -js_sign(dot(raydir, c.normal)));
	return c;
}
// BoxEntity.collision_info entity_box.ts:54-73
Collision box_collision(const Entity& e, const V3& raypos, const V3& raydir) {
	Collision c;
	const V3 sz = scale(V3{1, 1, 1}, e.extent);
	const BoxParams bp = box_line_intersection(e.pos, sz, raypos, raydir);
	if (bp.n == 0) return c;
	double t;
	int face;
	if (bp.u1 >= 0) { t = bp.u1; face = bp.i1; }
	else if (bp.u2 >= 0) { t = bp.u2; face = bp.i2; }
	else return c;
	c.hit = true;
	c.point = add(raypos, scale(raydir, t));
	const V3 n = face >= 0 ? FACE_NORMALS[face] : V3{0, 0, 0};
	c.normal = scale(n, -js_sign(dot(raydir, n)));
	return c;
}

// ---------------------------------------------------------------- textures, sky
// SolidTexture.get_color texture_solid.ts:33-35; ImageTexture.get_color texture_image.ts:40-63
bool texture_color(const Texture& t, double u, double v, double out[3]) {
	if (t.kind == 0 || !t.loaded) {
		out[0] = t.color[0]; out[1] = t.color[1]; out[2] = t.color[2];
		return true;
	}
	if (u < 0 - JS_EPSILON || u > 1 - JS_EPSILON || v < 0 - JS_EPSILON || v > 1 - JS_EPSILON)
		return false;  // Error('Texture coordinates out of bounds')
	const int32_t ui = js_int32(u * t.width);
	const int32_t vi = js_int32(v * t.height);
	const int64_t px = ((int64_t)vi * t.width + ui) * 3;
	if (px < 0 || px + 2 >= (int64_t)t.data.size()) return false;
	out[0] = t.data[px]; out[1] = t.data[px + 1]; out[2] = t.data[px + 2];
	return true;
}

struct Config {  // RaytracerConfig raytracer.ts:33-43 (+ SkySphere sky_sphere.ts:22-27)
	int refmax;
	int sky_texture;
	int default_substance;
	double distance_attenuation_factor;
};

struct PixelCounters {
	uint32_t nodes = 0, tests = 0, shades = 0, segments = 0;
};

struct Totals {
	uint64_t nodes = 0, tests = 0, shades = 0, segments = 0, cell_steps = 0, would_throw = 0,
	         within_tests = 0, texture_errors = 0, acute_warnings = 0, paths = 0;
};

// ---------------------------------------------------------------- Ray.trace raytracer.ts:168-277
struct RayResult {
	double color[3];
	int first_entity;  // entity of the first non-null collision_info of the path, -1 = none
};

// ORC_DEBUG_PATHS=1 in the environment: every collision of every traced path is printed (debugging single pixels)
static const bool g_debug_paths = getenv("ORC_DEBUG_PATHS") != nullptr;

RayResult trace_ray(const Scene& s, const Config& cfg, Walker& walker, FpLcg& rng, const V3& start_point,
                    const NodePos& start_node, const V3& dir_in, int start_substance, PixelCounters& pc,
                    Totals& tot) {
	// Ray constructor :79-99, keep_dir_unnormalized = true, colour = COLOR_WHITE
	int refcount = 0;
	const int refmax = cfg.refmax;
	V3 refpoint = start_point;
	V3 dir = dir_in;
	double col[3] = {1, 1, 1};
	double path_distance = 0;
	int cur_substance = start_substance;
	RayResult res;
	res.first_entity = -1;
	auto finish = [&]() {
		res.color[0] = col[0]; res.color[1] = col[1]; res.color[2] = col[2];
		return res;
	};

	walker.set_pos_and_dir(refpoint, dir, &start_node);  // :171
	pc.segments++;
	bool light_hit = false;
	Node* tree_node;
	NodePos tree_pos;
	while (walker.next(tree_node, tree_pos)) {  // :179
		pc.nodes++;
		Collision ci;
		int entity = -1;
		for (int id : tree_node->set) {  // :186-195, Set iteration order == insertion order
			entity = id;
			pc.tests++;
			const Entity& e = s.entities[id];
			ci = e.type == 0 ? sphere_collision(e, refpoint, dir) : box_collision(e, refpoint, dir);
			if (ci.hit) break;
		}
		if (!ci.hit) continue;
		if (g_debug_paths)
			fprintf(stderr, "[oracle] hit %d: entity %d from (%.17g %.17g %.17g) dir (%.17g %.17g %.17g) at (%.17g %.17g %.17g)\n", refcount,
			        entity, refpoint[0], refpoint[1], refpoint[2], dir[0], dir[1], dir[2], ci.point[0], ci.point[1], ci.point[2]);
		if (res.first_entity < 0) res.first_entity = entity;
		if (dot(dir, ci.normal) >= 0) {  // :200-203
			tot.acute_warnings++;
			return finish();
		}
		const Entity& e = s.entities[entity];
		const Material& m = s.materials[e.material];
		refcount++;  // :208
		{  // SolidMaterial.alter_ray material_solid.ts:30-36
			double u = 0, v = 0;
			if (e.type == 0) uv_map_sphere(sub(ci.point, e.pos), u, v);  // entity_sphere.ts:98-101
			// BoxEntity.map_uv returns [0,0] (entity_box.ts:104-107)
			double tc[3];
			if (!texture_color(s.textures[e.texture], u, v, tc)) {
				tot.texture_errors++;
				tc[0] = tc[1] = tc[2] = std::numeric_limits<double>::quiet_NaN();
			}
			col[0] *= tc[0]; col[1] *= tc[1]; col[2] *= tc[2];  // mul_color color.ts:50-52
			pc.shades++;
		}
		path_distance += length(sub(ci.point, refpoint));  // :210
		refpoint = ci.point;                                // :212
		if (m.light_source) {  // :215-218
			light_hit = true;
			break;
		}
		if (m.response == 0) {  // REFLECTION :221-237
			if (!m.mirror) return finish();
			dir = reflection(dir, ci.normal);  // reflect_ray :117-119
			if (m.roughness > 0.0) {           // scatter_ray :121-133
				V3 rv = isotropic_sphere_sample(rng);
				if (dot(rv, ci.normal) < 0) rv = scale(rv, -1);
				V3 ref_vec = add(scale(dir, 1 - m.roughness), scale(rv, m.roughness));
				dir = scale(ref_vec, 1.0 / length(ref_vec));
			}
			refpoint = add(refpoint, scale(dir, 1e-3));  // move_slightly_forward :158-164
		} else if (m.response == 1) {  // TRANSMISSION :238-249
			refpoint = add(refpoint, scale(dir, 1e-3));
			uint64_t wt = 0;
			const int rf = entity_at_pos(s, refpoint, &wt);
			tot.within_tests += wt;
			const int substance = rf >= 0 ? s.entities[rf].substance : cfg.default_substance;
			if (substance >= 0) {  // refract_ray :135-150
				const double r_ratio = s.substances[cur_substance] / s.substances[substance];
				const double r_ratio_sq = r_ratio * r_ratio;
				const double cosine = dot(dir, ci.normal);
				const double cosine_sq = cosine * cosine;
				const double ref_sine_sq = (1 - cosine_sq) * r_ratio_sq;
				if (ref_sine_sq <= 1) {
					const double ref_cosine = std::sqrt(1 - ref_sine_sq);
					const V3 adj = scale(ci.normal, ref_cosine - cosine);
					dir = scale(dir, r_ratio);
					dir = sub(dir, adj);
				} else {
					dir = reflection(dir, ci.normal);
				}
				cur_substance = substance;
			}
		} else {  // default :250-251
			return finish();
		}
		walker.set_pos_and_dir(refpoint, dir);  // :254
		if (refcount >= refmax) {               // :256-263
			col[0] = col[1] = col[2] = 0;
			return finish();
		}
		pc.segments++;
	}
	if (!light_hit) {  // :267-271, SkySphere.get_color sky_sphere.ts:22-27
		double u, v, sc[3];
		uv_map_sphere(dir, u, v);
		if (!texture_color(s.textures[cfg.sky_texture], u, v, sc)) {
			tot.texture_errors++;
			sc[0] = sc[1] = sc[2] = std::numeric_limits<double>::quiet_NaN();
		}
		col[0] *= sc[0]; col[1] *= sc[1]; col[2] *= sc[2];
		return finish();
	}
	// :273-276 inverse square law
	const double t = path_distance * cfg.distance_attenuation_factor;
	const double isl = 1.0 / (JS_EPSILON + t * t);
	col[0] *= isl; col[1] *= isl; col[2] *= isl;
	return finish();
}

// ---------------------------------------------------------------- src/view/camera.ts
struct Camera {
	double fov_v, fov_h;
	int screen_w, screen_h;
	double rot_v, rot_h;
	bool vertical_locked;
	V3 pos;
	V3 norm_fr{1, 0, 0}, norm_lf{0, 1, 0}, norm_up{0, 0, 1};
	V2 rot_v_v, rot_h_v, rot_scan_h_v, rot_scan_v_v;

	void init_rot_vectors() {  // :77-87
		rot_h_v = {std::cos(rot_h), std::sin(rot_h)};
		rot_v_v = {std::cos(rot_v), std::sin(rot_v)};
		const double rad_h = fov_h / screen_w;
		const double rad_v = fov_v / screen_h;
		rot_scan_h_v = {std::cos(rad_h), std::sin(rad_h)};
		rot_scan_v_v = {std::cos(rad_v), std::sin(rad_v)};
	}
	bool rotate_h_v(const V2& v) {  // :123-133
		V2 fr_xy = {norm_fr[0], norm_fr[1]};
		V2 lf_xy = {norm_lf[0], norm_lf[1]};
		fr_xy = rotate2_first(fr_xy, V2{-fr_xy[1], fr_xy[0]}, v);
		lf_xy = rotate2_first(lf_xy, V2{-lf_xy[1], lf_xy[0]}, v);
		norm_fr = {fr_xy[0], fr_xy[1], norm_fr[2]};
		norm_lf = {lf_xy[0], lf_xy[1], norm_lf[2]};
		norm_up = cross(norm_fr, norm_lf);
		return true;
	}
	bool rotate_v_v(const V2& v) {  // :137-149
		const double cmp_sign = v[1] < 0 ? 1 : -1;
		V3 fr, up;
		rotate_vectors(norm_fr, norm_up, v, fr, up);
		if (vertical_locked && js_sign(fr[2] - norm_fr[2]) == cmp_sign) return false;
		norm_fr = fr;
		norm_up = up;
		return true;
	}
	void rotate_h(double a) { rotate_h_v({std::cos(a), std::sin(a)}); }  // :90-93
	void rotate_v(double a) { rotate_v_v({std::cos(a), std::sin(a)}); }  // :96-99
	void rotate_h_step(int n) {  // :102-107 (uses rot_v_v, as the reference does)
		const V2 r = n >= 0 ? rot_v_v : V2{rot_v_v[0], -rot_v_v[1]};
		for (int i = 0; i < std::abs(n); i++) rotate_h_v(r);
	}
	void rotate_v_step(int n) {  // :110-118 (uses rot_h_v, as the reference does)
		const V2 r = n >= 0 ? rot_h_v : V2{rot_h_v[0], -rot_h_v[1]};
		for (int i = 0; i < std::abs(n); i++)
			if (!rotate_v_v(r)) break;
	}

	// get_dir_for_each_pixel :207-250.  `fixed_extents` = false restates the reference
	// literally (x over [0,screen_h), y over [0,screen_w): identical on squares);
	// true uses the evident intent (x over [0,screen_w), y over [0,screen_h)) so that
	// non-square frames can be timed.
	template <class F>
	void for_each_pixel(bool fixed_extents, F&& yield) const {
		const int x_extent = fixed_extents ? screen_w : screen_h;
		const int y_extent = fixed_extents ? screen_h : screen_w;
		const V2 rot_scan_h_counter = {rot_scan_h_v[0], -rot_scan_h_v[1]};
		const V2 rot_scan_v_counter = {rot_scan_v_v[0], -rot_scan_v_v[1]};
		auto iter_h = [&](int from_x, int to_x, int y, const V2& rot, const V3& beg_fr, int inc, bool first) {
			V3 fr = beg_fr, lf = norm_lf;
			if (first) rotate_vectors(fr, lf, rot, fr, lf);
			for (int i = from_x; i != to_x; i += inc) {
				yield(i, y, fr);
				rotate_vectors(fr, lf, rot, fr, lf);
			}
		};
		auto iter_v = [&](int from_y, int to_y, const V2& rot, int inc, bool first) {
			V3 fr = norm_fr, up = norm_up;
			if (first) rotate_vectors(fr, up, rot, fr, up);
			for (int i = from_y; i != to_y; i += inc) {
				iter_h(x_extent >> 1, x_extent, i, rot_scan_h_v, fr, 1, false);
				iter_h((x_extent >> 1) - 1, -1, i, rot_scan_h_counter, fr, -1, true);
				rotate_vectors(fr, up, rot, fr, up);
			}
		};
		iter_v(y_extent >> 1, y_extent, rot_scan_v_v, 1, false);
		iter_v((y_extent >> 1) - 1, -1, rot_scan_v_counter, -1, true);
	}
};

struct CamPixel {
	int x, y;
	V3 dir;
};

}  // namespace

// =============================================================================
// C ABI for the python test harness (ctypes).  Handles are opaque pointers.
// =============================================================================
extern "C" {

struct orc_config {
	int32_t refmax;
	int32_t sky_texture;
	int32_t default_substance;
	double distance_attenuation_factor;
};

struct orc_totals {
	uint64_t nodes, tests, shades, segments, cell_steps, would_throw, within_tests, texture_errors,
	    acute_warnings, paths;
};

void* orc_scene_new(const double* root_pos, double root_size) {
	Scene* s = new Scene();
	s->root = s->new_node({root_pos[0], root_pos[1], root_pos[2]}, root_size, nullptr);
	return s;
}
void orc_scene_free(void* h) { delete (Scene*)h; }
const char* orc_scene_error(void* h) { return ((Scene*)h)->error.c_str(); }

int32_t orc_add_material(void* h, int32_t response, int32_t light, int32_t mirror, double roughness) {
	Scene* s = (Scene*)h;
	s->materials.push_back({response, light != 0, mirror != 0, roughness});
	return (int32_t)s->materials.size() - 1;
}
int32_t orc_add_substance(void* h, double refractive_index) {
	Scene* s = (Scene*)h;
	s->substances.push_back(refractive_index);
	return (int32_t)s->substances.size() - 1;
}
int32_t orc_add_texture_solid(void* h, double r, double g, double b, double a) {
	Scene* s = (Scene*)h;
	Texture t;
	t.kind = 0;
	t.color[0] = r; t.color[1] = g; t.color[2] = b; t.color[3] = a;
	s->textures.push_back(std::move(t));
	return (int32_t)s->textures.size() - 1;
}
// rgb8 = w*h*3 bytes as decoded by load_image (texture_image.ts:112-124: data = u8/255.0); rgb8 == NULL
// models a texture that is not (yet) loaded, which answers with the fallback colour (:41-44).
int32_t orc_add_texture_image(void* h, int32_t w, int32_t hgt, const uint8_t* rgb8, const double* fallback) {
	Scene* s = (Scene*)h;
	Texture t;
	t.kind = 1;
	for (int i = 0; i < 4; i++) t.color[i] = fallback[i];
	t.width = w;
	t.height = hgt;
	if (rgb8) {
		t.loaded = true;
		t.data.resize((size_t)w * hgt * 3);
		for (size_t i = 0; i < t.data.size(); i++) t.data[i] = rgb8[i] / 255.0;
	}
	s->textures.push_back(std::move(t));
	return (int32_t)s->textures.size() - 1;
}
// add_entity_to_octree.  type 0 sphere (extent = diameter), 1 box (extent = size).
// Returns the entity id (insertion index) or -1 on TreeOutsideGrowError.
int32_t orc_add_entity(void* h, int32_t type, const double* pos, double extent, int32_t material,
                       int32_t texture, int32_t substance, int32_t max_in_depth, int32_t max_out_depth) {
	Scene* s = (Scene*)h;
	Entity e;
	e.type = type;
	e.pos = {pos[0], pos[1], pos[2]};
	e.extent = extent;
	e.material = material;
	e.texture = texture;
	e.substance = substance;
	return add_entity(*s, e, max_in_depth, max_out_depth);
}
// Moving an entity as the reference does it: BasicEntity._set_pos (src/entities/entity_basic.ts:38-42), then
// add_entity_to_octree again (src/octree_entity.ts:174-188), whose Entity.set_octree (src/entity.ts:50-56) deletes the
// entity from the Set of the node it was in and adds it to the END of the new node's Set.  0 / -1 (TreeOutsideGrowError).
int32_t orc_move_entity(void* h, int32_t id, const double* pos, int32_t max_in_depth, int32_t max_out_depth) {
	Scene& s = *(Scene*)h;
	if (id < 0 || id >= (int)s.entities.size()) return -2;
	Entity& e = s.entities[id];
	e.pos = {pos[0], pos[1], pos[2]};
	V3 apos;
	double asize;
	entity_aabb(e, apos, asize);
	Node* fitting = get_covering_node(s, apos, asize);
	if (fitting == nullptr) {
		fitting = extend_outside(s, s.root, s.abs_root(), apos, asize, max_out_depth);
		if (fitting == nullptr) {
			s.error = "TreeOutsideGrowError: The tree outside-depth limit exceeded";
			return -1;
		}
	}
	fitting = extend_inside(s, s.root, fitting, apos, asize, max_in_depth);
	if (e.node) {  // set_octree: this._octree.value.set.delete(this)
		std::vector<int>& old = e.node->set;
		old.erase(std::find(old.begin(), old.end(), id));
	}
	e.node = fitting;
	fitting->set.push_back(id);  // tree.value.set.add(this): insertion order, i.e. last
	return 0;
}
// bulk variant of orc_add_entity (same order, same semantics); stops at the first error.
int32_t orc_add_entities(void* h, int32_t n, const uint8_t* type, const double* pos, const double* extent,
                         const int32_t* material, const int32_t* texture, const int32_t* substance,
                         int32_t max_in_depth, int32_t max_out_depth) {
	for (int i = 0; i < n; i++) {
		const int32_t r = orc_add_entity(h, type[i], pos + 3 * (size_t)i, extent[i], material[i], texture[i],
		                                 substance[i], max_in_depth, max_out_depth);
		if (r < 0) return -1 - i;
	}
	return n;
}
// new_subtree octree_space.ts:95-108, addressed by a path of octants from the root. Returns 0 / -1.
int32_t orc_new_subtree(void* h, const int32_t* path, int32_t path_len, int32_t n) {
	Scene* s = (Scene*)h;
	Node* cur = s->abs_root();
	for (int i = 0; i < path_len; i++) {
		if (path[i] < 0 || path[i] > 7 || !cur->child[path[i]]) return -1;
		cur = cur->child[path[i]];
	}
	if (n < 0 || n > 7) return -2;          // Octree.check_bounds throws (octree.ts:41-44)
	if (cur->child[n] != nullptr) return -3;  // "Child already defined"
	new_subtree(*s, cur, n);
	return 0;
}

// ---- flat export: nodes in DFS pre-order (children 0..7), root = 0 ----------
static void flat_number(Node* n, std::vector<Node*>& order) {
	n->flat_index = (int)order.size();
	order.push_back(n);
	for (int i = 0; i < 8; i++)
		if (n->child[i]) flat_number(n->child[i], order);
}
void orc_flat_counts(void* h, uint32_t* n_nodes, uint32_t* n_list) {
	Scene* s = (Scene*)h;
	std::vector<Node*> order;
	flat_number(s->abs_root(), order);
	uint32_t nl = 0;
	for (Node* n : order) nl += (uint32_t)n->set.size();
	*n_nodes = (uint32_t)order.size();
	*n_list = nl;
}
void orc_flat_export(void* h, double* node_pos, double* node_size, int32_t* node_child, int32_t* node_parent,
                     int32_t* node_octant, uint32_t* node_list_off, uint32_t* list_entity, int32_t* root_index) {
	Scene* s = (Scene*)h;
	std::vector<Node*> order;
	flat_number(s->abs_root(), order);
	uint32_t off = 0;
	for (size_t i = 0; i < order.size(); i++) {
		Node* n = order[i];
		for (int d = 0; d < 3; d++) node_pos[3 * i + d] = n->pos[d];
		node_size[i] = n->size;
		for (int c = 0; c < 8; c++) node_child[8 * i + c] = n->child[c] ? n->child[c]->flat_index : -1;
		node_parent[i] = n->parent ? n->parent->flat_index : -1;
		node_octant[i] = index_within_parent(n);
		node_list_off[i] = off;
		for (int id : n->set) list_entity[off++] = (uint32_t)id;
	}
	node_list_off[order.size()] = off;
	*root_index = s->root->flat_index;
}
int32_t orc_entity_node(void* h, int32_t id) {
	Scene* s = (Scene*)h;
	std::vector<Node*> order;
	flat_number(s->abs_root(), order);
	return s->entities[id].node->flat_index;
}

// ---- unit-level probes for the golden tests ---------------------------------
// node_at_pos -> (flat node index, octant) or (-1,-1) for null
void orc_node_at_pos(void* h, const double* p, int32_t* node, int32_t* octant) {
	Scene* s = (Scene*)h;
	std::vector<Node*> order;
	flat_number(s->abs_root(), order);
	NodePos np = node_at_pos(s->root, {p[0], p[1], p[2]});
	*node = np.defined ? np.tree->flat_index : -1;
	*octant = np.defined ? np.octant : -1;
}
int32_t orc_entity_at_pos(void* h, const double* p) { return entity_at_pos(*(Scene*)h, {p[0], p[1], p[2]}); }
// OctreeWalker.each_stop(): for each stop writes (pos.tree flat index or -1 for the root stop,
// pos.octant or -1, node flat index or -1 when undefined).  Returns the number of stops.
int32_t orc_walk(void* h, const double* p, const double* d, int32_t include_undefined, int32_t use_start_node,
                 int32_t* stops, int32_t max_stops) {
	Scene* s = (Scene*)h;
	std::vector<Node*> order;
	flat_number(s->abs_root(), order);
	Walker w(s->root);
	w.include_undefined = include_undefined != 0;
	const V3 pos = {p[0], p[1], p[2]}, dir = {d[0], d[1], d[2]};
	if (use_start_node) {
		NodePos sn = node_at_pos(s->root, pos);
		w.set_pos_and_dir(pos, dir, &sn);
	} else {
		w.set_pos_and_dir(pos, dir);
	}
	int n = 0;
	Node* node;
	NodePos np;
	while (w.next(node, np)) {
		if (n < max_stops) {
			stops[3 * n + 0] = np.octant >= 0 ? np.tree->flat_index : -1;
			stops[3 * n + 1] = np.octant;
			stops[3 * n + 2] = node ? node->flat_index : -1;
		}
		n++;
		if (n > (1 << 24)) break;
	}
	return n;
}
// Entity.collision_info for one entity: returns 1 on hit and writes point[3], normal[3]
int32_t orc_collision(void* h, int32_t id, const double* p, const double* d, double* point, double* normal) {
	Scene* s = (Scene*)h;
	const Entity& e = s->entities[id];
	const V3 pos = {p[0], p[1], p[2]}, dir = {d[0], d[1], d[2]};
	Collision c = e.type == 0 ? sphere_collision(e, pos, dir) : box_collision(e, pos, dir);
	if (!c.hit) return 0;
	for (int i = 0; i < 3; i++) { point[i] = c.point[i]; normal[i] = c.normal[i]; }
	return 1;
}
void orc_uv_map_sphere(const double* d, double* uv) { uv_map_sphere({d[0], d[1], d[2]}, uv[0], uv[1]); }
void orc_fplcg(double seed, int32_t n, double* out) {
	FpLcg r;
	r.seed(seed);
	for (int i = 0; i < n; i++) out[i] = r.next();
}
// Box.line_intersection probe: returns n (0/2), writes u1,u2 and face indices
int32_t orc_box_line(const double* center, const double* size, const double* p, const double* d, double* u,
                     int32_t* faces) {
	BoxParams b = box_line_intersection({center[0], center[1], center[2]}, {size[0], size[1], size[2]},
	                                    {p[0], p[1], p[2]}, {d[0], d[1], d[2]});
	u[0] = b.u1; u[1] = b.u2; faces[0] = b.i1; faces[1] = b.i2;
	return b.n;
}

// ---- camera -------------------------------------------------------------------
void* orc_camera_new(double fov_v, double fov_h, int32_t screen_w, int32_t screen_h, double rot_v, double rot_h,
                     int32_t vertical_locked, const double* init_pos, int32_t has_v, double init_v_angle,
                     int32_t has_h, double init_h_angle) {
	Camera* c = new Camera();  // constructor camera.ts:61-75
	c->fov_v = fov_v; c->fov_h = fov_h; c->screen_w = screen_w; c->screen_h = screen_h;
	c->rot_v = rot_v; c->rot_h = rot_h; c->vertical_locked = vertical_locked != 0;
	c->pos = {init_pos[0], init_pos[1], init_pos[2]};
	c->init_rot_vectors();
	if (has_h) c->rotate_h(init_h_angle);
	if (has_v) c->rotate_v(init_v_angle);
	return c;
}
void orc_camera_free(void* c) { delete (Camera*)c; }
void orc_camera_rotate_h(void* c, double a) { ((Camera*)c)->rotate_h(a); }
void orc_camera_rotate_v(void* c, double a) { ((Camera*)c)->rotate_v(a); }
void orc_camera_rotate_h_step(void* c, int32_t n) { ((Camera*)c)->rotate_h_step(n); }
void orc_camera_rotate_v_step(void* c, int32_t n) { ((Camera*)c)->rotate_v_step(n); }
void orc_camera_set_pos(void* c, const double* p) { ((Camera*)c)->pos = {p[0], p[1], p[2]}; }
// pos[3], fr[3], lf[3], up[3]
void orc_camera_get_basis(void* c, double* out) {
	Camera* k = (Camera*)c;
	for (int i = 0; i < 3; i++) {
		out[i] = k->pos[i]; out[3 + i] = k->norm_fr[i]; out[6 + i] = k->norm_lf[i]; out[9 + i] = k->norm_up[i];
	}
}
// get_dir_for_each_pixel in yield order: xy[2*i], dir[3*i].  Returns the number of pixels yielded.
int64_t orc_camera_dirs(void* c, int32_t fixed_extents, int32_t* xy, double* dir, int64_t max_pixels) {
	int64_t n = 0;
	((Camera*)c)->for_each_pixel(fixed_extents != 0, [&](int x, int y, const V3& d) {
		if (n < max_pixels) {
			if (xy) { xy[2 * n] = x; xy[2 * n + 1] = y; }
			if (dir) { dir[3 * n] = d[0]; dir[3 * n + 1] = d[1]; dir[3 * n + 2] = d[2]; }
		}
		n++;
	});
	return n;
}

// ---- Raytracer.trace_frame raytracer.ts:308-330 + ExposureBuffer --------------
// rgb: float32 [screen_h][screen_w][3] ExposureBuffer.pixels, in/out (exposure_buffer.ts:27,68-91).
// n_frames calls of trace_frame are made; before call k (k >= 1, or k >= 0 when frame_first > 0 is the
// running count) the harness does what main.ts:210 does: next_frame().  frame_first is the
// ExposureBuffer.frame_count at the first call (0 after reset_exposure).
// rng_mode 0: one shared sequential FpLcg(seed) as the reference (single-threaded only).
// rng_mode 1: harness reseed before every pixel: rng.seed(seed + pixel_index + frame*W*H) with
//             pixel_index = y*W + x; rows may then run on n_threads threads.
// first_ids: entity of the first collision of each pixel's path in the LAST frame (-1 = none).
// counters: per pixel {nodes, tests, shades, segments} summed over frames (may be NULL).
// crop: if crop_w > 0 only pixels with x in [crop_x, crop_x+crop_w), y in [crop_y, crop_y+crop_h) are traced.
// Returns 0, or -1 "x or y out of bounds" (non-square frame with fixed_extents = 0), -2 bad arguments.
int32_t orc_render(void* scene_h, void* camera_h, const orc_config* cfg_in, int32_t fixed_extents,
                   int32_t n_frames, int32_t frame_first, int32_t rng_mode, double seed, int32_t n_threads,
                   int32_t crop_x, int32_t crop_y, int32_t crop_w, int32_t crop_h, float* rgb, int32_t* first_ids,
                   uint32_t* counters, orc_totals* totals_out) {
	Scene& s = *(Scene*)scene_h;
	const Camera& cam = *(Camera*)camera_h;
	Config cfg{cfg_in->refmax, cfg_in->sky_texture, cfg_in->default_substance, cfg_in->distance_attenuation_factor};
	const int W = cam.screen_w, H = cam.screen_h;
	if (!fixed_extents && W != H) return -1;  // ExposureBuffer.check_bounds throws (exposure_buffer.ts:181-186)
	if (n_frames < 1) return -2;
	if (rng_mode == 0) n_threads = 1;
	if (n_threads < 1) n_threads = 1;

	// the pixel list in generator order
	std::vector<CamPixel> px;
	px.reserve((size_t)W * H);
	cam.for_each_pixel(fixed_extents != 0, [&](int x, int y, const V3& d) {
		if (crop_w > 0 && (x < crop_x || x >= crop_x + crop_w || y < crop_y || y >= crop_y + crop_h)) return;
		px.push_back({x, y, d});
	});

	const V3 start_pos = cam.pos;
	const NodePos start_node = node_at_pos(s.root, start_pos);  // :310
	const int start_ent = entity_at_pos(s, start_pos);           // :312
	const int start_substance = start_ent >= 0 ? s.entities[start_ent].substance : cfg.default_substance;

	Totals grand;
	FpLcg shared_rng;
	shared_rng.seed(seed);
	for (int f = 0; f < n_frames; f++) {
		const int frame_count = frame_first + f;
		const double col_weight = frame_count == 0 ? 1.0 : 1.0 / (1 + frame_count);  // :53-66
		const bool last = f == n_frames - 1;
		std::vector<Totals> part(n_threads);
		auto work = [&](int tid) {
			Walker walker(s.root);  // new_entity_octree_walker, raytracer.ts:295
			WalkerCounters wc;
			walker.ctr = &wc;
			FpLcg local_rng;
			Totals& tot = part[tid];
			const size_t n = px.size();
			// rng_mode 1 only: pixels are independent, so threads take interleaved chunks of the scan order
			const size_t chunk = 256;
			for (size_t i = 0; i < n; i++) {
				if (n_threads > 1 && (i / chunk) % (size_t)n_threads != (size_t)tid) continue;
				const CamPixel& p = px[i];
				const size_t pix = (size_t)p.y * W + p.x;
				FpLcg* rng = &shared_rng;
				if (rng_mode == 1) {
					local_rng.seed(seed + (double)pix + (double)frame_count * (double)W * (double)H);
					rng = &local_rng;
				}
				PixelCounters pc;
				RayResult r = trace_ray(s, cfg, walker, *rng, start_pos, start_node, p.dir, start_substance, pc, tot);
				tot.paths++;
				tot.nodes += pc.nodes; tot.tests += pc.tests; tot.shades += pc.shades; tot.segments += pc.segments;
				// ExposureBuffer.set_color_i :77-91 (doubles, then the Float32Array store)
				float* o = rgb + pix * 3;
				for (int c = 0; c < 3; c++) {
					double cv = r.color[c] * col_weight;
					cv += (double)o[c] * (1 - col_weight);
					o[c] = (float)cv;
				}
				if (last && first_ids) first_ids[pix] = r.first_entity;
				if (counters) {
					uint32_t* k = counters + pix * 4;
					k[0] += pc.nodes; k[1] += pc.tests; k[2] += pc.shades; k[3] += pc.segments;
				}
			}
			tot.cell_steps += wc.cell_steps;
			tot.would_throw += wc.would_throw;
		};
		if (n_threads == 1) {
			work(0);
		} else {
			std::vector<std::thread> th;
			for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
			for (auto& t : th) t.join();
		}
		for (const Totals& t : part) {
			grand.nodes += t.nodes; grand.tests += t.tests; grand.shades += t.shades; grand.segments += t.segments;
			grand.cell_steps += t.cell_steps; grand.would_throw += t.would_throw;
			grand.within_tests += t.within_tests; grand.texture_errors += t.texture_errors;
			grand.acute_warnings += t.acute_warnings; grand.paths += t.paths;
		}
	}
	if (totals_out) {
		totals_out->nodes = grand.nodes; totals_out->tests = grand.tests; totals_out->shades = grand.shades;
		totals_out->segments = grand.segments; totals_out->cell_steps = grand.cell_steps;
		totals_out->would_throw = grand.would_throw; totals_out->within_tests = grand.within_tests;
		totals_out->texture_errors = grand.texture_errors; totals_out->acute_warnings = grand.acute_warnings;
		totals_out->paths = grand.paths;
	}
	return 0;
}

// ---- the demo scene of src/main.ts:60-147,393-396 ------------------------------
// Restates generate_some_aligned_entities + the enclosing scene box.  Image textures are never
// loaded in a headless run (assets are not in the tree; main.ts:381-388 logs and continues), and
// img_txt_prob is 0.0 (:395), so every entity gets a SolidTexture.  Materials / substances /
// textures are appended to the scene tables; returns the number of entities added (incl. the box)
// and writes the sky texture index (ImageTexture fallback colour (0.2,0.2,0.7), :378).
int32_t orc_build_demo_scene(void* h, double seed, int32_t n_entities, int32_t* sky_texture) {
	Scene& s = *(Scene*)h;
	FpLcg rng;
	rng.seed(seed);
	// substances [AIR, WATER, GLASS] (substance.ts:9-11), materials [LIGHT, ROUGH, SMOOTH, TRANSPARENT]
	const int sub0 = orc_add_substance(h, 1.0);
	orc_add_substance(h, 1.333);
	orc_add_substance(h, 1.5);
	const int mat_light = orc_add_material(h, 0, 1, 0, 0);    // SIMPLE_LIGHT_MATERIAL
	orc_add_material(h, 0, 0, 1, 0.5);                        // SIMPLE_ROUGH_MATERIAL
	orc_add_material(h, 0, 0, 1, 0);                          // SIMPLE_SMOOTH_MATERIAL
	orc_add_material(h, 1, 0, 0, 0);                          // SIMPLE_TRANSPARENT_MATERIAL
	const int mat_rough = mat_light + 1;
	*sky_texture = orc_add_texture_solid(h, 0.2, 0.2, 0.7, 1.0);
	const int white = orc_add_texture_solid(h, 1, 1, 1, 1);   // scene_box texture (:390)

	std::vector<std::array<int, 3>> existing;
	int added = 0;
	for (int i = 0; i < n_entities; i++) {
		const int level = 1 + (int)std::floor(rng.next() * 7);
		const int n_quant = 1 << level;
		const double size = 1.0 / n_quant;
		const int qx = js_int32(rng.next() * n_quant);
		const int qy = js_int32(rng.next() * n_quant);
		const int qz = js_int32(rng.next() * n_quant);
		const double x = qx * size + size / 2, y = qy * size + size / 2, z = qz * size + size / 2;
		bool dup = false;
		for (auto& q : existing)
			if (q[0] == qx && q[1] == qy && q[2] == qz) { dup = true; break; }
		if (dup) continue;
		existing.push_back({qx, qy, qz});
		const int ent_class = js_int32(rng.next() * 2);  // 0 SphereEntity, 1 BoxEntity
		const int substance = sub0 + js_int32(rng.next() * 3);
		// get_random_element_with_weights (:77-94) with weights [1,1,1,1]: the one-argument
		// comparator of :84 always returns a positive number, which leaves the index array in
		// its original order (single ascending run), so this is a cumulative pick in index order.
		int material;
		{
			const double wsum = ((0 + 1.0) + 1.0) + 1.0 + 1.0;
			const double wn = 1.0 * (1 / wsum);
			const double rnd = rng.next();
			double weight = 0;
			material = 3;
			for (int k = 0; k < 3; k++) {
				weight += wn;
				if (rnd <= weight) { material = k; break; }
			}
			material += mat_light;
		}
		// get_random_texture (:67-76) with probability 0 -> `rnd <= 0` only for rnd == 0
		int texture;
		{
			const double intensity = material == mat_light ? 5.0 : 1.0;
			const double rnd = rng.next();
			if (rnd <= 0.0) {
				// would pick an (unloaded) ImageTexture: fallback colour black (main.ts:56)
				rng.next();
				texture = orc_add_texture_solid(h, 0, 0, 0, 1);
			} else {
				const double r = rng.next(), g = rng.next(), b = rng.next();
				V3 c = {r, g, b};
				c = scale(c, 1.0 / length(c));  // normalize_self
				c = scale(c, intensity);
				texture = orc_add_texture_solid(h, c[0], c[1], c[2], 1.0);
			}
		}
		const double pos[3] = {x, y, z};
		if (orc_add_entity(h, ent_class, pos, size, material, texture, substance, 16, 0) < 0) return -1;
		added++;
	}
	const double c[3] = {0.5, 0.5, 0.5};
	if (orc_add_entity(h, 1, c, 1.0, mat_rough, white, sub0, 1, 0) < 0) return -1;  // scene_box :391,396
	(void)s;
	return added + 1;
}

// entity table read-back (for feeding the same scene to the product's host mirror)
int32_t orc_entity_count(void* h) { return (int32_t)((Scene*)h)->entities.size(); }
void orc_entity_get(void* h, int32_t id, int32_t* type, double* pos, double* extent, int32_t* material,
                    int32_t* texture, int32_t* substance) {
	const Entity& e = ((Scene*)h)->entities[id];
	*type = e.type;
	for (int i = 0; i < 3; i++) pos[i] = e.pos[i];
	*extent = e.extent; *material = e.material; *texture = e.texture; *substance = e.substance;
}
int32_t orc_material_count(void* h) { return (int32_t)((Scene*)h)->materials.size(); }
void orc_material_get(void* h, int32_t i, int32_t* response, int32_t* light, int32_t* mirror, double* roughness) {
	const Material& m = ((Scene*)h)->materials[i];
	*response = m.response; *light = m.light_source; *mirror = m.mirror; *roughness = m.roughness;
}
int32_t orc_texture_count(void* h) { return (int32_t)((Scene*)h)->textures.size(); }
void orc_texture_get(void* h, int32_t i, int32_t* kind, double* color, int32_t* w, int32_t* hgt, int32_t* loaded) {
	const Texture& t = ((Scene*)h)->textures[i];
	*kind = t.kind;
	for (int k = 0; k < 4; k++) color[k] = t.color[k];
	*w = t.width; *hgt = t.height; *loaded = t.loaded;
}
int32_t orc_substance_count(void* h) { return (int32_t)((Scene*)h)->substances.size(); }
double orc_substance_get(void* h, int32_t i) { return ((Scene*)h)->substances[i]; }

// ------------------------------------------------------------------ View.draw_ebuffer (src/view/view.ts:34-38)
// ExposureBuffer.rgb_to_y (src/view/exposure_buffer.ts:161-173): float64 arithmetic on the float32 pixels
static inline double rgb_to_y(const float* px) { return 0.299 * (double)px[0] + 0.587 * (double)px[1] + 0.114 * (double)px[2]; }
// mathutils.clamp (src/math/mathutils.ts:18-20) = Math.max(Math.min(x, max), min): NaN propagates
static inline double js_math_min(double a, double b) { return (std::isnan(a) || std::isnan(b)) ? NAN : (a < b ? a : b); }
static inline double js_math_max(double a, double b) { return (std::isnan(a) || std::isnan(b)) ? NAN : (a > b ? a : b); }
static inline double js_clamp(double x, double lo, double hi) { return js_math_max(js_math_min(x, hi), lo); }

// ExposureBuffer.get_mean / get_variance / get_absolute_dev (src/view/exposure_buffer.ts:93-142): sequential
// float64 sums in pixel order (the `_mean` style caches are never filled, so every call recomputes).
// out = {mean, variance, absolute deviation}
void orc_exposure_stats(const float* pixels, int64_t n_pixels, double* out) {
	double mean = 0;
	for (int64_t i = 0; i < n_pixels * 3; i += 3) mean += rgb_to_y(pixels + i);
	mean /= (double)n_pixels;
	double variance = 0, dev = 0;
	for (int64_t i = 0; i < n_pixels * 3; i += 3) {
		const double delta = rgb_to_y(pixels + i) - mean;
		variance += delta * delta;
		dev += std::fabs(delta);
	}
	out[0] = mean;
	out[1] = variance / (double)n_pixels;
	out[2] = dev / (double)n_pixels;
}

// ToneMapper.get_dynamic_range (src/view/tone_mapping.ts:24-79).  kind 0 ToneMapper_Identity, 1
// ToneMapper_StdDevAroundMean, 2 ToneMapper_AbsDevAroundMean; dynamic_coef = 1 << dynamic_range (:42).
void orc_dynamic_range(int32_t kind, int32_t dynamic_range, double min_dynamic, double max_dynamic, const double* stats,
                       double* out) {
	if (kind == 0) { out[0] = 0; out[1] = 1; return; }
	const double dynamic_coef = (double)(int32_t)(1u << (dynamic_range & 31));
	const double mean_br = stats[0];
	const double dev_br = kind == 1 ? std::sqrt(stats[1]) : stats[2];
	double drange_max = js_math_min(mean_br + dev_br, max_dynamic);
	double drange_min = drange_max / dynamic_coef;
	if (drange_min < min_dynamic) {
		drange_min = min_dynamic;
		drange_max = drange_min * dynamic_coef;
	}
	out[0] = drange_min;
	out[1] = drange_max;
}

// ExposureBuffer.discretize_to_screen (src/view/exposure_buffer.ts:145-158) into CanvasScreen.set_pixel_i /
// convert_color (src/view/screen_canvas.ts:45-56,101-103).  As written at HEAD: `pixels.slice(i, i+2)` keeps
// TWO channels, `.map` on a Float32Array rounds the clamped products to float32, and convert_color's third
// element is undefined, which a Uint8ClampedArray stores as 0 - the blue channel of every pixel is 0
// (SURVEY.md 8f N2).  rgba: [n_pixels][4] bytes.
void orc_discretize(const float* pixels, int64_t n_pixels, double drange_low, double drange_high, uint8_t* rgba) {
	const double drange = drange_high - drange_low;
	for (int64_t px_i = 0, i = 0; px_i < n_pixels; ++px_i, i += 3) {
		const double px_brightness = rgb_to_y(pixels + i);
		const double cmpr_brightness = (px_brightness - drange_low) / drange;
		const double scale_coef = cmpr_brightness / (px_brightness + std::numeric_limits<double>::epsilon());
		for (int k = 0; k < 2; k++) {
			const float compressed = (float)js_clamp((double)pixels[i + k] * scale_coef, 0.0, 1.0);  // Float32Array.map
			const double v = js_clamp((double)compressed, 0.0, 1.0) * 255.0;
			rgba[px_i * 4 + k] = (uint8_t)js_int32(v);  // (x*255) << 0, NaN -> 0
		}
		rgba[px_i * 4 + 2] = 0;
		rgba[px_i * 4 + 3] = 0xff;
	}
}

}  // extern "C"
