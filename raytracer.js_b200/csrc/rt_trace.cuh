// rt_trace.cuh — the per-ray hot path: octree walk in the reference's visit order, entity hit tests,
// shading, bounces.  Everything is RT_HD so that the same source is the CUDA kernel body and (for
// logic tests on a box without a GPU) a test-only host build; the shipped library only ever runs it
// on the device.
//
// Precision model ("float search, float64 confirm"):
//  * the octree walk and the scan of the entity lists run in float32 on conservative (inflated)
//    bounds: they can only produce false candidates, never lose a hit by more than rounding noise;
//  * every candidate is confirmed in float64 with the reference's own formula and operation order
//    (src/math/intersection.ts:109-128,150-204), so hit/miss decisions, hit points, normals, path
//    lengths and colours are the reference's float64 values;
//  * ray state (origin, direction, colour, path length, RNG) is float64; the float copies are
//    re-derived per segment.
//
// Reference citations are relative to /root/reference.
#pragma once
#include <math.h>
#include <string.h>

#include "rt_common.h"

#define RT_JS_EPSILON 2.220446049250313e-16
#define RT_JS_PI 3.141592653589793

// ------------------------------------------------------------------ exact float64 primitives
// No FMA contraction: the reference rounds after every multiply and add.
#if defined(__CUDACC__)
RT_HD double xmul(double a, double b) { return __dmul_rn(a, b); }
RT_HD double xadd(double a, double b) { return __dadd_rn(a, b); }
RT_HD double xsub(double a, double b) { return __dsub_rn(a, b); }
RT_HD double xdiv(double a, double b) { return __ddiv_rn(a, b); }
RT_HD double xsqrt(double a) { return __dsqrt_rn(a); }
#else
// host build: compiled with -ffp-contract=off
RT_HD double xmul(double a, double b) { return a * b; }
RT_HD double xadd(double a, double b) { return a + b; }
RT_HD double xsub(double a, double b) { return a - b; }
RT_HD double xdiv(double a, double b) { return a / b; }
RT_HD double xsqrt(double a) { return sqrt(a); }
#endif

// ------------------------------------------------------------------ loads (read-only path, 16 B)
// a record that will be fetched an iteration or more from now: start the fetch (no register, no dependency)
RT_HD void prefetch_record(const void* p) {
#ifdef __CUDA_ARCH__
	asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
	(void)p;
#endif
}

RT_HD RtF4 ld(const RtF4* p) {
#if defined(__CUDACC__)
	const float4 v = __ldg(reinterpret_cast<const float4*>(p));
	return RtF4{v.x, v.y, v.z, v.w};
#else
	return *p;
#endif
}
RT_HD RtI4 ld(const RtI4* p) {
#if defined(__CUDACC__)
	const int4 v = __ldg(reinterpret_cast<const int4*>(p));
	return RtI4{v.x, v.y, v.z, v.w};
#else
	return *p;
#endif
}
RT_HD RtD2 ld(const RtD2* p) {
#if defined(__CUDACC__)
	const double2 v = __ldg(reinterpret_cast<const double2*>(p));
	return RtD2{v.x, v.y};
#else
	return *p;
#endif
}
RT_HD RtD4 ld(const RtD4* p) {
#if defined(__CUDACC__)
	const double2 a = __ldg(reinterpret_cast<const double2*>(p));
	const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
	return RtD4{a.x, a.y, b.x, b.y};
#else
	return *p;
#endif
}
// ray-generation checkpoints: read once per packet - through the L2 only, the L1 is for the scene
RT_HD double ld_stream(const double* p) {
#if defined(__CUDACC__)
	return __ldcg(p);
#else
	return *p;
#endif
}
RT_HD RtD2 ld_stream(const RtD2* p) {
#if defined(__CUDACC__)
	const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
	return RtD2{v.x, v.y};
#else
	return *p;
#endif
}
RT_HD int ld(const int* p) {
#if defined(__CUDACC__)
	return __ldg(p);
#else
	return *p;
#endif
}

RT_HD float int_as_float(int v) {
#if defined(__CUDACC__)
	return __int_as_float(v);
#else
	float f;
	memcpy(&f, &v, sizeof f);
	return f;
#endif
}

// vector.dot (src/math/vector.ts:76-84): sum starts at 0, left to right
RT_HD double dot3(const double* a, const double* b) {
	double s = 0.0;
	s = xadd(s, xmul(a[0], b[0]));
	s = xadd(s, xmul(a[1], b[1]));
	s = xadd(s, xmul(a[2], b[2]));
	return s;
}
RT_HD double js_sign(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : x); }
RT_HD bool js_is_negative(double x) { return x < 0 || (x == 0 && signbit(x)); }  // mathutils.ts:45-47

struct RtCollision {
	double point[3];
	double normal[3];
};

struct RtCounts {
	unsigned int segments, nodes, tests, shades, confirms;
	unsigned int errors;  // RT_ERRFLAG_* raised below the functions that carry an error word
};

// ------------------------------------------------------------------ float64 confirmations
// Box.line_intersection (src/math/intersection.ts:150-204).  Returns false for `[]`.
RT_HD bool exact_box_params(const double* c, double size, const double* o, const double* d, double& u1,
                            double& u2, int& i1, int& i2) {
	double p[6], q[6];
#pragma unroll
	for (int k = 0; k < 3; k++) {
		const double tl = xsub(c[k], xmul(size, 0.5));
		p[2 * k] = -d[k];
		p[2 * k + 1] = d[k];
		q[2 * k] = xsub(o[k], tl);
		q[2 * k + 1] = xsub(xadd(tl, size), o[k]);
	}
	u1 = -INFINITY;
	u2 = INFINITY;
	i1 = -1;
	i2 = -1;
#pragma unroll
	for (int i = 0; i < 6; i++) {
		const double u = xdiv(q[i], p[i]);
		if (js_is_negative(p[i])) {
			if (u > u1) { u1 = u; i1 = i; }
		} else {
			if (u < u2) { u2 = u; i2 = i; }
		}
	}
	return !(u1 > u2);
}

// SphereEntity.collision_info (src/entities/entity_sphere.ts:68-88) over
// Sphere.line_intersection (src/math/intersection.ts:109-128), FORWARD selection (:207-216)
RT_HD bool exact_sphere(const RtD4& g, const double* o, const double* d, RtCollision& col) {
	const double pos[3] = {g.x, g.y, g.z};
	const double radius = xdiv(g.w, 2.0);
	const double radius_sq = xmul(radius, radius);
	const double dot_pp = dot3(pos, pos);
	const double dist[3] = {xsub(o[0], pos[0]), xsub(o[1], pos[1]), xsub(o[2], pos[2])};
	const double a = dot3(d, d);
	const double b = xmul(dot3(dist, d), 2.0);
	const double c = xsub(xsub(xadd(dot3(o, o), dot_pp), xmul(dot3(o, pos), 2.0)), radius_sq);
	const double delta = xsub(xmul(b, b), xmul(xmul(a, c), 4.0));
	if (delta < 0) return false;
	const double s_delta = xsqrt(delta);
	const double a2 = xmul(a, 2.0);
	const double tmp1 = xdiv(-b, a2);
	const double tmp2 = xdiv(s_delta, a2);
	const double t1 = xsub(tmp1, tmp2);
	const double t2 = xadd(tmp1, tmp2);
	double t;
	if (t1 >= 0) t = t1;
	else if (t2 >= 0) t = t2;
	else return false;
	const double k = xdiv(2.0, g.w);
#pragma unroll
	for (int i = 0; i < 3; i++) {
		col.point[i] = xadd(o[i], xmul(d[i], t));
		col.normal[i] = xmul(xsub(col.point[i], pos[i]), k);
	}
	const double sg = -js_sign(dot3(d, col.normal));
#pragma unroll
	for (int i = 0; i < 3; i++) col.normal[i] = xmul(col.normal[i], sg);
	return true;
}

// BoxEntity.collision_info (src/entities/entity_box.ts:54-73)
RT_HD bool exact_box(const RtD4& g, const double* o, const double* d, RtCollision& col) {
	const double pos[3] = {g.x, g.y, g.z};
	double u1, u2;
	int i1, i2;
	if (!exact_box_params(pos, g.w, o, d, u1, u2, i1, i2)) return false;
	double t;
	int face;
	if (u1 >= 0) { t = u1; face = i1; }
	else if (u2 >= 0) { t = u2; face = i2; }
	else return false;
#pragma unroll
	for (int i = 0; i < 3; i++) {
		col.point[i] = xadd(o[i], xmul(d[i], t));
		col.normal[i] = 0.0;
	}
	if (face >= 0) {
		double n[3] = {0.0, 0.0, 0.0};
		n[face >> 1] = (face & 1) ? 1.0 : -1.0;
		const double sg = -js_sign(dot3(d, n));
#pragma unroll
		for (int i = 0; i < 3; i++) col.normal[i] = xmul(n[i], sg);
	}
	return true;
}

// ------------------------------------------------------------------ float32 candidate filters
struct RtRayF {
	float ox, oy, oz;
	float dx, dy, dz;
	float ix, iy, iz;  // 1/d (±inf for zero components)
	float inv_a;       // 1/(d.d)
};

RT_HD RtRayF make_ray_f(const double* o, const double* d) {
	RtRayF r;
	r.ox = (float)o[0]; r.oy = (float)o[1]; r.oz = (float)o[2];
	r.dx = (float)d[0]; r.dy = (float)d[1]; r.dz = (float)d[2];
	r.ix = 1.0f / r.dx; r.iy = 1.0f / r.dy; r.iz = 1.0f / r.dz;
	r.inv_a = 1.0f / (r.dx * r.dx + r.dy * r.dy + r.dz * r.dz);
	return r;
}

// true if the entity MAY have a forward intersection (conservative by err_l)
RT_HD bool candidate(const RtF4& g, const RtRayF& r, float err_l) {
	const float cx = g.x - r.ox, cy = g.y - r.oy, cz = g.z - r.oz;
	if (g.w > 0.0f) {
		// sphere: squared distance from the centre to the ray's line against (radius + err)^2
		const float tca = cx * r.dx + cy * r.dy + cz * r.dz;
		const float s = tca * r.inv_a;
		const float lx = cx - s * r.dx, ly = cy - s * r.dy, lz = cz - s * r.dz;
		const float l2 = lx * lx + ly * ly + lz * lz;
		const float rr = g.w + err_l;
		const float rr2 = rr * rr;
		if (l2 > rr2) return false;
		if (tca >= 0.0f) return true;
		return cx * cx + cy * cy + cz * cz <= rr2;  // origin inside: the far root is forward
	}
	// box: slab test on the inflated box
	const float h = err_l - g.w;
	float t0 = (cx - h) * r.ix, t1 = (cx + h) * r.ix;
	float tmin = fminf(t0, t1), tmax = fmaxf(t0, t1);
	t0 = (cy - h) * r.iy; t1 = (cy + h) * r.iy;
	tmin = fmaxf(tmin, fminf(t0, t1)); tmax = fminf(tmax, fmaxf(t0, t1));
	t0 = (cz - h) * r.iz; t1 = (cz + h) * r.iz;
	tmin = fmaxf(tmin, fminf(t0, t1)); tmax = fminf(tmax, fmaxf(t0, t1));
	const float slack = err_l * fmaxf(fabsf(r.ix), fmaxf(fabsf(r.iy), fabsf(r.iz)));
	return !(tmin > tmax + slack) && !(tmax < -slack);
}

// ------------------------------------------------------------------ point location (float64)
// node_at_pos (src/octree_space.ts:61-93), half-open cube test of src/space.ts:55-66
RT_HD bool node_at_pos(const RtDevScene& S, const double* p, int& node, int& octant) {
	double np[3] = {S.root_pos[0], S.root_pos[1], S.root_pos[2]};
	double ns = S.root_size;
#pragma unroll
	for (int i = 0; i < 3; i++)
		if (!(p[i] >= np[i] && p[i] < xadd(np[i], ns))) return false;
	int cur = 0, next = 0, idx = 0;
	while (next >= 0) {
		const double k = xdiv(2.0, ns);
		const int ix = (int)xmul(xsub(p[0], np[0]), k);
		const int iy = (int)xmul(xsub(p[1], np[1]), k);
		const int iz = (int)xmul(xsub(p[2], np[2]), k);
		cur = next;
		idx = (iz << 2) + (iy << 1) + ix;
		if (idx < 0 || idx > 7) return false;  // the reference's Octree.get() would throw here
		next = ld(S.node_child + cur * 8 + idx);
		ns = xdiv(ns, 2.0);
		np[0] = xadd(np[0], xmul((double)ix, ns));
		np[1] = xadd(np[1], xmul((double)iy, ns));
		np[2] = xadd(np[2], xmul((double)iz, ns));
	}
	node = cur;
	octant = idx;
	return true;
}

// entity_at_pos (src/octree_entity.ts:191-202) -> slot or -1
RT_HD int entity_at_pos(const RtDevScene& S, const double* p) {
	int node, octant;
	if (!node_at_pos(S, p, node, octant)) return -1;
	while (node >= 0) {
		const RtI4 link = ld(S.node_link + node);
		for (int s = link.z; s < link.z + link.w; s++) {
			const RtD4 g = ld(S.slot_geom64 + s);
			const int type = ld(&S.slot_attr[s].y) >> RT_ATTR_TYPE_SHIFT;
			bool within;
			if (type == 0) {  // SphereEntity.is_within, entity_sphere.ts:63-66 (_radius_sq = d*d/4)
				const double dist[3] = {xsub(p[0], g.x), xsub(p[1], g.y), xsub(p[2], g.z)};
				within = dot3(dist, dist) <= xdiv(xmul(g.w, g.w), 4.0);
			} else {  // BoxEntity.is_within, entity_box.ts:47-52: [pos, pos+size) from the *centre*
				within = p[0] >= g.x && p[0] < xadd(g.x, g.w) && p[1] >= g.y && p[1] < xadd(g.y, g.w) &&
				         p[2] >= g.z && p[2] < xadd(g.z, g.w);
			}
			if (within) return s;
		}
		node = link.x;
	}
	return -1;
}

// ------------------------------------------------------------------ the walk
// Per-segment search context.  `rel` (optional) is the per-frame origin-relative copy of slot_geom for
// rays that start at the camera (see make_prim_record / rt_prepare_primary).
struct RtSearch {
	RtRayF r;
	const RtF4* rel;        // non-null: primary segment, origin-relative sphere tests
	unsigned chain_mask;    // bit k: the k-th origin-chain node's list may hold a hit (else skip its scan)
	int chain_levels;       // number of chain levels that were pre-tested (0: scan everything)
};

// The per-frame origin-relative record of one slot, for rays that start at the camera:
//   xyz = centre - camera, taken in float64 (no cancellation) and rounded once;
//   w   = +(radius+err)          sphere, camera outside it: the quick test of candidate_rel applies
//         -(bounding radius+err) box, camera outside its bounding sphere: packet cull on the bounding
//                                sphere, per-ray test = the generic slab test on slot_geom
//         +inf                   camera inside the (bounding) sphere: always a candidate
// (the radius itself, not its square: the packet's cone test - once per slot and packet - needs the radius,
// the per-ray test squares it once per survivor)
RT_HD RtF4 make_prim_record(const RtD4& g, bool is_sphere, double ox, double oy, double oz, float err_l) {
	const double cx = g.x - ox, cy = g.y - oy, cz = g.z - oz;
	const float rr = (is_sphere ? (float)(g.w * 0.5) : (float)(g.w * 0.8660254037844387) * 1.000001f) + err_l;
	const float rr2 = rr * rr;
	const bool inside = cx * cx + cy * cy + cz * cz <= (double)rr2 * 1.000001;
	return RtF4{(float)cx, (float)cy, (float)cz, inside ? INFINITY : (is_sphere ? rr : -rr)};
}

// candidate test against the origin-relative sphere record (w finite, > 0): ~12 flops, no
// centre - origin subtraction
RT_HD bool candidate_rel(const RtF4& g, const RtRayF& r) {
	const float tca = g.x * r.dx + g.y * r.dy + g.z * r.dz;
	const float s = tca * r.inv_a;
	const float lx = g.x - s * r.dx, ly = g.y - s * r.dy, lz = g.z - s * r.dz;
	const float l2 = lx * lx + ly * ly + lz * lz;
	return l2 <= g.w * g.w && tca >= 0.0f;  // camera outside the sphere: a forward root needs the centre ahead
}

RT_HD bool slot_candidate(const RtDevScene& S, const RtSearch& q, int s) {
	if (q.rel) {
		const RtF4 g = ld(q.rel + s);
		if (g.w > 0.0f && g.w != INFINITY) return candidate_rel(g, q.r);
	}
	return candidate(ld(S.slot_geom + s), q.r, S.err_l);
}

// float64 confirmation of slot s: Entity.collision_info(ray) != null, collision in `col`
RT_HD bool confirm_slot(const RtDevScene& S, int s, const double* o, const double* d, RtCollision& col) {
	const RtD4 g = ld(S.slot_geom64 + s);
	return ld(S.slot_geom + s).w > 0.0f ? exact_sphere(g, o, d, col) : exact_box(g, o, d, col);
}

// Long lists: every entity of the list the ray hits is found through the list's BVH and the lowest slot
// (= first in insertion order) wins, which is what the reference's linear scan with `break` returns.
RT_HD int scan_list_bvh(const RtDevScene& S, const RtSearch& q, int root, const double* o, const double* d,
                        RtCollision& col, RtCounts& cnt) {
	const RtRayF& r = q.r;
	int stack[RT_BVH_STACK];
	int sp = 0, best = 0x7fffffff;
	stack[sp++] = root;
	const float slack = S.err_l * fmaxf(fabsf(r.ix), fmaxf(fabsf(r.iy), fabsf(r.iz)));
	while (sp > 0) {
		const RtBvhNode* np = S.bvh_nodes + stack[--sp];
		const RtF4 n0 = ld(reinterpret_cast<const RtF4*>(np));
		const RtI4 n1 = ld(reinterpret_cast<const RtI4*>(np) + 1);
		// n0 = lo.xyz, hi.x ; n1 = hi.y, hi.z (as bits), a, b
		const float hy = int_as_float(n1.x), hz = int_as_float(n1.y);
		float t0 = (n0.x - r.ox) * r.ix, t1 = (n0.w - r.ox) * r.ix;
		float tmin = fminf(t0, t1), tmax = fmaxf(t0, t1);
		t0 = (n0.y - r.oy) * r.iy; t1 = (hy - r.oy) * r.iy;
		tmin = fmaxf(tmin, fminf(t0, t1)); tmax = fminf(tmax, fmaxf(t0, t1));
		t0 = (n0.z - r.oz) * r.iz; t1 = (hz - r.oz) * r.iz;
		tmin = fmaxf(tmin, fminf(t0, t1)); tmax = fminf(tmax, fmaxf(t0, t1));
		if (tmin > tmax + slack || tmax < -slack) continue;  // (NaN from 0 * inf compares false: kept)
		if (n1.w < 0) {
			if (-(n1.w + 1) >= best) continue;  // nothing below this node can beat the hit we have
			if (sp + 2 <= RT_BVH_STACK) {
				stack[sp++] = n1.z + 1;
				stack[sp++] = n1.z;  // the child with the lowest slot first
			} else {
				cnt.errors |= RT_ERRFLAG_STACK;  // (balanced median-split BVHs: 40 levels are never reached; never silent)
			}
			continue;
		}
		for (int k = 0; k < n1.w; k++) {
			const int s = ld(S.bvh_slots + n1.z + k);
			if (s >= best) break;  // ascending within the leaf
			if (slot_candidate(S, q, s)) {
				RtCollision c;
				if (confirm_slot(S, s, o, d, c)) {
					best = s;
					col = c;
					break;
				}
			}
		}
	}
	return best == 0x7fffffff ? -1 : best;
}

// Scans the list of `node`, slots [beg,end) in insertion order; returns the slot of the first entity whose
// float64 collision_info is non-null (src/raytracer.ts:186-195), or -1.
template <bool COUNT>
RT_HD int scan_list(const RtDevScene& S, const RtSearch& q, int node, int beg, int end, const double* o, const double* d,
                    RtCollision& col, RtCounts& cnt) {
	if (end - beg >= RT_BVH_MIN_LIST) {
		const int s = scan_list_bvh(S, q, ld(S.node_bvh + node), o, d, col, cnt);
		// the counters count the reference's linear scan: up to and including the hit, or the whole list
		if (COUNT) {
			cnt.tests += (unsigned)(s >= 0 ? s - beg + 1 : end - beg);
			cnt.confirms++;
		}
		return s;
	}
	int s = beg;
	while (true) {
		for (; s < end; ++s)
			if (slot_candidate(S, q, s)) break;
		if (s >= end) break;
		const bool hit = confirm_slot(S, s, o, d, col);
		if (COUNT) cnt.confirms++;
		if (hit) {
			if (COUNT) cnt.tests += (unsigned)(s - beg + 1);
			return s;
		}
		++s;
	}
	if (COUNT) cnt.tests += (unsigned)(end - beg);
	return -1;
}

// OctreeWalker.next() (src/octree_space.ts:316-361) fused with the list scan of Ray.trace
// (src/raytracer.ts:179-195).  State names follow the reference.  (node, octant) is the start
// position: node_at_pos of the origin, or octant < 0 for "origin outside the root: root only"
// (src/octree_space.ts:272-275,283-287).  Returns the hit slot or -1.
//
// Shape: "while-while".  The inner loop only advances the walker until it returns a node whose list
// has to be scanned; the scan runs after it, so the lanes of a warp re-converge between the two phases
// instead of interleaving cell steps of one ray with list scans of another.
// The nodes that contain the ray origin (the "origin chain") are returned in post-order, each when the
// ray leaves it: they are recognised by step_back at depth 0, and for primary rays
// their lists were pre-tested in lock-step by the whole warp (q.chain_mask).
template <bool COUNT>
RT_HD int walk_and_scan(const RtDevScene& S, const RtSearch& q, int node, int octant, const double* o,
                        const double* d, RtCollision& col, RtCounts& cnt) {
	const RtRayF& r = q.r;
	bool cur_returned = false, stepped_in = false, ahead = false;
	int depth = 0;
	int chain_level = 0;
	bool chain_pending = octant < 0;  // root-only mode: the single returned node is the chain's only node
	float npx = r.ox, npy = r.oy, npz = r.oz;  // next_pos[0]
	int face = 0;                               // next_pos[1] as a face index (axis*2 + positive)
	RtF4 g = ld(S.node_geom + node);
	bool done = false;
	while (true) {
		int scan_beg = 0, scan_end = 0, scan_node = 0;
		// ---- phase 1: advance the walker to the next node with a list to scan
		while (true) {
			const int child = octant >= 0 ? ld(S.node_child + node * 8 + octant) : node;
			if (!cur_returned && child >= 0) {
				cur_returned = true;
				if (COUNT) cnt.nodes++;
				const RtI4 link = ld(S.node_link + child);
				bool skip = link.w == 0;
				if (chain_pending) {
					chain_pending = false;
					if (chain_level < q.chain_levels && !((q.chain_mask >> chain_level) & 1u)) {
						skip = true;  // pre-tested: no entity of this list can be hit
						if (COUNT) cnt.tests += (unsigned)link.w;
					}
					chain_level++;
				}
				if (!skip) {
					scan_node = child;
					scan_beg = link.z;
					scan_end = link.z + link.w;
					break;
				}
			}
			if (octant >= 0) {
				if (!ahead) {
					if (!stepped_in && child >= 0) {  // step_in: octant_adj_pos(child, next_pos[0]) :41-50
						g = ld(S.node_geom + child);
						const float h = g.w * 0.5f;
						octant = (npx >= g.x + h ? 1 : 0) | (npy >= g.y + h ? 2 : 0) | (npz >= g.z + h ? 4 : 0);
						node = child;
						depth++;
						cur_returned = false;
						continue;
					}
					// update_next_pos :369-384: exit parameter and face of the cell (node, octant);
					// strict comparisons in face order -x,+x,-y,+y,-z,+z  =>  ties go x, then y, then z
					const float h = g.w * 0.5f;
					const float lox = g.x + ((octant & 1) ? h : 0.0f);
					const float loy = g.y + ((octant & 2) ? h : 0.0f);
					const float loz = g.z + ((octant & 4) ? h : 0.0f);
					const float tx = r.dx != 0.0f ? ((r.dx > 0.0f ? lox + h : lox) - r.ox) * r.ix : INFINITY;
					const float ty = r.dy != 0.0f ? ((r.dy > 0.0f ? loy + h : loy) - r.oy) * r.iy : INFINITY;
					const float tz = r.dz != 0.0f ? ((r.dz > 0.0f ? loz + h : loz) - r.oz) * r.iz : INFINITY;
					float t = tx;
					face = r.dx > 0.0f ? 1 : 0;
					if (ty < t) { t = ty; face = r.dy > 0.0f ? 3 : 2; }
					if (tz < t) { t = tz; face = r.dz > 0.0f ? 5 : 4; }
					npx = r.ox + r.dx * t;
					npy = r.oy + r.dy * t;
					npz = r.oz + r.dz * t;
				}
				const int axis = face >> 1;
				const int bit = (octant >> axis) & 1;
				if (bit != (face & 1)) {  // neighbour octant inside the parent cube :344-352
					octant ^= 1 << axis;
					cur_returned = false;
					stepped_in = false;
					ahead = false;
					continue;
				}
				ahead = true;
			}
			// step_back :280-308
			stepped_in = true;
			if (octant < 0) { done = true; break; }
			if (depth > 0) { depth--; cur_returned = true; }
			else { cur_returned = false; chain_pending = true; }  // the node we leave contained the origin
			const RtI4 link = ld(S.node_link + node);
			if (link.x >= 0) {
				octant = link.y;
				node = link.x;
				g = ld(S.node_geom + node);
			} else {
				octant = -1;
			}
		}
		if (done) break;
		// ---- phase 2: scan the list
		const int s = scan_list<COUNT>(S, q, scan_node, scan_beg, scan_end, o, d, col, cnt);
		if (s >= 0) return s;
	}
	return -1;
}

// RT_PRECISION_F64: the same walker with the reference's own float64 expressions - the cells are taken from
// node_geom64 (the OctreeDim values the reference holds), the exit point of a cell is Box.line_intersection's u2 and
// its face (update_next_pos, src/octree_space.ts:369-384 via dim_relative_to_parent :127-136), a step into a child
// is octant_adj_pos (:41-50) of that point.  No float32 value takes part in a decision of the walk, so the tree may
// be as deep as the reference allows; the list scan keeps its float32 FILTER (conservative by err_l at any depth)
// in front of the float64 confirmations.
template <bool COUNT>
RT_HD int walk_and_scan64(const RtDevScene& S, const RtSearch& q, int node, int octant, const double* o,
                          const double* d, RtCollision& col, RtCounts& cnt) {
	bool cur_returned = false, stepped_in = false, ahead = false;
	int depth = 0;
	double np[3] = {o[0], o[1], o[2]};  // next_pos[0]
	int face = -1;                       // next_pos[1] as a face index (axis*2 + positive), -1: the zero vector
	RtD4 g = ld(S.node_geom64 + node);
	while (true) {
		const int child = octant >= 0 ? ld(S.node_child + node * 8 + octant) : node;
		if (!cur_returned && child >= 0) {
			cur_returned = true;
			if (COUNT) cnt.nodes++;
			const RtI4 link = ld(S.node_link + child);
			if (link.w > 0) {
				const int s = scan_list<COUNT>(S, q, child, link.z, link.z + link.w, o, d, col, cnt);
				if (s >= 0) return s;
			}
		}
		if (octant >= 0) {
			if (!ahead) {
				if (!stepped_in && child >= 0) {  // step_in :310-314 with octant_adj_pos(child, next_pos[0])
					g = ld(S.node_geom64 + child);
					const double h = g.w / 2;
					octant = (np[0] >= xadd(g.x, h) ? 1 : 0) | (np[1] >= xadd(g.y, h) ? 2 : 0) | (np[2] >= xadd(g.z, h) ? 4 : 0);
					node = child;
					depth++;
					cur_returned = false;
					continue;
				}
				// update_next_pos: the cell (node, octant) as a box, Box.line_intersection, the far parameter and its face
				const double ph = g.w / 2;
				const double gp[3] = {g.x, g.y, g.z};
				double c[3];
#pragma unroll
				for (int k = 0; k < 3; k++) c[k] = xadd(xadd(gp[k], xmul((double)((octant >> k) & 1), ph)), xmul(0.5, ph));
				double u1, u2;
				int i1, i2;
				exact_box_params(c, ph, o, d, u1, u2, i1, i2);
				if (i2 < 0) return -1;  // no parameter at all (a zero direction): the reference throws a TypeError here
#pragma unroll
				for (int k = 0; k < 3; k++) np[k] = xadd(o[k], xmul(d[k], u2));
				face = i2;
			}
			const int axis = face >> 1;
			const int bit = (octant >> axis) & 1;
			if (bit != (face & 1)) {  // octant + normal stays inside the parent cube :344-352
				octant ^= 1 << axis;
				cur_returned = false;
				stepped_in = false;
				ahead = false;
				continue;
			}
			ahead = true;
		}
		// step_back :280-308
		stepped_in = true;
		if (octant < 0) return -1;
		if (depth > 0) { depth--; cur_returned = true; }
		else cur_returned = false;
		const RtI4 link = ld(S.node_link + node);
		if (link.x >= 0) {
			octant = link.y;
			node = link.x;
			g = ld(S.node_geom64 + node);
		} else {
			octant = -1;
		}
	}
}

// Lock-step pre-test of the origin-chain lists for a primary ray: every lane of the warp runs over the
// same slots (uniform 16 B loads).  Bit k of the result is set if the list of chain level k holds a
// (conservative) candidate.
RT_HD unsigned pretest_chain(const RtFrame& F, const RtDevScene& S, const RtSearch& q) {
	unsigned mask = 0;
	for (int k = 0; k < F.chain_levels; k++) {
		bool any = false;
		for (int s = F.chain_beg[k]; s < F.chain_end[k]; s++) any |= slot_candidate(S, q, s);
		mask |= (any ? 1u : 0u) << k;
	}
	return mask;
}

// ------------------------------------------------------------------ the packet walk (camera rays)
// All camera rays share one origin, so the 32 rays of an 8x4-pixel patch form a narrow packet.  The
// reference's visit order does not depend on the individual ray once the signs of the
// direction components are fixed: among the octants of a node a ray can only move from octant a to
// octant b if (a ^ neg) is a bitwise subset of (b ^ neg) (neg = mask of negative direction signs), so
// "children in ascending (octant ^ neg), pre-order" is a linear extension of every ray's own order, and
// the nodes that contain the origin are returned after everything below them (post-order).  A node the
// packet visits but one ray does not pierce cannot hold a hit for that ray (every entity lies inside its
// node's cube, src/octree_entity.ts:60-79), so visiting the UNION of the rays' nodes in that order gives
// every ray its own first hit.  The warp therefore walks ONE node sequence in lock-step, and the lanes
// are used twice over:
//   * list scan: lane j tests slot base+j against the packet's bounding cone (coalesced 16-byte loads,
//     one vote), and only the survivors are tested ray by ray (one uniform 16-byte load each);
//   * child selection: lanes 0..7 test the 8 octants of the popped node against the packet's slab
//     intervals and push the survivors with one vote.
// Written once for both builds: per-lane values are arrays of RT_NL elements (1 on the device, where a
// lane is a thread; 32 in the test-only host build, where a lane is a loop iteration).
#if defined(__CUDACC__)
#define RT_NL 1
#define RT_LANES(l, lane) for (int l = 0, lane = (int)(threadIdx.x & 31u); l < 1; ++l)
#else
#define RT_NL 32
#define RT_LANES(l, lane) for (int l = 0, lane = 0; l < 32; ++l, ++lane)
#endif

RT_HD unsigned warp_ballot(const bool (&p)[RT_NL]) {
#if defined(__CUDACC__)
	return __ballot_sync(0xffffffffu, p[0]);
#else
	unsigned m = 0;
	for (int l = 0; l < 32; l++) m |= (p[l] ? 1u : 0u) << l;
	return m;
#endif
}
// min / max of NON-NEGATIVE floats (their bit patterns order like unsigned integers): one redux.sync
RT_HD float warp_min_pos(const float (&v)[RT_NL]) {
#if defined(__CUDACC__)
	return __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(v[0])));
#else
	float m = v[0];
	for (int l = 1; l < 32; l++) m = v[l] < m ? v[l] : m;
	return m;
#endif
}
RT_HD float warp_max_pos(const float (&v)[RT_NL]) {
#if defined(__CUDACC__)
	return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v[0])));
#else
	float m = v[0];
	for (int l = 1; l < 32; l++) m = v[l] > m ? v[l] : m;
	return m;
#endif
}
RT_HD float warp_sum(const float (&v)[RT_NL]) {
#if defined(__CUDACC__)
	float s = v[0];
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	return s;
#else
	float s = 0.0f;
	for (int l = 0; l < 32; l++) s += v[l];
	return s;
#endif
}
RT_HD float warp_get(const float (&v)[RT_NL], int src) {
#if defined(__CUDACC__)
	return __shfl_sync(0xffffffffu, v[0], src);
#else
	return v[src];
#endif
}
RT_HD int warp_get(const int (&v)[RT_NL], int src) {
#if defined(__CUDACC__)
	return __shfl_sync(0xffffffffu, v[0], src);
#else
	return v[src];
#endif
}
RT_HD void warp_sync() {
#if defined(__CUDACC__)
	__syncwarp();
#endif
}
RT_HD int popc32(unsigned m) {
#if defined(__CUDACC__)
	return __popc(m);
#else
	return __builtin_popcount(m);
#endif
}
RT_HD int clz32(unsigned m) {  // m != 0
#if defined(__CUDACC__)
	return __clz((int)m);
#else
	return __builtin_clz(m);
#endif
}
// vote of a predicate over the warp; the host build (and ray-by-ray device callers, which pass LOCKSTEP =
// false and never reach it) see a single lane
RT_HD unsigned lane_vote(bool p) {
#if defined(__CUDACC__)
	return __ballot_sync(0xffffffffu, p);
#else
	return p ? 1u : 0u;
#endif
}
RT_HD int ffs32(unsigned m) {  // index of the lowest set bit, m != 0
#if defined(__CUDACC__)
	return __ffs((int)m) - 1;
#else
	return __builtin_ctz(m);
#endif
}

// Warp-uniform bounds of the packet.
struct RtPacket {
	float ox, oy, oz;      // the shared origin
	float ax, ay, az;      // unit axis of the bounding cone
	float sin_h, cos2_h;   // half-angle of the cone (inflated)
	float inv_lo[3];       // min over the rays of |1/d_k|
	float inv_hi[3];       // max over the rays of |1/d_k|
	int neg;               // bit k: d_k < 0 (for every ray of the class)
	float err_l;
};

// One camera ray of the packet (float copy; the float64 direction is re-derived from the scan tables
// when a candidate has to be confirmed).
struct RtPRay {
	float dx, dy, dz, inv_a;
};

// Conservative "may any ray of the packet pierce the cube [lo, lo+h]^3" by slab intervals.  With a
// shared origin the entry/exit parameters of axis k are u * |1/d_k| with u the signed distance to the
// near / far plane along the direction of travel, so their extremes over the packet are u times the
// extremes of |1/d_k|.  (0 * inf = NaN is dropped by fminf/fmaxf.)
RT_HD bool packet_pierces_cube(const RtPacket& P, float lox, float loy, float loz, float h) {
	const float lo[3] = {lox, loy, loz};
	const float o[3] = {P.ox, P.oy, P.oz};
	float tnear = 0.0f, tfar = INFINITY;
#pragma unroll
	for (int k = 0; k < 3; k++) {
		const float un = (((P.neg >> k) & 1) ? o[k] - (lo[k] + h) : lo[k] - o[k]) - P.err_l;
		const float uf = un + h + 2.0f * P.err_l;
		const float tn = un * (un >= 0.0f ? P.inv_lo[k] : P.inv_hi[k]);
		const float tf = uf * (uf >= 0.0f ? P.inv_hi[k] : P.inv_lo[k]);
		tnear = fmaxf(tnear, tn);
		tfar = fminf(tfar, tf);
	}
	return tnear <= tfar * 1.00001f;
}

// Conservative "may the (bounding) sphere of this origin-relative record meet the packet's cone".
// Distance from the centre to the cone's lateral surface is perp*cos - t*sin (t along the axis, perp
// across it) wherever that surface is the nearest part of the cone, and never more than the true
// distance elsewhere.
RT_HD bool packet_meets_record(const RtPacket& P, const RtF4& g) {
	if (g.w == INFINITY) return true;
	const float t = g.x * P.ax + g.y * P.ay + g.z * P.az;
	const float px = g.x - t * P.ax, py = g.y - t * P.ay, pz = g.z - t * P.az;
	const float perp2 = px * px + py * py + pz * pz;
	const float rhs = fabsf(g.w) + t * P.sin_h;
	return rhs >= 0.0f && perp2 * P.cos2_h <= rhs * rhs * 1.00001f;
}

// bit k of the result = bit (k ^ x) of m (8-bit mask): octant mask -> visit-order mask
RT_HD unsigned xor_permute8(unsigned m, int x) {
	if (x & 1) m = ((m & 0x55u) << 1) | ((m & 0xaau) >> 1);
	if (x & 2) m = ((m & 0x33u) << 2) | ((m & 0xccu) >> 2);
	if (x & 4) m = ((m & 0x0fu) << 4) | ((m & 0xf0u) >> 4);
	return m;
}

RT_HD void prefetch_l1(const void* p) {
#if defined(__CUDACC__)
	asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
	(void)p;
#endif
}
RT_HD RtPNode ld(const RtPNode* p) {
#if defined(__CUDACC__)
	const float4 a = __ldg(reinterpret_cast<const float4*>(p));
	const int4 b = __ldg(reinterpret_cast<const int4*>(p) + 1);
	return RtPNode{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#else
	return *p;
#endif
}

RT_HD RtWNode ld(const RtWNode* p) {
#if defined(__CUDACC__)
	const float4 a = __ldg(reinterpret_cast<const float4*>(p));
	const int4 b = __ldg(reinterpret_cast<const int4*>(p) + 1);
	const float4 c = __ldg(reinterpret_cast<const float4*>(p) + 2);
	const int4 d = __ldg(reinterpret_cast<const int4*>(p) + 3);
	RtWNode n;
	n.x = a.x; n.y = a.y; n.z = a.z; n.size = a.w;
	n.child_base = b.x; n.child_mask = b.y; n.up = b.z; n._pad = b.w;
	n.lo[0] = c.x; n.lo[1] = c.y; n.lo[2] = c.z; n.hi[0] = c.w;
	n.hi[1] = __int_as_float(d.x); n.hi[2] = __int_as_float(d.y); n.a = d.z; n.b = d.w;
	return n;
#else
	return *p;
#endif
}

// Geometry of a packet: RT_PPL rays per lane, ray j of lane `lane` is the pixel (lane & 7, lane >> 3) of
// the j-th 8x4 sub-patch; sub-patch q of a 16x16 tile sits at (8 * (q & 1), 4 * (q >> 1)).
struct RtPatch {
	int x0, y0;       // tile origin in the frame
	int sub0;         // first sub-patch of this packet within the tile
	size_t out_base;  // tile-compact output: index of the tile's first pixel
};
RT_HD int patch_x(const RtPatch& pt, int lane, int j) { return pt.x0 + (((pt.sub0 + j) & 1) << 3) + (lane & 7); }
RT_HD int patch_y(const RtPatch& pt, int lane, int j) { return pt.y0 + (((pt.sub0 + j) >> 1) << 2) + (lane >> 3); }

// rotate_vectors (src/math/vector.ts:318-323): (a, b) <- (a*c + b*s, a*-s + b*c), every product and sum rounded
RT_HD void rotate_pair(double* a, double* b, double c, double s) {
#pragma unroll
	for (int k = 0; k < 3; k++) {
		const double x = xadd(xmul(a[k], c), xmul(b[k], s));
		const double y = xadd(xmul(a[k], -s), xmul(b[k], c));
		a[k] = x;
		b[k] = y;
	}
}

// Ray generation for ONE component of one half of row y, exactly as iter_h of get_dir_for_each_pixel scans it
// (src/view/camera.ts:214-226,240-249): starting from the row's fr (row_fr[y]) and norm_lf, the right half
// (half == 0) yields x = width>>1, ..., width-1 rotating AFTER every yield by rot_scan_h_v, the left half
// (half == 1) rotates by the counter-clockwise rotation FIRST and yields x = (width>>1)-1, ..., 0.  rotate_vectors
// never mixes components, so (fr[k], lf[k]) is a recurrence of its own: one lane of the frame-setup kernel.  The
// state at every RT_RAYGEN_STRIDE-th yield goes to the checkpoint table (F.ray_ck).
RT_HD void raygen_half_row_component(const RtFrame& F, int y, int half, int comp, double* ck) {
	const RtD4 r = ld(F.row_fr + y);
	double a = comp == 0 ? r.x : comp == 1 ? r.y : r.z, b = F.lf[comp];
	const int x0 = F.width >> 1;
	const double c = F.scan_cos, s = half ? -F.scan_sin : F.scan_sin;
	const int n = half ? x0 : F.width - x0;
	if (half) {
		const double x = xadd(xmul(a, c), xmul(b, s)), yy = xadd(xmul(a, -s), xmul(b, c));
		a = x;
		b = yy;
	}
	double* out = ck + ((size_t)(y * 2 + half) * F.ray_ckh) * 6 + comp * 2;
	for (int i = 0; i < n; i++) {
		if ((i & (RT_RAYGEN_STRIDE - 1)) == 0) {
			out[0] = a;
			out[1] = b;
			out += 6;
		}
		const double na = xadd(xmul(a, c), xmul(b, s)), nb = xadd(xmul(a, -s), xmul(b, c));
		a = na;
		b = nb;
	}
}

// Camera direction of pixel (x,y): get_dir_for_each_pixel (src/view/camera.ts:207-250): the generator's state at
// the checkpoint before the pixel, iterated the remaining steps with the generator's own operations.
RT_HD void pixel_dir(const RtFrame& F, int x, int y, double* dir) {
	const int x0 = F.width >> 1;
	const int half = x < x0 ? 1 : 0;
	const int i = half ? x0 - 1 - x : x - x0;
	const RtD2* p = reinterpret_cast<const RtD2*>(F.ray_ck + ((size_t)(y * 2 + half) * F.ray_ckh + (i / RT_RAYGEN_STRIDE)) * 6);
	const RtD2 cx = ld_stream(p), cy = ld_stream(p + 1), cz = ld_stream(p + 2);
	double fr[3] = {cx.x, cy.x, cz.x}, lf[3] = {cx.y, cy.y, cz.y};
	const double s = half ? -F.scan_sin : F.scan_sin;
	for (int k = i & (RT_RAYGEN_STRIDE - 1); k > 0; k--) rotate_pair(fr, lf, F.scan_cos, s);
	dir[0] = fr[0]; dir[1] = fr[1]; dir[2] = fr[2];
}

// pixels outside the frame ride along with the direction of the nearest pixel inside it
RT_HD void pixel_dir_clamped(const RtFrame& F, int x, int y, double* dir) {
	pixel_dir(F, x < F.width ? x : F.width - 1, y < F.height ? y : F.height - 1, dir);
}

// float64 confirmation of a float32 candidate for the camera ray of pixel (x,y): 0 = collision_info is null,
// 1 = hit, 2 = hit whose normal does not face the ray (the guard of src/raytracer.ts:200-203 will end the path)
// A hit only counts if the reference's walker visits the entity's node for this ray.  The packet walk and the ordered
// walk visit a SUPERSET of the walker's nodes (conservative pierce tests), which cannot change a first hit as long as a
// ray that hits an entity really runs through its node's cube.  It does - except when it merely TOUCHES the cube: a
// corner or an edge (the float64 parameter interval of the cube has zero length: a camera standing on a corner of a
// box that fills its cell, a 45-degree ray through a lattice corner), or a ray that runs inside one of the cube's
// face planes.  Whether the walker visits such a node is decided by its half-open cells and its tie order (x, then
// y, then z), not by geometry, and the float64 hit test of an entity that touches the same corner says "hit, t = 0".
// Such a hit is not trusted: the segment is searched again by the float64 restatement of the walker itself
// (walk_and_scan64).  Generic scenes never get here (the coincidences have to be exact in float64); scenes built on a
// dyadic lattice with the camera on a lattice point - the demo pose (0.5, 0.5, 0.5) is one - do.
RT_COLD bool hit_only_touches_its_cell(const RtDevScene& S, int slot, const double* o, const double* d) {
	const RtD4 g = ld(S.node_geom64 + ld(S.slot_node + slot));
	const double lo[3] = {g.x, g.y, g.z};
	double c[3];
#pragma unroll
	for (int k = 0; k < 3; k++) {
		c[k] = xadd(lo[k], xmul(g.w, 0.5));
		if (d[k] == 0.0 && (o[k] == lo[k] || o[k] == xadd(lo[k], g.w))) return true;  // the ray runs inside a face plane
	}
	double u1, u2;
	int i1, i2;
	if (!exact_box_params(c, g.w, o, d, u1, u2, i1, i2)) return true;  // (the line misses the cube altogether: roundings only)
	// (not only an exactly empty interval: the walker decides by ITS arithmetic - exit points, octant_adj_pos - and
	// that can disagree with this slab test when the overlap is within rounding.  1e-9 of the parameter scale is far
	// above rounding and far below any overlap that matters: a generic ray is re-searched about once in a million hits.)
	const double lo_t = u1 > 0.0 ? u1 : 0.0;
	return !(u2 > lo_t + 1e-9 * (1.0 + (u2 < 0.0 ? -u2 : u2)));
}

// The same question in float32, answered "certainly not" for nearly every hit: the ray's overlap with the cell is
// longer than anything float32 rounding (`slack`: the walks' own bound on a parameter's error) or the float64
// threshold above could take away.  Only the others pay for the float64 test.
RT_HD bool hit_may_only_touch_its_cell(const RtDevScene& S, int slot, const RtRayF& r, float slack) {
	const RtF4 g = ld(S.node_geom + ld(S.slot_node + slot));
	float ta = (g.x - r.ox) * r.ix, tb = (g.x + g.w - r.ox) * r.ix;
	float tmin = fminf(ta, tb), tmax = fmaxf(ta, tb);
	ta = (g.y - r.oy) * r.iy; tb = (g.y + g.w - r.oy) * r.iy;
	tmin = fmaxf(tmin, fminf(ta, tb)); tmax = fminf(tmax, fmaxf(ta, tb));
	ta = (g.z - r.oz) * r.iz; tb = (g.z + g.w - r.oz) * r.iz;
	tmin = fmaxf(tmin, fminf(ta, tb)); tmax = fminf(tmax, fmaxf(ta, tb));
	return !(tmax - fmaxf(tmin, 0.0f) > 4.0f * slack + 1e-6f * (1.0f + fabsf(tmax)));  // (NaN from 0 * inf: "may")
}

// 0: no hit; 1: hit; 2: hit under an acute angle (:200-203)
RT_COLD int packet_confirm(const RtDevScene& S, const RtFrame& F, const double* dir, int slot) {
	RtCollision col;
	const RtD4 g64 = ld(S.slot_geom64 + slot);
	const bool hit = ld(S.slot_geom + slot).w > 0.0f ? exact_sphere(g64, F.pos, dir, col) : exact_box(g64, F.pos, dir, col);
	if (!hit) return 0;
	return dot3(dir, col.normal) >= 0 ? 2 : 1;
}
// (a function of its own, so that the confirmation above stays what the packet stage's hot loop was tuned around)
RT_COLD bool packet_hit_only_touches(const RtDevScene& S, const RtFrame& F, const double* dir, int slot) {
	const RtRayF r = make_ray_f(F.pos, dir);
	const float slack = S.err_l * fminf(fmaxf(fabsf(r.ix), fmaxf(fabsf(r.iy), fabsf(r.iz))), 1e7f);
	return hit_may_only_touch_its_cell(S, slot, r, slack) && hit_only_touches_its_cell(S, slot, F.pos, dir);
}

// Result of the primary search for one camera ray: the first-hit slot (bits 0..29) and RT_HIT_ACUTE, or
// RT_HIT_NONE (miss: sky), or RT_SLOT_UNKNOWN (the ray could not take part in a lock-step walk).
#define RT_HIT_NONE (-1)
#define RT_HIT_ACUTE 0x40000000
#define RT_HIT_SLOT_MASK 0x3fffffff
// stack entries: bit 8 of child_mask = the mask has been narrowed to the children the packet may pierce
#define RT_PNODE_TESTED 0x100

// One lock-step walk for the rays of one direction-sign class: bit j of open[l] = ray j of lane l is still
// searching.  `rays` is the packet's ray table, [PPL][32] (shared memory on the device).
template <int PPL>
RT_HD void packet_walk_class(const RtDevScene& S, const RtFrame& F, const RtPatch& pt, const RtPRay* rays, const double* dirs,
                             const RtPacket& P, unsigned (&open)[RT_NL], RtPNode* stack, int (&hit)[RT_NL][PPL],
                             bool& overflow) {
	int sp = 0;
	// stores the children of `nd` whose octant is in the 8-bit mask m (already narrowed to what the packet may
	// pierce) on the stack, last-visited first
	auto push_mask = [&](const RtPNode& nd, unsigned m) {
		if (!m) return;
		if (sp + 8 > RT_PACKET_STACK(PPL)) { overflow = true; return; }
		const unsigned mk = xor_permute8(m, P.neg);  // bit = visit key
		RT_LANES(l, lane) {
			(void)l;
			const int o = lane & 7;
			if (lane < 8 && ((m >> o) & 1u)) {
				const int key = o ^ P.neg;
				// The stack holds the children's RECORDS, not their indices: the (up to 8) records are
				// consecutive in memory (breadth-first numbering), so this is one coalesced fetch whose
				// latency is paid once per push instead of once per pop; and since the slots follow the same
				// order, the children's lists are neighbours too and can be prefetched right here.
				const RtPNode rec = ld(S.node_pk + nd.child_base + popc32(nd.child_mask & ((1u << o) - 1u)));
				if (rec.list_cnt > 0) prefetch_l1(F.prim_geom + rec.list_off);
				stack[sp + popc32(mk >> (key + 1))] = rec;  // smaller key = popped earlier = higher on the stack
			}
		}
		sp += popc32(m);
		warp_sync();
	};
	// children of an origin-chain node: only the octants a ray can still reach from octant after_oct (the
	// origin's cell / the chain child below), pierce test on lanes 0..7
	auto push_chain_children = [&](const RtPNode& nd, int after_oct) {
		if (!nd.child_mask) return;
		bool ok[RT_NL];
		const float h = nd.size * 0.5f;
		const int want = after_oct ^ P.neg;
		RT_LANES(l, lane) {
			const int o = lane & 7;
			const int key = o ^ P.neg;
			ok[l] = lane < 8 && ((nd.child_mask >> o) & 1) && (key & want) == want && o != after_oct &&
			        packet_pierces_cube(P, nd.x + ((o & 1) ? h : 0.0f), nd.y + ((o & 2) ? h : 0.0f),
			                            nd.z + ((o & 4) ? h : 0.0f), h);
		}
		push_mask(nd, warp_ballot(ok) & 0xffu);
	};
	// The pierce test of the (up to) 4 topmost stack entries at once: lane group g = lane >> 3 takes entry
	// sp-1-g, lane & 7 is the octant - 32 cube tests for the price of one, and the siblings below the top are
	// popped next anyway.  The entry's child_mask is narrowed in place and marked RT_PNODE_TESTED.
	auto test_top = [&]() {
		bool ok[RT_NL];
		RT_LANES(l, lane) {
			const int e = sp - 1 - (lane >> 3);
			ok[l] = false;
			if (e >= 0) {
				const RtPNode& nd = stack[e];
				const int o = lane & 7;
				if (!(nd.child_mask & RT_PNODE_TESTED) && ((nd.child_mask >> o) & 1)) {
					const float h = nd.size * 0.5f;
					ok[l] = packet_pierces_cube(P, nd.x + ((o & 1) ? h : 0.0f), nd.y + ((o & 2) ? h : 0.0f),
					                            nd.z + ((o & 4) ? h : 0.0f), h);
				}
			}
		}
		const unsigned m = warp_ballot(ok);
		warp_sync();
		RT_LANES(l, lane) {
			(void)l;
			const int e = sp - 1 - (lane >> 3);
			if ((lane & 7) == 0 && e >= 0 && !(stack[e].child_mask & RT_PNODE_TESTED)) {
				// (the full child_mask stays in bits 16..23: push_mask needs it to locate the records)
				const int full = stack[e].child_mask & 0xff;
				stack[e].child_mask = (int)((m >> (lane & 24)) & 0xffu) | RT_PNODE_TESTED | (full << 16);
			}
		}
		warp_sync();
	};
	// scans slots [beg,end) in list order; returns true when every ray has its hit
	auto scan = [&](int beg, int end) -> bool {
		for (int base = beg; base < end; base += 32) {
			bool pass[RT_NL];
			RT_LANES(l, lane) {
				const int s = base + lane;
				pass[l] = s < end && packet_meets_record(P, ld(F.prim_geom + s));
			}
			unsigned m = warp_ballot(pass);
			if (!m) continue;
			while (m) {
				const int slot = base + ffs32(m);
				m &= m - 1;
				const RtF4 g = ld(F.prim_geom + slot);
				const bool quick = g.w > 0.0f && g.w != INFINITY;
				const float rr2 = g.w * g.w;
				RT_LANES(l, lane) {
#pragma unroll
					for (int j = 0; j < PPL; j++) {
						if (!((open[l] >> j) & 1u)) continue;
						const RtPRay q = rays[j * 32 + lane];
						bool cand;
						if (quick) {  // candidate_rel
							const float tca = g.x * q.dx + g.y * q.dy + g.z * q.dz;
							const float sc = tca * q.inv_a;
							const float lx = g.x - sc * q.dx, ly = g.y - sc * q.dy, lz = g.z - sc * q.dz;
							cand = lx * lx + ly * ly + lz * lz <= rr2 && tca >= 0.0f;
						} else {
							RtRayF rf;
							rf.ox = P.ox; rf.oy = P.oy; rf.oz = P.oz;
							rf.dx = q.dx; rf.dy = q.dy; rf.dz = q.dz;
							rf.ix = 1.0f / q.dx; rf.iy = 1.0f / q.dy; rf.iz = 1.0f / q.dz;
							rf.inv_a = q.inv_a;
							cand = candidate(ld(S.slot_geom + slot), rf, S.err_l);
						}
						if (cand) {
							const int c = packet_confirm(S, F, dirs + (size_t)(j * 32 + lane) * 3, slot);
							if (c) {
								hit[l][j] = slot | (c == 2 ? RT_HIT_ACUTE : 0);
								open[l] &= ~(1u << j);
								// camera on a cell plane of the octree (F.tie_checks): a hit whose ray only touches the entity's cell is
								// not decided here - the bounce stage searches the ray again, alone and exactly (segment_found)
								if (F.tie_checks && packet_hit_only_touches(S, F, dirs + (size_t)(j * 32 + lane) * 3, slot)) hit[l][j] = RT_SLOT_UNKNOWN;
							}
						}
					}
				}
			}
			bool any[RT_NL];
			RT_LANES(l, lane) { (void)lane; any[l] = open[l] != 0u; }
			if (warp_ballot(any) == 0u) return true;
		}
		return false;
	};

	for (int k = 0; k < F.chain_levels; k++) {
		push_chain_children(ld(S.node_pk + F.chain_node[k]), F.chain_oct[k]);
		while (sp > 0 && !overflow) {
			if (!(stack[sp - 1].child_mask & RT_PNODE_TESTED)) test_top();
			RtPNode nd = stack[--sp];
			warp_sync();
			const unsigned pierced = (unsigned)nd.child_mask & 0xffu;
			nd.child_mask = (nd.child_mask >> 16) & 0xff;
			if (nd.list_cnt > 0 && scan(nd.list_off, nd.list_off + nd.list_cnt)) return;
			push_mask(nd, pierced);
		}
		if (overflow) return;
		if (F.chain_end[k] > F.chain_beg[k] && scan(F.chain_beg[k], F.chain_end[k])) return;
	}
}

// The directions of all the packet's pixels, produced cooperatively into `scratch` ([PPL][32][3] doubles: pixel
// `lane` of sub-patch j at scratch[(j * 32 + lane) * 3]).  A lane iterating its own pixel from the checkpoint
// (pixel_dir) repeats what its left neighbour did plus one step, and the FP64 pipe is narrow; here one lane takes
// one COMPONENT of one 8-pixel ROW of a sub-patch - 12 lanes per sub-patch, two sub-patches per round - and walks
// it once, in generator order (outwards from the middle column), leaving every intermediate state behind: a sixth
// of the FP64 instructions.  Same operations in the same order as the generator, hence the same bits.
template <int PPL>
RT_HD void packet_directions(const RtFrame& F, const RtPatch& pt, double* scratch) {
	const int x0 = F.width >> 1;
	for (int round = 0; round < (PPL + 1) / 2; round++) {
		RT_LANES(l, lane) {
			(void)l;
			const int jj = round * 2 + lane / 12, r = (lane % 12) / 3, c = lane % 3;
			if (lane >= 24 || jj >= PPL) continue;
			const int y = patch_y(pt, r * 8, jj) < F.height ? patch_y(pt, r * 8, jj) : F.height - 1;
			const int xa = patch_x(pt, 0, jj), xb = xa + 7 < F.width - 1 ? xa + 7 : F.width - 1;  // columns xa..xb are in the frame
			double* out = scratch + ((size_t)(jj * 32 + r * 8) * 3 + c);
			for (int half = 0; half < 2; half++) {
				// the pixels of this half in generator order: ascending x from the middle on the right, descending on the left
				const int lo = half ? xa : (xa > x0 ? xa : x0), hi = half ? (xb < x0 - 1 ? xb : x0 - 1) : xb;
				if (lo > hi) continue;
				const int i_lo = half ? x0 - 1 - hi : lo - x0, step = half ? -1 : 1;
				const double* ck = F.ray_ck + ((size_t)(y * 2 + half) * F.ray_ckh + i_lo / RT_RAYGEN_STRIDE) * 6 + c * 2;
				const RtD2 ab = ld_stream(reinterpret_cast<const RtD2*>(ck));
				double a = ab.x, b = ab.y;
				const double cs = F.scan_cos, sn = half ? -F.scan_sin : F.scan_sin;
				for (int k = i_lo & (RT_RAYGEN_STRIDE - 1); k > 0; k--) {
					const double na = xadd(xmul(a, cs), xmul(b, sn)), nb = xadd(xmul(a, -sn), xmul(b, cs));
					a = na;
					b = nb;
				}
				for (int x = half ? hi : lo, n = hi - lo + 1; n > 0; n--, x += step) {
					out[(x - xa) * 3] = a;
					const double na = xadd(xmul(a, cs), xmul(b, sn)), nb = xadd(xmul(a, -sn), xmul(b, cs));
					a = na;
					b = nb;
				}
			}
		}
	}
	warp_sync();
}

// First-hit code of every camera ray of the packet in the reference's visit order (RT_HIT_* above).  Rays with
// bit j of skip[l] set (pixels outside the frame) take no part.  `stack` is RT_PACKET_STACK(PPL) records, `dirs` PPL * 32 * 3 doubles and `rays`
// PPL * 32 ray records private to the warp.  A packet whose rays differ in the sign of a direction component
// (it straddles one of the three great circles through the axes) is walked once per sign class (a zero
// component counts as positive: such a ray never crosses a plane of that axis, so either order is its own).
// Rays that cannot take part in a lock-step walk (non-finite direction; node stack overflow; a packet that is
// not narrow) come back as RT_SLOT_UNKNOWN: the bounce stage searches those ray by ray.
template <int PPL>
RT_HD void packet_primary_hits(const RtDevScene& S, const RtFrame& F, const RtPatch& pt, const unsigned (&skip)[RT_NL],
                               RtPNode* stack, RtPRay* rays, double* dirs, int (&hit)[RT_NL][PPL]) {
	unsigned todo[RT_NL], negs[RT_NL];  // bit j: ray j still to do; 3 bits per ray: its direction-sign class
	packet_directions<PPL>(F, pt, dirs);  // float64, kept for the confirmations; the float32 copies go to `rays`
	RT_LANES(l, lane) {
		todo[l] = 0u;
		negs[l] = 0u;
#pragma unroll
		for (int j = 0; j < PPL; j++) {
			double dir[3];
			const int px = patch_x(pt, lane, j);
			double* d = dirs + (size_t)(j * 32 + lane) * 3;
			if (px < F.width) {  // (rows below the frame were produced with the last row's checkpoints)
				dir[0] = d[0]; dir[1] = d[1]; dir[2] = d[2];
			} else {
				pixel_dir_clamped(F, px, patch_y(pt, lane, j), dir);
				d[0] = dir[0]; d[1] = dir[1]; d[2] = dir[2];
			}
			RtPRay q;
			q.dx = (float)dir[0]; q.dy = (float)dir[1]; q.dz = (float)dir[2];
			q.inv_a = 1.0f / (q.dx * q.dx + q.dy * q.dy + q.dz * q.dz);
			rays[j * 32 + lane] = q;
			negs[l] |= (unsigned)((q.dx < 0.0f ? 1 : 0) | (q.dy < 0.0f ? 2 : 0) | (q.dz < 0.0f ? 4 : 0)) << (3 * j);
			const bool skipped = (skip[l] >> j) & 1u;
			const bool finite = q.inv_a > 0.0f && q.inv_a < INFINITY;
			hit[l][j] = !skipped && !finite ? RT_SLOT_UNKNOWN : RT_HIT_NONE;
			if (!skipped && finite) todo[l] |= 1u << j;
		}
	}
	warp_sync();
	while (true) {
		// ---- next sign class: that of the first ray still to do (lowest lane, then lowest j)
		bool any[RT_NL];
		RT_LANES(l, lane) { (void)lane; any[l] = todo[l] != 0u; }
		const unsigned todo_mask = warp_ballot(any);
		if (!todo_mask) break;
		const int first_lane = ffs32(todo_mask);
		RtPacket P;
		{
			int cls[RT_NL];
			RT_LANES(l, lane) { (void)lane; cls[l] = todo[l] ? (int)((negs[l] >> (3 * ffs32(todo[l]))) & 7u) : 0; }
			P.neg = warp_get(cls, first_lane);
		}
		unsigned open[RT_NL];
		float lo[3][RT_NL], hi[3][RT_NL], sx[RT_NL], sy[RT_NL], sz[RT_NL];
		RT_LANES(l, lane) {
			open[l] = 0u;
			lo[0][l] = lo[1][l] = lo[2][l] = INFINITY;
			hi[0][l] = hi[1][l] = hi[2][l] = 0.0f;
			sx[l] = sy[l] = sz[l] = 0.0f;
#pragma unroll
			for (int j = 0; j < PPL; j++) {
				if (!((todo[l] >> j) & 1u) || (int)((negs[l] >> (3 * j)) & 7u) != P.neg) continue;
				open[l] |= 1u << j;
				const RtPRay q = rays[j * 32 + lane];
				const float ix = fabsf(1.0f / q.dx), iy = fabsf(1.0f / q.dy), iz = fabsf(1.0f / q.dz);
				lo[0][l] = fminf(lo[0][l], ix); hi[0][l] = fmaxf(hi[0][l], ix);
				lo[1][l] = fminf(lo[1][l], iy); hi[1][l] = fmaxf(hi[1][l], iy);
				lo[2][l] = fminf(lo[2][l], iz); hi[2][l] = fmaxf(hi[2][l], iz);
				const float k = sqrtf(q.inv_a);  // sum of the unit directions: the cone axis
				sx[l] += q.dx * k; sy[l] += q.dy * k; sz[l] += q.dz * k;
			}
		}
#pragma unroll
		for (int k = 0; k < 3; k++) {
			P.inv_lo[k] = warp_min_pos(lo[k]) * 0.99999f;
			P.inv_hi[k] = warp_max_pos(hi[k]) * 1.00001f;
		}
		{
			const float ax = warp_sum(sx), ay = warp_sum(sy), az = warp_sum(sz);
			const float k = 1.0f / sqrtf(ax * ax + ay * ay + az * az);
			P.ax = ax * k; P.ay = ay * k; P.az = az * k;
			float sn[RT_NL];
			RT_LANES(l, lane) {
				sn[l] = 0.0f;
#pragma unroll
				for (int j = 0; j < PPL; j++) {
					if (!((open[l] >> j) & 1u)) continue;
					const RtPRay q = rays[j * 32 + lane];
					const float kk = sqrtf(q.inv_a);
					const float nx = q.dx * kk, ny = q.dy * kk, nz = q.dz * kk;
					const float t = nx * P.ax + ny * P.ay + nz * P.az;
					const float px = nx - t * P.ax, py = ny - t * P.ay, pz = nz - t * P.az;
					sn[l] = fmaxf(sn[l], px * px + py * py + pz * pz);
				}
			}
			P.sin_h = sqrtf(warp_max_pos(sn)) * 1.0001f + 1e-6f;
			P.cos2_h = 1.0f - P.sin_h * P.sin_h;
			P.ox = (float)F.pos[0]; P.oy = (float)F.pos[1]; P.oz = (float)F.pos[2];
			P.err_l = S.err_l;
		}
		unsigned cls_rays[RT_NL];
		RT_LANES(l, lane) { (void)lane; cls_rays[l] = open[l]; }
		bool overflow = !(P.sin_h < 0.5f);  // not a narrow packet (tiny frames, huge fov): ray by ray
		if (!overflow) packet_walk_class<PPL>(S, F, pt, rays, dirs, P, open, stack, hit, overflow);
		RT_LANES(l, lane) {
			(void)lane;
			todo[l] &= ~cls_rays[l];
			if (overflow) {
#pragma unroll
				for (int j = 0; j < PPL; j++)
					if ((cls_rays[l] >> j) & 1u) hit[l][j] = RT_SLOT_UNKNOWN;
			}
		}
	}
}

// ------------------------------------------------------------------ shading helpers
// uv_map_sphere (src/math/uv_mapping.ts:19-25)
RT_HD void uv_map_sphere(const double* v, double& u, double& w) {
	u = xsub(xadd(xdiv(xdiv(atan2(v[1], v[0]), RT_JS_PI), 2.0), 0.5), RT_JS_EPSILON);
	double s = 0.0;
	s = xadd(s, xmul(v[0], v[0]));
	s = xadd(s, xmul(v[1], v[1]));
	w = xsub(xadd(xdiv(atan2(v[2], xsqrt(s)), RT_JS_PI), 0.5), RT_JS_EPSILON);
}

// Texture.get_color: texture_solid.ts:33-35, texture_image.ts:40-63
RT_HD bool texture_color(const RtDevScene& S, int tex, bool is_sphere, const double* v, double* rgb) {
	const RtTexture* T = S.textures + tex;
	if (!T->image) {
		rgb[0] = T->r; rgb[1] = T->g; rgb[2] = T->b;
		return true;
	}
	double u = 0.0, w = 0.0;  // BoxEntity.map_uv returns [0,0] (entity_box.ts:104-107)
	if (is_sphere) uv_map_sphere(v, u, w);
	if (u < 0 - RT_JS_EPSILON || u > 1 - RT_JS_EPSILON || w < 0 - RT_JS_EPSILON || w > 1 - RT_JS_EPSILON) {
		rgb[0] = rgb[1] = rgb[2] = NAN;
		return false;
	}
	const long long ui = (long long)xmul(u, (double)T->width);   // (u*width) << 0
	const long long vi = (long long)xmul(w, (double)T->height);
	const uint8_t* px = S.texels + (T->texel_off + (unsigned long long)(vi * T->width + ui)) * 3ull;
	rgb[0] = xdiv((double)px[0], 255.0);
	rgb[1] = xdiv((double)px[1], 255.0);
	rgb[2] = xdiv((double)px[2], 255.0);
	return true;
}

// FpLcg (src/math/rng/fp-lcg.ts:20-82); `% 1.0` on non-negative values == x - floor(x), exact
struct RtRng {
	double s1, s2, s3;
	bool seeded;
};
#define RT_LCG_MUL1 (3532205053565347.0 / 3768278866164713.0)
#define RT_LCG_TERM1 (3773467585272041.0 / 4435662911655887.0)
#define RT_LCG_MUL2 (3632519696538149.0 / 4496133748415501.0)
#define RT_LCG_TERM2 (3396159042346757.0 / 4429161683464229.0)
#define RT_LCG_MUL3 (4056279137291581.0 / 4272384783187219.0)
#define RT_LCG_TERM3 (3685311960670787.0 / 3909517015383373.0)
RT_HD double frac1(double x) { return xsub(x, floor(x)); }
RT_HD void rng_seed(RtRng& g, double seed) {
	g.s1 = seed;
	g.s2 = xmul(seed, RT_LCG_MUL3);
	g.s3 = xmul(seed, RT_LCG_MUL2);
	g.seeded = true;
}
RT_HD double rng_next(RtRng& g) {
	const double a = frac1(xadd(xmul(g.s1, RT_LCG_MUL1), RT_LCG_TERM1));
	const double b = frac1(xadd(xmul(g.s2, RT_LCG_MUL2), RT_LCG_TERM2));
	const double c = frac1(xadd(xmul(g.s3, RT_LCG_MUL3), RT_LCG_TERM3));
	g.s1 = xadd(b, c);
	g.s2 = c;
	g.s3 = xadd(a, b);
	return frac1(xadd(xadd(a, b), c));
}

// ------------------------------------------------------------------ the ordered walk (single ray, bounce stage)
// The same order argument as the packet walk, for ONE ray: with the direction signs fixed, the reference
// visits the children a ray pierces in ascending (octant ^ neg), subtrees in pre-order, and the nodes that
// contain the ray origin in post-order.  Unlike the state machine of walk_and_scan (which follows the
// reference's walker step by step and is kept for the counting kernel), this formulation is a stack machine
// with two kinds of step only, so that the 32 independent rays of a warp in the bounce stage run it in
// lock-step (walk_iter) instead of serialising on the branches of 32 different walker states:
//   node step  pop the next octree node (or, with the stack empty, go one level up the origin chain), push
//              the children the ray pierces - last-visited first - and then the root of the node's list BVH;
//   list step  pop one node of the current list's BVH and fetch its two children together (64 bytes): slab
//              tests, then push the inner ones that are hit and test the up to RT_BVH_LEAF entities of the
//              leaves that are hit, keeping the LOWEST hit slot = the first entity in list order the ray hits
//              (src/raytracer.ts:186-195).
// The list BVH sits on top of the octree entries, so a list is finished before the next node is popped, and
// the walk ends at the first list that holds a hit.  A conservative (slack) pierce test can only add nodes,
// which cannot change a first hit (entities lie inside their node's cube).
#if defined(RT_WALK_STATS) && !defined(__CUDA_ARCH__)  // tests/hostsim: step counts of the ordered walk (tools, not product)
extern unsigned long long g_walk_stats[8];  // iterations, node steps, pair steps, leaf tests, confirmations, lists entered
#define RT_STAT(k) (__atomic_fetch_add(&g_walk_stats[k], 1ull, __ATOMIC_RELAXED))
#else
#define RT_STAT(k) ((void)0)
#endif
#if defined(RT_WALK_PROFILE) && defined(__CUDACC__)  // tools/ only: lane occupancy of the lock-step walk's phases
__device__ unsigned long long g_walk_prof[16];
#endif
#if defined(RT_WALK_PROFILE) && defined(__CUDA_ARCH__)
#define RT_PROF_LANES(k, cond)                                                                    \
	do {                                                                                          \
		const unsigned m_ = __ballot_sync(0xffffffffu, (cond));                                    \
		if ((threadIdx.x & 31) == 0 && m_) atomicAdd(&g_walk_prof[k], (unsigned long long)__popc(m_)); \
	} while (0)
#define RT_PROF_COUNT(k)                                                  \
	do {                                                                  \
		if ((threadIdx.x & 31) == 0) atomicAdd(&g_walk_prof[k], 1ull);     \
	} while (0)
#else
#define RT_PROF_LANES(k, cond) ((void)0)
#define RT_PROF_COUNT(k) ((void)0)
#endif
struct RtWalk {
	RtRayF r;
	int neg;          // bit k: d_k < 0
	int sp;
	int floor;        // stack height below the current list's BVH entries
	int in_list;      // 1: the entries above `floor` are BVH nodes of the current list
	int chain_node;   // origin-chain node whose list is returned once the stack is empty
	int chain_oct;    // octant of chain_node the ray leaves (its children before it cannot be reached); < 0: root only
	int chain_listed; // 1: chain_node's list has been scanned, next is its parent
	int chain_up;     // RtWNode.up of chain_node (known once its list has been started)
	int best;         // lowest hit slot of the current list, RT_NO_SLOT while none
	int hit;          // result: first-hit slot, -1 none, RT_WALK_OVERFLOW: the stack was too small for this ray - the
	                  // caller searches it with the reference-order walker instead (segment_found)
	float slack;
	int stack[RT_WALK_STACK];
};
#ifdef RT_TEST_WALK_CAP  // tests/hostsim only: a tiny stack, so that the overflow path (segment_found) is exercised
#define RT_WALK_CAP RT_TEST_WALK_CAP
#else
#define RT_WALK_CAP RT_WALK_STACK
#endif
// One iteration of the walk pushes at most 8 children, a list root and two BVH nodes: the room for that is checked
// ONCE per iteration (walk_iter), not per push - a ray that runs out of stack is never dropped silently, it is
// reported (W.hit == RT_WALK_OVERFLOW -> RT_ERRFLAG_STACK, segment_found) and the render call fails.
#define RT_WALK_PUSHES_PER_ITER 11
#define RT_WALK_OVERFLOW (-3)
#define RT_WALK_LEAF_TAG 0x40000000  // stack entry: a BVH leaf waiting to be tested
// Lanes that hold a hit leaf wait until this many do (lock-step walk), then test their leaves together; 0: every
// leaf is tested in the iteration that found it.  Measured on configs[2] (round 2): 4 -> -1.4 % at 1 spp, -3.5 % at
// 16 spp; 8 -> no gain.
#ifndef RT_LEAF_DEFER
#define RT_LEAF_DEFER 4
#endif
// Every stack entry is a dependent fetch when it is popped; the record is requested when the entry is PUSHED.
// (Measured on configs[2] / [3], round 2: 2-3 % SLOWER with it - the 32 resident warps already cover the fetch, the
// extra instructions do not pay; switched on at run time for the sparse warps of the stage's tail only, the test of
// the switch cost the full warps more than the tail gained.  Off.)
#ifndef RT_WALK_PREFETCH
#define RT_WALK_PREFETCH 0
#endif
static_assert(RT_WALK_CAP > RT_WALK_PUSHES_PER_ITER, "walk stack smaller than one iteration's pushes");
RT_HD void walk_push(RtWalk& W, int v) { W.stack[W.sp++] = v; }

// pushes the children of `nd` the ray may pierce, last-visited first.  In the ray's own frame (axis k
// mirrored when d_k < 0) the half of the cube entered first is "half 0": key bit k of a child says which
// half it is in, so the parameter interval of a child is a static selection among the intervals of the two
// halves per axis.  The eight interval tests are predicates only (no branches); the loop runs once per child
// actually pushed.
RT_HD void walk_push_children(RtWalk& W, const RtWNode& nd, int after_oct, const RtWNode* node_walk) {
	const RtRayF& r = W.r;
	const float h = nd.size * 0.5f;
	float n0[3], f0[3], n1[3], f1[3];  // [near, far] of half 0 and of half 1, per axis
	{
		const float lo[3] = {nd.x, nd.y, nd.z}, o[3] = {r.ox, r.oy, r.oz}, inv[3] = {r.ix, r.iy, r.iz};
#pragma unroll
		for (int k = 0; k < 3; k++) {
			const bool ng = (W.neg >> k) & 1;
			const float te = ((ng ? lo[k] + nd.size : lo[k]) - o[k]) * inv[k];  // entry plane
			const float tm = (lo[k] + h - o[k]) * inv[k];                        // mid plane
			const float tx = ((ng ? lo[k] : lo[k] + nd.size) - o[k]) * inv[k];  // exit plane
			// sorted pairs: robust against zero components of either sign (+-inf parameters), and NaN from
			// 0 * inf is dropped by fminf / fmaxf (the half then keeps the other bound: conservative)
			n0[k] = fminf(te, tm); f0[k] = fmaxf(te, tm);
			n1[k] = fminf(tm, tx); f1[k] = fmaxf(tm, tx);
		}
	}
	// x/y combinations shared by the two z halves
	const float nxy[4] = {fmaxf(n0[0], n0[1]), fmaxf(n1[0], n0[1]), fmaxf(n0[0], n1[1]), fmaxf(n1[0], n1[1])};
	const float fxy[4] = {fminf(f0[0], f0[1]), fminf(f1[0], f0[1]), fminf(f0[0], f1[1]), fminf(f1[0], f1[1])};
	const float nz[2] = {fmaxf(n0[2], 0.0f), fmaxf(n1[2], 0.0f)};
	const float fz[2] = {f0[2], f1[2]};
	unsigned m = 0;
#pragma unroll
	for (int key = 0; key < 8; key++) {
		const float tnear = fmaxf(nxy[key & 3], nz[key >> 2]);
		const float tfar = fminf(fxy[key & 3], fz[key >> 2]);
		m |= (tnear <= tfar * 1.00001f + W.slack ? 1u : 0u) << key;
	}
	m &= xor_permute8((unsigned)nd.child_mask, W.neg);  // bit key = the child of octant key ^ neg exists
	if (after_oct >= 0) {
		// only the octants a ray can still reach from after_oct: keys that are proper supersets of its key
		const int want = after_oct ^ W.neg;
		m &= ((want & 1) ? 0xaau : 0xffu) & ((want & 2) ? 0xccu : 0xffu) & ((want & 4) ? 0xf0u : 0xffu) & ~(1u << want);
	}
	while (m) {
		const int key = 31 - clz32(m);
		m ^= 1u << key;
		const int o = key ^ W.neg;
		const int child = nd.child_base + popc32((unsigned)nd.child_mask & ((1u << o) - 1u));
		walk_push(W, child);
		if (RT_WALK_PREFETCH) prefetch_record(node_walk + child);
	}
}

// (node, octant): node_at_pos of the ray origin, or octant < 0 for "origin outside the root: root only"
RT_HD void walk_begin(const RtDevScene& S, RtWalk& W, int node, int octant) {
	const RtRayF& r = W.r;
	W.neg = (r.dx < 0.0f ? 1 : 0) | (r.dy < 0.0f ? 2 : 0) | (r.dz < 0.0f ? 4 : 0);
	W.sp = 0;
	W.floor = 0;
	W.in_list = 0;
	W.best = RT_NO_SLOT;
	W.hit = -1;
	W.slack = S.err_l * fminf(fmaxf(fabsf(r.ix), fmaxf(fabsf(r.iy), fabsf(r.iz))), 1e7f);
	W.chain_node = node;
	W.chain_oct = octant;
	W.chain_listed = 0;
	W.chain_up = -1;
	if (octant >= 0) walk_push_children(W, ld(S.node_walk + node), octant, S.node_walk);
}

// float64 confirmation without the collision record (the caller recomputes it for the one slot it keeps)
RT_HD bool confirm_hit(const RtDevScene& S, int s, bool is_sphere, const double* o, const double* d) {
	RtCollision col;
	const RtD4 g = ld(S.slot_geom64 + s);
	return is_sphere ? exact_sphere(g, o, d, col) : exact_box(g, o, d, col);
}

// candidate() on the bounding sphere of the entity (the sphere itself, or the sphere around a box): one
// branch-free test for the entries of a leaf
RT_HD bool candidate_bounding(const RtF4& g, const RtRayF& r, float err_l) {
	const float cx = g.x - r.ox, cy = g.y - r.oy, cz = g.z - r.oz;
	const float tca = cx * r.dx + cy * r.dy + cz * r.dz;
	const float sc = tca * r.inv_a;
	const float lx = cx - sc * r.dx, ly = cy - sc * r.dy, lz = cz - sc * r.dz;
	const float l2 = lx * lx + ly * ly + lz * lz;
	const float rr = (g.w > 0.0f ? g.w : g.w * -1.7320526f) + err_l;
	const float rr2 = rr * rr;
	return l2 <= rr2 && (tca >= 0.0f || cx * cx + cy * cy + cz * cz <= rr2);
}

// slab test of a BVH box (conservative by W.slack; NaN from 0 * inf is dropped by fminf / fmaxf)
RT_HD bool walk_hits_box(const RtWalk& W, float lx, float ly, float lz, float hx, float hy, float hz) {
	const RtRayF& r = W.r;
	float ta = (lx - r.ox) * r.ix, tb = (hx - r.ox) * r.ix;
	float tmin = fminf(ta, tb), tmax = fmaxf(ta, tb);
	ta = (ly - r.oy) * r.iy; tb = (hy - r.oy) * r.iy;
	tmin = fmaxf(tmin, fminf(ta, tb)); tmax = fminf(tmax, fmaxf(ta, tb));
	ta = (lz - r.oz) * r.iz; tb = (hz - r.oz) * r.iz;
	tmin = fmaxf(tmin, fminf(ta, tb)); tmax = fminf(tmax, fmaxf(ta, tb));
	return !(tmin > tmax + W.slack || tmax < -W.slack);
}

// The entities of one BVH leaf (padded to RT_BVH_LEAF, ascending slots): the float32 tests of all of them
// first - independent loads, no branches -, then the float64 confirmation of the candidates in slot order.
RT_HD void walk_leaf(const RtDevScene& S, RtWalk& W, int leaf_a, const double* o, const double* d) {
#if RT_BVH_LEAF == 4
	const RtI4 sl = ld(reinterpret_cast<const RtI4*>(S.bvh_slots + leaf_a));
	const int slots[RT_BVH_LEAF] = {sl.x, sl.y, sl.z, sl.w};
#else
	int slots[RT_BVH_LEAF];
#pragma unroll
	for (int k = 0; k < RT_BVH_LEAF; k++) slots[k] = ld(S.bvh_slots + leaf_a + k);
#endif
	unsigned cand = 0, spheres = 0;
#pragma unroll
	for (int k = 0; k < RT_BVH_LEAF; k++) {
		const RtF4 g = ld(S.bvh_geom + leaf_a + k);
		cand |= (slots[k] < W.best && candidate_bounding(g, W.r, S.err_l) ? 1u : 0u) << k;
		spheres |= (g.w > 0.0f ? 1u : 0u) << k;
	}
	while (cand) {
		const int k = ffs32(cand);
		cand &= cand - 1;
		int s = slots[0];
#pragma unroll
		for (int q = 1; q < RT_BVH_LEAF; q++) s = k == q ? slots[q] : s;
		if (s >= W.best) break;
		RT_STAT(4);
		if (confirm_hit(S, s, (spheres >> k) & 1u, o, d)) {
			W.best = s;  // ascending slots: the later candidates of this leaf cannot beat it
			break;
		}
	}
}

// One iteration of the walk for every lane of the warp (`walking`: this lane takes part).  Returns false
// when this lane's walk is over (W.hit = slot or -1).  LOCKSTEP: all 32 lanes of the warp call this together
// (bounce stage) and the phases re-converge the warp between them; without it the caller may be one lane of a
// diverged warp (ray-by-ray kernels) and no warp-wide barrier is used.
template <bool LOCKSTEP>
RT_HD bool walk_iter(const RtDevScene& S, RtWalk& W, const double* o, const double* d, bool walking, int node_batch = 1,
                     int leaf_batch = RT_LEAF_DEFER) {
	if (walking) RT_STAT(0);
	if (walking && W.sp > RT_WALK_CAP - RT_WALK_PUSHES_PER_ITER) {  // no room for this iteration's pushes
		W.hit = RT_WALK_OVERFLOW;
		walking = false;
	}
	// ---- node step.  It is several times the cost of a list step, so in lock-step the lanes that need one
	// wait until `node_batch` of them do (or no lane is inside a list), and then take it together.
	bool node_step = walking && !W.in_list;
	if (LOCKSTEP) {
		const unsigned need = lane_vote(node_step), busy = lane_vote(walking && W.in_list);
		if (busy != 0u && popc32(need) < node_batch) node_step = false;
	}
	// (a) which node: the next one on the stack, or - stack empty - the next move along the origin chain
	int rec_node = -1, after = -1;
	bool push = false, list = false;
	if (node_step) {
		if (W.sp > 0) {
			rec_node = W.stack[--W.sp];
			push = list = true;
		} else if (!W.chain_listed) {
			// the ray leaves the current origin-chain node, which is returned now (post-order)
			W.chain_listed = 1;
			rec_node = W.chain_node;
			list = true;
		} else if (W.chain_oct < 0 || W.chain_up < 0) {
			walking = false;  // root-only mode, or the root has been returned: miss
		} else {
			W.chain_node = rec_node = W.chain_up & RT_WNODE_PARENT_MASK;
			W.chain_oct = after = (int)((unsigned)W.chain_up >> 28);
			W.chain_listed = 0;
			push = true;
		}
	}
	if (LOCKSTEP) {
		warp_sync();
		RT_PROF_COUNT(0);                      // warp iterations
		RT_PROF_LANES(1, walking);             // lanes walking
		RT_PROF_LANES(2, rec_node >= 0);       // lanes taking a node step
		RT_PROF_LANES(3, walking && !W.in_list && rec_node < 0);  // lanes waiting for a node batch
	}
	// (b) its 64-byte record: children, then the root of the list's BVH (box test right here)
	int leaf0 = -1, leaf1 = -1;  // leaves whose entities are to be tested in this iteration
	int nd_leaf = -1;
	(void)leaf1; (void)nd_leaf;
	if (rec_node >= 0) {
		RT_STAT(1);
		const RtWNode nd = ld(S.node_walk + rec_node);
		if (push) walk_push_children(W, nd, after, S.node_walk);
		else W.chain_up = nd.up;
		if (list && nd.b != 0 && walk_hits_box(W, nd.lo[0], nd.lo[1], nd.lo[2], nd.hi[0], nd.hi[1], nd.hi[2])) {
			W.floor = W.sp;
			W.in_list = 1;
			W.best = RT_NO_SLOT;
			if (nd.b < 0) walk_push(W, nd.a);
			else leaf0 = nd_leaf = nd.a;
			if (RT_WALK_PREFETCH) prefetch_record(nd.b < 0 ? (const void*)(S.bvh_nodes + nd.a) : (const void*)(S.bvh_geom + nd.a));
		}
	}
	if (LOCKSTEP) warp_sync();
#if RT_LEAF_DEFER > 0
	// ---- list step.  Leaves that are hit are not tested on the spot - that phase would run for the one or two lanes
	// that happen to have found one - but pushed like inner nodes (tagged); a lane whose top entry is a leaf waits
	// until RT_LEAF_DEFER lanes have one (or nothing else can move), and they test their leaves together.
	bool pair_step = false;
	if (nd_leaf >= 0) walk_push(W, nd_leaf | RT_WALK_LEAF_TAG);
	if (walking && W.in_list && W.sp > W.floor && !(W.stack[W.sp - 1] & RT_WALK_LEAF_TAG)) {
		pair_step = true;
		RT_STAT(2);
		const RtBvhNode* np = S.bvh_nodes + W.stack[--W.sp];
		const RtF4 a0 = ld(reinterpret_cast<const RtF4*>(np));
		const RtI4 a1 = ld(reinterpret_cast<const RtI4*>(np) + 1);
		const RtF4 b0 = ld(reinterpret_cast<const RtF4*>(np) + 2);
		const RtI4 b1 = ld(reinterpret_cast<const RtI4*>(np) + 3);
		const bool hit_a = walk_hits_box(W, a0.x, a0.y, a0.z, a0.w, int_as_float(a1.x), int_as_float(a1.y));
		const bool hit_b = walk_hits_box(W, b0.x, b0.y, b0.z, b0.w, int_as_float(b1.x), int_as_float(b1.y));
		// the right sibling first, so that the left one - which holds the lowest slot - is popped first
		if (hit_b && (b1.w > 0 || -(b1.w + 1) < W.best)) {
			walk_push(W, b1.w < 0 ? b1.z : (b1.z | RT_WALK_LEAF_TAG));
			if (RT_WALK_PREFETCH) prefetch_record(b1.w < 0 ? (const void*)(S.bvh_nodes + b1.z) : (const void*)(S.bvh_geom + b1.z));
		}
		if (hit_a && (a1.w > 0 || -(a1.w + 1) < W.best)) {
			walk_push(W, a1.w < 0 ? a1.z : (a1.z | RT_WALK_LEAF_TAG));
			if (RT_WALK_PREFETCH) prefetch_record(a1.w < 0 ? (const void*)(S.bvh_nodes + a1.z) : (const void*)(S.bvh_geom + a1.z));
		}
	}
	if (LOCKSTEP) {
		warp_sync();
		RT_PROF_LANES(4, pair_step);
	}
	// ---- leaf entities
	{
		bool leaf_step = walking && W.in_list && W.sp > W.floor && (W.stack[W.sp - 1] & RT_WALK_LEAF_TAG);
		if (LOCKSTEP) {
			const int n_leaf = popc32(lane_vote(leaf_step));
			const unsigned moved = lane_vote(pair_step || rec_node >= 0);
			if (n_leaf < leaf_batch && moved != 0u) leaf_step = false;
			RT_PROF_LANES(5, leaf_step);
#ifdef RT_WALK_PROFILE
			if (lane_vote(leaf_step)) RT_PROF_COUNT(6);  // iterations with a leaf phase
			if (lane_vote(rec_node >= 0)) RT_PROF_COUNT(7);  // iterations with a node phase
			if (lane_vote(pair_step)) RT_PROF_COUNT(8);
#endif
		}
		if (leaf_step) RT_STAT(3);
		if (leaf_step) walk_leaf(S, W, W.stack[--W.sp] & ~RT_WALK_LEAF_TAG, o, d);
	}
	if (LOCKSTEP) warp_sync();
#else
	// ---- list step: the two children of a BVH node are neighbours in memory and are fetched and tested
	// together (half as many dependent fetches as one node per step)
	if (walking && W.in_list && W.sp > W.floor) {
		const RtBvhNode* np = S.bvh_nodes + W.stack[--W.sp];
		const RtF4 a0 = ld(reinterpret_cast<const RtF4*>(np));
		const RtI4 a1 = ld(reinterpret_cast<const RtI4*>(np) + 1);
		const RtF4 b0 = ld(reinterpret_cast<const RtF4*>(np) + 2);
		const RtI4 b1 = ld(reinterpret_cast<const RtI4*>(np) + 3);
		// x0 = lo.xyz, hi.x ; x1 = hi.y, hi.z (as bits), a, b
		const bool hit_a = walk_hits_box(W, a0.x, a0.y, a0.z, a0.w, int_as_float(a1.x), int_as_float(a1.y));
		const bool hit_b = walk_hits_box(W, b0.x, b0.y, b0.z, b0.w, int_as_float(b1.x), int_as_float(b1.y));
		// the right sibling first, so that the left one - which holds the lowest slot - is popped first
		if (hit_b) {
			if (b1.w < 0) {
				if (-(b1.w + 1) < W.best) walk_push(W, b1.z);  // else: nothing below can beat the hit we have
			} else {
				leaf1 = b1.z;
			}
		}
		if (hit_a) {
			if (a1.w < 0) {
				if (-(a1.w + 1) < W.best) walk_push(W, a1.z);
			} else {
				leaf0 = a1.z;
			}
		}
	}
	if (LOCKSTEP) warp_sync();
	// ---- leaf entities
	if (leaf0 >= 0) walk_leaf(S, W, leaf0, o, d);
	if (LOCKSTEP) warp_sync();
	if (leaf1 >= 0) walk_leaf(S, W, leaf1, o, d);
	if (LOCKSTEP) warp_sync();
#endif
	if (walking && W.in_list && W.sp == W.floor) {  // the list is finished
		RT_STAT(5);
		W.in_list = 0;
		if (W.best != RT_NO_SLOT) {
			W.hit = W.best;
			walking = false;
		}
	}
	return walking;
}

// ------------------------------------------------------------------ Ray.trace (src/raytracer.ts:168-277)
// One path for one pixel of one exposure frame, as an explicit state machine: path_begin(), then
// path_segment() once per ray segment (walker re-seed + traversal to the first hit + the material's
// response) until it returns true.  The bounce stage keeps one RtPath per lane and refills finished lanes
// from the continuation queue between segments; trace_path() below is the plain loop.
struct RtPath {
	double refpoint[3], dir[3], col[3];
	double path_distance;
	int refcount, cur_substance;
	int node, octant;   // walker start: node_at_pos of refpoint
	bool have_node;     // false: refpoint outside the root cube
	bool light_hit, primary;
	int first_entity;   // entity of the first collision, -1 none
	RtRng rng;          // rng.seeded: the path drew from the RNG (a rough surface scattered it)
};

// `dir_in` is the camera direction (un-normalised, as the reference passes it).
RT_HD void path_begin(const RtFrame& F, const double* dir_in, RtPath& P) {
	for (int k = 0; k < 3; k++) {
		P.refpoint[k] = F.pos[k];
		P.dir[k] = dir_in[k];
		P.col[k] = 1.0;
	}
	P.path_distance = 0.0;
	P.refcount = 0;
	P.cur_substance = F.start_substance;
	P.rng.seeded = false;
	P.first_entity = -1;
	P.node = F.start_node;
	P.octant = F.start_octant;
	P.have_node = F.start_node >= 0;
	P.light_hit = false;
	P.primary = true;
}

// The end of Ray.trace after its loop: sky on a miss (:267-271), inverse square law on a light (:273-275).
RT_HD bool path_finish(const RtDevScene& S, const RtFrame& F, RtPath& P, double* out, uint32_t& err) {
	if (!P.light_hit) {  // :267-271, SkySphere.get_color (src/sky/sky_sphere.ts:22-27)
		double sc[3];
		if (!texture_color(S, F.sky_texture, true, P.dir, sc)) err |= RT_ERRFLAG_TEXTURE;
		out[0] = xmul(P.col[0], sc[0]); out[1] = xmul(P.col[1], sc[1]); out[2] = xmul(P.col[2], sc[2]);
		return true;
	}
	// inverse square law :273-275
	const double t = xmul(P.path_distance, F.attenuation);
	const double isl = xdiv(1.0, xadd(RT_JS_EPSILON, xmul(t, t)));
	out[0] = xmul(P.col[0], isl); out[1] = xmul(P.col[1], isl); out[2] = xmul(P.col[2], isl);
	return true;
}

// A segment in three parts, so that the bounce stage can run the search of 32 independent rays in lock-step:
//   segment_begin  walker re-seed; returns RT_SEG_DONE (path ended, colour in `out`), RT_SEG_SLOT (`slot` and
//                  `ci` are known: camera segment found by the primary stage, or searched right here with the
//                  reference-order state machine when W == nullptr), or RT_SEG_WALK (W is set up: the caller
//                  runs walk_iter() until it returns false, then calls segment_found());
//   segment_end    the material's response to the hit (or the sky / light ending); true when the path ended.
// `primary_slot`: the first-hit slot of the camera segment when the primary stage already found it (>= 0), or
// RT_SLOT_UNKNOWN to search.
// The tie machinery of a segment, all of it in two cold functions that take their arguments by value - nothing of the
// hot path's state has its address taken, and with F.tie_checks == 0 (every generic frame) nothing of this runs:
//   exact_research  the walker itself, in float64 (walk_and_scan64): the first-hit slot, or -1;
//   tie_recheck     `slot` is the hit a conservative walk found: if its ray only touches the entity's cell (float32
//                   pre-check, then float64), the segment is searched again exactly; returns the slot that counts.
RT_COLD int exact_research(const RtDevScene& S, double ox, double oy, double oz, double dx, double dy, double dz, int node, int octant) {
	const double o[3] = {ox, oy, oz}, d[3] = {dx, dy, dz};
	RtSearch q;
	q.r = make_ray_f(o, d);
	q.rel = nullptr;
	q.chain_mask = 0xffffffffu;
	q.chain_levels = 0;
	RtCounts cnt = {0, 0, 0, 0, 0, 0};
	RtCollision ci;
	return walk_and_scan64<false>(S, q, node, octant, o, d, ci, cnt);
}
RT_COLD int tie_recheck(const RtDevScene& S, int slot, double ox, double oy, double oz, double dx, double dy, double dz, int node, int octant) {
	const double o[3] = {ox, oy, oz}, d[3] = {dx, dy, dz};
	const RtRayF r = make_ray_f(o, d);
	const float slack = S.err_l * fminf(fmaxf(fabsf(r.ix), fmaxf(fabsf(r.iy), fabsf(r.iz))), 1e7f);
	if (!hit_may_only_touch_its_cell(S, slot, r, slack) || !hit_only_touches_its_cell(S, slot, o, d)) return slot;
	return exact_research(S, ox, oy, oz, dx, dy, dz, node, octant);
}

#define RT_SEG_DONE 0
#define RT_SEG_SLOT 1
#define RT_SEG_WALK 2
// (TIES = false: the bounce / resample kernels built for frames without exact ties (F.tie_checks == 0) - the mere presence
// of the tie branches, never taken, cost those kernels 4 % (same-box A/B); the frames that need them get the TIES build)
template <bool COUNT, bool TIES = true>
RT_HD int segment_begin(const RtDevScene& S, const RtFrame& F, RtPath& P, int primary_slot, double* out, RtCounts& cnt,
                        uint32_t& err, RtWalk* W, int& slot, RtCollision& ci) {
	// walker.set_pos_and_dir -> set_position -> setup_cur_node (src/octree_space.ts:188-205,251-278)
	if (COUNT) cnt.segments++;
	slot = -1;
	if (!P.have_node) {
		// origin outside the root cube: the root alone, if the ray meets its box going forward
		const double c[3] = {xadd(S.root_pos[0], xmul(0.5, S.root_size)), xadd(S.root_pos[1], xmul(0.5, S.root_size)),
		                     xadd(S.root_pos[2], xmul(0.5, S.root_size))};
		double u1, u2;
		int i1, i2;
		if (!exact_box_params(c, S.root_size, P.refpoint, P.dir, u1, u2, i1, i2) || !(u1 >= 0 || u2 >= 0)) {
			path_finish(S, F, P, out, err);
			return RT_SEG_DONE;
		}
		P.node = 0;
		P.octant = -1;
	}
	if (P.primary && primary_slot >= 0) {
		// found by the packet stage: only the collision itself is recomputed (same float64 formula)
		slot = confirm_slot(S, primary_slot, P.refpoint, P.dir, ci) ? primary_slot : -1;
		if (W) W->slack = -1.0f;  // (segment_found: this hit was looked at by packet_confirm)
		return RT_SEG_SLOT;
	}
	RtSearch q;
	q.r = make_ray_f(P.refpoint, P.dir);
	q.rel = nullptr;
	q.chain_mask = 0xffffffffu;
	q.chain_levels = 0;
	if (P.primary && F.prim_geom) q.rel = F.prim_geom;  // camera rays: origin-relative records
	if (W) {
		W->r = q.r;
		if (TIES && F.tie_checks && (P.dir[0] == 0.0 || P.dir[1] == 0.0 || P.dir[2] == 0.0)) {
			// a ray that stays inside an axis plane (the middle row or column of an axis-aligned camera; next to never a
			// bounced ray): where that plane is a cell boundary, which cells it belongs to is the walker's half-open rule,
			// and 0 * inf has no place in the conservative pierce tests - the float64 walker itself takes the segment
			W->slack = -1.0f;  // (segment_found: nothing left to check)
			slot = exact_research(S, P.refpoint[0], P.refpoint[1], P.refpoint[2], P.dir[0], P.dir[1], P.dir[2], P.node, P.octant);
			if (slot >= 0 && !confirm_slot(S, slot, P.refpoint, P.dir, ci)) slot = -1;
			return RT_SEG_SLOT;
		}
		walk_begin(S, *W, P.node, P.octant);
		return RT_SEG_WALK;
	}
	// RT_PRECISION_F64 - and the counting variant always: its counters are the reference's access pattern, and at an
	// exact tie (a ray inside a cell-boundary plane, through a lattice corner) only the float64 steps are the walker's
	if (F.search64 || COUNT) {
		slot = walk_and_scan64<COUNT>(S, q, P.node, P.octant, P.refpoint, P.dir, ci, cnt);
		err |= cnt.errors;
		return RT_SEG_SLOT;
	}
	if (q.rel && P.have_node) {  // lock-step pre-test of the shared origin chain
		q.chain_levels = F.chain_levels;
		q.chain_mask = pretest_chain(F, S, q);
	}
	slot = walk_and_scan<COUNT>(S, q, P.node, P.octant, P.refpoint, P.dir, ci, cnt);
	err |= cnt.errors;
	if (TIES && F.tie_checks && slot >= 0) {
		const int exact = tie_recheck(S, slot, P.refpoint[0], P.refpoint[1], P.refpoint[2], P.dir[0], P.dir[1], P.dir[2], P.node, P.octant);
		if (exact != slot) {
			slot = exact;
			if (slot >= 0 && !confirm_slot(S, slot, P.refpoint, P.dir, ci)) slot = -1;
		}
	}
	return RT_SEG_SLOT;
}

// after the ordered walk: the collision of the slot it found.  A walk that ran out of stack (W.hit ==
// RT_WALK_OVERFLOW; rt_ordered_walk_fits sizes the stack for the tree's depth at upload, so this is a defect, not an
// input) is not answered with a guess: the frame's error flag is raised and the render call fails.
template <bool TIES = true>
RT_HD void segment_found(const RtDevScene& S, const RtFrame& F, const RtPath& P, const RtWalk& W, int& slot, RtCollision& ci, uint32_t& err) {
	slot = W.hit;
	if (slot == RT_WALK_OVERFLOW) {
		err |= RT_ERRFLAG_STACK;
		slot = -1;
	}
	if (TIES && F.tie_checks && slot >= 0 && W.slack >= 0.0f)  // (exact ties: rt_b200.h RT_PARAM_EXACT_TIES; W.slack < 0: already exact)
		slot = tie_recheck(S, slot, P.refpoint[0], P.refpoint[1], P.refpoint[2], P.dir[0], P.dir[1], P.dir[2], P.node, P.octant);
	if (slot >= 0 && !confirm_slot(S, slot, P.refpoint, P.dir, ci)) slot = -1;  // (same formula as in the walk: cannot fail)
}

template <bool COUNT>
RT_HD bool segment_end(const RtDevScene& S, const RtFrame& F, RtPath& P, double pixel_seed, int slot, const RtCollision& ci,
                       double* out, RtCounts& cnt, uint32_t& err) {
	P.primary = false;
	if (slot < 0) return path_finish(S, F, P, out, err);  // miss: sky
	const RtI4 attr = ld(S.slot_attr + slot);
	if (P.first_entity < 0) P.first_entity = attr.x;
	if (dot3(P.dir, ci.normal) >= 0) {  // :200-203
		err |= RT_ERRFLAG_ACUTE;
		out[0] = P.col[0]; out[1] = P.col[1]; out[2] = P.col[2];
		return true;
	}
	const bool is_sphere = (attr.y >> RT_ATTR_TYPE_SHIFT) == 0;
	const RtMaterial m = S.materials[attr.y & RT_ATTR_MAT_MASK];
	P.refcount++;
	{  // SolidMaterial.alter_ray (src/materials/material_solid.ts:30-36)
		const RtD4 g = ld(S.slot_geom64 + slot);
		const double rel[3] = {xsub(ci.point[0], g.x), xsub(ci.point[1], g.y), xsub(ci.point[2], g.z)};
		double tc[3];
		if (!texture_color(S, attr.z, is_sphere, rel, tc)) err |= RT_ERRFLAG_TEXTURE;
		P.col[0] = xmul(P.col[0], tc[0]); P.col[1] = xmul(P.col[1], tc[1]); P.col[2] = xmul(P.col[2], tc[2]);
		if (COUNT) cnt.shades++;
	}
	{
		const double dd[3] = {xsub(ci.point[0], P.refpoint[0]), xsub(ci.point[1], P.refpoint[1]),
		                      xsub(ci.point[2], P.refpoint[2])};
		P.path_distance = xadd(P.path_distance, xsqrt(dot3(dd, dd)));  // :210
	}
	P.refpoint[0] = ci.point[0]; P.refpoint[1] = ci.point[1]; P.refpoint[2] = ci.point[2];  // :212
	if (m.flags & RT_MAT_LIGHT) {  // :215-218
		P.light_hit = true;
		return path_finish(S, F, P, out, err);
	}
	const uint32_t response = m.flags & RT_MAT_RESPONSE_MASK;
	if (response == 0u) {  // REFLECTION :221-237
		if (!(m.flags & RT_MAT_MIRROR)) {
			out[0] = P.col[0]; out[1] = P.col[1]; out[2] = P.col[2];
			return true;
		}
		{  // vector.reflection (src/math/vector.ts:263-268)
			const double k = xmul(-dot3(P.dir, ci.normal), 2.0);
			P.dir[0] = xadd(P.dir[0], xmul(ci.normal[0], k));
			P.dir[1] = xadd(P.dir[1], xmul(ci.normal[1], k));
			P.dir[2] = xadd(P.dir[2], xmul(ci.normal[2], k));
		}
		if (m.roughness > 0.0) {  // scatter_ray :121-133, isotropic_sphere_sample vector_utils.ts:8-14
			if (!P.rng.seeded) rng_seed(P.rng, pixel_seed);
			double rv[3];
			do {
				rv[0] = xsub(xmul(rng_next(P.rng), 2.0), 1.0);
				rv[1] = xsub(xmul(rng_next(P.rng), 2.0), 1.0);
				rv[2] = xsub(xmul(rng_next(P.rng), 2.0), 1.0);
			} while (dot3(rv, rv) > 1);
			if (dot3(rv, ci.normal) < 0) { rv[0] = -rv[0]; rv[1] = -rv[1]; rv[2] = -rv[2]; }
			const double ka = xsub(1.0, m.roughness);
			double rf[3] = {xadd(xmul(P.dir[0], ka), xmul(rv[0], m.roughness)),
			                xadd(xmul(P.dir[1], ka), xmul(rv[1], m.roughness)),
			                xadd(xmul(P.dir[2], ka), xmul(rv[2], m.roughness))};
			const double inv = xdiv(1.0, xsqrt(dot3(rf, rf)));
			P.dir[0] = xmul(rf[0], inv); P.dir[1] = xmul(rf[1], inv); P.dir[2] = xmul(rf[2], inv);
		}
		// move_slightly_forward :158-164
		P.refpoint[0] = xadd(P.refpoint[0], xmul(P.dir[0], 1e-3));
		P.refpoint[1] = xadd(P.refpoint[1], xmul(P.dir[1], 1e-3));
		P.refpoint[2] = xadd(P.refpoint[2], xmul(P.dir[2], 1e-3));
	} else if (response == 1u) {  // TRANSMISSION :238-249
		P.refpoint[0] = xadd(P.refpoint[0], xmul(P.dir[0], 1e-3));
		P.refpoint[1] = xadd(P.refpoint[1], xmul(P.dir[1], 1e-3));
		P.refpoint[2] = xadd(P.refpoint[2], xmul(P.dir[2], 1e-3));
		const int rf_slot = entity_at_pos(S, P.refpoint);
		const int substance = rf_slot >= 0 ? ld(&S.slot_attr[rf_slot].w) : F.default_substance;
		if (substance >= 0) {  // refract_ray :135-150
			const double r_ratio = xdiv(S.substances[P.cur_substance], S.substances[substance]);
			const double r_ratio_sq = xmul(r_ratio, r_ratio);
			const double cosine = dot3(P.dir, ci.normal);
			const double cosine_sq = xmul(cosine, cosine);
			const double ref_sine_sq = xmul(xsub(1.0, cosine_sq), r_ratio_sq);
			if (ref_sine_sq <= 1) {
				const double ref_cosine = xsqrt(xsub(1.0, ref_sine_sq));
				const double k = xsub(ref_cosine, cosine);
				P.dir[0] = xsub(xmul(P.dir[0], r_ratio), xmul(ci.normal[0], k));
				P.dir[1] = xsub(xmul(P.dir[1], r_ratio), xmul(ci.normal[1], k));
				P.dir[2] = xsub(xmul(P.dir[2], r_ratio), xmul(ci.normal[2], k));
			} else {
				const double k = xmul(-dot3(P.dir, ci.normal), 2.0);
				P.dir[0] = xadd(P.dir[0], xmul(ci.normal[0], k));
				P.dir[1] = xadd(P.dir[1], xmul(ci.normal[1], k));
				P.dir[2] = xadd(P.dir[2], xmul(ci.normal[2], k));
			}
			P.cur_substance = substance;
		}
	} else {  // BOTH / default :250-251
		out[0] = P.col[0]; out[1] = P.col[1]; out[2] = P.col[2];
		return true;
	}
	if (P.refcount >= F.refmax) {  // :256-263
		out[0] = out[1] = out[2] = 0.0;
		return true;
	}
	P.have_node = node_at_pos(S, P.refpoint, P.node, P.octant);  // :254 (re-seed from the root)
	return false;
}

// One whole segment on one thread (per-ray kernels, host build).  The counting variant follows the
// reference's walker step by step (its counters are the reference's access pattern); the plain variant uses
// the ordered walk, like the bounce stage.
template <bool COUNT>
RT_HD bool path_segment(const RtDevScene& S, const RtFrame& F, RtPath& P, double pixel_seed, int primary_slot, double* out,
                        RtCounts& cnt, uint32_t& err) {
	int slot;
	RtCollision ci;
	if (!COUNT && S.ordered_ok && !F.search64) {
		RtWalk W;
		const int r = segment_begin<COUNT>(S, F, P, primary_slot, out, cnt, err, &W, slot, ci);
		if (r == RT_SEG_DONE) return true;
		if (r == RT_SEG_WALK) {
			while (walk_iter<false>(S, W, P.refpoint, P.dir, true)) {
			}
			segment_found(S, F, P, W, slot, ci, err);
		}
	} else {
		if (segment_begin<COUNT>(S, F, P, primary_slot, out, cnt, err, nullptr, slot, ci) == RT_SEG_DONE) return true;
	}
	return segment_end<COUNT>(S, F, P, pixel_seed, slot, ci, out, cnt, err);
}

// Returns true when the path drew from the RNG: only such paths can differ from one exposure frame to the
// next (there is no pixel jitter).
template <bool COUNT>
RT_HD bool trace_path(const RtDevScene& S, const RtFrame& F, const double* dir_in, double pixel_seed, int primary_slot,
                      double* out, int& first_entity, RtCounts& cnt, uint32_t& err) {
	RtPath P;
	path_begin(F, dir_in, P);
	while (!path_segment<COUNT>(S, F, P, pixel_seed, primary_slot, out, cnt, err)) {
	}
	first_entity = P.first_entity;
	return P.rng.seeded;
}

// ------------------------------------------------------------------ one pixel, all exposure frames
// Raytracer.trace_frame body (src/raytracer.ts:318-329) + ExposureBuffer.set_color_i
// (src/view/exposure_buffer.ts:77-91) for n_frames consecutive frames.
template <bool COUNT>
RT_HD void render_pixel(const RtDevScene& S, const RtFrame& F, int x, int y, size_t out_index, RtCounts& cnt,
                        uint32_t& err, int primary_slot = RT_SLOT_UNKNOWN) {
	// camera direction: get_dir_for_each_pixel (src/view/camera.ts:207-250), from the ray-generation table
	double dir[3];
	pixel_dir(F, x, y, dir);
	const size_t pix = (size_t)y * F.width + x;
	float* o = F.rgb + out_index * 3;  // frame order, or tile-major when tile-sharded (seed stays per frame pixel)
	float px[3] = {0.f, 0.f, 0.f};
	if (F.frame_first > 0) { px[0] = o[0]; px[1] = o[1]; px[2] = o[2]; }
	int first_entity = -1;
	double c[3];
	bool varies = true;  // until proven otherwise by the first frame
	RtCounts per_frame = {0, 0, 0, 0, 0};
	for (uint32_t f = 0; f < F.n_frames; f++) {
		const uint32_t frame_count = F.frame_first + f;
		if (varies) {
			const double seed = xadd(xadd(F.rng_seed, (double)pix),
			                         xmul(xmul((double)frame_count, (double)F.width), (double)F.height));
			const RtCounts before = cnt;
			varies = trace_path<COUNT>(S, F, dir, seed, primary_slot, c, first_entity, cnt, err);
			if (COUNT && !varies) {
				per_frame.segments = cnt.segments - before.segments; per_frame.nodes = cnt.nodes - before.nodes;
				per_frame.tests = cnt.tests - before.tests; per_frame.shades = cnt.shades - before.shades;
				per_frame.confirms = cnt.confirms - before.confirms;
			}
		} else if (COUNT) {
			// a path that never drew from the RNG is the same path in every frame: the sample is reused, and
			// the counters keep counting what the reference would trace again
			cnt.segments += per_frame.segments; cnt.nodes += per_frame.nodes; cnt.tests += per_frame.tests;
			cnt.shades += per_frame.shades; cnt.confirms += per_frame.confirms;
		}
		const double w = xdiv(1.0, (double)(1u + frame_count));
		const double w1 = xsub(1.0, w);
#pragma unroll
		for (int k = 0; k < 3; k++) px[k] = (float)xadd(xmul(c[k], w), xmul((double)px[k], w1));
	}
	o[0] = px[0]; o[1] = px[1]; o[2] = px[2];
	if (F.first_ids) F.first_ids[out_index] = first_entity;
}

// ------------------------------------------------------------------ primary stage: one 8x4 patch of camera rays
// The first iteration of Ray.trace (src/raytracer.ts:179-263) for a camera ray whose first-hit slot is
// known, when the path ENDS there: miss -> sky (:267-271), acute-normal guard (:200-203), light (:215-218,
// :273-275), non-mirror REFLECTION (:222-225), BOTH/default (:250-251), refmax reached (:256-263).
// Returns false when the path continues (mirror bounce or transmission): the bounce stage traces it.
RT_HD bool primary_terminal(const RtDevScene& S, const RtFrame& F, const double* dir, int slot, double* out,
                            int& first_entity, uint32_t& err) {
	first_entity = -1;
	if (slot < 0) {  // SkySphere.get_color (src/sky/sky_sphere.ts:22-27); colour starts at (1,1,1)
		double sc[3];
		if (!texture_color(S, F.sky_texture, true, dir, sc)) err |= RT_ERRFLAG_TEXTURE;
		out[0] = xmul(1.0, sc[0]); out[1] = xmul(1.0, sc[1]); out[2] = xmul(1.0, sc[2]);
		return true;
	}
	const RtI4 attr = ld(S.slot_attr + slot);
	const RtMaterial m = S.materials[attr.y & RT_ATTR_MAT_MASK];
	const uint32_t response = m.flags & RT_MAT_RESPONSE_MASK;
	const bool bounces = !(m.flags & RT_MAT_LIGHT) && ((response == 0u && (m.flags & RT_MAT_MIRROR)) || response == 1u);
	if (bounces && F.refmax > 1) return false;
	first_entity = attr.x;
	const bool is_sphere = (attr.y >> RT_ATTR_TYPE_SHIFT) == 0;
	const RtD4 g = ld(S.slot_geom64 + slot);
	RtCollision ci;
	const bool hit = is_sphere ? exact_sphere(g, F.pos, dir, ci) : exact_box(g, F.pos, dir, ci);
	if (!hit) return false;  // cannot happen (same formula as the confirmation); let the bounce stage decide
	if (dot3(dir, ci.normal) >= 0) {
		err |= RT_ERRFLAG_ACUTE;
		out[0] = out[1] = out[2] = 1.0;
		return true;
	}
	const double rel[3] = {xsub(ci.point[0], g.x), xsub(ci.point[1], g.y), xsub(ci.point[2], g.z)};
	double tc[3];
	if (!texture_color(S, attr.z, is_sphere, rel, tc)) err |= RT_ERRFLAG_TEXTURE;
	double col[3] = {xmul(1.0, tc[0]), xmul(1.0, tc[1]), xmul(1.0, tc[2])};
	if (m.flags & RT_MAT_LIGHT) {
		const double dd[3] = {xsub(ci.point[0], F.pos[0]), xsub(ci.point[1], F.pos[1]), xsub(ci.point[2], F.pos[2])};
		const double t = xmul(xadd(0.0, xsqrt(dot3(dd, dd))), F.attenuation);
		const double isl = xdiv(1.0, xadd(RT_JS_EPSILON, xmul(t, t)));
		col[0] = xmul(col[0], isl); col[1] = xmul(col[1], isl); col[2] = xmul(col[2], isl);
	} else if (bounces) {  // refcount (1) >= refmax
		col[0] = col[1] = col[2] = 0.0;
	}
	out[0] = col[0]; out[1] = col[1]; out[2] = col[2];
	return true;
}

// ExposureBuffer.set_color_i (src/view/exposure_buffer.ts:77-91) for n_frames frames of the SAME sample
// (a path that ends at its first hit never draws from the RNG, and there is no pixel jitter: SURVEY.md F6):
// the blended pixel in px (not stored yet).
RT_HD void blend_constant_sample(const RtFrame& F, size_t out_index, const double* c, float* px) {
	px[0] = px[1] = px[2] = 0.f;
	if (F.frame_first > 0) {
		const float* o = F.rgb + out_index * 3;
		px[0] = o[0]; px[1] = o[1]; px[2] = o[2];
	}
	if (F.frame_first == 0 && F.n_frames == 1) {
		// the first frame of an exposure: weight 1 on the sample, 0 on the (zero) pixel - c*1 is c and 0*0 is +0 exactly,
		// so the blend is one addition
#pragma unroll
		for (int k = 0; k < 3; k++) px[k] = (float)xadd(c[k], 0.0);
		return;
	}
	for (uint32_t f = 0; f < F.n_frames; f++) {
		const double w = f == 0 ? F.w_first : xdiv(1.0, (double)(1u + F.frame_first + f));
		const double w1 = f == 0 ? F.w1_first : xsub(1.0, w);
#pragma unroll
		for (int k = 0; k < 3; k++) px[k] = (float)xadd(xmul(c[k], w), xmul((double)px[k], w1));
	}
}

// Appends `n` items to the continuation queue; returns the index of the first one.
RT_HD unsigned queue_reserve(unsigned* counter, unsigned n) {
#if defined(__CUDACC__)
	return atomicAdd(counter, n);
#else
	return __atomic_fetch_add(counter, n, __ATOMIC_RELAXED);
#endif
}

// The primary stage for one packet of PPL 8x4 sub-patches (rt_primary_kernel; tests/hostsim runs the same code
// with the lanes as loop iterations): the packet walk finds the first-hit code of every camera ray, then every
// pixel whose path ENDS at that first hit is shaded right here - the first iteration of Ray.trace
// (src/raytracer.ts:179-263) -, blended into the ExposureBuffer and stored; the others (mirror bounce,
// transmission, rays the lock-step walk could not take) go to the continuation queue with one warp-aggregated
// atomic per sub-patch.  The common endings need no float64 geometry at all: a miss is the sky texture's colour, a
// hit on a non-light surface with a SolidTexture is that texture's colour (the confirmation already told whether
// the acute-normal guard of :200-203 fires); lights, image textures and an image sky take primary_terminal().
// `stage`: 96 floats per warp (device: shared memory) through which the 32 pixels of a sub-patch leave as whole
// 16-byte words - full sectors, which is what matters when the frame lives in another GPU's or the host's memory.
template <int PPL>
RT_HD void primary_patch(const RtDevScene& S, const RtFrame& F, const RtPatch& pt, RtPNode* stack, RtPRay* rays, double* dirs,
                         float* stage, uint32_t& err_out) {
	unsigned skip[RT_NL];
	int hit[RT_NL][PPL];
	RT_LANES(l, lane) {
		skip[l] = 0u;
#pragma unroll
		for (int j = 0; j < PPL; j++) {
			const bool valid = patch_x(pt, lane, j) < F.width && patch_y(pt, lane, j) < F.height;
			if (!valid) skip[l] |= 1u << j;
			hit[l][j] = valid ? RT_SLOT_UNKNOWN : RT_HIT_NONE;
		}
	}
	if (F.packet_ok) packet_primary_hits<PPL>(S, F, pt, skip, stack, rays, dirs, hit);
	// the hit codes move from registers into the ray table (the rays are not needed any more), so that the loop
	// over the sub-patches below need not be unrolled
	int* const codes = reinterpret_cast<int*>(rays);
	warp_sync();
	RT_LANES(l, lane) {
#pragma unroll
		for (int j = 0; j < PPL; j++) codes[j * 32 + lane] = hit[l][j];
	}
	warp_sync();
	// ---- the paths that end at their first hit
	const RtTexture* sky = S.textures + F.sky_texture;
	const bool sky_simple = !sky->image;
#pragma unroll 1
	for (int j = 0; j < PPL; j++) {
#if defined(__CUDACC__)
		// A sub-patch of sky only (most of them, on the headline scene), first frame of an exposure, constant sky: every
		// pixel is the same colour - no per-pixel work at all, 24 lanes store the 384 bytes as repeating 16-byte words.
		if (sky_simple && F.frame_first == 0 && !F.queue_dense) {
			const int lane = (int)(threadIdx.x & 31u);
			const bool is_sky = !((skip[0] >> j) & 1u) && codes[j * 32 + lane] == RT_HIT_NONE;
			if (__ballot_sync(0xffffffffu, is_sky) == 0xffffffffu && (F.tile_compact || (F.width & 3) == 0) &&
			    (reinterpret_cast<uintptr_t>(F.rgb) & 15u) == 0) {
				const double c[3] = {xmul(1.0, sky->r), xmul(1.0, sky->g), xmul(1.0, sky->b)};
				float p[3];
				blend_constant_sample(F, 0, c, p);  // (frame_first == 0: the old pixel is not read)
				const int x0 = patch_x(pt, 0, j), y0 = patch_y(pt, 0, j);
				if (lane < 24) {
					const int row = lane / 6, chunk = lane - row * 6, o = (chunk * 4) % 3;
					const size_t first = F.tile_compact ? pt.out_base + (size_t)(((y0 + row) & 15) * 16 + (x0 & 15))
					                                    : (size_t)(y0 + row) * F.width + x0;
					const float a = p[o], b = p[(o + 1) % 3], cc = p[(o + 2) % 3];
					reinterpret_cast<float4*>(F.rgb + first * 3)[chunk] = make_float4(a, b, cc, a);
				}
				if (F.first_ids) {
					const int x = patch_x(pt, lane, j), y = patch_y(pt, lane, j);
					F.first_ids[F.tile_compact ? pt.out_base + (size_t)((y & 15) * 16 + (x & 15)) : (size_t)y * F.width + x] = -1;
				}
				continue;
			}
		}
#endif
		bool done[RT_NL], enqueue[RT_NL];
		float px[RT_NL][3];
		int first_entity[RT_NL];
		size_t out_index[RT_NL];
		RT_LANES(l, lane) {
			const int x = patch_x(pt, lane, j), y = patch_y(pt, lane, j);
			const bool valid = !((skip[l] >> j) & 1u);
			out_index[l] = F.tile_compact ? pt.out_base + (size_t)((y & 15) * 16 + (x & 15)) : (size_t)y * F.width + x;
			done[l] = enqueue[l] = false;
			first_entity[l] = -1;
			px[l][0] = px[l][1] = px[l][2] = 0.f;
			if (!valid) continue;
			const int code = codes[j * 32 + lane];
			double c[3];
			uint32_t err = 0;
			bool simple = false, full = false;
			if (code == RT_SLOT_UNKNOWN) {
				enqueue[l] = true;
			} else if (code < 0) {  // miss: SkySphere.get_color (src/sky/sky_sphere.ts:22-27) times the colour (1,1,1)
				if (sky_simple) {
					c[0] = xmul(1.0, sky->r); c[1] = xmul(1.0, sky->g); c[2] = xmul(1.0, sky->b);
					simple = true;
				} else {
					full = true;
				}
			} else {
				const int slot = code & RT_HIT_SLOT_MASK;
				const RtI4 attr = ld(S.slot_attr + slot);
				const RtMaterial m = S.materials[attr.y & RT_ATTR_MAT_MASK];
				const RtTexture* T = S.textures + attr.z;
				const uint32_t response = m.flags & RT_MAT_RESPONSE_MASK;
				const bool bounces = !(m.flags & RT_MAT_LIGHT) && ((response == 0u && (m.flags & RT_MAT_MIRROR)) || response == 1u);
				if (bounces && F.refmax > 1) {
					enqueue[l] = true;
				} else if (code & RT_HIT_ACUTE) {  // :200-203: the path keeps its colour (1,1,1)
					err |= RT_ERRFLAG_ACUTE;
					c[0] = c[1] = c[2] = 1.0;
					first_entity[l] = attr.x;
					simple = true;
				} else if ((m.flags & RT_MAT_LIGHT) || T->image) {
					full = true;
				} else {  // alter_ray: colour (1,1,1) times the SolidTexture's colour; refmax reached -> black (:256-263)
					c[0] = xmul(1.0, T->r); c[1] = xmul(1.0, T->g); c[2] = xmul(1.0, T->b);
					if (bounces) c[0] = c[1] = c[2] = 0.0;
					first_entity[l] = attr.x;
					simple = true;
				}
			}
			if (full) {
				double dir[3];
				pixel_dir(F, x, y, dir);
				if (primary_terminal(S, F, dir, code < 0 ? -1 : (code & RT_HIT_SLOT_MASK), c, first_entity[l], err)) simple = true;
				else enqueue[l] = true;  // (cannot happen; the bounce stage decides)
			}
			if (simple) {
				blend_constant_sample(F, out_index[l], c, px[l]);
				done[l] = true;
			}
			err_out |= err;
		}
		// ---- colours: the 32 pixels of the sub-patch are 4 rows of 8 pixels = 4 x 96 contiguous bytes
#if defined(__CUDACC__)
		{
			const int lane = (int)(threadIdx.x & 31u);
			const int x0 = patch_x(pt, 0, j), y0 = patch_y(pt, 0, j);
			const bool rows_aligned = F.tile_compact || (F.width & 3) == 0;
			const bool wide = __ballot_sync(0xffffffffu, done[0]) == 0xffffffffu && rows_aligned &&
			                  (reinterpret_cast<uintptr_t>(F.rgb) & 15u) == 0;
			if (wide) {
				stage[lane * 3] = px[0][0]; stage[lane * 3 + 1] = px[0][1]; stage[lane * 3 + 2] = px[0][2];
				__syncwarp();
				if (lane < 24) {
					const int row = lane / 6, chunk = lane - row * 6;
					const size_t first = F.tile_compact ? pt.out_base + (size_t)(((y0 + row) & 15) * 16 + (x0 & 15))
					                                    : (size_t)(y0 + row) * F.width + x0;
					reinterpret_cast<float4*>(F.rgb + first * 3)[chunk] = reinterpret_cast<const float4*>(stage + row * 24)[chunk];
				}
				__syncwarp();
			} else if (done[0]) {
				float* o = F.rgb + out_index[0] * 3;
				o[0] = px[0][0]; o[1] = px[0][1]; o[2] = px[0][2];
			}
			if (done[0] && F.first_ids) F.first_ids[out_index[0]] = first_entity[0];
			if (F.queue_dense) {
				// ordered continuation queue (refmax > 1): every pixel leaves its code in place, and a compaction pass
				// builds the queue in output order, so that the 32 pixels a bounce-stage warp takes - and the warps
				// that run beside it - are neighbours in the frame, whatever order the packets finished in
				if (!((skip[0] >> j) & 1u)) {
					const int code = codes[j * 32 + lane];
					F.queue_dense[out_index[0]] = !enqueue[0] ? RT_NOT_QUEUED : (code == RT_SLOT_UNKNOWN ? RT_SLOT_UNKNOWN : (code & RT_HIT_SLOT_MASK));
				}
				continue;
			}
			const unsigned m = __ballot_sync(0xffffffffu, enqueue[0]);
			if (m) {
				unsigned base = 0;
				if (lane == 0) base = queue_reserve(F.queue_count, (unsigned)__popc(m));
				base = __shfl_sync(0xffffffffu, base, 0);
				if (enqueue[0]) {
					const int code = codes[j * 32 + lane];
					F.queue[base + __popc(m & ((1u << lane) - 1u))] =
					    RtQueueItem{((uint32_t)patch_y(pt, lane, j) << 16) | (uint32_t)patch_x(pt, lane, j),
					                code == RT_SLOT_UNKNOWN ? RT_SLOT_UNKNOWN : (code & RT_HIT_SLOT_MASK)};
				}
			}
		}
#else
		(void)stage;
		RT_LANES(l, lane) {
			if (done[l]) {
				float* o = F.rgb + out_index[l] * 3;
				o[0] = px[l][0]; o[1] = px[l][1]; o[2] = px[l][2];
				if (F.first_ids) F.first_ids[out_index[l]] = first_entity[l];
			}
			if (enqueue[l]) {
				const int code = codes[j * 32 + lane];
				F.queue[queue_reserve(F.queue_count, 1u)] =
				    RtQueueItem{((uint32_t)patch_y(pt, lane, j) << 16) | (uint32_t)patch_x(pt, lane, j),
				                code == RT_SLOT_UNKNOWN ? RT_SLOT_UNKNOWN : (code & RT_HIT_SLOT_MASK)};
			}
		}
#endif
	}
}
