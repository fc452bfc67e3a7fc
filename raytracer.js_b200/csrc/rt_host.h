// rt_host.h — host-side (no CUDA) preparation shared by librt_b200 (rt_b200.cu) and the test-only
// host build of the kernel body (tests/hostsim): validation + packing of rt_scene_desc into the slot
// layout of rt_common.h, the camera scan tables, and the once-per-frame start state.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <exception>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/rt_b200.h"
#include "rt_common.h"

// The packed arrays are written once, in full, by the parallel loops of rt_pack_scene: a std::vector that zero-fills
// them first touches every page of a few hundred MB on ONE thread (half of the packing time at 1 M entities).  RtVec
// is std::vector with default-initialisation: resize() reserves, the loops' own stores touch the pages, in parallel.
template <class T>
struct RtNoInit {
	using value_type = T;
	RtNoInit() = default;
	template <class U> RtNoInit(const RtNoInit<U>&) {}
	T* allocate(size_t n) { return std::allocator<T>().allocate(n); }
	void deallocate(T* p, size_t n) { std::allocator<T>().deallocate(p, n); }
	template <class U, class... Args>
	void construct(U* p, Args&&... args) {
		if constexpr (sizeof...(Args) == 0) ::new ((void*)p) U;  // default-init: nothing written for trivial types
		else ::new ((void*)p) U(std::forward<Args>(args)...);
	}
	template <class U> bool operator==(const RtNoInit<U>&) const { return true; }
	template <class U> bool operator!=(const RtNoInit<U>&) const { return false; }
};
template <class T> using RtVec = std::vector<T, RtNoInit<T>>;

struct RtHostScene {
	RtVec<RtF4> node_geom;
	RtVec<RtD4> node_geom64;
	RtVec<RtI4> node_link;
	RtVec<int> node_child;
	RtVec<RtPNode> node_pk;
	RtVec<RtWNode> node_walk;
	RtVec<int> node_bvh;
	RtVec<RtBvhNode> bvh_nodes;
	RtVec<int> bvh_slots;
	RtVec<RtF4> bvh_geom;
	int max_bvh_depth = 0;  // deepest list-BVH level (root = 0)
	RtVec<RtF4> slot_geom;
	RtVec<RtD4> slot_geom64;
	RtVec<RtI4> slot_attr;
	RtVec<int> slot_node;
	std::vector<RtMaterial> materials;
	RtVec<RtTexture> textures;
	std::vector<double> substances;
	std::vector<uint8_t> texels;
	double root_pos[3] = {0, 0, 0};
	double root_size = 1;
	float err_l = 0;
	bool any_transmission = false;
	int max_depth = 0;  // deepest node level (root = 0)
	std::string f32_refusal;  // non-empty: why this tree cannot be searched in float32 (RT_PRECISION_F64 can)
	bool entities_on_cell_planes = false;  // some entity's bounding cube has a face on a cell plane of the tree: exact ties are possible
};

inline std::string rt_format(const char* fmt, ...);

// Packing a scene is O(nodes + entities) of independent work per node: it runs on the host's cores.
// RT_B200_PACK_THREADS overrides the thread count (default: the hardware's, at most 16); the packed arrays do not
// depend on it.  fn(begin, end) is called for blocks of `grain` indices of [0, n), handed out dynamically.
inline unsigned rt_pack_threads(size_t work) {
	const char* e = getenv("RT_B200_PACK_THREADS");
	int v = e ? atoi(e) : 0;
	if (v <= 0) v = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
	return work < 65536 ? 1u : (unsigned)v;
}

template <class F>
inline void rt_parallel_blocks(size_t n, size_t grain, size_t work, F&& fn) {
	const unsigned T = (unsigned)std::min<size_t>(rt_pack_threads(work), (n + grain - 1) / std::max<size_t>(grain, 1));
	if (T <= 1 || n <= grain) {
		if (n) fn((size_t)0, n);
		return;
	}
	std::atomic<size_t> next{0};
	std::exception_ptr thrown;  // (an allocation failure inside a block: re-thrown on the calling thread, after the join)
	std::mutex thrown_mu;
	auto run = [&] {
		try {
			for (;;) {
				const size_t b = next.fetch_add(grain);
				if (b >= n) return;
				fn(b, std::min(n, b + grain));
			}
		} catch (...) {
			next.store(n);  // the other threads stop at their next block
			std::lock_guard<std::mutex> g(thrown_mu);
			if (!thrown) thrown = std::current_exception();
		}
	};
	std::vector<std::thread> pool;
	for (unsigned t = 1; t < T; t++) pool.emplace_back(run);
	run();
	for (std::thread& t : pool) t.join();
	if (thrown) std::rethrow_exception(thrown);
}

// The error a single thread walking the indices in order would have met first: blocks report (key, status,
// message), the lowest key wins.
struct RtFirstError {
	std::mutex mu;
	std::atomic<bool> any{false};
	uint64_t key = ~0ull;
	rt_status st = RT_OK;
	std::string msg;
	void report(uint64_t k, rt_status s, std::string m) {
		std::lock_guard<std::mutex> g(mu);
		if (k < key) {
			key = k;
			st = s;
			msg = std::move(m);
		}
		any = true;
	}
};

// A deep, mutable copy of an rt_scene_desc: what rt_scene_update edits (the caller's arrays were only borrowed
// for the duration of rt_scene_upload).
struct RtSceneCopy {
	std::vector<double> node_pos, node_size, ent_pos, ent_extent, mat_roughness, tex_color, sub_refractive_index;
	std::vector<int32_t> node_child, node_parent, node_octant, ent_material, ent_texture, ent_substance, tex_width, tex_height;
	std::vector<uint32_t> node_list_off, list_entity;
	std::vector<uint8_t> ent_type, mat_response, mat_light, mat_mirror, tex_kind, tex_loaded, texels;
	std::vector<uint64_t> tex_texel_off;
	bool valid = false;

	// (a 1 M-entity description is ~110 MB in two dozen arrays: copied by a few threads, each taking whole arrays)
	void assign(const rt_scene_desc& d) {
		const size_t N = d.n_nodes, L = d.n_list, E = d.n_entities, M = d.n_materials, T = d.n_textures, S = d.n_substances;
		auto cp = [](auto& v, const auto* p, size_t n) { if (p && n) v.assign(p, p + n); else v.clear(); };
		const std::function<void()> groups[] = {
		    [&] { cp(node_pos, d.node_pos, 3 * N); cp(node_size, d.node_size, N); cp(node_child, d.node_child, 8 * N);
		          cp(node_parent, d.node_parent, N); cp(node_octant, d.node_octant, N); cp(node_list_off, d.node_list_off, N + 1); },
		    [&] { cp(ent_pos, d.ent_pos, 3 * E); cp(ent_type, d.ent_type, E); },
		    [&] { cp(list_entity, d.list_entity, L); cp(ent_extent, d.ent_extent, E); cp(ent_material, d.ent_material, E);
		          cp(ent_texture, d.ent_texture, E); cp(ent_substance, d.ent_substance, E); },
		    [&] { cp(tex_color, d.tex_color, 4 * T); cp(tex_kind, d.tex_kind, T); cp(tex_width, d.tex_width, T); cp(tex_height, d.tex_height, T);
		          cp(tex_loaded, d.tex_loaded, T); cp(tex_texel_off, d.tex_texel_off, T); cp(texels, d.texels, (size_t)d.n_texels * 3);
		          cp(mat_response, d.mat_response, M); cp(mat_light, d.mat_light, M); cp(mat_mirror, d.mat_mirror, M);
		          cp(mat_roughness, d.mat_roughness, M); cp(sub_refractive_index, d.sub_refractive_index, S); },
		};
		const size_t n_groups = sizeof groups / sizeof groups[0];
		rt_parallel_blocks(n_groups, 1, N + L + E + T, [&](size_t g0, size_t g1) {
			for (size_t g = g0; g < g1; g++) groups[g]();
		});
		valid = true;
	}
	rt_scene_desc desc() const {
		rt_scene_desc d;
		memset(&d, 0, sizeof d);
		d.struct_size = (uint32_t)sizeof d;
		auto ptr = [](const auto& v) { return v.empty() ? nullptr : v.data(); };
		d.n_nodes = (uint32_t)node_size.size();
		d.node_pos = ptr(node_pos); d.node_size = ptr(node_size); d.node_child = ptr(node_child); d.node_parent = ptr(node_parent);
		d.node_octant = ptr(node_octant); d.node_list_off = ptr(node_list_off);
		d.n_list = (uint32_t)list_entity.size(); d.list_entity = ptr(list_entity);
		d.n_entities = (uint32_t)ent_extent.size();
		d.ent_type = ptr(ent_type); d.ent_pos = ptr(ent_pos); d.ent_extent = ptr(ent_extent); d.ent_material = ptr(ent_material);
		d.ent_texture = ptr(ent_texture); d.ent_substance = ptr(ent_substance);
		d.n_materials = (uint32_t)mat_roughness.size();
		d.mat_response = ptr(mat_response); d.mat_light = ptr(mat_light); d.mat_mirror = ptr(mat_mirror); d.mat_roughness = ptr(mat_roughness);
		d.n_textures = (uint32_t)tex_kind.size();
		d.tex_kind = ptr(tex_kind); d.tex_color = ptr(tex_color); d.tex_width = ptr(tex_width); d.tex_height = ptr(tex_height);
		d.tex_loaded = ptr(tex_loaded); d.tex_texel_off = ptr(tex_texel_off);
		d.n_texels = texels.size() / 3; d.texels = ptr(texels);
		d.n_substances = (uint32_t)sub_refractive_index.size(); d.sub_refractive_index = ptr(sub_refractive_index);
		return d;
	}
};

// Moving entities the reference's way, on the flat tree: BasicEntity._set_pos (src/entities/entity_basic.ts:38-42)
// followed by add_entity_to_octree (src/octree_entity.ts:174-188), whose Entity.set_octree (src/entity.ts:50-56)
// deletes the entity from the Set of its node and adds it at the END of the Set of the node that now covers it
// (the deepest cube, up to max_in_depth levels below the root, that contains the entity's cubic AABB: the placement
// walk of rt_build.h, same float64 expressions; nodes are created on the way, never removed).  Moves are applied in
// the order given.  false + message where the reference throws TreeOutsideGrowError (max_out_depth = 0); the copy is
// then unchanged.
inline bool rt_scene_move_entities(RtSceneCopy& sc, uint32_t n, const uint32_t* ids, const double* new_pos, uint32_t max_in_depth,
                                   std::string& err) {
	const size_t N0 = sc.node_size.size(), E = sc.ent_extent.size();
	std::vector<double> node_pos = sc.node_pos, node_size = sc.node_size, ent_pos = sc.ent_pos;
	std::vector<int32_t> child = sc.node_child, parent = sc.node_parent, octant = sc.node_octant;
	std::vector<uint32_t> ent_node(E, 0), seq(E, 0);  // seq[e] = 0: still where the old lists have it
	for (size_t i = 0; i < N0; i++)
		for (uint32_t li = sc.node_list_off[i]; li < sc.node_list_off[i + 1]; li++) ent_node[sc.list_entity[li]] = (uint32_t)i;
	struct Placed { uint32_t node, entity, seq; };
	std::vector<Placed> placed;
	for (uint32_t m = 0; m < n; m++) {
		const uint32_t e = ids[m];
		if (e >= E) { err = rt_format("rt_scene_update: entity %u out of range (%zu entities)", e, E); return false; }
		const double ext = sc.ent_extent[e];
		double mn[3];
		for (int k = 0; k < 3; k++) {
			ent_pos[3 * (size_t)e + k] = new_pos[3 * (size_t)m + k];
			mn[k] = sc.ent_type[e] == 0 ? new_pos[3 * (size_t)m + k] - ext * 0.5 : new_pos[3 * (size_t)m + k] - ext / 2;
		}
		auto fits = [&](const double* p, double s) {
			for (int k = 0; k < 3; k++)
				if (!(mn[k] >= p[k] && mn[k] + ext <= p[k] + s)) return false;
			return true;
		};
		bool inside = true;
		for (int k = 0; k < 3; k++) inside = inside && mn[k] >= node_pos[k] && mn[k] < node_pos[k] + node_size[0];
		if (!inside || !fits(node_pos.data(), node_size[0])) {
			err = rt_format("The tree outside-depth limit exceeded (entity %u does not fit the root cube at its new position; max_out_depth is 0)", e);
			return false;
		}
		uint32_t node = 0;
		for (uint32_t depth = 0; depth < max_in_depth; depth++) {
			const double* np = &node_pos[3 * (size_t)node];
			const double ns = node_size[node], k2 = 2.0 / ns, hs = ns / 2;
			int o[3];
			double cp[3];
			for (int k = 0; k < 3; k++) {
				o[k] = (int)((mn[k] - np[k]) * k2);
				cp[k] = np[k] + o[k] * hs;
			}
			if (!fits(cp, hs)) break;
			const int idx = (o[2] << 2) | (o[1] << 1) | o[0];
			int32_t ch = child[(size_t)node * 8 + idx];
			if (ch < 0) {
				ch = (int32_t)node_size.size();
				child[(size_t)node * 8 + idx] = ch;
				node_pos.insert(node_pos.end(), cp, cp + 3);
				node_size.push_back(hs);
				child.insert(child.end(), 8, -1);
				parent.push_back((int32_t)node);
				octant.push_back(idx);
			}
			node = (uint32_t)ch;
		}
		ent_node[e] = node;
		seq[e] = m + 1;
		placed.push_back(Placed{node, e, m + 1});
	}
	// the new lists: what is left of the old ones, in their order, then the arrivals in the order they came
	const size_t N = node_size.size();
	std::vector<uint32_t> off(N + 1, 0);
	for (size_t e = 0; e < E; e++) off[ent_node[e] + 1]++;
	for (size_t i = 0; i < N; i++) off[i + 1] += off[i];
	std::vector<uint32_t> cur(off.begin(), off.end() - 1), list(sc.list_entity.size());
	for (size_t i = 0; i < N0; i++)
		for (uint32_t li = sc.node_list_off[i]; li < sc.node_list_off[i + 1]; li++) {
			const uint32_t e = sc.list_entity[li];
			if (seq[e] == 0) list[cur[i]++] = e;
		}
	for (const Placed& p : placed)
		if (seq[p.entity] == p.seq) list[cur[p.node]++] = p.entity;
	sc.node_pos.swap(node_pos); sc.node_size.swap(node_size); sc.node_child.swap(child); sc.node_parent.swap(parent);
	sc.node_octant.swap(octant); sc.ent_pos.swap(ent_pos); sc.node_list_off.swap(off); sc.list_entity.swap(list);
	return true;
}

inline std::string rt_format(const char* fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	return buf;
}

// Binary BVHs over the entity lists (rt_common.h: RtBvhNode), one per non-empty list.  Median split of the
// centres along the widest axis; boxes are the entities' float64 boxes inflated by err_l and rounded outwards
// to float32.  Every inner node records the lowest slot below it, and its LEFT child is the one that holds
// that slot: a traversal that keeps the lowest hit slot so far skips whole subtrees that cannot beat it, and
// reaches the low slots first.
inline void rt_build_list_bvhs(RtHostScene& hs) {
	const size_t N = hs.node_link.size();
	hs.node_bvh.assign(N, -1);
	hs.max_bvh_depth = 0;
	struct Box { float lo[3], hi[3]; };
	const double infl = (double)hs.err_l;
	auto box_of = [&](int s) {
		const RtD4& g = hs.slot_geom64[s];
		const double h = (hs.slot_geom[s].w > 0.0f ? g.w * 0.5 : g.w / 2) + infl;
		const double c[3] = {g.x, g.y, g.z};
		Box b;
		for (int k = 0; k < 3; k++) {
			b.lo[k] = std::nextafter((float)(c[k] - h), -INFINITY);
			b.hi[k] = std::nextafter((float)(c[k] + h), INFINITY);
		}
		return b;
	};
	// The shape of a list's BVH depends on its length alone (median splits down to RT_BVH_LEAF entries), so every
	// list knows where its records go before any is built, and the lists are built in parallel straight into place:
	// the arrays are the ones a single thread produces, whatever the thread count.
	struct Shape { int nodes, leaves; };
	std::vector<Shape> memo(4096, Shape{0, 0});
	std::function<Shape(int)> shape = [&](int cnt) -> Shape {
		if (cnt <= RT_BVH_LEAF) return Shape{1, 1};
		if ((size_t)cnt < memo.size() && memo[cnt].nodes) return memo[cnt];
		const Shape l = shape(cnt / 2), r = shape(cnt - cnt / 2);
		const Shape t{1 + l.nodes + r.nodes, l.leaves + r.leaves};
		if ((size_t)cnt < memo.size()) memo[cnt] = t;
		return t;
	};
	std::vector<int> leaf_base(N, 0);
	size_t n_bvh = 0, n_leaves = 0;
	for (size_t n = 0; n < N; n++) {
		const int cnt = hs.node_link[n].w;
		if (cnt <= 0) continue;
		const Shape t = shape(cnt);
		hs.node_bvh[n] = (int)n_bvh;
		leaf_base[n] = (int)n_leaves;
		n_bvh += (size_t)t.nodes;
		n_leaves += (size_t)t.leaves;
	}
	// (every record is written by the build below; only the place-holders of a scene without entities are filled here)
	hs.bvh_nodes.clear(); hs.bvh_slots.clear(); hs.bvh_geom.clear();
	if (n_bvh == 0) hs.bvh_nodes.assign(1, RtBvhNode{});
	else hs.bvh_nodes.resize(n_bvh);
	if (n_leaves == 0) {
		hs.bvh_slots.assign(RT_BVH_LEAF, RT_NO_SLOT);
		hs.bvh_geom.assign(RT_BVH_LEAF, RtF4{0, 0, 0, 0});
	} else {
		hs.bvh_slots.resize(n_leaves * RT_BVH_LEAF);
		hs.bvh_geom.resize(n_leaves * RT_BVH_LEAF);
	}
	std::mutex mu;
	rt_parallel_blocks(N, 256, hs.slot_geom.size(), [&](size_t n0, size_t n1) {
		std::vector<int> idx;
		struct Job { int node, beg, end, depth; };
		std::vector<Job> jobs;
		int deepest = 0;
		for (size_t n = n0; n < n1; n++) {
			const int off = hs.node_link[n].z, cnt = hs.node_link[n].w;
			if (cnt <= 0) continue;
			idx.resize(cnt);
			for (int i = 0; i < cnt; i++) idx[i] = off + i;
			int next_node = hs.node_bvh[n] + 1, next_leaf = leaf_base[n];  // where this list's next records go
			jobs.clear();
			jobs.push_back(Job{hs.node_bvh[n], 0, cnt, 0});
			while (!jobs.empty()) {
				const Job j = jobs.back();
				jobs.pop_back();
				deepest = std::max(deepest, j.depth);
				Box bb;
				double cmin[3] = {1e300, 1e300, 1e300}, cmax[3] = {-1e300, -1e300, -1e300};
				int min_slot = 0x7fffffff;
				for (int k = 0; k < 3; k++) { bb.lo[k] = INFINITY; bb.hi[k] = -INFINITY; }
				for (int i = j.beg; i < j.end; i++) {
					const Box b = box_of(idx[i]);
					const RtD4& g = hs.slot_geom64[idx[i]];
					const double c[3] = {g.x, g.y, g.z};
					min_slot = std::min(min_slot, idx[i]);
					for (int k = 0; k < 3; k++) {
						bb.lo[k] = std::min(bb.lo[k], b.lo[k]);
						bb.hi[k] = std::max(bb.hi[k], b.hi[k]);
						cmin[k] = std::min(cmin[k], c[k]);
						cmax[k] = std::max(cmax[k], c[k]);
					}
				}
				RtBvhNode nd;
				for (int k = 0; k < 3; k++) { nd.lo[k] = bb.lo[k]; nd.hi[k] = bb.hi[k]; }
				const int count = j.end - j.beg;
				int axis = 0;
				for (int k = 1; k < 3; k++)
					if (cmax[k] - cmin[k] > cmax[axis] - cmin[axis]) axis = k;
				if (count <= RT_BVH_LEAF) {  // (leaves never hold more: the leaf test is unrolled over RT_BVH_LEAF entries)
					std::sort(idx.begin() + j.beg, idx.begin() + j.end);  // ascending slots: list order within the leaf
					nd.a = next_leaf * RT_BVH_LEAF;
					nd.b = count;
					for (int i = 0; i < RT_BVH_LEAF; i++) {  // padded to RT_BVH_LEAF entries (finite geometry, slot = none)
						const bool real = j.beg + i < j.end;
						hs.bvh_slots[(size_t)nd.a + i] = real ? idx[j.beg + i] : RT_NO_SLOT;
						hs.bvh_geom[(size_t)nd.a + i] = hs.slot_geom[idx[real ? j.beg + i : j.beg]];
					}
					next_leaf++;
				} else {
					const int mid = j.beg + count / 2;
					// (coincident centres: the comparator falls back to the slot numbers, any split is as good)
					std::nth_element(idx.begin() + j.beg, idx.begin() + mid, idx.begin() + j.end, [&](int p, int q) {
						const double a = axis == 0 ? hs.slot_geom64[p].x : axis == 1 ? hs.slot_geom64[p].y : hs.slot_geom64[p].z;
						const double b = axis == 0 ? hs.slot_geom64[q].x : axis == 1 ? hs.slot_geom64[q].y : hs.slot_geom64[q].z;
						return a < b || (a == b && p < q);
					});
					const bool min_left = std::find(idx.begin() + j.beg, idx.begin() + mid, min_slot) != idx.begin() + mid;
					nd.a = next_node;
					nd.b = -(min_slot + 1);
					next_node += 2;
					jobs.push_back(Job{nd.a + (min_left ? 0 : 1), j.beg, mid, j.depth + 1});
					jobs.push_back(Job{nd.a + (min_left ? 1 : 0), mid, j.end, j.depth + 1});
				}
				hs.bvh_nodes[j.node] = nd;
			}
		}
		std::lock_guard<std::mutex> g(mu);
		hs.max_bvh_depth = std::max(hs.max_bvh_depth, deepest);
	});
	// the walk records (rt_common.h: RtWNode)
	hs.node_walk.resize(N);
	rt_parallel_blocks(N, 4096, N, [&](size_t n0, size_t n1) {
		for (size_t n = n0; n < n1; n++) {
			const RtPNode& pk = hs.node_pk[n];
			const RtI4& link = hs.node_link[n];
			RtWNode w{};
			w.x = pk.x; w.y = pk.y; w.z = pk.z; w.size = pk.size;
			w.child_base = pk.child_base; w.child_mask = pk.child_mask;
			w.up = link.x < 0 ? -1 : (int)((unsigned)link.x | ((unsigned)link.y << 28));
			w.a = -1;
			w.b = 0;
			if (hs.node_bvh[n] >= 0) {
				const RtBvhNode& r = hs.bvh_nodes[hs.node_bvh[n]];
				for (int k = 0; k < 3; k++) { w.lo[k] = r.lo[k]; w.hi[k] = r.hi[k]; }
				w.a = r.a;
				w.b = r.b;
			}
			hs.node_walk[n] = w;
		}
	});
}

// 1 when the stack of the bounce stage's ordered walk (RT_WALK_STACK entries) holds the deepest case: up to 7
// siblings per octree level + 8, and one list BVH on top (depth + 1 entries, binary DFS)
inline int rt_ordered_walk_fits(const RtHostScene& hs) {
	if (hs.node_link.size() > (size_t)RT_WNODE_PARENT_MASK) return 0;  // RtWNode.up packs the parent in 28 bits
	return 7 * hs.max_depth + 8 + hs.max_bvh_depth + 2 <= RT_WALK_STACK ? 1 : 0;
}

// RT_B200_PACK_TIMING=1: the stages of rt_pack_scene on stderr
struct RtPackClock {
	bool on = getenv("RT_B200_PACK_TIMING") != nullptr;
	std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
	void mark(const char* what) {
		if (!on) return;
		const auto now = std::chrono::steady_clock::now();
		fprintf(stderr, "[rt_pack_scene] %-16s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
		t = now;
	}
};

// Validates `sc` and fills `hs`.  On failure returns the status and a message in `err`.
inline rt_status rt_pack_scene(const rt_scene_desc* sc, RtHostScene& hs, std::string& err) {
#define RT_FAIL(st, ...)              \
	do {                              \
		err = rt_format(__VA_ARGS__); \
		return st;                    \
	} while (0)
	if (!sc) RT_FAIL(RT_ERR_INVALID, "rt_scene_upload: scene is NULL");
	if (sc->struct_size != sizeof(rt_scene_desc))
		RT_FAIL(RT_ERR_INVALID, "rt_scene_upload: struct_size %u != %zu (ABI mismatch)", sc->struct_size,
		        sizeof(rt_scene_desc));
	const uint32_t N = sc->n_nodes, L = sc->n_list, E = sc->n_entities;
	if (N == 0) RT_FAIL(RT_ERR_INVALID, "scene has no root node");
	if (!sc->node_pos || !sc->node_size || !sc->node_child || !sc->node_parent || !sc->node_octant || !sc->node_list_off)
		RT_FAIL(RT_ERR_INVALID, "node arrays must not be NULL");
	if (L && !sc->list_entity) RT_FAIL(RT_ERR_INVALID, "list_entity is NULL");
	if (E && (!sc->ent_type || !sc->ent_pos || !sc->ent_extent || !sc->ent_material || !sc->ent_texture || !sc->ent_substance))
		RT_FAIL(RT_ERR_INVALID, "entity arrays must not be NULL");
	if (sc->n_materials && (!sc->mat_response || !sc->mat_light || !sc->mat_mirror || !sc->mat_roughness))
		RT_FAIL(RT_ERR_INVALID, "material arrays must not be NULL");
	if (sc->n_textures && (!sc->tex_kind || !sc->tex_color)) RT_FAIL(RT_ERR_INVALID, "texture arrays must not be NULL");
	if (sc->n_substances && !sc->sub_refractive_index) RT_FAIL(RT_ERR_INVALID, "substance array is NULL");
	if (sc->n_materials >= (1u << RT_ATTR_TYPE_SHIFT)) RT_FAIL(RT_ERR_UNSUPPORTED, "too many materials");
	if (sc->node_parent[0] != -1)
		RT_FAIL(RT_ERR_UNSUPPORTED, "unsupported octree: node 0 must be the absolute root (tree.parent == undefined)");
	if (sc->node_list_off[0] != 0 || sc->node_list_off[N] != L)
		RT_FAIL(RT_ERR_INVALID, "node_list_off must start at 0 and end at n_list");

	RtPackClock clk;
	RtFirstError fe;
#define RT_FAIL_AT(key, st, ...)                                   \
	do {                                                           \
		fe.report((uint64_t)(key), st, rt_format(__VA_ARGS__)); \
		return;                                                    \
	} while (0)
#define RT_STAGE_END()         \
	do {                       \
		if (fe.any) {          \
			err = fe.msg;      \
			return fe.st;      \
		}                      \
	} while (0)
	const size_t work = (size_t)N + L;
	// ---- validate the links, then renumber breadth-first (children of a node become consecutive)
	// node_octant is the reference's index_within_parent() (src/octree_space.ts:110-125), which derives the
	// index from the cell positions ("FIXME: ... not merely by checking geometry"): for a root whose cell
	// corners are not exactly representable it can disagree with the slot the node really sits in, and the
	// reference's own walker then steps back into the wrong octant.  Such a tree is refused, not guessed at.
#define RT_OCTANT_HINT "index_within_parent() disagrees with the child slot (src/octree_space.ts:110-125 computes it from positions): " \
	"use a root cube whose size is a power of two, placed at a multiple of it"
	rt_parallel_blocks(N, 4096, work, [&](size_t i0, size_t i1) {
		for (size_t i = i0; i < i1; i++) {
			const double s = sc->node_size[i];
			if (!(s > 0) || !std::isfinite(s)) RT_FAIL_AT(i, RT_ERR_INVALID, "node %zu: bad size", i);
			if (sc->node_list_off[i + 1] < sc->node_list_off[i]) RT_FAIL_AT(i, RT_ERR_INVALID, "node %zu: list offsets not monotone", i);
			const int par = sc->node_parent[i];
			if (i > 0 && (par < 0 || (uint32_t)par >= N)) RT_FAIL_AT(i, RT_ERR_INVALID, "node %zu: bad parent %d", i, par);
			const int oc = sc->node_octant[i];
			if (i > 0 && (oc < 0 || oc > 7)) RT_FAIL_AT(i, RT_ERR_INVALID, "node %zu: bad octant %d", i, oc);
			if (i > 0 && sc->node_child[(size_t)par * 8 + oc] != (int)i)
				RT_FAIL_AT(i, RT_ERR_INVALID, "node %zu: parent %d does not list it as child %d; " RT_OCTANT_HINT, i, par, oc);
			for (int c = 0; c < 8; c++) {
				const int ch = sc->node_child[i * 8 + c];
				if (ch < -1 || ch >= (int)N || ch == 0) RT_FAIL_AT(i, RT_ERR_INVALID, "node %zu: bad child %d", i, ch);
				if (ch > 0 && (sc->node_parent[ch] != (int)i || sc->node_octant[ch] != c))
					RT_FAIL_AT(i, RT_ERR_INVALID, "node %zu: child %d does not point back; " RT_OCTANT_HINT, i, ch);
			}
		}
	});
	RT_STAGE_END();
	clk.mark("links");
	std::vector<int> perm(N, -1), order;  // perm[old] = new, order[new] = old
	order.reserve(N);
	order.push_back(0);
	perm[0] = 0;
	for (size_t head = 0; head < order.size(); head++) {
		const int i = order[head];
		for (int c = 0; c < 8; c++) {
			const int ch = sc->node_child[(size_t)i * 8 + c];
			if (ch > 0) {
				if (perm[ch] >= 0) RT_FAIL(RT_ERR_INVALID, "node %d is reachable twice", ch);
				perm[ch] = (int)order.size();
				order.push_back(ch);
			}
		}
	}
	if (order.size() != N) RT_FAIL(RT_ERR_INVALID, "%zu of %u nodes are not reachable from the root", (size_t)N - order.size(), N);
	hs.node_geom.resize(N);
	hs.node_geom64.resize(N);
	hs.node_link.resize(N);
	hs.node_child.resize((size_t)N * 8);
	hs.node_pk.resize(N);
	// Slots follow the breadth-first node numbering (each node's run keeps its insertion order), so the lists
	// of sibling nodes are neighbours in memory, like their records.
	std::vector<uint32_t> slot_base(N + 1, 0);
	for (uint32_t ni = 0; ni < N; ni++) slot_base[ni + 1] = slot_base[ni] + (sc->node_list_off[order[ni] + 1] - sc->node_list_off[order[ni]]);
	double scale = 0;
	std::mutex scale_mu;
	auto merge_scale = [&](double v) {
		std::lock_guard<std::mutex> g(scale_mu);
		scale = std::max(scale, v);
	};
	rt_parallel_blocks(N, 4096, work, [&](size_t n0, size_t n1) {
		double sc_max = 0;
		for (size_t ni = n0; ni < n1; ni++) {
			const int i = order[ni];
			const double* p = sc->node_pos + 3 * (size_t)i;
			const double s = sc->node_size[i];
			hs.node_geom[ni] = RtF4{(float)p[0], (float)p[1], (float)p[2], (float)s};
			hs.node_geom64[ni] = RtD4{p[0], p[1], p[2], s};
			const int par = sc->node_parent[i];
			const int off = (int)slot_base[ni], cnt = (int)(slot_base[ni + 1] - slot_base[ni]);
			hs.node_link[ni] = RtI4{i > 0 ? perm[par] : -1, i > 0 ? sc->node_octant[i] : -1, off, cnt};
			int base = -1, mask = 0;
			for (int c = 0; c < 8; c++) {
				const int ch = sc->node_child[(size_t)i * 8 + c];
				hs.node_child[ni * 8 + c] = ch > 0 ? perm[ch] : -1;
				if (ch > 0) {
					if (base < 0) base = perm[ch];
					mask |= 1 << c;
				}
			}
			hs.node_pk[ni] = RtPNode{(float)p[0], (float)p[1], (float)p[2], (float)s, off, cnt, base, mask};
			for (int k = 0; k < 3; k++) sc_max = std::max(sc_max, std::fabs(p[k]) + s);
		}
		merge_scale(sc_max);
	});
	{
		std::vector<int> depth(N, 0);  // breadth-first numbering: a parent always precedes its children
		hs.max_depth = 0;
		for (uint32_t i = 1; i < N; i++) {
			depth[i] = depth[hs.node_link[i].x] + 1;
			hs.max_depth = std::max(hs.max_depth, depth[i]);
		}
	}
	clk.mark("renumber");
	hs.slot_geom.resize(L);
	hs.slot_geom64.resize(L);
	hs.slot_attr.resize(L);
	hs.slot_node.resize(L);
	// (error keys: 4 * slot + the place of the check in a single thread's order)
	// An entity whose bounding cube has a face ON a cell plane of the tree (a coordinate that is a whole number of the
	// smallest cells, counted from the root's corner) fills or touches cells exactly: rays can merely touch such cells,
	// and what the reference does then is its tie rules (rt_trace.cuh: hit_only_touches_its_cell).  Random scenes have
	// none; scenes placed on a grid - a floor on the root's face is enough - do, and get the tie checks by themselves.
	std::atomic<bool> on_planes{false};
	const double root_size = sc->node_size[0];
	const int plane_depth = std::min(hs.max_depth, 48);
	rt_parallel_blocks(N, 1024, work, [&](size_t n0, size_t n1) {
		double sc_max = 0;
		bool planes = false;
		for (size_t ni = n0; ni < n1; ni++) {
			const int i = order[ni];
			const uint32_t beg = sc->node_list_off[i], end = sc->node_list_off[i + 1];
			uint32_t s = slot_base[ni];
			for (uint32_t li = beg; li < end; li++, s++) {
				const uint32_t e = sc->list_entity[li];
				if (e >= E) RT_FAIL_AT(4ull * s, RT_ERR_INVALID, "list slot %u: entity %u out of range", li, e);
				const double* p = sc->ent_pos + 3 * (size_t)e;
				const double ext = sc->ent_extent[e];
				const uint32_t type = sc->ent_type[e];
				if (type > RT_ENTITY_BOX) RT_FAIL_AT(4ull * s + 2, RT_ERR_UNSUPPORTED, "unsupported Entity subclass (type %u) for entity %u", type, e);
				if (!(ext > 0) || !std::isfinite(ext)) RT_FAIL_AT(4ull * s + 2, RT_ERR_INVALID, "entity %u: bad extent", e);
				const int m = sc->ent_material[e], t = sc->ent_texture[e], sb = sc->ent_substance[e];
				if (m < 0 || (uint32_t)m >= sc->n_materials) RT_FAIL_AT(4ull * s + 2, RT_ERR_INVALID, "entity %u: material %d", e, m);
				if (t < 0 || (uint32_t)t >= sc->n_textures) RT_FAIL_AT(4ull * s + 2, RT_ERR_INVALID, "entity %u: texture %d", e, t);
				if (sb < -1 || sb >= (int)sc->n_substances) RT_FAIL_AT(4ull * s + 2, RT_ERR_INVALID, "entity %u: substance %d", e, sb);
				const float w = type == RT_ENTITY_SPHERE ? (float)(ext / 2) : -(float)(ext / 2);
				hs.slot_geom[s] = RtF4{(float)p[0], (float)p[1], (float)p[2], w};
				hs.slot_geom64[s] = RtD4{p[0], p[1], p[2], ext};
				hs.slot_attr[s] = RtI4{(int)e, m | (int)(type << RT_ATTR_TYPE_SHIFT), t, sb};
				hs.slot_node[s] = (int)ni;
				for (int k = 0; k < 3; k++) {
					sc_max = std::max(sc_max, std::fabs(p[k]) + ext);
					const double c_lo = std::ldexp((p[k] - ext / 2 - sc->node_pos[k]) / root_size, plane_depth);
					const double c_hi = std::ldexp((p[k] + ext / 2 - sc->node_pos[k]) / root_size, plane_depth);
					planes |= c_lo == std::floor(c_lo) || c_hi == std::floor(c_hi);
				}
			}
		}
		merge_scale(sc_max);
		if (planes) on_planes.store(true);
	});
	hs.entities_on_cell_planes = on_planes.load();
	if (clk.on) fprintf(stderr, "[rt_pack_scene] entity faces on cell planes: %s\n", hs.entities_on_cell_planes ? "yes (exact-tie checks on)" : "no");
	{
		std::vector<uint8_t> seen(E, 0);  // an entity sits in ONE node's Set (src/entity.ts:50-56)
		uint32_t s = 0;
		for (uint32_t ni = 0; ni < N && !(fe.any && 4ull * s > fe.key); ni++) {
			const int i = order[ni];
			for (uint32_t li = sc->node_list_off[i]; li < sc->node_list_off[i + 1]; li++, s++) {
				const uint32_t e = sc->list_entity[li];
				if (e >= E) continue;
				if (seen[e]) {
					fe.report(4ull * s + 1, RT_ERR_INVALID, rt_format("entity %u is listed in more than one node", e));
					break;
				}
				seen[e] = 1;
			}
		}
	}
	RT_STAGE_END();
	hs.err_l = (float)(16.0 * 1.1920929e-7 * scale);
	clk.mark("slots");
	{
		// The two geometric invariants every walker here relies on to be exact (the reference's own walker only needs
		// the links): a child's cell IS the parent's octant, and every entity's cubic AABB lies inside its node's cube
		// (what add_entity_to_octree guarantees, src/octree_entity.ts:60-79,174-188).  A tree from another flattener,
		// or one whose entities were moved without being re-inserted, would render differently here without an error:
		// it is refused instead.  float64, O(nodes + entities), tolerance = a few ulps of the coordinate scale.
		const double tol = 1e-12 * (scale > 0 ? scale : 1.0);
		rt_parallel_blocks(N, 4096, work, [&](size_t i0, size_t i1) {
			for (size_t i = std::max<size_t>(i0, 1); i < i1; i++) {
				const int par = sc->node_parent[i], oc = sc->node_octant[i];
				const double ps = sc->node_size[par], half = ps / 2;
				if (std::fabs(sc->node_size[i] - half) > tol)
					RT_FAIL_AT(i, RT_ERR_INVALID, "node %zu: size %.17g is not half of its parent's %.17g", i, sc->node_size[i], ps);
				for (int k = 0; k < 3; k++) {
					const double want = sc->node_pos[3 * (size_t)par + k] + (((oc >> k) & 1) ? half : 0.0);
					if (std::fabs(sc->node_pos[3 * i + k] - want) > tol)
						RT_FAIL_AT(i, RT_ERR_INVALID, "node %zu: its cell is not octant %d of its parent's cube (axis %d: %.17g, expected %.17g)", i, oc, k,
						           sc->node_pos[3 * i + k], want);
				}
			}
		});
		RT_STAGE_END();
		rt_parallel_blocks(N, 1024, work, [&](size_t i0, size_t i1) {
			for (size_t i = i0; i < i1; i++) {
				const double* np = sc->node_pos + 3 * i;
				const double ns = sc->node_size[i];
				for (uint32_t li = sc->node_list_off[i]; li < sc->node_list_off[i + 1]; li++) {
					const uint32_t e = sc->list_entity[li];
					const double* p = sc->ent_pos + 3 * (size_t)e;
					const double h = sc->ent_extent[e] / 2;
					for (int k = 0; k < 3; k++)
						if (p[k] - h < np[k] - tol || p[k] + h > np[k] + ns + tol)
							RT_FAIL_AT(li, RT_ERR_INVALID, "entity %u sticks out of the cube of node %zu that lists it (axis %d): re-insert moved entities "
							                               "(add_entity_to_octree) before uploading", e, i, k);
				}
			}
		});
		RT_STAGE_END();
	}
	{
		// The search runs in float32 (RT_PRECISION_F32): a cell's corners and centre must be distinct float32
		// numbers at the scene's coordinate scale, or the walkers cannot tell neighbouring cells apart.  With a
		// unit root that is 23 levels; the reference (float64 throughout) has no such limit.
		double min_size = sc->node_size[0];
		for (uint32_t i = 0; i < N; i++) min_size = std::min(min_size, sc->node_size[i]);
		const double ulp = std::ldexp(1.0, std::ilogb(scale > 0 ? scale : 1.0) - 23);
		hs.f32_refusal.clear();
		if (min_size * 0.5 < ulp)  // (not an upload error: RT_PRECISION_F64 renders such a tree; a float32 render call is refused)
			hs.f32_refusal = rt_format("octree cells of size %.3g are below the float32 resolution of the search at coordinate scale %.3g "
			                           "(a unit root allows 23 levels): lower max_in_depth, or render with RT_PRECISION_F64", min_size, scale);
	}
	clk.mark("geometry checks");
	rt_build_list_bvhs(hs);
	clk.mark("list BVHs");
	hs.materials.resize(sc->n_materials);
	hs.any_transmission = false;
	for (uint32_t i = 0; i < sc->n_materials; i++) {
		if (sc->mat_response[i] > RT_RESPONSE_BOTH) RT_FAIL(RT_ERR_INVALID, "material %u: response", i);
		hs.materials[i].flags =
		    sc->mat_response[i] | (sc->mat_light[i] ? RT_MAT_LIGHT : 0u) | (sc->mat_mirror[i] ? RT_MAT_MIRROR : 0u);
		hs.materials[i]._pad = 0;
		hs.materials[i].roughness = sc->mat_roughness[i];
		hs.any_transmission |= sc->mat_response[i] == RT_RESPONSE_TRANSMISSION;
	}
	hs.textures.resize(sc->n_textures);
	rt_parallel_blocks(sc->n_textures, 8192, sc->n_textures, [&](size_t i0, size_t i1) {  // (one SolidTexture per entity is common)
		for (size_t i = i0; i < i1; i++) {
			RtTexture& T = hs.textures[i];
			memset(&T, 0, sizeof T);
			T.r = sc->tex_color[4 * i];
			T.g = sc->tex_color[4 * i + 1];
			T.b = sc->tex_color[4 * i + 2];
			if (sc->tex_kind[i] > RT_TEXTURE_IMAGE) RT_FAIL_AT(2 * i, RT_ERR_UNSUPPORTED, "unsupported Texture subclass (kind %u)", sc->tex_kind[i]);
			if (sc->tex_kind[i] == RT_TEXTURE_IMAGE && sc->tex_loaded && sc->tex_loaded[i]) {
				if (!sc->tex_width || !sc->tex_height || !sc->tex_texel_off || !sc->texels)
					RT_FAIL_AT(2 * i + 1, RT_ERR_INVALID, "image texture arrays must not be NULL");
				const int64_t w = sc->tex_width[i], h = sc->tex_height[i];
				if (w <= 0 || h <= 0 || sc->tex_texel_off[i] + (uint64_t)(w * h) > sc->n_texels)
					RT_FAIL_AT(2 * i + 1, RT_ERR_INVALID, "texture %zu: %lldx%lld texels out of the pool", i, (long long)w, (long long)h);
				T.image = 1;
				T.width = (int)w;
				T.height = (int)h;
				T.texel_off = sc->tex_texel_off[i];
			}
		}
	});
	RT_STAGE_END();
	hs.substances.assign(sc->sub_refractive_index, sc->sub_refractive_index + sc->n_substances);
	hs.texels.assign(sc->texels, sc->texels + (sc->texels ? sc->n_texels * 3 : 0));
	clk.mark("materials+texels");
	for (int k = 0; k < 3; k++) hs.root_pos[k] = sc->node_pos[k];
	hs.root_size = sc->node_size[0];
	// bound on the float32 error of a centre-to-line distance for coordinates up to `scale`
	hs.err_l = (float)(16.0 * 1.1920929e-7 * scale);
	return RT_OK;
#undef RT_FAIL
#undef RT_FAIL_AT
#undef RT_STAGE_END
}

// ---- host float64 restatements used once per frame (src/raytracer.ts:309-313) -----------------
// node_at_pos (src/octree_space.ts:61-93)
inline bool rt_host_node_at_pos(const RtHostScene& c, const double* p, int& node, int& octant) {
	double np[3] = {c.root_pos[0], c.root_pos[1], c.root_pos[2]};
	double ns = c.root_size;
	for (int i = 0; i < 3; i++)
		if (!(p[i] >= np[i] && p[i] < np[i] + ns)) return false;
	int cur = 0, next = 0, idx = 0;
	while (next >= 0) {
		const double k = 2.0 / ns;
		const int ix = (int)((p[0] - np[0]) * k), iy = (int)((p[1] - np[1]) * k), iz = (int)((p[2] - np[2]) * k);
		cur = next;
		idx = (iz << 2) + (iy << 1) + ix;
		if (idx < 0 || idx > 7) return false;
		next = c.node_child[(size_t)cur * 8 + idx];
		ns = ns / 2.0;
		np[0] += ix * ns;
		np[1] += iy * ns;
		np[2] += iz * ns;
	}
	node = cur;
	octant = idx;
	return true;
}
// entity_at_pos (src/octree_entity.ts:191-202) -> slot or -1
inline int rt_host_entity_at_pos(const RtHostScene& c, const double* p) {
	int node, octant;
	if (!rt_host_node_at_pos(c, p, node, octant)) return -1;
	while (node >= 0) {
		const RtI4 link = c.node_link[node];
		for (int s = link.z; s < link.z + link.w; s++) {
			const RtD4& g = c.slot_geom64[s];
			const int type = c.slot_attr[s].y >> RT_ATTR_TYPE_SHIFT;
			bool within;
			if (type == 0) {  // entity_sphere.ts:63-66
				const double d0 = p[0] - g.x, d1 = p[1] - g.y, d2 = p[2] - g.z;
				double s2 = 0;
				s2 += d0 * d0;
				s2 += d1 * d1;
				s2 += d2 * d2;
				within = s2 <= g.w * g.w / 4;
			} else {  // entity_box.ts:47-52
				within = p[0] >= g.x && p[0] < g.x + g.w && p[1] >= g.y && p[1] < g.y + g.w && p[2] >= g.z &&
				         p[2] < g.z + g.w;
			}
			if (within) return s;
		}
		node = link.x;
	}
	return -1;
}

// The row part of Camera.get_dir_for_each_pixel (src/view/camera.ts:207-250, iter_v): fr rotated towards up,
// iterated exactly like the generator, outwards from the middle row.  The scans along the rows are iterated on
// the device (rt_trace.cuh: raygen_half_row) from these and (scan_cos, scan_sin) = rot_scan_h_v.
inline void rt_build_camera_rows(const rt_camera& cam, std::vector<RtD4>& row, double& scan_cos, double& scan_sin) {
	const int W = (int)cam.width, H = (int)cam.height;
	row.assign(H, RtD4{0, 0, 0, 0});
	const double rad_h = cam.fov_h / W, rad_v = cam.fov_v / H;
	scan_cos = std::cos(rad_h);
	scan_sin = std::sin(rad_h);
	const double cv = std::cos(rad_v), sv = std::sin(rad_v);
	auto rot3 = [](double* bx, double* by, double c, double s) {  // rotate_vectors vector.ts:318-323
		for (int k = 0; k < 3; k++) {
			const double x = bx[k] * c + by[k] * s;
			const double y = bx[k] * -s + by[k] * c;
			bx[k] = x;
			by[k] = y;
		}
	};
	const int y0 = H >> 1;
	double fr[3] = {cam.fr[0], cam.fr[1], cam.fr[2]}, up[3] = {cam.up[0], cam.up[1], cam.up[2]};
	for (int y = y0; y < H; y++) {
		row[y] = RtD4{fr[0], fr[1], fr[2], 0};
		rot3(fr, up, cv, sv);
	}
	double fr2[3] = {cam.fr[0], cam.fr[1], cam.fr[2]}, up2[3] = {cam.up[0], cam.up[1], cam.up[2]};
	rot3(fr2, up2, cv, -sv);
	for (int y = y0 - 1; y >= 0; y--) {
		row[y] = RtD4{fr2[0], fr2[1], fr2[2], 0};
		rot3(fr2, up2, cv, -sv);
	}
}

// Argument checks shared by every render entry point.
inline rt_status rt_check_render_args(bool has_scene, uint32_t n_textures, uint32_t n_substances, const rt_camera* cam,
                                      const rt_params* prm, std::string& err) {
	if (!cam || !prm) { err = "rt_render: camera and params must not be NULL"; return RT_ERR_INVALID; }
	if (!has_scene) { err = "rt_render: no scene uploaded"; return RT_ERR_NO_SCENE; }
	if (cam->width == 0 || cam->height == 0 || cam->width > 65536 || cam->height > 65536) {
		err = rt_format("rt_render: bad frame size %ux%u", cam->width, cam->height);
		return RT_ERR_INVALID;
	}
	if ((cam->flags & RT_CAM_REFERENCE_EXTENTS) && cam->width != cam->height) {
		err = "x or y out of bounds";
		return RT_ERR_BOUNDS;
	}
	if (prm->precision != RT_PRECISION_F32 && prm->precision != RT_PRECISION_F64) { err = rt_format("unsupported precision %u", prm->precision); return RT_ERR_UNSUPPORTED; }
	if (prm->n_frames == 0) { err = "rt_render: n_frames must be >= 1"; return RT_ERR_INVALID; }
	if (prm->sky_texture < 0 || (uint32_t)prm->sky_texture >= n_textures) {
		err = rt_format("rt_render: sky_texture %d out of range", prm->sky_texture);
		return RT_ERR_INVALID;
	}
	if (prm->default_substance < 0 || (uint32_t)prm->default_substance >= n_substances) {
		err = rt_format("rt_render: default_substance %d out of range", prm->default_substance);
		return RT_ERR_INVALID;
	}
	return RT_OK;
}

// Fills the camera / start-state / config part of an RtFrame (pointers are left to the caller).
inline rt_status rt_fill_frame(const RtHostScene& hs, const rt_camera* cam, const rt_params* prm, RtFrame& F,
                               std::string& err) {
	if (prm->precision == RT_PRECISION_F32 && !hs.f32_refusal.empty()) {
		err = hs.f32_refusal;
		return RT_ERR_UNSUPPORTED;
	}
	F.search64 = prm->precision == RT_PRECISION_F64;
	// exact ties (rt_b200.h: RT_PARAM_EXACT_TIES): asked for, or the scene has entities whose faces lie on cell planes
	// (rt_pack_scene), or possible for camera rays - the camera stands on a cell plane of the octree (a coordinate that is a multiple of the smallest cell's size, counted from the root's corner)
	F.tie_checks = ((prm->flags & RT_PARAM_EXACT_TIES) || hs.entities_on_cell_planes) ? 1 : 0;
	for (int k = 0; k < 3 && !F.tie_checks; k++) {
		const double cells = std::ldexp((cam->pos[k] - hs.root_pos[k]) / hs.root_size, std::min(hs.max_depth, 48));
		if (cells == std::floor(cells)) F.tie_checks = 1;
	}
	for (int i = 0; i < 3; i++) {
		F.pos[i] = cam->pos[i];
		F.lf[i] = cam->lf[i];
	}
	F.width = (int)cam->width;
	F.height = (int)cam->height;
	int node = -1, octant = -1;
	if (!rt_host_node_at_pos(hs, cam->pos, node, octant)) {
		node = -1;
		octant = -1;
	}
	F.start_node = node;
	F.start_octant = octant;
	const int start_slot = rt_host_entity_at_pos(hs, cam->pos);
	F.start_substance = start_slot >= 0 ? hs.slot_attr[start_slot].w : prm->default_substance;
	if (F.start_substance < 0) {
		if (hs.any_transmission) {
			err = "camera is inside an entity whose substance is undefined and the scene has TRANSMISSION materials "
			      "(the reference throws a TypeError at the first refraction)";
			return RT_ERR_UNSUPPORTED;
		}
		F.start_substance = prm->default_substance;
	}
	F.refmax = prm->refmax;
	F.sky_texture = prm->sky_texture;
	F.default_substance = prm->default_substance;
	F.attenuation = prm->distance_attenuation_factor;
	F.n_frames = prm->n_frames;
	F.frame_first = prm->frame_first;
	F.w_first = 1.0 / (double)(1u + prm->frame_first);  // ExposureBuffer.next_frame (src/view/exposure_buffer.ts:58)
	F.w1_first = 1.0 - F.w_first;
	F.rng_seed = prm->rng_seed;
	return RT_OK;
}

// The origin chain of the camera rays: start node, its parent, ..., root, in the order the walker
// returns them (post-order), as slot ranges for the lock-step pre-test (rt_trace.cuh: pretest_chain).
inline void rt_fill_chain(const RtHostScene& hs, RtFrame& F) {
	F.chain_levels = 0;
	F.packet_ok = 0;
	int n = F.start_node, oct = F.start_octant;
	for (; n >= 0 && F.chain_levels < RT_MAX_CHAIN; n = hs.node_link[n].x) {
		const int k = F.chain_levels++;
		F.chain_beg[k] = hs.node_link[n].z;
		F.chain_end[k] = hs.node_link[n].z + hs.node_link[n].w;
		F.chain_node[k] = n;
		F.chain_oct[k] = oct;
		oct = hs.node_link[n].y;
	}
	// the packet stage needs the whole chain (up to the root); its node stack is checked as it grows
	F.packet_ok = (F.start_node >= 0 && n < 0) ? 1 : 0;
}
