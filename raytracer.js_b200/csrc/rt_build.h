// rt_build.h — host-side bulk octree builder (C++, no CUDA): SURVEY §8f row N1, "host octree build -> flatten".
//
// Restates add_entity_to_octree (src/octree_entity.ts:174-188) for a whole entity array at once, for trees
// that never grow outwards (max_out_depth = 0, what the demo and every benchmark scene use).  For one entity
// the reference does: get_aabb (cubic: min corner + edge; entity_sphere.ts:90-96, entity_box.ts:75-82) ->
// get_covering_node_for_entity (:60-79: deepest existing node that holds the min corner, then up until the
// AABB fits, closed upper bound as in space_in_space, src/space.ts:85-97) -> extend_tree_inside_to_fit_up_to_depth
// (:92-114: keep creating the child octant of the min corner while the AABB fits it, up to max_in_depth) ->
// set_octree (src/entity.ts:50-56: append to that node's Set).  The node an entity ends in therefore depends on
// the entity alone: walk down from the root, octant = ((min - node.pos) * (2 / node.size)) << 0 per axis, child
// cube = node.pos + octant * (node.size / 2), stop when the AABB does not fit the child or the depth limit is
// reached.  Inserting the entities in index order reproduces the reference's tree (same nodes, same float64
// positions, which are computed by the same expressions) and its per-node insertion order.
#pragma once
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

struct rt_tree {
	std::vector<double> pos;     // [n*3]
	std::vector<double> size;    // [n]
	std::vector<int32_t> child;  // [n*8]
	std::vector<int32_t> parent, octant;
	std::vector<uint32_t> ent_node;  // [n_entities] node each entity landed in
	std::vector<uint32_t> list_off, list_entity;
};

// false + message when an entity cannot be placed without growing the tree outwards
// (the reference throws TreeOutsideGrowError with max_out_depth: 0, src/octree_entity.ts:116-171)
inline bool rt_tree_build_impl(rt_tree& T, const double* root_pos, double root_size, uint32_t n, const uint8_t* type,
                               const double* epos, const double* extent, uint32_t max_in_depth, std::string& err) {
	T.pos.assign(root_pos, root_pos + 3);
	T.size.assign(1, root_size);
	T.child.assign(8, -1);
	T.parent.assign(1, -1);
	T.octant.assign(1, -1);
	T.ent_node.resize(n);
	for (uint32_t e = 0; e < n; e++) {
		const double ext = extent[e];
		double mn[3];
		if (type[e] == 0) {  // SphereEntity.get_aabb: pos - (d,d,d) * 0.5
			for (int k = 0; k < 3; k++) mn[k] = epos[3 * (size_t)e + k] - ext * 0.5;
		} else {  // BoxEntity.get_aabb: pos - (size/2)
			const double h = ext / 2;
			for (int k = 0; k < 3; k++) mn[k] = epos[3 * (size_t)e + k] - h;
		}
		auto fits = [&](const double* p, double s) {  // aabb_in_space -> space_in_space (closed upper bound)
			for (int k = 0; k < 3; k++)
				if (!(mn[k] >= p[k] && mn[k] + ext <= p[k] + s)) return false;
			return true;
		};
		// node_at_pos needs the min corner inside the root's half-open cube, the climb needs the AABB to fit it
		bool inside = true;
		for (int k = 0; k < 3; k++) inside = inside && mn[k] >= root_pos[k] && mn[k] < root_pos[k] + root_size;
		if (!inside || !fits(root_pos, root_size)) {
			err = "The tree outside-depth limit exceeded (entity " + std::to_string(e) + " does not fit the root cube; max_out_depth is 0)";
			return false;
		}
		uint32_t node = 0;
		for (uint32_t depth = 0; depth < max_in_depth; depth++) {
			const double* np = &T.pos[3 * (size_t)node];
			const double ns = T.size[node];
			const double k2 = 2.0 / ns, hs = ns / 2;
			int o[3];
			double cp[3];
			for (int k = 0; k < 3; k++) {
				o[k] = (int)((mn[k] - np[k]) * k2);  // `<< 0`
				cp[k] = np[k] + o[k] * hs;
			}
			if (!fits(cp, hs)) break;
			const int idx = (o[2] << 2) | (o[1] << 1) | o[0];
			int32_t ch = T.child[(size_t)node * 8 + idx];
			if (ch < 0) {
				ch = (int32_t)T.size.size();
				T.child[(size_t)node * 8 + idx] = ch;
				T.pos.insert(T.pos.end(), cp, cp + 3);
				T.size.push_back(hs);
				T.child.insert(T.child.end(), 8, -1);
				T.parent.push_back((int32_t)node);
				T.octant.push_back(idx);
			}
			node = (uint32_t)ch;
		}
		T.ent_node[e] = node;
	}
	// per-node lists in insertion order (counting sort by node keeps the entity order within a node)
	const size_t N = T.size.size();
	T.list_off.assign(N + 1, 0);
	for (uint32_t e = 0; e < n; e++) T.list_off[T.ent_node[e] + 1]++;
	for (size_t i = 0; i < N; i++) T.list_off[i + 1] += T.list_off[i];
	T.list_entity.resize(n);
	std::vector<uint32_t> cur(T.list_off.begin(), T.list_off.end() - 1);
	for (uint32_t e = 0; e < n; e++) T.list_entity[cur[T.ent_node[e]]++] = e;
	return true;
}
