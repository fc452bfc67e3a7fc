// rt_build_gpu.cuh — device-side bulk octree builder: SURVEY §8f row N1, "then GPU-side build".
//
// The same restatement of add_entity_to_octree (src/octree_entity.ts:174-188, max_out_depth = 0) as the host
// builder in rt_build.h - the node an entity ends in depends on the entity alone - laid out for the GPU:
//   1. one thread per entity walks down from the root with the reference's float64 expressions (octant =
//      ((min - node.pos) * (2 / node.size)) << 0, child cube = node.pos + octant * (node.size / 2), stop when
//      the AABB does not fit the child or at max_in_depth) and emits the PATH KEY of its node - one 4-bit digit
//      (octant + 1) per level, most significant first - together with the keys of all its ancestors: the
//      reference creates every node on the way (extend_tree_inside_to_fit_up_to_depth, :92-114);
//   2. radix sort + unique of all emitted keys = the set of nodes, and since a parent's key is its child's key
//      with the last digit cleared, the sorted order IS the depth-first pre-order with children 0..7;
//   3. one thread per node replays the arithmetic of its path for the node's float64 position and size (the
//      expressions the reference evaluates when it creates the node) and finds its parent by binary search;
//   4. a stable radix sort of the entity indices by node keeps the insertion order inside every list
//      (EntitySet is an insertion-ordered Set, src/octree_entity.ts:32-49).
// Sorting, unique and the scans are CUB's (the CUDA toolkit's device-wide primitives); the placement, node and
// list kernels are this file's.  No FMA contraction anywhere: __dmul_rn / __dadd_rn, as in the ray path.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include "rt_build.h"

#define RT_GPU_BUILD_MAX_DEPTH 16  // 16 digits of 4 bits

namespace rt_gpu_build {

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

struct Root { double pos[3]; double size; };

// 1. placement.  keys_all: [n][max_in_depth + 1] the key of every node on the entity's path (unused tail =
// ~0 sentinel); ent_key[e] = the key of its own node.  first_bad: lowest entity index that does not fit the root.
__global__ void place_kernel(Root root, uint32_t n, const uint8_t* __restrict__ type, const double* __restrict__ epos,
                             const double* __restrict__ extent, uint32_t max_in_depth, unsigned long long* __restrict__ keys_all,
                             unsigned long long* __restrict__ ent_key, unsigned* __restrict__ first_bad) {
	const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n) return;
	const double ext = extent[e];
	double mn[3];
	if (type[e] == 0) {  // SphereEntity.get_aabb: pos - (d,d,d) * 0.5
		for (int k = 0; k < 3; k++) mn[k] = dsub(epos[3 * (size_t)e + k], dmul(ext, 0.5));
	} else {  // BoxEntity.get_aabb: pos - (size/2)
		const double h = ddiv(ext, 2.0);
		for (int k = 0; k < 3; k++) mn[k] = dsub(epos[3 * (size_t)e + k], h);
	}
	auto fits = [&](const double* p, double s) {  // aabb_in_space -> space_in_space (closed upper bound)
		for (int k = 0; k < 3; k++)
			if (!(mn[k] >= p[k] && dadd(mn[k], ext) <= dadd(p[k], s))) return false;
		return true;
	};
	bool inside = true;
	for (int k = 0; k < 3; k++) inside = inside && mn[k] >= root.pos[k] && mn[k] < dadd(root.pos[k], root.size);
	unsigned long long* out = keys_all + (size_t)e * (max_in_depth + 1);
	if (!inside || !fits(root.pos, root.size)) {
		atomicMin(first_bad, e);
		for (uint32_t d = 0; d <= max_in_depth; d++) out[d] = ~0ull;
		ent_key[e] = 0ull;
		return;
	}
	double np[3] = {root.pos[0], root.pos[1], root.pos[2]};
	double ns = root.size;
	unsigned long long key = 0ull;
	uint32_t depth = 0;
	out[0] = 0ull;  // the root
	for (; depth < max_in_depth; depth++) {
		const double k2 = ddiv(2.0, ns), hs = ddiv(ns, 2.0);
		int o[3];
		double cp[3];
		for (int k = 0; k < 3; k++) {
			o[k] = (int)dmul(dsub(mn[k], np[k]), k2);  // `<< 0`
			cp[k] = dadd(np[k], dmul((double)o[k], hs));
		}
		if (!fits(cp, hs)) break;
		const int idx = (o[2] << 2) | (o[1] << 1) | o[0];
		key |= (unsigned long long)(idx + 1) << (60 - 4 * depth);
		out[depth + 1] = key;
		np[0] = cp[0]; np[1] = cp[1]; np[2] = cp[2];
		ns = hs;
	}
	for (uint32_t d = depth + 1; d <= max_in_depth; d++) out[d] = ~0ull;
	ent_key[e] = key;
}

__device__ __forceinline__ int lower_bound(const unsigned long long* a, int n, unsigned long long v) {
	int lo = 0, hi = n;
	while (lo < hi) {
		const int mid = (lo + hi) >> 1;
		if (a[mid] < v) lo = mid + 1; else hi = mid;
	}
	return lo;
}

// 3. nodes: position and size by replaying the path, parent by search, the parent's child table
__global__ void node_kernel(Root root, int n_nodes, const unsigned long long* __restrict__ node_key, double* __restrict__ pos,
                            double* __restrict__ size, int32_t* __restrict__ parent, int32_t* __restrict__ octant,
                            int32_t* __restrict__ child) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_nodes) return;
	const unsigned long long key = node_key[i];
	double np[3] = {root.pos[0], root.pos[1], root.pos[2]};
	double ns = root.size;
	int depth = 0, last = -1;
	for (; depth < RT_GPU_BUILD_MAX_DEPTH; depth++) {
		const int digit = (int)((key >> (60 - 4 * depth)) & 15ull);
		if (!digit) break;
		last = digit - 1;
		const double hs = ddiv(ns, 2.0);
		np[0] = dadd(np[0], dmul((double)(last & 1), hs));
		np[1] = dadd(np[1], dmul((double)((last >> 1) & 1), hs));
		np[2] = dadd(np[2], dmul((double)((last >> 2) & 1), hs));
		ns = hs;
	}
	pos[3 * (size_t)i] = np[0]; pos[3 * (size_t)i + 1] = np[1]; pos[3 * (size_t)i + 2] = np[2];
	size[i] = ns;
	if (depth == 0) {
		parent[i] = -1;
		octant[i] = -1;
	} else {
		const unsigned long long pkey = key & ~(15ull << (60 - 4 * (depth - 1)));
		const int p = lower_bound(node_key, n_nodes, pkey);
		parent[i] = p;
		octant[i] = last;
		child[(size_t)p * 8 + last] = i;
	}
}

// 4a. entity -> node index
__global__ void entity_node_kernel(uint32_t n, const unsigned long long* __restrict__ ent_key, const unsigned long long* __restrict__ node_key,
                                   int n_nodes, uint32_t* __restrict__ ent_node, uint32_t* __restrict__ ent_index) {
	const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n) return;
	ent_node[e] = (uint32_t)lower_bound(node_key, n_nodes, ent_key[e]);
	ent_index[e] = e;
}

// 4b. list offsets: first position of each node in the node-sorted entity array
__global__ void list_off_kernel(int n_nodes, uint32_t n, const uint32_t* __restrict__ sorted_node, uint32_t* __restrict__ list_off) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i > n_nodes) return;
	uint32_t lo = 0, hi = n;
	while (lo < hi) {
		const uint32_t mid = (lo + hi) >> 1;
		if (sorted_node[mid] < (uint32_t)i) lo = mid + 1; else hi = mid;
	}
	list_off[i] = lo;
}

struct Buf {
	void* p = nullptr;
	~Buf() { if (p) cudaFree(p); }
	cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
	template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

#define RT_GB(call)                                                                            \
	do {                                                                                       \
		cudaError_t e_ = (call);                                                               \
		if (e_ != cudaSuccess) {                                                               \
			err = std::string(#call) + ": " + cudaGetErrorString(e_);                          \
			return 2;                                                                          \
		}                                                                                      \
	} while (0)

// returns 0 ok, 1 an entity does not fit the root (err = the reference's message), 2 CUDA failure
inline int build(rt_tree& T, cudaStream_t stream, const double* root_pos, double root_size, uint32_t n, const uint8_t* type,
                 const double* epos, const double* extent, uint32_t D, std::string& err, uint64_t* launches) {
	Root root{{root_pos[0], root_pos[1], root_pos[2]}, root_size};
	const size_t n_all = (size_t)n * (D + 1) + 1;  // + the root's own key, so that an empty scene still has its root
	Buf d_type, d_pos, d_ext, d_all, d_all_sorted, d_ent_key, d_node_key, d_count, d_bad, d_tmp;
	RT_GB(d_type.alloc(n)); RT_GB(d_pos.alloc((size_t)n * 24)); RT_GB(d_ext.alloc((size_t)n * 8));
	RT_GB(d_all.alloc(n_all * 8)); RT_GB(d_all_sorted.alloc(n_all * 8)); RT_GB(d_ent_key.alloc((size_t)n * 8));
	RT_GB(d_node_key.alloc(n_all * 8)); RT_GB(d_count.alloc(8)); RT_GB(d_bad.alloc(4));
	if (n) {
		RT_GB(cudaMemcpyAsync(d_type.p, type, n, cudaMemcpyHostToDevice, stream));
		RT_GB(cudaMemcpyAsync(d_pos.p, epos, (size_t)n * 24, cudaMemcpyHostToDevice, stream));
		RT_GB(cudaMemcpyAsync(d_ext.p, extent, (size_t)n * 8, cudaMemcpyHostToDevice, stream));
	}
	RT_GB(cudaMemsetAsync(d_bad.p, 0xff, 4, stream));
	RT_GB(cudaMemsetAsync(d_all.as<unsigned long long>() + (n_all - 1), 0, 8, stream));  // the root key
	if (n) {
		place_kernel<<<(n + 255) / 256, 256, 0, stream>>>(root, n, d_type.as<uint8_t>(), d_pos.as<double>(), d_ext.as<double>(), D,
		                                                 d_all.as<unsigned long long>(), d_ent_key.as<unsigned long long>(), d_bad.as<unsigned>());
		(*launches)++;
	}
	unsigned bad = 0xffffffffu;
	RT_GB(cudaMemcpyAsync(&bad, d_bad.p, 4, cudaMemcpyDeviceToHost, stream));
	// 2. sort + unique
	size_t tmp_sort = 0, tmp_uniq = 0, tmp_pairs = 0;
	RT_GB(cub::DeviceRadixSort::SortKeys(nullptr, tmp_sort, d_all.as<unsigned long long>(), d_all_sorted.as<unsigned long long>(), (int)n_all, 0, 64, stream));
	RT_GB(cub::DeviceSelect::Unique(nullptr, tmp_uniq, d_all_sorted.as<unsigned long long>(), d_node_key.as<unsigned long long>(), d_count.as<int>(), (int)n_all, stream));
	Buf d_ent_node, d_ent_node_sorted, d_ent_index, d_list_entity;
	RT_GB(d_ent_node.alloc((size_t)n * 4)); RT_GB(d_ent_node_sorted.alloc((size_t)n * 4)); RT_GB(d_ent_index.alloc((size_t)n * 4));
	RT_GB(d_list_entity.alloc((size_t)n * 4));
	RT_GB(cub::DeviceRadixSort::SortPairs(nullptr, tmp_pairs, d_ent_node.as<uint32_t>(), d_ent_node_sorted.as<uint32_t>(), d_ent_index.as<uint32_t>(),
	                                      d_list_entity.as<uint32_t>(), (int)n, 0, 32, stream));
	RT_GB(d_tmp.alloc(std::max(tmp_sort, std::max(tmp_uniq, tmp_pairs))));
	RT_GB(cub::DeviceRadixSort::SortKeys(d_tmp.p, tmp_sort, d_all.as<unsigned long long>(), d_all_sorted.as<unsigned long long>(), (int)n_all, 0, 64, stream));
	RT_GB(cub::DeviceSelect::Unique(d_tmp.p, tmp_uniq, d_all_sorted.as<unsigned long long>(), d_node_key.as<unsigned long long>(), d_count.as<int>(), (int)n_all, stream));
	int n_keys = 0;
	RT_GB(cudaMemcpyAsync(&n_keys, d_count.p, 4, cudaMemcpyDeviceToHost, stream));
	RT_GB(cudaStreamSynchronize(stream));
	if (bad != 0xffffffffu) {
		err = "The tree outside-depth limit exceeded (entity " + std::to_string(bad) + " does not fit the root cube; max_out_depth is 0)";
		return 1;
	}
	// the sentinel (unused path slots) sorts last: drop it
	unsigned long long last_key = 0;
	RT_GB(cudaMemcpy(&last_key, d_node_key.as<unsigned long long>() + (n_keys - 1), 8, cudaMemcpyDeviceToHost));
	const int N = last_key == ~0ull ? n_keys - 1 : n_keys;
	// 3. nodes
	Buf d_npos, d_nsize, d_parent, d_octant, d_child, d_list_off;
	RT_GB(d_npos.alloc((size_t)N * 24)); RT_GB(d_nsize.alloc((size_t)N * 8)); RT_GB(d_parent.alloc((size_t)N * 4));
	RT_GB(d_octant.alloc((size_t)N * 4)); RT_GB(d_child.alloc((size_t)N * 32)); RT_GB(d_list_off.alloc((size_t)(N + 1) * 4));
	RT_GB(cudaMemsetAsync(d_child.p, 0xff, (size_t)N * 32, stream));
	node_kernel<<<(N + 255) / 256, 256, 0, stream>>>(root, N, d_node_key.as<unsigned long long>(), d_npos.as<double>(), d_nsize.as<double>(),
	                                                d_parent.as<int32_t>(), d_octant.as<int32_t>(), d_child.as<int32_t>());
	(*launches)++;
	// 4. lists
	if (n) {
		entity_node_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, d_ent_key.as<unsigned long long>(), d_node_key.as<unsigned long long>(), N,
		                                                       d_ent_node.as<uint32_t>(), d_ent_index.as<uint32_t>());
		(*launches)++;
		RT_GB(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_pairs, d_ent_node.as<uint32_t>(), d_ent_node_sorted.as<uint32_t>(), d_ent_index.as<uint32_t>(),
		                                      d_list_entity.as<uint32_t>(), (int)n, 0, 32, stream));  // stable: insertion order inside a list
	}
	list_off_kernel<<<(N + 1 + 255) / 256, 256, 0, stream>>>(N, n, d_ent_node_sorted.as<uint32_t>(), d_list_off.as<uint32_t>());
	(*launches)++;
	RT_GB(cudaGetLastError());
	// 5. back to the host, into the same rt_tree the host builder fills
	T.pos.resize((size_t)N * 3); T.size.resize(N); T.child.resize((size_t)N * 8); T.parent.resize(N); T.octant.resize(N);
	T.ent_node.resize(n); T.list_off.resize(N + 1); T.list_entity.resize(n);
	RT_GB(cudaMemcpyAsync(T.pos.data(), d_npos.p, (size_t)N * 24, cudaMemcpyDeviceToHost, stream));
	RT_GB(cudaMemcpyAsync(T.size.data(), d_nsize.p, (size_t)N * 8, cudaMemcpyDeviceToHost, stream));
	RT_GB(cudaMemcpyAsync(T.child.data(), d_child.p, (size_t)N * 32, cudaMemcpyDeviceToHost, stream));
	RT_GB(cudaMemcpyAsync(T.parent.data(), d_parent.p, (size_t)N * 4, cudaMemcpyDeviceToHost, stream));
	RT_GB(cudaMemcpyAsync(T.octant.data(), d_octant.p, (size_t)N * 4, cudaMemcpyDeviceToHost, stream));
	RT_GB(cudaMemcpyAsync(T.list_off.data(), d_list_off.p, (size_t)(N + 1) * 4, cudaMemcpyDeviceToHost, stream));
	if (n) {
		RT_GB(cudaMemcpyAsync(T.ent_node.data(), d_ent_node.p, (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
		RT_GB(cudaMemcpyAsync(T.list_entity.data(), d_list_entity.p, (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
	}
	RT_GB(cudaStreamSynchronize(stream));
	return 0;
}
#undef RT_GB

}  // namespace rt_gpu_build
