// rt_common.h — device-visible scene/frame structures of librt_b200 (HBM layout).
//
// Layout idea: everything the per-ray loop touches is addressed by *slot*, not by entity id.
// A slot is a position in the concatenation of all node entity lists (each run in EntitySet
// insertion order, src/octree_entity.ts:32-49).  Every entity lives in exactly one node
// (src/octree_entity.ts:174-188), so slots are a permutation of the entities and a node's list is
// one contiguous run of 16-byte records: a warp scanning a list issues consecutive LDG.128s.
#pragma once
#include <stdint.h>

// The kernel body (rt_trace.cuh) is device code in the shipped library; the same source is compiled as
// plain inline C++ by the test-only host build (tests/hostsim), which has no nvcc in the loop.
#if defined(__CUDACC__)
#define RT_HD __device__ __forceinline__
#define RT_D __device__ __forceinline__
#define RT_COLD __device__ __noinline__  // rarely executed float64 blocks: keep them out of the hot loop's registers
#else
#define RT_HD inline
#define RT_D inline
#define RT_COLD inline
#endif

struct alignas(16) RtF4 { float x, y, z, w; };
struct alignas(16) RtI4 { int x, y, z, w; };
struct alignas(16) RtD2 { double x, y; };
struct alignas(32) RtD4 { double x, y, z, w; };
struct RtD3 { double x, y, z; };
struct RtF3 { float x, y, z; };

// Packet-walk node record.  Nodes are renumbered breadth-first at upload, so the existing children of a
// node are consecutive: child of octant o = child_base + popc(child_mask & ((1 << o) - 1)).
struct alignas(32) RtPNode {
	float x, y, z, size;  // cube
	int list_off, list_cnt;
	int child_base, child_mask;
};

// Octree node record of the bounce stage's ordered walk, 64 bytes: the cube, the children (breadth-first
// numbering, as in RtPNode) and the way up, then the ROOT record of the node's list BVH (same layout as
// RtBvhNode), so that a node step tests the list's bounding box without another dependent fetch.
struct alignas(32) RtWNode {
	float x, y, z, size;
	int child_base, child_mask;
	int up;        // parent | index_within_parent << 28, -1: root
	int _pad;
	float lo[3], hi[3];  // box of the list's entities
	int a;         // list BVH root: inner: index of its child pair; leaf: first entry in bvh_slots / bvh_geom
	int b;         // 0: empty list; < 0 inner (-(lowest slot + 1)); > 0 leaf (entries)
};
#define RT_WNODE_PARENT_MASK 0x0fffffff

// Per-list bounding volume hierarchy.  The reference scans a node's entity list linearly and takes the FIRST
// entity (in insertion order) the ray hits; any structure that finds every hit entity of the list and keeps
// the lowest slot gives the same answer.  Every non-empty list has a binary BVH over the entities' float32
// boxes (inflated by err_l): a ray then tests a few dozen boxes instead of hundreds or thousands of entities
// of the big straddler lists near the root.  The bounce stage walks every list through its BVH (one uniform
// kind of step for the 32 independent rays of a warp); the ray-by-ray kernels use it for lists of
// RT_BVH_MIN_LIST or more entries and scan shorter ones linearly.
struct alignas(16) RtBvhNode {
	float lo[3], hi[3];
	int a;  // inner: index of the left child (right = a + 1), the left one holds the lowest slot;
	        // leaf: first entry in bvh_slots / bvh_geom (a multiple of RT_BVH_LEAF: leaves are padded to
	        // RT_BVH_LEAF entries, the padding carries slot RT_NO_SLOT)
	int b;  // inner: -(lowest slot below this node + 1) (< 0); leaf: number of entries (> 0), ascending slots
};
#define RT_BVH_MIN_LIST 24
#ifndef RT_BVH_LEAF
#define RT_BVH_LEAF 4  // entities per BVH leaf (2 or 4)
#endif
#define RT_BVH_STACK 40
#define RT_NO_SLOT 0x7fffffff

// material flags
#define RT_MAT_RESPONSE_MASK 3u
#define RT_MAT_LIGHT 4u
#define RT_MAT_MIRROR 8u

struct alignas(16) RtMaterial {
	uint32_t flags;
	uint32_t _pad;
	double roughness;
};

struct alignas(16) RtTexture {
	double r, g, b;      // SolidTexture colour or ImageTexture fallback
	uint32_t image;      // 1: loaded image -> texel lookup, 0: constant colour
	int32_t width, height;
	uint32_t _pad;
	uint64_t texel_off;  // first texel in the pool
};

// slot_attr.y packs material | type<<24
#define RT_ATTR_MAT_MASK 0x00ffffff
#define RT_ATTR_TYPE_SHIFT 24

struct RtDevScene {
	// octree, 64 B per node across three arrays
	const RtF4* node_geom;   // pos.xyz, size  (float copy of OctreeDim)
	const RtD4* node_geom64; // pos.xyz, size as the reference holds them (RT_PRECISION_F64: the float64 walker)
	const RtI4* node_link;   // parent, index_within_parent, list_off, list_cnt
	const int* node_child;   // [n*8], -1 none
	const RtPNode* node_pk;  // the same nodes as one 32-byte record each, for the packet walk
	const RtWNode* node_walk; // ... and for the bounce stage's ordered walk
	const int* node_bvh;     // [n] root of the node's list BVH in bvh_nodes, or -1 (short list)
	const RtBvhNode* bvh_nodes;
	const int* bvh_slots;    // leaf entries: slot numbers
	const RtF4* bvh_geom;    // leaf entries: copy of slot_geom in leaf order (no indirection in the leaf test)
	// entity lists, slot order
	const RtF4* slot_geom;   // centre.xyz, w = radius (>0, sphere) | -half_size (<0, box)
	const RtD4* slot_geom64; // centre.xyz, w = diameter | size  (the reference's float64 values)
	const RtI4* slot_attr;   // entity id, material|type<<24, texture, substance (-1 undefined)
	const int* slot_node;    // the node whose list holds the slot (hit_only_touches_its_cell)
	// tables
	const RtMaterial* materials;
	const RtTexture* textures;
	const double* substances;
	const uint8_t* texels;   // RGB8
	double root_pos[3];
	double root_size;
	int n_nodes, n_slots;
	int ordered_ok;          // 1: the tree depth fits RT_WALK_STACK, rays may use the ordered walk
	float err_l;             // bound on the float error of a point-to-line distance in this scene
	float _pad;
};

#define RT_RAYGEN_STRIDE 8  // yields between two checkpoints of the ray generation (a power of two)
// checkpoints per half row (the right half is the longer one when the width is odd)
inline int raygen_checkpoints_per_half(int width) { return ((width - (width >> 1)) + RT_RAYGEN_STRIDE - 1) / RT_RAYGEN_STRIDE; }
#define RT_MAX_CHAIN 32

// One pixel handed from the primary stage to the bounce stage: its path continues after the first hit
// (mirror / transmission), or the primary search itself has to be done per ray (slot == RT_SLOT_UNKNOWN).
struct alignas(8) RtQueueItem {
	uint32_t xy;  // y << 16 | x
	int32_t slot; // first-hit slot of the camera ray, or RT_SLOT_UNKNOWN
};
#define RT_SLOT_UNKNOWN (-2)
#define RT_NOT_QUEUED (-3)
// Node stack of the packet walk, in records per warp (shared memory): a packet usually keeps at most 3-4 pending
// siblings per level (the packet is narrow), so 96 records cover trees far deeper than float32 can resolve; a walk
// that would overflow hands its rays to the bounce stage instead (RT_SLOT_UNKNOWN), never drops a node silently.
#define RT_PACKET_STACK(ppl) 96
#define RT_WALK_STACK 160    // stack of the per-ray ordered walk (bounce stage): octree nodes (same bound) + one list BVH on top

struct RtFrame {
	// camera
	double pos[3];
	double lf[3];
	// Camera.get_dir_for_each_pixel (src/view/camera.ts:207-250) iterates 2-D rotations: rows outwards from the middle
	// row (fr towards up), then along every row outwards from the middle column (fr towards lf).  Floating-point
	// rotations do not commute with any closed form, so the directions are produced by the very same iteration:
	// row_fr on the host (height steps); the scans along the rows by the ray-generation lanes of the frame-setup kernel,
	// which keep the generator's state (fr, lf) of every RT_RAYGEN_STRIDE-th yield of each half row in `ray_ck`; a
	// pixel's direction is its checkpoint iterated the remaining 0..7 steps (rt_trace.cuh: pixel_dir) - the same
	// operations in the same order as the generator's, hence the same bits.  (A full table of directions would be
	// 32 B per pixel written and read per camera pose: the scan is then bound by the L2's request rate, not by its
	// own dependent chain.)
	const RtD4* row_fr;    // [height] fr rotated towards up by the accumulated vertical scan rotation (host-built)
	const double* ray_ck;  // [height][2 halves][ray_ckh] records of 6 doubles: fr.x, lf.x, fr.y, lf.y, fr.z, lf.z
	int ray_ckh;           // checkpoints per half row
	double scan_cos, scan_sin;  // rot_scan_h_v: cos / sin of fov_h / width
	int width, height;
	// start state shared by all primary rays (src/raytracer.ts:309-313)
	int start_node, start_octant;  // start_node < 0: camera outside the root cube
	int start_substance;
	// RaytracerConfig
	int refmax, sky_texture, default_substance;
	double attenuation;
	// exposure + rng
	uint32_t n_frames, frame_first;
	double w_first, w1_first;  // ExposureBuffer.col_weight of the first frame, 1/(1+frame_first), and 1 - it (host-computed: same IEEE ops)
	double rng_seed;
	// outputs
	float* rgb;
	int* first_ids;
	unsigned long long* counters;  // 8 x u64 or null
	uint32_t* error_flags;
	// pixel subset for tile-sharded rendering: tiles t with t % tile_world == tile_rank
	int tile_rank, tile_world;
	int tile_compact;  // 1: outputs are tile-major [own tile k][16*16] instead of [height][width]
	// band of the frame this launch covers: tiles [tile_begin, tile_end) (row-major tile numbers, whole tile
	// rows), so that a finished band can travel to the host while the next one is rendered
	int tile_begin, tile_end;
	unsigned long long out_first;  // index of the band's first output pixel (frame order; 0 when tile-compact)
	// primary-ray acceleration (exact: only skips work that provably cannot produce a hit)
	const RtF4* prim_geom;  // [n_slots] origin-relative records (make_prim_record), or null
	int chain_levels;       // origin chain: start_node, its parent, ..., root (post-order return order)
	int chain_beg[RT_MAX_CHAIN], chain_end[RT_MAX_CHAIN];  // slot ranges of the chain nodes' lists
	int chain_node[RT_MAX_CHAIN];  // the chain nodes themselves
	int chain_oct[RT_MAX_CHAIN];   // octant of chain_node[k] that holds the origin (k = 0: the start cell)
	int packet_ok;           // 1: camera rays may use the packet stage (chain complete, tree depth fits the stack)
	unsigned* work_counter;  // persistent-warp patch dispenser
	int* hit_slots;          // [pixels of this rank, output order] first-hit slots: primary stage -> shade stage
	// continuation queue between the shade stage and the bounce stage
	RtQueueItem* queue;      // [capacity]
	int* queue_dense;        // [pixels, output order] or null: per-pixel code (first-hit slot, RT_SLOT_UNKNOWN, RT_NOT_QUEUED) the primary stage leaves for rt_queue_compact_kernel
	unsigned* queue_count;   // items appended by the primary stage
	unsigned* queue_taken;   // consumer cursor of the bounce stage
	// resample queue between the bounce stage and the resample stage (n_frames > 1): the pixels whose path
	// draws from the RNG, i.e. whose every exposure frame is a different path
	RtQueueItem* vqueue;     // [capacity], or null: the bounce stage traces all frames of such a pixel itself
	unsigned* vqueue_count;
	unsigned* vqueue_taken;
	double* samples;         // resample stage: [pixels of a round][n_frames][3] path colours
	unsigned sample_chunk;   // pixels the table holds (a longer resample queue takes several rounds)
	int tie_checks;          // 1: exact ties are possible (rt_fill_frame): hits that only touch their cell are searched again by the float64 walker
	int search64;            // RT_PRECISION_F64: the cell-by-cell walker in float64 (walk_and_scan64), one ray per lane
	int bounce_min_walking;  // bounce stage: leave the lock-step walk when fewer lanes than this are still walking
	int bounce_node_batch;   // bounce stage: lanes that need a node step wait until this many do
	int bounce_sparse;       // ... unless fewer lanes than this are walking at all (the stage's tail)
};

#define RT_ERRFLAG_TEXTURE 1u
#define RT_ERRFLAG_ACUTE 2u
#define RT_ERRFLAG_STACK 4u  // a traversal stack was too small for a ray: the frame is refused (RT_ERR_UNSUPPORTED), never silently wrong
