/* rt_b200.h — C ABI of librt_b200.so, the B200 (sm_100a) drop-in for the per-pixel ray hot
 * path of Dark565/raytracer.js.
 *
 * What it replaces (paths relative to the reference tree):
 *   Raytracer.trace_frame()            src/raytracer.ts:308-330   -> rt_render*
 *   new Raytracer(config, otree, ...)  src/raytracer.ts:291-298   -> rt_create + rt_scene_upload
 *   RaytracerConfig                    src/raytracer.ts:33-43     -> rt_params
 *   Camera pose + CameraConfig         src/view/camera.ts:27-75   -> rt_camera
 *   ExposureBuffer.pixels / set_color  src/view/exposure_buffer.ts:27,68-91 -> the rgb in/out buffer
 *   EntityOtree (Octree<EntitySet,OctreeDim>) + Entity/Material/Texture/Substance objects
 *                                      src/octree.ts:25-126, src/octree_entity.ts:32-49,
 *                                      src/entities/..., src/material.ts:67-103, src/texture/...,
 *                                      src/substance.ts:1-11      -> rt_scene_desc (flat SoA)
 *
 * The host keeps building the octree exactly as today (add_entity_to_octree,
 * src/octree_entity.ts:174-188); a flattener walks it (DFS pre-order, children 0..7) and fills
 * rt_scene_desc.  The binding a maintainer adds on the reference side (N-API addon + the
 * GpuRaytracer TypeScript class) is shown in INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every array is caller-owned and COPIED during the
 * call; every function returns an rt_status (0 = ok) and never aborts the process; the message of
 * the last failure is rt_last_error(ctx) (ctx may be NULL for rt_create failures).  A ctx is not
 * thread-safe; distinct ctxs are independent.  There is no CPU fallback: without a CUDA device
 * rt_create fails with RT_ERR_CUDA.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_B200_ABI_VERSION 1

typedef struct rt_ctx rt_ctx;

typedef enum rt_status {
	RT_OK = 0,
	RT_ERR_INVALID = 1,     /* bad argument / inconsistent scene description */
	RT_ERR_CUDA = 2,        /* CUDA runtime failure (message has the CUDA error string) */
	RT_ERR_NO_SCENE = 3,    /* rt_render before rt_scene_upload */
	RT_ERR_UNSUPPORTED = 4, /* a feature the reference would accept but this path does not */
	RT_ERR_BOUNDS = 5,      /* "x or y out of bounds": non-square frame with RT_CAM_REFERENCE_EXTENTS
	                           (ExposureBuffer.check_bounds, src/view/exposure_buffer.ts:181-186) */
	RT_ERR_TEXTURE = 6      /* 'Texture coordinates out of bounds' (src/texture/texture_image.ts:49-50) */
} rt_status;

/* Entity kinds: SphereEntity (src/entities/entity_sphere.ts), BoxEntity (src/entities/entity_box.ts) */
#define RT_ENTITY_SPHERE 0u
#define RT_ENTITY_BOX 1u
/* ResponseType, src/material.ts:22-26 */
#define RT_RESPONSE_REFLECTION 0u
#define RT_RESPONSE_TRANSMISSION 1u
#define RT_RESPONSE_BOTH 2u
/* Texture kinds: SolidTexture (src/texture/texture_solid.ts), ImageTexture (src/texture/texture_image.ts) */
#define RT_TEXTURE_SOLID 0u
#define RT_TEXTURE_IMAGE 1u

/* The flattened scene.  Node 0 is the octree root handed to `new Raytracer(...)`, which must also be
 * the absolute root (tree.parent == undefined); nodes are numbered in DFS pre-order.  All reals are
 * the reference's float64 values, unchanged. */
typedef struct rt_scene_desc {
	uint32_t struct_size; /* sizeof(rt_scene_desc), for ABI checking */

	/* octree nodes: Octree.id = OctreeDim{pos,size} (src/octree_space.ts:30-33) */
	uint32_t n_nodes;
	const double* node_pos;        /* [n_nodes*3] */
	const double* node_size;       /* [n_nodes]   */
	const int32_t* node_child;     /* [n_nodes*8] child node index or -1 (Octree.get(n)) */
	const int32_t* node_parent;    /* [n_nodes]   -1 for the root */
	const int32_t* node_octant;    /* [n_nodes]   index_within_parent() (src/octree_space.ts:113-125), -1 root */
	const uint32_t* node_list_off; /* [n_nodes+1] CSR offsets into list_entity */
	uint32_t n_list;               /* node_list_off[n_nodes] */
	const uint32_t* list_entity;   /* [n_list] entity ids; each node's run is EntitySet.set in insertion order */

	/* entities; the id of an entity is its index here */
	uint32_t n_entities;
	const uint8_t* ent_type;       /* [n_entities] RT_ENTITY_* */
	const double* ent_pos;         /* [n_entities*3] get_pos() */
	const double* ent_extent;      /* [n_entities] get_diameter() | get_size() */
	const int32_t* ent_material;   /* [n_entities] */
	const int32_t* ent_texture;    /* [n_entities] */
	const int32_t* ent_substance;  /* [n_entities] index or -1 for `undefined` (src/raytracer.ts:243-248) */

	/* StaticMaterial fields, src/material.ts:73-81 */
	uint32_t n_materials;
	const uint8_t* mat_response;   /* RT_RESPONSE_* */
	const uint8_t* mat_light;      /* light_source */
	const uint8_t* mat_mirror;     /* mirror */
	const double* mat_roughness;   /* roughness_index */

	/* textures */
	uint32_t n_textures;
	const uint8_t* tex_kind;       /* RT_TEXTURE_* */
	const double* tex_color;       /* [n_textures*4] SolidTexture.color | ImageTexture.fallback_color (rgba) */
	const int32_t* tex_width;      /* image only */
	const int32_t* tex_height;     /* image only */
	const uint8_t* tex_loaded;     /* image only: 0 -> get_color answers the fallback colour */
	const uint64_t* tex_texel_off; /* image only: first texel of this image in `texels` (in texels) */
	uint64_t n_texels;
	const uint8_t* texels;         /* [n_texels*3] RGB8 as decoded by load_image (value/255.0 on use) */

	/* Substance.refractive_index, src/substance.ts:1-11 */
	uint32_t n_substances;
	const double* sub_refractive_index;
} rt_scene_desc;

/* rt_camera.flags */
#define RT_CAM_REFERENCE_EXTENTS 1u /* keep the reference's swapped scan extents (src/view/camera.ts:242-249):
                                       identical on square frames, RT_ERR_BOUNDS on non-square ones.
                                       Without it x runs over [0,width) and y over [0,height). */

/* Camera pose (private fields norm_fr/norm_lf/norm_up/pos/conf of src/view/camera.ts:51-59).  The
 * per-pixel direction is the generator's: fr rotated (y - (height>>1)) steps of fov_v/height towards up,
 * then (x - (width>>1)) steps of fov_h/width towards lf (src/view/camera.ts:207-250), un-normalised. */
typedef struct rt_camera {
	double pos[3];
	double fr[3], lf[3], up[3];
	double fov_h, fov_v;
	uint32_t width, height; /* CameraConfig.screen_w / screen_h == ExposureBuffer width / height */
	uint32_t flags;
	uint32_t _pad;
} rt_camera;

#define RT_PRECISION_F32 0u /* float search + float64 confirmation/shading of the found hit (default) */
/* float64 search: every ray is walked by the reference's own walker (OctreeWalker.next, src/octree_space.ts:316-361)
 * restated in float64 - cell by cell, same expressions - one ray per lane.  No packet stage and no lock-step walk:
 * several times slower, but without the float32 search's limit on the depth of the octree (cells smaller than a
 * float32 ulp of the coordinates, about 23 levels below a unit root, are refused with RT_PRECISION_F32). */
#define RT_PRECISION_F64 1u

/* rt_params.flags (validation / A-B measurement switches; results are identical either way) */
#define RT_PARAM_NO_PRIMARY_RECORDS 1u /* no per-frame origin-relative records: every ray takes the generic path */
#define RT_PARAM_PER_RAY 2u            /* camera rays are walked one by one too (no packet stage) */
/* Exact ties.  The packet walk and the ordered walk visit a superset of the nodes the reference's walker visits, which
 * cannot change a first hit unless a ray merely TOUCHES the cell of an entity it "hits" (a corner, an edge, a face plane
 * it runs inside): whether the walker visits such a cell is its half-open rule and its tie order, not geometry.  With
 * this flag - and by itself whenever the camera stands on a cell plane of the octree (the demo pose (0.5, 0.5, 0.5) does)
 * or the uploaded scene has an entity whose bounding cube has a face on a cell plane (scenes placed on a grid) - every
 * such hit is searched again by the float64 restatement of the walker, and rays with an exactly zero direction
 * component are walked by it from the start.  A few per cent slower; generic scenes and cameras never get it. */
#define RT_PARAM_EXACT_TIES 4u

/* RaytracerConfig (src/raytracer.ts:33-43) + exposure state + the harness RNG policy */
typedef struct rt_params {
	int32_t refmax;
	int32_t sky_texture;        /* SkySphere(texture) (src/sky/sky_sphere.ts); index into textures */
	int32_t default_substance;  /* index into substances */
	int32_t _pad0;
	double distance_attenuation_factor;
	/* n_frames consecutive trace_frame() calls, the first one at ExposureBuffer.frame_count ==
	 * frame_first (0 right after reset_exposure()), with next_frame() between them
	 * (src/view/exposure_buffer.ts:53-66; src/main.ts:210).  This is how the reference does spp > 1. */
	uint32_t n_frames;
	uint32_t frame_first;
	/* Rough materials draw from one shared sequential FpLcg in pixel-scan order in the reference
	 * (SURVEY.md F6), which no parallel renderer can reproduce.  This path reseeds per pixel through
	 * the public PRNG.seed(): seed = rng_seed + (y*width + x) + frame_count*width*height. */
	double rng_seed;
	uint32_t precision;         /* RT_PRECISION_* */
	uint32_t flags;
} rt_params;

/* Work counters of the last render (what the per-ray-bytes roofline is built from). */
typedef struct rt_counters {
	uint64_t paths;     /* pixels x frames */
	uint64_t segments;  /* traversals started (primary + continued bounces) */
	uint64_t nodes;     /* octree nodes returned by the walker order */
	uint64_t tests;     /* entity hit tests executed */
	uint64_t shades;    /* alter_ray calls */
	uint64_t confirms;  /* float64 confirmations of float candidates */
	uint64_t texture_errors;
	uint64_t acute_warnings; /* the console.warn path of src/raytracer.ts:200-203 */
} rt_counters;

/* rt_render flags */
#define RT_RENDER_COUNTERS 1u  /* collect rt_counters (slower kernel variant) */

/* ---- lifetime ------------------------------------------------------------------------------ */
/* device < 0: current CUDA device. */
rt_status rt_create(int32_t device, rt_ctx** out);
/* ONE process, n_gpus GPUs: the ctx handed back drives all of them behind the same entry points,
 * so that the reference's call site stays one `trace_frame()` (src/main.ts:408-414).  devices: n_gpus CUDA device
 * numbers, or NULL for 0 .. n_gpus-1 (a device may be named more than once: several members on one GPU, which is
 * how the sharding is tested on a one-GPU box).  Peer access is enabled between all of them.
 *   rt_scene_upload   validates and packs the scene ONCE on the host, uploads it to the first GPU and replicates it
 *                     from there device to device (cudaMemcpyPeerAsync over NVLink / NVSwitch);
 *   rt_render         every member renders its interleaved 16x16 tiles (tile t -> member t % n_gpus) and stores them
 *                     straight into the caller's host frame, page-locked and mapped into every GPU's address space,
 *                     over that GPU's own PCIe link; the launch sequences of the members are enqueued by one worker
 *                     thread per GPU; counters are summed;
 *   rt_render_device / rt_render_present
 *                     the frame lives on the first GPU, the others store their tiles into it over NVLink (peer
 *                     memory), the first GPU's stream waits for their events (no host round trip);
 *   rt_get_counters, rt_synchronize cover all members; everything else acts on the first GPU.
 * n_gpus == 1 is rt_create(devices ? devices[0] : 0). */
rt_status rt_create_multi(int32_t n_gpus, const int32_t* devices, rt_ctx** out);
uint32_t rt_group_size(const rt_ctx* ctx); /* GPUs (members) behind this ctx: 1 for rt_create */
void rt_destroy(rt_ctx* ctx);
const char* rt_last_error(const rt_ctx* ctx);
uint32_t rt_abi_version(void);
/* Use an existing CUDA stream (cudaStream_t as void*) for all work of this ctx; NULL = the ctx's own. */
rt_status rt_set_stream(rt_ctx* ctx, void* cuda_stream);

/* ---- scene ---------------------------------------------------------------------------------- */
/* Validates and copies the scene to the device (replaces any previous scene of the ctx). */
rt_status rt_scene_upload(rt_ctx* ctx, const rt_scene_desc* scene);

/* Dynamic scenes: move entities the way the reference does - BasicEntity._set_pos
 * (src/entities/entity_basic.ts:38-42) followed by add_entity_to_octree again (src/octree_entity.ts:174-188), whose
 * Entity.set_octree (src/entity.ts:50-56) takes the entity out of its node's Set and adds it to the END of the Set of the
 * node that now covers it; nodes are created on the way (max_in_depth as in AddEntityToOctreeFlags, max_out_depth = 0)
 * and never removed.  The library applies the moves, in the order given, to the copy of the description it kept at
 * rt_scene_upload - the caller passes only the moved entities, no re-flattening of its own tree - and refreshes the
 * device scene (of every GPU of a group).  entity_ids index the entity arrays of the uploaded description.  Where the
 * reference throws TreeOutsideGrowError the call fails with RT_ERR_UNSUPPORTED and the scene is unchanged. */
rt_status rt_scene_update(rt_ctx* ctx, uint32_t n_moved, const uint32_t* entity_ids, const double* new_pos /* [n_moved*3] */,
                          uint32_t max_in_depth);

/* ---- render: the trace_frame() drop-in ------------------------------------------------------- */
/* Host buffers.  rgb: float32 [height][width][3], the ExposureBuffer pixel store, read when
 * frame_first > 0 and always written.  first_ids (optional): int32 [height][width], the entity of
 * the first collision of each pixel's path in the last frame, -1 = none.  Synchronous. */
rt_status rt_render(rt_ctx* ctx, const rt_camera* cam, const rt_params* params, uint32_t render_flags,
                    float* rgb, int32_t* first_ids, rt_counters* counters);

/* Pipelined frames, for a host loop that renders frame after frame (the tick handler of src/main.ts:210,410-414 with
 * two ExposureBuffers): rt_render_begin enqueues a frame and returns at once; rt_render_end waits for the OLDEST frame
 * begun and not yet ended - its pixels (and ids) are then in the host buffers given to ITS rt_render_begin, which must
 * stay valid and untouched until then.  At most two frames in flight: the device-to-host copy of frame k (the larger
 * part of a 1080p frame's end-to-end time) overlaps the rendering of frame k + 1.  Counters need RT_RENDER_COUNTERS in
 * `flags`.  One GPU only (a group's members already deliver their tiles concurrently, over one PCIe link each). */
rt_status rt_render_begin(rt_ctx* ctx, const rt_camera* cam, const rt_params* params, uint32_t render_flags, float* rgb,
                          int32_t* first_ids);
rt_status rt_render_end(rt_ctx* ctx, rt_counters* counters);

/* Device buffers (same layouts), enqueued on the ctx stream, asynchronous.  For callers that keep
 * the ExposureBuffer resident on the GPU (tone mapping on device, multi-GPU tile exchange). */
rt_status rt_render_device(rt_ctx* ctx, const rt_camera* cam, const rt_params* params, uint32_t render_flags,
                           float* rgb_dev, int32_t* first_ids_dev);
/* Camera.get_dir_for_each_pixel (src/view/camera.ts:207-250) alone: the direction the generator yields for every
 * pixel, float64 [height][width][3], produced on the device by the same iterated rotations (bit for bit the
 * generator's values; un-normalised, as Raytracer.trace_frame passes them on, src/raytracer.ts:323-324).  Host
 * buffer, synchronous.  Every render call runs this pass itself; the entry point exists for checking it. */
rt_status rt_camera_directions(rt_ctx* ctx, const rt_camera* cam, double* dirs);
/* Counters of the last rt_render_device with RT_RENDER_COUNTERS (synchronises the stream). */
rt_status rt_get_counters(rt_ctx* ctx, rt_counters* counters);
rt_status rt_synchronize(rt_ctx* ctx);

/* ---- present: View.draw_ebuffer() (src/view/view.ts:34-38) on the device -------------------------- */
/* The step right after trace_frame(): exposure statistics (ExposureBuffer.get_mean / get_variance /
 * get_absolute_dev, src/view/exposure_buffer.ts:93-142), the tone mapper's dynamic range
 * (src/view/tone_mapping.ts:24-79) and the range compression to 8-bit pixels
 * (ExposureBuffer.discretize_to_screen, src/view/exposure_buffer.ts:145-158, into CanvasScreen.set_pixel_i /
 * convert_color, src/view/screen_canvas.ts:45-56,101-103), so that only RGBA8 leaves the GPU.  The output is
 * the CanvasScreen's ImageData: [height][width][4] bytes, alpha 255, and - as the reference is written at
 * HEAD (`pixels.slice(i, i+2)`) - blue 0.  The frame-wide sums are deterministic parallel float64
 * reductions (the reference adds the pixels one by one; the results agree to ~1e-13 relative). */
#define RT_TONE_IDENTITY 0u /* ToneMapper_Identity: [0, 1] */
#define RT_TONE_STDDEV 1u   /* ToneMapper_StdDevAroundMean */
#define RT_TONE_ABSDEV 2u   /* ToneMapper_AbsDevAroundMean */
typedef struct rt_tone {
	uint32_t kind;          /* RT_TONE_* */
	uint32_t dynamic_range; /* stops: dynamic_coef = 1 << dynamic_range (Screen.dynamic_range, 8 for a canvas) */
	double min_dynamic, max_dynamic;
} rt_tone;
typedef struct rt_exposure_stats {
	double mean, variance, absolute_dev; /* of the luma Y = 0.299 R + 0.587 G + 0.114 B */
	double drange_low, drange_high;      /* ToneMapper.get_dynamic_range */
} rt_exposure_stats;
/* Device buffers, asynchronous on the ctx stream.  rgb_dev: float32 [height][width][3]; rgba_dev: [height][width][4]. */
rt_status rt_present_device(rt_ctx* ctx, const float* rgb_dev, uint32_t width, uint32_t height, const rt_tone* tone,
                            uint8_t* rgba_dev);
/* Statistics and range of the last rt_present_device (synchronises the stream). */
rt_status rt_present_stats(rt_ctx* ctx, rt_exposure_stats* out);
/* Host buffers: H2D of the float frame, present, D2H of the RGBA8 image.  Synchronous.  stats may be NULL. */
rt_status rt_present(rt_ctx* ctx, const float* rgb, uint32_t width, uint32_t height, const rt_tone* tone, uint8_t* rgba,
                     rt_exposure_stats* stats);
/* trace_frame() + draw_ebuffer() with the ExposureBuffer RESIDENT on the device: the float frame stays in HBM
 * between calls (frame_first > 0 continues it; the first call of an exposure must use frame_first = 0) and only
 * the 8-bit image travels to the host.  Synchronous.  stats and counters may be NULL. */
rt_status rt_render_present(rt_ctx* ctx, const rt_camera* cam, const rt_params* params, uint32_t render_flags,
                            const rt_tone* tone, uint8_t* rgba, rt_exposure_stats* stats, rt_counters* counters);
/* Copies the resident ExposureBuffer to the host (float32 [height][width][3]).  Synchronous. */
rt_status rt_exposure_download(rt_ctx* ctx, float* rgb, uint32_t width, uint32_t height);

/* ---- multi-GPU tile sharding ------------------------------------------------------------------ */
/* The frame is cut into 16x16-pixel tiles numbered row-major; rank r of `world` owns the tiles t with
 * t % world == r (interleaved, to balance sky against dense regions).  One ctx (one process) per GPU
 * holds a replica of the scene.  rt_render_tiles_device writes this rank's tiles tile-major:
 * tiles_dev[k][16*16][3] for its k-th tile (t = r + k*world), rt_tiles_per_rank() tiles in all (the
 * last may be unused); exposure accumulation (frame_first > 0) stays local in that buffer.  After the
 * ranks' buffers were gathered rank-major (an NCCL all-gather / gather over NVLink, done by the
 * caller), rt_untile_device scatters them into the [height][width][3] frame. */
uint32_t rt_tiles_per_rank(uint32_t width, uint32_t height, uint32_t world);
rt_status rt_render_tiles_device(rt_ctx* ctx, const rt_camera* cam, const rt_params* params, uint32_t render_flags,
                                 uint32_t rank, uint32_t world, float* tiles_dev, int32_t* tile_ids_dev);
rt_status rt_untile_device(rt_ctx* ctx, uint32_t width, uint32_t height, uint32_t world, const float* gathered_dev,
                           float* rgb_dev);

/* ---- multi-GPU without a gather step: every rank writes its tiles straight into ONE frame over NVLink -- */
/* rt_render_shard_device renders the tiles of `rank` (t % world == rank) into a buffer in FRAME layout
 * ([height][width][3], ids [height][width]) and touches no other pixel.  frame_dev may be a peer-mapped
 * pointer into another GPU's memory (rt_peer_open): the kernels then store their pixels across NVLink /
 * NVSwitch as they are produced, so the "gather" is fused into the render and no tile exchange or
 * de-interleave pass is needed.  rt_peer_barrier closes the frame: it returns (in stream order) once every
 * rank's stores have arrived. */
rt_status rt_render_shard_device(rt_ctx* ctx, const rt_camera* cam, const rt_params* params, uint32_t render_flags,
                                 uint32_t rank, uint32_t world, float* frame_dev, int32_t* ids_dev);
#define RT_PEER_HANDLE_BYTES 64
/* cudaMalloc'ed, zero-filled device memory that other processes on this node can map (CUDA IPC). */
rt_status rt_peer_alloc(rt_ctx* ctx, size_t bytes, void** dev_ptr, unsigned char handle_out[RT_PEER_HANDLE_BYTES]);
rt_status rt_peer_free(rt_ctx* ctx, void* dev_ptr);
/* Map / unmap an allocation of another rank (its handle travels through any host channel). */
rt_status rt_peer_open(rt_ctx* ctx, const unsigned char handle[RT_PEER_HANDLE_BYTES], void** dev_ptr);
rt_status rt_peer_close(rt_ctx* ctx, void* dev_ptr);
/* All-ranks barrier on the ctx stream, one rank per GPU.  flags[r] = device pointer (own or peer-mapped) to
 * rank r's flag array of `world` uint32 (rt_peer_alloc gives zeroed memory); epoch must grow by one per
 * call, identically on every rank, starting at 1.  Everything enqueued before it on any rank's stream is
 * visible to whatever is enqueued after it on every rank. */
rt_status rt_peer_barrier(rt_ctx* ctx, uint32_t rank, uint32_t world, uint32_t* const* flags, uint32_t epoch);

/* ---- host-side helpers (no GPU work): texture ingest ---------------------------------------------- */
/* ImageTexture's decode step (src/texture/texture_image.ts:76-136) for hosts without a DOM: the reference draws the
 * file into a canvas and keeps the RGB bytes of getImageData (alpha dropped); here the bytes of a PNG (8 bits per
 * channel, non-interlaced: grey, grey + alpha, RGB, RGBA, palette), an uncompressed 24 / 32-bit BMP or a binary PPM
 * (P6) are decoded into RGB8, rows top to bottom - the layout of rt_scene_desc.texels.  *rgb is malloc'ed, release it
 * with rt_image_free.  RT_ERR_UNSUPPORTED for anything else (the caller then uses the texture's fallback colour, as
 * the reference does while an image is not loaded, :45-47); the message is rt_last_error(NULL). */
rt_status rt_image_decode(const uint8_t* bytes, uint64_t n_bytes, uint32_t* width, uint32_t* height, uint8_t** rgb);
void rt_image_free(uint8_t* rgb);

/* ---- host-side helpers (no GPU work): bulk scene construction ------------------------------------ */
/* Bulk restatement of add_entity_to_octree (src/octree_entity.ts:174-188) with max_out_depth = 0: inserts
 * the entities in index order into an empty octree rooted at (root_pos, root_size) and keeps the result in
 * an rt_tree.  Same nodes, same float64 node positions and same per-node insertion order as the host API
 * builds one entity at a time, without a pointer tree (1 M entities in well under a second).  Fails with
 * RT_ERR_UNSUPPORTED where the reference throws TreeOutsideGrowError (an entity that does not fit the root);
 * the message is rt_last_error(NULL). */
typedef struct rt_tree rt_tree;
rt_status rt_tree_build(const double root_pos[3], double root_size, uint32_t n_entities, const uint8_t* ent_type,
                        const double* ent_pos, const double* ent_extent, uint32_t max_in_depth, rt_tree** out);
/* The same builder on the GPU (max_in_depth <= 16): one thread per entity computes the path key of its node
 * with the reference's float64 expressions, the node set is the sort + unique of all path prefixes (which is
 * the depth-first pre-order, children 0..7), node positions are replayed per node, and a stable sort by node
 * keeps the insertion order inside every list.  Same tree as rt_tree_build (same nodes, positions, lists);
 * the node NUMBERING differs (pre-order here, creation order there), which rt_scene_upload does not care about.
 * Errors as rt_tree_build, message in rt_last_error(ctx). */
rt_status rt_tree_build_gpu(rt_ctx* ctx, const double root_pos[3], double root_size, uint32_t n_entities, const uint8_t* ent_type,
                            const double* ent_pos, const double* ent_extent, uint32_t max_in_depth, rt_tree** out);
uint32_t rt_tree_node_count(const rt_tree* t);
/* Fills the node arrays of an rt_scene_desc (sizes from rt_tree_node_count / n_entities); entity ids in
 * list_entity are the indices of the arrays given to rt_tree_build. */
void rt_tree_export(const rt_tree* t, double* node_pos, double* node_size, int32_t* node_child, int32_t* node_parent,
                    int32_t* node_octant, uint32_t* node_list_off, uint32_t* list_entity);
void rt_tree_free(rt_tree* t);
/* FpLcg (src/math/rng/fp-lcg.ts:62-82): seed(seed), then n x next() into out. */
void rt_fplcg_fill(double seed, uint64_t n, double* out);

/* ---- host buffer pinning ---------------------------------------------------------------------- */
/* Page-lock a caller-owned buffer (e.g. the ExposureBuffer's Float32Array backing store) so that
 * rt_render's copies run at full PCIe rate.  Optional; unregister before freeing the buffer. */
rt_status rt_host_register(rt_ctx* ctx, void* ptr, size_t bytes);
rt_status rt_host_unregister(rt_ctx* ctx, void* ptr);
/* Page-lock AND map a caller-owned host buffer into the device's address space (zero-copy): *dev_ptr can be
 * handed to rt_render_device / rt_render_shard_device as the frame, and the kernels then store their pixels
 * straight into host memory over PCIe.  With one process per GPU and the buffer in shared memory, every rank
 * delivers its own tiles over its own PCIe link - N links instead of one gather plus one copy.  The pixels are
 * complete on the host after rt_synchronize on every rank.  Release with rt_host_unregister. */
rt_status rt_host_map(rt_ctx* ctx, void* ptr, size_t bytes, void** dev_ptr);

/* ---- device timing on the ctx stream (CUDA events) ------------------------------------------- */
/* Benchmark hygiene: overwrite a scratch buffer larger than the 126 MB L2 on the ctx stream. */
rt_status rt_flush_l2(rt_ctx* ctx);
rt_status rt_timer_start(rt_ctx* ctx);
rt_status rt_timer_stop(rt_ctx* ctx, float* elapsed_ms); /* synchronises */
/* Number of kernels this library launched on the ctx since rt_create. */
uint64_t rt_launch_count(const rt_ctx* ctx);
/* Per-stage device time of a frame (what the reference's tick handler brackets as ONE trace_frame() with
 * performance.now(), src/main.ts:244-263, split by kernel): with profiling on, every non-replayed
 * rt_render_device / rt_render_shard_device call records CUDA events between its kernels on the ctx stream, and
 * rt_stage_times (synchronises) returns the milliseconds of each stage of the LAST such call, -1 for a stage that
 * did not run.  The events add a little launch gap: keep it off in the timed region of a benchmark. */
#define RT_STAGE_PREPARE 0  /* per-camera origin-relative records */
#define RT_STAGE_PRIMARY 1  /* camera rays: packet walk + shading of the paths that end at their first hit */
#define RT_STAGE_SHADE 2    /* refmax > 1: the continuation queue is put in output order (the primary stage itself shades) */
#define RT_STAGE_BOUNCE 3   /* continued paths */
#define RT_STAGE_RESAMPLE 4 /* (pixel, frame) samples of rough pixels, n_frames >= 4 */
#define RT_N_STAGES 5
rt_status rt_set_profiling(rt_ctx* ctx, int32_t on);
rt_status rt_stage_times(rt_ctx* ctx, float ms[RT_N_STAGES]);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
