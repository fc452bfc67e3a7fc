// rt_b200.cu — librt_b200.so: the C ABI of include/rt_b200.h and the sm_100a kernels behind it.
// No PyTorch, no CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

#include "rt_build.h"
#include "rt_build_gpu.cuh"
#include "rt_host.h"
#include "rt_image.h"
#include "rt_trace.cuh"

// ================================================================== kernels
#define RT_TILE_W 16
#define RT_TILE_H 16
#define RT_BLOCK (RT_TILE_W * RT_TILE_H)

// Ray generation: Camera.get_dir_for_each_pixel (src/view/camera.ts:207-250) for the whole frame.  The generator
// ITERATES a rotation along every row, outwards from the middle column; in floating point that recurrence has no
// closed form with the same bits, and after three mirror bounces off millimetre spheres a last-bit difference of a
// camera direction is a different path.  So the recurrence itself runs here.  rotate_vectors works component by
// component - (fr[k], lf[k]) <- (fr[k] c + lf[k] s, -fr[k] s + lf[k] c) -, so a half row is THREE independent
// recurrences: one lane per (row, half, component), 6 x height lanes, width / 2 dependent steps of 4 multiplications
// and 2 additions each (latency bound: ~25 cycles per step); every 8th state is kept (rt_common.h: RtFrame.ray_ck).
#define RT_SETUP_THREADS 96
RT_D void raygen_lane(const RtFrame& F, double* __restrict__ ck, int t) {
	if (t >= 6 * F.height) return;
	raygen_half_row_component(F, t / 6, (t / 3) & 1, t % 3, ck);
}
__global__ void __launch_bounds__(RT_SETUP_THREADS) rt_raygen_kernel(const __grid_constant__ RtFrame F, double* __restrict__ ck) {
	raygen_lane(F, ck, blockIdx.x * blockDim.x + threadIdx.x);
}
// the direction of every pixel, expanded from the checkpoints (rt_camera_directions)
__global__ void rt_expand_dirs_kernel(const __grid_constant__ RtFrame F, double* __restrict__ out) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (size_t)F.width * F.height) return;
	double d[3];
	pixel_dir(F, (int)(i % F.width), (int)(i / F.width), d);
	out[3 * i] = d[0]; out[3 * i + 1] = d[1]; out[3 * i + 2] = d[2];
}

// Everything a frame needs before its first ray, in ONE launch (three launches and their gaps would be a tenth of
// the headline frame): block 0 zeroes the control cells (work counters, error flags, dispensers, queue cursors); the
// next raygen_blocks blocks are the ray generation (skipped when the camera's basis has not changed: 0 blocks); the
// rest compute the per-frame origin-relative copy of the slot geometry for camera rays (centre - camera in float64,
// no cancellation, rounded once; skipped without primary records: 0 blocks).
__global__ void __launch_bounds__(RT_SETUP_THREADS)
    rt_frame_setup_kernel(const __grid_constant__ RtDevScene S, const __grid_constant__ RtFrame F, unsigned long long* __restrict__ cells,
                          int n_cells, int raygen_blocks, double* __restrict__ ray_ck, RtF4* __restrict__ prim_out) {
	int b = (int)blockIdx.x;
	if (b == 0) {
		for (int i = threadIdx.x; i < n_cells; i += RT_SETUP_THREADS) cells[i] = 0ull;
		return;
	}
	b -= 1;
	if (b < raygen_blocks) {
		raygen_lane(F, ray_ck, b * RT_SETUP_THREADS + (int)threadIdx.x);
		return;
	}
	const int s = (b - raygen_blocks) * RT_SETUP_THREADS + (int)threadIdx.x;
	if (s >= S.n_slots) return;
	prim_out[s] = make_prim_record(ld(S.slot_geom64 + s), ld(S.slot_geom + s).w > 0.0f, F.pos[0], F.pos[1], F.pos[2], S.err_l);
}

// Persistent warps: the grid is sized to the resident capacity of the GPU (SMs x CTAs/SM) and every
// warp pulls 8x4-pixel patches from an atomic counter until the frame (or this rank's share of its
// 16x16 tiles, t % tile_world == tile_rank) is done.  One thread per pixel; the 32 rays of a patch
// walk nearly the same octree cells and scan the same entity lists (uniform LDG.128 addresses), and a
// warp that finishes a cheap sky patch immediately takes the next one instead of idling in its CTA.
#define RT_WARPS_PER_CTA 4
template <bool COUNT>
__global__ void __launch_bounds__(RT_WARPS_PER_CTA * 32, 4)
    rt_render_kernel(const __grid_constant__ RtDevScene S, const __grid_constant__ RtFrame F, int tiles_x, int n_patches) {
	const int lane = threadIdx.x & 31;
	RtCounts cnt = {0, 0, 0, 0, 0};
	unsigned long long paths = 0;
	uint32_t err = 0;
	while (true) {
		unsigned p = 0;
		if (lane == 0) p = atomicAdd(F.work_counter, 1u);
		p = __shfl_sync(0xffffffffu, p, 0);
		if (p >= (unsigned)n_patches) break;
		const int k = (int)(p >> 3), sub = (int)(p & 7);  // k-th own tile, patch within the tile
		const int tile = F.tile_begin + F.tile_rank + k * F.tile_world;
		if (tile >= F.tile_end) continue;
		const int tx = tile % tiles_x, ty = tile / tiles_x;
		const int x = tx * RT_TILE_W + (sub & 1) * 8 + (lane & 7);
		const int y = ty * RT_TILE_H + (sub >> 1) * 4 + (lane >> 3);
		if (x < F.width && y < F.height) {
			const size_t out_index =
			    F.tile_compact ? (size_t)k * RT_BLOCK + ((y & (RT_TILE_H - 1)) * RT_TILE_W + (x & (RT_TILE_W - 1)))
			                   : (size_t)y * F.width + x;
			render_pixel<COUNT>(S, F, x, y, out_index, cnt, err);
			paths += F.n_frames;
		}
		__syncwarp();
	}
	if (COUNT) {
		unsigned long long v[6] = {paths, cnt.segments, cnt.nodes, cnt.tests, cnt.shades, cnt.confirms};
#pragma unroll
		for (int k = 0; k < 6; k++) {
			unsigned long long s = v[k];
			for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
			if (lane == 0 && s) atomicAdd(F.counters + k, s);
		}
	}
	if (err) atomicOr(F.error_flags, err);
}

// ---- the pipeline (default): primary stage = packet walk + shading of the paths that end at their first hit,
// bounce stage = the continued paths, ray by ray
// Primary stage.  Same persistent-warp dispenser as above, handing out packets of PPL 8x4 sub-patches of a 16x16
// tile; every warp walks the octree ONCE for the 32 x PPL camera rays of its packet (rt_trace.cuh:
// packet_primary_hits), shades and stores the pixels whose path ends at the first hit and appends the others to
// the continuation queue (primary_patch).  Per warp in shared memory: the node-record stack of the walk, the
// packet's ray table (32 x PPL x 16 B: rays are touched only when a list entry survives the packet's cone test,
// and keeping them out of the registers is what lets 5-6 CTAs be resident per SM), and the staging of a
// sub-patch's pixels for 16-byte stores.
#define RT_A_WARPS 4
#ifndef RT_PPL
#define RT_PPL 4  // rays per lane in the primary stage: a packet is RT_PPL 8x4 sub-patches of a 16x16 tile
#endif
#ifndef RT_A_MINB
#define RT_A_MINB 5
#endif
#ifdef RT_WALK_TIMELINE  // tools/ only: when do the warps of the bounce stage run out of queue, and when do they finish
#define RT_PROF_WARPS 8192
__device__ unsigned long long g_prof_start[RT_PROF_WARPS], g_prof_exhaust[RT_PROF_WARPS], g_prof_exit[RT_PROF_WARPS];
__device__ __forceinline__ unsigned long long prof_now() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
#endif
#ifdef RT_WALK_TIMELINE  // tools/ only: start and duration of every packet of the primary stage
#define RT_PROF_PACKETS 65536
__device__ unsigned long long g_pk_start[RT_PROF_PACKETS];
__device__ unsigned int g_pk_dur[RT_PROF_PACKETS];
#endif
template <int PPL, int MINB>
__global__ void __launch_bounds__(RT_A_WARPS * 32, MINB)
    rt_primary_kernel(const __grid_constant__ RtDevScene S, const __grid_constant__ RtFrame F, int tiles_x, int n_packets) {
	__shared__ RtPNode stacks[RT_A_WARPS][RT_PACKET_STACK(PPL)];
	__shared__ __align__(16) RtPRay rays[RT_A_WARPS][PPL * 32];
	__shared__ __align__(16) double dirs[RT_A_WARPS][PPL * 32 * 3];
	__shared__ __align__(16) float stages[RT_A_WARPS][96];
	constexpr int PER_TILE = 8 / PPL;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t err = 0;
	while (true) {
		unsigned p = 0;
		if (lane == 0) p = atomicAdd(F.work_counter, 1u);
		p = __shfl_sync(0xffffffffu, p, 0);
		if (p >= (unsigned)n_packets) break;
		const int k = (int)(p / PER_TILE);
		const int tile = F.tile_begin + F.tile_rank + k * F.tile_world;
		if (tile >= F.tile_end) continue;
		RtPatch pt;
		pt.x0 = (tile % tiles_x) * RT_TILE_W;
		pt.y0 = (tile / tiles_x) * RT_TILE_H;
		pt.sub0 = (int)(p % PER_TILE) * PPL;
		pt.out_base = (size_t)k * RT_BLOCK;
#ifdef RT_WALK_TIMELINE
		const unsigned long long pk_t0 = prof_now();
#endif
		primary_patch<PPL>(S, F, pt, stacks[warp], rays[warp], dirs[warp], stages[warp], err);
#ifdef RT_WALK_TIMELINE
		if (lane == 0 && p < RT_PROF_PACKETS) {
			g_pk_start[p] = pk_t0;
			g_pk_dur[p] = (unsigned)(prof_now() - pk_t0);
		}
#endif
	}
	if (err) atomicOr(F.error_flags, err);
}

// The continuation queue in OUTPUT ORDER (refmax > 1): the primary stage left one code per pixel; one thread per pixel,
// one warp-aggregated atomic per 32 pixels.  Blocks start in index order, so the queue comes out (nearly) sorted by
// pixel: a bounce-stage warp's 32 rays, and the warps beside it, start from neighbouring pixels - which is what keeps
// their node and list fetches in the L1 (measured: appending in packet-completion order cost the bounce stage 17 %).
__global__ void __launch_bounds__(256)
    rt_queue_compact_kernel(const __grid_constant__ RtFrame F, int tiles_x, size_t n_out) {
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int lane = threadIdx.x & 31;
	int code = RT_NOT_QUEUED, x = 0, y = 0;
	if (t < n_out) {
		size_t i = F.out_first + t;
		bool valid = true;
		if (F.tile_compact || F.tile_world > 1) {
			const int k = (int)(t / RT_BLOCK), in = (int)(t % RT_BLOCK);
			const int tile = F.tile_begin + F.tile_rank + k * F.tile_world;
			valid = tile < F.tile_end;
			x = (tile % tiles_x) * RT_TILE_W + (in & (RT_TILE_W - 1));
			y = (tile / tiles_x) * RT_TILE_H + (in / RT_TILE_W);
			valid = valid && x < F.width && y < F.height;
			i = F.tile_compact ? t : (size_t)y * F.width + x;
		} else {
			y = (int)(i / (size_t)F.width);
			x = (int)(i - (size_t)y * F.width);
		}
		if (valid) code = F.queue_dense[i];
	}
	const bool enqueue = code != RT_NOT_QUEUED;
	const unsigned m = __ballot_sync(0xffffffffu, enqueue);
	if (m) {
		unsigned base = 0;
		if (lane == 0) base = atomicAdd(F.queue_count, (unsigned)__popc(m));
		base = __shfl_sync(0xffffffffu, base, 0);
		if (enqueue) F.queue[base + __popc(m & ((1u << lane) - 1u))] = RtQueueItem{((uint32_t)y << 16) | (uint32_t)x, code};
	}
}

// Bounce stage: persistent wavefront with a per-warp ray queue.  Every lane owns one pixel job (all its
// exposure frames) and is in one of four states:
//   IDLE   no job: refilled from the continuation queue (one atomic per refill, ranks by popc of the vote);
//   BEGIN  a ray segment starts: walker re-seed (segment_begin);
//   WALK   the ordered walk, run by the whole warp in lock-step (walk_iter: every lane does one node step
//          and/or one list-BVH step per iteration, phases re-converged);
//   END    the search is over: collision + the material's response (segment_end), which starts the next
//          segment (BEGIN), the next exposure frame, or ends the job (IDLE).
// The warp leaves the walk loop as soon as a quarter of the lanes that entered it have finished (or fewer
// than F.bounce_min_walking are left), so that a path that bounces four times or crosses a long list does
// not hold finished lanes hostage: those shade, start their next segment or fetch a new pixel, and re-join.
// Frames whose path never drew from the RNG reuse the first frame's sample.
#define RT_ST_IDLE 0
#define RT_ST_BEGIN 1
#define RT_ST_WALK 2
#define RT_ST_END 3
template <int MINB, bool TIES>
__global__ void __launch_bounds__(RT_WARPS_PER_CTA * 32, MINB)
    rt_bounce_kernel(const __grid_constant__ RtDevScene S, const __grid_constant__ RtFrame F, int tiles_x) {
	const int lane = threadIdx.x & 31;
	const unsigned lt_mask = (1u << lane) - 1u;
	const unsigned n = *F.queue_count;
	RtCounts cnt = {0, 0, 0, 0, 0};
	uint32_t err = 0;
	// the lane's pixel job
	int st = RT_ST_IDLE;
	int x = 0, y = 0, slot = RT_SLOT_UNKNOWN;
	size_t out_index = 0;
	uint32_t frame = 0;
	float px[3] = {0.f, 0.f, 0.f};
	RtPath P = {};  // (set by path_begin before any use; value-initialised to keep the compiler quiet)
	RtWalk W;
	bool exhausted = n == 0;
#ifdef RT_WALK_TIMELINE
	const unsigned prof_warp = min((unsigned)RT_PROF_WARPS - 1u, blockIdx.x * RT_WARPS_PER_CTA + (threadIdx.x >> 5));
	if (lane == 0) g_prof_start[prof_warp] = prof_now();
	bool prof_seen_exhausted = false;
#endif
	// a sample (path colour c) of exposure frame `frame` is complete: ExposureBuffer.set_color_i
	// (src/view/exposure_buffer.ts:77-91) for this frame, and for all remaining ones when the path cannot
	// change (it never drew from the RNG); then the next frame's path, or the end of the job
	auto sample_done = [&](const double* c) {
		const bool varies = P.rng.seeded;
		if (varies && F.vqueue && F.n_frames > 1) {
			// every frame of this pixel is a different path: the resample stage traces them as independent samples.
			// This one - the first frame's - goes into the sample table right away (when the pixel falls into the
			// table's first round), so that it is not traced a second time.
			const unsigned qi = atomicAdd(F.vqueue_count, 1u);
			F.vqueue[qi] = RtQueueItem{((uint32_t)y << 16) | (uint32_t)x, slot};
			if (qi < F.sample_chunk) {
				double* o = F.samples + (size_t)qi * F.n_frames * 3;
				o[0] = c[0]; o[1] = c[1]; o[2] = c[2];
				if (F.first_ids) F.first_ids[out_index] = P.first_entity;
			}
			st = RT_ST_IDLE;
			return;
		}
		const uint32_t last = varies ? frame + 1 : F.n_frames;
		for (; frame < last; frame++) {
			const double w = xdiv(1.0, (double)(1u + F.frame_first + frame));
			const double w1 = xsub(1.0, w);
#pragma unroll
			for (int k = 0; k < 3; k++) px[k] = (float)xadd(xmul(c[k], w), xmul((double)px[k], w1));
		}
		if (frame < F.n_frames) {
			double dir[3];
			pixel_dir(F, x, y, dir);
			path_begin(F, dir, P);
			st = RT_ST_BEGIN;
		} else {
			float* o = F.rgb + out_index * 3;
			o[0] = px[0]; o[1] = px[1]; o[2] = px[2];
			if (F.first_ids) F.first_ids[out_index] = P.first_entity;
			st = RT_ST_IDLE;
		}
	};
	while (true) {
		// ---- refill the idle lanes
		const unsigned idle = __ballot_sync(0xffffffffu, st == RT_ST_IDLE);
		if (idle && !exhausted) {
			const unsigned want = (unsigned)__popc(idle);
			unsigned base = 0;
			if (lane == 0) base = atomicAdd(F.queue_taken, want);
			base = __shfl_sync(0xffffffffu, base, 0);
			exhausted = base + want >= n;
			const unsigned i = base + (unsigned)__popc(idle & lt_mask);
			if (st == RT_ST_IDLE && i < n) {
				const RtQueueItem it = F.queue[i];
				x = (int)(it.xy & 0xffffu);
				y = (int)(it.xy >> 16);
				slot = it.slot;
				out_index = (size_t)y * F.width + x;
				if (F.tile_compact) {
					const int tile = (y / RT_TILE_H) * tiles_x + (x / RT_TILE_W);
					out_index = (size_t)(tile / F.tile_world) * RT_BLOCK + ((y & (RT_TILE_H - 1)) * RT_TILE_W + (x & (RT_TILE_W - 1)));
				}
				frame = 0;
				px[0] = px[1] = px[2] = 0.f;
				if (F.frame_first > 0) {
					const float* o = F.rgb + out_index * 3;
					px[0] = o[0]; px[1] = o[1]; px[2] = o[2];
				}
				double dir[3];
				pixel_dir(F, x, y, dir);
				path_begin(F, dir, P);
				st = RT_ST_BEGIN;
			}
		}
		if (__ballot_sync(0xffffffffu, st != RT_ST_IDLE) == 0u) break;
#ifdef RT_WALK_TIMELINE
		if (exhausted && !prof_seen_exhausted) {
			prof_seen_exhausted = true;
			if (lane == 0) g_prof_exhaust[prof_warp] = prof_now();
		}
#endif
		RT_PROF_COUNT(9);                        // outer passes
		RT_PROF_LANES(10, st == RT_ST_BEGIN);
		RT_PROF_LANES(11, st == RT_ST_IDLE);
		RT_PROF_LANES(12, st == RT_ST_END);
		// ---- BEGIN: walker re-seed
		if (st == RT_ST_BEGIN) {
			double c[3];
			int hit = -1;
			RtCollision ci;
			const int r = segment_begin<false, TIES>(S, F, P, slot, c, cnt, err, S.ordered_ok ? &W : nullptr, hit, ci);
			if (r == RT_SEG_DONE) {
				sample_done(c);
			} else if (r == RT_SEG_WALK) {
				st = RT_ST_WALK;
			} else {
				W.hit = hit;  // known from the primary stage (or searched by the fallback walker)
				st = RT_ST_END;
			}
		}
		__syncwarp();
		// ---- WALK in lock-step
		{
			bool walking = st == RT_ST_WALK;
			int nw = __popc(__ballot_sync(0xffffffffu, walking));
			if (nw > 0) {
				const int limit = max(1, min(F.bounce_min_walking, nw - (nw >> 2)));
				do {
#ifdef RT_WALK_PROFILE
					if (!exhausted) {
						RT_PROF_COUNT(14);
						RT_PROF_LANES(15, walking);
					}
#endif
					// (a warp that is running empty - the stage's tail - batches nothing: its few paths are the critical path)
					const bool sparse = nw < F.bounce_sparse;
					walking = walk_iter<true>(S, W, P.refpoint, P.dir, walking, sparse ? 1 : F.bounce_node_batch, sparse ? 1 : RT_LEAF_DEFER);
					nw = __popc(__ballot_sync(0xffffffffu, walking));
				} while (nw >= limit);
				if (st == RT_ST_WALK && !walking) st = RT_ST_END;
			}
		}
		RT_PROF_LANES(13, st == RT_ST_END);
		// ---- END: collision, material response
		if (st == RT_ST_END) {
			const uint32_t frame_count = F.frame_first + frame;
			const double seed = xadd(xadd(F.rng_seed, (double)((size_t)y * F.width + x)),
			                         xmul(xmul((double)frame_count, (double)F.width), (double)F.height));
			double c[3];
			int hit;
			RtCollision ci;
			segment_found<TIES>(S, F, P, W, hit, ci, err);
			if (segment_end<false>(S, F, P, seed, hit, ci, c, cnt, err)) sample_done(c);
			else st = RT_ST_BEGIN;
		}
		__syncwarp();
	}
#ifdef RT_WALK_TIMELINE
	if (lane == 0) g_prof_exit[prof_warp] = prof_now();
#endif
	if (err) atomicOr(F.error_flags, err);
}

// Resample stage (n_frames > 1).  A pixel whose path scatters off a rough surface is a different path in
// every exposure frame, and ExposureBuffer.set_color_i (src/view/exposure_buffer.ts:77-91) blends the frames
// one after the other into a float32 pixel.  One lane tracing all frames of such a pixel in a row is the
// longest job of the whole render (n_frames x bounces walks) and would be its tail on any number of GPUs.
// Here the job is one (pixel, frame) SAMPLE: the samples of all queued pixels are one stream of jobs
// (job j = frame j % n_frames of pixel j / n_frames, so neighbouring lanes start on the same ray), the lanes of
// the persistent warps run the bounce stage's state machine over it - IDLE lanes take the next samples, BEGIN /
// lock-step WALK / END as in rt_bounce_kernel - and park every path colour as float64 in the frame's sample table;
// rt_resample_blend_kernel then blends the samples of each pixel in frame order, so the pixel is the sequential
// one, bit for bit.  (Round 1 gave every warp its own pool of 256 samples and blended when the pool was done: each
// pool ended with its own tail of a few long paths in a nearly empty warp.)  The table holds `chunk` pixels; a queue
// longer than that takes several rounds of the two kernels (`first` = the round's first queued pixel).
RT_HD size_t frame_out_index(const RtFrame& F, int x, int y, int tiles_x) {
	if (!F.tile_compact) return (size_t)y * F.width + x;
	const int tile = (y / RT_TILE_H) * tiles_x + (x / RT_TILE_W);
	return (size_t)(tile / F.tile_world) * RT_BLOCK + ((y & (RT_TILE_H - 1)) * RT_TILE_W + (x & (RT_TILE_W - 1)));
}

template <int MINB, bool TIES>
__global__ void __launch_bounds__(RT_WARPS_PER_CTA * 32, MINB)
    rt_resample_kernel(const __grid_constant__ RtDevScene S, const __grid_constant__ RtFrame F, int tiles_x, unsigned first, unsigned chunk) {
	const int lane = threadIdx.x & 31;
	const unsigned lt_mask = (1u << lane) - 1u;
	const unsigned queued = *F.vqueue_count;
	if (queued <= first) return;
	const unsigned nf = F.n_frames;
	// the first round's pixels come with their first frame's sample (the bounce stage traced it): nf - 1 jobs per pixel
	const unsigned per_pixel = first == 0 ? nf - 1 : nf;
	const unsigned n = min(queued - first, chunk) * per_pixel;  // samples of this round (chunk * nf < 2^32, launch_render)
	RtCounts cnt = {0, 0, 0, 0, 0};
	uint32_t err = 0;
	int st = RT_ST_IDLE;
	unsigned job = 0;
	int x = 0, y = 0, slot = RT_SLOT_UNKNOWN;
	RtPath P = {};
	RtWalk W;
	bool exhausted = false;
	auto sample_done = [&](const double* c) {
		double* o = F.samples + (size_t)job * 3;
		o[0] = c[0]; o[1] = c[1]; o[2] = c[2];
		if (job % nf == 0 && F.first_ids) F.first_ids[frame_out_index(F, x, y, tiles_x)] = P.first_entity;
		st = RT_ST_IDLE;
	};
	while (true) {
		// ---- idle lanes take the next samples
		const unsigned idle = __ballot_sync(0xffffffffu, st == RT_ST_IDLE);
		if (idle && !exhausted) {
			const unsigned want = (unsigned)__popc(idle);
			unsigned base = 0;
			if (lane == 0) base = atomicAdd(F.vqueue_taken, want);
			base = __shfl_sync(0xffffffffu, base, 0);
			exhausted = base + want >= n;
			const unsigned j = base + (unsigned)__popc(idle & lt_mask);
			if (st == RT_ST_IDLE && j < n) {
				const unsigned pixel = j / per_pixel;
				job = pixel * nf + (nf - per_pixel) + j % per_pixel;  // place in the sample table: [pixel][frame]
				const RtQueueItem it = F.vqueue[first + pixel];
				x = (int)(it.xy & 0xffffu);
				y = (int)(it.xy >> 16);
				slot = it.slot;
				double dir[3];
				pixel_dir(F, x, y, dir);
				path_begin(F, dir, P);
				st = RT_ST_BEGIN;
			}
		}
		if (__ballot_sync(0xffffffffu, st != RT_ST_IDLE) == 0u) break;
		// ---- BEGIN: walker re-seed
		if (st == RT_ST_BEGIN) {
			double c[3];
			int hit = -1;
			RtCollision ci;
			const int r = segment_begin<false, TIES>(S, F, P, slot, c, cnt, err, S.ordered_ok ? &W : nullptr, hit, ci);
			if (r == RT_SEG_DONE) {
				sample_done(c);
			} else if (r == RT_SEG_WALK) {
				st = RT_ST_WALK;
			} else {
				W.hit = hit;
				st = RT_ST_END;
			}
		}
		__syncwarp();
		// ---- WALK in lock-step
		{
			bool walking = st == RT_ST_WALK;
			int nw = __popc(__ballot_sync(0xffffffffu, walking));
			if (nw > 0) {
				const int limit = max(1, min(F.bounce_min_walking, nw - (nw >> 2)));
				do {
					// (a warp that is running empty - the stage's tail - batches nothing: its few paths are the critical path)
					const bool sparse = nw < F.bounce_sparse;
					walking = walk_iter<true>(S, W, P.refpoint, P.dir, walking, sparse ? 1 : F.bounce_node_batch, sparse ? 1 : RT_LEAF_DEFER);
					nw = __popc(__ballot_sync(0xffffffffu, walking));
				} while (nw >= limit);
				if (st == RT_ST_WALK && !walking) st = RT_ST_END;
			}
		}
		// ---- END: collision, material response
		if (st == RT_ST_END) {
			const uint32_t frame_count = F.frame_first + job % nf;
			const double seed = xadd(xadd(F.rng_seed, (double)((size_t)y * F.width + x)),
			                         xmul(xmul((double)frame_count, (double)F.width), (double)F.height));
			double c[3];
			int hit;
			RtCollision ci;
			segment_found<TIES>(S, F, P, W, hit, ci, err);
			if (segment_end<false>(S, F, P, seed, hit, ci, c, cnt, err)) sample_done(c);
			else st = RT_ST_BEGIN;
		}
		__syncwarp();
	}
	if (err) atomicOr(F.error_flags, err);
}

// ExposureBuffer.set_color_i (src/view/exposure_buffer.ts:77-91), frame after frame, for the pixels of one round of
// the resample stage: one thread per queued pixel.  Re-arms the sample counter for the next round.
__global__ void rt_resample_blend_kernel(const __grid_constant__ RtFrame F, int tiles_x, unsigned first, unsigned chunk) {
	const unsigned queued = *F.vqueue_count;
	const unsigned nf = F.n_frames;
	const unsigned n = queued <= first ? 0u : min(queued - first, chunk);
	for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const RtQueueItem it = F.vqueue[first + i];
		const size_t out_index = frame_out_index(F, (int)(it.xy & 0xffffu), (int)(it.xy >> 16), tiles_x);
		float* o = F.rgb + out_index * 3;
		float px[3] = {0.f, 0.f, 0.f};
		if (F.frame_first > 0) { px[0] = o[0]; px[1] = o[1]; px[2] = o[2]; }
		const double* c = F.samples + (size_t)i * nf * 3;
		for (unsigned f = 0; f < nf; f++, c += 3) {
			const double w = xdiv(1.0, (double)(1u + F.frame_first + f));
			const double w1 = xsub(1.0, w);
#pragma unroll
			for (int q = 0; q < 3; q++) px[q] = (float)xadd(xmul(c[q], w), xmul((double)px[q], w1));
		}
		o[0] = px[0]; o[1] = px[1]; o[2] = px[2];
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) *F.vqueue_taken = 0u;
}

// Tile-major buffers of all ranks, concatenated [world][tiles_per_rank][16*16][3]  ->  frame [H][W][3].
__global__ void rt_untile_kernel(const float* __restrict__ gathered, float* __restrict__ rgb, int width, int height,
                                 int tiles_x, int world, int tiles_per_rank) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= width || y >= height) return;
	const int tile = (y / RT_TILE_H) * tiles_x + (x / RT_TILE_W);
	const int rank = tile % world, k = tile / world;
	const size_t src = ((size_t)(rank * tiles_per_rank + k) * RT_BLOCK + (y % RT_TILE_H) * RT_TILE_W + (x % RT_TILE_W)) * 3;
	const size_t dst = ((size_t)y * width + x) * 3;
	rgb[dst] = gathered[src];
	rgb[dst + 1] = gathered[src + 1];
	rgb[dst + 2] = gathered[src + 2];
}

// All-ranks barrier over peer-mapped flags (one rank per GPU, one CTA of `world` threads per rank).
// Thread t tells rank t "rank `rank` has reached epoch e" with a system-scope release store into rank t's
// flag array, then waits until rank t has told us the same.  The fence orders every store this GPU issued
// before (the pixels written into the peer's frame) ahead of the flag.
__global__ void rt_peer_barrier_kernel(uint32_t* const* __restrict__ flags, int rank, int world, uint32_t epoch) {
	const int t = threadIdx.x;
	if (t >= world) return;
	__threadfence_system();
	uint32_t* theirs = flags[t] + rank;
	asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
	const uint32_t* mine = flags[rank] + t;
	uint32_t v;
	do {
		asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
	} while ((int32_t)(v - epoch) < 0);
}

// ---- present: View.draw_ebuffer() (src/view/view.ts:34-38) ------------------------------------------
// Three HBM-bound passes over the float32 frame (12 B/pixel read each; the frame stays in the 126 MB L2
// between them up to 4K): luma sum, deviation sums, range compression to RGBA8.  Every block reduces its
// share in a fixed order and every consumer adds the block partials in index order, so the results do not
// depend on scheduling.
#define RT_PRESENT_THREADS 256
RT_D double luma(const float* px) {  // ExposureBuffer.rgb_to_y (src/view/exposure_buffer.ts:161-173), float64, no FMA
	return xadd(xadd(xmul(0.299, (double)px[0]), xmul(0.587, (double)px[1])), xmul(0.114, (double)px[2]));
}
RT_D double block_sum(double v, double* sh) {  // fixed tree: xor shuffles inside a warp, warps in order
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
	__syncthreads();
	double s = 0.0;
	for (int w = 0; w < RT_PRESENT_THREADS / 32; w++) s += sh[w];
	__syncthreads();
	return s;
}
RT_D double partial_sum(const double* partial, int n_blocks, double* sh) {  // the same value in every block
	double v = 0.0;
	for (int i = threadIdx.x; i < n_blocks; i += RT_PRESENT_THREADS) v += partial[i];
	return block_sum(v, sh);
}
__global__ void __launch_bounds__(RT_PRESENT_THREADS)
    rt_luma_sum_kernel(const float* __restrict__ rgb, size_t n_pixels, double* __restrict__ partial) {
	__shared__ double sh[RT_PRESENT_THREADS / 32];
	double v = 0.0;
	for (size_t i = (size_t)blockIdx.x * RT_PRESENT_THREADS + threadIdx.x; i < n_pixels; i += (size_t)gridDim.x * RT_PRESENT_THREADS)
		v += luma(rgb + i * 3);
	const double s = block_sum(v, sh);
	if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// ExposureBuffer.get_variance / get_absolute_dev (:111-142): sums of (y - mean)^2 and |y - mean|
__global__ void __launch_bounds__(RT_PRESENT_THREADS)
    rt_luma_dev_kernel(const float* __restrict__ rgb, size_t n_pixels, const double* __restrict__ partial_sum_in,
                       double* __restrict__ partial_var, double* __restrict__ partial_abs) {
	__shared__ double sh[RT_PRESENT_THREADS / 32];
	const double mean = xdiv(partial_sum(partial_sum_in, gridDim.x, sh), (double)n_pixels);
	double v = 0.0, a = 0.0;
	for (size_t i = (size_t)blockIdx.x * RT_PRESENT_THREADS + threadIdx.x; i < n_pixels; i += (size_t)gridDim.x * RT_PRESENT_THREADS) {
		const double delta = xsub(luma(rgb + i * 3), mean);
		v += xmul(delta, delta);
		a += fabs(delta);
	}
	const double sv = block_sum(v, sh), sa = block_sum(a, sh);
	if (threadIdx.x == 0) { partial_var[blockIdx.x] = sv; partial_abs[blockIdx.x] = sa; }
}
// mathutils.clamp = Math.max(Math.min(x, max), min) (src/math/mathutils.ts:18-20): NaN propagates
RT_D double js_clamp01(double x) {
	if (x != x) return x;
	return x > 1.0 ? 1.0 : (x < 0.0 ? 0.0 : x);
}
// ToneMapper.get_dynamic_range (src/view/tone_mapping.ts:24-79) + ExposureBuffer.discretize_to_screen
// (src/view/exposure_buffer.ts:145-158) + CanvasScreen.convert_color (src/view/screen_canvas.ts:101-103)
__global__ void __launch_bounds__(RT_PRESENT_THREADS)
    rt_discretize_kernel(const float* __restrict__ rgb, size_t n_pixels, const double* __restrict__ partial, int n_partial,
                         rt_tone tone, uchar4* __restrict__ rgba, double* __restrict__ stats_out) {
	__shared__ double sh[RT_PRESENT_THREADS / 32];
	const double mean = xdiv(partial_sum(partial, n_partial, sh), (double)n_pixels);
	const double variance = xdiv(partial_sum(partial + n_partial, n_partial, sh), (double)n_pixels);
	const double absdev = xdiv(partial_sum(partial + 2 * n_partial, n_partial, sh), (double)n_pixels);
	double lo = 0.0, hi = 1.0;
	if (tone.kind != RT_TONE_IDENTITY) {
		const double coef = (double)(int)(1u << (tone.dynamic_range & 31u));
		const double dev = tone.kind == RT_TONE_STDDEV ? xsqrt(variance) : absdev;
		const double m = xadd(mean, dev);
		hi = (m != m || tone.max_dynamic != tone.max_dynamic) ? NAN : (m < tone.max_dynamic ? m : tone.max_dynamic);  // Math.min
		lo = xdiv(hi, coef);
		if (lo < tone.min_dynamic) {
			lo = tone.min_dynamic;
			hi = xmul(lo, coef);
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		stats_out[0] = mean; stats_out[1] = variance; stats_out[2] = absdev; stats_out[3] = lo; stats_out[4] = hi;
	}
	const double drange = xsub(hi, lo);
	for (size_t i = (size_t)blockIdx.x * RT_PRESENT_THREADS + threadIdx.x; i < n_pixels; i += (size_t)gridDim.x * RT_PRESENT_THREADS) {
		const float* px = rgb + i * 3;
		const double y = luma(px);
		const double scale = xdiv(xdiv(xsub(y, lo), drange), xadd(y, RT_JS_EPSILON));
		unsigned char out[2];
#pragma unroll
		for (int k = 0; k < 2; k++) {  // `pixels.slice(i, i+2)`: two channels
			const float compressed = (float)js_clamp01(xmul((double)px[k], scale));  // Float32Array.prototype.map
			const double v = xmul(js_clamp01((double)compressed), 255.0);
			out[k] = v != v ? 0 : (unsigned char)(int)v;  // (x * 255) << 0
		}
		rgba[i] = make_uchar4(out[0], out[1], 0, 0xff);  // convert_color()[2] is undefined -> 0 in the Uint8ClampedArray
	}
}

// ================================================================== host side
namespace {

thread_local std::string g_create_error;

template <class T>
struct DevBuf {
	T* p = nullptr;
	size_t n = 0;
	bool owned = true;  // false: a slice of a pool (adopt), freed with the pool
	cudaError_t alloc(size_t count) {
		if (count <= n && p) return cudaSuccess;
		release();
		cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
		if (e == cudaSuccess) n = count;
		else p = nullptr;
		return e;
	}
	void adopt(void* slice, size_t count) {
		release();
		p = static_cast<T*>(slice);
		n = count;
		owned = false;
	}
	void release() {
		if (p && owned) cudaFree(p);
		p = nullptr;
		n = 0;
		owned = true;
	}
};

}  // namespace

// Identity of a device-resident render call: equal keys enqueue byte-identical work.
struct RenderKey {
	rt_camera cam;
	rt_params prm;
	uint32_t flags;
	int rank, world, compact;
	void* rgb;
	void* ids;
	uint64_t scene_version;
	void* stream;
};

// Identity of a ray-generation table: the camera's basis and frame, not its position.
struct RaygenKey {
	double basis[9], fov_h, fov_v;
	uint32_t width, height, flags;
	void* dirs;
};

struct rt_ctx {
	int device = 0;
	RaygenKey raygen_key;
	bool raygen_key_valid = false;
	cudaStream_t own_stream = nullptr, stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	cudaStream_t copy_stream = nullptr;          // device->host band copies of rt_render
	cudaEvent_t band_done[16] = {};              // band b rendered (RT_MAX_BANDS)
	cudaEvent_t stage_free = nullptr;            // camera-table staging may be rewritten
	void* stage = nullptr;                       // pinned staging of the camera scan tables
	size_t stage_cap = 0;
	// CUDA-graph cache of the last device-resident render call (launch_render)
	RenderKey last_key;
	bool key_valid = false, last_was_graph = false, use_graphs = true;
	cudaGraphExec_t graph_exec = nullptr;
	uint64_t graph_kernels = 0, scene_version = 0;
	int n_bands = 4;                             // rt_render: bands of tile rows (tuning knob RT_B200_BANDS)
	bool n_bands_forced = false;
	std::string err;
	uint64_t launches = 0;
	bool has_scene = false;
	// per-stage device timing of the last eager render (rt_set_profiling / rt_stage_times)
	bool profile = false;
	cudaEvent_t stage_ev[RT_N_STAGES + 1] = {};
	bool stage_ran[RT_N_STAGES] = {};

	// packed host copy (also serves the once-per-frame start state); the members of a multi-GPU group share one
	std::shared_ptr<RtHostScene> host_p = std::make_shared<RtHostScene>();
	RtSceneCopy scene_copy;  // the uploaded description, kept for rt_scene_update
	DevBuf<unsigned char> scene_pool;            // ONE allocation behind all the scene arrays (17 cudaMallocs of a 1 M-entity scene: 77 ms)
	RtHostScene& host_ref() { return *host_p; }
	// ---- multi-GPU group (rt_create_multi): ONE process drives every GPU.  The ctx handed to the caller is the
	// leader (group[0] == this); members[1..] own a device, a stream and a worker thread that enqueues their share
	// of a frame (the per-call CPU cost of a launch sequence times 8 GPUs from one thread would be the frame time).
	std::vector<rt_ctx*> group;
	rt_ctx* leader = nullptr;
	int rank = 0;
	cudaEvent_t fork_ev = nullptr, done_ev = nullptr;
	std::thread worker;
	std::mutex wm;
	std::condition_variable wcv;
	std::atomic<uint64_t> w_posted{0}, w_finished{0};
	std::function<rt_status()> w_job;
	rt_status w_status = RT_OK;
	bool w_quit = false;
	// pipelined frames (rt_render_begin / rt_render_end): two device frames, two pinned counter snapshots
	DevBuf<float> pipe_rgb[2];
	DevBuf<int> pipe_ids[2];
	unsigned long long* pipe_counters = nullptr;  // pinned, [2][9]
	cudaEvent_t pipe_kernels[2] = {}, pipe_done[2] = {};
	uint64_t pipe_begun = 0, pipe_ended = 0;
	struct HostMap { void* host; size_t bytes; bool registered; std::vector<void*> dev; };
	std::vector<HostMap> hostmaps;  // caller-owned host frames mapped into every member's address space
	DevBuf<RtF4> node_geom;
	DevBuf<RtD4> node_geom64;
	DevBuf<RtI4> node_link;
	DevBuf<int> node_child;
	DevBuf<RtPNode> node_pk;
	DevBuf<RtWNode> node_walk;
	DevBuf<int> node_bvh;
	DevBuf<RtBvhNode> bvh_nodes;
	DevBuf<int> bvh_slots;
	DevBuf<RtF4> bvh_geom;
	DevBuf<RtF4> slot_geom;
	DevBuf<RtD4> slot_geom64;
	DevBuf<RtI4> slot_attr;
	DevBuf<int> slot_node;
	DevBuf<RtMaterial> materials;
	DevBuf<RtTexture> textures;
	DevBuf<double> substances;
	DevBuf<uint8_t> texels;
	RtDevScene dev{};

	// per-frame
	DevBuf<RtD4> row_fr;
	DevBuf<double> ray_ck;                       // ray-generation checkpoints of the current camera basis (RtFrame.ray_ck)
	DevBuf<float> rgb;
	DevBuf<int> ids;
	DevBuf<unsigned long long> counters;
	std::vector<RtD4> h_row_fr;
	DevBuf<uint8_t> l2_scratch;
	DevBuf<RtF4> prim_geom;
	DevBuf<RtQueueItem> queue;
	DevBuf<RtQueueItem> vqueue;
	DevBuf<int> queue_dense;
	DevBuf<double> present_partial;              // 3 x blocks partial sums + 8 stats (rt_present_device)
	DevBuf<uint8_t> rgba;                        // RGBA8 image of rt_present / rt_render_present
	int present_blocks = 0;
	uint32_t exposure_w = 0, exposure_h = 0;     // size of the resident ExposureBuffer in ctx->rgb (rt_render_present)
	DevBuf<double> samples;
	DevBuf<uint32_t*> peer_flags;
	std::vector<uint32_t*> peer_flags_host;
	int render_grid[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // persistent grid sizes: rt_render_kernel<false/true>, -, bounce, primary<1,2,4,8>
	int ppl = RT_PPL;                            // sub-patches (rays per lane) of a packet: 4
	int primary_minb = RT_A_MINB;                // tuning knob RT_B200_PRIMARY_MINB=4|5|6: resident CTAs per SM the primary stage is compiled for
	int bounce_min_walking = 16;                 // tuning knob RT_B200_BOUNCE_MIN (rt_bounce_kernel)
	size_t sample_bytes = (size_t)2048 << 20;    // bound on the resample stage's sample table (RT_B200_SAMPLE_MIB)
	int resample_min_frames = 4;                 // tuning knob RT_B200_RESAMPLE_MIN
	bool ordered_queue = true;                   // tuning knob RT_B200_ORDERED_QUEUE=0: packets append to the continuation queue as they finish
	bool resample = true;                        // tuning knob RT_B200_RESAMPLE=0: the bounce stage traces all frames of a rough pixel
	int bounce_minb = 8;                         // tuning knob RT_B200_BOUNCE_MINB (rt_bounce_kernel<MINB>)
	int bounce_node_batch = 8;                   // tuning knob RT_B200_NODE_BATCH (walk_iter)
	int bounce_sparse = 8;                       // tuning knob RT_B200_SPARSE: below this many walking lanes a warp stops batching node / leaf steps
};

namespace {

rt_status fail(rt_ctx* ctx, rt_status st, const std::string& msg) {
	if (ctx) ctx->err = msg;
	else g_create_error = msg;
	return st;
}

#define RT_CUDA(ctx, call)                                                                               \
	do {                                                                                                 \
		cudaError_t e_ = (call);                                                                         \
		if (e_ != cudaSuccess)                                                                           \
			return fail(ctx, RT_ERR_CUDA, rt_format("%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); \
	} while (0)

template <class T, class A>
rt_status upload(rt_ctx* ctx, DevBuf<T>& buf, const std::vector<T, A>& host) {
	RT_CUDA(ctx, buf.alloc(host.size()));
	if (!host.empty())
		RT_CUDA(ctx, cudaMemcpyAsync(buf.p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
	return RT_OK;
}

rt_status check_args(rt_ctx* ctx, const rt_camera* cam, const rt_params* prm) {
	if (!ctx) return RT_ERR_INVALID;
	std::string err;
	const rt_status st = rt_check_render_args(ctx->has_scene, (uint32_t)ctx->host_ref().textures.size(),
	                                          (uint32_t)ctx->host_ref().substances.size(), cam, prm, err);
	return st ? fail(ctx, st, err) : RT_OK;
}

const void* primary_kernel_of(int ppl, int minb) {
	(void)ppl;  // (8 rays per lane was measured no faster than 4 in round 2 and no longer fits the shared memory of a CTA)
	return minb == 4 ? (const void*)rt_primary_kernel<4, 4> : minb == 6 ? (const void*)rt_primary_kernel<4, 6> : (const void*)rt_primary_kernel<4, 5>;
}

// Builds the RtFrame (camera tables, start state) and enqueues the kernels on the ctx stream.  The frame
// can be cut into `n_bands` groups of whole tile rows, each rendered by its own launches; after_band(b,
// row_begin, row_end) is called right after band b's kernels were enqueued (rt_render uses it to start the
// device->host copy of that band behind an event while the next band renders).
//
// Device-resident entry points (no band hook) repeat byte-identical work whenever the camera stands still
// (exposure accumulation of identical frames, multi-GPU frames): the second identical call is captured
// into a CUDA graph and later ones replay it with one launch, which takes the per-launch CPU cost (two
// table copies, a memset, four kernels) out of the frame time.
typedef std::function<rt_status(int, int, int)> BandHook;
#define RT_MAX_BANDS 16
rt_status launch_render(rt_ctx* ctx, const rt_camera* cam, const rt_params* prm, uint32_t flags, float* rgb_dev,
                        int* ids_dev, int tile_rank, int tile_world, bool tile_compact = false, int n_bands = 1,
                        const BandHook& after_band = BandHook()) {
	// ---- graph cache
	RenderKey key;
	memset(&key, 0, sizeof key);
	key.cam = *cam; key.prm = *prm; key.flags = flags; key.rank = tile_rank; key.world = tile_world;
	key.compact = tile_compact ? 1 : 0; key.rgb = rgb_dev; key.ids = ids_dev; key.scene_version = ctx->scene_version;
	key.stream = ctx->stream;
	const bool cacheable = ctx->use_graphs && !after_band && prm->frame_first == 0;
	const bool same = cacheable && ctx->key_valid && memcmp(&key, &ctx->last_key, sizeof key) == 0;
	if (same && ctx->graph_exec) {
		RT_CUDA(ctx, cudaGraphLaunch(ctx->graph_exec, ctx->stream));
		ctx->launches += ctx->graph_kernels;
		ctx->last_was_graph = true;
		return RT_OK;
	}
	if (ctx->graph_exec) {
		cudaGraphExecDestroy(ctx->graph_exec);
		ctx->graph_exec = nullptr;
	}
	ctx->last_key = key;
	ctx->key_valid = cacheable;
	const bool capture = same;  // second identical call in a row
	if (ctx->last_was_graph) {
		// a replayed graph may still be reading the table staging
		RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		ctx->last_was_graph = false;
	}

	// ---- host preparation and every allocation (nothing below this block allocates: it may be captured)
	// Ray generation is per camera BASIS (and frame, and shard): the directions do not depend on where the camera
	// stands, so a camera that merely moved (the reference's WASD keys, src/main.ts:296-330) keeps its table.
	const size_t n_row = cam->height;
	RT_CUDA(ctx, ctx->row_fr.alloc(n_row));
	const int ray_ckh = raygen_checkpoints_per_half((int)cam->width);
	RT_CUDA(ctx, ctx->ray_ck.alloc((size_t)cam->height * 2 * ray_ckh * 6));
	RaygenKey rk;
	memset(&rk, 0, sizeof rk);
	memcpy(rk.basis, cam->fr, sizeof cam->fr); memcpy(rk.basis + 3, cam->lf, sizeof cam->lf); memcpy(rk.basis + 6, cam->up, sizeof cam->up);
	rk.fov_h = cam->fov_h; rk.fov_v = cam->fov_v; rk.width = cam->width; rk.height = cam->height; rk.flags = cam->flags;
	rk.dirs = ctx->ray_ck.p;
	const bool need_raygen = !capture && (!ctx->raygen_key_valid || memcmp(&rk, &ctx->raygen_key, sizeof rk) != 0);
	double scan_cos = std::cos(cam->fov_h / cam->width), scan_sin = std::sin(cam->fov_h / cam->width);  // (as rt_build_camera_rows)
	RtD4* st_row = nullptr;
	if (need_raygen) {
		if (ctx->stage_cap < n_row * sizeof(RtD4)) {
			if (ctx->stage) cudaFreeHost(ctx->stage);
			ctx->stage = nullptr;
			ctx->stage_cap = 0;
			RT_CUDA(ctx, cudaMallocHost(&ctx->stage, n_row * sizeof(RtD4)));
			ctx->stage_cap = n_row * sizeof(RtD4);
		}
		RT_CUDA(ctx, cudaEventSynchronize(ctx->stage_free));  // the previous table copy has left the staging
		rt_build_camera_rows(*cam, ctx->h_row_fr, scan_cos, scan_sin);
		st_row = reinterpret_cast<RtD4*>(ctx->stage);
		memcpy(st_row, ctx->h_row_fr.data(), n_row * sizeof(RtD4));
	}
	RtFrame F{};
	std::string err;
	if (rt_status st = rt_fill_frame(ctx->host_ref(), cam, prm, F, err)) return fail(ctx, st, err);
	F.row_fr = ctx->row_fr.p;
	F.ray_ck = ctx->ray_ck.p;
	F.ray_ckh = ray_ckh;
	F.scan_cos = scan_cos;
	F.scan_sin = scan_sin;
	F.rgb = rgb_dev;
	F.first_ids = ids_dev;
	F.tile_rank = tile_rank;
	F.tile_world = tile_world;
	F.tile_compact = tile_compact ? 1 : 0;
	F.bounce_min_walking = ctx->bounce_min_walking;
	F.bounce_node_batch = ctx->bounce_node_batch;
	F.bounce_sparse = ctx->bounce_sparse;
	// u64 cells: [0..7] work counters, [8] error flags, then per band {patch dispenser, queue count, queue cursor,
	// resample queue count, resample queue cursor}
	n_bands = std::max(1, std::min(n_bands, RT_MAX_BANDS));
	const size_t n_cells = 9 + 5 * RT_MAX_BANDS;
	RT_CUDA(ctx, ctx->counters.alloc(n_cells));
	F.counters = ctx->counters.p;
	F.error_flags = reinterpret_cast<uint32_t*>(ctx->counters.p + 8);
	const bool count = (flags & RT_RENDER_COUNTERS) != 0;
	const int tiles_x = (F.width + RT_TILE_W - 1) / RT_TILE_W, tiles_y = (F.height + RT_TILE_H - 1) / RT_TILE_H;
	const int n_tiles = tiles_x * tiles_y;
	if ((n_tiles - tile_rank + tile_world - 1) / tile_world <= 0) return RT_OK;
	if (tile_compact || tile_world > 1) n_bands = 1;
	n_bands = std::min(n_bands, tiles_y);
	// primary-ray preparation: origin-relative slot records + the origin chain (start node ... root)
	const RtHostScene& H = ctx->host_ref();
	const int n_slots = (int)H.slot_geom.size();
	const bool prim = n_slots > 0 && !(prm->flags & RT_PARAM_NO_PRIMARY_RECORDS);
	F.prim_geom = nullptr;
	F.chain_levels = 0;
	F.packet_ok = 0;
	if (prim) {
		RT_CUDA(ctx, ctx->prim_geom.alloc((size_t)n_slots));
		F.prim_geom = ctx->prim_geom.p;
		rt_fill_chain(H, F);
	}
	const bool pipeline = F.packet_ok && !count && !(prm->flags & RT_PARAM_PER_RAY) && !F.search64;
	if (!pipeline) F.packet_ok = 0;
	// rough pixels: from resample_min_frames exposure frames on (4), their frames are traced as independent samples by
	// the resample stage (below that the bounce stage's lane traces them in a row: the stage's launches and its own
	// tail cost more than the short in-lane tail they remove)
	const bool resample = pipeline && prm->n_frames >= (uint32_t)ctx->resample_min_frames && ctx->resample;
	if (pipeline) {
		const size_t cap = tile_compact ? (size_t)((n_tiles - tile_rank + tile_world - 1) / tile_world) * RT_BLOCK
		                                : (size_t)F.width * F.height;
		RT_CUDA(ctx, ctx->queue.alloc(cap));
		if (resample) RT_CUDA(ctx, ctx->vqueue.alloc(cap));
		if (prm->refmax > 1 && ctx->ordered_queue) RT_CUDA(ctx, ctx->queue_dense.alloc(cap));
	}
	auto grid_of = [&](int which, const void* kernel, int threads, int& out) -> rt_status {
		int& grid = ctx->render_grid[which];
		if (grid == 0) {
			int per_sm = 0, sms = 0;
			RT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0));
			RT_CUDA(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
			grid = std::max(1, per_sm) * std::max(1, sms);
		}
		out = grid;
		return RT_OK;
	};
	// rays per lane of the packet stage (tuning knob: RT_B200_PPL=1|2|4|8 in the environment)
	const int ppl = ctx->ppl, pminb = ctx->primary_minb;
	const void* primary_kernel = primary_kernel_of(ppl, pminb);
	// resident CTAs per SM the bounce stage is compiled for (tuning knob: RT_B200_BOUNCE_MINB=4|5|6|8)
	const int minb = ctx->bounce_minb;
	// (frames with exact ties, F.tie_checks, take the builds that carry the tie branches: one, for 8 resident CTAs)
	const bool ties = F.tie_checks != 0;
	const void* bounce_kernel = ties ? (const void*)rt_bounce_kernel<8, true>
	                          : minb == 4 ? (const void*)rt_bounce_kernel<4, false> : minb == 5 ? (const void*)rt_bounce_kernel<5, false>
	                          : minb == 6 ? (const void*)rt_bounce_kernel<6, false> : (const void*)rt_bounce_kernel<8, false>;
	const void* resample_kernel = ties ? (const void*)rt_resample_kernel<8, true>
	                            : minb == 4 ? (const void*)rt_resample_kernel<4, false> : minb == 5 ? (const void*)rt_resample_kernel<5, false>
	                            : minb == 6 ? (const void*)rt_resample_kernel<6, false> : (const void*)rt_resample_kernel<8, false>;
	int grid_primary = 0, grid_bounce = 0, grid_ray = 0, grid_resample = 0;
	unsigned resample_chunk = 1;
	F.samples = nullptr;
	F.sample_chunk = 0;
	if (pipeline) {
		if (rt_status st = grid_of(4, primary_kernel, RT_A_WARPS * 32, grid_primary)) return st;
		if (rt_status st = grid_of(ties ? 5 : 3, bounce_kernel, RT_WARPS_PER_CTA * 32, grid_bounce)) return st;
		if (resample) {
			if (rt_status st = grid_of(ties ? 6 : 2, resample_kernel, RT_WARPS_PER_CTA * 32, grid_resample)) return st;
			// the sample table: [pixels of a round][n_frames][3] float64 path colours; a queue longer than the table
			// takes several rounds (RT_B200_SAMPLE_MIB bounds the table, default 2 GiB)
			const size_t cap = tile_compact ? (size_t)((n_tiles - tile_rank + tile_world - 1) / tile_world) * RT_BLOCK
			                                : (size_t)F.width * F.height;
			const size_t per_pixel = (size_t)prm->n_frames * 3 * sizeof(double);
			resample_chunk = (unsigned)std::max<size_t>(1, std::min<size_t>(cap, ctx->sample_bytes / per_pixel));
			RT_CUDA(ctx, ctx->samples.alloc((size_t)resample_chunk * prm->n_frames * 3));
			F.samples = ctx->samples.p;
			F.sample_chunk = resample_chunk;
		}
	} else if (count) {
		if (rt_status st = grid_of(1, (const void*)rt_render_kernel<true>, RT_WARPS_PER_CTA * 32, grid_ray)) return st;
	} else {
		if (rt_status st = grid_of(0, (const void*)rt_render_kernel<false>, RT_WARPS_PER_CTA * 32, grid_ray)) return st;
	}

	// ---- the stream work
	const uint64_t launches_before = ctx->launches;
	auto enqueue = [&]() -> rt_status {
		// A captured frame repeats the call before it (same camera, same scene): the scan tables and the
		// origin-relative records that call left on the device are this frame's, so the graph holds neither.
		if (need_raygen) {
			RT_CUDA(ctx, cudaMemcpyAsync(ctx->row_fr.p, st_row, n_row * sizeof(RtD4), cudaMemcpyHostToDevice, ctx->stream));
			RT_CUDA(ctx, cudaEventRecord(ctx->stage_free, ctx->stream));
		}
		const bool prof = ctx->profile && !capture && n_bands == 1;
		for (int k = 0; k < RT_N_STAGES; k++) ctx->stage_ran[k] = false;
		auto mark = [&](int stage) -> cudaError_t {  // event before stage `stage` (= after the one before it)
			return prof ? cudaEventRecord(ctx->stage_ev[stage], ctx->stream) : cudaSuccess;
		};
		RT_CUDA(ctx, mark(0));
		// Frame setup, one launch: the control cells are zeroed; per camera POSE (a captured frame repeats the call
		// before it and inherits it) the origin-relative records; per camera BASIS the ray generation - the directions
		// do not depend on where the camera stands, so a camera that merely moved (the reference's WASD keys,
		// src/main.ts:296-330) keeps its table.
		{
			int raygen_blocks = 0, prep_blocks = 0;
			if (!capture && prim) prep_blocks = (n_slots + RT_SETUP_THREADS - 1) / RT_SETUP_THREADS;
			if (need_raygen) {
				raygen_blocks = (6 * F.height + RT_SETUP_THREADS - 1) / RT_SETUP_THREADS;
				ctx->raygen_key = rk;
				ctx->raygen_key_valid = true;
			}
			rt_frame_setup_kernel<<<1 + raygen_blocks + prep_blocks, RT_SETUP_THREADS, 0, ctx->stream>>>(
			    ctx->dev, F, ctx->counters.p, (int)n_cells, raygen_blocks, ctx->ray_ck.p, ctx->prim_geom.p);
			ctx->launches++;
			ctx->stage_ran[0] = prof;
			RT_CUDA(ctx, cudaGetLastError());
		}
		RT_CUDA(ctx, mark(1));
		for (int band = 0; band < n_bands; band++) {
			const int row_begin = (int)((long long)tiles_y * band / n_bands), row_end = (int)((long long)tiles_y * (band + 1) / n_bands);
			F.tile_begin = row_begin * tiles_x;
			F.tile_end = row_end * tiles_x;
			const int band_tiles = F.tile_end - F.tile_begin;
			const int my_tiles = (band_tiles - tile_rank + tile_world - 1) / tile_world;
			if (my_tiles <= 0) continue;
			const int y_begin = row_begin * RT_TILE_H, y_end = std::min(F.height, row_end * RT_TILE_H);
			F.out_first = tile_compact ? 0ull : (unsigned long long)y_begin * F.width;
			const size_t n_out = (tile_compact || tile_world > 1) ? (size_t)my_tiles * RT_BLOCK : (size_t)(y_end - y_begin) * F.width;
			unsigned long long* cells = ctx->counters.p + 9 + 5 * band;
			F.work_counter = reinterpret_cast<unsigned*>(cells);
			F.queue_count = reinterpret_cast<unsigned*>(cells + 1);
			F.queue_taken = reinterpret_cast<unsigned*>(cells + 2);
			F.vqueue_count = reinterpret_cast<unsigned*>(cells + 3);
			F.vqueue_taken = reinterpret_cast<unsigned*>(cells + 4);
			const int n_patches = my_tiles * 8;
			if (pipeline) {
				// primary stage (packet walk) -> shade stage -> bounce stage over the continuation queue
				F.queue = ctx->queue.p + F.out_first;
				F.queue_dense = prm->refmax > 1 && ctx->ordered_queue ? ctx->queue_dense.p : nullptr;
				F.vqueue = resample ? ctx->vqueue.p + F.out_first : nullptr;
				const int n_packets = my_tiles * (8 / ppl);
				const int blocks = std::min(grid_primary, (n_packets + RT_A_WARPS - 1) / RT_A_WARPS);
				void* args[] = {(void*)&ctx->dev, (void*)&F, (void*)&tiles_x, (void*)&n_packets};
				RT_CUDA(ctx, cudaLaunchKernel(primary_kernel, dim3(blocks), dim3(RT_A_WARPS * 32), args, 0, ctx->stream));
				ctx->launches++;
				RT_CUDA(ctx, mark(2));
				if (F.queue_dense) {
					rt_queue_compact_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, ctx->stream>>>(F, tiles_x, n_out);
					ctx->launches++;
					ctx->stage_ran[2] = prof;
					RT_CUDA(ctx, cudaGetLastError());
				}
				RT_CUDA(ctx, mark(3));
				void* bargs[] = {(void*)&ctx->dev, (void*)&F, (void*)&tiles_x};
				// with refmax <= 1 no path continues after its first hit: the queue only ever holds rays the lock-step
				// walk could not take, and one CTA per SM drains it
				const int bounce_blocks = prm->refmax <= 1 ? std::max(1, grid_bounce / std::max(1, minb)) : grid_bounce;
				RT_CUDA(ctx, cudaLaunchKernel(bounce_kernel, dim3(std::min(bounce_blocks, (n_patches + RT_WARPS_PER_CTA - 1) / RT_WARPS_PER_CTA)),
				                              dim3(RT_WARPS_PER_CTA * 32), bargs, 0, ctx->stream));
				RT_CUDA(ctx, mark(4));
				ctx->stage_ran[1] = ctx->stage_ran[3] = prof;
				if (resample) {
					const int blocks = std::min(grid_resample, (n_patches + RT_WARPS_PER_CTA - 1) / RT_WARPS_PER_CTA);
					for (size_t first = 0; first < n_out; first += resample_chunk) {
						unsigned first_u = (unsigned)first, chunk_u = resample_chunk;
						void* rargs[] = {(void*)&ctx->dev, (void*)&F, (void*)&tiles_x, (void*)&first_u, (void*)&chunk_u};
						RT_CUDA(ctx, cudaLaunchKernel(resample_kernel, dim3(blocks), dim3(RT_WARPS_PER_CTA * 32), rargs, 0, ctx->stream));
						rt_resample_blend_kernel<<<std::min(blocks, (int)((std::min<size_t>(n_out - first, resample_chunk) + 255) / 256)), 256, 0, ctx->stream>>>(
						    F, tiles_x, first_u, chunk_u);
						ctx->launches += 2;
						RT_CUDA(ctx, cudaGetLastError());
					}
					ctx->stage_ran[4] = prof;
				}
				RT_CUDA(ctx, mark(5));
			} else if (count) {
				rt_render_kernel<true><<<std::min(grid_ray, (n_patches + RT_WARPS_PER_CTA - 1) / RT_WARPS_PER_CTA), RT_WARPS_PER_CTA * 32, 0,
				                         ctx->stream>>>(ctx->dev, F, tiles_x, n_patches);
			} else {
				rt_render_kernel<false><<<std::min(grid_ray, (n_patches + RT_WARPS_PER_CTA - 1) / RT_WARPS_PER_CTA), RT_WARPS_PER_CTA * 32, 0,
				                          ctx->stream>>>(ctx->dev, F, tiles_x, n_patches);
			}
			ctx->launches++;
			RT_CUDA(ctx, cudaGetLastError());
			if (after_band)
				if (rt_status st = after_band(band, y_begin, y_end)) return st;
		}
		return RT_OK;
	};
	if (!capture) return enqueue();
	RT_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
	const rt_status st = enqueue();
	cudaGraph_t graph = nullptr;
	const cudaError_t e_end = cudaStreamEndCapture(ctx->stream, &graph);
	ctx->graph_kernels = ctx->launches - launches_before;
	if (st != RT_OK || e_end != cudaSuccess || !graph) {
		if (graph) cudaGraphDestroy(graph);
		cudaGetLastError();
		ctx->use_graphs = false;  // capture is not possible in this environment: stay eager from now on
		ctx->key_valid = false;
		ctx->launches = launches_before;
		if (st != RT_OK) return st;
		return enqueue();
	}
	const cudaError_t e_inst = cudaGraphInstantiate(&ctx->graph_exec, graph, 0);
	cudaGraphDestroy(graph);
	if (e_inst != cudaSuccess) {
		ctx->graph_exec = nullptr;
		ctx->use_graphs = false;
		ctx->key_valid = false;
		ctx->launches = launches_before;
		cudaGetLastError();
		return enqueue();
	}
	RT_CUDA(ctx, cudaGraphLaunch(ctx->graph_exec, ctx->stream));
	ctx->last_was_graph = true;
	return RT_OK;
}

// Bands overlap the host copy of a finished part of the frame with the rendering of the next - worth it while a frame
// renders about as fast as it travels (refmax <= 1: 0.2 ms against 0.45 ms of copy).  With bouncing paths every band
// would run its own bounce / resample stage over a quarter of the queue, and those stages end in a tail of long paths
// whose length does not shrink with the queue: measured on configs[2], 4 bands cost 4.99 ms against 3.37 ms for one.
int host_bands(const rt_ctx* ctx, const rt_params* prm) { return ctx->n_bands_forced || prm->refmax <= 1 ? ctx->n_bands : 1; }

rt_status read_counters(rt_ctx* ctx, rt_counters* out) {
	unsigned long long h[9] = {0};
	if (ctx->counters.p)
		RT_CUDA(ctx, cudaMemcpyAsync(h, ctx->counters.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
	RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	const uint32_t ef = (uint32_t)h[8];
	out->paths = h[0]; out->segments = h[1]; out->nodes = h[2]; out->tests = h[3]; out->shades = h[4];
	out->confirms = h[5];
	out->texture_errors = (ef & RT_ERRFLAG_TEXTURE) ? 1 : 0;
	out->acute_warnings = (ef & RT_ERRFLAG_ACUTE) ? 1 : 0;
	if (ef & RT_ERRFLAG_STACK)  // never a silently wrong pixel: the frame is refused
		return fail(ctx, RT_ERR_UNSUPPORTED, "a traversal stack was too small for this scene (internal limit): the frame was not rendered correctly");
	return RT_OK;
}

}  // namespace

// ================================================================== multi-GPU group (one process, every GPU)
namespace {

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
	__builtin_ia32_pause();
#else
	std::this_thread::yield();
#endif
}

// A member's worker: sleeps on its condition variable, spins for a moment before that (a frame is a fraction of a
// millisecond: the next job usually arrives before a sleeping thread could be woken).
void worker_loop(rt_ctx* m) {
	cudaSetDevice(m->device);
	uint64_t seen = 0;
	for (;;) {
		for (int i = 0; i < 40000 && m->w_posted.load(std::memory_order_acquire) == seen; i++) cpu_relax();
		if (m->w_posted.load(std::memory_order_acquire) == seen) {
			std::unique_lock<std::mutex> lk(m->wm);
			m->wcv.wait(lk, [&] { return m->w_posted.load(std::memory_order_acquire) != seen; });
		}
		std::function<rt_status()> job;
		{
			std::lock_guard<std::mutex> lk(m->wm);
			seen = m->w_posted.load(std::memory_order_acquire);
			if (m->w_quit) return;
			job = m->w_job;
		}
		m->w_status = job ? job() : RT_OK;
		m->w_finished.store(seen, std::memory_order_release);
	}
}

void worker_post(rt_ctx* m, std::function<rt_status()> job) {
	{
		std::lock_guard<std::mutex> lk(m->wm);
		m->w_job = std::move(job);
		m->w_posted.fetch_add(1, std::memory_order_release);
	}
	m->wcv.notify_one();
}

rt_status worker_wait(rt_ctx* m) {
	const uint64_t want = m->w_posted.load(std::memory_order_acquire);
	for (int i = 0; m->w_finished.load(std::memory_order_acquire) != want; i++)
		if (i < 100000) cpu_relax();
		else std::this_thread::yield();
	return m->w_status;
}

// fn(rank, member) for every member of the leader's group: members 1.. on their workers, member 0 on the calling
// thread.  Returns the first failure, with the member's message copied to the leader.
rt_status group_run(rt_ctx* L, const std::function<rt_status(int, rt_ctx*)>& fn) {
	const int n = (int)L->group.size();
	for (int r = 1; r < n; r++) {
		rt_ctx* m = L->group[r];
		worker_post(m, [&fn, r, m]() { return fn(r, m); });
	}
	rt_status st = fn(0, L);
	for (int r = 1; r < n; r++) {
		const rt_status sr = worker_wait(L->group[r]);
		if (sr != RT_OK && st == RT_OK) {
			st = sr;
			L->err = rt_format("GPU %d of the group: %s", L->group[r]->device, L->group[r]->err.c_str());
		}
	}
	cudaSetDevice(L->device);
	return st;
}

// A caller-owned host buffer, page-locked and mapped into every member's address space (zero copy): each GPU's
// kernels store its tiles straight into it over that GPU's own PCIe link.
rt_status group_map_host(rt_ctx* L, void* ptr, size_t bytes, rt_ctx::HostMap** out) {
	for (auto& hm : L->hostmaps)
		if (hm.host == ptr && hm.bytes == bytes) { *out = &hm; return RT_OK; }
	for (size_t i = 0; i < L->hostmaps.size();)  // the same address with another size: the old buffer is gone
		if (L->hostmaps[i].host == ptr) {
			if (L->hostmaps[i].registered) cudaHostUnregister(ptr);
			L->hostmaps.erase(L->hostmaps.begin() + (long)i);
		} else i++;
	if (L->hostmaps.size() >= 8) {
		if (L->hostmaps[0].registered) cudaHostUnregister(L->hostmaps[0].host);
		L->hostmaps.erase(L->hostmaps.begin());
	}
	rt_ctx::HostMap hm{ptr, bytes, true, {}};
	RT_CUDA(L, cudaSetDevice(L->device));
	cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable);
	if (e == cudaErrorHostMemoryAlreadyRegistered) {  // pinned by the caller (rt_host_register): with UVA it is mapped already
		cudaGetLastError();
		hm.registered = false;
	} else if (e != cudaSuccess) {
		return fail(L, RT_ERR_CUDA, rt_format("cudaHostRegister(%zu bytes): %s", bytes, cudaGetErrorString(e)));
	}
	for (rt_ctx* m : L->group) {
		void* d = nullptr;
		cudaSetDevice(m->device);
		e = cudaHostGetDevicePointer(&d, ptr, 0);
		if (e != cudaSuccess) {
			if (hm.registered) cudaHostUnregister(ptr);
			cudaSetDevice(L->device);
			return fail(L, RT_ERR_CUDA, rt_format("cudaHostGetDevicePointer on GPU %d: %s", m->device, cudaGetErrorString(e)));
		}
		hm.dev.push_back(d);
	}
	cudaSetDevice(L->device);
	L->hostmaps.push_back(hm);
	*out = &L->hostmaps.back();
	return RT_OK;
}

// trace_frame() of a group into HOST buffers: every member renders its interleaved 16x16 tiles (tile t -> member
// t % n) from its replica of the scene and stores them into the mapped host frame.  Synchronous.
rt_status group_render_host(rt_ctx* L, const rt_camera* cam, const rt_params* prm, uint32_t flags, float* rgb, int32_t* ids,
                            rt_counters* counters) {
	const int n = (int)L->group.size();
	const size_t npx = (size_t)cam->width * cam->height;
	rt_ctx::HostMap *mr = nullptr, *mi = nullptr;
	if (rt_status st = group_map_host(L, rgb, npx * 3 * sizeof(float), &mr)) return st;
	const std::vector<void*> rgb_dev = mr->dev;  // (copied: a second mapping may move the table)
	std::vector<void*> ids_dev(n, nullptr);
	if (ids) {
		if (rt_status st = group_map_host(L, ids, npx * sizeof(int32_t), &mi)) return st;
		ids_dev = mi->dev;
	}
	if (counters) flags |= RT_RENDER_COUNTERS;
	std::vector<rt_counters> part(n);
	const rt_status st = group_run(L, [&](int r, rt_ctx* m) -> rt_status {
		RT_CUDA(m, cudaSetDevice(m->device));
		if (rt_status s = launch_render(m, cam, prm, flags, (float*)rgb_dev[r], (int*)ids_dev[r], r, n, false)) return s;
		return read_counters(m, &part[r]);  // synchronises the member's stream and fetches its error flags
	});
	if (st != RT_OK) return st;
	rt_counters tot;
	memset(&tot, 0, sizeof tot);
	for (const rt_counters& c : part) {
		tot.paths += c.paths; tot.segments += c.segments; tot.nodes += c.nodes; tot.tests += c.tests; tot.shades += c.shades;
		tot.confirms += c.confirms; tot.texture_errors |= c.texture_errors; tot.acute_warnings |= c.acute_warnings;
	}
	if (counters) *counters = tot;
	if (tot.texture_errors) return fail(L, RT_ERR_TEXTURE, "Texture coordinates out of bounds");
	return RT_OK;
}

// ... into a DEVICE frame that lives on the leader's GPU: the other members store their tiles into it over NVLink
// (peer access inside one process: plain pointers), the leader's stream waits for their events.  Asynchronous.
rt_status group_render_device(rt_ctx* L, const rt_camera* cam, const rt_params* prm, uint32_t flags, float* rgb_dev, int32_t* ids_dev) {
	const int n = (int)L->group.size();
	RT_CUDA(L, cudaSetDevice(L->device));
	RT_CUDA(L, cudaEventRecord(L->fork_ev, L->stream));  // whatever the leader's stream did to the frame so far comes first
	const rt_status st = group_run(L, [&](int r, rt_ctx* m) -> rt_status {
		RT_CUDA(m, cudaSetDevice(m->device));
		if (r > 0) RT_CUDA(m, cudaStreamWaitEvent(m->stream, L->fork_ev, 0));
		if (rt_status s = launch_render(m, cam, prm, flags, rgb_dev, ids_dev, r, n, false)) return s;
		if (r > 0) RT_CUDA(m, cudaEventRecord(m->done_ev, m->stream));
		return RT_OK;
	});
	if (st != RT_OK) return st;
	for (int r = 1; r < n; r++) RT_CUDA(L, cudaStreamWaitEvent(L->stream, L->group[r]->done_ev, 0));
	return RT_OK;
}

// one render into a device frame, single GPU or group
rt_status render_device_any(rt_ctx* ctx, const rt_camera* cam, const rt_params* prm, uint32_t flags, float* rgb_dev, int32_t* ids_dev) {
	if (ctx->group.size() > 1) return group_render_device(ctx, cam, prm, flags, rgb_dev, ids_dev);
	return launch_render(ctx, cam, prm, flags, rgb_dev, ids_dev, 0, 1);
}

rt_status read_counters_any(rt_ctx* ctx, rt_counters* out) {
	if (ctx->group.size() <= 1) return read_counters(ctx, out);
	memset(out, 0, sizeof *out);
	for (rt_ctx* m : ctx->group) {
		rt_counters c;
		RT_CUDA(ctx, cudaSetDevice(m->device));
		if (rt_status st = read_counters(m, &c)) return st;
		out->paths += c.paths; out->segments += c.segments; out->nodes += c.nodes; out->tests += c.tests; out->shades += c.shades;
		out->confirms += c.confirms; out->texture_errors |= c.texture_errors; out->acute_warnings |= c.acute_warnings;
	}
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	return RT_OK;
}

}  // namespace

// ================================================================== C ABI
// No C++ exception crosses the C ABI: the entry points that build large host structures (a 1 M-entity scene is a few
// hundred MB of vectors, on several threads) turn an allocation failure into a status and a message.
template <class F>
static rt_status abi_guard(rt_ctx* ctx, const char* what, F&& body) {
	try {
		return body();
	} catch (const std::bad_alloc&) {
		return fail(ctx, RT_ERR_INVALID, rt_format("%s: out of host memory", what));
	} catch (const std::exception& e) {
		return fail(ctx, RT_ERR_INVALID, rt_format("%s: %s", what, e.what()));
	} catch (...) {
		return fail(ctx, RT_ERR_INVALID, rt_format("%s: unknown failure", what));
	}
}

extern "C" {

uint32_t rt_abi_version(void) { return RT_B200_ABI_VERSION; }

const char* rt_last_error(const rt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

rt_status rt_create(int32_t device, rt_ctx** out) {
	if (!out) return fail(nullptr, RT_ERR_INVALID, "rt_create: out is NULL");
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
		return fail(nullptr, RT_ERR_CUDA,
		            rt_format("rt_create: no CUDA device (%s); this library has no CPU path",
		                      e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)));
	if (device < 0) {
		if (cudaGetDevice(&device) != cudaSuccess) device = 0;
	}
	if (device >= count) return fail(nullptr, RT_ERR_INVALID, rt_format("rt_create: device %d of %d", device, count));
	e = cudaSetDevice(device);
	if (e != cudaSuccess) return fail(nullptr, RT_ERR_CUDA, rt_format("cudaSetDevice(%d): %s", device, cudaGetErrorString(e)));
	rt_ctx* ctx = new rt_ctx();
	ctx->device = device;
	e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
	if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
	if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->stage_free, cudaEventDisableTiming);
	for (int b = 0; b < RT_MAX_BANDS && e == cudaSuccess; b++) e = cudaEventCreateWithFlags(&ctx->band_done[b], cudaEventDisableTiming);
	for (int k = 0; k <= RT_N_STAGES && e == cudaSuccess; k++) e = cudaEventCreate(&ctx->stage_ev[k]);
	if (e != cudaSuccess) {
		rt_status st = fail(nullptr, RT_ERR_CUDA, rt_format("rt_create: %s", cudaGetErrorString(e)));
		rt_destroy(ctx);
		return st;
	}
	ctx->stream = ctx->own_stream;
	if (const char* e = getenv("RT_B200_GRAPHS")) ctx->use_graphs = atoi(e) != 0;
	if (const char* e = getenv("RT_B200_BANDS")) {
		const int v = atoi(e);
		if (v >= 1 && v <= RT_MAX_BANDS) {
			ctx->n_bands = v;
			ctx->n_bands_forced = true;
		}
	}
	if (const char* e = getenv("RT_B200_BOUNCE_MIN")) {
		const int v = atoi(e);
		if (v >= 1 && v <= 32) ctx->bounce_min_walking = v;
	}
	if (const char* e = getenv("RT_B200_RESAMPLE")) ctx->resample = atoi(e) != 0;
	if (const char* e = getenv("RT_B200_ORDERED_QUEUE")) ctx->ordered_queue = atoi(e) != 0;
	if (const char* e = getenv("RT_B200_RESAMPLE_MIN")) ctx->resample_min_frames = std::max(2, atoi(e));
	if (const char* e = getenv("RT_B200_SAMPLE_MIB")) ctx->sample_bytes = (size_t)std::max(1, atoi(e)) << 20;
	if (const char* e = getenv("RT_B200_SAMPLE_KIB")) ctx->sample_bytes = (size_t)std::max(1, atoi(e)) << 10;  // (tests: several rounds on a small frame)
	if (const char* e = getenv("RT_B200_BOUNCE_MINB")) {
		const int v = atoi(e);
		if (v == 4 || v == 5 || v == 6 || v == 8) ctx->bounce_minb = v;
	}
	if (const char* e = getenv("RT_B200_NODE_BATCH")) {
		const int v = atoi(e);
		if (v >= 1 && v <= 32) ctx->bounce_node_batch = v;
	}
	if (const char* e = getenv("RT_B200_SPARSE")) {
		const int v = atoi(e);
		if (v >= 0 && v <= 33) ctx->bounce_sparse = v;
	}
	if (const char* e = getenv("RT_B200_PPL")) {
		const int v = atoi(e);
		if (v == 4) ctx->ppl = v;
	}
	if (const char* e = getenv("RT_B200_PRIMARY_MINB")) {
		const int v = atoi(e);
		if (v == 4 || v == 5 || v == 6) ctx->primary_minb = v;
	}
	*out = ctx;
	return RT_OK;
}

rt_status rt_create_multi(int32_t n_gpus, const int32_t* devices, rt_ctx** out) {
	if (!out) return fail(nullptr, RT_ERR_INVALID, "rt_create_multi: out is NULL");
	*out = nullptr;
	if (n_gpus < 1 || n_gpus > 64) return fail(nullptr, RT_ERR_INVALID, rt_format("rt_create_multi: n_gpus %d", n_gpus));
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
		return fail(nullptr, RT_ERR_CUDA, rt_format("rt_create_multi: no CUDA device (%s); this library has no CPU path",
		                                            e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)));
	if (!devices && n_gpus > count)
		return fail(nullptr, RT_ERR_INVALID, rt_format("rt_create_multi: %d GPUs asked for, %d present", n_gpus, count));
	std::vector<rt_ctx*> members;
	auto undo = [&]() {
		for (rt_ctx* m : members) rt_destroy(m);
	};
	for (int i = 0; i < n_gpus; i++) {
		rt_ctx* m = nullptr;
		const rt_status st = rt_create(devices ? devices[i] : i, &m);
		if (st != RT_OK) {
			undo();
			return st;
		}
		members.push_back(m);
	}
	rt_ctx* L = members[0];
	if (n_gpus == 1) {
		*out = L;
		return RT_OK;
	}
	// every GPU reaches every other GPU's memory (NVLink / NVSwitch): scene replication and the device-resident frame
	for (rt_ctx* a : members)
		for (rt_ctx* b : members) {
			if (a->device == b->device) continue;
			int can = 0;
			cudaDeviceCanAccessPeer(&can, a->device, b->device);
			cudaSetDevice(a->device);
			e = can ? cudaDeviceEnablePeerAccess(b->device, 0) : cudaErrorPeerAccessUnsupported;
			if (e == cudaErrorPeerAccessAlreadyEnabled) {
				cudaGetLastError();
			} else if (e != cudaSuccess) {
				const rt_status st = fail(nullptr, RT_ERR_CUDA, rt_format("rt_create_multi: GPU %d cannot map GPU %d's memory: %s", a->device,
				                                                          b->device, cudaGetErrorString(e)));
				cudaGetLastError();
				undo();
				return st;
			}
		}
	for (int r = 0; r < n_gpus; r++) {
		rt_ctx* m = members[r];
		m->rank = r;
		m->leader = r ? L : nullptr;
		cudaSetDevice(m->device);
		e = cudaEventCreateWithFlags(&m->fork_ev, cudaEventDisableTiming);
		if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->done_ev, cudaEventDisableTiming);
		if (e != cudaSuccess) {
			const rt_status st = fail(nullptr, RT_ERR_CUDA, rt_format("rt_create_multi: %s", cudaGetErrorString(e)));
			undo();
			return st;
		}
	}
	L->group = members;
	for (int r = 1; r < n_gpus; r++) members[r]->worker = std::thread(worker_loop, members[r]);
	cudaSetDevice(L->device);
	*out = L;
	return RT_OK;
}

uint32_t rt_group_size(const rt_ctx* ctx) { return ctx ? (uint32_t)std::max<size_t>(1, ctx->group.size()) : 0; }

void rt_destroy(rt_ctx* ctx) {
#ifdef RT_WALK_PROFILE
	if (ctx) {
		unsigned long long h[16];
		cudaSetDevice(ctx->device);
		cudaDeviceSynchronize();
		if (cudaMemcpyFromSymbol(h, g_walk_prof, sizeof h) == cudaSuccess) {
			fprintf(stderr, "[walk profile]");
			for (int k = 0; k < 16; k++) fprintf(stderr, " %llu", h[k]);
			fprintf(stderr, "\n");
		}
	}
#endif
#ifdef RT_WALK_TIMELINE
	if (ctx) {
		cudaSetDevice(ctx->device);
		cudaDeviceSynchronize();
		static unsigned long long t_start[RT_PROF_WARPS], t_exh[RT_PROF_WARPS], t_exit[RT_PROF_WARPS];
		if (cudaMemcpyFromSymbol(t_start, g_prof_start, sizeof t_start) == cudaSuccess && cudaMemcpyFromSymbol(t_exh, g_prof_exhaust, sizeof t_exh) == cudaSuccess &&
		    cudaMemcpyFromSymbol(t_exit, g_prof_exit, sizeof t_exit) == cudaSuccess) {
			unsigned long long t0 = ~0ull;
			for (int k = 0; k < RT_PROF_WARPS; k++)
				if (t_start[k]) t0 = std::min(t0, t_start[k]);
			unsigned hist[3][256] = {{0}};
			for (int k = 0; k < RT_PROF_WARPS; k++) {
				if (!t_start[k]) continue;
				hist[0][std::min<unsigned long long>(255, (t_start[k] - t0) / 50000)]++;
				if (t_exh[k]) hist[1][std::min<unsigned long long>(255, (t_exh[k] - t0) / 50000)]++;
				if (t_exit[k]) hist[2][std::min<unsigned long long>(255, (t_exit[k] - t0) / 50000)]++;
			}
			{
				static unsigned long long pk_start[RT_PROF_PACKETS];
				static unsigned int pk_dur[RT_PROF_PACKETS];
				if (cudaMemcpyFromSymbol(pk_start, g_pk_start, sizeof pk_start) == cudaSuccess && cudaMemcpyFromSymbol(pk_dur, g_pk_dur, sizeof pk_dur) == cudaSuccess) {
					unsigned long long p0 = ~0ull, p1 = 0;
					int n = 0;
					for (int k = 0; k < RT_PROF_PACKETS; k++)
						if (pk_start[k]) { p0 = std::min(p0, pk_start[k]); p1 = std::max(p1, pk_start[k] + pk_dur[k]); n = k + 1; }
					if (n) {
						fprintf(stderr, "[primary stage (last launch): %d packets, span %.1f us]\n", n, (p1 - p0) / 1e3);
						unsigned started[64] = {0}, ended[64] = {0};
						double dur_sum[64] = {0};
						for (int k = 0; k < n; k++) {
							if (!pk_start[k]) continue;
							const int b0 = (int)std::min<unsigned long long>(63, (pk_start[k] - p0) / 10000), b1 = (int)std::min<unsigned long long>(63, (pk_start[k] + pk_dur[k] - p0) / 10000);
							started[b0]++; ended[b1]++; dur_sum[b0] += pk_dur[k] / 1e3;
						}
						fprintf(stderr, "  10 us buckets: packets started / finished / mean duration (us) of those started\n");
						for (int b = 0; b < 64; b++)
							if (started[b] || ended[b]) fprintf(stderr, "  %4d us %6u %6u %8.1f\n", b * 10, started[b], ended[b], started[b] ? dur_sum[b] / started[b] : 0.0);
						// by position in the dispenser's order (= tile order, row-major): mean duration of each 1/16 of the packets
						fprintf(stderr, "  mean packet duration (us) by sixteenth of the frame:");
						for (int q = 0; q < 16; q++) {
							double sum = 0; int c = 0;
							for (int k = n * q / 16; k < n * (q + 1) / 16; k++) if (pk_start[k]) { sum += pk_dur[k] / 1e3; c++; }
							fprintf(stderr, " %.1f", c ? sum / c : 0.0);
						}
						fprintf(stderr, "\n");
					}
				}
			}
			fprintf(stderr, "[bounce warps (last launch), 50 us buckets: started / queue exhausted / finished]\n");
			for (int k = 0; k < 256; k++)
				if (hist[0][k] || hist[1][k] || hist[2][k]) fprintf(stderr, "  %5d us  %6u %6u %6u\n", k * 50, hist[0][k], hist[1][k], hist[2][k]);
		}
	}
#endif
	if (!ctx) return;
	if (ctx->group.size() > 1) {  // a leader: stop and free the other members first
		std::vector<rt_ctx*> members = ctx->group;
		ctx->group.clear();
		for (size_t r = 1; r < members.size(); r++) {
			rt_ctx* m = members[r];
			if (m->worker.joinable()) {
				{
					std::lock_guard<std::mutex> lk(m->wm);
					m->w_quit = true;
					m->w_posted.fetch_add(1, std::memory_order_release);
				}
				m->wcv.notify_one();
				m->worker.join();
			}
			m->leader = nullptr;
			rt_destroy(m);
		}
		cudaSetDevice(ctx->device);
		for (auto& hm : ctx->hostmaps)
			if (hm.registered) cudaHostUnregister(hm.host);
		ctx->hostmaps.clear();
	}
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
	for (int k = 0; k < 2; k++) {
		ctx->pipe_rgb[k].release();
		ctx->pipe_ids[k].release();
		if (ctx->pipe_kernels[k]) cudaEventDestroy(ctx->pipe_kernels[k]);
		if (ctx->pipe_done[k]) cudaEventDestroy(ctx->pipe_done[k]);
	}
	if (ctx->pipe_counters) cudaFreeHost(ctx->pipe_counters);
	if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
	if (ctx->done_ev) cudaEventDestroy(ctx->done_ev);
	cudaStreamSynchronize(ctx->stream);
	ctx->node_geom.release(); ctx->node_geom64.release(); ctx->node_link.release(); ctx->node_child.release(); ctx->node_pk.release(); ctx->node_walk.release(); ctx->node_bvh.release(); ctx->bvh_nodes.release(); ctx->bvh_slots.release(); ctx->bvh_geom.release();
	ctx->slot_geom.release(); ctx->slot_geom64.release(); ctx->slot_attr.release(); ctx->slot_node.release();
	ctx->materials.release(); ctx->textures.release(); ctx->substances.release(); ctx->texels.release();
	ctx->ray_ck.release(); ctx->row_fr.release(); ctx->rgb.release(); ctx->ids.release();
	ctx->counters.release(); ctx->l2_scratch.release(); ctx->prim_geom.release(); ctx->queue.release(); ctx->vqueue.release(); ctx->queue_dense.release(); ctx->present_partial.release(); ctx->rgba.release(); ctx->samples.release(); ctx->scene_pool.release(); ctx->peer_flags.release();
	if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
	if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
	for (int b = 0; b < RT_MAX_BANDS; b++)
		if (ctx->band_done[b]) cudaEventDestroy(ctx->band_done[b]);
	for (int k = 0; k <= RT_N_STAGES; k++)
		if (ctx->stage_ev[k]) cudaEventDestroy(ctx->stage_ev[k]);
	if (ctx->stage_free) cudaEventDestroy(ctx->stage_free);
	if (ctx->stage) cudaFreeHost(ctx->stage);
	if (ctx->ev0) cudaEventDestroy(ctx->ev0);
	if (ctx->ev1) cudaEventDestroy(ctx->ev1);
	if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
	delete ctx;
}

rt_status rt_set_stream(rt_ctx* ctx, void* cuda_stream) {
	if (!ctx) return RT_ERR_INVALID;
	RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
	return RT_OK;
}

rt_status rt_synchronize(rt_ctx* ctx) {
	if (!ctx) return RT_ERR_INVALID;
	for (size_t r = 1; r < ctx->group.size(); r++) {
		RT_CUDA(ctx, cudaSetDevice(ctx->group[r]->device));
		RT_CUDA(ctx, cudaStreamSynchronize(ctx->group[r]->stream));
	}
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return RT_OK;
}

rt_status rt_timer_start(rt_ctx* ctx) {
	if (!ctx) return RT_ERR_INVALID;
	RT_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	return RT_OK;
}

rt_status rt_timer_stop(rt_ctx* ctx, float* elapsed_ms) {
	if (!ctx || !elapsed_ms) return RT_ERR_INVALID;
	RT_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
	RT_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
	RT_CUDA(ctx, cudaEventElapsedTime(elapsed_ms, ctx->ev0, ctx->ev1));
	return RT_OK;
}

uint64_t rt_launch_count(const rt_ctx* ctx) { return ctx ? ctx->launches : 0; }

rt_status rt_set_profiling(rt_ctx* ctx, int32_t on) {
	if (!ctx) return RT_ERR_INVALID;
	ctx->profile = on != 0;
	return RT_OK;
}

rt_status rt_stage_times(rt_ctx* ctx, float ms[RT_N_STAGES]) {
	if (!ctx || !ms) return RT_ERR_INVALID;
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	for (int k = 0; k < RT_N_STAGES; k++) {
		ms[k] = -1.0f;
		if (ctx->stage_ran[k]) RT_CUDA(ctx, cudaEventElapsedTime(&ms[k], ctx->stage_ev[k], ctx->stage_ev[k + 1]));
	}
	return RT_OK;
}

rt_status rt_flush_l2(rt_ctx* ctx) {
	if (!ctx) return RT_ERR_INVALID;
	const size_t bytes = 256u << 20;
	RT_CUDA(ctx, ctx->l2_scratch.alloc(bytes));
	RT_CUDA(ctx, cudaMemsetAsync(ctx->l2_scratch.p, 0x5a, bytes, ctx->stream));
	return RT_OK;
}

rt_status rt_host_register(rt_ctx* ctx, void* ptr, size_t bytes) {
	if (!ctx || !ptr || !bytes) return RT_ERR_INVALID;
	RT_CUDA(ctx, cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
	return RT_OK;
}

rt_status rt_host_map(rt_ctx* ctx, void* ptr, size_t bytes, void** dev_ptr) {
	if (!ctx || !ptr || !bytes || !dev_ptr) return RT_ERR_INVALID;
	*dev_ptr = nullptr;
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	RT_CUDA(ctx, cudaHostRegister(ptr, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
	const cudaError_t e = cudaHostGetDevicePointer(dev_ptr, ptr, 0);
	if (e != cudaSuccess) {
		cudaHostUnregister(ptr);
		return fail(ctx, RT_ERR_CUDA, rt_format("cudaHostGetDevicePointer: %s", cudaGetErrorString(e)));
	}
	return RT_OK;
}

rt_status rt_host_unregister(rt_ctx* ctx, void* ptr) {
	if (!ctx || !ptr) return RT_ERR_INVALID;
	RT_CUDA(ctx, cudaHostUnregister(ptr));
	return RT_OK;
}

// every device array of a scene, with its host vector of the same name in RtHostScene
#define RT_SCENE_ARRAYS(X) \
	X(node_geom) X(node_geom64) X(node_link) X(node_child) X(node_pk) X(node_walk) X(node_bvh) X(bvh_nodes) X(bvh_slots) X(bvh_geom) \
	X(slot_geom) X(slot_geom64) X(slot_attr) X(slot_node) X(materials) X(textures) X(substances) X(texels)

// All scene arrays of a ctx as slices of one pool allocation (grow-only: a re-upload of a scene that fits allocates
// nothing), each aligned to 256 bytes.
static cudaError_t alloc_scene_arrays(rt_ctx* m, const RtHostScene& H) {
	size_t total = 0;
#define X(name) total += (std::max<size_t>(H.name.size(), 1) * sizeof(H.name[0]) + 255) & ~(size_t)255;
	RT_SCENE_ARRAYS(X)
#undef X
	if (total > m->scene_pool.n || !m->scene_pool.p) {
#define X(name) m->name.release();
		RT_SCENE_ARRAYS(X)
#undef X
		if (cudaError_t e = m->scene_pool.alloc(total + total / 8)) return e;  // (headroom: a dynamic scene grows by a few nodes per update)
	}
	size_t off = 0;
#define X(name)                                                                        \
	m->name.adopt(m->scene_pool.p + off, H.name.size());                               \
	off += (std::max<size_t>(H.name.size(), 1) * sizeof(H.name[0]) + 255) & ~(size_t)255;
	RT_SCENE_ARRAYS(X)
#undef X
	return cudaSuccess;
}

static void set_dev_scene(rt_ctx* ctx) {
	const RtHostScene& H = ctx->host_ref();
	RtDevScene& D = ctx->dev;
#define X(name) D.name = ctx->name.p;
	RT_SCENE_ARRAYS(X)
#undef X
	for (int k = 0; k < 3; k++) D.root_pos[k] = H.root_pos[k];
	D.root_size = H.root_size;
	D.n_nodes = (int)H.node_geom.size();
	D.n_slots = (int)H.slot_geom.size();
	D.err_l = H.err_l;
	D.ordered_ok = rt_ordered_walk_fits(H);
}

static rt_status scene_upload_impl(rt_ctx* ctx, const rt_scene_desc* sc);

rt_status rt_scene_upload(rt_ctx* ctx, const rt_scene_desc* sc) {
	if (!ctx) return RT_ERR_INVALID;
	return abi_guard(ctx, "rt_scene_upload", [&] { return scene_upload_impl(ctx, sc); });
}

static rt_status scene_upload_impl(rt_ctx* ctx, const rt_scene_desc* sc) {
	if (ctx->leader) return fail(ctx, RT_ERR_INVALID, "rt_scene_upload: this ctx is a member of a multi-GPU group; upload through its leader");
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	std::string err;
	RtPackClock clk;  // (RT_B200_PACK_TIMING=1: the stages of the upload on stderr)
	auto hs = std::make_shared<RtHostScene>();
	// the library's own copy of the description (what rt_scene_update edits) is taken beside the packing, on a thread
	// of its own - the caller's arrays are only borrowed for the duration of this call - and kept only if the scene
	// is accepted (not when rt_scene_update hands its own copy back)
	const bool take_copy = sc && sc->struct_size == sizeof(rt_scene_desc) && sc->n_nodes > 0 && sc->node_size &&
	                       sc->node_size != ctx->scene_copy.node_size.data();
	RtSceneCopy fresh;
	std::thread copier;
	std::exception_ptr copy_thrown, pack_thrown;
	if (take_copy)
		copier = std::thread([&] {
			try {
				fresh.assign(*sc);
			} catch (...) {
				copy_thrown = std::current_exception();
			}
		});
	rt_status pack_st = RT_OK;
	try {
		pack_st = rt_pack_scene(sc, *hs, err);  // validated and packed ONCE, on the host
	} catch (...) {
		pack_thrown = std::current_exception();
	}
	if (copier.joinable()) copier.join();
	if (pack_thrown) std::rethrow_exception(pack_thrown);  // (caught at the ABI: abi_guard)
	if (copy_thrown) std::rethrow_exception(copy_thrown);
	if (pack_st) return fail(ctx, pack_st, err);
	if (take_copy) ctx->scene_copy = std::move(fresh);
	clk.mark("= pack + description copy");
	std::vector<rt_ctx*> all = ctx->group.empty() ? std::vector<rt_ctx*>{ctx} : ctx->group;
	for (rt_ctx* m : all) {
		m->has_scene = false;
		m->scene_version++;
		m->key_valid = false;
		m->host_p = hs;
	}
	const RtHostScene& H = *hs;
	rt_status st = RT_OK;
	RT_CUDA(ctx, alloc_scene_arrays(ctx, H));
	clk.mark("device alloc");
#define X(name) if (!st) st = upload(ctx, ctx->name, H.name);
	RT_SCENE_ARRAYS(X)
#undef X
	if (st) return st;
	RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	clk.mark("H2D");
	set_dev_scene(ctx);
	ctx->has_scene = true;
	// the other GPUs of a group get their replica device to device (NVLink / NVSwitch), not from the host again
	for (size_t r = 1; r < all.size(); r++) {
		rt_ctx* m = all[r];
		RT_CUDA(ctx, cudaSetDevice(m->device));
		RT_CUDA(ctx, alloc_scene_arrays(m, H));
#define X(name)                                                                                                         \
		if (!H.name.empty())                                                                                                \
			RT_CUDA(ctx, cudaMemcpyPeerAsync(m->name.p, m->device, ctx->name.p, ctx->device, H.name.size() * sizeof(H.name[0]), m->stream));
		RT_SCENE_ARRAYS(X)
#undef X
		RT_CUDA(ctx, cudaStreamSynchronize(m->stream));
		set_dev_scene(m);
		m->has_scene = true;
	}
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	if (all.size() > 1) clk.mark("replication");
	return RT_OK;
}

rt_status rt_scene_update(rt_ctx* ctx, uint32_t n_moved, const uint32_t* entity_ids, const double* new_pos, uint32_t max_in_depth) {
	if (!ctx) return RT_ERR_INVALID;
	if (ctx->leader) return fail(ctx, RT_ERR_INVALID, "rt_scene_update: this ctx is a member of a multi-GPU group; update through its leader");
	if (!ctx->has_scene || !ctx->scene_copy.valid) return fail(ctx, RT_ERR_NO_SCENE, "rt_scene_update: no scene uploaded");
	if (n_moved && (!entity_ids || !new_pos)) return fail(ctx, RT_ERR_INVALID, "rt_scene_update: NULL argument");
	if (n_moved == 0) return RT_OK;
	return abi_guard(ctx, "rt_scene_update", [&]() -> rt_status {
		std::string err;
		if (!rt_scene_move_entities(ctx->scene_copy, n_moved, entity_ids, new_pos, max_in_depth, err))
			return fail(ctx, err.find("outside-depth") != std::string::npos ? RT_ERR_UNSUPPORTED : RT_ERR_INVALID, err);
		const rt_scene_desc d = ctx->scene_copy.desc();
		return scene_upload_impl(ctx, &d);
	});
}

rt_status rt_render_device(rt_ctx* ctx, const rt_camera* cam, const rt_params* prm, uint32_t flags, float* rgb_dev,
                           int32_t* ids_dev) {
	if (rt_status st = check_args(ctx, cam, prm)) return st;
	if (!rgb_dev) return fail(ctx, RT_ERR_INVALID, "rt_render_device: rgb_dev is NULL");
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	return render_device_any(ctx, cam, prm, flags, rgb_dev, ids_dev);
}

rt_status rt_camera_directions(rt_ctx* ctx, const rt_camera* cam, double* dirs) {
	if (!ctx) return RT_ERR_INVALID;
	if (!cam || !dirs) return fail(ctx, RT_ERR_INVALID, "rt_camera_directions: NULL argument");
	if (cam->width == 0 || cam->height == 0 || cam->width > 65536 || cam->height > 65536)
		return fail(ctx, RT_ERR_INVALID, rt_format("rt_camera_directions: bad frame size %ux%u", cam->width, cam->height));
	if ((cam->flags & RT_CAM_REFERENCE_EXTENTS) && cam->width != cam->height) return fail(ctx, RT_ERR_BOUNDS, "x or y out of bounds");
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	RtFrame F{};
	for (int i = 0; i < 3; i++) F.lf[i] = cam->lf[i];
	F.width = (int)cam->width;
	F.height = (int)cam->height;
	F.tile_world = 1;
	std::vector<RtD4> rows;
	rt_build_camera_rows(*cam, rows, F.scan_cos, F.scan_sin);
	const size_t npx = (size_t)F.width * F.height;
	F.ray_ckh = raygen_checkpoints_per_half(F.width);
	RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // (pageable source below: nothing may still read row_fr)
	RT_CUDA(ctx, ctx->row_fr.alloc(rows.size()));
	RT_CUDA(ctx, ctx->ray_ck.alloc((size_t)F.height * 2 * F.ray_ckh * 6));
	DevBuf<double> out;
	RT_CUDA(ctx, out.alloc(npx * 3));
	ctx->key_valid = false;  // the tables of a cached frame are gone
	ctx->raygen_key_valid = false;
	RT_CUDA(ctx, cudaMemcpyAsync(ctx->row_fr.p, rows.data(), rows.size() * sizeof(RtD4), cudaMemcpyHostToDevice, ctx->stream));
	F.row_fr = ctx->row_fr.p;
	F.ray_ck = ctx->ray_ck.p;
	rt_raygen_kernel<<<(6 * F.height + RT_SETUP_THREADS - 1) / RT_SETUP_THREADS, RT_SETUP_THREADS, 0, ctx->stream>>>(F, ctx->ray_ck.p);
	rt_expand_dirs_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, ctx->stream>>>(F, out.p);
	ctx->launches += 2;
	cudaError_t e = cudaGetLastError();
	if (e == cudaSuccess) e = cudaMemcpyAsync(dirs, out.p, npx * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	out.release();
	if (e != cudaSuccess) return fail(ctx, RT_ERR_CUDA, rt_format("rt_camera_directions: %s", cudaGetErrorString(e)));
	return RT_OK;
}

uint32_t rt_tiles_per_rank(uint32_t width, uint32_t height, uint32_t world) {
	if (!world) return 0;
	const uint32_t n = ((width + RT_TILE_W - 1) / RT_TILE_W) * ((height + RT_TILE_H - 1) / RT_TILE_H);
	return (n + world - 1) / world;
}

rt_status rt_render_tiles_device(rt_ctx* ctx, const rt_camera* cam, const rt_params* prm, uint32_t flags,
                                 uint32_t rank, uint32_t world, float* tiles_dev, int32_t* tile_ids_dev) {
	if (rt_status st = check_args(ctx, cam, prm)) return st;
	if (!tiles_dev) return fail(ctx, RT_ERR_INVALID, "rt_render_tiles_device: tiles_dev is NULL");
	if (world == 0 || rank >= world) return fail(ctx, RT_ERR_INVALID, rt_format("bad tile shard %u of %u", rank, world));
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	return launch_render(ctx, cam, prm, flags, tiles_dev, tile_ids_dev, (int)rank, (int)world, true);
}

rt_status rt_render_shard_device(rt_ctx* ctx, const rt_camera* cam, const rt_params* prm, uint32_t flags, uint32_t rank,
                                 uint32_t world, float* frame_dev, int32_t* ids_dev) {
	if (rt_status st = check_args(ctx, cam, prm)) return st;
	if (!frame_dev) return fail(ctx, RT_ERR_INVALID, "rt_render_shard_device: frame_dev is NULL");
	if (world == 0 || rank >= world) return fail(ctx, RT_ERR_INVALID, rt_format("bad tile shard %u of %u", rank, world));
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	return launch_render(ctx, cam, prm, flags, frame_dev, ids_dev, (int)rank, (int)world, false);
}

rt_status rt_peer_alloc(rt_ctx* ctx, size_t bytes, void** dev_ptr, unsigned char handle_out[RT_PEER_HANDLE_BYTES]) {
	static_assert(sizeof(cudaIpcMemHandle_t) == RT_PEER_HANDLE_BYTES, "CUDA IPC handle size");
	if (!ctx || !dev_ptr || !handle_out || !bytes) return RT_ERR_INVALID;
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	void* p = nullptr;
	RT_CUDA(ctx, cudaMalloc(&p, bytes));
	cudaError_t e = cudaMemset(p, 0, bytes);
	cudaIpcMemHandle_t h;
	if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
	if (e != cudaSuccess) {
		cudaFree(p);
		return fail(ctx, RT_ERR_CUDA, rt_format("rt_peer_alloc: %s", cudaGetErrorString(e)));
	}
	memcpy(handle_out, &h, sizeof h);
	*dev_ptr = p;
	return RT_OK;
}

rt_status rt_peer_free(rt_ctx* ctx, void* dev_ptr) {
	if (!ctx || !dev_ptr) return RT_ERR_INVALID;
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	RT_CUDA(ctx, cudaFree(dev_ptr));
	return RT_OK;
}

rt_status rt_peer_open(rt_ctx* ctx, const unsigned char handle[RT_PEER_HANDLE_BYTES], void** dev_ptr) {
	if (!ctx || !handle || !dev_ptr) return RT_ERR_INVALID;
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, sizeof h);
	RT_CUDA(ctx, cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
	return RT_OK;
}

rt_status rt_peer_close(rt_ctx* ctx, void* dev_ptr) {
	if (!ctx || !dev_ptr) return RT_ERR_INVALID;
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	RT_CUDA(ctx, cudaIpcCloseMemHandle(dev_ptr));
	return RT_OK;
}

rt_status rt_peer_barrier(rt_ctx* ctx, uint32_t rank, uint32_t world, uint32_t* const* flags, uint32_t epoch) {
	if (!ctx || !flags || world == 0 || world > 64 || rank >= world) return RT_ERR_INVALID;
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	RT_CUDA(ctx, ctx->peer_flags.alloc(64));
	if (ctx->peer_flags_host.size() != world || memcmp(ctx->peer_flags_host.data(), flags, world * sizeof(uint32_t*)) != 0) {
		ctx->peer_flags_host.assign(flags, flags + world);
		RT_CUDA(ctx, cudaMemcpyAsync(ctx->peer_flags.p, ctx->peer_flags_host.data(), world * sizeof(uint32_t*),
		                             cudaMemcpyHostToDevice, ctx->stream));
	}
	rt_peer_barrier_kernel<<<1, 64, 0, ctx->stream>>>(ctx->peer_flags.p, (int)rank, (int)world, epoch);
	ctx->launches++;
	RT_CUDA(ctx, cudaGetLastError());
	return RT_OK;
}

rt_status rt_image_decode(const uint8_t* bytes, uint64_t n_bytes, uint32_t* width, uint32_t* height, uint8_t** rgb) {
	if (!bytes || !width || !height || !rgb) return fail(nullptr, RT_ERR_INVALID, "rt_image_decode: NULL argument");
	*rgb = nullptr;
	*width = *height = 0;
	return abi_guard(nullptr, "rt_image_decode", [&]() -> rt_status {
		std::vector<uint8_t> px;
		std::string err;
		uint32_t w = 0, h = 0;
		if (!rt_image::decode(bytes, (size_t)n_bytes, w, h, px, err)) return fail(nullptr, RT_ERR_UNSUPPORTED, err);
		uint8_t* out = static_cast<uint8_t*>(malloc(px.size()));
		if (!out) return fail(nullptr, RT_ERR_INVALID, "rt_image_decode: out of host memory");
		memcpy(out, px.data(), px.size());
		*rgb = out;
		*width = w;
		*height = h;
		return RT_OK;
	});
}

void rt_image_free(uint8_t* rgb) { free(rgb); }

rt_status rt_tree_build(const double root_pos[3], double root_size, uint32_t n, const uint8_t* type, const double* pos,
                        const double* extent, uint32_t max_in_depth, rt_tree** out) {
	if (!out) return fail(nullptr, RT_ERR_INVALID, "rt_tree_build: out is NULL");
	*out = nullptr;
	if (!root_pos || !(root_size > 0) || (n && (!type || !pos || !extent)))
		return fail(nullptr, RT_ERR_INVALID, "rt_tree_build: bad arguments");
	return abi_guard(nullptr, "rt_tree_build", [&]() -> rt_status {
		std::unique_ptr<rt_tree> t(new rt_tree());
		std::string err;
		if (!rt_tree_build_impl(*t, root_pos, root_size, n, type, pos, extent, max_in_depth, err)) return fail(nullptr, RT_ERR_UNSUPPORTED, err);
		*out = t.release();
		return RT_OK;
	});
}

rt_status rt_tree_build_gpu(rt_ctx* ctx, const double root_pos[3], double root_size, uint32_t n, const uint8_t* type,
                            const double* pos, const double* extent, uint32_t max_in_depth, rt_tree** out) {
	if (!ctx) return RT_ERR_INVALID;
	if (!out) return fail(ctx, RT_ERR_INVALID, "rt_tree_build_gpu: out is NULL");
	*out = nullptr;
	if (!root_pos || !(root_size > 0) || (n && (!type || !pos || !extent)))
		return fail(ctx, RT_ERR_INVALID, "rt_tree_build_gpu: bad arguments");
	if (max_in_depth > RT_GPU_BUILD_MAX_DEPTH)
		return fail(ctx, RT_ERR_UNSUPPORTED, rt_format("rt_tree_build_gpu: max_in_depth %u > %d (path keys hold 16 levels); use rt_tree_build", max_in_depth, RT_GPU_BUILD_MAX_DEPTH));
	if ((uint64_t)n * (max_in_depth + 1) + 1 > 0x7fffffffull)
		return fail(ctx, RT_ERR_UNSUPPORTED, "rt_tree_build_gpu: too many entities for one sort");
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	return abi_guard(ctx, "rt_tree_build_gpu", [&]() -> rt_status {
		std::unique_ptr<rt_tree> t(new rt_tree());
		std::string err;
		const int rc = rt_gpu_build::build(*t, ctx->stream, root_pos, root_size, n, type, pos, extent, max_in_depth, err, &ctx->launches);
		if (rc != 0) return fail(ctx, rc == 1 ? RT_ERR_UNSUPPORTED : RT_ERR_CUDA, err);
		*out = t.release();
		return RT_OK;
	});
}

uint32_t rt_tree_node_count(const rt_tree* t) { return t ? (uint32_t)t->size.size() : 0; }

void rt_tree_export(const rt_tree* t, double* node_pos, double* node_size, int32_t* node_child, int32_t* node_parent,
                    int32_t* node_octant, uint32_t* node_list_off, uint32_t* list_entity) {
	if (!t) return;
	if (node_pos) memcpy(node_pos, t->pos.data(), t->pos.size() * sizeof(double));
	if (node_size) memcpy(node_size, t->size.data(), t->size.size() * sizeof(double));
	if (node_child) memcpy(node_child, t->child.data(), t->child.size() * sizeof(int32_t));
	if (node_parent) memcpy(node_parent, t->parent.data(), t->parent.size() * sizeof(int32_t));
	if (node_octant) memcpy(node_octant, t->octant.data(), t->octant.size() * sizeof(int32_t));
	if (node_list_off) memcpy(node_list_off, t->list_off.data(), t->list_off.size() * sizeof(uint32_t));
	if (list_entity && !t->list_entity.empty()) memcpy(list_entity, t->list_entity.data(), t->list_entity.size() * sizeof(uint32_t));
}

void rt_tree_free(rt_tree* t) { delete t; }

void rt_fplcg_fill(double seed, uint64_t n, double* out) {
	// host restatement of FpLcg (src/math/rng/fp-lcg.ts:62-82); `% 1.0` on non-negative values
	double s1 = seed, s2 = seed * RT_LCG_MUL3, s3 = seed * RT_LCG_MUL2;
	for (uint64_t i = 0; i < n; i++) {
		const double a = fmod(s1 * RT_LCG_MUL1 + RT_LCG_TERM1, 1.0);
		const double b = fmod(s2 * RT_LCG_MUL2 + RT_LCG_TERM2, 1.0);
		const double c = fmod(s3 * RT_LCG_MUL3 + RT_LCG_TERM3, 1.0);
		s1 = b + c;
		s2 = c;
		s3 = a + b;
		out[i] = fmod(a + b + c, 1.0);
	}
}

rt_status rt_untile_device(rt_ctx* ctx, uint32_t width, uint32_t height, uint32_t world, const float* gathered_dev,
                           float* rgb_dev) {
	if (!ctx) return RT_ERR_INVALID;
	if (!gathered_dev || !rgb_dev || !world || !width || !height) return fail(ctx, RT_ERR_INVALID, "rt_untile_device: bad arguments");
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	const int tiles_x = (width + RT_TILE_W - 1) / RT_TILE_W;
	const dim3 block(32, 8), grid((width + 31) / 32, (height + 7) / 8);
	rt_untile_kernel<<<grid, block, 0, ctx->stream>>>(gathered_dev, rgb_dev, (int)width, (int)height, tiles_x, (int)world,
	                                                  (int)rt_tiles_per_rank(width, height, world));
	ctx->launches++;
	RT_CUDA(ctx, cudaGetLastError());
	return RT_OK;
}

rt_status rt_get_counters(rt_ctx* ctx, rt_counters* out) {
	if (!ctx || !out) return RT_ERR_INVALID;
	return read_counters_any(ctx, out);
}

rt_status rt_present_device(rt_ctx* ctx, const float* rgb_dev, uint32_t width, uint32_t height, const rt_tone* tone,
                            uint8_t* rgba_dev) {
	if (!ctx || !rgb_dev || !rgba_dev || !tone) return fail(ctx, RT_ERR_INVALID, "rt_present_device: NULL argument");
	if (!width || !height) return fail(ctx, RT_ERR_INVALID, "rt_present_device: empty frame");
	if (tone->kind > RT_TONE_ABSDEV) return fail(ctx, RT_ERR_UNSUPPORTED, rt_format("unsupported ToneMapper subclass (kind %u)", tone->kind));
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	if (ctx->present_blocks == 0) {
		int sms = 0;
		RT_CUDA(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
		ctx->present_blocks = std::max(1, sms) * 8;  // 8 CTAs of 256 threads per SM: a multiple of the SM count
	}
	const size_t npx = (size_t)width * height;
	const int blocks = (int)std::min<size_t>((size_t)ctx->present_blocks, (npx + RT_PRESENT_THREADS - 1) / RT_PRESENT_THREADS);
	RT_CUDA(ctx, ctx->present_partial.alloc((size_t)3 * ctx->present_blocks + 8));
	double* part = ctx->present_partial.p;
	rt_luma_sum_kernel<<<blocks, RT_PRESENT_THREADS, 0, ctx->stream>>>(rgb_dev, npx, part);
	rt_luma_dev_kernel<<<blocks, RT_PRESENT_THREADS, 0, ctx->stream>>>(rgb_dev, npx, part, part + blocks, part + 2 * blocks);
	rt_discretize_kernel<<<blocks, RT_PRESENT_THREADS, 0, ctx->stream>>>(rgb_dev, npx, part, blocks, *tone, reinterpret_cast<uchar4*>(rgba_dev),
	                                                                     part + (size_t)3 * ctx->present_blocks);
	ctx->launches += 3;
	RT_CUDA(ctx, cudaGetLastError());
	return RT_OK;
}

rt_status rt_present_stats(rt_ctx* ctx, rt_exposure_stats* out) {
	if (!ctx || !out) return RT_ERR_INVALID;
	if (!ctx->present_partial.p) return fail(ctx, RT_ERR_INVALID, "rt_present_stats: nothing was presented yet");
	double h[5];
	RT_CUDA(ctx, cudaMemcpyAsync(h, ctx->present_partial.p + (size_t)3 * ctx->present_blocks, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
	RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	out->mean = h[0]; out->variance = h[1]; out->absolute_dev = h[2]; out->drange_low = h[3]; out->drange_high = h[4];
	return RT_OK;
}

rt_status rt_present(rt_ctx* ctx, const float* rgb, uint32_t width, uint32_t height, const rt_tone* tone, uint8_t* rgba,
                     rt_exposure_stats* stats) {
	if (!ctx || !rgb || !rgba) return fail(ctx, RT_ERR_INVALID, "rt_present: NULL argument");
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	const size_t npx = (size_t)width * height;
	RT_CUDA(ctx, ctx->rgb.alloc(npx * 3));
	RT_CUDA(ctx, ctx->rgba.alloc(npx * 4));
	ctx->exposure_w = ctx->exposure_h = 0;  // ctx->rgb no longer holds a resident exposure
	RT_CUDA(ctx, cudaMemcpyAsync(ctx->rgb.p, rgb, npx * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
	if (rt_status st = rt_present_device(ctx, ctx->rgb.p, width, height, tone, ctx->rgba.p)) return st;
	RT_CUDA(ctx, cudaMemcpyAsync(rgba, ctx->rgba.p, npx * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (stats) return rt_present_stats(ctx, stats);
	RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return RT_OK;
}

rt_status rt_render_present(rt_ctx* ctx, const rt_camera* cam, const rt_params* prm, uint32_t flags, const rt_tone* tone,
                            uint8_t* rgba, rt_exposure_stats* stats, rt_counters* counters) {
	if (rt_status st = check_args(ctx, cam, prm)) return st;
	if (!rgba || !tone) return fail(ctx, RT_ERR_INVALID, "rt_render_present: NULL argument");
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	const size_t npx = (size_t)cam->width * cam->height;
	if (prm->frame_first > 0 && (ctx->exposure_w != cam->width || ctx->exposure_h != cam->height))
		return fail(ctx, RT_ERR_INVALID, "rt_render_present: frame_first > 0 but no resident ExposureBuffer of this size (start the exposure with frame_first = 0)");
	RT_CUDA(ctx, ctx->rgb.alloc(npx * 3));
	RT_CUDA(ctx, ctx->rgba.alloc(npx * 4));
	ctx->exposure_w = cam->width;
	ctx->exposure_h = cam->height;
	if (counters) flags |= RT_RENDER_COUNTERS;
	if (rt_status st = render_device_any(ctx, cam, prm, flags, ctx->rgb.p, nullptr)) return st;
	if (rt_status st = rt_present_device(ctx, ctx->rgb.p, cam->width, cam->height, tone, ctx->rgba.p)) return st;
	RT_CUDA(ctx, cudaMemcpyAsync(rgba, ctx->rgba.p, npx * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (stats)
		if (rt_status st = rt_present_stats(ctx, stats)) return st;
	rt_counters tmp;
	if (rt_status st = read_counters_any(ctx, &tmp)) return st;  // also synchronises and fetches the error flags
	if (counters) *counters = tmp;
	if (tmp.texture_errors) return fail(ctx, RT_ERR_TEXTURE, "Texture coordinates out of bounds");
	return RT_OK;
}

rt_status rt_exposure_download(rt_ctx* ctx, float* rgb, uint32_t width, uint32_t height) {
	if (!ctx || !rgb) return RT_ERR_INVALID;
	if (!ctx->exposure_w || ctx->exposure_w != width || ctx->exposure_h != height)
		return fail(ctx, RT_ERR_INVALID, "rt_exposure_download: no resident ExposureBuffer of this size");
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	RT_CUDA(ctx, cudaMemcpyAsync(rgb, ctx->rgb.p, (size_t)width * height * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return RT_OK;
}

rt_status rt_render(rt_ctx* ctx, const rt_camera* cam, const rt_params* prm, uint32_t flags, float* rgb,
                    int32_t* first_ids, rt_counters* counters) {
	if (rt_status st = check_args(ctx, cam, prm)) return st;
	if (!rgb) return fail(ctx, RT_ERR_INVALID, "rt_render: rgb is NULL");
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	if (ctx->group.size() > 1) return group_render_host(ctx, cam, prm, flags, rgb, first_ids, counters);
	const size_t npx = (size_t)cam->width * cam->height;
	RT_CUDA(ctx, ctx->rgb.alloc(npx * 3));
	ctx->exposure_w = ctx->exposure_h = 0;  // ctx->rgb is this call's staging, not a resident exposure
	if (first_ids) RT_CUDA(ctx, ctx->ids.alloc(npx));
	if (prm->frame_first > 0)
		RT_CUDA(ctx, cudaMemcpyAsync(ctx->rgb.p, rgb, npx * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
	if (counters) flags |= RT_RENDER_COUNTERS;
	// The frame is rendered in bands of tile rows; each finished band starts its way to the host on the copy
	// stream (behind an event) while the next band renders, so the PCIe transfer overlaps the kernels.
	const size_t W = cam->width;
	const BandHook copy_band = [&](int band, int y_begin, int y_end) -> rt_status {
		RT_CUDA(ctx, cudaEventRecord(ctx->band_done[band], ctx->stream));
		RT_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->band_done[band], 0));
		const size_t off = (size_t)y_begin * W, n = (size_t)(y_end - y_begin) * W;
		RT_CUDA(ctx, cudaMemcpyAsync(rgb + off * 3, ctx->rgb.p + off * 3, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_stream));
		if (first_ids)
			RT_CUDA(ctx, cudaMemcpyAsync(first_ids + off, ctx->ids.p + off, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->copy_stream));
		return RT_OK;
	};
	if (rt_status st = launch_render(ctx, cam, prm, flags, ctx->rgb.p, first_ids ? ctx->ids.p : nullptr, 0, 1, false,
	                                 host_bands(ctx, prm), copy_band))
		return st;
	RT_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
	rt_counters tmp;
	if (rt_status st = read_counters(ctx, &tmp)) return st;  // also synchronises and fetches the error flags
	if (counters) *counters = tmp;
	if (tmp.texture_errors) return fail(ctx, RT_ERR_TEXTURE, "Texture coordinates out of bounds");
	return RT_OK;
}

// ---- pipelined frames: the copy of frame k to the host overlaps the rendering of frame k + 1 -----------------------
rt_status rt_render_begin(rt_ctx* ctx, const rt_camera* cam, const rt_params* prm, uint32_t flags, float* rgb, int32_t* first_ids) {
	if (rt_status st = check_args(ctx, cam, prm)) return st;
	if (!rgb) return fail(ctx, RT_ERR_INVALID, "rt_render_begin: rgb is NULL");
	if (ctx->group.size() > 1) return fail(ctx, RT_ERR_UNSUPPORTED, "rt_render_begin: not on a multi-GPU group (its members already deliver their tiles concurrently)");
	if (ctx->pipe_begun - ctx->pipe_ended >= 2) return fail(ctx, RT_ERR_INVALID, "rt_render_begin: two frames are already in flight; call rt_render_end first");
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	const int k = (int)(ctx->pipe_begun & 1);
	if (!ctx->pipe_counters) {
		RT_CUDA(ctx, cudaMallocHost((void**)&ctx->pipe_counters, 2 * 9 * sizeof(unsigned long long)));
		for (int i = 0; i < 2; i++) {
			RT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pipe_kernels[i], cudaEventDisableTiming));
			RT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pipe_done[i], cudaEventDisableTiming));
		}
	}
	const size_t npx = (size_t)cam->width * cam->height;
	RT_CUDA(ctx, ctx->pipe_rgb[k].alloc(npx * 3));
	if (first_ids) RT_CUDA(ctx, ctx->pipe_ids[k].alloc(npx));
	float* rgb_dev = ctx->pipe_rgb[k].p;
	int* ids_dev = first_ids ? ctx->pipe_ids[k].p : nullptr;
	if (prm->frame_first > 0)
		RT_CUDA(ctx, cudaMemcpyAsync(rgb_dev, rgb, npx * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
	const size_t W = cam->width;
	const BandHook copy_band = [&](int band, int y_begin, int y_end) -> rt_status {
		RT_CUDA(ctx, cudaEventRecord(ctx->band_done[band], ctx->stream));
		RT_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->band_done[band], 0));
		const size_t off = (size_t)y_begin * W, n = (size_t)(y_end - y_begin) * W;
		RT_CUDA(ctx, cudaMemcpyAsync(rgb + off * 3, rgb_dev + off * 3, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_stream));
		if (first_ids)
			RT_CUDA(ctx, cudaMemcpyAsync(first_ids + off, ids_dev + off, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->copy_stream));
		return RT_OK;
	};
	if (rt_status st = launch_render(ctx, cam, prm, flags, rgb_dev, ids_dev, 0, 1, false, host_bands(ctx, prm), copy_band)) return st;
	// this frame's counters and error flags, before the next frame's setup kernel zeroes the cells
	RT_CUDA(ctx, cudaMemcpyAsync(ctx->pipe_counters + 9 * k, ctx->counters.p, 9 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
	RT_CUDA(ctx, cudaEventRecord(ctx->pipe_kernels[k], ctx->stream));
	RT_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->pipe_kernels[k], 0));
	RT_CUDA(ctx, cudaEventRecord(ctx->pipe_done[k], ctx->copy_stream));
	ctx->pipe_begun++;
	return RT_OK;
}

rt_status rt_render_end(rt_ctx* ctx, rt_counters* counters) {
	if (!ctx) return RT_ERR_INVALID;
	if (ctx->pipe_begun == ctx->pipe_ended) return fail(ctx, RT_ERR_INVALID, "rt_render_end: no frame in flight");
	RT_CUDA(ctx, cudaSetDevice(ctx->device));
	const int k = (int)(ctx->pipe_ended & 1);
	RT_CUDA(ctx, cudaEventSynchronize(ctx->pipe_done[k]));
	ctx->pipe_ended++;
	const unsigned long long* h = ctx->pipe_counters + 9 * k;
	const uint32_t ef = (uint32_t)h[8];
	if (counters) {
		counters->paths = h[0]; counters->segments = h[1]; counters->nodes = h[2]; counters->tests = h[3]; counters->shades = h[4];
		counters->confirms = h[5];
		counters->texture_errors = (ef & RT_ERRFLAG_TEXTURE) ? 1 : 0;
		counters->acute_warnings = (ef & RT_ERRFLAG_ACUTE) ? 1 : 0;
	}
	if (ef & RT_ERRFLAG_STACK)
		return fail(ctx, RT_ERR_UNSUPPORTED, "a traversal stack was too small for this scene (internal limit): the frame was not rendered correctly");
	if (ef & RT_ERRFLAG_TEXTURE) return fail(ctx, RT_ERR_TEXTURE, "Texture coordinates out of bounds");
	return RT_OK;
}

}  // extern "C"
