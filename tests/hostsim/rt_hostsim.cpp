// tests/hostsim/rt_hostsim.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE, NOT A FALLBACK.
//
// Compiles the kernel body of librt_b200 (raytracer.js_b200/csrc/rt_trace.cuh, all RT_HD) for the
// host so that its traversal / confirmation / shading logic can be checked against the oracle on a
// box without a GPU (`pytest -m "not gpu"`).  It is built only by tests/ into tests/hostsim/, is never
// linked into or loaded by the package, and the C ABI in include/rt_b200.h has no path to it.
#include <thread>

#ifndef HOSTSIM_PPL
#define HOSTSIM_PPL 4  // rays per lane of the packet stage, as RT_PPL in rt_b200.cu
#endif

#include "../../raytracer.js_b200/csrc/rt_host.h"
#include "../../raytracer.js_b200/csrc/rt_trace.cuh"

// Camera.get_dir_for_each_pixel through the kernel body's ray generation (raygen_half_row): dirs[h][w][3].
extern "C" void hostsim_raygen(const rt_camera* cam, double* dirs) {
	RtFrame F{};
	std::vector<RtD4> row;
	rt_build_camera_rows(*cam, row, F.scan_cos, F.scan_sin);
	for (int i = 0; i < 3; i++) F.lf[i] = cam->lf[i];
	F.width = (int)cam->width;
	F.height = (int)cam->height;
	F.row_fr = row.data();
	F.ray_ckh = raygen_checkpoints_per_half(F.width);
	std::vector<double> ck((size_t)F.height * 2 * F.ray_ckh * 6);
	for (int t = 0; t < 6 * F.height; t++) raygen_half_row_component(F, t / 6, (t / 3) & 1, t % 3, ck.data());
	F.ray_ck = ck.data();
	for (int y = 0; y < F.height; y++)
		for (int x = 0; x < F.width; x++) pixel_dir(F, x, y, dirs + ((size_t)y * F.width + x) * 3);
}

// Entity moves (rt_host.h: rt_scene_move_entities, what rt_scene_update runs) applied to the description before the
// next hostsim_render packs it; n == 0 clears them.
static std::vector<uint32_t> g_move_ids;
static std::vector<double> g_move_pos;
static uint32_t g_move_depth = 16;
extern "C" void hostsim_set_moves(uint32_t n, const uint32_t* ids, const double* pos, uint32_t max_in_depth) {
	g_move_ids.assign(ids, ids + n);
	g_move_pos.assign(pos, pos + 3 * (size_t)n);
	g_move_depth = max_in_depth;
}

// Debugging aid: restrict the ray-by-ray mode (pipeline == 0) to a crop of the frame (w == 0: whole frame).
static int g_crop[4] = {0, 0, 0, 0};
// A digest (FNV-1a, 64 bit) of everything rt_pack_scene produces, for the test that the packed scene does not depend
// on the number of packing threads.  Returns the status; the message goes to errbuf.
extern "C" int hostsim_pack_digest(const rt_scene_desc* sc, uint64_t* digest, char* errbuf, int errlen) {
	RtHostScene hs;
	std::string err;
	const rt_status st = rt_pack_scene(sc, hs, err);
	if (errbuf && errlen > 0) snprintf(errbuf, errlen, "%s", err.c_str());
	if (st) return st;
	uint64_t h = 1469598103934665603ull;
	auto eat = [&](const void* p, size_t n) {
		const unsigned char* b = (const unsigned char*)p;
		for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
	};
#define EAT(v) eat(hs.v.data(), hs.v.size() * sizeof(hs.v[0]));
	EAT(node_geom) EAT(node_geom64) EAT(node_link) EAT(node_child) EAT(node_pk) EAT(node_walk) EAT(node_bvh) EAT(bvh_nodes) EAT(bvh_slots) EAT(bvh_geom)
	EAT(slot_geom) EAT(slot_geom64) EAT(slot_attr) EAT(slot_node) EAT(materials) EAT(textures) EAT(substances) EAT(texels)
#undef EAT
	eat(&hs.max_bvh_depth, sizeof hs.max_bvh_depth);
	eat(&hs.max_depth, sizeof hs.max_depth);
	eat(&hs.err_l, sizeof hs.err_l);
	*digest = h;
	return RT_OK;
}

#ifdef RT_WALK_STATS
unsigned long long g_walk_stats[8];
extern "C" void hostsim_walk_stats(unsigned long long* out, int reset) {
	for (int k = 0; k < 8; k++) {
		out[k] = g_walk_stats[k];
		if (reset) g_walk_stats[k] = 0;
	}
}
#endif

extern "C" void hostsim_set_crop(int x, int y, int w, int h) {
	g_crop[0] = x; g_crop[1] = y; g_crop[2] = w; g_crop[3] = h;
}

// tile_world <= 1: full frame into rgb[height][width][3].  tile_world > 1: only the tiles of tile_rank,
// tile-major into rgb[k][16*16][3] (the layout of rt_render_tiles_device).
// pipeline != 0: the pipeline of rt_b200.cu (primary_patch per packet with the 32 lanes as loop iterations: packet
// walk + shading of the paths that end at their first hit; then the bounce stage over the continuation queue); no
// work counters in that mode.
// pipeline == 0: every path ray by ray with the counting variant (what rt_render_kernel<true> runs).
extern "C" int hostsim_render(const rt_scene_desc* sc, const rt_camera* cam, const rt_params* prm, int n_threads,
                              int tile_rank, int tile_world, int pipeline, float* rgb, int32_t* ids,
                              rt_counters* counters, char* errbuf, int errlen) {
	std::string err;
	RtHostScene hs;
	RtSceneCopy moved;
	rt_scene_desc moved_desc;
	if (!g_move_ids.empty()) {
		moved.assign(*sc);
		if (!rt_scene_move_entities(moved, (uint32_t)g_move_ids.size(), g_move_ids.data(), g_move_pos.data(), g_move_depth, err)) {
			snprintf(errbuf, errlen, "%s", err.c_str());
			return (int)RT_ERR_UNSUPPORTED;
		}
		moved_desc = moved.desc();
		sc = &moved_desc;
	}
	rt_status st = rt_pack_scene(sc, hs, err);
	if (!st) st = rt_check_render_args(true, (uint32_t)hs.textures.size(), (uint32_t)hs.substances.size(), cam, prm, err);
	RtFrame F{};
	if (!st) st = rt_fill_frame(hs, cam, prm, F, err);
	if (st) {
		snprintf(errbuf, errlen, "%s", err.c_str());
		return (int)st;
	}
	std::vector<RtD4> row;
	rt_build_camera_rows(*cam, row, F.scan_cos, F.scan_sin);
	RtDevScene S{};
	S.node_geom = hs.node_geom.data(); S.node_geom64 = hs.node_geom64.data(); S.node_link = hs.node_link.data(); S.node_child = hs.node_child.data(); S.node_pk = hs.node_pk.data(); S.node_walk = hs.node_walk.data(); S.node_bvh = hs.node_bvh.data(); S.bvh_nodes = hs.bvh_nodes.data(); S.bvh_slots = hs.bvh_slots.data(); S.bvh_geom = hs.bvh_geom.data();
	S.slot_geom = hs.slot_geom.data(); S.slot_geom64 = hs.slot_geom64.data(); S.slot_attr = hs.slot_attr.data(); S.slot_node = hs.slot_node.data();
	S.materials = hs.materials.data(); S.textures = hs.textures.data(); S.substances = hs.substances.data();
	S.texels = hs.texels.data();
	for (int k = 0; k < 3; k++) S.root_pos[k] = hs.root_pos[k];
	S.root_size = hs.root_size;
	S.n_nodes = (int)hs.node_geom.size();
	S.n_slots = (int)hs.slot_geom.size();
	S.err_l = hs.err_l;
	S.ordered_ok = rt_ordered_walk_fits(hs);
	F.row_fr = row.data();
	// ray generation (the body of the frame-setup kernel's ray-generation lanes): checkpoints of the generator's
	// iterated rotations along every row
	F.ray_ckh = raygen_checkpoints_per_half(F.width);
	std::vector<double> ray_ck((size_t)F.height * 2 * F.ray_ckh * 6);
	for (int t = 0; t < 6 * F.height; t++) raygen_half_row_component(F, t / 6, (t / 3) & 1, t % 3, ray_ck.data());
	F.ray_ck = ray_ck.data();
	F.rgb = rgb;
	F.first_ids = ids;
	// bit 1 of `pipeline`: a shard (tile_world > 1) writes into a FRAME-layout buffer (rt_render_shard_device)
	const bool tiled = tile_world > 1 && !(pipeline & 2);
	pipeline &= 1;
	F.tile_rank = tile_world > 1 ? tile_rank : 0;
	F.tile_world = tile_world > 1 ? tile_world : 1;
	F.tile_compact = tiled ? 1 : 0;
	// the primary-ray preparation of launch_render (rt_b200.cu), on the host
	std::vector<RtF4> prim(hs.slot_geom.size());
	for (size_t s = 0; s < prim.size(); s++)
		prim[s] = make_prim_record(hs.slot_geom64[s], hs.slot_geom[s].w > 0.0f, cam->pos[0], cam->pos[1], cam->pos[2], hs.err_l);
	F.prim_geom = prim.empty() || (prm->flags & RT_PARAM_NO_PRIMARY_RECORDS) ? nullptr : prim.data();
	F.packet_ok = 0;
	if (F.prim_geom) rt_fill_chain(hs, F);
	if (n_threads < 1) n_threads = 1;
	if (pipeline && F.packet_ok && !(prm->flags & RT_PARAM_PER_RAY)) {
		// ---- primary stage, patch by patch in the kernel's own patch geometry
		const int tiles_x = (F.width + 15) / 16, tiles_y = (F.height + 15) / 16, n_tiles = tiles_x * tiles_y;
		const int my_tiles = (n_tiles - F.tile_rank + F.tile_world - 1) / F.tile_world;
		std::vector<uint32_t> errs(n_threads, 0);
		auto out_index_of = [&](int x, int y, int k) {
			return F.tile_compact ? (size_t)k * 256 + ((y & 15) * 16 + (x & 15)) : (size_t)y * F.width + x;
		};
		constexpr int PPL = HOSTSIM_PPL, PER_TILE = 8 / PPL;
		// the continuation queue of the primary stage (rt_trace.cuh: primary_patch appends with an atomic)
		std::vector<RtQueueItem> queue(tiled ? (size_t)my_tiles * 256 : (size_t)F.width * F.height);
		unsigned queue_count = 0;
		F.queue = queue.data();
		F.queue_count = &queue_count;
		// ---- primary stage: packet walk + shading of the paths that end at their first hit
		auto stage_a = [&](int t) {
			std::vector<RtPNode> stack(RT_PACKET_STACK(HOSTSIM_PPL));
			std::vector<RtPRay> rays(PPL * 32);
			std::vector<double> dirs(PPL * 32 * 3);
			float stage[96];
			for (int p = t; p < my_tiles * PER_TILE; p += n_threads) {
				const int k = p / PER_TILE;
				const int tile = F.tile_rank + k * F.tile_world;
				if (tile >= n_tiles) continue;
				RtPatch pt;
				pt.x0 = (tile % tiles_x) * 16;
				pt.y0 = (tile / tiles_x) * 16;
				pt.sub0 = (p % PER_TILE) * PPL;
				pt.out_base = (size_t)k * 256;
				primary_patch<PPL>(S, F, pt, stack.data(), rays.data(), dirs.data(), stage, errs[t]);
			}
		};
		std::vector<std::thread> th;
		for (int t = 0; t < n_threads; t++) th.emplace_back(stage_a, t);
		for (auto& t : th) t.join();
		th.clear();
		// ---- bounce stage over the continuation queue
		const uint64_t queued = queue_count;
		auto stage_b = [&](int t) {
			for (size_t i = t; i < queued; i += n_threads) {
				const RtQueueItem& it = queue[i];
				const int x = (int)(it.xy & 0xffffu), y = (int)(it.xy >> 16);
				const int tile = (y / 16) * tiles_x + (x / 16);
				RtCounts c = {0, 0, 0, 0, 0};
				render_pixel<false>(S, F, x, y, out_index_of(x, y, tile / F.tile_world), c, errs[t], it.slot);
			}
		};
		for (int t = 0; t < n_threads; t++) th.emplace_back(stage_b, t);
		for (auto& t : th) t.join();
		if (counters) {
			memset(counters, 0, sizeof *counters);
			counters->paths = (uint64_t)F.width * F.height * F.n_frames;
			counters->confirms = queued;  // (test hook) number of pixels that went through the queue
			for (int t = 0; t < n_threads; t++) {
				if (errs[t] & RT_ERRFLAG_TEXTURE) counters->texture_errors = 1;
				if (errs[t] & RT_ERRFLAG_ACUTE) counters->acute_warnings = 1;
			}
		}
		for (int t = 0; t < n_threads; t++)
			if (errs[t] & RT_ERRFLAG_STACK) {
				snprintf(errbuf, errlen, "a traversal stack was too small for this scene");
				return (int)RT_ERR_UNSUPPORTED;
			}
		for (int t = 0; t < n_threads; t++)
			if (errs[t] & RT_ERRFLAG_STACK) {
				snprintf(errbuf, errlen, "a traversal stack was too small for this scene");
				return (int)RT_ERR_UNSUPPORTED;
			}
		return 0;
	}
	F.packet_ok = 0;
	std::vector<RtCounts> part(n_threads, RtCounts{0, 0, 0, 0, 0});
	std::vector<uint32_t> errs(n_threads, 0);
	auto work = [&](int t) {
		for (int y = t; y < F.height; y += n_threads)
			for (int x = 0; x < F.width; x++) {
				RtCounts c = {0, 0, 0, 0, 0};
				size_t out_index = (size_t)y * F.width + x;
				if (g_crop[2] > 0 && (x < g_crop[0] || x >= g_crop[0] + g_crop[2] || y < g_crop[1] || y >= g_crop[1] + g_crop[3])) continue;
				if (F.tile_world > 1) {  // same mapping as rt_render_kernel
					const int tiles_x = (F.width + 15) / 16;
					const int tile = (y / 16) * tiles_x + (x / 16);
					if (tile % tile_world != tile_rank) continue;
					if (tiled) out_index = (size_t)(tile / tile_world) * 256 + (y % 16) * 16 + (x % 16);
				}
				render_pixel<true>(S, F, x, y, out_index, c, errs[t]);
				part[t].segments += c.segments; part[t].nodes += c.nodes; part[t].tests += c.tests;
				part[t].shades += c.shades; part[t].confirms += c.confirms;
			}
	};
	std::vector<std::thread> th;
	for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
	for (auto& t : th) t.join();
	if (counters) {
		memset(counters, 0, sizeof *counters);
		counters->paths = (uint64_t)F.width * F.height * F.n_frames;
		for (int t = 0; t < n_threads; t++) {
			counters->segments += part[t].segments; counters->nodes += part[t].nodes; counters->tests += part[t].tests;
			counters->shades += part[t].shades; counters->confirms += part[t].confirms;
			if (errs[t] & RT_ERRFLAG_TEXTURE) counters->texture_errors = 1;
			if (errs[t] & RT_ERRFLAG_ACUTE) counters->acute_warnings = 1;
		}
	}
	for (int t = 0; t < n_threads; t++)
		if ((errs[t] | part[t].errors) & RT_ERRFLAG_STACK) {
			snprintf(errbuf, errlen, "a traversal stack was too small for this scene");
			return (int)RT_ERR_UNSUPPORTED;
		}
	return 0;
}
