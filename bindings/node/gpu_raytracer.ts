/* gpu_raytracer.ts — `GpuRaytracer`, a drop-in for `Raytracer` (src/raytracer.ts:281-339) that runs
 * trace_frame() on a B200 through the N-API addon rt_b200.node (rt_napi.c -> librt_b200.so).
 *
 * Same constructor arguments and public members as the reference class: config, set_camera,
 * set_ebuffer, trace_frame, tree, rng.  In main.ts (src/main.ts:408) the only change is
 *     const raytracer = new GpuRaytracer(conf, otree, camera, ebuffer, prng);
 * Unknown Entity / Material / Texture / Sky subclasses throw Error("unsupported ..."): there is no
 * CPU fallback.  (Written against the reference's sources; it cannot be compiled or run in the build
 * image, which has no Node / tsc — see INTEGRATION.md.) */
import { RaytracerConfig } from '@app/raytracer';
import { EntityOtree } from '@app/octree_entity';
import { Camera } from '@app/view/camera';
import ExposureBuffer from '@app/view/exposure_buffer';
import RNG from '@app/math/rng/rng';
import { Entity } from '@app/entity';
import { SphereEntity } from '@app/entities/entity_sphere';
import { BoxEntity } from '@app/entities/entity_box';
import { StaticMaterial } from '@app/material';
import { SolidTexture } from '@app/texture/texture_solid';
import { ImageTexture } from '@app/texture/texture_image';
import { SkySphere } from '@app/sky/sky_sphere';
import Substance from '@app/substance';

// eslint-disable-next-line @typescript-eslint/no-var-requires
const native = require('./rt_b200.node');

/** ExposureBuffer whose pixel store is handed to the GPU (protected members are reachable from a subclass). */
export class GpuExposureBuffer extends ExposureBuffer {
	get store(): Float32Array { return this.pixels; }
	get w(): number { return this.width; }
	get h(): number { return this.height; }
}

interface FlatScene { [name: string]: ArrayBufferView }

/** DFS pre-order walk of the pointer octree (children 0..7) into the arrays of rt_scene_desc
 *  (include/rt_b200.h).  Each node's list keeps the EntitySet's insertion order (F1). */
export function flatten(root: EntityOtree, extra_textures: any[], extra_substances: Substance[]): { flat: FlatScene, tex: Map<any, number>, sub: Map<any, number> } {
	if (root.parent !== undefined) throw Error('unsupported octree: the tree handed to the raytracer must be the absolute root');
	const nodes: EntityOtree[] = [], parent: number[] = [], octant: number[] = [];
	const stack: [EntityOtree, number, number][] = [[root, -1, -1]];
	while (stack.length) {
		const [n, p, o] = stack.pop()!;
		const idx = nodes.length;
		nodes.push(n); parent.push(p); octant.push(o);
		for (let c = 7; c >= 0; --c) { const ch = n.get(c); if (ch !== undefined) stack.push([ch, idx, c]); }
	}
	// the pop order above is pre-order with children ascending; child indices are resolved afterwards
	const N = nodes.length, index = new Map<EntityOtree, number>(nodes.map((n, i) => [n, i]));
	const mats = new Map<any, number>(), tex = new Map<any, number>(), sub = new Map<any, number>();
	const intern = (m: Map<any, number>, k: any) => { if (!m.has(k)) m.set(k, m.size); return m.get(k)!; };
	extra_textures.forEach(t => intern(tex, t));
	extra_substances.forEach(s => intern(sub, s));
	const ents: Entity[] = [], list_off = new Uint32Array(N + 1);
	const node_pos = new Float64Array(3 * N), node_size = new Float64Array(N), node_child = new Int32Array(8 * N).fill(-1);
	nodes.forEach((n, i) => {
		node_pos.set(n.id.pos.v, 3 * i); node_size[i] = n.id.size;
		for (let c = 0; c < 8; ++c) { const ch = n.get(c); if (ch !== undefined) node_child[8 * i + c] = index.get(ch)!; }
		list_off[i] = ents.length;
		if (n.value) for (const e of n.value.set) ents.push(e);
	});
	list_off[N] = ents.length;
	const E = ents.length;
	const f: FlatScene = {
		node_pos, node_size, node_child, node_parent: Int32Array.from(parent), node_octant: Int32Array.from(octant),
		node_list_off: list_off, list_entity: Uint32Array.from(ents.keys()),
		ent_type: new Uint8Array(E), ent_pos: new Float64Array(3 * E), ent_extent: new Float64Array(E),
		ent_material: new Int32Array(E), ent_texture: new Int32Array(E), ent_substance: new Int32Array(E),
	};
	ents.forEach((e, i) => {
		if (e instanceof SphereEntity) { (f.ent_type as Uint8Array)[i] = 0; (f.ent_extent as Float64Array)[i] = e.get_diameter(); }
		else if (e instanceof BoxEntity) { (f.ent_type as Uint8Array)[i] = 1; (f.ent_extent as Float64Array)[i] = e.get_size(); }
		else throw Error(`unsupported Entity subclass ${e.constructor.name}`);
		(f.ent_pos as Float64Array).set((e as any).get_pos().v, 3 * i);
		const m = (e as any).get_material();
		if (!(m instanceof StaticMaterial)) throw Error(`unsupported Material subclass ${m.constructor.name}`);
		(f.ent_material as Int32Array)[i] = intern(mats, m);
		(f.ent_texture as Int32Array)[i] = intern(tex, (e as any).get_texture());
		const s = e.get_substance();
		(f.ent_substance as Int32Array)[i] = s === undefined ? -1 : intern(sub, s);
	});
	const M = mats.size, T = tex.size;
	f.mat_response = new Uint8Array(M); f.mat_light = new Uint8Array(M); f.mat_mirror = new Uint8Array(M); f.mat_roughness = new Float64Array(M);
	mats.forEach((i, m: StaticMaterial) => {
		(f.mat_response as Uint8Array)[i] = m.response; (f.mat_light as Uint8Array)[i] = +m.light_source;
		(f.mat_mirror as Uint8Array)[i] = +m.mirror; (f.mat_roughness as Float64Array)[i] = m.roughness_index;
	});
	f.tex_kind = new Uint8Array(T); f.tex_color = new Float64Array(4 * T); f.tex_width = new Int32Array(T); f.tex_height = new Int32Array(T);
	f.tex_loaded = new Uint8Array(T); f.tex_texel_off = new BigUint64Array(T);
	const pool: number[][] = []; let texels = 0;
	tex.forEach((i, t: any) => {
		if (t instanceof SolidTexture) { const c = (t as any).color; (f.tex_color as Float64Array).set([c.r, c.g, c.b, c.a], 4 * i); }
		else if (t instanceof ImageTexture) {
			const c = (t as any).fallback_color, data: number[] | undefined = (t as any).image_data; // TS-private, plain at run time
			(f.tex_kind as Uint8Array)[i] = 1; (f.tex_color as Float64Array).set([c.r, c.g, c.b, c.a], 4 * i);
			if (data && data.length) {
				(f.tex_loaded as Uint8Array)[i] = 1; (f.tex_width as Int32Array)[i] = (t as any).width; (f.tex_height as Int32Array)[i] = (t as any).height;
				(f.tex_texel_off as BigUint64Array)[i] = BigInt(texels); pool.push(data); texels += data.length / 3;
			}
		} else throw Error(`unsupported Texture subclass ${t.constructor.name}`);
	});
	// ImageTexture.image_data holds the decoded channels as doubles in [0,1] (`image_data[index] / 255.0`,
	// src/texture/texture_image.ts:117-119): the texel pool wants the 8-bit values back (the device divides by 255.0
	// again on use, the reference's own expression).  Math.round(v * 255) is exact: the values are k / 255.
	const tx = new Uint8Array(texels * 3); let o = 0;
	for (const d of pool) for (const v of d) tx[o++] = Math.round(v * 255);
	f.texels = tx;
	f.sub_refractive_index = new Float64Array(sub.size);
	sub.forEach((i, s: Substance) => { (f.sub_refractive_index as Float64Array)[i] = s.refractive_index; });
	return { flat: f, tex, sub };
}

/** `new ImageTexture(url, fallback, hflip, vflip)` for a host without a DOM (SURVEY.md 8f N3).  The reference's
 *  constructor decodes through `new Image()` + a canvas (src/texture/texture_image.ts:76-136) and cannot run under
 *  Node; this builds the same object - same prototype, same fields, so `get_color` and the flattener above see no
 *  difference - from the library's decoder (native.decodeImage: PNG, BMP, binary PPM; RGB kept, alpha dropped), with
 *  the flips done by the reference's own index walk.  A file that does not decode leaves `image_data` undefined: the
 *  texture answers its fallback colour, as while the reference's loading promise is pending or rejected. */
export function image_texture_from_file(path: string, fallback_color: any, hflip = false, vflip = false): ImageTexture {
	const t: any = Object.create(ImageTexture.prototype);
	t.image_url = path; t.image_data = undefined; t.fallback_color = fallback_color; t.loading_promise = Promise.resolve();
	try {
		const { width, height, rgb } = native.decodeImage(new Uint8Array(require('fs').readFileSync(path)));
		const data: number[] = Array(width * height * 3);
		let j = 0;
		for (let y = vflip ? height - 1 : 0; y !== (vflip ? -1 : height); y += vflip ? -1 : 1)
			for (let x = hflip ? width - 1 : 0; x !== (hflip ? -1 : width); x += hflip ? -1 : 1) {
				const k = (y * width + x) * 3;
				data[j] = rgb[k] / 255.0; data[j + 1] = rgb[k + 1] / 255.0; data[j + 2] = rgb[k + 2] / 255.0;
				j += 3;
			}
		t.image_data = data; t.width = width; t.height = height;
	} catch (e) { /* not decodable: fallback colour */ }
	return t as ImageTexture;
}

export class GpuRaytracer {
	private otree: EntityOtree;
	private camera: Camera;
	private ebuffer: GpuExposureBuffer;
	private _rng: RNG;
	private ctx: any;
	private tex!: Map<any, number>;
	private sub!: Map<any, number>;
	private pinned?: Float32Array;
	private n_gpus = 1;
	/** seed handed to the per-pixel reseed policy (rt_b200.h: rt_params.rng_seed) */
	rng_seed = 1.0;
	/** 0: float32 search + float64 confirmation (default); 1: RT_PRECISION_F64, the walker in float64, ray by ray -
	 *  for octrees deeper than float32 resolves (max_in_depth beyond ~23 below a unit root) */
	precision = 0;
	/** RT_PARAM_EXACT_TIES: for scenes built on a dyadic lattice (entities that fill or touch their cells exactly); on by
	 *  itself whenever the camera stands on a cell plane of the octree */
	exact_ties = false;

	config: RaytracerConfig;

	/** n_gpus > 1: ONE process drives that many GPUs behind the same calls (rt_create_multi): the scene is packed once
	 *  and replicated device to device, trace_frame() shards the frame into interleaved 16x16 tiles and every GPU
	 *  stores its tiles straight into the ExposureBuffer's Float32Array (page-locked and mapped by the library). */
	constructor(config: RaytracerConfig, otree: EntityOtree, camera: Camera, ebuffer: GpuExposureBuffer, rng: RNG, device = -1, n_gpus = 1) {
		this.camera = camera;
		this.ebuffer = ebuffer;
		this.otree = otree;
		this._rng = rng;
		this.config = Object.assign({}, config);
		this.n_gpus = n_gpus;
		this.ctx = n_gpus > 1 ? native.createMulti(n_gpus) : native.create(device);  // throws without a CUDA device: no CPU fallback
		this.refresh_scene();
	}

	/** Re-flatten and upload the octree (call after entities were added or moved). */
	refresh_scene() {
		if (!(this.config.sky instanceof SkySphere)) throw Error(`unsupported Sky subclass ${this.config.sky.constructor.name}`);
		const { flat, tex, sub } = flatten(this.otree, [(this.config.sky as any).texture], [this.config.default_substance]);
		this.tex = tex; this.sub = sub;
		native.uploadScene(this.ctx, flat);
	}

	set_camera(camera: Camera) { this.camera = camera; }

	set_ebuffer(ebuffer: GpuExposureBuffer) { this.ebuffer = ebuffer; }

	/** One Raytracer.trace_frame() (src/raytracer.ts:308-330) into the ExposureBuffer's pixel store. */
	trace_frame(n_frames = 1) {
		const cam: any = this.camera, conf = cam.conf;    // norm_fr/lf/up, pos, conf are TS-private, plain at run time
		const eb = this.ebuffer;
		// one GPU: page-lock the pixel store once so that the copy runs at PCIe rate; a group maps it itself
		if (this.n_gpus === 1 && this.pinned !== eb.store) { if (this.pinned) native.unpin(this.ctx, this.pinned); native.pin(this.ctx, eb.store); this.pinned = eb.store; }
		native.render(this.ctx,
			{ pos: Float64Array.from(cam.pos.v), fr: Float64Array.from(cam.norm_fr.v), lf: Float64Array.from(cam.norm_lf.v),
			  up: Float64Array.from(cam.norm_up.v), fov_h: conf.fov_h, fov_v: conf.fov_v, width: conf.screen_w, height: conf.screen_h,
			  flags: 1 /* RT_CAM_REFERENCE_EXTENTS: behave exactly like the reference, including its throw on non-square frames */ },
			{ refmax: this.config.refmax, sky_texture: this.tex.get((this.config.sky as any).texture),
			  default_substance: this.sub.get(this.config.default_substance),
			  distance_attenuation_factor: this.config.distance_attenuation_factor,
			  n_frames, frame_first: eb.current_frame, rng_seed: this.rng_seed, want_counters: 0, precision: this.precision, exact_ties: +this.exact_ties },
			eb.store);
		for (let i = 1; i < n_frames; ++i) eb.next_frame();
	}

	get tree() { return this.otree; }

	get rng() { return this._rng; }
}

/** View (src/view/view.ts:24-41) with draw_ebuffer() on the GPU: exposure statistics, the tone mapper's
 *  dynamic range and the range compression run in librt_b200 (rt_present), and the CanvasScreen's ImageData is
 *  filled in place.  Same constructor arguments as View plus the GpuRaytracer whose context does the work.
 *  Unknown ToneMapper subclasses throw (no CPU fallback). */
export class GpuView {
	ebuffer: GpuExposureBuffer;
	screen: any;          // CanvasScreen: `image` (ImageData) and `flush()` are reached with a cast, like the camera's fields
	tone_mapper: any;     // ToneMapper_Identity | ToneMapper_StdDevAroundMean | ToneMapper_AbsDevAroundMean
	last_stats?: { mean: number, variance: number, absolute_dev: number, drange_low: number, drange_high: number };
	private raytracer: GpuRaytracer;

	constructor(ebuffer: GpuExposureBuffer, screen: any, tone_mapper: any, raytracer: GpuRaytracer) {
		this.ebuffer = ebuffer; this.screen = screen; this.tone_mapper = tone_mapper; this.raytracer = raytracer;
	}

	private tone() {
		const tm = this.tone_mapper, name = tm.constructor.name;
		if (name === 'ToneMapper_Identity') return { kind: 0, dynamic_range: 0, min_dynamic: 0, max_dynamic: 0 };
		const kind = name === 'ToneMapper_StdDevAroundMean' ? 1 : name === 'ToneMapper_AbsDevAroundMean' ? 2 : -1;
		if (kind < 0) throw Error(`unsupported ToneMapper subclass ${name}`);
		return { kind, dynamic_range: Math.log2(tm.dynamic_coef), min_dynamic: tm.min_dynamic, max_dynamic: tm.max_dynamic };
	}

	/** View.draw_ebuffer() (src/view/view.ts:34-38) */
	draw_ebuffer() {
		const eb = this.ebuffer, image: ImageData = (this.screen as any).image;
		this.last_stats = native.present((this.raytracer as any).ctx, eb.store, eb.w, eb.h, this.tone(), image.data);
		if (!(this.screen as any).flags.buffer_pixels) this.screen.flush();
	}
}
