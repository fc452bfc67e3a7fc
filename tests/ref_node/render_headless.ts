// Headless render of the REFERENCE (Dark565/raytracer.js) in Node.  Never executed in this repo's environments
// (no JavaScript runtime): see README.md.  Bundle with esbuild, aliasing @app -> <reference>/src.
import * as fs from 'fs';
import FpLcg from '@app/math/rng/fp-lcg';
import { point } from '@app/math/geometry';
import { Camera, CameraConfig } from '@app/view/camera';
import ExposureBuffer from '@app/view/exposure_buffer';
import { Raytracer, RaytracerConfig, Ray } from '@app/raytracer';
import { new_entity_octree, add_entity_to_octree } from '@app/octree_entity';
import { SphereEntity } from '@app/entities/entity_sphere';
import { BoxEntity } from '@app/entities/entity_box';
import { SolidMaterial } from '@app/materials/material_solid';
import { ResponseType } from '@app/material';
import { SolidTexture } from '@app/texture/texture_solid';
import { Texture } from '@app/texture/texture';
import { SkySphere } from '@app/sky/sky_sphere';
import { SUBSTANCE_AIR } from '@app/substance';
import { Entity } from '@app/entity';
import { Point } from '@app/math/geometry';

function arg(name: string, def?: string): string {
	const i = process.argv.indexOf('--' + name);
	if (i < 0) { if (def === undefined) throw Error(`missing --${name}`); return def; }
	return process.argv[i + 1];
}
const flag = (name: string) => process.argv.includes('--' + name);

/** SolidMaterial that notes, per Ray, the first entity alter_ray() is called with: the first collision of the
 *  path (src/raytracer.ts:206).  Rays are distinguished by object identity. */
const first_hit = new WeakMap<Ray, Entity>();
class RecordingMaterial extends SolidMaterial {
	alter_ray(ray: Ray, entity: Entity, texture: Texture, p: Point): boolean {
		if (!first_hit.has(ray)) first_hit.set(ray, entity);
		return super.alter_ray(ray, entity, texture, p);
	}
}

const n = parseInt(arg('n')), dmin = parseFloat(arg('dmin')), dmax = parseFloat(arg('dmax'));
const seed = parseFloat(arg('seed', '42')), mix = arg('mix', 'diffuse'), box_fraction = parseFloat(arg('boxes', '0'));
const size = parseInt(arg('size')), frames = parseInt(arg('frames', '1')), rng_seed = parseFloat(arg('rng-seed', '1'));

// ---- the scene of raytracer.js_b200/scenes.py:random_spheres (same draws, same order)
const rng = new FpLcg(seed);
const otree = new_entity_octree({ pos: point(0, 0, 0), size: 1 }, undefined);
const mats = mix == 'diffuse'
	? [new RecordingMaterial(ResponseType.REFLECTION, false, false, 0.0)]
	: [new RecordingMaterial(ResponseType.REFLECTION, false, true, 0.0), new RecordingMaterial(ResponseType.REFLECTION, false, false, 0.0),
	   new RecordingMaterial(ResponseType.REFLECTION, false, true, 0.5), new RecordingMaterial(ResponseType.REFLECTION, true, false, 0.0)];
const cum = mix == 'diffuse' ? [1.0] : [0.70, 0.85, 0.95, 1.0];
const entities: Entity[] = [];
for (let i = 0; i < n; ++i) {
	const d = dmin + rng.next() * (dmax - dmin);
	const c = [0, 0, 0].map(() => d / 2 + rng.next() * (1 - d));
	const is_box = box_fraction > 0 && rng.next() < box_fraction;
	let mi = 0;
	if (mats.length > 1) { const u = rng.next(); while (mi < cum.length - 1 && u > cum[mi]) mi++; }
	const k = mats[mi].light_source ? 5.0 : 1.0;
	const tex = new SolidTexture({ r: rng.next() * k, g: rng.next() * k, b: rng.next() * k, a: 1.0 });
	const e = is_box ? new BoxEntity(undefined, mats[mi], tex, SUBSTANCE_AIR, point(c[0], c[1], c[2]), d)
	                 : new SphereEntity(undefined, mats[mi], tex, SUBSTANCE_AIR, point(c[0], c[1], c[2]), d);
	add_entity_to_octree(otree, e, { max_in_depth: 16, max_out_depth: 0 });
	entities.push(e);
}
const index_of = new Map<Entity, number>(entities.map((e, i) => [e, i]));

// ---- camera (the pose of scenes.bench_camera), exposure buffer, raytracer
const conf: CameraConfig = { fov_v: Math.PI * 0.5, fov_h: Math.PI * 0.5, screen_w: size, screen_h: size,
                             rot_v: Math.PI / 30, rot_h: Math.PI / 30, flags: { vertical_locked: true } };
const camera = new Camera(conf, point(0.5013, 0.4987, 0.5021), 0, Math.PI / 180 * 30);
class DumpBuffer extends ExposureBuffer { get store(): Float32Array { return this.pixels; } }
const ebuffer = new DumpBuffer(size, size, -1);
const prng = new FpLcg(rng_seed);
const rconf: RaytracerConfig = { refmax: mix == 'diffuse' ? 1 : 4, default_substance: SUBSTANCE_AIR, distance_attenuation_factor: 1,
                                 sky: new SkySphere(new SolidTexture({ r: 0.2, g: 0.2, b: 0.7, a: 1.0 })) };
const raytracer = new Raytracer(rconf, otree, camera, ebuffer, prng);

// ---- first-hit ids: trace_frame() creates one Ray per pixel in the generator's order; hook Ray.trace to learn which
const ids = new Int32Array(size * size).fill(-1);
let frame = 0, cur_px = -1;
const order: [number, number][] = [];
for (const px of camera.get_dir_for_each_pixel()) order.push([px.x, px.y]);
let ray_no = 0;
const orig_trace = (Ray.prototype as any).trace;
(Ray.prototype as any).trace = function () {
	const [x, y] = order[ray_no++ % order.length];
	cur_px = y * size + x;
	if (!flag('shared-rng')) prng.seed(rng_seed + cur_px + frame * size * size);  // the harness RNG policy (rt_b200.h)
	const r = orig_trace.call(this);
	const e = first_hit.get(this);
	ids[cur_px] = e !== undefined ? index_of.get(e)! : -1;
	return r;
};
for (frame = 0; frame < frames; ++frame) {
	if (frame) ebuffer.next_frame();
	raytracer.trace_frame();
}

const head = new Int32Array([size, size, frames]);
fs.writeFileSync(arg('out'), Buffer.concat([Buffer.from(head.buffer), Buffer.from(ebuffer.store.buffer), Buffer.from(ids.buffer)]));
console.log(`wrote ${arg('out')}: ${size}x${size}, ${frames} frame(s), ${entities.length} entities`);
