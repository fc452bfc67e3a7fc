/* rt_napi.c — Node N-API addon over the C ABI of librt_b200.so (include/rt_b200.h).
 *
 * This is the binding a maintainer of Dark565/raytracer.js adds on the reference side; the TypeScript
 * class that uses it (gpu_raytracer.ts) drops in for `Raytracer` (src/raytracer.ts:281-339).  The addon is
 * deliberately thin: every typed array is passed by pointer (zero copy: the Float32Array of the
 * ExposureBuffer is the render target, page-locked once with rt_host_register), every non-zero rt_status
 * becomes a thrown JS Error carrying rt_last_error(), and the process is never aborted.
 *
 * Build where Node exists:   cc -shared -fPIC -o rt_b200.node rt_napi.c -I$(NODE)/include/node -L. -lrt_b200
 * Build check here (no Node): gcc -fsyntax-only -Wall -I../../include rt_napi.c   (uses node_api_min.h)
 *
 * Exports:  create(device) -> ctx      (freed by the finalizer of the external)
 *           createMulti(nGpus) -> ctx  one process, nGpus GPUs behind the same calls (rt_create_multi)
 *           uploadScene(ctx, flat)     flat = the object gpu_raytracer.ts: flatten() builds (typed arrays)
 *           render(ctx, camera, params, pixels: Float32Array, ids?: Int32Array) -> counters | undefined
 *           pin(ctx, Float32Array)     unpin(ctx, Float32Array)
 */
#if defined(__has_include)
#if __has_include(<node_api.h>)
#include <node_api.h>
#else
#include "node_api_min.h"
#endif
#else
#include "node_api_min.h"
#endif
#include <stdio.h>
#include <string.h>

#include "rt_b200.h"

#define RT_NAPI_TRY(env, call)                                  \
	do {                                                        \
		if ((call) != napi_ok) {                                \
			napi_throw_error((env), NULL, "rt_b200: bad argument (" #call ")"); \
			return NULL;                                        \
		}                                                       \
	} while (0)

static napi_value throw_rt(napi_env env, rt_ctx* ctx, rt_status st) {
	const char* msg = rt_last_error(ctx);
	/* keep the reference's own messages: "x or y out of bounds", "Texture coordinates out of bounds" */
	napi_throw_error(env, st == RT_ERR_BOUNDS ? "ERR_OUT_OF_RANGE" : NULL, msg && *msg ? msg : "rt_b200 failed");
	return NULL;
}

static rt_ctx* ctx_of(napi_env env, napi_value v) {
	void* p = NULL;
	if (napi_get_value_external(env, v, &p) != napi_ok) return NULL;
	return (rt_ctx*)p;
}

/* typed array property `name` of `obj` -> data pointer (+ element count, + element type) */
static void* ta_typed(napi_env env, napi_value obj, const char* name, size_t* len, napi_typedarray_type* type) {
	napi_value v;
	bool is = false;
	void* data = NULL;
	size_t n = 0;
	napi_typedarray_type t;
	if (len) *len = 0;
	if (napi_get_named_property(env, obj, name, &v) != napi_ok || napi_is_typedarray(env, v, &is) != napi_ok || !is) return NULL;
	if (napi_get_typedarray_info(env, v, &t, &n, &data, NULL, NULL) != napi_ok) return NULL;
	if (len) *len = n;
	if (type) *type = t;
	return data;
}
static void* ta(napi_env env, napi_value obj, const char* name, size_t* len) { return ta_typed(env, obj, name, len, NULL); }
/* The flat scene's arrays are indexed by rt_pack_scene with the counts taken from a few of them: every array
 * must have the element type and EXACTLY the length those counts imply, or the library would read past its end. */
static void* ta_checked(napi_env env, napi_value obj, const char* name, napi_typedarray_type want, size_t want_len, int* bad) {
	size_t n = 0;
	napi_typedarray_type t = want;
	void* p = ta_typed(env, obj, name, &n, &t);
	if ((!p && want_len) || t != want || n != want_len) {
		if (!*bad) {
			char msg[160];
			snprintf(msg, sizeof msg, "rt_b200: flat.%s must be a typed array of kind %d with %zu elements (got kind %d, %zu)", name, (int)want,
			         want_len, (int)t, n);
			napi_throw_error(env, "ERR_OUT_OF_RANGE", msg);
		}
		*bad = 1;
		return NULL;
	}
	return p;
}
static double num(napi_env env, napi_value obj, const char* name) {
	napi_value v;
	double d = 0;
	if (napi_get_named_property(env, obj, name, &v) == napi_ok) napi_get_value_double(env, v, &d);
	return d;
}
static void vec3(napi_env env, napi_value obj, const char* name, double* out) {
	size_t n = 0;
	const double* p = (const double*)ta(env, obj, name, &n); /* Float64Array(3) */
	if (p && n >= 3) memcpy(out, p, 3 * sizeof(double));
}

static void finalize_ctx(napi_env env, void* data, void* hint) {
	(void)env; (void)hint;
	rt_destroy((rt_ctx*)data);
}

static napi_value Create(napi_env env, napi_callback_info info) {
	size_t argc = 1;
	napi_value argv[1], out;
	int32_t device = -1;
	rt_ctx* ctx = NULL;
	RT_NAPI_TRY(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	if (argc >= 1) napi_get_value_int32(env, argv[0], &device);
	rt_status st = rt_create(device, &ctx);
	if (st != RT_OK) return throw_rt(env, NULL, st); /* no CUDA device => Error: there is no CPU fallback */
	RT_NAPI_TRY(env, napi_create_external(env, ctx, finalize_ctx, NULL, &out));
	return out;
}

/* createMulti(nGpus) -> ctx driving GPUs 0 .. nGpus-1 from this one process (rt_create_multi) */
static napi_value CreateMulti(napi_env env, napi_callback_info info) {
	size_t argc = 1;
	napi_value argv[1], out;
	int32_t n_gpus = 1;
	rt_ctx* ctx = NULL;
	RT_NAPI_TRY(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	if (argc >= 1) napi_get_value_int32(env, argv[0], &n_gpus);
	rt_status st = rt_create_multi(n_gpus, NULL, &ctx);
	if (st != RT_OK) return throw_rt(env, NULL, st);
	RT_NAPI_TRY(env, napi_create_external(env, ctx, finalize_ctx, NULL, &out));
	return out;
}

static napi_value UploadScene(napi_env env, napi_callback_info info) {
	size_t argc = 2, n = 0;
	napi_value argv[2], flat;
	RT_NAPI_TRY(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	rt_ctx* ctx = ctx_of(env, argv[0]);
	flat = argv[1];
	rt_scene_desc d;
	int bad = 0;
	size_t N = 0, L = 0, E = 0, M = 0, T = 0, S = 0, X = 0;
	memset(&d, 0, sizeof d);
	d.struct_size = (uint32_t)sizeof d;
	/* the counts come from one array per table; every other array is checked against them */
	ta(env, flat, "node_size", &N);
	ta(env, flat, "list_entity", &L);
	ta(env, flat, "ent_type", &E);
	ta(env, flat, "mat_response", &M);
	ta(env, flat, "tex_kind", &T);
	ta(env, flat, "sub_refractive_index", &S);
	ta(env, flat, "texels", &X);
	(void)n;
	d.n_nodes = (uint32_t)N; d.n_list = (uint32_t)L; d.n_entities = (uint32_t)E; d.n_materials = (uint32_t)M;
	d.n_textures = (uint32_t)T; d.n_substances = (uint32_t)S; d.n_texels = X / 3;
#define RT_TA(field, ctype, kind, len) d.field = (const ctype*)ta_checked(env, flat, #field, kind, (len), &bad)
	RT_TA(node_pos, double, napi_float64_array, 3 * N);
	RT_TA(node_size, double, napi_float64_array, N);
	RT_TA(node_child, int32_t, napi_int32_array, 8 * N);
	RT_TA(node_parent, int32_t, napi_int32_array, N);
	RT_TA(node_octant, int32_t, napi_int32_array, N);
	RT_TA(node_list_off, uint32_t, napi_uint32_array, N + 1);
	RT_TA(list_entity, uint32_t, napi_uint32_array, L);
	RT_TA(ent_type, uint8_t, napi_uint8_array, E);
	RT_TA(ent_pos, double, napi_float64_array, 3 * E);
	RT_TA(ent_extent, double, napi_float64_array, E);
	RT_TA(ent_material, int32_t, napi_int32_array, E);
	RT_TA(ent_texture, int32_t, napi_int32_array, E);
	RT_TA(ent_substance, int32_t, napi_int32_array, E);
	RT_TA(mat_response, uint8_t, napi_uint8_array, M);
	RT_TA(mat_light, uint8_t, napi_uint8_array, M);
	RT_TA(mat_mirror, uint8_t, napi_uint8_array, M);
	RT_TA(mat_roughness, double, napi_float64_array, M);
	RT_TA(tex_kind, uint8_t, napi_uint8_array, T);
	RT_TA(tex_color, double, napi_float64_array, 4 * T);
	RT_TA(tex_width, int32_t, napi_int32_array, T);
	RT_TA(tex_height, int32_t, napi_int32_array, T);
	RT_TA(tex_loaded, uint8_t, napi_uint8_array, T);
	RT_TA(tex_texel_off, uint64_t, napi_biguint64_array, T);
	RT_TA(texels, uint8_t, napi_uint8_array, X);
	RT_TA(sub_refractive_index, double, napi_float64_array, S);
#undef RT_TA
	if (bad || X % 3 != 0) {
		if (!bad) napi_throw_error(env, "ERR_OUT_OF_RANGE", "rt_b200: flat.texels must hold RGB triples");
		return NULL;
	}
	rt_status st = rt_scene_upload(ctx, &d);
	if (st != RT_OK) return throw_rt(env, ctx, st);
	return NULL;
}

static napi_value Render(napi_env env, napi_callback_info info) {
	size_t argc = 5, npx = 0, nid = 0;
	napi_value argv[5], out = NULL, v;
	napi_typedarray_type t;
	void *rgb = NULL, *ids = NULL;
	bool want_counters;
	RT_NAPI_TRY(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	rt_ctx* ctx = ctx_of(env, argv[0]);
	rt_camera cam;
	rt_params prm;
	memset(&cam, 0, sizeof cam);
	memset(&prm, 0, sizeof prm);
	vec3(env, argv[1], "pos", cam.pos); vec3(env, argv[1], "fr", cam.fr);
	vec3(env, argv[1], "lf", cam.lf);   vec3(env, argv[1], "up", cam.up);
	cam.fov_h = num(env, argv[1], "fov_h"); cam.fov_v = num(env, argv[1], "fov_v");
	cam.width = (uint32_t)num(env, argv[1], "width"); cam.height = (uint32_t)num(env, argv[1], "height");
	cam.flags = (uint32_t)num(env, argv[1], "flags");
	prm.refmax = (int32_t)num(env, argv[2], "refmax");
	prm.sky_texture = (int32_t)num(env, argv[2], "sky_texture");
	prm.default_substance = (int32_t)num(env, argv[2], "default_substance");
	prm.distance_attenuation_factor = num(env, argv[2], "distance_attenuation_factor");
	prm.n_frames = (uint32_t)num(env, argv[2], "n_frames");
	prm.frame_first = (uint32_t)num(env, argv[2], "frame_first");
	prm.rng_seed = num(env, argv[2], "rng_seed");
	prm.precision = num(env, argv[2], "precision") == 1 ? RT_PRECISION_F64 : RT_PRECISION_F32; /* absent: float32 search */
	prm.flags = num(env, argv[2], "exact_ties") != 0 ? RT_PARAM_EXACT_TIES : 0u;                /* lattice scenes (rt_b200.h) */
	want_counters = num(env, argv[2], "want_counters") != 0;
	RT_NAPI_TRY(env, napi_get_typedarray_info(env, argv[3], &t, &npx, &rgb, NULL, NULL));
	if (t != napi_float32_array || npx != (size_t)cam.width * cam.height * 3) {
		napi_throw_error(env, "ERR_OUT_OF_RANGE", "x or y out of bounds"); /* ExposureBuffer.check_bounds */
		return NULL;
	}
	if (argc >= 5) {
		bool is = false;
		if (napi_is_typedarray(env, argv[4], &is) == napi_ok && is) {
			RT_NAPI_TRY(env, napi_get_typedarray_info(env, argv[4], &t, &nid, &ids, NULL, NULL));
			/* rt_render writes width*height int32 into it: anything else would be a write past the array's end */
			if (t != napi_int32_array || nid != npx / 3) {
				napi_throw_error(env, "ERR_OUT_OF_RANGE", "rt_b200: ids must be an Int32Array of width * height elements");
				return NULL;
			}
		}
	}
	rt_counters cnt;
	rt_status st = rt_render(ctx, &cam, &prm, 0, (float*)rgb, (int32_t*)ids, want_counters ? &cnt : NULL);
	if (st != RT_OK) return throw_rt(env, ctx, st);
	if (!want_counters) return NULL;
	RT_NAPI_TRY(env, napi_create_object(env, &out));
#define PUT(name) napi_create_double(env, (double)cnt.name, &v); napi_set_named_property(env, out, #name, v)
	PUT(paths); PUT(segments); PUT(nodes); PUT(tests); PUT(shades); PUT(confirms);
#undef PUT
	return out;
}

/* present(ctx, pixels: Float32Array, width, height, tone: {kind, dynamic_range, min_dynamic, max_dynamic},
 *         image: Uint8ClampedArray) -> {mean, variance, absolute_dev, drange_low, drange_high}
 * View.draw_ebuffer() (src/view/view.ts:34-38): the ImageData backing store of the canvas is filled on the GPU. */
static napi_value Present(napi_env env, napi_callback_info info) {
	size_t argc = 6, npx = 0, nimg = 0;
	napi_value argv[6], out = NULL, v;
	napi_typedarray_type t;
	void *rgb = NULL, *img = NULL;
	RT_NAPI_TRY(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	rt_ctx* ctx = ctx_of(env, argv[0]);
	uint32_t w = 0, h = 0;
	RT_NAPI_TRY(env, napi_get_typedarray_info(env, argv[1], &t, &npx, &rgb, NULL, NULL));
	RT_NAPI_TRY(env, napi_get_value_uint32(env, argv[2], &w));
	RT_NAPI_TRY(env, napi_get_value_uint32(env, argv[3], &h));
	rt_tone tone;
	memset(&tone, 0, sizeof tone);
	tone.kind = (uint32_t)num(env, argv[4], "kind");
	tone.dynamic_range = (uint32_t)num(env, argv[4], "dynamic_range");
	tone.min_dynamic = num(env, argv[4], "min_dynamic");
	tone.max_dynamic = num(env, argv[4], "max_dynamic");
	RT_NAPI_TRY(env, napi_get_typedarray_info(env, argv[5], &t, &nimg, &img, NULL, NULL));
	if (npx != (size_t)w * h * 3 || nimg != (size_t)w * h * 4) {
		napi_throw_error(env, "ERR_OUT_OF_RANGE", "x or y out of bounds");
		return NULL;
	}
	rt_exposure_stats st;
	rt_status rc = rt_present(ctx, (const float*)rgb, w, h, &tone, (uint8_t*)img, &st);
	if (rc != RT_OK) return throw_rt(env, ctx, rc);
	RT_NAPI_TRY(env, napi_create_object(env, &out));
#define PUT(name) napi_create_double(env, st.name, &v); napi_set_named_property(env, out, #name, v)
	PUT(mean); PUT(variance); PUT(absolute_dev); PUT(drange_low); PUT(drange_high);
#undef PUT
	return out;
}

static napi_value PinOrUnpin(napi_env env, napi_callback_info info, int pin) {
	size_t argc = 2, n = 0;
	napi_value argv[2];
	napi_typedarray_type t;
	void* data = NULL;
	RT_NAPI_TRY(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	rt_ctx* ctx = ctx_of(env, argv[0]);
	RT_NAPI_TRY(env, napi_get_typedarray_info(env, argv[1], &t, &n, &data, NULL, NULL));
	rt_status st = pin ? rt_host_register(ctx, data, n * sizeof(float)) : rt_host_unregister(ctx, data);
	if (st != RT_OK) return throw_rt(env, ctx, st);
	return NULL;
}
static napi_value Pin(napi_env env, napi_callback_info info) { return PinOrUnpin(env, info, 1); }
static napi_value Unpin(napi_env env, napi_callback_info info) { return PinOrUnpin(env, info, 0); }

/* decodeImage(bytes: Uint8Array) -> { width, height, rgb: Uint8Array }: ImageTexture's decode step for a host without
 * a DOM (src/texture/texture_image.ts:76-136 draws the file into a canvas): rt_image_decode, PNG / BMP / binary PPM ->
 * RGB8 rows top to bottom, alpha dropped.  Throws for anything else; the adapter then keeps the fallback colour. */
static napi_value DecodeImage(napi_env env, napi_callback_info info) {
	size_t argc = 1, n = 0;
	napi_value argv[1], out = NULL, ab = NULL, arr = NULL, v;
	napi_typedarray_type t;
	void *bytes = NULL, *dst = NULL;
	uint32_t w = 0, h = 0;
	uint8_t* rgb = NULL;
	RT_NAPI_TRY(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	RT_NAPI_TRY(env, napi_get_typedarray_info(env, argv[0], &t, &n, &bytes, NULL, NULL));
	if (t != napi_uint8_array) {
		napi_throw_error(env, "ERR_INVALID_ARG_TYPE", "rt_b200: decodeImage takes a Uint8Array");
		return NULL;
	}
	rt_status st = rt_image_decode((const uint8_t*)bytes, (uint64_t)n, &w, &h, &rgb);
	if (st != RT_OK) return throw_rt(env, NULL, st);
	if (napi_create_arraybuffer(env, (size_t)w * h * 3, &dst, &ab) != napi_ok ||
	    napi_create_typedarray(env, napi_uint8_array, (size_t)w * h * 3, ab, 0, &arr) != napi_ok) {
		rt_image_free(rgb);
		napi_throw_error(env, NULL, "rt_b200: decodeImage: cannot allocate the texel array");
		return NULL;
	}
	memcpy(dst, rgb, (size_t)w * h * 3);
	rt_image_free(rgb);
	RT_NAPI_TRY(env, napi_create_object(env, &out));
	napi_create_double(env, (double)w, &v); napi_set_named_property(env, out, "width", v);
	napi_create_double(env, (double)h, &v); napi_set_named_property(env, out, "height", v);
	napi_set_named_property(env, out, "rgb", arr);
	return out;
}

napi_value napi_register_module_v1(napi_env env, napi_value exports) {
	const napi_property_descriptor props[] = {
	    {"create", NULL, Create, NULL, NULL, NULL, napi_default, NULL},
	    {"createMulti", NULL, CreateMulti, NULL, NULL, NULL, napi_default, NULL},
	    {"uploadScene", NULL, UploadScene, NULL, NULL, NULL, napi_default, NULL},
	    {"render", NULL, Render, NULL, NULL, NULL, napi_default, NULL},
	    {"present", NULL, Present, NULL, NULL, NULL, napi_default, NULL},
	    {"pin", NULL, Pin, NULL, NULL, NULL, napi_default, NULL},
	    {"unpin", NULL, Unpin, NULL, NULL, NULL, napi_default, NULL},
	    {"decodeImage", NULL, DecodeImage, NULL, NULL, NULL, napi_default, NULL},
	};
	napi_define_properties(env, exports, sizeof props / sizeof props[0], props);
	return exports;
}
