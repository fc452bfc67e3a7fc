/* node_api_min.h — the handful of Node-API (N-API v8) declarations rt_napi.c uses, hand-declared because
 * this build image has no Node headers.  On a machine with Node, delete this file and
 * `#include <node_api.h>` instead (the names and signatures below are Node's own, ABI-stable). */
#ifndef NODE_API_MIN_H
#define NODE_API_MIN_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef struct napi_ref__* napi_ref;
typedef enum { napi_ok = 0 } napi_status;
typedef enum { napi_default = 0 } napi_property_attributes;
typedef enum {
	napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array, napi_int32_array,
	napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array, napi_biguint64_array
} napi_typedarray_type;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* data, void* hint);
typedef struct {
	const char* utf8name; napi_value name; napi_callback method; napi_callback getter; napi_callback setter;
	napi_value value; napi_property_attributes attributes; void* data;
} napi_property_descriptor;

napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_get_named_property(napi_env, napi_value object, const char* name, napi_value* result);
napi_status napi_get_value_double(napi_env, napi_value, double* result);
napi_status napi_get_value_int32(napi_env, napi_value, int32_t* result);
napi_status napi_get_value_uint32(napi_env, napi_value, uint32_t* result);
napi_status napi_get_typedarray_info(napi_env, napi_value, napi_typedarray_type* type, size_t* length, void** data,
                                     napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_is_typedarray(napi_env, napi_value, bool* result);
napi_status napi_create_external(napi_env, void* data, napi_finalize finalize_cb, void* hint, napi_value* result);
napi_status napi_get_value_external(napi_env, napi_value, void** result);
napi_status napi_create_object(napi_env, napi_value* result);
napi_status napi_create_arraybuffer(napi_env, size_t byte_length, void** data, napi_value* result);
napi_status napi_create_typedarray(napi_env, napi_typedarray_type type, size_t length, napi_value arraybuffer, size_t byte_offset,
                                   napi_value* result);
napi_status napi_create_double(napi_env, double, napi_value* result);
napi_status napi_set_named_property(napi_env, napi_value object, const char* name, napi_value value);
napi_status napi_get_undefined(napi_env, napi_value* result);
napi_status napi_throw_error(napi_env, const char* code, const char* msg);
napi_status napi_define_properties(napi_env, napi_value object, size_t count, const napi_property_descriptor* props);
#endif
