/*
 * node_api.h -- TEST STUB, not Node's header.  Declares exactly the N-API entities that
 * addon/jsmpeg_b200_napi.c uses, with the signatures of Node's js_native_api.h / node_api.h
 * (N-API version 1-3 subset), so that `gcc -fsyntax-only` can prove the addon source parses and
 * type-checks in an image that has no Node toolchain (tests/test_host_logic.py).
 */
#ifndef JSMPEG_B200_TEST_STUB_NODE_API_H
#define JSMPEG_B200_TEST_STUB_NODE_API_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_callback_info__ *napi_callback_info;

typedef enum { napi_ok = 0, napi_invalid_arg, napi_object_expected, napi_generic_failure = 9 } napi_status;
typedef enum {
	napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array,
	napi_int32_array, napi_uint32_array, napi_float32_array, napi_float64_array
} napi_typedarray_type;

typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void *finalize_data, void *finalize_hint);

#define NAPI_AUTO_LENGTH SIZE_MAX

napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t *argc, napi_value *argv, napi_value *this_arg, void **data);
napi_status napi_get_value_external(napi_env env, napi_value value, void **result);
napi_status napi_get_value_uint32(napi_env env, napi_value value, uint32_t *result);
napi_status napi_create_external(napi_env env, void *data, napi_finalize finalize_cb, void *finalize_hint, napi_value *result);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type *type, size_t *length, void **data,
                                     napi_value *arraybuffer, size_t *byte_offset);
napi_status napi_create_uint32(napi_env env, uint32_t value, napi_value *result);
napi_status napi_create_int32(napi_env env, int32_t value, napi_value *result);
napi_status napi_create_double(napi_env env, double value, napi_value *result);
napi_status napi_get_boolean(napi_env env, bool value, napi_value *result);
napi_status napi_create_external_arraybuffer(napi_env env, void *external_data, size_t byte_length, napi_finalize finalize_cb,
                                             void *finalize_hint, napi_value *result);
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value arraybuffer, size_t byte_offset,
                                   napi_value *result);
napi_status napi_create_object(napi_env env, napi_value *result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char *utf8name, napi_value value);
napi_status napi_create_function(napi_env env, const char *utf8name, size_t length, napi_callback cb, void *data, napi_value *result);
napi_status napi_throw_error(napi_env env, const char *code, const char *msg);

typedef napi_value (*napi_addon_register_func)(napi_env env, napi_value exports);
#ifndef NODE_GYP_MODULE_NAME
#define NODE_GYP_MODULE_NAME jsmpeg_b200
#endif
/* Node: registers `regfunc` as the module initialiser.  Stub: reference it so that its type is checked. */
#define NAPI_MODULE(modname, regfunc) napi_addon_register_func jsmpeg_b200_stub_registered_init = (regfunc);

#endif
