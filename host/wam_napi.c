/*
 * wam_napi.c — N-API addon (plain C, node_api.h only) binding libwam.so's C ABI (include/wam.h)
 * for host/fsk_core_gpu.ts.  NOT COMPILED in the build environment (no Node headers in the image).
 *
 * Shape of the binding:
 *   - handles are napi_wrap'ped objects whose finalizers call wam_fsk_destroy / wam_fsk_batch_destroy;
 *   - demodulate / modulate run on napi_async_work so the Promise settles off the JS thread; the
 *     ArrayBuffers are pinned with napi_create_reference for the duration of the work item, and the
 *     single-stream demodulate writes the AGC-scaled samples back into the caller's Float32Array
 *     (the reference mutates its input, fsk.ts:55);
 *   - a negative wam_error rejects the Promise with wam_last_error(); WAM_E_NOT_CONFIGURED carries the
 *     reference's message ('FSK modulator not configured' / 'FSK demodulator not configured').
 * Only the single-stream pair is spelled out; the batch entry points follow the same pattern around
 * wam_fsk_batch_demodulate / wam_fsk_batch_modulate / wam_xmodem_batch_check.
 */
#include <node_api.h>
#include <stdlib.h>
#include <string.h>

#include "../include/wam.h"

#define NAPI_OK(call) do { if ((call) != napi_ok) { napi_throw_error(env, NULL, #call " failed"); return NULL; } } while (0)

static void fsk_finalize(napi_env env, void* data, void* hint) { (void)env; (void)hint; wam_fsk_destroy((wam_fsk*)data); }

static napi_value FskCreate(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1]; int32_t device = 0;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc > 0) NAPI_OK(napi_get_value_int32(env, argv[0], &device));
  wam_fsk* m = NULL;
  if (wam_fsk_create(device, &m) != WAM_OK) { napi_throw_error(env, NULL, wam_last_error()); return NULL; }
  napi_value obj; NAPI_OK(napi_create_object(env, &obj));
  NAPI_OK(napi_wrap(env, obj, m, fsk_finalize, NULL, NULL));
  return obj;
}

/* reads {sampleRate, baudRate, markFrequency, ...} (src/modems/fsk.ts:5-17) into wam_fsk_config */
static int read_config(napi_env env, napi_value js, wam_fsk_config* c, uint8_t* pre, uint8_t* sfd) {
  napi_value v; double d; bool b; uint32_t n;
#define NUM(field) if (napi_get_named_property(env, js, #field, &v) != napi_ok || napi_get_value_double(env, v, &d) != napi_ok) return -1; c->field = d
  NUM(sampleRate); NUM(baudRate); NUM(markFrequency); NUM(spaceFrequency); NUM(syncThreshold); NUM(preFilterBandwidth);
#undef NUM
#define INT(field) if (napi_get_named_property(env, js, #field, &v) != napi_ok || napi_get_value_double(env, v, &d) != napi_ok) return -1; c->field = (int32_t)d
  INT(startBits); INT(stopBits);
#undef INT
  if (napi_get_named_property(env, js, "agcEnabled", &v) != napi_ok || napi_get_value_bool(env, v, &b) != napi_ok) return -1;
  c->agcEnabled = b;
  if (napi_get_named_property(env, js, "adaptiveThreshold", &v) == napi_ok && napi_get_value_bool(env, v, &b) == napi_ok) c->adaptiveThreshold = b;
  char parity[8] = {0}; size_t len = 0;
  if (napi_get_named_property(env, js, "parity", &v) != napi_ok || napi_get_value_string_utf8(env, v, parity, sizeof(parity), &len) != napi_ok) return -1;
  c->parity = strcmp(parity, "even") == 0 ? 1 : strcmp(parity, "odd") == 0 ? 2 : 0;
  const char* names[2] = {"preamblePattern", "sfdPattern"}; uint8_t* dst[2] = {pre, sfd}; int32_t* lens[2] = {&c->preambleLength, &c->sfdLength};
  for (int k = 0; k < 2; k++) {
    napi_value arr, e;
    if (napi_get_named_property(env, js, names[k], &arr) != napi_ok || napi_get_array_length(env, arr, &n) != napi_ok || n > 32) return -1;
    for (uint32_t i = 0; i < n; i++) { if (napi_get_element(env, arr, i, &e) != napi_ok || napi_get_value_double(env, e, &d) != napi_ok) return -1; dst[k][i] = (uint8_t)d; }
    *lens[k] = (int32_t)n;
  }
  c->preamblePattern = pre; c->sfdPattern = sfd;
  return 0;
}

static napi_value FskConfigure(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value argv[2]; wam_fsk* m;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  NAPI_OK(napi_unwrap(env, argv[0], (void**)&m));
  wam_fsk_config c; uint8_t pre[32], sfd[32];
  memset(&c, 0, sizeof(c));
  if (read_config(env, argv[1], &c, pre, sfd) != 0) { napi_throw_type_error(env, NULL, "bad FSKConfig"); return NULL; }
  if (wam_fsk_configure(m, &c) != WAM_OK) napi_throw_error(env, NULL, wam_last_error());
  return NULL;
}

typedef struct {
  napi_async_work work; napi_deferred deferred; napi_ref samples_ref;
  wam_fsk* m; float* samples; long n; uint8_t* out; long cap, n_out; int rc; char err[256]; double eod_before, eod_after;
} demod_job;

static void demod_execute(napi_env env, void* data) {
  (void)env; demod_job* j = (demod_job*)data; wam_fsk_status st;
  wam_fsk_status_get(j->m, &st); j->eod_before = st.eodEvents;
  j->rc = wam_fsk_demodulate(j->m, j->samples, j->n, j->out, j->cap, &j->n_out);   /* mutates j->samples when AGC is on */
  if (j->rc != WAM_OK) { strncpy(j->err, wam_last_error(), sizeof(j->err) - 1); return; }
  wam_fsk_status_get(j->m, &st); j->eod_after = st.eodEvents;
}

static void demod_complete(napi_env env, napi_status status, void* data) {
  demod_job* j = (demod_job*)data; napi_value result, bytes, eod, msg, err; void* dst;
  if (status == napi_ok && j->rc == WAM_OK) {
    napi_value ab; napi_create_arraybuffer(env, (size_t)j->n_out, &dst, &ab);
    memcpy(dst, j->out, (size_t)j->n_out);
    napi_create_typedarray(env, napi_uint8_array, (size_t)j->n_out, ab, 0, &bytes);
    napi_create_double(env, j->eod_after - j->eod_before, &eod);
    napi_create_object(env, &result);
    napi_set_named_property(env, result, "bytes", bytes); napi_set_named_property(env, result, "eod", eod);
    napi_resolve_deferred(env, j->deferred, result);
  } else {
    napi_create_string_utf8(env, j->rc == WAM_E_NOT_CONFIGURED ? "FSK demodulator not configured" : j->err, NAPI_AUTO_LENGTH, &msg);
    napi_create_error(env, NULL, msg, &err); napi_reject_deferred(env, j->deferred, err);
  }
  napi_delete_reference(env, j->samples_ref); napi_delete_async_work(env, j->work); free(j->out); free(j);
}

static napi_value FskDemodulate(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value argv[2], promise, name; wam_fsk* m;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  NAPI_OK(napi_unwrap(env, argv[0], (void**)&m));
  napi_typedarray_type t; size_t len; void* data; napi_value ab; size_t off;
  NAPI_OK(napi_get_typedarray_info(env, argv[1], &t, &len, &data, &ab, &off));
  if (t != napi_float32_array) { napi_throw_type_error(env, NULL, "samples must be a Float32Array"); return NULL; }
  demod_job* j = (demod_job*)calloc(1, sizeof(*j));
  j->m = m; j->samples = (float*)data; j->n = (long)len; j->cap = (long)len / 8 + 16; j->out = (uint8_t*)malloc((size_t)j->cap);
  NAPI_OK(napi_create_reference(env, argv[1], 1, &j->samples_ref));
  NAPI_OK(napi_create_promise(env, &j->deferred, &promise));
  NAPI_OK(napi_create_string_utf8(env, "wam_fsk_demodulate", NAPI_AUTO_LENGTH, &name));
  NAPI_OK(napi_create_async_work(env, NULL, name, demod_execute, demod_complete, j, &j->work));
  NAPI_OK(napi_queue_async_work(env, j->work));
  return promise;
}

static napi_value FskReset(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1]; wam_fsk* m;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  NAPI_OK(napi_unwrap(env, argv[0], (void**)&m));
  wam_fsk_reset(m);
  return NULL;
}

static napi_value Init(napi_env env, napi_value exports) {
  napi_property_descriptor props[] = {
      {"fskCreate", NULL, FskCreate, NULL, NULL, NULL, napi_default, NULL},
      {"fskConfigure", NULL, FskConfigure, NULL, NULL, NULL, napi_default, NULL},
      {"fskDemodulate", NULL, FskDemodulate, NULL, NULL, NULL, napi_default, NULL},
      {"fskReset", NULL, FskReset, NULL, NULL, NULL, napi_default, NULL},
      /* fskModulate, fskStatus, batchCreate, batchDemodulate, batchModulate, xmodemBatchCheck: same pattern */
  };
  napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
  return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
