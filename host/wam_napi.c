/*
 * wam_napi.c — N-API addon (plain C, node_api.h only) binding libwam.so's C ABI (include/wam.h)
 * for host/fsk_core_gpu.ts.  There is no Node toolchain in the build image: the file is checked for
 * syntax and for agreement with include/wam.h against a stub node_api.h (tests/stubs/node_api.h,
 * tests/test_lib_and_host.py), it is not executed here.
 *
 * Shape of the binding:
 *   - handles are napi_wrap'ped objects whose finalizers call the matching wam_*_destroy;
 *   - everything that launches GPU work runs on napi_async_work, so the Promise settles off the JS
 *     thread; typed arrays handed to a job are pinned with napi_create_reference until it completes;
 *     the single-stream demodulate writes the AGC-scaled samples back into the caller's Float32Array
 *     (the reference mutates its input, fsk.ts:55);
 *   - a negative wam_error rejects the Promise with wam_last_error(); WAM_E_NOT_CONFIGURED carries the
 *     reference's messages ('FSK modulator not configured' / 'FSK demodulator not configured').
 *
 * Natives (the `native` object of host/fsk_core_gpu.ts):
 *   fskCreate fskConfigure fskModulate fskDemodulate fskReset fskStatus
 *   batchCreate batchDemodulate batchModulate batchStatus xmodemBatchCheck
 *   muxCreate muxPush muxFlush muxSend muxModulate muxPull
 *   xmodemReceiverCreate xmodemReceiverFeed xmodemReceiverData
 */
#include <node_api.h>
#include <stdlib.h>
#include <string.h>

#include "../include/wam.h"

#define NAPI_OK(call) do { if ((call) != napi_ok) { napi_throw_error(env, NULL, #call " failed"); return NULL; } } while (0)
#define WAM_OK_OR_THROW(call) do { if ((call) < 0) { napi_throw_error(env, NULL, wam_last_error()); return NULL; } } while (0)
#define MAX_CFGS 16

/* ---- arguments ---------------------------------------------------------------------------------------------- */
static int get_args(napi_env env, napi_callback_info info, size_t want, napi_value* argv) {
  size_t argc = want;
  for (size_t i = 0; i < want; i++) argv[i] = NULL;
  return napi_get_cb_info(env, info, &argc, argv, NULL, NULL) == napi_ok ? (int)argc : -1;
}
static int is_nullish(napi_env env, napi_value v) {
  napi_valuetype t;
  return v == NULL || napi_typeof(env, v, &t) != napi_ok || t == napi_undefined || t == napi_null;
}
static long get_long(napi_env env, napi_value v, long dflt) {
  double d;
  return (!is_nullish(env, v) && napi_get_value_double(env, v, &d) == napi_ok) ? (long)d : dflt;
}
/* typed array of the wanted element type (or any when want < 0): data pointer and element count */
static void* typed(napi_env env, napi_value v, int want, size_t* len) {
  napi_typedarray_type t; void* data = NULL; napi_value ab; size_t off;
  *len = 0;
  if (is_nullish(env, v) || napi_get_typedarray_info(env, v, &t, len, &data, &ab, &off) != napi_ok) return NULL;
  if (want >= 0 && (int)t != want) { *len = 0; return NULL; }
  return data;
}
static napi_value make_typed(napi_env env, napi_typedarray_type t, size_t elem, const void* src, size_t count) {
  napi_value ab, out; void* dst;
  if (napi_create_arraybuffer(env, count * elem, &dst, &ab) != napi_ok) return NULL;
  if (count) memcpy(dst, src, count * elem);
  return napi_create_typedarray(env, t, count, ab, 0, &out) == napi_ok ? out : NULL;
}
static void set_num(napi_env env, napi_value obj, const char* k, double v) {
  napi_value n;
  if (napi_create_double(env, v, &n) == napi_ok) napi_set_named_property(env, obj, k, n);
}

/* reads {sampleRate, baudRate, markFrequency, ...} (src/modems/fsk.ts:5-17) into wam_fsk_config */
static int read_config(napi_env env, napi_value js, wam_fsk_config* c, uint8_t* pre, uint8_t* sfd) {
  napi_value v; double d; bool b; uint32_t n;
  memset(c, 0, sizeof(*c));
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
/* FSKConfig[] -> cfgs (pattern bytes live in `pat`, 64 per config) */
static int read_configs(napi_env env, napi_value arr, wam_fsk_config* cfgs, uint8_t* pat) {
  uint32_t n = 0; napi_value e;
  if (napi_get_array_length(env, arr, &n) != napi_ok || n == 0 || n > MAX_CFGS) return -1;
  for (uint32_t i = 0; i < n; i++)
    if (napi_get_element(env, arr, i, &e) != napi_ok || read_config(env, e, &cfgs[i], pat + 64 * i, pat + 64 * i + 32) != 0) return -1;
  return (int)n;
}
static napi_value status_object(napi_env env, const wam_fsk_status* s) {  /* getStatus(), fsk.ts:481-493 */
  napi_value o, b;
  if (napi_create_object(env, &o) != napi_ok) return NULL;
  if (napi_get_boolean(env, s->ready != 0, &b) == napi_ok) napi_set_named_property(env, o, "ready", b);
  if (napi_get_boolean(env, s->frameStarted != 0, &b) == napi_ok) napi_set_named_property(env, o, "frameStarted", b);
  set_num(env, o, "globalSampleCounter", s->globalSampleCounter); set_num(env, o, "receivedBitsLength", s->receivedBitsLength);
  set_num(env, o, "byteBufferLength", s->byteBufferLength); set_num(env, o, "demodulationCalls", s->demodulationCalls);
  set_num(env, o, "syncDetections", s->syncDetections); set_num(env, o, "silenceThreshold", s->silenceThreshold);
  set_num(env, o, "totalSamplesProcessed", s->totalSamplesProcessed); set_num(env, o, "eodEvents", s->eodEvents);
  set_num(env, o, "errorEvents", s->errorEvents);
  return o;
}

/* ---- asynchronous jobs ---------------------------------------------------------------------------------------- */
typedef struct job {
  napi_async_work work; napi_deferred deferred; napi_ref keep[3]; int nkeep;
  void (*run)(struct job*);                 /* worker thread: calls into libwam, sets rc */
  napi_value (*result)(napi_env, struct job*);  /* JS thread: builds the resolved value */
  void* h; void* p[3]; long n[6]; void* out[3]; long n_out[3];
  int rc; const char* not_configured; char err[256];
} job;

static void job_execute(napi_env env, void* data) {
  (void)env; job* j = (job*)data;
  j->run(j);
  if (j->rc < 0) strncpy(j->err, wam_last_error(), sizeof(j->err) - 1);
}
static void job_complete(napi_env env, napi_status status, void* data) {
  job* j = (job*)data; napi_value v = NULL, msg, err;
  if (status == napi_ok && j->rc >= 0) v = j->result(env, j);
  if (v) napi_resolve_deferred(env, j->deferred, v);
  else {
    const char* m = (j->rc == WAM_E_NOT_CONFIGURED && j->not_configured) ? j->not_configured : (j->err[0] ? j->err : "wam: job failed");
    napi_create_string_utf8(env, m, NAPI_AUTO_LENGTH, &msg); napi_create_error(env, NULL, msg, &err);
    napi_reject_deferred(env, j->deferred, err);
  }
  for (int i = 0; i < j->nkeep; i++) napi_delete_reference(env, j->keep[i]);
  napi_delete_async_work(env, j->work);
  for (int i = 0; i < 3; i++) free(j->out[i]);
  free(j);
}
static job* job_new(void) { return (job*)calloc(1, sizeof(job)); }
static int job_keep(napi_env env, job* j, napi_value v) {
  return (is_nullish(env, v) || napi_create_reference(env, v, 1, &j->keep[j->nkeep++]) == napi_ok) ? 0 : -1;
}
static napi_value job_start(napi_env env, job* j, const char* what) {
  napi_value promise, name;
  if (napi_create_promise(env, &j->deferred, &promise) != napi_ok || napi_create_string_utf8(env, what, NAPI_AUTO_LENGTH, &name) != napi_ok ||
      napi_create_async_work(env, NULL, name, job_execute, job_complete, j, &j->work) != napi_ok || napi_queue_async_work(env, j->work) != napi_ok) {
    napi_throw_error(env, NULL, "wam: cannot queue the job");
    return NULL;
  }
  return promise;
}

/* ---- FSKCore: one stream ------------------------------------------------------------------------------------ */
static void fsk_finalize(napi_env env, void* data, void* hint) { (void)env; (void)hint; wam_fsk_destroy((wam_fsk*)data); }

static napi_value FskCreate(napi_env env, napi_callback_info info) {
  napi_value argv[1], obj; wam_fsk* m = NULL;
  if (get_args(env, info, 1, argv) < 0) return NULL;
  WAM_OK_OR_THROW(wam_fsk_create((int)get_long(env, argv[0], 0), &m));
  NAPI_OK(napi_create_object(env, &obj));
  NAPI_OK(napi_wrap(env, obj, m, fsk_finalize, NULL, NULL));
  return obj;
}
static napi_value FskConfigure(napi_env env, napi_callback_info info) {
  napi_value argv[2]; wam_fsk* m; wam_fsk_config c; uint8_t pre[32], sfd[32];
  if (get_args(env, info, 2, argv) < 2) return NULL;
  NAPI_OK(napi_unwrap(env, argv[0], (void**)&m));
  if (read_config(env, argv[1], &c, pre, sfd) != 0) { napi_throw_type_error(env, NULL, "bad FSKConfig"); return NULL; }
  WAM_OK_OR_THROW(wam_fsk_configure(m, &c));
  return NULL;
}
static napi_value FskReset(napi_env env, napi_callback_info info) {
  napi_value argv[1]; wam_fsk* m;
  if (get_args(env, info, 1, argv) < 1) return NULL;
  NAPI_OK(napi_unwrap(env, argv[0], (void**)&m));
  wam_fsk_reset(m);
  return NULL;
}
static napi_value FskStatus(napi_env env, napi_callback_info info) {
  napi_value argv[1]; wam_fsk* m; wam_fsk_status st;
  if (get_args(env, info, 1, argv) < 1) return NULL;
  NAPI_OK(napi_unwrap(env, argv[0], (void**)&m));
  WAM_OK_OR_THROW(wam_fsk_status_get(m, &st));
  return status_object(env, &st);
}

static void fsk_modulate_run(job* j) {
  long cap = wam_fsk_modulate_size((wam_fsk*)j->h, j->n[0]);
  if (cap < 0) { j->rc = (int)cap; return; }
  j->out[0] = malloc(sizeof(float) * (size_t)(cap > 0 ? cap : 1));
  j->rc = wam_fsk_modulate((wam_fsk*)j->h, (const uint8_t*)j->p[0], j->n[0], (float*)j->out[0], cap, &j->n_out[0]);
}
static napi_value fsk_modulate_result(napi_env env, job* j) { return make_typed(env, napi_float32_array, sizeof(float), j->out[0], (size_t)j->n_out[0]); }
static napi_value FskModulate(napi_env env, napi_callback_info info) {  /* modulateData(), fsk.ts:377-462 */
  napi_value argv[2]; size_t len; job* j = job_new();
  if (get_args(env, info, 2, argv) < 2) return NULL;
  NAPI_OK(napi_unwrap(env, argv[0], &j->h));
  j->p[0] = typed(env, argv[1], napi_uint8_array, &len); j->n[0] = (long)len;
  j->run = fsk_modulate_run; j->result = fsk_modulate_result; j->not_configured = "FSK modulator not configured";
  if (job_keep(env, j, argv[1]) != 0) return NULL;
  return job_start(env, j, "wam_fsk_modulate");
}

static void fsk_demodulate_run(job* j) {
  wam_fsk_status st;
  wam_fsk_status_get((wam_fsk*)j->h, &st); double eod0 = st.eodEvents;
  j->out[0] = malloc((size_t)j->n[1]);
  j->rc = wam_fsk_demodulate((wam_fsk*)j->h, (float*)j->p[0], j->n[0], (uint8_t*)j->out[0], j->n[1], &j->n_out[0]);  /* mutates the samples when AGC is on */
  if (j->rc >= 0 && wam_fsk_status_get((wam_fsk*)j->h, &st) == WAM_OK) j->n_out[1] = (long)(st.eodEvents - eod0);
}
static napi_value fsk_demodulate_result(napi_env env, job* j) {
  napi_value o;
  if (napi_create_object(env, &o) != napi_ok) return NULL;
  napi_set_named_property(env, o, "bytes", make_typed(env, napi_uint8_array, 1, j->out[0], (size_t)j->n_out[0]));
  set_num(env, o, "eod", (double)j->n_out[1]);  /* 'eod' events to re-emit, fsk.ts:289 */
  return o;
}
static napi_value FskDemodulate(napi_env env, napi_callback_info info) {  /* demodulateData(), fsk.ts:190-222 */
  napi_value argv[2]; size_t len; job* j = job_new();
  if (get_args(env, info, 2, argv) < 2) return NULL;
  NAPI_OK(napi_unwrap(env, argv[0], &j->h));
  j->p[0] = typed(env, argv[1], napi_float32_array, &len);
  if (!j->p[0]) {  /* an empty Float32Array has no data pointer; anything else is a type error */
    size_t any; napi_typedarray_type t; void* d; napi_value ab; size_t off;
    if (napi_get_typedarray_info(env, argv[1], &t, &any, &d, &ab, &off) != napi_ok || t != napi_float32_array) { free(j); napi_throw_type_error(env, NULL, "samples must be a Float32Array"); return NULL; }
  }
  j->n[0] = (long)len; j->n[1] = (long)len / 8 + 16;
  j->run = fsk_demodulate_run; j->result = fsk_demodulate_result; j->not_configured = "FSK demodulator not configured";
  if (job_keep(env, j, argv[1]) != 0) return NULL;
  return job_start(env, j, "wam_fsk_demodulate");
}

/* ---- batch -------------------------------------------------------------------------------------------------- */
typedef struct { wam_fsk_batch* b; long n_streams; } batch_handle;
static void batch_finalize(napi_env env, void* data, void* hint) { (void)env; (void)hint; batch_handle* h = (batch_handle*)data; wam_fsk_batch_destroy(h->b); free(h); }

static napi_value BatchCreate(napi_env env, napi_callback_info info) {  /* (device, nStreams, cfgs, cfgIndex?) */
  napi_value argv[4], obj; wam_fsk_config cfgs[MAX_CFGS]; uint8_t pat[64 * MAX_CFGS]; size_t len;
  if (get_args(env, info, 4, argv) < 3) return NULL;
  int n_cfgs = read_configs(env, argv[2], cfgs, pat);
  if (n_cfgs < 0) { napi_throw_type_error(env, NULL, "bad FSKConfig list"); return NULL; }
  batch_handle* h = (batch_handle*)calloc(1, sizeof(*h));
  h->n_streams = get_long(env, argv[1], 0);
  const int32_t* idx = (const int32_t*)typed(env, argv[3], napi_int32_array, &len);
  if (idx && (long)len != h->n_streams) { free(h); napi_throw_type_error(env, NULL, "cfgIndex needs one entry per stream"); return NULL; }
  if (wam_fsk_batch_create((int)get_long(env, argv[0], 0), h->n_streams, cfgs, n_cfgs, idx, &h->b) < 0) { free(h); napi_throw_error(env, NULL, wam_last_error()); return NULL; }
  NAPI_OK(napi_create_object(env, &obj));
  NAPI_OK(napi_wrap(env, obj, h, batch_finalize, NULL, NULL));
  return obj;
}
static void batch_demodulate_run(job* j) {  /* n: 0 nSamples, 1 stride, 2 is_pcm16 */
  batch_handle* h = (batch_handle*)j->h;
  long cap = wam_fsk_batch_out_capacity(h->b, j->n[0]);
  if (cap < 0) { j->rc = (int)cap; return; }
  j->n_out[0] = cap;
  j->out[0] = calloc((size_t)h->n_streams, (size_t)(cap > 0 ? cap : 1));
  j->out[1] = calloc((size_t)h->n_streams, sizeof(int32_t));
  j->rc = j->n[2] ? wam_fsk_batch_demodulate_pcm16(h->b, (const int16_t*)j->p[0], j->n[1], j->n[0], (uint8_t*)j->out[0], cap, (int32_t*)j->out[1], 0)
                  : wam_fsk_batch_demodulate(h->b, (float*)j->p[0], j->n[1], j->n[0], (uint8_t*)j->out[0], cap, (int32_t*)j->out[1], 0);
}
static napi_value rows_result(napi_env env, job* j) {  /* {bytes, lengths, stride} */
  batch_handle* h = (batch_handle*)j->h; napi_value o;
  if (napi_create_object(env, &o) != napi_ok) return NULL;
  napi_set_named_property(env, o, "bytes", make_typed(env, napi_uint8_array, 1, j->out[0], (size_t)(h->n_streams * j->n_out[0])));
  napi_set_named_property(env, o, "lengths", make_typed(env, napi_int32_array, sizeof(int32_t), j->out[1], (size_t)h->n_streams));
  set_num(env, o, "stride", (double)j->n_out[0]);
  return o;
}
static napi_value BatchDemodulate(napi_env env, napi_callback_info info) {  /* (h, Float32Array | Int16Array [n][nSamples], nSamples) */
  napi_value argv[3]; size_t len; job* j = job_new();
  if (get_args(env, info, 3, argv) < 3) return NULL;
  NAPI_OK(napi_unwrap(env, argv[0], &j->h));
  j->n[0] = get_long(env, argv[2], 0); j->n[1] = j->n[0];
  j->p[0] = typed(env, argv[1], napi_float32_array, &len);
  if (!j->p[0]) { j->p[0] = typed(env, argv[1], napi_int16_array, &len); j->n[2] = 1; }
  if ((long)len != ((batch_handle*)j->h)->n_streams * j->n[0]) { free(j); napi_throw_type_error(env, NULL, "samples must hold nStreams x nSamples float32 or int16 values"); return NULL; }
  j->run = batch_demodulate_run; j->result = rows_result;
  if (job_keep(env, j, argv[1]) != 0) return NULL;
  return job_start(env, j, "wam_fsk_batch_demodulate");
}
static void batch_modulate_run(job* j) {  /* n: 0 nBytes per row; p: 0 data, 1 lengths (nullable) */
  batch_handle* h = (batch_handle*)j->h;
  long per = j->n[1];  /* samples per row, from the caller (FSKBatchGPU computes it with modulate_size) */
  j->out[0] = calloc((size_t)h->n_streams * (size_t)(per > 0 ? per : 1), sizeof(float));
  j->out[1] = calloc((size_t)h->n_streams, sizeof(int32_t));
  j->n_out[0] = per;
  j->rc = wam_fsk_batch_modulate(h->b, (const uint8_t*)j->p[0], j->n[0], (const int32_t*)j->p[1], j->n[0], (float*)j->out[0], per, (int32_t*)j->out[1]);
}
static napi_value batch_modulate_result(napi_env env, job* j) {  /* {samples, lengths, stride} */
  batch_handle* h = (batch_handle*)j->h; napi_value o;
  if (napi_create_object(env, &o) != napi_ok) return NULL;
  napi_set_named_property(env, o, "samples", make_typed(env, napi_float32_array, sizeof(float), j->out[0], (size_t)(h->n_streams * j->n_out[0])));
  napi_set_named_property(env, o, "lengths", make_typed(env, napi_int32_array, sizeof(int32_t), j->out[1], (size_t)h->n_streams));
  set_num(env, o, "stride", (double)j->n_out[0]);
  return o;
}
static napi_value BatchModulate(napi_env env, napi_callback_info info) {  /* (h, Uint8Array [n][nBytes], nBytes, samplesPerRow, lengths?) */
  napi_value argv[5]; size_t len; job* j = job_new();
  if (get_args(env, info, 5, argv) < 4) return NULL;
  NAPI_OK(napi_unwrap(env, argv[0], &j->h));
  j->p[0] = typed(env, argv[1], napi_uint8_array, &len);
  j->n[0] = get_long(env, argv[2], 0); j->n[1] = get_long(env, argv[3], 0);
  j->p[1] = typed(env, argv[4], napi_int32_array, &len);
  j->run = batch_modulate_run; j->result = batch_modulate_result; j->not_configured = "FSK modulator not configured";
  if (job_keep(env, j, argv[1]) != 0 || job_keep(env, j, argv[4]) != 0) return NULL;
  return job_start(env, j, "wam_fsk_batch_modulate");
}
static napi_value BatchStatus(napi_env env, napi_callback_info info) {  /* getStatus() of every stream */
  napi_value argv[1], arr; batch_handle* h;
  if (get_args(env, info, 1, argv) < 1) return NULL;
  NAPI_OK(napi_unwrap(env, argv[0], (void**)&h));
  wam_fsk_status* st = (wam_fsk_status*)calloc((size_t)h->n_streams, sizeof(*st));
  if (wam_fsk_batch_status(h->b, st) < 0) { free(st); napi_throw_error(env, NULL, wam_last_error()); return NULL; }
  NAPI_OK(napi_create_array_with_length(env, (size_t)h->n_streams, &arr));
  for (long s = 0; s < h->n_streams; s++) napi_set_element(env, arr, (uint32_t)s, status_object(env, &st[s]));
  free(st);
  return arr;
}
static napi_value XmodemBatchCheck(napi_env env, napi_callback_info info) {  /* (device, bytes, stride, lengths, expectedSeq?) -> Int32Array [n][7] */
  napi_value argv[5]; size_t nb, nl, ne;
  if (get_args(env, info, 5, argv) < 4) return NULL;
  const uint8_t* bytes = (const uint8_t*)typed(env, argv[1], napi_uint8_array, &nb);
  const int32_t* lens = (const int32_t*)typed(env, argv[3], napi_int32_array, &nl);
  const int32_t* exp = (const int32_t*)typed(env, argv[4], napi_int32_array, &ne);
  long stride = get_long(env, argv[2], 0);
  if (!lens || (exp && ne != nl) || (long)nb < stride * (long)nl) { napi_throw_type_error(env, NULL, "bad packet buffers"); return NULL; }
  wam_pkt_result* r = (wam_pkt_result*)calloc(nl ? nl : 1, sizeof(*r));
  if (wam_xmodem_batch_check((int)get_long(env, argv[0], 0), bytes, stride, lens, exp, (long)nl, r) < 0) { free(r); napi_throw_error(env, NULL, wam_last_error()); return NULL; }
  napi_value out = make_typed(env, napi_int32_array, sizeof(int32_t), r, nl * (sizeof(*r) / sizeof(int32_t)));
  free(r);
  return out;
}

/* ---- session multiplexer (wam_fsk_mux_*) --------------------------------------------------------------------- */
typedef struct { wam_fsk_mux* m; long n_sessions; } mux_handle;
static void mux_finalize(napi_env env, void* data, void* hint) { (void)env; (void)hint; mux_handle* h = (mux_handle*)data; wam_fsk_mux_destroy(h->m); free(h); }

static napi_value MuxCreate(napi_env env, napi_callback_info info) {  /* (device, nSessions, cfgs, cfgIndex?, maxBlock) */
  napi_value argv[5], obj; wam_fsk_config cfgs[MAX_CFGS]; uint8_t pat[64 * MAX_CFGS]; size_t len;
  if (get_args(env, info, 5, argv) < 3) return NULL;
  int n_cfgs = read_configs(env, argv[2], cfgs, pat);
  if (n_cfgs < 0) { napi_throw_type_error(env, NULL, "bad FSKConfig list"); return NULL; }
  mux_handle* h = (mux_handle*)calloc(1, sizeof(*h));
  h->n_sessions = get_long(env, argv[1], 0);
  const int32_t* idx = (const int32_t*)typed(env, argv[3], napi_int32_array, &len);
  if (wam_fsk_mux_create((int)get_long(env, argv[0], 0), h->n_sessions, cfgs, n_cfgs, idx, get_long(env, argv[4], 1024), &h->m) < 0) { free(h); napi_throw_error(env, NULL, wam_last_error()); return NULL; }
  NAPI_OK(napi_create_object(env, &obj));
  NAPI_OK(napi_wrap(env, obj, h, mux_finalize, NULL, NULL));
  return obj;
}
static napi_value MuxPush(napi_env env, napi_callback_info info) {  /* (h, session, Float32Array quantum) — process(), fsk-processor.ts:152-167 */
  napi_value argv[3]; mux_handle* h; size_t len;
  if (get_args(env, info, 3, argv) < 3) return NULL;
  NAPI_OK(napi_unwrap(env, argv[0], (void**)&h));
  const float* q = (const float*)typed(env, argv[2], napi_float32_array, &len);
  WAM_OK_OR_THROW(wam_fsk_mux_push(h->m, get_long(env, argv[1], -1), q, (long)len));
  return NULL;
}
static void mux_flush_run(job* j) {
  mux_handle* h = (mux_handle*)j->h;
  long cap = wam_fsk_mux_out_capacity(h->m);
  if (cap < 0) { j->rc = (int)cap; return; }
  j->n_out[0] = cap;
  j->out[0] = calloc((size_t)h->n_sessions, (size_t)(cap > 0 ? cap : 1));
  j->out[1] = calloc((size_t)h->n_sessions, sizeof(int32_t));
  j->rc = wam_fsk_mux_flush(h->m, (uint8_t*)j->out[0], cap, (int32_t*)j->out[1]);
}
static napi_value mux_rows_result(napi_env env, job* j) {
  mux_handle* h = (mux_handle*)j->h; napi_value o;
  if (napi_create_object(env, &o) != napi_ok) return NULL;
  napi_set_named_property(env, o, "bytes", make_typed(env, napi_uint8_array, 1, j->out[0], (size_t)(h->n_sessions * j->n_out[0])));
  napi_set_named_property(env, o, "lengths", make_typed(env, napi_int32_array, sizeof(int32_t), j->out[1], (size_t)h->n_sessions));
  set_num(env, o, "stride", (double)j->n_out[0]);
  return o;
}
static napi_value MuxFlush(napi_env env, napi_callback_info info) {
  napi_value argv[1]; job* j = job_new();
  if (get_args(env, info, 1, argv) < 1) return NULL;
  NAPI_OK(napi_unwrap(env, argv[0], &j->h));
  j->run = mux_flush_run; j->result = mux_rows_result;
  return job_start(env, j, "wam_fsk_mux_flush");
}
static napi_value MuxSend(napi_env env, napi_callback_info info) {  /* (h, session, Uint8Array) — startModulation(), chunked-modulator.ts:31-39 */
  napi_value argv[3]; mux_handle* h; size_t len;
  if (get_args(env, info, 3, argv) < 3) return NULL;
  NAPI_OK(napi_unwrap(env, argv[0], (void**)&h));
  const uint8_t* d = (const uint8_t*)typed(env, argv[2], napi_uint8_array, &len);
  WAM_OK_OR_THROW(wam_fsk_mux_send(h->m, get_long(env, argv[1], -1), d, (long)len));
  return NULL;
}
static void mux_modulate_run(job* j) { j->rc = wam_fsk_mux_modulate(((mux_handle*)j->h)->m); }
static napi_value undefined_result(napi_env env, job* j) { napi_value u; (void)j; return napi_get_undefined(env, &u) == napi_ok ? u : NULL; }
static napi_value MuxModulate(napi_env env, napi_callback_info info) {  /* one batched modulateData() for every queued session */
  napi_value argv[1]; job* j = job_new();
  if (get_args(env, info, 1, argv) < 1) return NULL;
  NAPI_OK(napi_unwrap(env, argv[0], &j->h));
  j->run = mux_modulate_run; j->result = undefined_result;
  return job_start(env, j, "wam_fsk_mux_modulate");
}
static napi_value MuxPull(napi_env env, napi_callback_info info) {  /* (h, session, sampleCount) — getNextSamples(), chunked-modulator.ts:41-81 */
  napi_value argv[3], o, b; mux_handle* h; wam_chunk_result res;
  if (get_args(env, info, 3, argv) < 3) return NULL;
  NAPI_OK(napi_unwrap(env, argv[0], (void**)&h));
  long want = get_long(env, argv[2], 128);
  float* buf = (float*)malloc(sizeof(float) * (size_t)(want > 0 ? want : 1));
  int rc = wam_fsk_mux_pull(h->m, get_long(env, argv[1], -1), buf, want, &res);
  if (rc < 0) { free(buf); napi_throw_error(env, NULL, wam_last_error()); return NULL; }
  if (rc == 0) { free(buf); NAPI_OK(napi_get_null(env, &o)); return o; }  /* not modulating: null */
  NAPI_OK(napi_create_object(env, &o));
  napi_set_named_property(env, o, "signal", make_typed(env, napi_float32_array, sizeof(float), buf, (size_t)res.samples));
  free(buf);
  if (napi_get_boolean(env, res.isComplete != 0, &b) == napi_ok) napi_set_named_property(env, o, "isComplete", b);
  set_num(env, o, "samplesConsumed", (double)res.samplesConsumed); set_num(env, o, "totalSamples", (double)res.totalSamples);
  return o;
}

/* ---- batched XModem receiver (wam_xmodem_batch_receive; xmodem.ts:232-321) ----------------------------------- */
typedef struct { int device; long n; int max_retries; wam_xmodem_rx_state* st; uint8_t* data; long data_stride; } xrx_handle;
static void xrx_finalize(napi_env env, void* data, void* hint) { (void)env; (void)hint; xrx_handle* h = (xrx_handle*)data; free(h->st); free(h->data); free(h); }

static napi_value XmodemReceiverCreate(napi_env env, napi_callback_info info) {  /* (device, nSessions, maxRetries, maxDataBytes?) */
  napi_value argv[4], obj;
  if (get_args(env, info, 4, argv) < 2) return NULL;
  xrx_handle* h = (xrx_handle*)calloc(1, sizeof(*h));
  h->device = (int)get_long(env, argv[0], 0); h->n = get_long(env, argv[1], 0); h->max_retries = (int)get_long(env, argv[2], 10);
  h->data_stride = get_long(env, argv[3], 65536);
  h->st = (wam_xmodem_rx_state*)calloc((size_t)(h->n > 0 ? h->n : 1), sizeof(*h->st));
  h->data = (uint8_t*)calloc((size_t)(h->n > 0 ? h->n : 1), (size_t)h->data_stride);
  for (long s = 0; s < h->n; s++) h->st[s].expectedSequence = 1;
  NAPI_OK(napi_create_object(env, &obj));
  NAPI_OK(napi_wrap(env, obj, h, xrx_finalize, NULL, NULL));
  return obj;
}
static void xrx_feed_run(job* j) {  /* p: 0 bytes [n][stride], 1 lengths; n: 0 stride */
  xrx_handle* h = (xrx_handle*)j->h;
  const int reply_cap = 64;
  j->n_out[0] = reply_cap;
  j->out[0] = calloc((size_t)h->n, (size_t)reply_cap);      /* replies */
  j->out[1] = calloc((size_t)h->n, sizeof(int32_t));        /* n_replies */
  j->out[2] = calloc((size_t)h->n, sizeof(int32_t));        /* consumed */
  j->rc = wam_xmodem_batch_receive(h->device, (const uint8_t*)j->p[0], j->n[0], (const int32_t*)j->p[1], h->n, h->max_retries, h->st,
                                   (uint8_t*)j->out[0], reply_cap, (int32_t*)j->out[1], (int32_t*)j->out[2], h->data, h->data_stride);
}
static napi_value xrx_feed_result(napi_env env, job* j) {  /* {replies, replyCounts, replyStride, consumed, done} */
  xrx_handle* h = (xrx_handle*)j->h; napi_value o;
  if (napi_create_object(env, &o) != napi_ok) return NULL;
  napi_set_named_property(env, o, "replies", make_typed(env, napi_uint8_array, 1, j->out[0], (size_t)(h->n * j->n_out[0])));
  napi_set_named_property(env, o, "replyCounts", make_typed(env, napi_int32_array, sizeof(int32_t), j->out[1], (size_t)h->n));
  napi_set_named_property(env, o, "consumed", make_typed(env, napi_int32_array, sizeof(int32_t), j->out[2], (size_t)h->n));
  set_num(env, o, "replyStride", (double)j->n_out[0]);
  int32_t* done = (int32_t*)malloc(sizeof(int32_t) * (size_t)(h->n > 0 ? h->n : 1));
  for (long s = 0; s < h->n; s++) done[s] = h->st[s].done;
  napi_set_named_property(env, o, "done", make_typed(env, napi_int32_array, sizeof(int32_t), done, (size_t)h->n));
  free(done);
  return o;
}
static napi_value XmodemReceiverFeed(napi_env env, napi_callback_info info) {  /* (h, Uint8Array [n][stride], stride, Int32Array lengths) */
  napi_value argv[4]; size_t nb, nl; job* j = job_new();
  if (get_args(env, info, 4, argv) < 4) return NULL;
  NAPI_OK(napi_unwrap(env, argv[0], &j->h));
  j->p[0] = typed(env, argv[1], napi_uint8_array, &nb); j->n[0] = get_long(env, argv[2], 0);
  j->p[1] = typed(env, argv[3], napi_int32_array, &nl);
  if ((long)nl != ((xrx_handle*)j->h)->n || (long)nb < j->n[0] * (long)nl) { free(j); napi_throw_type_error(env, NULL, "bad burst buffers"); return NULL; }
  j->run = xrx_feed_run; j->result = xrx_feed_result;
  if (job_keep(env, j, argv[1]) != 0 || job_keep(env, j, argv[3]) != 0) return NULL;
  return job_start(env, j, "wam_xmodem_batch_receive");
}
static napi_value XmodemReceiverData(napi_env env, napi_callback_info info) {  /* (h, session) -> assembleData(receive.data), xmodem.ts:323-334 */
  napi_value argv[2]; xrx_handle* h;
  if (get_args(env, info, 2, argv) < 2) return NULL;
  NAPI_OK(napi_unwrap(env, argv[0], (void**)&h));
  long s = get_long(env, argv[1], -1);
  if (s < 0 || s >= h->n) { napi_throw_range_error(env, NULL, "no such session"); return NULL; }
  long n = h->st[s].dataLen < h->data_stride ? h->st[s].dataLen : h->data_stride;
  return make_typed(env, napi_uint8_array, 1, h->data + s * h->data_stride, (size_t)n);
}

static napi_value Init(napi_env env, napi_value exports) {
  napi_property_descriptor props[] = {
      {"fskCreate", NULL, FskCreate, NULL, NULL, NULL, napi_default, NULL},
      {"fskConfigure", NULL, FskConfigure, NULL, NULL, NULL, napi_default, NULL},
      {"fskModulate", NULL, FskModulate, NULL, NULL, NULL, napi_default, NULL},
      {"fskDemodulate", NULL, FskDemodulate, NULL, NULL, NULL, napi_default, NULL},
      {"fskReset", NULL, FskReset, NULL, NULL, NULL, napi_default, NULL},
      {"fskStatus", NULL, FskStatus, NULL, NULL, NULL, napi_default, NULL},
      {"batchCreate", NULL, BatchCreate, NULL, NULL, NULL, napi_default, NULL},
      {"batchDemodulate", NULL, BatchDemodulate, NULL, NULL, NULL, napi_default, NULL},
      {"batchModulate", NULL, BatchModulate, NULL, NULL, NULL, napi_default, NULL},
      {"batchStatus", NULL, BatchStatus, NULL, NULL, NULL, napi_default, NULL},
      {"xmodemBatchCheck", NULL, XmodemBatchCheck, NULL, NULL, NULL, napi_default, NULL},
      {"muxCreate", NULL, MuxCreate, NULL, NULL, NULL, napi_default, NULL},
      {"muxPush", NULL, MuxPush, NULL, NULL, NULL, napi_default, NULL},
      {"muxFlush", NULL, MuxFlush, NULL, NULL, NULL, napi_default, NULL},
      {"muxSend", NULL, MuxSend, NULL, NULL, NULL, napi_default, NULL},
      {"muxModulate", NULL, MuxModulate, NULL, NULL, NULL, napi_default, NULL},
      {"muxPull", NULL, MuxPull, NULL, NULL, NULL, napi_default, NULL},
      {"xmodemReceiverCreate", NULL, XmodemReceiverCreate, NULL, NULL, NULL, napi_default, NULL},
      {"xmodemReceiverFeed", NULL, XmodemReceiverFeed, NULL, NULL, NULL, napi_default, NULL},
      {"xmodemReceiverData", NULL, XmodemReceiverData, NULL, NULL, NULL, napi_default, NULL},
  };
  napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
  return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
