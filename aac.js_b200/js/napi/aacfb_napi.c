/*
 * aacfb_napi.c -- N-API addon: the thin JS <-> C-ABI seam of include/aacfb.h.
 *
 * One JS function per C entry point, typed arrays are borrowed for the
 * duration of the call (no copies, no retained references), library errors
 * become `throw new Error(msg)` -- the reference's only error convention
 * (filter_bank.js:26, mdct.js:49, fft.js:42, tns.js:85).
 *
 * NOT BUILT IN THIS REPOSITORY'S IMAGE: no node / node_api.h exists there.
 * The same C entry points are exercised by the Python mirror (ctypes), which
 * is the tested surface.  Build on a Node machine with binding.gyp.
 */
#include <node_api.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../../include/aacfb.h"

#define NAPI_OK(env, call) do { if ((call) != napi_ok) { napi_throw_error((env), NULL, "N-API call failed: " #call); return NULL; } } while (0)

static napi_value fail(napi_env env, aacfb_ctx *ctx, int rc) {
    const char *msg = aacfb_last_error(ctx);
    napi_throw_error(env, NULL, (msg && *msg) ? msg : "aacfb error");
    (void)rc;
    return NULL;
}

static void *typed(napi_env env, napi_value v, size_t *bytes) { /* NULL for null/undefined */
    napi_valuetype t;
    if (napi_typeof(env, v, &t) != napi_ok || t == napi_null || t == napi_undefined) { if (bytes) *bytes = 0; return NULL; }
    napi_typedarray_type ty; size_t len; void *data; napi_value ab; size_t off;
    if (napi_get_typedarray_info(env, v, &ty, &len, &data, &ab, &off) != napi_ok) { if (bytes) *bytes = 0; return NULL; }
    static const size_t width[] = {1, 1, 1, 2, 2, 4, 4, 4, 8, 8, 8};
    if (bytes) *bytes = len * width[ty];
    return data;
}

/* The handle JS holds: the context plus the shape every buffer size is checked against. */
typedef struct { aacfb_ctx *ctx; int S, C; } handle_t;

static handle_t *handle(napi_env env, napi_value v) {
    void *p = NULL;
    if (napi_get_value_external(env, v, &p) != napi_ok || !p) { napi_throw_error(env, NULL, "not an aacfb handle"); return NULL; }
    return (handle_t *)p;
}

static void finalize(napi_env env, void *data, void *hint) {
    (void)env; (void)hint;
    handle_t *h = (handle_t *)data;
    aacfb_destroy(h->ctx);
    free(h);
}

/* A typed array that must hold at least `need` bytes (NULL allowed when `optional`).  A short array from JS
 * would otherwise become an out-of-bounds read or write inside cudaMemcpy. */
static int sized(napi_env env, napi_value v, size_t need, int optional, const char *what, void **out) {
    size_t have = 0;
    *out = typed(env, v, &have);
    if (!*out) {
        if (optional) return 1;
        napi_throw_error(env, NULL, what);
        return 0;
    }
    if (have < need) { napi_throw_error(env, NULL, what); return 0; }
    return 1;
}

/* create(device, nStreams, channels, sampleIndex, smallFrames, flags) -> handle */
static napi_value js_create(napi_env env, napi_callback_info info) {
    size_t argc = 6; napi_value a[6]; int32_t v[6] = {0, 1, 2, 4, 0, 0};
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    for (size_t i = 0; i < argc && i < 6; ++i) napi_get_value_int32(env, a[i], &v[i]);
    aacfb_ctx *ctx = NULL;
    int rc = aacfb_create(&ctx, v[0], v[1], v[2], v[3], v[4], (uint32_t)v[5]);
    if (rc != AACFB_OK) return fail(env, NULL, rc);   /* e.g. "WHA?? No small frames allowed." */
    handle_t *h = (handle_t *)malloc(sizeof *h);
    if (!h) { aacfb_destroy(ctx); napi_throw_error(env, NULL, "out of memory"); return NULL; }
    h->ctx = ctx; h->S = v[1]; h->C = v[2];
    napi_value ext;
    NAPI_OK(env, napi_create_external(env, h, finalize, NULL, &ext));
    return ext;
}

/* The one batched call everything else is a special case of:
 * processIo(handle, input (Float32Array spectra | Uint8Array of aacfb_qframe records), inFormat 0|1, info u8,
 *           stereoOps u8|null, tnsBlob u8|null, tnsOffsets u32|null, pcm (Float32Array | Int16Array), pcmFormat 0|1, nFrames) */
static napi_value process_any(napi_env env, handle_t *h, napi_value input, uint32_t in_format, napi_value info,
                              napi_value stereo, napi_value blob, napi_value offsets, napi_value pcm, uint32_t pcm_format,
                              int32_t n) {
    if (!h) return NULL;
    if (n < 0 || in_format > AACFB_IN_Q16 || pcm_format > AACFB_PCM_S16) { napi_throw_error(env, NULL, "bad frame count / format"); return NULL; }
    const size_t n_cf = (size_t)h->S * (size_t)n * (size_t)h->C;
    void *in_p, *info_p, *st_p, *blob_p, *off_p, *pcm_p;
    if (!sized(env, input, n_cf * (in_format == AACFB_IN_Q16 ? sizeof(aacfb_qframe) : 4096), 0, "input array too small for nFrames", &in_p)) return NULL;
    if (!sized(env, info, n_cf * sizeof(aacfb_frame_info), 0, "info array too small for nFrames", &info_p)) return NULL;
    if (!sized(env, stereo, n_cf / 2 * sizeof(aacfb_stereo_ops), 1, "stereo-ops array too small for nFrames", &st_p)) return NULL;
    if (!sized(env, offsets, (n_cf + 1) * sizeof(uint32_t), 1, "TNS offsets array too small for nFrames", &off_p)) return NULL;
    const size_t blob_need = off_p ? ((const uint32_t *)off_p)[n_cf] : 0;
    if (!sized(env, blob, blob_need, 1, "TNS blob shorter than its offsets say", &blob_p)) return NULL;
    if (!sized(env, pcm, n_cf * (pcm_format == AACFB_PCM_S16 ? 2048 : 4096), 0, "pcm array too small for nFrames", &pcm_p)) return NULL;
    int rc = aacfb_process_io(h->ctx, in_p, in_format, (const aacfb_frame_info *)info_p, (const aacfb_stereo_ops *)st_p,
                              (const uint8_t *)blob_p, (const uint32_t *)off_p, pcm_p, pcm_format, n);
    if (rc != AACFB_OK) return fail(env, h->ctx, rc);
    return NULL;
}

static napi_value js_process_io(napi_env env, napi_callback_info info) {
    size_t argc = 10; napi_value a[10];
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    uint32_t inf = 0, outf = 0; int32_t n = 0;
    napi_get_value_uint32(env, a[2], &inf); napi_get_value_uint32(env, a[8], &outf); napi_get_value_int32(env, a[9], &n);
    return process_any(env, handle(env, a[0]), a[1], inf, a[3], a[4], a[5], a[6], a[7], outf, n);
}

/* allocPinned(bytes) -> ArrayBuffer on page-locked memory (aacfb_host_alloc), freed with the buffer: the
 * decoder's staging typed arrays live here, so the library's copies run at the PCIe rate. */
static void free_pinned(napi_env env, void *data, void *hint) { (void)env; (void)hint; aacfb_host_free(data); }
static napi_value js_alloc_pinned(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value a[1];
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    uint32_t bytes = 0; napi_get_value_uint32(env, a[0], &bytes);
    void *p = aacfb_host_alloc(bytes ? bytes : 16);
    if (!p) return fail(env, NULL, AACFB_ERR_CUDA);
    memset(p, 0, bytes ? bytes : 16);
    napi_value ab;
    if (napi_create_external_arraybuffer(env, p, bytes, free_pinned, NULL, &ab) != napi_ok) {
        aacfb_host_free(p);
        napi_throw_error(env, NULL, "external ArrayBuffers are not available in this runtime");
        return NULL;
    }
    return ab;
}

/* process(handle, spectra f32, info u8 (8 B per channel-frame), tnsBlob u8|null, tnsOffsets u32|null, pcm f32, nFrames) */
static napi_value js_process(napi_env env, napi_callback_info info) {
    size_t argc = 7; napi_value a[7];
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    int32_t n = 0; napi_get_value_int32(env, a[6], &n);
    napi_value none; napi_get_null(env, &none);
    return process_any(env, handle(env, a[0]), a[1], AACFB_IN_F32, a[2], none, a[3], a[4], a[5], AACFB_PCM_F32, n);
}

/* processStereo(handle, spectra f32, info u8, stereoOps u8 (768 B per pair-frame), tnsBlob, tnsOffsets, pcm f32, nFrames):
 * spectra are ics.data BEFORE processMS / processIS (decoder.js:294-301); see js/stereo_pack.js */
static napi_value js_process_stereo(napi_env env, napi_callback_info info) {
    size_t argc = 8; napi_value a[8];
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    int32_t n = 0; napi_get_value_int32(env, a[7], &n);
    return process_any(env, handle(env, a[0]), a[1], AACFB_IN_F32, a[2], a[3], a[4], a[5], a[6], AACFB_PCM_F32, n);
}

/* swbOffsets(sampleIndex, isShort) -> Uint16Array copy of info.swbOffsets (tables.js:126-154); the JS host
 * has the reference's own tables, this is for hosts that do not */
static napi_value js_swb_offsets(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value a[2];
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    int32_t si = 0, sh = 0; napi_get_value_int32(env, a[0], &si); napi_get_value_int32(env, a[1], &sh);
    uint16_t tmp[64];
    int n = aacfb_get_swb_offsets(si, sh, tmp, 64);
    if (n < 0) return fail(env, NULL, n);
    void *data; napi_value ab, out;
    NAPI_OK(env, napi_create_arraybuffer(env, sizeof(uint16_t) * (size_t)(n + 1), &data, &ab));
    memcpy(data, tmp, sizeof(uint16_t) * (size_t)(n + 1));
    NAPI_OK(env, napi_create_typedarray(env, napi_uint16_array, (size_t)(n + 1), ab, 0, &out));
    return out;
}

/* adtsIndex(bytes u8, out u32 (3 words per frame: offset, frameLength, headerBytes)) -> number of complete
 * frames found; out[3n] receives the offset of the first incomplete frame (the rewind point).
 * ADTSDemuxer.readHeader (adts_demuxer.js:28-52) for every access unit of the buffer at once. */
static napi_value js_adts_index(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value a[2];
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    size_t nbytes = 0, obytes = 0;
    const uint8_t *data = (const uint8_t *)typed(env, a[0], &nbytes);
    uint32_t *out = (uint32_t *)typed(env, a[1], &obytes);
    int cap = (int)(obytes / 12);
    if (cap > 0) cap -= 1;                     /* the last triple holds the rewind point */
    aacfb_adts_frame fr[256];
    size_t pos = 0, consumed = 0;
    int total = 0;
    while (total < cap) {
        int want = cap - total < 256 ? cap - total : 256;
        int n = aacfb_adts_index(data + pos, nbytes - pos, fr, want, &consumed);
        if (n < 0) return fail(env, NULL, n);
        for (int i = 0; i < n; i++) {
            out[3 * (total + i)] = (uint32_t)(pos + fr[i].offset);
            out[3 * (total + i) + 1] = fr[i].frame_length;
            out[3 * (total + i) + 2] = fr[i].header_bytes;
        }
        total += n; pos += consumed;
        if (n < want) break;
    }
    if (obytes >= 12) out[3 * total] = (uint32_t)pos;
    napi_value r;
    NAPI_OK(env, napi_create_int32(env, total, &r));
    return r;
}

/* filterbankProcess(handle, stream, channel, info u8[8], input f32[1024], output f32[1024]) */
static napi_value js_filterbank(napi_env env, napi_callback_info info) {
    size_t argc = 6; napi_value a[6];
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    handle_t *h = handle(env, a[0]);
    if (!h) return NULL;
    int32_t s = 0, c = 0; napi_get_value_int32(env, a[1], &s); napi_get_value_int32(env, a[2], &c);
    void *fi, *in, *out;
    if (!sized(env, a[3], sizeof(aacfb_frame_info), 0, "info must hold 8 bytes", &fi)) return NULL;
    if (!sized(env, a[4], 4096, 0, "input must hold 1024 floats", &in) || !sized(env, a[5], 4096, 0, "output must hold 1024 floats", &out)) return NULL;
    int rc = aacfb_filterbank_process(h->ctx, s, c, (const aacfb_frame_info *)fi, (const float *)in, (float *)out);
    if (rc != AACFB_OK) return fail(env, h->ctx, rc);
    return NULL;
}

/* tnsProcess(handle, info u8[8], block u8, data f32[1024], mode) */
static napi_value js_tns(napi_env env, napi_callback_info info) {
    size_t argc = 5; napi_value a[5];
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    handle_t *h = handle(env, a[0]);
    if (!h) return NULL;
    size_t nb = 0; const uint8_t *blk = (const uint8_t *)typed(env, a[2], &nb);
    uint32_t mode = 0; napi_get_value_uint32(env, a[4], &mode);
    void *fi, *data;
    if (!sized(env, a[1], sizeof(aacfb_frame_info), 0, "info must hold 8 bytes", &fi)) return NULL;
    if (!sized(env, a[3], 4096, 0, "data must hold 1024 floats", &data)) return NULL;
    int rc = aacfb_tns_process(h->ctx, (const aacfb_frame_info *)fi, blk, nb, (float *)data, mode);
    if (rc != AACFB_OK) return fail(env, h->ctx, rc);
    return NULL;
}

static napi_value js_reset(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value a[1];
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    handle_t *h = handle(env, a[0]);
    if (!h) return NULL;
    int rc = aacfb_reset(h->ctx);
    if (rc != AACFB_OK) return fail(env, h->ctx, rc);
    return NULL;
}

/* getOverlap(handle, f32[S*C*1024]) / setOverlap(handle, f32[...]) */
static napi_value js_get_overlap(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value a[2];
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    handle_t *h = handle(env, a[0]);
    if (!h) return NULL;
    void *ov;
    if (!sized(env, a[1], (size_t)h->S * h->C * 4096, 0, "overlap array must hold S*C*1024 floats", &ov)) return NULL;
    int rc = aacfb_get_overlap(h->ctx, (float *)ov);
    if (rc != AACFB_OK) return fail(env, h->ctx, rc);
    return NULL;
}
static napi_value js_set_overlap(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value a[2];
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    handle_t *h = handle(env, a[0]);
    if (!h) return NULL;
    void *ov;
    if (!sized(env, a[1], (size_t)h->S * h->C * 4096, 0, "overlap array must hold S*C*1024 floats", &ov)) return NULL;
    int rc = aacfb_set_overlap(h->ctx, (const float *)ov);
    if (rc != AACFB_OK) return fail(env, h->ctx, rc);
    return NULL;
}

static napi_value init(napi_env env, napi_value exports) {
    const napi_property_descriptor props[] = {
        {"create", NULL, js_create, NULL, NULL, NULL, napi_default, NULL},
        {"process", NULL, js_process, NULL, NULL, NULL, napi_default, NULL},
        {"processStereo", NULL, js_process_stereo, NULL, NULL, NULL, napi_default, NULL},
        {"processIo", NULL, js_process_io, NULL, NULL, NULL, napi_default, NULL},
        {"allocPinned", NULL, js_alloc_pinned, NULL, NULL, NULL, napi_default, NULL},
        {"swbOffsets", NULL, js_swb_offsets, NULL, NULL, NULL, napi_default, NULL},
        {"adtsIndex", NULL, js_adts_index, NULL, NULL, NULL, napi_default, NULL},
        {"filterbankProcess", NULL, js_filterbank, NULL, NULL, NULL, napi_default, NULL},
        {"tnsProcess", NULL, js_tns, NULL, NULL, NULL, napi_default, NULL},
        {"reset", NULL, js_reset, NULL, NULL, NULL, napi_default, NULL},
        {"getOverlap", NULL, js_get_overlap, NULL, NULL, NULL, napi_default, NULL},
        {"setOverlap", NULL, js_set_overlap, NULL, NULL, NULL, napi_default, NULL},
    };
    napi_define_properties(env, exports, sizeof props / sizeof props[0], props);
    return exports;
}
NAPI_MODULE(aacfb, init)
