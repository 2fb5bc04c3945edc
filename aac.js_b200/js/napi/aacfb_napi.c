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

static aacfb_ctx *handle(napi_env env, napi_value v) {
    void *p = NULL;
    napi_get_value_external(env, v, &p);
    return (aacfb_ctx *)p;
}

static void finalize(napi_env env, void *data, void *hint) { (void)env; (void)hint; aacfb_destroy((aacfb_ctx *)data); }

/* create(device, nStreams, channels, sampleIndex, smallFrames, flags) -> handle */
static napi_value js_create(napi_env env, napi_callback_info info) {
    size_t argc = 6; napi_value a[6]; int32_t v[6] = {0, 1, 2, 4, 0, 0};
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    for (size_t i = 0; i < argc && i < 6; ++i) napi_get_value_int32(env, a[i], &v[i]);
    aacfb_ctx *ctx = NULL;
    int rc = aacfb_create(&ctx, v[0], v[1], v[2], v[3], v[4], (uint32_t)v[5]);
    if (rc != AACFB_OK) return fail(env, NULL, rc);   /* e.g. "WHA?? No small frames allowed." */
    napi_value ext;
    NAPI_OK(env, napi_create_external(env, ctx, finalize, NULL, &ext));
    return ext;
}

/* process(handle, spectra f32, info u8 (8 B per channel-frame), tnsBlob u8|null, tnsOffsets u32|null, pcm f32, nFrames) */
static napi_value js_process(napi_env env, napi_callback_info info) {
    size_t argc = 7; napi_value a[7];
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    aacfb_ctx *ctx = handle(env, a[0]);
    int32_t n = 0; napi_get_value_int32(env, a[6], &n);
    int rc = aacfb_process(ctx, (const float *)typed(env, a[1], NULL), (const aacfb_frame_info *)typed(env, a[2], NULL),
                           (const uint8_t *)typed(env, a[3], NULL), (const uint32_t *)typed(env, a[4], NULL),
                           (float *)typed(env, a[5], NULL), n);
    if (rc != AACFB_OK) return fail(env, ctx, rc);
    return NULL;
}

/* processStereo(handle, spectra f32, info u8, stereoOps u8 (768 B per pair-frame), tnsBlob, tnsOffsets, pcm f32, nFrames):
 * spectra are ics.data BEFORE processMS / processIS (decoder.js:294-301); see js/stereo_pack.js */
static napi_value js_process_stereo(napi_env env, napi_callback_info info) {
    size_t argc = 8; napi_value a[8];
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    aacfb_ctx *ctx = handle(env, a[0]);
    int32_t n = 0; napi_get_value_int32(env, a[7], &n);
    int rc = aacfb_process_stereo(ctx, (const float *)typed(env, a[1], NULL), (const aacfb_frame_info *)typed(env, a[2], NULL),
                                  (const aacfb_stereo_ops *)typed(env, a[3], NULL), (const uint8_t *)typed(env, a[4], NULL),
                                  (const uint32_t *)typed(env, a[5], NULL), (float *)typed(env, a[6], NULL), n);
    if (rc != AACFB_OK) return fail(env, ctx, rc);
    return NULL;
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
    aacfb_ctx *ctx = handle(env, a[0]);
    int32_t s = 0, c = 0; napi_get_value_int32(env, a[1], &s); napi_get_value_int32(env, a[2], &c);
    int rc = aacfb_filterbank_process(ctx, s, c, (const aacfb_frame_info *)typed(env, a[3], NULL),
                                      (const float *)typed(env, a[4], NULL), (float *)typed(env, a[5], NULL));
    if (rc != AACFB_OK) return fail(env, ctx, rc);
    return NULL;
}

/* tnsProcess(handle, info u8[8], block u8, data f32[1024], mode) */
static napi_value js_tns(napi_env env, napi_callback_info info) {
    size_t argc = 5; napi_value a[5];
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    aacfb_ctx *ctx = handle(env, a[0]);
    size_t nb = 0; const uint8_t *blk = (const uint8_t *)typed(env, a[2], &nb);
    uint32_t mode = 0; napi_get_value_uint32(env, a[4], &mode);
    int rc = aacfb_tns_process(ctx, (const aacfb_frame_info *)typed(env, a[1], NULL), blk, nb, (float *)typed(env, a[3], NULL), mode);
    if (rc != AACFB_OK) return fail(env, ctx, rc);
    return NULL;
}

static napi_value js_reset(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value a[1];
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    aacfb_ctx *ctx = handle(env, a[0]);
    int rc = aacfb_reset(ctx);
    if (rc != AACFB_OK) return fail(env, ctx, rc);
    return NULL;
}

/* getOverlap(handle, f32[S*C*1024]) / setOverlap(handle, f32[...]) */
static napi_value js_get_overlap(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value a[2];
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    aacfb_ctx *ctx = handle(env, a[0]);
    int rc = aacfb_get_overlap(ctx, (float *)typed(env, a[1], NULL));
    if (rc != AACFB_OK) return fail(env, ctx, rc);
    return NULL;
}
static napi_value js_set_overlap(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value a[2];
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    aacfb_ctx *ctx = handle(env, a[0]);
    int rc = aacfb_set_overlap(ctx, (const float *)typed(env, a[1], NULL));
    if (rc != AACFB_OK) return fail(env, ctx, rc);
    return NULL;
}

static napi_value init(napi_env env, napi_value exports) {
    const napi_property_descriptor props[] = {
        {"create", NULL, js_create, NULL, NULL, NULL, napi_default, NULL},
        {"process", NULL, js_process, NULL, NULL, NULL, napi_default, NULL},
        {"processStereo", NULL, js_process_stereo, NULL, NULL, NULL, napi_default, NULL},
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
