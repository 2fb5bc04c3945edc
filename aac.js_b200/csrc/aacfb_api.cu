// aacfb_api.cu -- the C ABI of include/aacfb.h on top of the sm_100a kernels.
//
// No CPU fallback lives here: every compute entry point needs a CUDA device
// that can run the sm_100a image and fails with AACFB_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/aacfb.h"
#include "aacfb_kernels.h"
#include "aacfb_tables.h"

using namespace aacfb;

namespace {

constexpr int kLanes = 4;        // host-path pipeline: sub-batches in flight (H2D | kernel | D2H overlap)
constexpr int kCounters = 64;

thread_local char g_err[256] = "";

struct Lane {                    // device staging of one in-flight sub-batch of the host path
    cudaStream_t stream = nullptr;
    // copy streams (aacfb_process_io): inputs landed / kernels done (inputs free, PCM ready) / PCM copied out
    cudaEvent_t ev_in = nullptr, ev_k = nullptr, ev_out = nullptr;
    float *d_spectra = nullptr, *d_pcm = nullptr, *d_scratch = nullptr;
    aacfb_frame_info *d_info = nullptr;
    uint32_t *d_offsets = nullptr;
    aacfb_stereo_ops *d_stereo = nullptr;  // [cap_stereo] records of the sub-batch's pair-frames
    float *d_stereo_out = nullptr;         // output of the stereo pre-pass (TNS modes only)
    float *d_deq = nullptr;                // output of the inverse-quantisation pre-pass (Q16 input + TNS modes only)
    size_t cap_cf = 0;           // capacity in channel-frames
    size_t cap_scratch = 0;
    size_t cap_stereo = 0, cap_stereo_out = 0, cap_deq = 0;
    // scratch holds cap_scratch rows followed by cap_scratch range words (see tns_kernel)
    uint32_t *ranges() const { return reinterpret_cast<uint32_t *>(d_scratch + cap_scratch * 1024); }
};

}  // namespace

struct aacfb_ctx {
    int device = 0, S = 0, C = 0, sample_index = 0, num_sms = 0;
    uint32_t flags = 0;
    float *d_ovl[2] = {nullptr, nullptr};
    int cur = 0;
    SynthTables *d_tab = nullptr;       // windows carry the 2^-15 output scale (decoder.js:210)
    SynthTables *d_tab_unit = nullptr;  // unscaled windows, for the inner seam (FilterBank.process output)
    TnsBandTables *d_bands = nullptr;
    DequantTables *d_dq = nullptr;      // inverse-quantisation tables of this sample rate (ics.js:203-266)
    unsigned *d_counters = nullptr;
    unsigned counter_next = 0;
    Lane lane[kLanes];
    cudaStream_t h2d = nullptr, d2h = nullptr;   // one stream per PCIe direction for the host path's copies
    uint8_t *d_blob = nullptr;
    size_t cap_blob = 0;
    float *d_dev_scratch = nullptr;  // scratch of the device-pointer path
    size_t cap_dev_scratch = 0;
    float *d_dev_stereo_out = nullptr;  // stereo pre-pass output of the device-pointer path
    size_t cap_dev_stereo_out = 0;
    float *d_dev_deq = nullptr;         // inverse-quantisation pre-pass output of the device-pointer path
    size_t cap_dev_deq = 0;
    float *d_tns_ring[kLanes + 1] = {};  // synth_tns_kernel: the CTAs' rings of filtered rows, one set per lane stream
                                         // (+ one for the device-pointer path), allocated on first use
    uint64_t launches = 0;
    char err[256] = "";
};

namespace {

int fail(aacfb_ctx *ctx, int code, const char *fmt, ...) {
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    std::snprintf(g_err, sizeof g_err, "%s", buf);
    if (ctx) std::snprintf(ctx->err, sizeof ctx->err, "%s", buf);
    return code;
}

#define CU(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(ctx, AACFB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),   \
                        __FILE__, __LINE__);                                                            \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Slice length: the flattened sequence of pair-frames is cut into equal slices, one per worker
// (frames cost the same, so equal static shares balance perfectly and cost one halo frame each);
// tiny batches use fewer, longer slices so that the halo stays below ~1/8 of the work.
// AACFB_SLICE_LEN overrides (tuning aid).
// exact: the number of items must not exceed `workers` (synth_tns_kernel: one item per client), so no override
int pick_slice(int n_pairs, int T, int workers, bool exact = false) {
    if (const char *env = exact ? nullptr : std::getenv("AACFB_SLICE_LEN")) {
        const int v = std::atoi(env);
        if (v >= 1) return v;
    }
    const long total = (long)n_pairs * T;
    long q = (total + workers - 1) / workers;
    if (q < 8) q = std::min<long>(8, total);
    return (int)std::max<long>(1, q);
}

bool stereo_needs_prepass(int nc, bool tns_on) { return tns_on || nc != 2; }

// What one enqueue works on (device pointers).
struct Job {
    const float *d_spectra = nullptr;         // float rows (AACFB_IN_F32) ...
    const aacfb_qframe *d_q = nullptr;        // ... or aacfb_qframe records (AACFB_IN_Q16)
    float *d_deq = nullptr;                   // Q16 + TNS modes: where the inverse-quantisation pre-pass writes
    const aacfb_frame_info *d_info = nullptr;
    const aacfb_stereo_ops *d_stereo = nullptr;
    float *d_stereo_out = nullptr;
    const uint8_t *d_blob = nullptr;
    const uint32_t *d_offsets = nullptr;
    size_t blob_bytes = 0;
    float *d_scratch = nullptr;
    uint32_t *d_ranges = nullptr;
    void *d_pcm = nullptr;
    bool s16 = false;                         // AACFB_PCM_S16
    int S_sub = 0, s_base = 0, T = 0, nc = 0, c0 = 0;
    float scale = 1.0f;                       // 2^-15 for float PCM (decoder.js:210), 1 for the inner seam and int16
    bool in_place_state = false;
    bool no_short = false;                    // the caller has seen every window_sequence: no EIGHT_SHORT in the batch
};

bool tns_active(const aacfb_ctx *ctx, const Job &j) {
    const uint32_t mode = ctx->flags & AACFB_TNS_MODE_MASK;
    return mode != AACFB_TNS_AS_SHIPPED && j.d_blob && j.d_offsets && j.blob_bytes > 0 && j.d_scratch;
}

// Enqueue the pre-passes the context's mode asks for and the synthesis kernel for S_sub streams
// starting at stream s_base, all on `stream`.
//
// Quantised input (d_q): synth_kernel dequantises the staged records itself (ics.js:203-266) unless a
// TNS pass has to see float rows first; then dequant_kernel writes them to d_deq.
// Stereo tools (d_stereo != nullptr): with two channels and no TNS pass, synth_kernel applies the
// ops to the staged rows (no extra traffic but the records).  When TNS has to run between the
// stereo tools and the IMDCT (decoder.js:300-319), or with more than one pair per stream, a
// pre-pass writes the processed spectra to d_stereo_out and everything downstream reads that.
int enqueue(aacfb_ctx *ctx, Job j, cudaStream_t stream) {
    const uint32_t mode = ctx->flags & AACFB_TNS_MODE_MASK;
    const size_t n_cf = (size_t)j.S_sub * j.T * j.nc;
    if ((long long)j.S_sub * j.T * j.nc >= (1ll << 30)) return fail(ctx, AACFB_ERR_ARG, "batch too large: split it (channel-frames < 2^30)");
    const bool tns_on = tns_active(ctx, j);
    const bool stereo_pre = j.d_stereo && stereo_needs_prepass(j.nc, tns_on);
    if (j.d_q && (tns_on || stereo_pre)) {
        if (!j.d_deq) return fail(ctx, AACFB_ERR_ARG, "internal: no buffer for the inverse-quantisation pre-pass");
        DequantParams dp{};
        dp.qframes = reinterpret_cast<const uint8_t *>(j.d_q); dp.info = j.d_info; dp.dq = ctx->d_dq; dp.out = j.d_deq; dp.n_cf = n_cf;
        CU(ctx, launch_dequant(dp, stream));
        ctx->launches++;
        j.d_spectra = j.d_deq;
        j.d_q = nullptr;
    }
    if (stereo_pre) {
        if (!j.d_stereo_out) return fail(ctx, AACFB_ERR_ARG, "internal: no buffer for the stereo pre-pass");
        StereoParams st{};
        st.spectra = j.d_spectra; st.out = j.d_stereo_out; st.info = j.d_info; st.stereo = j.d_stereo; st.n_pairs_frames = n_cf / 2;
        CU(ctx, launch_stereo(st, stream));
        ctx->launches++;
        j.d_spectra = j.d_stereo_out;
        j.d_stereo = nullptr;
    }
    // TNS inside the synthesis kernel (synth_tns_kernel) when the rows are floats of two-channel streams with float
    // PCM: no round trip of the filtered coefficients through HBM.  Items with EIGHT_SHORT frames are left to the
    // pre-pass + generic instantiation, which then run gated on the count the fused kernel leaves (a batch whose
    // window sequences the caller has seen -- no_short -- spares those launches).
    // OPT-IN (AACFB_TNS_FUSED=1): parity-green, but measured 2 x slower than pre-pass + synthesis on config 4
    // (0.82 ms against 0.41 ms: the one filtering worker per CTA needs ~8700 cycles per 32-coefficient tile where the
    // chain alone takes 1570 -- DESIGN.md section 8), so the pre-pass stays the default.
    static const bool fuse_env = [] { const char *e = std::getenv("AACFB_TNS_FUSED"); return e && std::atoi(e) != 0; }();
    const bool fused = tns_on && fuse_env && j.nc == 2 && !j.d_q && !j.d_stereo && !j.s16 && !j.in_place_state;
    int ring_i = kLanes;   // kernels of different lanes may overlap at their tails: each lane stream has its own rings
    for (int i = 0; i < kLanes; ++i)
        if (ctx->lane[i].stream == stream) ring_i = i;
    if (fused && !ctx->d_tns_ring[ring_i])
        CU(ctx, cudaMalloc(&ctx->d_tns_ring[ring_i], tns_ring_floats(ctx->num_sms) * sizeof(float)));
    unsigned *slots = ctx->d_counters + 4 * (ctx->counter_next++ % (kCounters / 4));
    TnsParams tp{};
    tp.spectra = j.d_spectra; tp.scratch = j.d_scratch; tp.ranges = j.d_ranges; tp.info = j.d_info; tp.blob = j.d_blob;
    tp.offsets = j.d_offsets;
    tp.blob_bytes = j.blob_bytes; tp.n_cf = n_cf; tp.sample_index = ctx->sample_index;
    tp.ar = mode == AACFB_TNS_FIXED_AR; tp.bands = ctx->d_bands;
    tp.gate = fused ? slots + 2 : nullptr;
    if (tns_on && !fused) {
        CU(ctx, launch_tns(tp, stream));
        ctx->launches++;
    }
    SynthParams sp{};
    sp.spectra = j.d_spectra; sp.scratch = tns_on ? j.d_scratch : nullptr; sp.ranges = j.d_ranges; sp.info = j.d_info;
    sp.pcm = static_cast<float *>(j.d_pcm);
    sp.qframes = reinterpret_cast<const uint8_t *>(j.d_q); sp.dq = ctx->d_dq; sp.pcm_s16 = j.s16 ? 1 : 0;
    sp.stereo = j.d_stereo;
    sp.ovl_in = ctx->d_ovl[ctx->cur];
    sp.ovl_out = j.in_place_state ? ctx->d_ovl[ctx->cur] : ctx->d_ovl[ctx->cur ^ 1];
    sp.tab = j.scale == 1.0f ? ctx->d_tab_unit : ctx->d_tab;
    const int n_pairs = (j.S_sub * j.nc + 1) / 2;
    // in-place state (inner seam): one item, so the state is read before it is written.  Fused: one item per client.
    const int Q = j.in_place_state ? n_pairs * j.T
                                   : pick_slice(n_pairs, j.T, ctx->num_sms * (fused ? kFusedClients : kWorkers), fused);
    sp.g = make_geometry(j.S_sub, j.T, j.nc, ctx->C, j.c0, j.s_base, Q);
    sp.scale = j.scale;
    // Two instantiations walk the same item list: the long-only one takes the items without
    // EIGHT_SHORT frames, the generic one the rest (each item is classified on the device).
    // If the long-only pass finds no such item the generic pass exits at once; a caller that has
    // looked at every window_sequence itself (the host path's validation) spares that launch.
    CU(ctx, cudaMemsetAsync(slots, 0, 4 * sizeof(unsigned), stream));
    sp.short_items = slots + 2;
    if (fused) {
        sp.tns_ring = ctx->d_tns_ring[ring_i]; sp.tns_blob = j.d_blob; sp.tns_offsets = j.d_offsets; sp.tns_blob_bytes = j.blob_bytes;
        sp.tns_bands = ctx->d_bands; sp.tns_ar = tp.ar; sp.sample_index = ctx->sample_index;
        sp.counter = slots;
        CU(ctx, launch_synth_tns(sp, ctx->num_sms, stream));
        ctx->launches++;
        if (!j.no_short) {   // only does anything if the fused kernel counted an item with an EIGHT_SHORT frame
            CU(ctx, launch_tns(tp, stream));
            sp.counter = slots + 1;
            CU(ctx, launch_synth(sp, ctx->num_sms, true, stream));
            ctx->launches += 2;
        }
        return AACFB_OK;
    }
    for (int generic = 0; generic < (j.no_short ? 1 : 2); ++generic) {
        sp.counter = slots + generic;
        CU(ctx, launch_synth(sp, ctx->num_sms, generic != 0, stream));
        ctx->launches++;
    }
    return AACFB_OK;
}

int grow_lane_stereo(aacfb_ctx *ctx, Lane &ln, size_t n_cf, bool prepass) {
    if (n_cf / 2 > ln.cap_stereo) {
        cudaFree(ln.d_stereo); ln.d_stereo = nullptr; ln.cap_stereo = 0;
        CU(ctx, cudaMalloc(&ln.d_stereo, (n_cf / 2) * sizeof(aacfb_stereo_ops)));
        ln.cap_stereo = n_cf / 2;
    }
    if (prepass && n_cf > ln.cap_stereo_out) {
        cudaFree(ln.d_stereo_out); ln.d_stereo_out = nullptr; ln.cap_stereo_out = 0;
        CU(ctx, cudaMalloc(&ln.d_stereo_out, n_cf * 4096));
        ln.cap_stereo_out = n_cf;
    }
    return AACFB_OK;
}

// d_spectra holds float rows or aacfb_qframe records (2304 <= 4096 bytes), d_pcm float or int16 samples.
int grow_lane(aacfb_ctx *ctx, Lane &ln, size_t n_cf, bool need_scratch, bool need_deq = false) {
    if (n_cf > ln.cap_cf) {
        cudaFree(ln.d_spectra); cudaFree(ln.d_pcm); cudaFree(ln.d_info); cudaFree(ln.d_offsets);
        ln.d_spectra = ln.d_pcm = nullptr; ln.d_info = nullptr; ln.d_offsets = nullptr; ln.cap_cf = 0;
        CU(ctx, cudaMalloc(&ln.d_spectra, n_cf * 4096));
        CU(ctx, cudaMalloc(&ln.d_pcm, n_cf * 4096));
        CU(ctx, cudaMalloc(&ln.d_info, n_cf * sizeof(aacfb_frame_info)));
        CU(ctx, cudaMalloc(&ln.d_offsets, (n_cf + 1) * sizeof(uint32_t)));
        ln.cap_cf = n_cf;
    }
    if (need_scratch && n_cf > ln.cap_scratch) {
        cudaFree(ln.d_scratch); ln.d_scratch = nullptr; ln.cap_scratch = 0;
        CU(ctx, cudaMalloc(&ln.d_scratch, n_cf * 4100));
        ln.cap_scratch = n_cf;
    }
    if (need_deq && n_cf > ln.cap_deq) {
        cudaFree(ln.d_deq); ln.d_deq = nullptr; ln.cap_deq = 0;
        CU(ctx, cudaMalloc(&ln.d_deq, n_cf * 4096));
        ln.cap_deq = n_cf;
    }
    return AACFB_OK;
}

// Host-side validation of the side info of channel-frames [i0, i1) (the reference throws on these:
// tns.js:84-85; unknown window sequences cannot occur there, ics.js:282).  `blob_bytes` = offsets[n_cf]:
// every block has to lie inside it and the offsets must not decrease.  *any_short is raised when an
// EIGHT_SHORT frame is seen (spares the generic kernel launch when none is).
int validate(aacfb_ctx *ctx, const aacfb_frame_info *info, const uint8_t *blob, const uint32_t *offsets, size_t blob_bytes,
             size_t i0, size_t i1, bool *any_short) {
    for (size_t i = i0; i < i1; ++i) {
        if (info[i].window_sequence > 3)
            return fail(ctx, AACFB_ERR_SEQUENCE, "channel-frame %zu: window_sequence %u out of range", i,
                        (unsigned)info[i].window_sequence);
        if (info[i].window_sequence == AACFB_EIGHT_SHORT_SEQUENCE && any_short) *any_short = true;
        if (!blob || !offsets) continue;
        const uint32_t o0 = offsets[i], o1 = offsets[i + 1];
        if (o1 < o0 || o1 > blob_bytes) return fail(ctx, AACFB_ERR_TNS, "channel-frame %zu: TNS offsets decrease or leave the blob", i);
        if (!info[i].tns_present) continue;
        if (o0 & 3) return fail(ctx, AACFB_ERR_TNS, "channel-frame %zu: bad TNS offsets", i);
        if (o1 == o0) continue;
        if (o1 - o0 < 8) return fail(ctx, AACFB_ERR_TNS, "channel-frame %zu: TNS block too short", i);
        const uint8_t *b = blob + o0;
        uint32_t pos = 8;
        for (int w = 0; w < 8; ++w)
            for (int f = 0; f < b[w]; ++f) {
                if (pos + 4 > o1 - o0) return fail(ctx, AACFB_ERR_TNS, "channel-frame %zu: truncated TNS block", i);
                const unsigned order = b[pos + 1];
                if (order > AACFB_TNS_MAX_ORDER)
                    return fail(ctx, AACFB_ERR_TNS, "TNS filter out of range: %u", order);  // tns.js:85
                pos += 4 + 4 * order;
                if (pos > o1 - o0) return fail(ctx, AACFB_ERR_TNS, "channel-frame %zu: truncated TNS block", i);
            }
    }
    return AACFB_OK;
}

// Stereo side info: the flag belongs to the left channel of a pair, op codes index scale[128].
int validate_stereo(aacfb_ctx *ctx, const aacfb_frame_info *info, const aacfb_stereo_ops *ops, size_t i0, size_t i1) {
    for (size_t i = i0; i < i1; ++i) {
        if (!info[i].stereo_present) continue;
        if (i & 1) return fail(ctx, AACFB_ERR_ARG, "channel-frame %zu: stereo_present set on a right channel", i);
        const aacfb_stereo_ops &r = ops[i >> 1];
        for (int g = 0; g < 256; ++g)
            if (r.op[g] > AACFB_STEREO_IS + 127)
                return fail(ctx, AACFB_ERR_ARG, "channel-frame %zu: stereo op %u out of range", i, (unsigned)r.op[g]);
    }
    return AACFB_OK;
}

// Quantised input: maxSFB must fit the band table of the window length (the reference would read
// swbOffsets past its end, ics.js:217-219) and the groups of an EIGHT_SHORT frame must cover its 8 windows.
int validate_q(aacfb_ctx *ctx, const aacfb_frame_info *info, const aacfb_qframe *q, size_t i0, size_t i1) {
    const TnsBandTables &B = tns_band_tables();
    for (size_t i = i0; i < i1; ++i) {
        const bool is_short = info[i].window_sequence == AACFB_EIGHT_SHORT_SEQUENCE;
        const int nb = is_short ? B.swb_short_count[ctx->sample_index] : B.swb_long_count[ctx->sample_index];
        if (info[i].max_sfb > nb)
            return fail(ctx, AACFB_ERR_ARG, "channel-frame %zu: max_sfb %u exceeds the %d bands of this window length", i,
                        (unsigned)info[i].max_sfb, nb);
        int windows = 0, groups = 0;
        for (int g = 0; g < 8 && q[i].group_len[g]; ++g) { windows += q[i].group_len[g]; ++groups; }
        if (windows != (is_short ? 8 : 1))
            return fail(ctx, AACFB_ERR_ARG, "channel-frame %zu: group_len covers %d windows", i, windows);
        if (groups * (int)info[i].max_sfb > 120)
            return fail(ctx, AACFB_ERR_ARG, "channel-frame %zu: more than 120 sections", i);
    }
    return AACFB_OK;
}

}  // namespace

extern "C" {

#define API __attribute__((visibility("default")))

API int aacfb_version(void) { return 100; }

API const char *aacfb_last_error(const aacfb_ctx *ctx) { return ctx ? ctx->err : g_err; }

API uint64_t aacfb_launch_count(const aacfb_ctx *ctx) { return ctx ? ctx->launches : 0; }

API int aacfb_get_table(int which, float *dst, int capacity) {
    if (!dst) return AACFB_ERR_ARG;
    const HostTables &H = host_tables();
    const float *src = nullptr;
    int n = 0;
    switch (which) {
    case 0: src = &H.roots512[0][0]; n = 1024; break;
    case 1: src = &H.roots64[0][0]; n = 128; break;
    case 2: src = &H.synth.cs2048[0].x; n = 1024; break;
    case 3: src = &H.synth.cs256[0].x; n = 128; break;
    case 4: src = H.sine1024; n = 1024; break;
    case 5: src = H.kbd1024; n = 1024; break;
    case 6: src = H.sine128; n = 128; break;
    case 7: src = H.kbd128; n = 128; break;
    case 8: case 9: case 10: {   // inverse-quantisation tables (the same for every sample rate)
        static DequantTables *D = nullptr;
        static std::once_flag once;
        std::call_once(once, [] { D = new DequantTables; build_dequant_tables(4, *D); });
        src = which == 8 ? D->iq : which == 9 ? D->sf : D->noise;
        n = which == 8 ? 8192 : which == 9 ? 428 : 32;
        break;
    }
    default: return AACFB_ERR_ARG;
    }
    if (capacity < n) return AACFB_ERR_ARG;
    std::memcpy(dst, src, sizeof(float) * n);
    return n;
}

API int aacfb_get_swb_offsets(int sample_index, int is_short, uint16_t *dst, int capacity) {
    if (!dst || sample_index < 0 || sample_index > 11) return AACFB_ERR_ARG;
    const TnsBandTables &B = tns_band_tables();
    const int n = is_short ? B.swb_short_count[sample_index] : B.swb_long_count[sample_index];
    if (capacity < n + 1) return AACFB_ERR_ARG;
    std::memcpy(dst, is_short ? B.swb_short[sample_index] : B.swb_long[sample_index], sizeof(uint16_t) * (n + 1));
    return n;
}

// adts_demuxer.js:28-52 without a bit reader: the header is 7 (or 9) whole bytes.
static_assert(sizeof(aacfb_adts_frame) == 24, "aacfb_adts_frame layout");
API int aacfb_adts_index(const uint8_t *data, size_t size, aacfb_adts_frame *frames, int capacity, size_t *consumed) {
    if (!data && size) return fail(nullptr, AACFB_ERR_ARG, "null buffer");
    if (capacity < 0) return fail(nullptr, AACFB_ERR_ARG, "negative capacity");
    size_t p = 0;
    int n = 0;
    while (size - p >= 7 && (!frames || n < capacity)) {
        const uint8_t *h = data + p;
        if (h[0] != 0xff || (h[1] & 0xf0) != 0xf0) {
            if (consumed) *consumed = p;
            return fail(nullptr, AACFB_ERR_ADTS, "Invalid ADTS header.");   // adts_demuxer.js:30
        }
        const bool protection_absent = (h[1] & 1) != 0;
        const uint32_t frame_length = ((uint32_t)(h[3] & 3) << 11) | ((uint32_t)h[4] << 3) | (h[5] >> 5);
        const uint32_t header_bytes = protection_absent ? 7 : 9;
        if (frame_length < header_bytes) {
            if (consumed) *consumed = p;
            return fail(nullptr, AACFB_ERR_ADTS, "ADTS frame at offset %zu is shorter than its header", p);
        }
        if (frame_length > size - p) break;   // incomplete: the next buffer starts here
        if (frames) {
            aacfb_adts_frame f{};
            f.offset = p; f.frame_length = frame_length; f.header_bytes = (uint8_t)header_bytes;
            f.profile = (uint8_t)(((h[2] >> 6) & 3) + 1);
            f.sampling_index = (uint8_t)((h[2] >> 2) & 15);
            f.chan_config = (uint8_t)(((h[2] & 1) << 2) | (h[3] >> 6));
            f.num_frames = (uint8_t)((h[6] & 3) + 1);
            frames[n] = f;
        }
        ++n;
        p += frame_length;
    }
    if (consumed) *consumed = p;
    return n;
}

API int aacfb_create(aacfb_ctx **out, int device, int n_streams, int channels, int sample_index, int small_frames,
                     uint32_t flags) {
    if (!out) return fail(nullptr, AACFB_ERR_ARG, "null out pointer");
    *out = nullptr;
    if (small_frames) return fail(nullptr, AACFB_ERR_SMALL, "WHA?? No small frames allowed.");  // filter_bank.js:26
    if (n_streams < 1 || channels < 1 || channels > AACFB_MAX_CHANNELS)
        return fail(nullptr, AACFB_ERR_ARG, "n_streams/channels out of range");
    if (sample_index < 0 || sample_index > 11) return fail(nullptr, AACFB_ERR_ARG, "sample_index must be 0..11");
    if ((flags & AACFB_TNS_MODE_MASK) == 3u || (flags & ~AACFB_TNS_MODE_MASK))
        return fail(nullptr, AACFB_ERR_ARG, "unknown flags");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, AACFB_ERR_CUDA, "no CUDA device: %s (this library has no CPU path)", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, AACFB_ERR_ARG, "device %d out of range", device);
    aacfb_ctx *ctx = new (std::nothrow) aacfb_ctx;
    if (!ctx) return fail(nullptr, AACFB_ERR_NOMEM, "out of memory");
    ctx->device = device; ctx->S = n_streams; ctx->C = channels; ctx->sample_index = sample_index; ctx->flags = flags;
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    auto bail = [&](cudaError_t err, const char *what) {
        const int rc = fail(nullptr, AACFB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err));
        aacfb_destroy(ctx);
        return rc;
    };
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    if (prop.major != 10) {
        aacfb_destroy(ctx);
        return fail(nullptr, AACFB_ERR_CUDA, "device %d is sm_%d%d; this library ships sm_100a code only", device,
                    prop.major, prop.minor);
    }
    ctx->num_sms = prop.multiProcessorCount;
    const size_t ovl_bytes = (size_t)n_streams * channels * 4096;
    for (int i = 0; i < 2; ++i) {
        if ((e = cudaMalloc(&ctx->d_ovl[i], ovl_bytes)) != cudaSuccess) return bail(e, "cudaMalloc overlap");
        if ((e = cudaMemset(ctx->d_ovl[i], 0, ovl_bytes)) != cudaSuccess) return bail(e, "cudaMemset overlap");
    }
    {
        SynthTables *scaled = new (std::nothrow) SynthTables(host_tables().synth);
        if (!scaled) { aacfb_destroy(ctx); return fail(nullptr, AACFB_ERR_NOMEM, "out of memory"); }
        scale_windows(*scaled, 1.0f / 32768.0f);
        e = cudaMalloc(&ctx->d_tab, sizeof(SynthTables));
        if (e == cudaSuccess) e = cudaMemcpy(ctx->d_tab, scaled, sizeof(SynthTables), cudaMemcpyHostToDevice);
        delete scaled;
        if (e != cudaSuccess) return bail(e, "device tables");
        if ((e = cudaMalloc(&ctx->d_tab_unit, sizeof(SynthTables))) != cudaSuccess) return bail(e, "cudaMalloc tables");
        if ((e = cudaMemcpy(ctx->d_tab_unit, &host_tables().synth, sizeof(SynthTables), cudaMemcpyHostToDevice)) != cudaSuccess)
            return bail(e, "cudaMemcpy tables");
    }
    if ((e = cudaMalloc(&ctx->d_bands, sizeof(TnsBandTables))) != cudaSuccess) return bail(e, "cudaMalloc bands");
    if ((e = cudaMemcpy(ctx->d_bands, &tns_band_tables(), sizeof(TnsBandTables), cudaMemcpyHostToDevice)) != cudaSuccess)
        return bail(e, "cudaMemcpy bands");
    {
        DequantTables *dq = new (std::nothrow) DequantTables;
        if (!dq) { aacfb_destroy(ctx); return fail(nullptr, AACFB_ERR_NOMEM, "out of memory"); }
        build_dequant_tables(sample_index, *dq);
        e = cudaMalloc(&ctx->d_dq, sizeof(DequantTables));
        if (e == cudaSuccess) e = cudaMemcpy(ctx->d_dq, dq, sizeof(DequantTables), cudaMemcpyHostToDevice);
        delete dq;
        if (e != cudaSuccess) return bail(e, "dequant tables");
    }
    if ((e = cudaMalloc(&ctx->d_counters, kCounters * sizeof(unsigned))) != cudaSuccess) return bail(e, "cudaMalloc");
    for (int i = 0; i < kLanes; ++i) {
        Lane &ln = ctx->lane[i];
        if ((e = cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
        for (cudaEvent_t *ev : {&ln.ev_in, &ln.ev_k, &ln.ev_out})
            if ((e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    }
    for (cudaStream_t *st : {&ctx->h2d, &ctx->d2h})
        if ((e = cudaStreamCreateWithFlags(st, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    *out = ctx;
    return AACFB_OK;
}

API int aacfb_destroy(aacfb_ctx *ctx) {
    if (!ctx) return AACFB_OK;
    DeviceGuard guard(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < kLanes; ++i) {
        Lane &ln = ctx->lane[i];
        if (ln.stream) cudaStreamDestroy(ln.stream);
        for (cudaEvent_t ev : {ln.ev_in, ln.ev_k, ln.ev_out})
            if (ev) cudaEventDestroy(ev);
        cudaFree(ln.d_spectra); cudaFree(ln.d_pcm); cudaFree(ln.d_scratch); cudaFree(ln.d_info); cudaFree(ln.d_offsets);
        cudaFree(ln.d_stereo); cudaFree(ln.d_stereo_out); cudaFree(ln.d_deq);
    }
    if (ctx->h2d) cudaStreamDestroy(ctx->h2d);
    if (ctx->d2h) cudaStreamDestroy(ctx->d2h);
    cudaFree(ctx->d_dev_stereo_out); cudaFree(ctx->d_dev_deq); cudaFree(ctx->d_dq);
    for (float *r : ctx->d_tns_ring) cudaFree(r);
    cudaFree(ctx->d_ovl[0]); cudaFree(ctx->d_ovl[1]); cudaFree(ctx->d_tab); cudaFree(ctx->d_tab_unit); cudaFree(ctx->d_bands);
    cudaFree(ctx->d_counters); cudaFree(ctx->d_blob); cudaFree(ctx->d_dev_scratch);
    delete ctx;
    return AACFB_OK;
}

API int aacfb_reset(aacfb_ctx *ctx) {
    if (!ctx) return fail(nullptr, AACFB_ERR_ARG, "null context");
    DeviceGuard guard(ctx->device);
    CU(ctx, cudaDeviceSynchronize());
    CU(ctx, cudaMemset(ctx->d_ovl[ctx->cur], 0, (size_t)ctx->S * ctx->C * 4096));
    return AACFB_OK;
}

API int aacfb_get_overlap(aacfb_ctx *ctx, float *overlap) {
    if (!ctx || !overlap) return fail(ctx, AACFB_ERR_ARG, "null argument");
    DeviceGuard guard(ctx->device);
    CU(ctx, cudaDeviceSynchronize());
    CU(ctx, cudaMemcpy(overlap, ctx->d_ovl[ctx->cur], (size_t)ctx->S * ctx->C * 4096, cudaMemcpyDeviceToHost));
    return AACFB_OK;
}

API int aacfb_set_overlap(aacfb_ctx *ctx, const float *overlap) {
    if (!ctx || !overlap) return fail(ctx, AACFB_ERR_ARG, "null argument");
    DeviceGuard guard(ctx->device);
    CU(ctx, cudaDeviceSynchronize());
    CU(ctx, cudaMemcpy(ctx->d_ovl[ctx->cur], overlap, (size_t)ctx->S * ctx->C * 4096, cudaMemcpyHostToDevice));
    return AACFB_OK;
}

API int aacfb_process_device(aacfb_ctx *ctx, const float *d_spectra, const aacfb_frame_info *d_info,
                             const uint8_t *d_tns_blob, const uint32_t *d_tns_offsets, size_t tns_blob_bytes,
                             float *d_pcm, int n_frames, void *stream) {
    return aacfb_process_device_io(ctx, d_spectra, AACFB_IN_F32, d_info, nullptr, d_tns_blob, d_tns_offsets, tns_blob_bytes,
                                   d_pcm, AACFB_PCM_F32, n_frames, stream);
}

API int aacfb_process_device_stereo(aacfb_ctx *ctx, const float *d_spectra, const aacfb_frame_info *d_info,
                                    const aacfb_stereo_ops *d_stereo_ops, const uint8_t *d_tns_blob,
                                    const uint32_t *d_tns_offsets, size_t tns_blob_bytes, float *d_pcm, int n_frames,
                                    void *stream) {
    return aacfb_process_device_io(ctx, d_spectra, AACFB_IN_F32, d_info, d_stereo_ops, d_tns_blob, d_tns_offsets,
                                   tns_blob_bytes, d_pcm, AACFB_PCM_F32, n_frames, stream);
}

// The device-pointer entry points TRUST their side info (it lives in device memory: no host-side
// validation; the kernels mask window_sequence to 2 bits and bound every TNS block by the blob size).
// Use them from ONE stream per context: the pre-pass scratch buffers belong to the context, and they
// are (re)allocated -- with a device synchronisation -- whenever a call needs more than any before it.
API int aacfb_process_device_io(aacfb_ctx *ctx, const void *d_input, uint32_t in_format, const aacfb_frame_info *d_info,
                                const aacfb_stereo_ops *d_stereo_ops, const uint8_t *d_tns_blob,
                                const uint32_t *d_tns_offsets, size_t tns_blob_bytes, void *d_pcm, uint32_t pcm_format,
                                int n_frames, void *stream) {
    if (!ctx) return fail(nullptr, AACFB_ERR_ARG, "null context");
    if (n_frames < 0) return fail(ctx, AACFB_ERR_ARG, "negative frame count");
    if (in_format > AACFB_IN_Q16 || pcm_format > AACFB_PCM_S16) return fail(ctx, AACFB_ERR_ARG, "unknown input / PCM format");
    if (n_frames == 0) return AACFB_OK;
    if (!d_input || !d_info || !d_pcm) return fail(ctx, AACFB_ERR_ARG, "null buffer");
    if ((reinterpret_cast<uintptr_t>(d_input) | reinterpret_cast<uintptr_t>(d_pcm)) & 15)
        return fail(ctx, AACFB_ERR_ARG, "input/pcm must be 16-byte aligned");
    DeviceGuard guard(ctx->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint32_t mode = ctx->flags & AACFB_TNS_MODE_MASK;
    const size_t n_cf = (size_t)ctx->S * n_frames * ctx->C;
    auto grow = [&](float *&buf, size_t &cap, size_t bytes_per_cf) -> int {
        if (n_cf > cap) {
            CU(ctx, cudaDeviceSynchronize());
            cudaFree(buf); buf = nullptr; cap = 0;
            CU(ctx, cudaMalloc(&buf, n_cf * bytes_per_cf));
            cap = n_cf;
        }
        return AACFB_OK;
    };
    int rc;
    float *scratch = nullptr;
    if (mode != AACFB_TNS_AS_SHIPPED && d_tns_blob && d_tns_offsets && tns_blob_bytes) {
        if ((rc = grow(ctx->d_dev_scratch, ctx->cap_dev_scratch, 4100)) != AACFB_OK) return rc;
        scratch = ctx->d_dev_scratch;
    }
    uint32_t *ranges = scratch ? reinterpret_cast<uint32_t *>(scratch + ctx->cap_dev_scratch * 1024) : nullptr;
    float *stereo_out = nullptr;
    bool stereo_pre = false;
    if (d_stereo_ops) {
        if (ctx->C & 1) return fail(ctx, AACFB_ERR_ARG, "stereo tools need an even channel count (pairs are channels 2j, 2j+1)");
        if (reinterpret_cast<uintptr_t>(d_stereo_ops) & 15) return fail(ctx, AACFB_ERR_ARG, "stereo_ops must be 16-byte aligned");
        if ((stereo_pre = stereo_needs_prepass(ctx->C, scratch != nullptr))) {
            if ((rc = grow(ctx->d_dev_stereo_out, ctx->cap_dev_stereo_out, 4096)) != AACFB_OK) return rc;
            stereo_out = ctx->d_dev_stereo_out;
        }
    }
    Job j;
    if (in_format == AACFB_IN_Q16) {
        j.d_q = static_cast<const aacfb_qframe *>(d_input);
        if (scratch || stereo_pre) {
            if ((rc = grow(ctx->d_dev_deq, ctx->cap_dev_deq, 4096)) != AACFB_OK) return rc;
            j.d_deq = ctx->d_dev_deq;
        }
    } else {
        j.d_spectra = static_cast<const float *>(d_input);
    }
    j.d_info = d_info; j.d_stereo = d_stereo_ops; j.d_stereo_out = stereo_out;
    j.d_blob = d_tns_blob; j.d_offsets = d_tns_offsets; j.blob_bytes = tns_blob_bytes;
    j.d_scratch = scratch; j.d_ranges = ranges; j.d_pcm = d_pcm; j.s16 = pcm_format == AACFB_PCM_S16;
    j.S_sub = ctx->S; j.s_base = 0; j.T = n_frames; j.nc = ctx->C; j.c0 = 0;
    j.scale = j.s16 ? 1.0f : 1.0f / 32768.0f;
    if ((rc = enqueue(ctx, j, st)) != AACFB_OK) return rc;
    ctx->cur ^= 1;
    return AACFB_OK;
}

API int aacfb_process(aacfb_ctx *ctx, const float *spectra, const aacfb_frame_info *info, const uint8_t *tns_blob,
                      const uint32_t *tns_offsets, float *pcm, int n_frames) {
    return aacfb_process_io(ctx, spectra, AACFB_IN_F32, info, nullptr, tns_blob, tns_offsets, pcm, AACFB_PCM_F32, n_frames);
}

API int aacfb_process_stereo(aacfb_ctx *ctx, const float *spectra, const aacfb_frame_info *info,
                             const aacfb_stereo_ops *stereo_ops, const uint8_t *tns_blob, const uint32_t *tns_offsets,
                             float *pcm, int n_frames) {
    return aacfb_process_io(ctx, spectra, AACFB_IN_F32, info, stereo_ops, tns_blob, tns_offsets, pcm, AACFB_PCM_F32, n_frames);
}

API int aacfb_process_io(aacfb_ctx *ctx, const void *input, uint32_t in_format, const aacfb_frame_info *info,
                         const aacfb_stereo_ops *stereo_ops, const uint8_t *tns_blob, const uint32_t *tns_offsets,
                         void *pcm, uint32_t pcm_format, int n_frames) {
    if (!ctx) return fail(nullptr, AACFB_ERR_ARG, "null context");
    if (n_frames < 0) return fail(ctx, AACFB_ERR_ARG, "negative frame count");
    if (in_format > AACFB_IN_Q16 || pcm_format > AACFB_PCM_S16) return fail(ctx, AACFB_ERR_ARG, "unknown input / PCM format");
    if (n_frames == 0) return AACFB_OK;
    if (!input || !info || !pcm) return fail(ctx, AACFB_ERR_ARG, "null buffer");
    const int S = ctx->S, C = ctx->C, T = n_frames;
    if (stereo_ops && (C & 1)) return fail(ctx, AACFB_ERR_ARG, "stereo tools need an even channel count (pairs are channels 2j, 2j+1)");
    const size_t per_stream = (size_t)T * C, n_all = (size_t)S * per_stream;
    const bool q16 = in_format == AACFB_IN_Q16, s16 = pcm_format == AACFB_PCM_S16;
    const size_t in_bytes = q16 ? sizeof(aacfb_qframe) : 4096, out_bytes = s16 ? 2048 : 4096;   // per channel-frame
    const uint8_t *in8 = static_cast<const uint8_t *>(input);
    uint8_t *out8 = static_cast<uint8_t *>(pcm);
    DeviceGuard guard(ctx->device);
    const uint32_t mode = ctx->flags & AACFB_TNS_MODE_MASK;
    const bool tns_on = mode != AACFB_TNS_AS_SHIPPED && tns_blob && tns_offsets && tns_offsets[n_all] > 0;
    const size_t blob_bytes = (tns_blob && tns_offsets) ? tns_offsets[n_all] : 0;
    int rc;
    // Sub-batches of whole streams, two in flight: the copy-in of one overlaps
    // the kernel and copy-out of the other (PCIe is the bottleneck end to end).
    // Equal sub-batches (a ramp of small first/last sub-batches: no gain); their number: below.
    // The copies of ALL sub-batches run on two streams of their own, one per PCIe direction, tied to the lanes'
    // kernels by events: with everything of a sub-batch on its lane's stream the next copy-in of a lane had to
    // wait for the lane's copy-out (stream order), and the H2D link idled for (kernel + launch gaps) per
    // sub-batch -- 7.2 ms instead of the ~6.5 ms the link allows for 303 MB in / 268 MB out.
    // AACFB_COPY_STREAMS=0: the lane-ordered pipeline (A/B).
    static const bool copy_streams = [] { const char *e = std::getenv("AACFB_COPY_STREAMS"); return !e || std::atoi(e) != 0; }();
    int lanes = 2;
    std::vector<int> parts;   // streams per sub-batch
    // Sub-batch count: about 68 MB (in + out) each, between 8 and 16 -- measured on config 2: float in / float out
    // (1075 MB) 16 sub-batches 12.22 ms, 8: 12.50; aacfb_qframe in / int16 out (571 MB) 8: 6.95 ms, 16: 7.02, 32: 7.89.
    const size_t total_bytes = n_all * (in_bytes + out_bytes);
    int n_sub = std::min(S, (int)std::min<size_t>(16, std::max<size_t>(8, (total_bytes + (34u << 20)) / (68u << 20))));
    if (const char *env = std::getenv("AACFB_SUB_BATCHES")) n_sub = std::max(1, std::min(S, std::atoi(env)));   // tuning aids
    if (const char *env = std::getenv("AACFB_LANES")) lanes = std::atoi(env) >= 4 ? 4 : std::atoi(env) >= 2 ? 2 : 1;  // divisors of the counter ring
    if (total_bytes < (size_t)(16u << 20)) n_sub = 1;
    for (int i = 0, done = 0; i < n_sub; ++i) {
        const int upto = (int)((long long)S * (i + 1) / n_sub);
        if (upto > done) parts.push_back(upto - done);
        done = upto;
    }
    // The copy-out of the LAST sub-batch overlaps nothing (and so does the copy-in of the first): taper both ends
    // by halving the outermost parts AACFB_TAPER times (tuning aid; 0 = equal parts).
    int taper = 0;
    if (const char *env = std::getenv("AACFB_TAPER")) taper = std::max(0, std::min(4, std::atoi(env)));
    for (int k = 0; k < taper && parts.size() > 1; ++k) {
        const int last = parts.back();
        if (last >= 2) { parts.back() = last - last / 2; parts.push_back(last / 2); }
        const int first = parts.front();
        if (first >= 2) { parts.front() = first - first / 2; parts.insert(parts.begin(), first / 2); }
    }
    // The side info is validated sub-batch by sub-batch, each right before its copies are queued, so
    // only the first one's check delays the first H2D; the rest overlaps the transfers in flight.
    // A small call (one sub-batch) is checked as a whole, which also tells whether the generic
    // kernel instantiation is needed at all.
    bool any_short = false;
    auto check = [&](size_t i0, size_t i1) -> int {
        int r = validate(ctx, info, tns_blob, tns_offsets, blob_bytes, i0, i1, &any_short);
        if (r == AACFB_OK && stereo_ops) r = validate_stereo(ctx, info, stereo_ops, i0, i1);
        if (r == AACFB_OK && q16) r = validate_q(ctx, info, static_cast<const aacfb_qframe *>(input), i0, i1);
        return r;
    };
    const bool whole_check = parts.size() == 1;
    if (whole_check && (rc = check(0, n_all)) != AACFB_OK) return rc;
    // From here on work is queued on the lanes: an error must drain them before returning (the
    // caller's buffers are referenced by copies in flight) and leaves ctx->cur -- the overlap state
    // the next call starts from -- untouched.
    auto drain = [&](int code) {
        cudaStreamSynchronize(ctx->h2d);
        for (int i = 0; i < kLanes; ++i) cudaStreamSynchronize(ctx->lane[i].stream);
        cudaStreamSynchronize(ctx->d2h);
        return code;
    };
    if (tns_on) {
        if (blob_bytes > ctx->cap_blob) {
            CU(ctx, cudaDeviceSynchronize());
            cudaFree(ctx->d_blob); ctx->d_blob = nullptr; ctx->cap_blob = 0;
            CU(ctx, cudaMalloc(&ctx->d_blob, blob_bytes + 16));
            ctx->cap_blob = blob_bytes;
        }
        CU(ctx, cudaMemcpyAsync(ctx->d_blob, tns_blob, blob_bytes, cudaMemcpyHostToDevice, ctx->lane[0].stream));
        CU(ctx, cudaStreamSynchronize(ctx->lane[0].stream));
    }
    const int s_max = *std::max_element(parts.begin(), parts.end());
    const bool stereo_pre = stereo_ops && stereo_needs_prepass(C, tns_on);
    for (int i = 0; i < lanes; ++i) {
        if ((rc = grow_lane(ctx, ctx->lane[i], (size_t)s_max * per_stream, tns_on, q16 && (tns_on || stereo_pre))) != AACFB_OK) return rc;
        if (stereo_ops && (rc = grow_lane_stereo(ctx, ctx->lane[i], (size_t)s_max * per_stream, stereo_pre)) != AACFB_OK) return rc;
    }
#define CUD(call)                                                                                                    \
    do {                                                                                                             \
        cudaError_t e__ = (call);                                                                                    \
        if (e__ != cudaSuccess)                                                                                      \
            return drain(fail(ctx, AACFB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__)); \
    } while (0)
    int li = 0, s0 = 0;
    bool used[kLanes] = {};
    for (size_t pi = 0; pi < parts.size(); s0 += parts[pi], ++pi, li = (li + 1) % lanes) {
        Lane &ln = ctx->lane[li];
        const int sn = parts[pi];
        const size_t n_cf = (size_t)sn * per_stream, off = (size_t)s0 * per_stream;
        if (!whole_check && (rc = check(off, off + n_cf)) != AACFB_OK) return drain(rc);
        // lane-ordered: stream order makes reuse of this lane's buffers safe.  Copy streams: the lane's inputs are
        // free once its previous kernels are done, its PCM buffer once the previous copy-out is.
        cudaStream_t sin = copy_streams ? ctx->h2d : ln.stream, sout = copy_streams ? ctx->d2h : ln.stream;
        if (copy_streams && used[li]) CUD(cudaStreamWaitEvent(sin, ln.ev_k, 0));
        CUD(cudaMemcpyAsync(ln.d_spectra, in8 + off * in_bytes, n_cf * in_bytes, cudaMemcpyHostToDevice, sin));
        CUD(cudaMemcpyAsync(ln.d_info, info + off, n_cf * sizeof(aacfb_frame_info), cudaMemcpyHostToDevice, sin));
        if (tns_on)
            CUD(cudaMemcpyAsync(ln.d_offsets, tns_offsets + off, (n_cf + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, sin));
        if (stereo_ops)
            CUD(cudaMemcpyAsync(ln.d_stereo, stereo_ops + off / 2, (n_cf / 2) * sizeof(aacfb_stereo_ops),
                                cudaMemcpyHostToDevice, sin));
        if (copy_streams) {
            CUD(cudaEventRecord(ln.ev_in, sin));
            CUD(cudaStreamWaitEvent(ln.stream, ln.ev_in, 0));
            if (used[li]) CUD(cudaStreamWaitEvent(ln.stream, ln.ev_out, 0));
        }
        Job j;
        if (q16) { j.d_q = reinterpret_cast<const aacfb_qframe *>(ln.d_spectra); j.d_deq = ln.d_deq; }
        else j.d_spectra = ln.d_spectra;
        j.d_info = ln.d_info; j.d_stereo = stereo_ops ? ln.d_stereo : nullptr; j.d_stereo_out = ln.d_stereo_out;
        j.d_blob = tns_on ? ctx->d_blob : nullptr; j.d_offsets = ln.d_offsets; j.blob_bytes = tns_on ? blob_bytes : 0;
        j.d_scratch = ln.d_scratch; j.d_ranges = ln.d_scratch ? ln.ranges() : nullptr;
        j.d_pcm = ln.d_pcm; j.s16 = s16;
        j.S_sub = sn; j.s_base = s0; j.T = T; j.nc = C; j.c0 = 0;
        j.scale = s16 ? 1.0f : 1.0f / 32768.0f;
        j.no_short = whole_check && !any_short;
        if ((rc = enqueue(ctx, j, ln.stream)) != AACFB_OK) return drain(rc);
        if (copy_streams) {
            CUD(cudaEventRecord(ln.ev_k, ln.stream));
            CUD(cudaStreamWaitEvent(sout, ln.ev_k, 0));
        }
        CUD(cudaMemcpyAsync(out8 + off * out_bytes, ln.d_pcm, n_cf * out_bytes, cudaMemcpyDeviceToHost, sout));
        if (copy_streams) CUD(cudaEventRecord(ln.ev_out, sout));
        used[li] = true;
    }
#undef CUD
    CU(ctx, cudaStreamSynchronize(ctx->h2d));
    for (int i = 0; i < kLanes; ++i) CU(ctx, cudaStreamSynchronize(ctx->lane[i].stream));
    CU(ctx, cudaStreamSynchronize(ctx->d2h));
    ctx->cur ^= 1;
    return AACFB_OK;
}

// ---- page-locked host memory for the caller's staging buffers (include/aacfb.h) -------------------
API void *aacfb_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (bytes == 0) return nullptr;
    const cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { fail(nullptr, AACFB_ERR_CUDA, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e)); return nullptr; }
    return p;
}
API int aacfb_host_free(void *p) {
    if (!p) return AACFB_OK;
    const cudaError_t e = cudaFreeHost(p);
    return e == cudaSuccess ? AACFB_OK : fail(nullptr, AACFB_ERR_CUDA, "cudaFreeHost: %s", cudaGetErrorString(e));
}
API int aacfb_host_register(void *p, size_t bytes) {
    if (!p || !bytes) return fail(nullptr, AACFB_ERR_ARG, "null buffer");
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    return e == cudaSuccess ? AACFB_OK : fail(nullptr, AACFB_ERR_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e));
}
API int aacfb_host_unregister(void *p) {
    if (!p) return AACFB_OK;
    const cudaError_t e = cudaHostUnregister(p);
    return e == cudaSuccess ? AACFB_OK : fail(nullptr, AACFB_ERR_CUDA, "cudaHostUnregister: %s", cudaGetErrorString(e));
}

API int aacfb_filterbank_process(aacfb_ctx *ctx, int stream, int channel, const aacfb_frame_info *info,
                                 const float *input, float *output) {
    if (!ctx || !info || !input || !output) return fail(ctx, AACFB_ERR_ARG, "null argument");
    if (stream < 0 || stream >= ctx->S || channel < 0 || channel >= ctx->C)
        return fail(ctx, AACFB_ERR_ARG, "stream/channel out of range");
    if (info->window_sequence > 3) {  // filter_bank.js:104 has no default case: zero output, state untouched
        std::memset(output, 0, 4096);
        return AACFB_OK;
    }
    DeviceGuard guard(ctx->device);
    Lane &ln = ctx->lane[0];
    int rc = grow_lane(ctx, ln, 1, false);
    if (rc != AACFB_OK) return rc;
    aacfb_frame_info fi = *info;
    fi.tns_present = 0;  // the inner seam is the filterbank alone
    CU(ctx, cudaMemcpyAsync(ln.d_spectra, input, 4096, cudaMemcpyHostToDevice, ln.stream));
    CU(ctx, cudaMemcpyAsync(ln.d_info, &fi, sizeof fi, cudaMemcpyHostToDevice, ln.stream));
    {
        Job j;
        j.d_spectra = ln.d_spectra; j.d_info = ln.d_info; j.d_pcm = ln.d_pcm;
        j.S_sub = 1; j.s_base = stream; j.T = 1; j.nc = 1; j.c0 = channel; j.scale = 1.0f; j.in_place_state = true;
        j.no_short = fi.window_sequence != AACFB_EIGHT_SHORT_SEQUENCE;
        rc = enqueue(ctx, j, ln.stream);
    }
    if (rc != AACFB_OK) return rc;
    CU(ctx, cudaMemcpyAsync(output, ln.d_pcm, 4096, cudaMemcpyDeviceToHost, ln.stream));
    CU(ctx, cudaStreamSynchronize(ln.stream));
    return AACFB_OK;
}

API int aacfb_tns_process(aacfb_ctx *ctx, const aacfb_frame_info *info, const uint8_t *tns_block, size_t block_bytes,
                          float *data, uint32_t mode) {
    if (!ctx || !info || !data) return fail(ctx, AACFB_ERR_ARG, "null argument");
    mode &= AACFB_TNS_MODE_MASK;
    if (mode == 3u) return fail(ctx, AACFB_ERR_ARG, "unknown TNS mode");
    if (mode == AACFB_TNS_AS_SHIPPED || !tns_block || block_bytes == 0) return AACFB_OK;  // identity, tns.js:122
    aacfb_frame_info fi = *info;
    fi.tns_present = 1;
    const uint32_t offs[2] = {0, (uint32_t)block_bytes};
    int rc = validate(ctx, &fi, tns_block, offs, block_bytes, 0, 1, nullptr);
    if (rc != AACFB_OK) return rc;
    DeviceGuard guard(ctx->device);
    Lane &ln = ctx->lane[0];
    if ((rc = grow_lane(ctx, ln, 1, true)) != AACFB_OK) return rc;
    if (block_bytes > ctx->cap_blob) {
        CU(ctx, cudaDeviceSynchronize());
        cudaFree(ctx->d_blob); ctx->d_blob = nullptr; ctx->cap_blob = 0;
        CU(ctx, cudaMalloc(&ctx->d_blob, block_bytes + 16));
        ctx->cap_blob = block_bytes;
    }
    CU(ctx, cudaMemcpyAsync(ln.d_spectra, data, 4096, cudaMemcpyHostToDevice, ln.stream));
    CU(ctx, cudaMemcpyAsync(ln.d_info, &fi, sizeof fi, cudaMemcpyHostToDevice, ln.stream));
    CU(ctx, cudaMemcpyAsync(ln.d_offsets, offs, sizeof offs, cudaMemcpyHostToDevice, ln.stream));
    CU(ctx, cudaMemcpyAsync(ctx->d_blob, tns_block, block_bytes, cudaMemcpyHostToDevice, ln.stream));
    TnsParams tp{};
    // the kernel writes only the filtered interval of the row: start from a copy
    CU(ctx, cudaMemcpyAsync(ln.d_scratch, ln.d_spectra, 4096, cudaMemcpyDeviceToDevice, ln.stream));
    tp.spectra = ln.d_spectra; tp.scratch = ln.d_scratch; tp.ranges = ln.ranges(); tp.info = ln.d_info; tp.blob = ctx->d_blob;
    tp.offsets = ln.d_offsets; tp.blob_bytes = block_bytes; tp.n_cf = 1; tp.sample_index = ctx->sample_index;
    tp.ar = mode == AACFB_TNS_FIXED_AR; tp.bands = ctx->d_bands;
    CU(ctx, launch_tns(tp, ln.stream));
    ctx->launches++;
    CU(ctx, cudaMemcpyAsync(data, ln.d_scratch, 4096, cudaMemcpyDeviceToHost, ln.stream));
    CU(ctx, cudaStreamSynchronize(ln.stream));
    return AACFB_OK;
}

}  // extern "C"
