// aacfb_emul.cpp -- CPU emulation of the synthesis kernel's worker schedule.
//
// TEST-ONLY.  Runs the very same __host__ __device__ phase code the CUDA
// kernel runs (aacfb_core.cuh / aacfb_worker.cuh): 64 host threads play the
// 64 threads of one worker, a std::barrier plays the named barrier, and a
// plain memcpy plays the TMA load.  It lets the CPU-only test tier check the
// kernel's index maps, swizzles, window-switching and chunk/halo logic against
// the oracle bit patterns without a GPU.  It is never linked into
// libaacfb.so and is not a fallback: the shipped library has no CPU path.
#include <algorithm>
#include <barrier>
#include <cstring>
#include <thread>
#include <vector>

#include <cuda_runtime.h>  // float2/float4 only

#include "aacfb_tables.h"
#include "aacfb_worker.cuh"

using namespace aacfb;

namespace {
struct HostSync {
    std::barrier<> *bar;
    float *mailbox;  // 64 slots shared by the worker's threads
    void barrier() { bar->arrive_and_wait(); }
    void stage_free() { bar->arrive_and_wait(); }
    float partner(int u, float v) {  // the kernel's __shfl_xor_sync(.., 31)
        mailbox[u] = v;
        bar->arrive_and_wait();
        const float r = mailbox[63 - u];
        bar->arrive_and_wait();
        return r;
    }
    float partner7(int u, float v) {  // the kernel's __shfl_xor_sync(.., 7)
        mailbox[u] = v;
        bar->arrive_and_wait();
        const float r = mailbox[u ^ 7];
        bar->arrive_and_wait();
        return r;
    }
};
}  // namespace

// input: float spectra (AACFB_IN_F32) or aacfb_qframe records (AACFB_IN_Q16); pcm: float or int16.
extern "C" __attribute__((visibility("default"))) int aacfb_emul_process_io(
    const void *input, uint32_t in_format, const aacfb_frame_info *info, const uint8_t *tns_blob, const uint32_t *tns_offsets,
    float *overlap /*[S][C][1024] in/out*/, void *pcm_out, uint32_t pcm_format, int S, int T, int C, int sample_index,
    uint32_t flags, int chunk_len) {
    const HostTables &H = host_tables();
    const bool s16 = pcm_format == AACFB_PCM_S16;
    const float kScale = s16 ? 1.0f : 1.0f / 32768.0f;   // int16 PCM: unit-scale windows, like the library
    float *pcm = static_cast<float *>(pcm_out);
    static DequantTables dq_tab;
    build_dequant_tables(sample_index, dq_tab);
    const uint8_t *qframes = in_format == AACFB_IN_Q16 ? static_cast<const uint8_t *>(input) : nullptr;
    const float *spectra = qframes ? nullptr : static_cast<const float *>(input);
    std::vector<float> deq;
    // the kernel's tables: every window carries the output scale (scale_windows)
    static SynthTables scaled;
    scaled = H.synth;
    scale_windows(scaled, kScale);
    const SynthTables *tab = &scaled;
    const TnsBandTables &bt = tns_band_tables();
    const uint32_t mode = flags & AACFB_TNS_MODE_MASK;

    // TNS pre-pass (the kernel's tns_kernel): filtered copies of the rows that carry TNS
    std::vector<float> scratch;
    const bool tns_on = mode != AACFB_TNS_AS_SHIPPED && tns_blob && tns_offsets;
    if (tns_on && qframes) {   // the library's dequant_kernel pre-pass: TNS needs float rows
        deq.resize((size_t)S * T * C * 1024);
        for (size_t cf = 0; cf < (size_t)S * T * C; ++cf)
            dequant_row(&dq_tab, qframes + cf * kQFrameBytes, fb_pack(info[cf]), deq.data() + cf * 1024);
        spectra = deq.data();
        qframes = nullptr;
    }
    if (tns_on) {
        scratch.assign(spectra, spectra + (size_t)S * T * C * 1024);
        for (size_t cf = 0; cf < (size_t)S * T * C; ++cf) {
            if (!info[cf].tns_present) continue;
            const uint32_t o0 = tns_offsets[cf], o1 = tns_offsets[cf + 1];
            std::fill_n(scratch.begin() + cf * 1024, 1024, -1.0e30f);  // tns_apply must write every coefficient
            tns_apply(fb_pack(info[cf]) & 0xffffff03u, tns_blob + o0, o1 > o0 ? o1 - o0 : 0, sample_index,
                      mode == AACFB_TNS_FIXED_AR, bt, spectra + cf * 1024, scratch.data() + cf * 1024);
        }
    }
    const float *rows = tns_on ? scratch.data() : spectra;

    Geometry g = make_geometry(S, T, C, C, 0, 0, chunk_len);
    g.stride = 1;  // in-place overlap state below: visit the slices in memory order

    for (int item = 0; item < g.n_items; ++item) {
        const int f0 = item_begin(g, item), f1 = item_end(g, item);
        const int fb = (f0 % g.T) != 0 ? f0 - 1 : f0;
        const int nf = f1 - fb;
        std::barrier<> bar(kWorkerThreads);
        std::vector<float> stage_mem(3 * kStageFloats + 4);
        float *stage = stage_mem.data();
        while (reinterpret_cast<uintptr_t>(stage) & 15) ++stage;
        float *scratch2[2] = {stage + kStageFloats, stage + 2 * kStageFloats};
        float mailbox[kWorkerThreads];
        bool item_has_short = false;  // the kernel's per-item classification
        for (int f = 0; f < nf; ++f) {
            const Pair pp = make_pair(g, (int)((fb + f) / g.T));
            const int t = (int)((fb + f) % g.T);
            for (int c = 0; c < pp.nch; ++c)
                item_has_short |= info[cf_index(g, pp.s[c], t, pp.j[c])].window_sequence == AACFB_EIGHT_SHORT_SEQUENCE;
        }

        auto body = [&](int u) {
            HostSync sync{&bar, mailbox};
            Pts z;
            Ovl ov;
            std::memset(&ov, 0, sizeof ov);
            for (int f = 0; f < nf; ++f) {
                const Pair pr = make_pair(g, (int)((fb + f) / g.T));
                const int t = (int)((fb + f) % g.T);
                if (t == 0) {
                    ovl_load<0>(u, overlap + state_index(g, pr.s[0], pr.j[0]), ov, kScale);
                    if (pr.nch == 2) ovl_load<1>(u, overlap + state_index(g, pr.s[1], pr.j[1]), ov, kScale);
                }
                // "TMA": thread 0 fills the stage, everyone waits
                if (u == 0)
                    for (int c = 0; c < pr.nch; ++c) {
                        const size_t cf = cf_index(g, pr.s[c], t, pr.j[c]);
                        if (qframes) std::memcpy(reinterpret_cast<uint8_t *>(stage + 1024 * c) + kQLandOffset, qframes + cf * kQFrameBytes, kQFrameBytes);
                        else std::memcpy(stage + 1024 * c, rows + cf * 1024, 4096);
                    }
                bar.arrive_and_wait();
                FrameIO io;
                io.stage = stage;
                io.scratch = scratch2[f & 1];
                io.nch = pr.nch;
                io.ops = nullptr;  // the stereo tools are exercised on the device only
                DqCtx dqc;
                dqc.D = &dq_tab; dqc.iq_lo = dq_tab.iq; dqc.sf = dq_tab.sf;
                dq_thread_consts(&dq_tab, u, dqc);
                io.dq = qframes ? &dqc : nullptr;
                io.dst.s16 = s16;
                io.dst.emit = fb + f >= f0;
                io.dst.interleaved = pr.interleaved;
                io.dst.scale = kScale;
                io.dst.inv_scale = 1.0f / kScale;
                io.dst.ostride = g.nc;
                for (int c = 0; c < 2; ++c) io.fi[c] = fb_pack(info[cf_index(g, pr.s[c], t, pr.j[c])]);
                const size_t oa = ((size_t)pr.s[0] * g.T + t) * 1024 * g.nc + pr.j[0], ob = ((size_t)pr.s[1] * g.T + t) * 1024 * g.nc + pr.j[1];
                io.dst.out0 = s16 ? reinterpret_cast<float *>(static_cast<int16_t *>(pcm_out) + oa) : pcm + oa;
                io.dst.out1 = s16 ? reinterpret_cast<float *>(static_cast<int16_t *>(pcm_out) + ob) : pcm + ob;
                if (item_has_short) worker_frame<true, false, true>(u, sync, io, tab, tab, z, ov);
                else worker_frame<false, false, true>(u, sync, io, tab, tab, z, ov);
                if (t == g.T - 1) {
                    // NOTE: in place -- safe here because items run one after the other, in order
                    ovl_store<0>(u, ov, overlap + state_index(g, pr.s[0], pr.j[0]), 1.0f / kScale);
                    if (pr.nch == 2) ovl_store<1>(u, ov, overlap + state_index(g, pr.s[1], pr.j[1]), 1.0f / kScale);
                }
            }
        };
        std::vector<std::thread> th;
        for (int u = 0; u < kWorkerThreads; ++u) th.emplace_back(body, u);
        for (auto &t : th) t.join();
    }
    return 0;
}

extern "C" __attribute__((visibility("default"))) int aacfb_emul_process(
    const float *spectra, const aacfb_frame_info *info, const uint8_t *tns_blob, const uint32_t *tns_offsets,
    float *overlap, float *pcm, int S, int T, int C, int sample_index, uint32_t flags, int chunk_len) {
    return aacfb_emul_process_io(spectra, AACFB_IN_F32, info, tns_blob, tns_offsets, overlap, pcm, AACFB_PCM_F32, S, T, C,
                                 sample_index, flags, chunk_len);
}

// dequant_row / pcm_s16 alone (unit tests of the phase code against the oracle)
extern "C" __attribute__((visibility("default"))) void aacfb_emul_dequant(const uint8_t *qframe, const aacfb_frame_info *info,
                                                                         int sample_index, float *out) {
    DequantTables *D = new DequantTables;
    build_dequant_tables(sample_index, *D);
    dequant_row(D, qframe, fb_pack(*info), out);
    delete D;
}
extern "C" __attribute__((visibility("default"))) void aacfb_emul_pcm_s16(const float *x, int16_t *out, int n) {
    for (int i = 0; i < n; ++i) out[i] = (int16_t)pcm_s16(x[i]);
}

extern "C" __attribute__((visibility("default"))) int aacfb_emul_table(int which, float *dst, int cap) {
    const HostTables &H = host_tables();
    auto put = [&](const float *src, int n) { if (cap < n) return -1; std::memcpy(dst, src, sizeof(float) * n); return n; };
    switch (which) {
    case 0: return put(&H.roots512[0][0], 1024);
    case 1: return put(&H.roots64[0][0], 128);
    case 2: return put(&H.synth.cs2048[0].x, 1024);
    case 3: return put(&H.synth.cs256[0].x, 128);
    case 4: return put(H.sine1024, 1024);
    case 5: return put(H.kbd1024, 1024);
    case 6: return put(H.sine128, 128);
    case 7: return put(H.kbd128, 128);
    }
    return -1;
}
