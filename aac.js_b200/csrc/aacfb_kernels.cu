// aacfb_kernels.cu -- sm_100a kernels of the filterbank-synthesis path.
//
//   synth_kernel : TNS-filtered spectra -> IMDCT -> window -> overlap-add ->
//                  interleave, x 2^-15          (reference filter_bank.js:88-204,
//                  mdct.js:62-115, fft.js:105-192, decoder.js:204-213)
//   tns_kernel   : TNS.process               (reference tns.js:105-177)
//
// synth_kernel is persistent: one CTA per SM, kWorkers workers of 64 threads
// per CTA.  A worker pulls (chain pair, time chunk) items from a global counter
// and streams through the chunk's frames.  Spectrum rows are brought from
// HBM by 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP) into a ring of
// 8 KiB staging buffers per worker, completion signalled on mbarriers, so
// the next frame's rows are always in flight while the current one is
// transformed.  All FFT data movement stays in registers, the staging buffer
// and a scratch buffer; the only global stores are fully coalesced float4
// rows of finished PCM.  See aacfb_worker.cuh for
// the per-frame schedule and aacfb_core.cuh for the arithmetic.
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "aacfb_kernels.h"
#include "aacfb_worker.cuh"

namespace aacfb {

namespace {

constexpr int kStageBytes = kStageFloats * 4;
constexpr int kTabBytes = (kSmemTableBytes + 127) & ~127;

// Shared-memory layout of a CTA with W workers and a TMA ring of ST stages per worker.
template <int W, int ST>
struct Layout {
    static constexpr int kBufsPerWorker = ST + 2;  // TMA ring + two alternating scratch buffers
    static constexpr int kOffStages = kTabBytes;
    static constexpr int kOffBars = kOffStages + W * kBufsPerWorker * kStageBytes;
    static constexpr int kOffSlots = kOffBars + W * ST * 8;
    static constexpr int kTotal = kOffSlots + W * 32;
    static_assert(kTotal <= 227 * 1024, "shared memory budget");
    static_assert(2 * W + 1 <= 16, "named barriers");
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// 1-D TMA: global -> shared, completes `bytes` on the mbarrier.
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

struct RowSource {
    const float *spectra, *scratch;
    const aacfb_frame_info *info;
};

struct DevSync {
    uint32_t bar_id;      // named barrier of this worker (all 64 threads block)
    uint32_t free_id;     // named barrier "stage is free": followers arrive, the leader's warp waits
    bool leader_warp, leader;
    // the refill this frame's stage_free() has to issue (leader only)
    bool next_valid;
    uint32_t dst, mbar;
    const float *src[2];
    int nrows;
    __device__ __forceinline__ void barrier() { asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory"); }
    __device__ __forceinline__ void issue() {
        mbar_expect_tx(mbar, (uint32_t)nrows * 4096u);
        bulk_load(dst, src[0], 4096u, mbar);
        if (nrows == 2) bulk_load(dst + 4096u, src[1], 4096u, mbar);
    }
    __device__ __forceinline__ void stage_free() {
        // order this thread's generic-proxy accesses to the stage before the async-proxy refill
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (leader_warp) {
            asm volatile("bar.sync %0, 64;" ::"r"(free_id) : "memory");
            if (leader && next_valid) issue();
        } else {
            asm volatile("bar.arrive %0, 64;" ::"r"(free_id) : "memory");
        }
    }
    // value of thread 63-u: lane ^ 31 of the same warp (worker_thread_index)
    __device__ __forceinline__ float partner(int, float v) { return __shfl_xor_sync(0xffffffffu, v, 31); }
};

// aacfb_frame_info is 8 bytes: one 64-bit load, low word = FrameBits, byte 4 = tns_present
__device__ __forceinline__ uint2 info_raw(const SynthParams &P, size_t cf) {
    return __ldg(reinterpret_cast<const uint2 *>(P.info) + cf);
}
__device__ __forceinline__ FrameBits info_lo(const SynthParams &P, size_t cf) { return info_raw(P, cf).x; }
__device__ __forceinline__ const float *row_ptr(const SynthParams &P, size_t cf) {
    const bool tns = P.scratch != nullptr && (info_raw(P, cf).y & 0xffu) != 0;
    return (tns ? P.scratch : P.spectra) + cf * 1024;
}

}  // namespace

// GENERIC = false: takes only work items without EIGHT_SHORT frames (the long-transform code
// alone, best register allocation); GENERIC = true: takes only the items that have one.
// Both instantiations are launched back to back and walk the same item list.
template <bool GENERIC, int W, int ST>
__global__ void __launch_bounds__(W * 64, 1) synth_kernel(const __grid_constant__ SynthParams P) {
    using L = Layout<W, ST>;
    constexpr int kCtaThreads = W * 64, kWorkers = W, kStages = ST, kBufsPerWorker = L::kBufsPerWorker;
    constexpr int kOffStages = L::kOffStages, kOffBars = L::kOffBars, kOffSlots = L::kOffSlots;
    extern __shared__ __align__(128) uint8_t smem[];
    // the long-only instantiation (launched first) counts the items it had to leave to us
    if (GENERIC && *P.short_items == 0u) return;
    const int tid = threadIdx.x, w = tid >> 6;
    const int u = worker_thread_index((tid >> 5) & 1, tid & 31);
    const bool leader = u == 0;

    // constant tables -> shared memory (once per CTA)
    {
        const float4 *src = reinterpret_cast<const float4 *>(P.tab);
        float4 *dst = reinterpret_cast<float4 *>(smem);
        for (int i = tid; i < kSmemTableBytes / 16; i += kCtaThreads) dst[i] = src[i];
        __syncthreads();
        // the long windows carry the output scale (a power of two; see win_first in aacfb_core.cuh)
        float *wz = reinterpret_cast<float *>(smem + offsetof(SynthTables, wz));
        for (int i = tid; i < 2 * 512 * 2; i += kCtaThreads) wz[i] *= P.scale;
    }
    const SynthTables *ts = reinterpret_cast<const SynthTables *>(smem);
    float *stages = reinterpret_cast<float *>(smem + kOffStages) + (size_t)w * kBufsPerWorker * kStageFloats;
    float *scratch = stages + kStages * kStageFloats;
    const uint32_t bars = smem_u32(smem + kOffBars) + w * kStages * 8;
    volatile int *slot = reinterpret_cast<volatile int *>(smem + kOffSlots) + 8 * w;  // [0] item, [1] short flag, [2..6] prefetch cursor
    if (leader) {
        for (int s = 0; s < kStages; ++s) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    DevSync sync;
    sync.bar_id = 1 + w;
    sync.free_id = 1 + kWorkers + w;
    sync.leader_warp = ((tid >> 5) & 1) == 0;
    sync.leader = leader;
    const Geometry g = P.g;
    uint32_t fc = 0;  // frames this worker has staged so far: ring position and mbarrier phase
    Pts z;
    Ovl ov;

    // First item of every worker is assigned statically, interleaved over the CTAs, so that a
    // batch with few items still spreads over all SMs; further items come from the counter.
    bool first = true;
    for (;;) {
        if (leader) {
            slot[0] = first ? w * (int)gridDim.x + (int)blockIdx.x
                            : kWorkers * (int)gridDim.x + (int)atomicAdd(P.counter, 1u);
            slot[1] = 0;
        }
        first = false;
        sync.barrier();
        const int item = slot[0];
        if (item >= g.n_items) break;
        const int f0 = item_begin(g, item), f1 = item_end(g, item);
        const int p0 = f0 / g.T, t0 = f0 - p0 * g.T;
        const int fb = t0 != 0 ? f0 - 1 : f0;   // halo frame first, unless the slice starts a pair
        const int tb = t0 != 0 ? t0 - 1 : 0;
        const int nf = f1 - fb;
        {   // does this item contain an EIGHT_SHORT frame?  (64 threads scan its side info)
            bool mine = false;
            for (int f = tid & 63; f < nf; f += 64) {
                const int ff = fb + f, pi = ff / g.T;
                const Pair pp = make_pair(g, pi);
                const int t = ff - pi * g.T;
                mine |= is_short(info_lo(P, cf_index(g, pp.s[0], t, pp.j[0])));
                if (pp.nch == 2) mine |= is_short(info_lo(P, cf_index(g, pp.s[1], t, pp.j[1])));
            }
            if (__any_sync(0xffffffffu, mine) && (tid & 31) == 0) slot[1] = 1;
            sync.barrier();
            const bool has_short = slot[1] != 0;
            sync.barrier();  // slot is rewritten by the leader at the next item
            if (has_short != GENERIC) {
                if (!GENERIC && leader) atomicAdd(P.short_items, 1u);
                continue;
            }
        }

        // Prefetch cursor: the frame whose rows go into the ring next.  Only the leader thread
        // uses it, so it lives in shared memory rather than in every thread's registers:
        // [0] pair, [1] t, [2],[3] channel-frame index of the two chains (advance by nc per frame).
        volatile int *cur = slot + 2;
        auto cursor_to = [&](int pair, int t) {
            const Pair pn = make_pair(g, pair);
            cur[0] = pair; cur[1] = t;
            cur[2] = (int)cf_index(g, pn.s[0], t, pn.j[0]);
            cur[3] = (int)cf_index(g, pn.s[1], t, pn.j[1]);
            cur[4] = pn.nch;
        };
        if (leader) cursor_to(p0, tb);
        auto refill = [&](uint32_t st) {
            const int tn = cur[1], ca = cur[2], cb = cur[3];
            sync.dst = smem_u32(stages + st * kStageFloats);
            sync.mbar = bars + 8 * st;
            sync.nrows = cur[4];
            sync.src[0] = row_ptr(P, (size_t)ca);
            sync.src[1] = row_ptr(P, (size_t)cb);
            if (tn + 1 == g.T) { if (cur[0] + 1 < g.n_pairs) cursor_to(cur[0] + 1, 0); }
            else { cur[1] = tn + 1; cur[2] = ca + g.nc; cur[3] = cb + g.nc; }
        };
        if (leader) {  // prologue: fill the ring
            for (int i = 0; i < kStages && i < nf; ++i) {
                refill((fc + i) % kStages);
                sync.issue();
            }
        }

        Pair pr = make_pair(g, p0);
        int t = tb, pi = p0;
        // channel-frame indices and PCM offsets of the two chains advance linearly inside a pair
        size_t cfa = cf_index(g, pr.s[0], t, pr.j[0]), cfb = cf_index(g, pr.s[1], t, pr.j[1]);
        size_t oa = ((size_t)pr.s[0] * g.T + t) * 1024 * g.nc + pr.j[0], ob = ((size_t)pr.s[1] * g.T + t) * 1024 * g.nc + pr.j[1];
        for (int f = 0; f < nf; ++f, ++fc) {
            if (t == 0) {  // a pair starts here: its overlap comes from the state
                ovl_load<0>(u, P.ovl_in + state_index(g, pr.s[0], pr.j[0]), ov, P.scale);
                if (pr.nch == 2) ovl_load<1>(u, P.ovl_in + state_index(g, pr.s[1], pr.j[1]), ov, P.scale);
            }
            const uint32_t st = fc % kStages;
            FrameIO io;
            io.stage = stages + st * kStageFloats;
            io.scratch = scratch + (fc & 1u) * kStageFloats;
            io.nch = pr.nch;
            io.dst.emit = fb + f >= f0;
            io.dst.interleaved = pr.interleaved;
            io.dst.scale = P.scale;
            io.dst.inv_scale = 1.0f / P.scale;
            io.dst.ostride = g.nc;
            io.fi[0] = info_lo(P, cfa);
            io.fi[1] = info_lo(P, cfb);
            io.dst.out0 = P.pcm + oa;
            io.dst.out1 = P.pcm + ob;
            sync.next_valid = f + kStages < nf;
            if (leader && sync.next_valid) refill(st);
            mbar_wait(bars + 8 * st, (fc / kStages) & 1u);
            worker_frame<GENERIC>(u, sync, io, ts, P.tab, z, ov);
            cfa += g.nc; cfb += g.nc;
            oa += (size_t)1024 * g.nc; ob += (size_t)1024 * g.nc;
            if (++t == g.T) {  // the pair is complete: its overlap goes back to the state
                ovl_store<0>(u, ov, P.ovl_out + state_index(g, pr.s[0], pr.j[0]), 1.0f / P.scale);
                if (pr.nch == 2) ovl_store<1>(u, ov, P.ovl_out + state_index(g, pr.s[1], pr.j[1]), 1.0f / P.scale);
                t = 0;
                if (f + 1 < nf) {
                    pr = make_pair(g, ++pi);
                    cfa = cf_index(g, pr.s[0], 0, pr.j[0]); cfb = cf_index(g, pr.s[1], 0, pr.j[1]);
                    oa = (size_t)pr.s[0] * g.T * 1024 * g.nc + pr.j[0]; ob = (size_t)pr.s[1] * g.T * 1024 * g.nc + pr.j[1];
                }
            }
        }
    }
}

// One thread per channel-frame that carries TNS: write its filtered copy to
// `scratch` (every float4 exactly once).  The chain is serial by construction
// (see tns_run), parallelism is across channel-frames.
__global__ void __launch_bounds__(kTnsThreads, 8) tns_kernel(const __grid_constant__ TnsParams P) {
    const size_t cf = (size_t)blockIdx.x * kTnsThreads + threadIdx.x;
    if (cf >= P.n_cf) return;
    const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(P.info) + cf);
    if ((raw.y & 0xffu) == 0) return;
    const uint32_t o0 = P.offsets[cf], o1 = P.offsets[cf + 1];
    const bool has_block = o1 > o0 && o1 <= P.blob_bytes;
    tns_apply((FrameBits)raw.x & 0xffffff03u, P.blob + (has_block ? o0 : 0), has_block ? o1 - o0 : 0, P.sample_index,
              P.ar != 0, *P.bands, P.spectra + cf * 1024, P.scratch + cf * 1024);
}

template <bool GENERIC, int W, int ST>
static cudaError_t launch_one(const SynthParams &P, int num_sms, cudaStream_t stream) {
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(synth_kernel<GENERIC, W, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Layout<W, ST>::kTotal);
        if (e != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    const int grid = num_sms < (P.g.n_items + W - 1) / W ? num_sms : (P.g.n_items + W - 1) / W;
    synth_kernel<GENERIC, W, ST><<<grid < 1 ? 1 : grid, W * 64, Layout<W, ST>::kTotal, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_synth(const SynthParams &P, int num_sms, bool generic, cudaStream_t stream) {
    return generic ? launch_one<true, kWorkersGeneric, kStagesGeneric>(P, num_sms, stream)
                   : launch_one<false, kWorkers, kStages>(P, num_sms, stream);
}

cudaError_t launch_tns(const TnsParams &P, cudaStream_t stream) {
    const unsigned grid = (unsigned)((P.n_cf + kTnsThreads - 1) / kTnsThreads);
    if (grid == 0) return cudaSuccess;
    tns_kernel<<<grid, kTnsThreads, 0, stream>>>(P);
    return cudaGetLastError();
}

int synth_smem_bytes() { return Layout<kWorkers, kStages>::kTotal; }

}  // namespace aacfb
