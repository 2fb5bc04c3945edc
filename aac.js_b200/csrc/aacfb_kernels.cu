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
#include <cstdio>
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
// STEREO: the instantiation that also applies the stereo tools (wider side-info ring entries,
// one aacfb_stereo_ops record per stage); the plain one is byte for byte what it was without them.
template <int W, int ST, bool STEREO = false, bool IOV = false>
struct Layout {
    static constexpr int kBufsPerWorker = ST + 2;  // TMA ring + two alternating scratch buffers
    static constexpr int kOffStages = kTabBytes;
    static constexpr int kOffBars = kOffStages + W * kBufsPerWorker * kStageBytes;
    static constexpr int kOffSlots = kOffBars + W * ST * 8;
    static constexpr int kRingWords = STEREO ? 4 : 2;   // per ring entry: 2 chains x (FrameBits [, flags word])
    static constexpr int kPendInts = 12;                // the refill the leader has prepared (see DevSync)
    static constexpr int kSlotInts = 8 + kPendInts + 2 * kRingWords * ST;   // per worker: item, flag, cursor[6], pending refill, side-info ring [2 ST]
    static constexpr int kOffOps = (kOffSlots + W * kSlotInts * 4 + 15) & ~15;   // aacfb_stereo_ops per worker and stage
    static constexpr int kEndOps = STEREO ? kOffOps + W * ST * (int)sizeof(aacfb_stereo_ops) : kOffSlots + W * kSlotInts * 4;
    // IOV: SCALEFACTOR_TABLE (512 entries, NaN tail) and IQ_TABLE[0 .. kDqIqLo) for dequant_stage
    static constexpr int kOffDq = (kEndOps + 15) & ~15;
    static constexpr int kDqBytes = (512 + kDqIqLo) * 4;
    static constexpr int kTotal = IOV ? kOffDq + kDqBytes : kEndOps;
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

// PARK: keep the prepared refill in shared memory instead of registers.  Costs the leader's warp
// a few dozen instructions per frame (measured -4 % on the long-only instantiations) but frees
// ~8 registers per thread, which is what keeps the generic instantiations from spilling.
// RING_ROWS: the filtered part of a row comes from a row of `scratch` named by the refill record
// (synth_tns_kernel: a slot of the CTA's ring) instead of from the row's own index.
template <bool STEREO, bool PARK, bool IOV, bool RING_ROWS = false>
struct DevSync {
    uint32_t bar_id;      // named barrier of this worker (all 64 threads block)
    uint32_t free_id;     // named barrier "stage is free": followers arrive, the leader's warp waits
    bool leader_warp, leader;
    volatile uint32_t *ring;  // side-info ring of the worker (see synth_kernel)
    // The refill this frame's stage_free() has to issue is prepared by the leader thread before the
    // frame's arithmetic and issued after it.  Only that one thread needs it, so it is parked in
    // shared memory instead of occupying registers of all 64 threads across the whole frame:
    // pend[0] valid, [1] dst, [2] mbar, [3] nrows, [4..5] channel-frame index of the rows,
    // [6..7] rng: lo4 | hi4 << 16 = the float4 interval that comes from the TNS scratch (0: none),
    // [8] ring entry of the frame's side info, [9] shared-memory home of its stereo record,
    // [10] the rows are (left, right) of one stream.
    volatile uint32_t *pend;
    uint32_t reg[11];     // the same record in registers (!PARK)
    // [11] (RING_ROWS: scratch row of the frame's first chain) always lives in shared memory
    __device__ __forceinline__ void put(int i, uint32_t v) { if (PARK || i >= 11) pend[i] = v; else reg[i] = v; }
    __device__ __forceinline__ uint32_t get(int i) const { return (PARK || i >= 11) ? pend[i] : reg[i]; }
    const float *spectra, *scratch;
    const uint8_t *qframes;           // IOV: aacfb_qframe records instead of float rows (nullptr: float rows)
    const aacfb_stereo_ops *stereo;   // global records (nullptr: none)
    __device__ __forceinline__ void barrier() { asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory"); }
    // One row = 4096 bytes on the mbarrier, fetched as up to three 1-D bulk copies: the interval
    // the TNS pass filtered comes from its scratch, the rest straight from the spectra.
    __device__ __forceinline__ void issue_row(uint32_t d, uint32_t mbar, int cfi, uint32_t r, int srow) {
        if (IOV && qframes) {   // one 2304-byte record, landing in the upper part of the row slot (dequant_stage)
            bulk_load(d + (uint32_t)kQLandOffset, qframes + (size_t)cfi * kQFrameBytes, (uint32_t)kQFrameBytes, mbar);
            return;
        }
        const float *a = spectra + (size_t)cfi * 1024;
        const uint32_t lo = r & 0xffffu, hi = r >> 16;
        if (hi <= lo) { bulk_load(d, a, 4096u, mbar); return; }
        const float *b = scratch + (size_t)srow * 1024;
        if (lo) bulk_load(d, a, 16u * lo, mbar);
        bulk_load(d + 16u * lo, b + 4 * lo, 16u * (hi - lo), mbar);
        if (hi < 256u) bulk_load(d + 16u * hi, a + 4 * hi, 16u * (256u - hi), mbar);
    }
    // Leader only: issue the prepared refill.  STEREO: the side info of the frame must have landed
    // in the ring (stereo_present = byte 5 of the left channel's aacfb_frame_info = second word of
    // its entry).
    __device__ __forceinline__ void issue() {
        const uint32_t dst = get(1), mbar = get(2), nrows = get(3);
        bool ops = false;
        if (STEREO) ops = get(10) != 0 && ((ring[4 * get(8) + 1] >> 8) & 0xffu) != 0;
        const uint32_t row_bytes = (IOV && qframes) ? (uint32_t)kQFrameBytes : 4096u;
        mbar_expect_tx(mbar, nrows * row_bytes + (ops ? (uint32_t)sizeof(aacfb_stereo_ops) : 0u));
        issue_row(dst, mbar, (int)get(4), get(6), RING_ROWS ? (int)get(11) : (int)get(4));
        if (nrows == 2) issue_row(dst + 4096u, mbar, (int)get(5), get(7), RING_ROWS ? (int)get(11) + 1 : (int)get(5));
        if (STEREO && ops) bulk_load(get(9), stereo + (get(4) >> 1), (uint32_t)sizeof(aacfb_stereo_ops), mbar);
    }
    __device__ __forceinline__ void stage_free() {
        // order this thread's generic-proxy accesses to the stage before the async-proxy refill
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (leader_warp) {
            asm volatile("bar.sync %0, 64;" ::"r"(free_id) : "memory");
            if (leader && get(0) != 0u) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");  // the prefetched side info has landed
                issue();
            }
        } else {
            asm volatile("bar.arrive %0, 64;" ::"r"(free_id) : "memory");
        }
    }
    // value of thread 63-u: lane ^ 31 of the same warp (worker_thread_index)
    __device__ __forceinline__ float partner(int, float v) { return __shfl_xor_sync(0xffffffffu, v, 31); }
    // value of thread u ^ 7 (same window of an EIGHT_SHORT frame, mirrored position): lane ^ 7
    __device__ __forceinline__ float partner7(int, float v) { return __shfl_xor_sync(0xffffffffu, v, 7); }
};

// aacfb_frame_info is 8 bytes: one 64-bit load, low word = FrameBits, byte 4 = tns_present
__device__ __forceinline__ uint2 info_raw(const SynthParams &P, size_t cf) {
    return __ldg(reinterpret_cast<const uint2 *>(P.info) + cf);
}
__device__ __forceinline__ FrameBits info_lo(const SynthParams &P, size_t cf) { return info_raw(P, cf).x; }
// which float4 interval of row cf lives in the TNS scratch (0: the whole row comes from the spectra)
// (tns_kernel writes ranges[cf] for EVERY row, 0 for those without TNS, so tns_present need not be looked at first:
// one global load on the leader's path instead of two dependent ones)
__device__ __forceinline__ uint32_t row_range(const SynthParams &P, size_t cf) {
    return P.scratch != nullptr ? __ldg(P.ranges + cf) : 0u;
}

}  // namespace

// GENERIC = false: takes only work items without EIGHT_SHORT frames (the long-transform code
// alone, best register allocation); GENERIC = true: takes only the items that have one.
// Both instantiations are launched back to back and walk the same item list.
// IOV: the instantiations that also take aacfb_qframe input and / or write int16 PCM.
template <bool GENERIC, bool STEREO, bool IOV, int W, int ST>
__global__ void __launch_bounds__(W * 64, 1) synth_kernel(const __grid_constant__ SynthParams P) {
    using L = Layout<W, ST, STEREO, IOV>;
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
    }
    DqCtx dqc;
    if (IOV && P.qframes != nullptr) {   // lookup tables of the inverse quantisation -> shared memory
        float *dst = reinterpret_cast<float *>(smem + L::kOffDq);
        for (int i = tid; i < 512; i += kCtaThreads) dst[i] = P.dq->sf[i];
        for (int i = tid; i < kDqIqLo; i += kCtaThreads) dst[512 + i] = P.dq->iq[i];
        dqc.D = P.dq; dqc.sf = dst; dqc.iq_lo = dst + 512;
        dq_thread_consts(P.dq, u, dqc);
    }
    const SynthTables *ts = reinterpret_cast<const SynthTables *>(smem);
    float *stages = reinterpret_cast<float *>(smem + kOffStages) + (size_t)w * kBufsPerWorker * kStageFloats;
    float *scratch = stages + kStages * kStageFloats;
    const uint32_t bars = smem_u32(smem + kOffBars) + w * kStages * 8;
    // per worker: [0] item, [1] short flag, [2..6] prefetch cursor, [8..] side-info ring
    volatile int *slot = reinterpret_cast<volatile int *>(smem + kOffSlots) + L::kSlotInts * w;
    // Side info (packed aacfb_frame_info) of the frames in flight.  The leader fetches it with a
    // 4-byte cp.async when it prefetches the frame's rows, kStages frames ahead, so no thread waits
    // on a global load at the top of a frame; 2 kStages entries because the entry of the frame
    // being worked on is still being read when the next prefetch is issued.
    volatile uint32_t *fi_ring = reinterpret_cast<volatile uint32_t *>(slot + 8 + L::kPendInts);  // [entry][chain][1 or 2 words]
    constexpr uint32_t kFiRing = 2 * kStages, kRW = L::kRingWords;
    const aacfb_stereo_ops *ops_area =
        reinterpret_cast<const aacfb_stereo_ops *>(smem + L::kOffOps) + (size_t)w * kStages;
    if (leader) {
        for (int s = 0; s < kStages; ++s) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // PARK: the plain generic instantiation has had registers to spare since the product-array addresses became XORs
    // (config 5 0.2727 -> 0.2699 ms without it); the stereo / quantised-input ones still spill and keep it.
#ifndef AACFB_GENERIC_PARK
#define AACFB_GENERIC_PARK 0   // tuning switch: 1 = also the plain generic instantiation parks the refill in shared memory
#endif
    DevSync<STEREO, GENERIC && (STEREO || IOV || AACFB_GENERIC_PARK != 0), IOV> sync;
    sync.bar_id = 1 + w;
    sync.free_id = 1 + kWorkers + w;
    sync.leader_warp = ((tid >> 5) & 1) == 0;
    sync.leader = leader;
    sync.spectra = P.spectra;
    sync.scratch = P.scratch;
    sync.qframes = IOV ? P.qframes : nullptr;
    sync.stereo = P.stereo;
    sync.ring = fi_ring;
    sync.pend = reinterpret_cast<volatile uint32_t *>(slot + 8);
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
            if (STEREO) cur[5] = pn.interleaved ? 1 : 0;
        };
        if (leader) cursor_to(p0, tb);
        auto refill = [&](uint32_t st, uint32_t frame_no) {
            const int tn = cur[1], ca = cur[2], cb = cur[3];
            {
                const uint32_t ring_e = frame_no % kFiRing;
                const uint32_t e = smem_u32(const_cast<uint32_t *>(fi_ring)) + 4u * kRW * ring_e;
                const uint32_t *inf = reinterpret_cast<const uint32_t *>(P.info);
                if (STEREO) {   // the whole 8-byte record: the stereo flag sits in its second word
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(e), "l"(inf + 2 * (size_t)ca) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(e + 8u), "l"(inf + 2 * (size_t)cb) : "memory");
                } else {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(e), "l"(inf + 2 * (size_t)ca) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(e + 4u), "l"(inf + 2 * (size_t)cb) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (STEREO) {
                    sync.put(8, ring_e);
                    sync.put(9, smem_u32(ops_area + st));
                    sync.put(10, cur[5] != 0 ? 1u : 0u);
                }
            }
            sync.put(1, smem_u32(stages + st * kStageFloats));
            sync.put(2, bars + 8 * st);
            sync.put(3, (uint32_t)cur[4]);
            sync.put(4, (uint32_t)ca); sync.put(5, (uint32_t)cb);
            sync.put(6, row_range(P, (size_t)ca));
            sync.put(7, row_range(P, (size_t)cb));
            if (tn + 1 == g.T) { if (cur[0] + 1 < g.n_pairs) cursor_to(cur[0] + 1, 0); }
            else { cur[1] = tn + 1; cur[2] = ca + g.nc; cur[3] = cb + g.nc; }
        };
        if (leader) {  // prologue: fill the ring
            for (int i = 0; i < kStages && i < nf; ++i) {
                refill((fc + i) % kStages, fc + i);
                if (STEREO) asm volatile("cp.async.wait_group 0;" ::: "memory");
                sync.issue();
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        sync.barrier();  // the side info of the first frames is in place

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
            io.dst.s16 = IOV && P.pcm_s16 != 0;
            io.dq = (IOV && P.qframes != nullptr) ? &dqc : nullptr;
            io.fi[0] = fi_ring[kRW * (fc % kFiRing)];
            io.fi[1] = fi_ring[kRW * (fc % kFiRing) + kRW / 2];
            io.ops = nullptr;
            if (STEREO && pr.interleaved && ((fi_ring[4 * (fc % kFiRing) + 1] >> 8) & 0xffu) != 0) io.ops = ops_area + st;
            if (IOV && P.pcm_s16 != 0) {   // same element offsets, 2-byte samples
                io.dst.out0 = reinterpret_cast<float *>(reinterpret_cast<int16_t *>(P.pcm) + oa);
                io.dst.out1 = reinterpret_cast<float *>(reinterpret_cast<int16_t *>(P.pcm) + ob);
            } else {
                io.dst.out0 = P.pcm + oa;
                io.dst.out1 = P.pcm + ob;
            }
            if (leader) {
                const bool next_valid = f + kStages < nf;
                sync.put(0, next_valid ? 1u : 0u);
                if (next_valid) refill(st, fc + kStages);
            }
            mbar_wait(bars + 8 * st, (fc / kStages) & 1u);
            worker_frame<GENERIC, STEREO, IOV>(u, sync, io, ts, P.tab, z, ov);
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

// ------------------------------------------------------------------ TNS
// tns_kernel: TNS.process (tns.js:105-177) for every channel-frame that carries TNS.
//
// The chain of one filter run is serial by construction (each tap is a rounded f32
// read-modify-write in the reference's order), so parallelism is across channel-frames:
// one LANE per row, each WARP autonomous over 32 consecutive rows (no CTA-level sync).
// What is new against a thread-per-row loop is the data movement: a lane never touches
// global memory for its own row.  The warp moves *tiles* of 32 rows x 32 coefficients
//   global --cp.async 16 B, 4 whole 128-byte lines per instruction--> shared ring (3 tiles)
//   shared: lane r filters row r of the tile in place (LDS.128 / STS.128, pitch 36 floats:
//           conflict-free both for the row-per-lane and the line-per-quarter-warp pattern)
//   shared --LDS.128 / STG.128, whole lines--> scratch
// so every global access is a full line and the next two tiles are in flight while one is
// being filtered.  A tile's column window is per row: block b of a run that starts at
// `start` and walks in direction `inc` covers coefficients start + inc*(32b .. 32b+31).
//
// Only the filtered coefficients are written: scratch[cf] is valid on the bounding interval
// [lo4, hi4) (float4 units) of the row's runs -- coefficients inside it that no filter touches
// are copied -- and `ranges[cf]` = lo4 | hi4 << 16 tells synth_kernel which part of the row to
// fetch from scratch (the rest comes straight from the spectra).
#ifndef AACFB_TNS_COLS
#define AACFB_TNS_COLS 32   // coefficients per tile row (32: 128-byte pieces of a row, 64: 256-byte pieces)
#endif
constexpr int kTnsCols = AACFB_TNS_COLS;
constexpr int kTnsQuads = kTnsCols / 4;                 // float4s per tile row
constexpr int kTnsRowsPerCopy = 32 / kTnsQuads;         // tile rows one warp-wide 16-byte copy covers
constexpr int kTnsCopies = 32 / kTnsRowsPerCopy;        // copies per tile and lane
constexpr int kTnsPitch = kTnsCols + 4;                 // floats per tile row (+ 16 B: conflict-free, see above)
static_assert(kTnsCols == 32 || kTnsCols == 64, "tile width");
#ifndef AACFB_TNS_RING
#define AACFB_TNS_RING 3    // tiles per warp: one being filtered, AACFB_TNS_RING - 1 in flight
#endif
#ifndef AACFB_TNS_CTAS
#define AACFB_TNS_CTAS 4    // resident CTAs per SM the register budget is set for (4: 128 registers per thread)
#endif
constexpr int kTnsRing = AACFB_TNS_RING;
constexpr int kTnsAhead = kTnsRing - 1;
constexpr int kTnsTileFloats = 32 * kTnsPitch;
constexpr int kTnsSmemRing = kTnsWarps * kTnsRing * kTnsTileFloats * 4;
constexpr int kTnsSmemBytes = kTnsSmemRing + 256;      // + the band tables of one sample rate

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
#ifndef AACFB_TNS_L2_PREFETCH
#define AACFB_TNS_L2_PREFETCH 256   // tuning switch: 0 = plain
#endif
#if AACFB_TNS_L2_PREFETCH == 256
    // A tile row is one 128-byte line and the next tile of the same row is its neighbour: asking L2
    // for the whole 256-byte block turns the DRAM access into 256-byte runs (tools/ubench/stride_copy.cu:
    // 6.3 instead of 4.9 TB/s for this pattern) and the second tile then hits in L2.
    asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
#elif AACFB_TNS_L2_PREFETCH == 128
    asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
#endif
}
// the same with a source size: 0 reads nothing and zero-fills the 16 bytes
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void *src, uint32_t src_bytes) {
#if AACFB_TNS_L2_PREFETCH == 256
    asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// One run slot of the warp: lane r filters its run (start, size, inc; size = 0: none) of row r
// (x_base / y_base point at row 0 of the warp).
// ORD >= every lane's order (coefficients beyond a lane's order are zero: no-ops).
template <int ORD, bool AR>
__device__ __noinline__ void tns_tile_run(float *ring, int lane, const float *x_base, float *y_base, int start, int size,
                                          int inc, const float *lpc, int order) {
    float h[ORD], c[ORD];
#pragma unroll
    for (int i = 0; i < ORD; ++i) { h[i] = 0.f; c[i] = i < order ? lpc[i] : 0.f; }
    const bool nan_from = order == AACFB_TNS_MAX_ORDER;
    // The rows this lane moves: r = lane / kTnsQuads + kTnsRowsPerCopy * k, 16-byte column cc = lane % kTnsQuads.
    //   g_off : element offset (from row 0 of the warp) of float4 `cc` of the current block
    //   g_step: +- kTnsCols elements per block (direction of the served row's run)
    //   g_nv  : my float4 is inside the run for blocks b < g_nv (its position in run order is
    //           quad cc of an upward run, quad kTnsQuads - 1 - cc of a downward one)
    const int cc = lane & (kTnsQuads - 1), rr = lane / kTnsQuads;
    int g_off[kTnsCopies], g_step[kTnsCopies], g_nv[kTnsCopies];
    int nv_min = 1 << 30;
#pragma unroll
    for (int k = 0; k < kTnsCopies; ++k) {
        const int r = rr + kTnsRowsPerCopy * k;
        const int st_k = __shfl_sync(0xffffffffu, start, r);
        const int sz_k = __shfl_sync(0xffffffffu, size, r), in_k = __shfl_sync(0xffffffffu, inc, r);
        g_off[k] = r * 1024 + (in_k > 0 ? st_k : st_k - (kTnsCols - 1)) + 4 * cc;
        g_step[k] = in_k > 0 ? kTnsCols : -kTnsCols;
        const int quad = in_k > 0 ? cc : kTnsQuads - 1 - cc;
        g_nv[k] = sz_k > 4 * quad ? (sz_k - 4 * quad + kTnsCols - 1) / kTnsCols : 0;
        nv_min = min(nv_min, g_nv[k]);
    }
    int nblk = (size + kTnsCols - 1) / kTnsCols;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nblk = max(nblk, __shfl_xor_sync(0xffffffffu, nblk, o));
        nv_min = min(nv_min, __shfl_xor_sync(0xffffffffu, nv_min, o));
    }
    const uint32_t ring_s = smem_u32(ring) + 4u * (uint32_t)(rr * kTnsPitch + 4 * cc);
    const float *tile_rd = ring + rr * kTnsPitch + 4 * cc;  // my float4 of served row k = 0
    constexpr int kCopyStride = kTnsRowsPerCopy * kTnsPitch;   // floats between the rows of copy k and k + 1
    // Blocks b < nv_min are whole for every row of the warp: no predicates there.
    auto load_tile = [&](int b, int slot) {  // tile b -> ring slot
        if (b < nv_min) {
#pragma unroll
            for (int k = 0; k < kTnsCopies; ++k)
                cp_async16(ring_s + 4u * (uint32_t)(slot * kTnsTileFloats + k * kCopyStride),
                           x_base + (g_off[k] + kTnsAhead * g_step[k]));
        } else if (b < nblk) {
#pragma unroll
            for (int k = 0; k < kTnsCopies; ++k)
                if (b < g_nv[k])
                    cp_async16(ring_s + 4u * (uint32_t)(slot * kTnsTileFloats + k * kCopyStride),
                               x_base + (g_off[k] + kTnsAhead * g_step[k]));
        }
        cp_async_commit();
    };
    // g_off describes block b while tile b + kTnsAhead is being fetched: start it that many blocks back
#pragma unroll
    for (int k = 0; k < kTnsCopies; ++k) g_off[k] -= kTnsAhead * g_step[k];
#pragma unroll
    for (int i = 0; i < kTnsAhead; ++i) {
        load_tile(i, i);
#pragma unroll
        for (int k = 0; k < kTnsCopies; ++k) g_off[k] += g_step[k];
    }
    int slot = 0;
    for (int b = 0; b < nblk; ++b) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kTnsAhead - 1) : "memory");
        __syncwarp();  // tile b has landed for every lane; tile b-1's write-out has been read
        load_tile(b + kTnsAhead, slot >= 1 ? slot - 1 : kTnsRing - 1);  // == (b + kTnsAhead) % kTnsRing
        float *tile = ring + slot * kTnsTileFloats;
        float4 *row = reinterpret_cast<float4 *>(tile + lane * kTnsPitch);
        const int left = size - kTnsCols * b;
        if (left >= kTnsCols) {  // straight-line: the history shift is pure register renaming
            // all loads first (shared-memory latency is paid once per group of quads, not per quad)
            constexpr int kHalf = ORD > 12 ? 4 : 8;
#pragma unroll
            for (int q0 = 0; q0 < kTnsQuads; q0 += kHalf) {
                float4 v[kHalf];
#pragma unroll
                for (int q = 0; q < kHalf; ++q) v[q] = row[inc > 0 ? q0 + q : kTnsQuads - 1 - q0 - q];
#pragma unroll
                for (int q = 0; q < kHalf; ++q) v[q] = tns_quad<ORD, AR>(v[q], inc, h, c, nan_from, kTnsCols * b + 4 * (q0 + q));
#pragma unroll
                for (int q = 0; q < kHalf; ++q) row[inc > 0 ? q0 + q : kTnsQuads - 1 - q0 - q] = v[q];
            }
        } else if (left > 0) {
            for (int q = 0; 4 * q < left; ++q) {
                float4 *p = row + (inc > 0 ? q : kTnsQuads - 1 - q);
                *p = tns_quad<ORD, AR>(*p, inc, h, c, nan_from, kTnsCols * b + 4 * q);
            }
        }
        __syncwarp();
        if (b < nv_min) {
#pragma unroll
            for (int k = 0; k < kTnsCopies; ++k)
                *reinterpret_cast<float4 *>(y_base + g_off[k]) =
                    *reinterpret_cast<const float4 *>(tile_rd + slot * kTnsTileFloats + k * kCopyStride);
        } else {
#pragma unroll
            for (int k = 0; k < kTnsCopies; ++k)
                if (b < g_nv[k])
                    *reinterpret_cast<float4 *>(y_base + g_off[k]) =
                        *reinterpret_cast<const float4 *>(tile_rd + slot * kTnsTileFloats + k * kCopyStride);
        }
#pragma unroll
        for (int k = 0; k < kTnsCopies; ++k) g_off[k] += g_step[k];
        slot = slot + 1 == kTnsRing ? 0 : slot + 1;
    }
    cp_async_wait0();
    __syncwarp();
}

template <bool AR>
__device__ __forceinline__ void tns_tile_dispatch(int ord_max, float *ring, int lane, const float *x, float *y, int start,
                                                  int size, int inc, const float *lpc, int order) {
    if (ord_max <= 4) tns_tile_run<4, AR>(ring, lane, x, y, start, size, inc, lpc, order);
    else if (ord_max <= 8) tns_tile_run<8, AR>(ring, lane, x, y, start, size, inc, lpc, order);
    else if (ord_max <= 12) tns_tile_run<12, AR>(ring, lane, x, y, start, size, inc, lpc, order);
    else if (ord_max <= 16) tns_tile_run<16, AR>(ring, lane, x, y, start, size, inc, lpc, order);
    else tns_tile_run<20, AR>(ring, lane, x, y, start, size, inc, lpc, order);
}

__global__ void __launch_bounds__(kTnsWarps * 32, AACFB_TNS_CTAS) tns_kernel(const __grid_constant__ TnsParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    if (P.gate != nullptr && *P.gate == 0u) return;   // nothing was left to the pre-pass (see launch_synth_tns)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float *ring = reinterpret_cast<float *>(smem) + wid * kTnsRing * kTnsTileFloats;
    // scalefactor-band tables of this sample rate -> shared memory (tables.js:34-163, tns.js:65)
    uint16_t *bands = reinterpret_cast<uint16_t *>(smem + kTnsSmemRing);  // [0,52) long, [52,68) short, [68..70] counts
    {
        const int si = P.sample_index, t = threadIdx.x;
        if (t < 52) bands[t] = P.bands->swb_long[si][t];
        else if (t < 68) bands[t] = P.bands->swb_short[si][t - 52];
        else if (t == 68) bands[68] = P.bands->swb_long_count[si];
        else if (t == 69) bands[69] = P.bands->swb_short_count[si];
        else if (t == 70) bands[70] = P.bands->tns_max_bands[si];
    }
    const size_t cf0 = ((size_t)blockIdx.x * kTnsWarps + wid) * 32;  // row 0 of this warp
    const bool in_range = cf0 + lane < P.n_cf;
    const size_t cf = in_range ? cf0 + lane : 0;
    // side info: the loads below do not depend on each other
    uint2 raw = make_uint2(0u, 0u);
    uint32_t o0 = 0, o1 = 0;
    if (in_range) {
        raw = __ldg(reinterpret_cast<const uint2 *>(P.info) + cf);
        o0 = __ldg(P.offsets + cf); o1 = __ldg(P.offsets + cf + 1);
    }
    __syncthreads();  // band tables are in place (the only CTA-wide synchronisation)
    if (cf0 >= P.n_cf) return;
    const bool present = in_range && (raw.y & 0xffu) != 0;
    const FrameBits fi = (FrameBits)raw.x & 0xffffff03u;
    const bool has_block = present && o1 > o0 && o1 <= P.blob_bytes && (o0 & 3u) == 0;
    const uint32_t block_bytes = has_block ? o1 - o0 : 0;
    // The blocks of consecutive channel-frames are contiguous in the blob: the warp copies the
    // whole region [rlo, rhi) into its (idle) tile ring with coalesced loads and parses it there.
    uint32_t rlo = has_block ? o0 : 0xffffffffu, rhi = has_block ? o1 : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        rlo = min(rlo, __shfl_xor_sync(0xffffffffu, rlo, o));
        rhi = max(rhi, __shfl_xor_sync(0xffffffffu, rhi, o));
    }
    const bool staged = rhi > rlo && rhi - rlo <= (uint32_t)(kTnsRing * kTnsTileFloats * 4);
    auto stage_blob = [&]() {
        if (!staged) return;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(P.blob + rlo);
        uint32_t *dst = reinterpret_cast<uint32_t *>(ring);
        const uint32_t words = (rhi - rlo + 3u) >> 2;
#pragma unroll 4
        for (uint32_t i = lane; i < words; i += 32) dst[i] = __ldg(src + i);
        __syncwarp();
    };
    const uint8_t *block = staged ? reinterpret_cast<const uint8_t *>(ring) + (has_block ? o0 - rlo : 0)
                                  : P.blob + (has_block ? o0 : 0);

    // One walk over the filters: each run slot of the warp is filtered as it is found; on the
    // way every lane records which float4s of its row the runs cover and their bounding interval.
    uint32_t covered[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) covered[i] = 0u;
    int lo4 = 256, hi4 = 0;
    {
        stage_blob();
        TnsWalker wk(fi, block, block_bytes, bands, bands[68], bands + 52, bands[69], bands[70]);
        bool more = has_block;
        const float *x0 = P.spectra + cf0 * 1024;
        float *y0 = P.scratch + cf0 * 1024;
        for (bool first = true;; first = false) {
            if (!first) stage_blob();  // the previous run used the ring
            TnsFilter ft;
            ft.active = false; ft.valid = false; ft.size = 0; ft.start = 0; ft.inc = 1; ft.order = 0; ft.coef = nullptr;
            while (more) {
                ft = wk.next();
                if (!ft.valid) { more = false; ft.active = false; }
                if (ft.active || !more) break;
            }
            float lpc[AACFB_TNS_MAX_ORDER];
            int order = 0, size = 0;
            if (ft.active) {
                order = ft.order; size = ft.size;
                const int lo = ft.inc > 0 ? ft.start : ft.start - ft.size + 1;
                const int a = lo >> 2, b = (lo + ft.size) >> 2;
                lo4 = min(lo4, a); hi4 = max(hi4, b);
#pragma unroll
                for (int g = 0; g < 8; ++g) {  // bits [a, b) of the 256-bit map
                    const int s = max(a - 32 * g, 0), e = min(b - 32 * g, 32);
                    if (e > s) covered[g] |= (e - s == 32 ? 0xffffffffu : ((1u << (e - s)) - 1u) << s);
                }
                for (int i = 0; i < order; ++i) {  // reflection -> direct form, tns.js:128-140
                    const float r = -ft.coef[i];
                    lpc[i] = r;
                    for (int j = 0, len = (i + 1) >> 1; j < len; ++j) {
                        const float fwd = lpc[j], bwd = lpc[i - 1 - j];
                        lpc[j] = f_fma(r, bwd, fwd);
                        lpc[i - 1 - j] = f_fma(r, fwd, bwd);
                    }
                }
            }
            if (!__any_sync(0xffffffffu, ft.active)) break;
            __syncwarp();  // every lane has parsed its block: the ring may be reused
            int ord_max = order;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ord_max = max(ord_max, __shfl_xor_sync(0xffffffffu, ord_max, o));
            if (P.ar) tns_tile_dispatch<true>(ord_max, ring, lane, x0, y0, ft.start, size, ft.inc, lpc, order);
            else tns_tile_dispatch<false>(ord_max, ring, lane, x0, y0, ft.start, size, ft.inc, lpc, order);
        }
    }
    if (hi4 <= lo4) lo4 = hi4 = 0;
    if (in_range) P.ranges[cf] = (uint32_t)lo4 | ((uint32_t)hi4 << 16);

    // Coefficients inside the interval that no run covers are copied (whole lines, row by row;
    // only rows that have such gaps are visited).
    bool gaps = false;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const int s = max(lo4 - 32 * g, 0), e = min(hi4 - 32 * g, 32);
        const uint32_t inside = e > s ? (e - s == 32 ? 0xffffffffu : ((1u << (e - s)) - 1u) << s) : 0u;
        gaps |= (inside & ~covered[g]) != 0u;
    }
    const float4 *x4 = reinterpret_cast<const float4 *>(P.spectra);
    float4 *y4 = reinterpret_cast<float4 *>(P.scratch);
    for (uint32_t todo = __ballot_sync(0xffffffffu, gaps); todo; todo &= todo - 1) {
        const int r = __ffs(todo) - 1;
        const int lo_r = __shfl_sync(0xffffffffu, lo4, r), hi_r = __shfl_sync(0xffffffffu, hi4, r);
        const size_t cf_r = cf0 + r;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const uint32_t cv = __shfl_sync(0xffffffffu, covered[g], r);
            const int idx = 32 * g + lane;
            if (!((cv >> lane) & 1u) && idx >= lo_r && idx < hi_r) y4[cf_r * 256 + idx] = x4[cf_r * 256 + idx];
        }
    }
}

// ------------------------------------------------------------------ TNS inside the synthesis kernel
// synth_tns_kernel: the pre-pass above costs a round trip of every filtered coefficient through HBM
// (config 4: 1.67 x the algorithmic traffic).  Here the last worker of every CTA does not synthesise:
// it is the CTA's *TNS worker* and filters the rows its kClients = W - 1 neighbours (the *clients*)
// are about to transform, a block of kFusedBlockFrames frames per client at a time.
//
//   * One lane per FRAME: the lane carries the two rows of a pair-frame through the serial chain at
//     once, packed (fma.rn.f32x2: one FFMA2 per tap for both rows, per-lane IEEE rounding, so the
//     result is bit for bit that of tns_kernel).  60 of the worker's 64 lanes are busy: 120 rows per
//     block, and a block takes about as long as a single row (the chain is latency-bound).
//   * Rows move as in tns_kernel, tile by tile (32 coefficients), but every tile row belongs to one
//     lane alone: the lane fetches it with a 1-D bulk copy (TMA) into a padded slot (pitch 144 B:
//     the row-per-lane LDS.128 / STS.128 are conflict-free), filters it in place and sends it on
//     with a bulk store; two stages per row, the refill of a stage is issued in the middle of the
//     next tile.  No address arithmetic per 16 bytes, no cross-lane hand-over inside a run.
//   * The filtered interval of a row goes to a slot of the CTA's *ring* in global memory (2 blocks x
//     128 rows; written and read back within ~50 us, so it lives in L2) and the client's TMA refill
//     assembles the row from it and from the spectra exactly as it does after the pre-pass
//     (DevSync<RING_ROWS>).  Block b + 1 is filtered while the clients transform block b.
//   * ready[slot] (TNS worker -> clients) and consumed[slot] (clients -> TNS worker) are words in
//     shared memory; every client has exactly ONE item (the host sizes the slices for kClients x
//     grid workers), so both sides derive the same schedule from the item geometry alone.
//   * Items with an EIGHT_SHORT frame are skipped by both sides and counted in P.short_items: the
//     gated pre-pass + generic instantiation that follow take them (launch order in aacfb_api.cu).
namespace {

constexpr int kFusedBlockFrames = 12;
constexpr int kFTilePitch = 144;                        // bytes per tile row: 128 + 16
constexpr int kFTileWarpBytes = 64 * kFTilePitch;       // one stage of one warp: rows A of its 32 lanes, then rows B
constexpr int kFBlobCap = 2 * kFTileWarpBytes / 64;     // bytes per row when the blocks are parsed in the idle tile stages
static_assert(kFusedClients * kFusedBlockFrames <= 64 && 2 * 64 <= kFusedRingRows, "lanes / ring rows");

template <int W, int ST>
struct FusedLayout {
    using Base = Layout<W - 1, ST>;                     // the clients' part is that of a (W - 1)-worker synth_kernel
    static constexpr int kOffTiles = (Base::kTotal + 127) & ~127;        // 2 warps x 2 stages
    static constexpr int kOffTBars = kOffTiles + 4 * kFTileWarpBytes;    // full[warp][stage]
    static constexpr int kOffRng = kOffTBars + 4 * 8;                    // u32 [2][kFusedRingRows]: interval of each ring row
    static constexpr int kOffFlags = kOffRng + 2 * kFusedRingRows * 4;   // ready[2], consumed[2]
    static constexpr int kOffPlan = kOffFlags + 16;                      // per client: first frame (halo included), frames
    static constexpr int kOffBands = kOffPlan + (W - 1) * 8;             // band tables of the sample rate (71 x u16)
    static constexpr int kTotal = kOffBands + 160;
    static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

// Waits of the fused kernel's hand-shakes are bounded: a protocol error traps (the launch fails with an
// error the C-ABI reports) instead of hanging the device.
constexpr uint32_t kSpinLimit = 1u << 24;
#ifndef AACFB_SPIN_NS
#define AACFB_SPIN_NS 64
#endif
__device__ __forceinline__ void spin_until_ge(uint32_t addr, uint32_t want);
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    uint32_t done, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > kSpinLimit) __trap();
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t lds_acquire(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void spin_until_ge(uint32_t addr, uint32_t want) {
    uint32_t spins = 0;
    while (lds_acquire(addr) < want) {
        __nanosleep(AACFB_SPIN_NS);
        if (++spins > kSpinLimit / 64) __trap();
    }
}
__device__ __forceinline__ void sts_release(uint32_t addr, uint32_t v) {
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add(uint32_t addr, uint32_t v) {
    asm volatile("red.release.cta.shared::cta.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Four coefficients of both rows through the serial chain.  UP: both runs walk upward.
template <int ORD, bool AR, bool UP>
__device__ __forceinline__ void pair_quad(float4 &a, float4 &b, int ia, int ib, F2 (&h)[ORD], const F2 (&c)[ORD], bool nana,
                                          bool nanb, int m0) {
    float xa[4], xb[4];
    if (UP || ia > 0) { xa[0] = a.x; xa[1] = a.y; xa[2] = a.z; xa[3] = a.w; }
    else              { xa[0] = a.w; xa[1] = a.z; xa[2] = a.y; xa[3] = a.x; }
    if (UP || ib > 0) { xb[0] = b.x; xb[1] = b.y; xb[2] = b.z; xb[3] = b.w; }
    else              { xb[0] = b.w; xb[1] = b.z; xb[2] = b.y; xb[3] = b.x; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        F2 acc{xa[j], xb[j]};
#pragma unroll
        for (int i = 0; i < ORD; ++i) acc = f_fma(h[i], c[i], acc);   // c holds -lpc in the all-pole branch
        if (!AR) {   // MA branch, order 20: tmp[20] reads undefined -> NaN once m >= 20 (tns.js:43,169)
            if (nana && m0 + j >= AACFB_TNS_MAX_ORDER) acc.x = NAN;
            if (nanb && m0 + j >= AACFB_TNS_MAX_ORDER) acc.y = NAN;
        }
        const F2 push = AR ? acc : F2{xa[j], xb[j]};
#pragma unroll
        for (int i = ORD - 1; i > 0; --i) h[i] = h[i - 1];
        h[0] = push;
        xa[j] = acc.x; xb[j] = acc.y;
    }
    if (UP || ia > 0) { a.x = xa[0]; a.y = xa[1]; a.z = xa[2]; a.w = xa[3]; }
    else              { a.x = xa[3]; a.y = xa[2]; a.z = xa[1]; a.w = xa[0]; }
    if (UP || ib > 0) { b.x = xb[0]; b.y = xb[1]; b.z = xb[2]; b.w = xb[3]; }
    else              { b.x = xb[3]; b.y = xb[2]; b.z = xb[1]; b.w = xb[0]; }
}

// One run of row A and one of row B of every lane (size 0: none), filtered side by side.
// A tile = 32 coefficients of the warp's 64 rows (rows A of its lanes, then rows B), each row at its own
// column window.  Tiles move like those of tns_kernel: cp.async 16 B per lane, four whole 128-byte lines per
// instruction, into a two-stage ring (pitch 144 B: conflict-free for the line-per-quarter-warp copy and for the
// row-per-lane LDS.128 / STS.128 of the filter); the filtered tile leaves with whole-line STG.128 to the ring rows.
// (Per-lane bulk copies -- every tile row its own 128-byte TMA transfer -- were measured first: the copy engine
// serialises them, 4 x slower than the pre-pass.)
//   warp_g    : the warp's two stages
//   cfa, cfb  : channel-frame index of the lane's rows (rows of `spectra`)
//   ring_row0 : ring row of lane 0's row A (lane i: + 2 i, row B: + 1)
// reflection -> direct form (tns.js:128-140) in registers: the loops are unrolled to ORD and predicated on the
// row's own order (a local array here would live in local memory, behind an L1 that the staging buffers have
// squeezed to a few KiB: measured 45 000 cycles per row).
template <int ORD>
__device__ __forceinline__ void tns_lpc_regs(const float *coef, int order, float (&l)[ORD]) {
#pragma unroll
    for (int i = 0; i < ORD; ++i) l[i] = 0.f;
#pragma unroll
    for (int i = 0; i < ORD; ++i) {
        if (i < order) {
            const float r = -coef[i];
            l[i] = r;
#pragma unroll
            for (int j = 0; j < (i + 1) / 2; ++j) {
                const float fwd = l[j], bwd = l[i - 1 - j];
                l[j] = f_fma(r, bwd, fwd);
                l[i - 1 - j] = f_fma(r, fwd, bwd);
            }
        }
    }
}

template <int ORD, bool AR, bool UP>
__device__ __noinline__ void tns_pair_run(uint8_t *warp_g, int lam, const float *spectra, float *ring, uint32_t cfa, uint32_t cfb,
                                          uint32_t ring_row0, int sa, int na, int ia, int sb, int nb, int ib,
                                          const float *coefa, int oa, const float *coefb, int ob) {
    F2 h[ORD], c[ORD];
    {
        float la[ORD], lb[ORD];
        tns_lpc_regs<ORD>(coefa, oa, la);
        tns_lpc_regs<ORD>(coefb, ob, lb);
#pragma unroll
        for (int i = 0; i < ORD; ++i) {
            h[i] = F2{0.f, 0.f};
            c[i] = AR ? F2{-la[i], -lb[i]} : F2{la[i], lb[i]};   // orders below ORD: zero coefficients, no-ops
        }
    }
    const bool nana = oa == AACFB_TNS_MAX_ORDER, nanb = ob == AACFB_TNS_MAX_ORDER;
    int nblk = max((na + 31) >> 5, (nb + 31) >> 5);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nblk = max(nblk, __shfl_xor_sync(0xffffffffu, nblk, o));
    // Copy instruction k moves float4 `qi` of tile row 4 k + ol: row A (k < 8) / row B (k >= 8) of lane 4 (k & 7) + ol.
    const int qi = lam & 7, ol = lam >> 3;
    uint32_t cfk[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        cfk[k] = __shfl_sync(0xffffffffu, cfa, 4 * k + ol);
        cfk[8 + k] = __shfl_sync(0xffffffffu, cfb, 4 * k + ol);
    }
    // window of tile t of a run, as its owner publishes it: (column of the window's lowest float4 + 32) | float4s << 12 | upward << 16
    auto window = [&](int t, int start, int size, int inc) -> uint32_t {
        const int left = size - 32 * t;
        const int n4 = left <= 0 ? 0 : (left >= 32 ? 8 : left >> 2);
        const bool up = UP || inc > 0;
        const int col_lo = up ? start + 32 * t : start - 32 * t - 31;
        // (a run that is over -- n4 = 0 -- has a meaningless, possibly negative column: keep it inside its field)
        return ((uint32_t)(col_lo + 32) & 0xfffu) | ((uint32_t)n4 << 12) | (up ? 1u << 16 : 0u);
    };
    const uint32_t tile_s = smem_u32(warp_g);
    // The movers first collect the 16 windows (independent shuffles), then issue 16 independent copies without a
    // branch per row, so that their latencies overlap -- with each other and with the filter chain they are
    // scheduled next to (one basic block per tile).
    auto windows = [&](int t, uint32_t (&wi)[16]) {
        const uint32_t wa = window(t, sa, na, ia), wb = window(t, sb, nb, ib);
#pragma unroll
        for (int k = 0; k < 16; ++k) wi[k] = __shfl_sync(0xffffffffu, k < 8 ? wa : wb, 4 * (k & 7) + ol);
    };
    auto load = [&](int t, int s) {
        uint32_t wi[16];
        windows(t, wi);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int n4 = (int)((wi[k] >> 12) & 15u), col = (int)(wi[k] & 0xfffu) - 32 + 4 * qi;
            const bool valid = (wi[k] >> 16) ? qi < n4 : qi >= 8 - n4;
            // src-size 0: nothing is read (the 16 bytes are zero-filled), so no branch around the copy
            cp_async16_zfill(tile_s + (uint32_t)(s * kFTileWarpBytes + (4 * k + ol) * kFTilePitch + 16 * qi),
                             spectra + (valid ? ((size_t)cfk[k] << 10) + col : (size_t)0), valid ? 16u : 0u);
        }
        cp_async_commit();
    };
    auto store = [&](int t, int s) {
        uint32_t wi[16];
        windows(t, wi);
#pragma unroll
        for (int k0 = 0; k0 < 16; k0 += 8) {
            float4 v[8];
#pragma unroll
            for (int k = k0; k < k0 + 8; ++k)
                v[k - k0] = *reinterpret_cast<const float4 *>(warp_g + s * kFTileWarpBytes + (4 * k + ol) * kFTilePitch + 16 * qi);
#pragma unroll
            for (int k = k0; k < k0 + 8; ++k) {
                const int n4 = (int)((wi[k] >> 12) & 15u), col = (int)(wi[k] & 0xfffu) - 32 + 4 * qi;
                const bool valid = (wi[k] >> 16) ? qi < n4 : qi >= 8 - n4;
                const size_t row = ring_row0 + 2u * (uint32_t)(4 * (k & 7) + ol) + (uint32_t)(k >> 3);
                if (valid) __stcg(reinterpret_cast<float4 *>(ring + ((row << 10) + col)), v[k - k0]);
            }
        }
    };
    auto filter = [&](int t, int s) {
        float4 *ra = reinterpret_cast<float4 *>(warp_g + s * kFTileWarpBytes + lam * kFTilePitch);
        float4 *rb = reinterpret_cast<float4 *>(warp_g + s * kFTileWarpBytes + (32 + lam) * kFTilePitch);
        const int lefta = na - 32 * t, leftb = nb - 32 * t;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float4 va[4], vb[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int qq = 4 * half + q;
                va[q] = ra[(UP || ia > 0) ? qq : 7 - qq];
                vb[q] = rb[(UP || ib > 0) ? qq : 7 - qq];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) pair_quad<ORD, AR, UP>(va[q], vb[q], ia, ib, h, c, nana, nanb, 32 * t + 16 * half + 4 * q);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int qq = 4 * half + q;
                if (4 * qq < lefta) ra[(UP || ia > 0) ? qq : 7 - qq] = va[q];
                if (4 * qq < leftb) rb[(UP || ib > 0) ? qq : 7 - qq] = vb[q];
            }
        }
    };
#ifdef AACFB_FUSED_TRACE
    const long long tc0 = clock64();
#endif
    __syncwarp();   // the blocks were parsed in this memory: every lane is done with its copy
    if (nblk > 0) {
        // Tile t is filtered in stage t & 1 while, in the same basic block, tile t - 1 leaves the other stage and
        // tile t + 1 is fetched into it (the same lane reads and then refills a given 16 bytes, in program order).
        load(0, 0);
        cp_async_wait0();
        __syncwarp();
        if (nblk > 1) load(1, 1);
        filter(0, 0);
        cp_async_wait0();
        __syncwarp();
        for (int t = 1; t + 1 < nblk; ++t) {
            const int s = t & 1;
            store(t - 1, s ^ 1);
            load(t + 1, s ^ 1);
            filter(t, s);
            cp_async_wait0();
            __syncwarp();   // tile t + 1 has landed for every lane and tile t is filtered
        }
        if (nblk > 1) {
            const int t = nblk - 1, s = t & 1;
            store(t - 1, s ^ 1);
            filter(t, s);
            __syncwarp();
        }
        store(nblk - 1, (nblk - 1) & 1);
    }
    __syncwarp();   // the stages may be reused (next run: parsed blocks of other lanes)
#ifdef AACFB_FUSED_TRACE
    if (blockIdx.x == 0 && (threadIdx.x & 63) == 0) printf("  run of %d tiles: %lld cycles\n", nblk, clock64() - tc0);
#endif
}

template <bool AR, bool UP>
__device__ __forceinline__ void tns_pair_dispatch(int ord_max, uint8_t *warp_g, int lam, const float *spectra, float *ring,
                                                  uint32_t cfa, uint32_t cfb, uint32_t ring_row0, int sa, int na, int ia,
                                                  int sb, int nb, int ib, const float *coefa, int oa, const float *coefb, int ob) {
    if (ord_max <= 8) tns_pair_run<8, AR, UP>(warp_g, lam, spectra, ring, cfa, cfb, ring_row0, sa, na, ia, sb, nb, ib, coefa, oa, coefb, ob);
    else if (ord_max <= 12) tns_pair_run<12, AR, UP>(warp_g, lam, spectra, ring, cfa, cfb, ring_row0, sa, na, ia, sb, nb, ib, coefa, oa, coefb, ob);
    else tns_pair_run<20, AR, UP>(warp_g, lam, spectra, ring, cfa, cfb, ring_row0, sa, na, ia, sb, nb, ib, coefa, oa, coefb, ob);
}

// The TNS worker of a CTA (64 threads).
template <int W, int ST>
__device__ __noinline__ void tns_server(const SynthParams &P, uint8_t *smem) {
    using L = FusedLayout<W, ST>;
    constexpr int kClients = W - 1;
    const int l = threadIdx.x & 63, wv = l >> 5, lam = l & 31;
    auto server_barrier = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(W) : "memory"); };   // ids 1..W-1: clients
    uint16_t *bands = reinterpret_cast<uint16_t *>(smem + L::kOffBands);   // [0,52) long, [52,68) short, [68..70] counts
    for (int t = l; t < 71; t += 64) {
        const int si = P.sample_index;
        bands[t] = t < 52 ? P.tns_bands->swb_long[si][t] : t < 68 ? P.tns_bands->swb_short[si][t - 52]
                 : t == 68 ? P.tns_bands->swb_long_count[si] : t == 69 ? P.tns_bands->swb_short_count[si]
                           : P.tns_bands->tns_max_bands[si];
    }
    const volatile int *plan = reinterpret_cast<const volatile int *>(smem + L::kOffPlan);
    const int cw = l / kFusedBlockFrames, ci = l - cw * kFusedBlockFrames;
    const bool lane_on = cw < kClients;
    const int fb = lane_on ? plan[2 * cw] : 0, nf = lane_on ? plan[2 * cw + 1] : 0;
    int nf_max = 0;
    for (int c = 0; c < kClients; ++c) nf_max = max(nf_max, plan[2 * c + 1]);
    const int nblocks = (nf_max + kFusedBlockFrames - 1) / kFusedBlockFrames;
    server_barrier();   // band tables in place

    uint8_t *warp_g = smem + L::kOffTiles + wv * 2 * kFTileWarpBytes;     // the warp's two stages
    volatile uint32_t *rng = reinterpret_cast<volatile uint32_t *>(smem + L::kOffRng);
    const uint32_t flags = smem_u32(smem + L::kOffFlags);   // ready[2] | consumed[2]
    const Geometry g = P.g;
    const bool ar = P.tns_ar != 0;

#ifdef AACFB_FUSED_TRACE
#define AACFB_TR(i) if (blockIdx.x == 0 && l == 0) tr[i] = clock64()
    long long tr[6] = {0, 0, 0, 0, 0, 0};
#else
#define AACFB_TR(i)
#endif
    for (int b = 0; b < nblocks; ++b) {
        const int slot = b & 1;
        AACFB_TR(0);
        if (b >= 2) {   // the clients have fetched the last rows of the block that used this slot
            if (l == 0) {
                uint32_t need = 0;
                for (int c = 0; c < kClients; ++c) need += plan[2 * c + 1] > kFusedBlockFrames * (b - 2) ? 1u : 0u;
                spin_until_ge(flags + 8u + 4u * slot, need);
                sts_release(flags + 8u + 4u * slot, 0u);
            }
            server_barrier();
        }
        AACFB_TR(1);
        const int n = kFusedBlockFrames * b + ci;
        const bool on = n < nf;
        size_t cf[2] = {0, 0};
        if (on) {
            const int ff = fb + n, pi = ff / g.T, t = ff - pi * g.T;
            const Pair pp = make_pair(g, pi);
            cf[0] = cf_index(g, pp.s[0], t, pp.j[0]);
            cf[1] = cf_index(g, pp.s[1], t, pp.j[1]);
        }
        const int rr = slot * kFusedRingRows + 2 * l;
        const uint32_t ring_row0 = (uint32_t)blockIdx.x * 2 * kFusedRingRows + (uint32_t)(slot * kFusedRingRows + 64 * wv);
        float *y[2];
        y[0] = P.tns_ring + ((size_t)blockIdx.x * 2 * kFusedRingRows + rr) * 1024;
        y[1] = y[0] + 1024;
        const float *x[2] = {P.spectra + cf[0] * 1024, P.spectra + cf[1] * 1024};

        // ---- side info of my two rows
        uint2 raw[2] = {make_uint2(0u, 0u), make_uint2(0u, 0u)};
        uint32_t o0[2] = {0u, 0u}, o1[2] = {0u, 0u};
        if (on) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                raw[c] = __ldg(reinterpret_cast<const uint2 *>(P.info) + cf[c]);
                o0[c] = __ldg(P.tns_offsets + cf[c]);
                o1[c] = __ldg(P.tns_offsets + cf[c] + 1);
            }
        }
        bool has_block[2];
        uint32_t bytes[2];
        const uint8_t *block[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const bool present = on && (raw[c].y & 0xffu) != 0;
            has_block[c] = present && o1[c] > o0[c] && o1[c] <= P.tns_blob_bytes && (o0[c] & 3u) == 0;
            bytes[c] = has_block[c] ? o1[c] - o0[c] : 0u;
        }
        // blocks are parsed from a private copy in the (idle) tile stages: kFBlobCap bytes per row
        auto stage_blob = [&]() {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t *dst = reinterpret_cast<uint32_t *>(warp_g + (2 * lam + c) * kFBlobCap);
                block[c] = P.tns_blob + (has_block[c] ? o0[c] : 0u);
                if (bytes[c] != 0u && bytes[c] <= (uint32_t)kFBlobCap) {
                    const uint32_t *src = reinterpret_cast<const uint32_t *>(P.tns_blob + o0[c]);
                    const uint32_t words = (bytes[c] + 3u) >> 2;
#pragma unroll 8
                    for (uint32_t i = 0; i < words; ++i) dst[i] = __ldg(src + i);
                    block[c] = reinterpret_cast<const uint8_t *>(dst);
                }
            }
        };
        stage_blob();
        TnsWalker wk0((FrameBits)raw[0].x & 0xffffff03u, block[0], bytes[0], bands, bands[68], bands + 52, bands[69], bands[70]);
        TnsWalker wk1((FrameBits)raw[1].x & 0xffffff03u, block[1], bytes[1], bands, bands[68], bands + 52, bands[69], bands[70]);
        bool more[2] = {has_block[0], has_block[1]};
        int lo[2] = {1024, 1024}, hi[2] = {0, 0}, covered[2] = {0, 0};
        AACFB_TR(2);
        for (bool first = true;; first = false) {
            if (!first) {
                if (!__any_sync(0xffffffffu, more[0] || more[1])) break;
                stage_blob();   // the run before this one used the stages
                wk0.block = block[0]; wk1.block = block[1];
            }
            TnsFilter ft[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                ft[c].active = false; ft[c].valid = false; ft[c].size = 0; ft[c].start = 0; ft[c].inc = 1; ft[c].order = 0; ft[c].coef = nullptr;
                while (more[c]) {
                    ft[c] = c == 0 ? wk0.next() : wk1.next();
                    if (!ft[c].valid) { more[c] = false; ft[c].active = false; }
                    if (ft[c].active || !more[c]) break;
                }
                if (ft[c].active) {
                    const int a = ft[c].inc > 0 ? ft[c].start : ft[c].start - ft[c].size + 1;
                    lo[c] = min(lo[c], a); hi[c] = max(hi[c], a + ft[c].size); covered[c] += ft[c].size;
                } else {
                    ft[c].size = 0; ft[c].order = 0;
                }
            }
            if (!__any_sync(0xffffffffu, ft[0].active || ft[1].active)) break;
            int ord_max = max(ft[0].order, ft[1].order);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ord_max = max(ord_max, __shfl_xor_sync(0xffffffffu, ord_max, o));
            const bool up = __all_sync(0xffffffffu, ft[0].inc > 0 && ft[1].inc > 0);
#define AACFB_PAIR_ARGS warp_g, lam, P.spectra, P.tns_ring, (uint32_t)cf[0], (uint32_t)cf[1], ring_row0, ft[0].start, ft[0].size, \
                        ft[0].inc, ft[1].start, ft[1].size, ft[1].inc, ft[0].coef, ft[0].order, ft[1].coef, ft[1].order
            if (ar) { if (up) tns_pair_dispatch<true, true>(ord_max, AACFB_PAIR_ARGS); else tns_pair_dispatch<true, false>(ord_max, AACFB_PAIR_ARGS); }
            else    { if (up) tns_pair_dispatch<false, true>(ord_max, AACFB_PAIR_ARGS); else tns_pair_dispatch<false, false>(ord_max, AACFB_PAIR_ARGS); }
#undef AACFB_PAIR_ARGS
        }
        AACFB_TR(3);
        // ---- which part of each ring row is valid; coefficients inside it that no run covers are copied
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t r = 0u;
            if (hi[c] > lo[c]) {
                r = (uint32_t)(lo[c] >> 2) | ((uint32_t)(hi[c] >> 2) << 16);
                if (covered[c] != hi[c] - lo[c]) {   // rare: a filter of order 0 between two active ones
                    uint32_t cv[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                    TnsWalker wk((FrameBits)raw[c].x & 0xffffff03u, P.tns_blob + o0[c], bytes[c], bands, bands[68], bands + 52, bands[69], bands[70]);
                    for (TnsFilter f = wk.next(); f.valid; f = wk.next()) {
                        if (!f.active) continue;
                        const int a = f.inc > 0 ? f.start : f.start - f.size + 1;
                        for (int q = a >> 2; q < (a + f.size) >> 2; ++q) cv[q >> 5] |= 1u << (q & 31);
                    }
                    const float4 *x4 = reinterpret_cast<const float4 *>(x[c]);
                    float4 *y4 = reinterpret_cast<float4 *>(y[c]);
                    for (int q = lo[c] >> 2; q < hi[c] >> 2; ++q)
                        if (!((cv[q >> 5] >> (q & 31)) & 1u)) {
                            y4[q] = x4[q];
                        }
                }
            }
            rng[rr + c] = r;
        }
        __threadfence();
        fence_async_all();   // my stores to the ring are ordered before the clients' bulk loads (async proxy)
        AACFB_TR(4);
        server_barrier();
        if (l == 0) sts_release(flags + 4u * slot, (uint32_t)(b + 1));
#ifdef AACFB_FUSED_TRACE
        AACFB_TR(5);
        if (blockIdx.x == 0 && l == 0)
            printf("block %d: wait %lld  side info %lld  runs %lld  tail %lld  publish %lld cycles\n", b, tr[1] - tr[0],
                   tr[2] - tr[1], tr[3] - tr[2], tr[4] - tr[3], tr[5] - tr[4]);
#endif
    }
#undef AACFB_TR
}

}  // namespace

template <int W, int ST>
__global__ void __launch_bounds__(W * 64, 1) synth_tns_kernel(const __grid_constant__ SynthParams P) {
    using L = FusedLayout<W, ST>;
    using LB = typename L::Base;
    constexpr int kClients = W - 1, kStagesK = ST, kBufsPerWorker = LB::kBufsPerWorker;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, w = tid >> 6;
    const int u = worker_thread_index((tid >> 5) & 1, tid & 31);
    const bool leader = u == 0;
    {
        const float4 *src = reinterpret_cast<const float4 *>(P.tab);
        float4 *dst = reinterpret_cast<float4 *>(smem);
        for (int i = tid; i < kSmemTableBytes / 16; i += W * 64) dst[i] = src[i];
    }
    const Geometry g = P.g;
    volatile int *plan = reinterpret_cast<volatile int *>(smem + L::kOffPlan);
    volatile int *slot = reinterpret_cast<volatile int *>(smem + LB::kOffSlots) + LB::kSlotInts * (w < kClients ? w : 0);
    // ---- every client has one item, assigned statically; classify it
    int tb = 0, p0 = 0, nf = 0;
    bool halo = false;   // the item's first frame only rebuilds the overlap
    if (w < kClients) {
        int f0 = 0, fb = 0;
        const int item = w * (int)gridDim.x + (int)blockIdx.x;
        if (leader) slot[1] = 0;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + w) : "memory");
        if (item < g.n_items) {
            f0 = item_begin(g, item);
            const int f1 = item_end(g, item);
            p0 = f0 / g.T;
            const int t0 = f0 - p0 * g.T;
            fb = t0 != 0 ? f0 - 1 : f0;
            tb = t0 != 0 ? t0 - 1 : 0;
            halo = t0 != 0;
            nf = f1 - fb;
            bool mine = false;
            for (int f = tid & 63; f < nf; f += 64) {
                const int ff = fb + f, pi = ff / g.T;
                const Pair pp = make_pair(g, pi);
                const int t = ff - pi * g.T;
                mine |= is_short(info_lo(P, cf_index(g, pp.s[0], t, pp.j[0])));
                if (pp.nch == 2) mine |= is_short(info_lo(P, cf_index(g, pp.s[1], t, pp.j[1])));
            }
            if (__any_sync(0xffffffffu, mine) && (tid & 31) == 0) slot[1] = 1;
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + w) : "memory");
        if (slot[1] != 0) {   // left to the pre-pass + generic instantiation
            if (leader) atomicAdd(P.short_items, 1u);
            nf = 0;
        }
        if (leader) { plan[2 * w] = fb; plan[2 * w + 1] = nf; }
    }
    const uint32_t bars = smem_u32(smem + LB::kOffBars) + (w < kClients ? w : 0) * kStagesK * 8;
    if (w < kClients && leader) {
        for (int s = 0; s < kStagesK; ++s) mbar_init(bars + 8 * s, 1);
    }
    if (tid == kClients * 64) {
        volatile uint32_t *fl = reinterpret_cast<volatile uint32_t *>(smem + L::kOffFlags);
        fl[0] = fl[1] = fl[2] = fl[3] = 0u;
    }
    if (leader) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (w == kClients) { tns_server<W, ST>(P, smem); return; }
    if (nf == 0) return;

    // ---- client: the frame loop of synth_kernel<false, false, false> over its one item
    const SynthTables *ts = reinterpret_cast<const SynthTables *>(smem);
    float *stages = reinterpret_cast<float *>(smem + LB::kOffStages) + (size_t)w * kBufsPerWorker * kStageFloats;
    float *scratch = stages + kStagesK * kStageFloats;
    volatile uint32_t *fi_ring = reinterpret_cast<volatile uint32_t *>(slot + 8 + LB::kPendInts);
    constexpr uint32_t kFiRing = 2 * kStagesK, kRW = LB::kRingWords;
    DevSync<false, false, false, true> sync;
    sync.bar_id = 1 + w;
    sync.free_id = 1 + W + w;
    sync.leader_warp = ((tid >> 5) & 1) == 0;
    sync.leader = leader;
    sync.spectra = P.spectra;
    sync.scratch = P.tns_ring + (size_t)blockIdx.x * 2 * kFusedRingRows * 1024;
    sync.qframes = nullptr;
    sync.stereo = nullptr;
    sync.ring = fi_ring;
    sync.pend = reinterpret_cast<volatile uint32_t *>(slot + 8);
    Pts z;
    Ovl ov;
    // prefetch cursor (leader only): [0] pair, [1] t, [2],[3] channel-frame indices, [4] chains, [5] frame number in the item
    volatile int *cur = slot + 2;
    auto cursor_to = [&](int pair, int t) {
        const Pair pn = make_pair(g, pair);
        cur[0] = pair; cur[1] = t;
        cur[2] = (int)cf_index(g, pn.s[0], t, pn.j[0]);
        cur[3] = (int)cf_index(g, pn.s[1], t, pn.j[1]);
        cur[4] = pn.nch;
    };
    if (leader) { cursor_to(p0, tb); cur[5] = 0; }
    auto refill = [&](uint32_t st, uint32_t frame_no) {
        const int tn = cur[1], ca = cur[2], cb = cur[3], n = cur[5];
        {
            const uint32_t e = smem_u32(const_cast<uint32_t *>(fi_ring)) + 4u * kRW * (frame_no % kFiRing);
            const uint32_t *inf = reinterpret_cast<const uint32_t *>(P.info);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(e), "l"(inf + 2 * (size_t)ca) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(e + 4u), "l"(inf + 2 * (size_t)cb) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        // the frame's filtered rows: slot (block & 1) of the ring, once the TNS worker has published the block
        const int blk = n / kFusedBlockFrames, bslot = blk & 1;
        if (n - blk * kFusedBlockFrames == 0) {
            const uint32_t ready = smem_u32(smem + L::kOffFlags) + 4u * bslot;
            spin_until_ge(ready, (uint32_t)(blk + 1));
            fence_async_all();
        }
        const volatile uint32_t *rng = reinterpret_cast<const volatile uint32_t *>(smem + L::kOffRng);
        const int rr = bslot * kFusedRingRows + 2 * (w * kFusedBlockFrames + (n - blk * kFusedBlockFrames));
        sync.put(1, smem_u32(stages + st * kStageFloats));
        sync.put(2, bars + 8 * st);
        sync.put(3, (uint32_t)cur[4]);
        sync.put(4, (uint32_t)ca); sync.put(5, (uint32_t)cb);
        sync.put(6, rng[rr]);
        sync.put(7, rng[rr + 1]);
        sync.put(11, (uint32_t)rr);
        cur[5] = n + 1;
        if (tn + 1 == g.T) { if (cur[0] + 1 < g.n_pairs) cursor_to(cur[0] + 1, 0); }
        else { cur[1] = tn + 1; cur[2] = ca + g.nc; cur[3] = cb + g.nc; }
    };
    uint32_t fc = 0;
    if (leader) {
        for (int i = 0; i < kStagesK && i < nf; ++i) {
            refill((uint32_t)i % kStagesK, (uint32_t)i);
            sync.issue();
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    sync.barrier();
    Pair pr = make_pair(g, p0);
    int t = tb, pi = p0;
    size_t oa = ((size_t)pr.s[0] * g.T + t) * 1024 * g.nc + pr.j[0], ob = ((size_t)pr.s[1] * g.T + t) * 1024 * g.nc + pr.j[1];
    for (int f = 0; f < nf; ++f, ++fc) {
        if (t == 0) {
            ovl_load<0>(u, P.ovl_in + state_index(g, pr.s[0], pr.j[0]), ov, P.scale);
            if (pr.nch == 2) ovl_load<1>(u, P.ovl_in + state_index(g, pr.s[1], pr.j[1]), ov, P.scale);
        }
        const uint32_t st = fc % kStagesK;
        FrameIO io;
        io.stage = stages + st * kStageFloats;
        io.scratch = scratch + (fc & 1u) * kStageFloats;
        io.nch = pr.nch;
        io.dst.emit = !(halo && f == 0);
        io.dst.interleaved = pr.interleaved;
        io.dst.scale = P.scale;
        io.dst.inv_scale = 1.0f / P.scale;
        io.dst.ostride = g.nc;
        io.dst.s16 = false;
        io.dq = nullptr;
        io.fi[0] = fi_ring[kRW * (fc % kFiRing)];
        io.fi[1] = fi_ring[kRW * (fc % kFiRing) + kRW / 2];
        io.ops = nullptr;
        io.dst.out0 = P.pcm + oa;
        io.dst.out1 = P.pcm + ob;
        if (leader) {
            const bool next_valid = f + kStagesK < nf;
            sync.put(0, next_valid ? 1u : 0u);
            if (next_valid) refill(st, fc + kStagesK);
        }
        mbar_wait(bars + 8 * st, (fc / kStagesK) & 1u);
        // the last frame of a block has landed: its ring slot may be overwritten
        if (leader && ((f + 1) % kFusedBlockFrames == 0 || f + 1 == nf))
            red_release_add(smem_u32(smem + L::kOffFlags) + 8u + 4u * ((f / kFusedBlockFrames) & 1), 1u);
        // two channels per stream (the launch condition): every frame has two long chains
        frame_all_long<2, true, AACFB_PK != 0, AACFB_ROT != 0, AACFB_ROT != 0>(u, sync, io, ts, P.tab, z, ov);
        oa += (size_t)1024 * g.nc; ob += (size_t)1024 * g.nc;
        if (++t == g.T) {
            ovl_store<0>(u, ov, P.ovl_out + state_index(g, pr.s[0], pr.j[0]), 1.0f / P.scale);
            if (pr.nch == 2) ovl_store<1>(u, ov, P.ovl_out + state_index(g, pr.s[1], pr.j[1]), 1.0f / P.scale);
            t = 0;
            if (f + 1 < nf) {
                pr = make_pair(g, ++pi);
                oa = (size_t)pr.s[0] * g.T * 1024 * g.nc + pr.j[0]; ob = (size_t)pr.s[1] * g.T * 1024 * g.nc + pr.j[1];
            }
        }
    }
}

// Stereo tools as a pre-pass: out = spectra with processMS / processIS applied (rows of pair-frames
// without a record are copied).  Used only when TNS has to run between the stereo tools and the
// IMDCT (TNS_FIXED_* modes); otherwise synth_kernel applies the ops on the staged rows.
// One warp per channel-pair frame.
__global__ void __launch_bounds__(128) stereo_kernel(const __grid_constant__ StereoParams P) {
    const size_t pf = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pf >= P.n_pairs_frames) return;
    const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(P.info) + 2 * pf);
    const bool has = ((raw.y >> 8) & 0xffu) != 0;
    const aacfb_stereo_ops *ops = P.stereo + pf;
    const float4 *L = reinterpret_cast<const float4 *>(P.spectra) + 2 * pf * 256, *R = L + 256;
    float4 *Lo = reinterpret_cast<float4 *>(P.out) + 2 * pf * 256, *Ro = Lo + 256;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int g = lane + 32 * i;
        const float4 l = L[g];
        float4 r = R[g], m = l;
        const int op = has ? ops->op[g] : AACFB_STEREO_NONE;
        if (op == AACFB_STEREO_MS) {
            m.x = f_add(l.x, r.x); m.y = f_add(l.y, r.y); m.z = f_add(l.z, r.z); m.w = f_add(l.w, r.w);
            r.x = f_sub(l.x, r.x); r.y = f_sub(l.y, r.y); r.z = f_sub(l.z, r.z); r.w = f_sub(l.w, r.w);
        } else if (op >= AACFB_STEREO_IS) {
            const float sc = ops->scale[(op - AACFB_STEREO_IS) & 127];
            r.x = f_mul(l.x, sc); r.y = f_mul(l.y, sc); r.z = f_mul(l.z, sc); r.w = f_mul(l.w, sc);
        }
        Lo[g] = m;
        Ro[g] = r;
    }
}

cudaError_t launch_stereo(const StereoParams &P, cudaStream_t stream) {
    const size_t threads = P.n_pairs_frames * 32;
    if (threads == 0) return cudaSuccess;
    stereo_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, stream>>>(P);
    return cudaGetLastError();
}

template <bool GENERIC, bool STEREO, bool IOV, int W, int ST>
static cudaError_t launch_one(const SynthParams &P, int num_sms, cudaStream_t stream) {
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(synth_kernel<GENERIC, STEREO, IOV, W, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Layout<W, ST, STEREO, IOV>::kTotal);
        if (e != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    const int grid = num_sms < (P.g.n_items + W - 1) / W ? num_sms : (P.g.n_items + W - 1) / W;
    synth_kernel<GENERIC, STEREO, IOV, W, ST><<<grid < 1 ? 1 : grid, W * 64, Layout<W, ST, STEREO, IOV>::kTotal, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_synth(const SynthParams &P, int num_sms, bool generic, cudaStream_t stream) {
    const bool iov = P.qframes != nullptr || P.pcm_s16 != 0;   // quantised input and / or int16 PCM
    if (P.stereo != nullptr) {   // the instantiations that also apply the stereo tools of the pair-frames
        if (iov) return generic ? launch_one<true, true, true, kWorkersGeneric, kStagesGeneric>(P, num_sms, stream)
                                : launch_one<false, true, true, kWorkers, kStages>(P, num_sms, stream);
        return generic ? launch_one<true, true, false, kWorkersGeneric, kStagesGeneric>(P, num_sms, stream)
                       : launch_one<false, true, false, kWorkers, kStages>(P, num_sms, stream);
    }
    if (iov) return generic ? launch_one<true, false, true, kWorkersGeneric, kStagesGeneric>(P, num_sms, stream)
                            : launch_one<false, false, true, kWorkers, kStages>(P, num_sms, stream);
    return generic ? launch_one<true, false, false, kWorkersGeneric, kStagesGeneric>(P, num_sms, stream)
                   : launch_one<false, false, false, kWorkers, kStages>(P, num_sms, stream);
}

// Inverse quantisation as a pre-pass: qframes -> float rows.  Only used when a TNS pass has to run
// between it and the IMDCT (TNS_FIXED_* modes); otherwise synth_kernel dequantises the staged records.
// One warp per channel-frame: lane l owns groups l + 32 j.
__global__ void __launch_bounds__(128) dequant_kernel(const __grid_constant__ DequantParams P) {
    const size_t cf = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (cf >= P.n_cf) return;
    const uint8_t *rec = P.qframes + cf * kQFrameBytes;
    const FrameBits fi = __ldg(reinterpret_cast<const uint32_t *>(P.info) + 2 * cf);
    float4 *out = reinterpret_cast<float4 *>(P.out) + cf * 256;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c4 = lane + 32 * j;
        const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(rec + 256 + 8 * c4));
        int16_t q4[4];
        q4[0] = (int16_t)(raw.x & 0xffffu); q4[1] = (int16_t)(raw.x >> 16);
        q4[2] = (int16_t)(raw.y & 0xffffu); q4[3] = (int16_t)(raw.y >> 16);
        float v[4];
        dequant4(P.dq, rec, fi, c4, q4, v);
        out[c4] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

cudaError_t launch_dequant(const DequantParams &P, cudaStream_t stream) {
    const size_t threads = P.n_cf * 32;
    if (threads == 0) return cudaSuccess;
    dequant_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_tns(const TnsParams &P, cudaStream_t stream) {
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(tns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTnsSmemBytes);
        if (e != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    const size_t rows_per_cta = (size_t)kTnsWarps * 32;
    const unsigned grid = (unsigned)((P.n_cf + rows_per_cta - 1) / rows_per_cta);
    if (grid == 0) return cudaSuccess;
    tns_kernel<<<grid, kTnsWarps * 32, kTnsSmemBytes, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_synth_tns(const SynthParams &P, int num_sms, cudaStream_t stream) {
    using L = FusedLayout<kWorkers, kStages>;
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(synth_tns_kernel<kWorkers, kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
        if (e != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    const int need = (P.g.n_items + kFusedClients - 1) / kFusedClients;
    const int grid = num_sms < need ? num_sms : need;
    synth_tns_kernel<kWorkers, kStages><<<grid < 1 ? 1 : grid, kWorkers * 64, L::kTotal, stream>>>(P);
    return cudaGetLastError();
}

size_t tns_ring_floats(int num_sms) { return (size_t)num_sms * 2 * kFusedRingRows * 1024; }

int synth_smem_bytes() { return Layout<kWorkers, kStages>::kTotal; }

}  // namespace aacfb
