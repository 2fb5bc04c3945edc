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
#include <stdint.h>

#include "aacfb_kernels.h"
#include "aacfb_worker.cuh"

namespace aacfb {

namespace {

constexpr int kStageBytes = kStageFloats * 4;
constexpr int kTabBytes = (kSmemTableBytes + 127) & ~127;
constexpr int kOffStages = kTabBytes;
constexpr int kBufsPerWorker = kStages + 2;  // TMA ring + two alternating scratch buffers
constexpr int kOffBars = kOffStages + kWorkers * kBufsPerWorker * kStageBytes;
constexpr int kOffSlots = kOffBars + kWorkers * kStages * 8;
constexpr int kSmemTotal = kOffSlots + kWorkers * 4;
static_assert(kSmemTotal <= 227 * 1024, "shared memory budget");
static_assert(2 * kWorkers + 1 <= 16, "named barriers");

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

__global__ void __launch_bounds__(kCtaThreads, 1) synth_kernel(const __grid_constant__ SynthParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, w = tid >> 6;
    const int u = worker_thread_index((tid >> 5) & 1, tid & 31);
    const bool leader = u == 0;

    // constant tables -> shared memory (once per CTA)
    {
        const float4 *src = reinterpret_cast<const float4 *>(P.tab);
        float4 *dst = reinterpret_cast<float4 *>(smem);
        for (int i = tid; i < kSmemTableBytes / 16; i += kCtaThreads) dst[i] = src[i];
    }
    const SynthTables *ts = reinterpret_cast<const SynthTables *>(smem);
    float *stages = reinterpret_cast<float *>(smem + kOffStages) + (size_t)w * kBufsPerWorker * kStageFloats;
    float *scratch = stages + kStages * kStageFloats;
    const uint32_t bars = smem_u32(smem + kOffBars) + w * kStages * 8;
    volatile int *slot = reinterpret_cast<volatile int *>(smem + kOffSlots) + w;
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
        if (leader)
            *slot = first ? w * (int)gridDim.x + (int)blockIdx.x
                          : kWorkers * (int)gridDim.x + (int)atomicAdd(P.counter, 1u);
        first = false;
        sync.barrier();
        const int item = *slot;
        if (item >= P.n_items) break;
        const Item it = make_item(g, item);
        const int f_begin = it.t0 > 0 ? it.t0 - 1 : 0;
        const int nf = it.t1 - f_begin;

        if (leader) {  // prologue: fill the ring
            for (int i = 0; i < kStages && i < nf; ++i) {
                const uint32_t st = (fc + i) % kStages;
                sync.dst = smem_u32(stages + st * kStageFloats);
                sync.mbar = bars + 8 * st;
                sync.nrows = it.nch;
                sync.src[0] = row_ptr(P, cf_index(g, it.s[0], f_begin + i, it.j[0]));
                sync.src[1] = row_ptr(P, cf_index(g, it.s[1], f_begin + i, it.j[1]));
                sync.issue();
            }
        }
        if (it.t0 == 0) {
            ovl_load<0>(u, P.ovl_in + state_index(g, it.s[0], it.j[0]), ov);
            if (it.nch == 2) ovl_load<1>(u, P.ovl_in + state_index(g, it.s[1], it.j[1]), ov);
        }

        for (int f = 0; f < nf; ++f, ++fc) {
            const int t = f_begin + f;
            const uint32_t st = fc % kStages;
            FrameIO io;
            io.stage = stages + st * kStageFloats;
            io.scratch = scratch + (fc & 1u) * kStageFloats;
            io.nch = it.nch;
            io.dst.emit = t >= it.t0;
            io.dst.interleaved = it.interleaved;
            io.dst.scale = P.scale;
            io.dst.ostride = g.nc;
            io.fi[0] = info_lo(P, cf_index(g, it.s[0], t, it.j[0]));
            io.fi[1] = info_lo(P, cf_index(g, it.s[1], t, it.j[1]));
            io.dst.out0 = P.pcm + ((size_t)it.s[0] * g.T + t) * 1024 * g.nc + it.j[0];
            io.dst.out1 = P.pcm + ((size_t)it.s[1] * g.T + t) * 1024 * g.nc + it.j[1];
            sync.next_valid = f + kStages < nf;
            if (leader && sync.next_valid) {
                sync.dst = smem_u32(io.stage);
                sync.mbar = bars + 8 * st;
                sync.nrows = it.nch;
                sync.src[0] = row_ptr(P, cf_index(g, it.s[0], t + kStages, it.j[0]));
                sync.src[1] = row_ptr(P, cf_index(g, it.s[1], t + kStages, it.j[1]));
            }
            mbar_wait(bars + 8 * st, (fc / kStages) & 1u);
            worker_frame(u, sync, io, ts, P.tab, z, ov);
        }
        if (it.t1 == g.T) {
            ovl_store<0>(u, ov, P.ovl_out + state_index(g, it.s[0], it.j[0]));
            if (it.nch == 2) ovl_store<1>(u, ov, P.ovl_out + state_index(g, it.s[1], it.j[1]));
        }
    }
}

// One thread per channel-frame that carries TNS: copy the row to `scratch`
// (cooperatively, coalesced) and run its filters there.  The chain is serial
// by construction (see tns_run), parallelism is across channel-frames.
__global__ void __launch_bounds__(kTnsThreads) tns_kernel(const __grid_constant__ TnsParams P) {
    const size_t row0 = (size_t)blockIdx.x * kTnsThreads;
    for (int r = 0; r < kTnsThreads; ++r) {
        const size_t cf = row0 + r;
        if (cf >= P.n_cf) break;
        if (!P.info[cf].tns_present) continue;
        const float4 *x4 = reinterpret_cast<const float4 *>(P.spectra + cf * 1024);
        float4 *y4 = reinterpret_cast<float4 *>(P.scratch + cf * 1024);
        for (int i = threadIdx.x; i < 256; i += kTnsThreads) y4[i] = x4[i];
    }
    __syncthreads();
    const size_t cf = row0 + threadIdx.x;
    if (cf >= P.n_cf) return;
    const aacfb_frame_info fi = P.info[cf];
    if (!fi.tns_present) return;
    const uint32_t o0 = P.offsets[cf], o1 = P.offsets[cf + 1];
    if (o1 <= o0 || o1 > P.blob_bytes) return;
    aacfb_frame_info f2 = fi;
    f2.window_sequence &= 3;
    tns_apply(f2, P.blob + o0, o1 - o0, P.sample_index, P.ar != 0, *P.bands, P.spectra + cf * 1024,
              P.scratch + cf * 1024);
}

cudaError_t launch_synth(const SynthParams &P, int grid, cudaStream_t stream) {
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
        if (e != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    synth_kernel<<<grid, kCtaThreads, kSmemTotal, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_tns(const TnsParams &P, cudaStream_t stream) {
    const unsigned grid = (unsigned)((P.n_cf + kTnsThreads - 1) / kTnsThreads);
    if (grid == 0) return cudaSuccess;
    tns_kernel<<<grid, kTnsThreads, 0, stream>>>(P);
    return cudaGetLastError();
}

int synth_smem_bytes() { return kSmemTotal; }

}  // namespace aacfb
