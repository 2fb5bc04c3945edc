// aacfb_worker.cuh -- the per-frame schedule of one worker thread.
//
// A worker (64 threads) owns one staging buffer per frame: two spectrum rows
// of 1024 floats (one per chain), filled by TMA in the kernel.  Once a row
// has been consumed into registers the same 4 KiB serve as that chain's FFT
// exchange buffer; for EIGHT_SHORT frames the whole 8 KiB serve as the
// 2048-sample IMDCT buffer `buf` of filter_bank.js:43; at the end of the
// frame they serve as the transpose buffer of the PCM write-out.
//
// `Sync` provides the two synchronisation points the schedule needs:
//   sync.barrier()     all 64 threads of the worker
//   sync.stage_free()  barrier + "this frame's staging buffer may be refilled"
// The kernel implements them with a named barrier and a TMA issue, the CPU
// emulation (tests) with a std::barrier.
#pragma once
#include "aacfb_core.cuh"
#include "aacfb_geometry.h"

namespace aacfb {

struct FrameIO {
    float *stage;                 // 2048 floats: row of chain 0, row of chain 1
    aacfb_frame_info fi[2];
    int nch;                      // 1 or 2 live chains
    bool emit;                    // false for the halo frame of a chunk
    bool interleaved;             // chains are channels c, c+1 of one stereo stream
    float scale;                  // 2^-15 (decoder.js:210) or 1 for the inner seam
    float *out[2];                // sample 0 of this frame for each chain
    int ostride;
};

// ONLY_LONG / LONG_START / LONG_STOP for chains C0..C0+NCH-1.  Leaves the
// frame's PCM in `o`; the staging buffer is no longer read when it returns
// (callers still need a barrier before overwriting it).
template <int C0, int NCH, class Sync>
AACFB_HD void frame_long(int u, Sync &sync, const FrameIO &io, const SynthTables *ts, const SynthTables *tg,
                         Pts &z, Ovl &ov, Out &o) {
    const float *row[2] = {io.stage, io.stage + kRowFloats};
    float2 *buf[2] = {reinterpret_cast<float2 *>(io.stage), reinterpret_cast<float2 *>(io.stage + kRowFloats)};
    long_load<C0, NCH>(u, row, ts->cs2048, z);
    pass_a<C0, NCH>(z, ts->rootsA);
    sync.barrier();  // every thread has consumed the rows
    ex1_write<C0, NCH>(u, z, buf);
    sync.barrier();
    ex1_read<C0, NCH>(u, buf, z);
    pass_3stage<C0, NCH>(z, ts->twB + 7 * passb_blo(u));
    sync.barrier();
    ex2_write<C0, NCH>(u, z, buf);
    sync.barrier();
    ex2_read<C0, NCH>(u, buf, z);
    float2 twc[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) twc[j] = ts->twC[j][u];
    pass_3stage<C0, NCH>(z, twc);
    LongWin win[2];
#pragma unroll
    for (int c = C0; c < C0 + NCH; ++c) win[c] = long_windows(io.fi[c], ts->wz, tg);
    long_finish<C0, NCH>(u, z, ov, ts->cs2048, win, io.emit, io.scale, o);
}

// The 8 x 64-point FFTs of EIGHT_SHORT for chains C0..C0+NCH-1; leaves the
// un-twiddled bins in z.
template <int C0, int NCH, class Sync>
AACFB_HD void frame_short_fft(int u, Sync &sync, const FrameIO &io, const SynthTables *ts, Pts &z) {
    const float *row[2] = {io.stage, io.stage + kRowFloats};
    float2 *buf[2] = {reinterpret_cast<float2 *>(io.stage), reinterpret_cast<float2 *>(io.stage + kRowFloats)};
    short_load<C0, NCH>(u, row, ts->cs256, z);
    pass_a<C0, NCH>(z, ts->roots64A);
    sync.barrier();
    exs_write<C0, NCH>(u, z, buf);
    sync.barrier();
    exs_read<C0, NCH>(u, buf, z);
    pass_3stage<C0, NCH>(z, ts->twS + 7 * (u & 7));
}

// Window + overlap-add of chain C of an EIGHT_SHORT frame.  Needs the whole
// staging buffer: barrier first, since other threads may still be reading
// exchange data out of it.
template <int C, class Sync>
AACFB_HD void frame_short_ola(int u, Sync &sync, const FrameIO &io, const SynthTables *ts, const Pts &z, Ovl &ov,
                              Out &o) {
    sync.barrier();
    short_scatter<C>(u, z, ts->cs256, io.stage);
    sync.barrier();
    short_finish<C>(u, io.stage, ov, io.fi[C], ts->wshort, io.emit, io.scale, o);
}

AACFB_HD bool is_short(const aacfb_frame_info &fi) { return fi.window_sequence == AACFB_EIGHT_SHORT_SEQUENCE; }

// One frame of the worker's (up to) two chains.
template <class Sync>
AACFB_HD void worker_frame(int u, Sync &sync, const FrameIO &io, const SynthTables *ts, const SynthTables *tg, Pts &z,
                           Ovl &ov) {
    Out o;
    const bool s0 = is_short(io.fi[0]);
    const bool s1 = io.nch == 2 && is_short(io.fi[1]);
    if (io.nch == 1) {
        if (!s0) frame_long<0, 1>(u, sync, io, ts, tg, z, ov, o);
        else { frame_short_fft<0, 1>(u, sync, io, ts, z); frame_short_ola<0>(u, sync, io, ts, z, ov, o); }
    } else if (!s0 && !s1) {
        frame_long<0, 2>(u, sync, io, ts, tg, z, ov, o);
    } else if (s0 && s1) {
        frame_short_fft<0, 2>(u, sync, io, ts, z);
        frame_short_ola<0>(u, sync, io, ts, z, ov, o);
        frame_short_ola<1>(u, sync, io, ts, z, ov, o);
    } else if (s0) {  // chain 1 long first (touches only its own row), then chain 0 short
        frame_long<1, 1>(u, sync, io, ts, tg, z, ov, o);
        frame_short_fft<0, 1>(u, sync, io, ts, z);
        frame_short_ola<0>(u, sync, io, ts, z, ov, o);
    } else {
        frame_long<0, 1>(u, sync, io, ts, tg, z, ov, o);
        frame_short_fft<1, 1>(u, sync, io, ts, z);
        frame_short_ola<1>(u, sync, io, ts, z, ov, o);
    }
    if (io.emit) {
        sync.barrier();  // nobody reads exchange / IMDCT data out of the stage any more
        if (io.nch == 2) out_stage<0, 2>(u, o, io.stage, io.interleaved);
        else out_stage<0, 1>(u, o, io.stage, false);
        sync.barrier();
        out_copy(u, io.stage, io.out, io.ostride, 0, io.nch, io.interleaved && io.nch == 2);
    }
    sync.stage_free();
}

}  // namespace aacfb
