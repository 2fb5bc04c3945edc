// aacfb_worker.cuh -- the per-frame schedule of one worker thread.
//
// A worker (64 threads) owns one staging buffer per frame: two spectrum rows
// of 1024 floats (one per chain), filled by TMA in the kernel.  Once a row
// has been consumed into registers the same 4 KiB serve as that chain's FFT
// second exchange buffer (the first exchange goes through a per-worker scratch
// buffer); for EIGHT_SHORT frames stage and scratch serve as the 2048-sample
// IMDCT buffers `buf` of filter_bank.js:43; the PCM write-out is transposed
// through whichever of the two is free.
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
    float *stage;                 // 2048 floats filled by TMA: row of chain 0, row of chain 1
    float *scratch;               // 2048 floats, alternates between two buffers from frame to frame
    FrameBits fi[2];              // packed aacfb_frame_info of each chain
    int nch;                      // 1 or 2 live chains
    bool emit;                    // false for the halo frame of a chunk
    bool interleaved;             // chains are channels 0, 1 of one stereo stream
    float scale;                  // 2^-15 (decoder.js:210) or 1 for the inner seam
    float *out0, *out1;           // sample 0 of this frame for each chain
    int ostride;
};

AACFB_HD bool is_short(FrameBits fi) { return fb_seq(fi) == AACFB_EIGHT_SHORT_SEQUENCE; }

// Buffer plan of a frame (S = stage, X = scratch).  Every arrow that crosses
// threads is separated by exactly one worker barrier; because the scratch
// buffer alternates between frames, nothing of frame f+1 can collide with the
// write-out of frame f.
//
//   long:   rows S -> regs -> X (exchange 1) -> regs -> S (exchange 2) -> regs -> PCM X -> global
//   short:  rows S -> regs -> X (exchange)   -> regs -> IMDCT buffers S / X -> regs -> PCM -> global

// 512-point inverse FFT of chains C0..C0+NCH-1 (fft.js:105-192 on the
// pre-twiddled rows, mdct.js:73-79): two barriers.
template <int C0, int NCH, class Sync>
AACFB_HD void long_fft(int u, Sync &sync, const FrameIO &io, const SynthTables *ts, Pts &z) {
    const float *row[2] = {io.stage, io.stage + kRowFloats};
    float2 *bufx[2] = {reinterpret_cast<float2 *>(io.scratch), reinterpret_cast<float2 *>(io.scratch + kRowFloats)};
    float2 *bufs[2] = {reinterpret_cast<float2 *>(io.stage), reinterpret_cast<float2 *>(io.stage + kRowFloats)};
    long_load<C0, NCH>(u, row, ts->cs2048, z);
    pass_a<C0, NCH>(z, ts->rootsA);
    ex1_write<C0, NCH>(u, z, bufx);
    sync.barrier();  // exchange 1 complete; every thread has consumed its part of the rows
    ex1_read<C0, NCH>(u, bufx, z);
    pass_3stage<C0, NCH>(z, ts->twB + 7 * passb_blo(u));
    ex2_write<C0, NCH>(u, z, bufs);
    sync.barrier();  // exchange 2 complete; exchange-1 data is dead
    ex2_read<C0, NCH>(u, bufs, z);
    float2 twc[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) twc[j] = ts->twC[j][u];
    pass_3stage<C0, NCH>(z, twc);
}

// The 8 x 64-point FFTs of EIGHT_SHORT for chains C0..C0+NCH-1: one barrier.
template <int C0, int NCH, class Sync>
AACFB_HD void short_fft(int u, Sync &sync, const FrameIO &io, const SynthTables *ts, Pts &z) {
    const float *row[2] = {io.stage, io.stage + kRowFloats};
    float2 *bufx[2] = {reinterpret_cast<float2 *>(io.scratch), reinterpret_cast<float2 *>(io.scratch + kRowFloats)};
    short_load<C0, NCH>(u, row, ts->cs256, z);
    pass_a<C0, NCH>(z, ts->roots64A);
    exs_write<C0, NCH>(u, z, bufx);
    sync.barrier();
    exs_read<C0, NCH>(u, bufx, z);
    pass_3stage<C0, NCH>(z, ts->twS + 7 * (u & 7));
}

// Write-out of a finished frame from buffer `src`, after which the stage may
// be refilled.  `src_is_stage`: the PCM sits in the stage itself, so the copy
// has to finish before the refill.
template <class Sync>
AACFB_HD void frame_tail(int u, Sync &sync, const FrameIO &io, const float *src, bool src_is_stage) {
    if (!io.emit) { sync.stage_free(); return; }
    if (src_is_stage) sync.barrier(); else sync.stage_free();
    if (io.nch == 2 && io.interleaved) out_copy_interleaved(u, src, io.out0);
    else {
        out_copy_planar(u, src, io.out0, io.ostride);
        if (io.nch == 2) out_copy_planar(u, src + 1024, io.out1, io.ostride);
    }
    if (src_is_stage) sync.stage_free();
}

// A frame whose chains are all long transforms: 3 barriers.
template <int NCH, class Sync>
AACFB_HD void frame_all_long(int u, Sync &sync, const FrameIO &io, const SynthTables *ts, const SynthTables *tg,
                             Pts &z, Ovl &ov) {
    long_fft<0, NCH>(u, sync, io, ts, z);
    Out none;
    const bool uniform = fb_seq(io.fi[0]) == AACFB_ONLY_LONG_SEQUENCE &&
                         (NCH == 1 || ((io.fi[0] ^ io.fi[1]) & 0x00ffffffu) == 0);
    // exchange-1 data in the scratch buffer is dead since the second barrier: the PCM goes there
    if (uniform) long_finish<0, NCH, true, true>(u, z, ov, ts, tg, io.fi, io.emit, io.scale, io.scratch, io.interleaved, none);
    else long_finish<0, NCH, false, true>(u, z, ov, ts, tg, io.fi, io.emit, io.scale, io.scratch, io.interleaved, none);
    frame_tail(u, sync, io, io.scratch, false);
}

// A frame with at least one EIGHT_SHORT chain: results pass through registers.
template <class Sync>
AACFB_HD void frame_with_short(int u, Sync &sync, const FrameIO &io, const SynthTables *ts, const SynthTables *tg,
                               Pts &z, Ovl &ov) {
    Out o;
    const bool s0 = is_short(io.fi[0]);
    const bool s1 = io.nch == 2 && is_short(io.fi[1]);
    if (io.nch == 2 && s0 && s1) {
        short_fft<0, 2>(u, sync, io, ts, z);
        short_scatter<0>(u, z, ts->cs256, io.stage);        // rows are dead since the exchange barrier
        sync.barrier();
        short_finish<0>(u, io.stage, ov, io.fi[0], ts->wshort, io.emit, io.scale, o);
        short_scatter<1>(u, z, ts->cs256, io.scratch);      // exchange data dead since the last barrier
        sync.barrier();
        short_finish<1>(u, io.scratch, ov, io.fi[1], ts->wshort, io.emit, io.scale, o);
        if (io.emit) out_stage<0, 2>(u, o, io.stage, io.interleaved);  // chain 0's buffer: dead since the last barrier
        frame_tail(u, sync, io, io.stage, true);
        return;
    }
    if (io.nch == 1) {
        short_fft<0, 1>(u, sync, io, ts, z);
        short_scatter<0>(u, z, ts->cs256, io.stage);
        sync.barrier();
        short_finish<0>(u, io.stage, ov, io.fi[0], ts->wshort, io.emit, io.scale, o);
        if (io.emit) out_stage<0, 1>(u, o, io.scratch, false);
    } else if (s0) {  // chain 1 long first (it only touches its own halves), then chain 0 short
        long_fft<1, 1>(u, sync, io, ts, z);
        long_finish<1, 1, false, false>(u, z, ov, ts, tg, io.fi, io.emit, io.scale, io.stage, false, o);
        short_fft<0, 1>(u, sync, io, ts, z);
        short_scatter<0>(u, z, ts->cs256, io.stage);
        sync.barrier();
        short_finish<0>(u, io.stage, ov, io.fi[0], ts->wshort, io.emit, io.scale, o);
        if (io.emit) out_stage<0, 2>(u, o, io.scratch, io.interleaved);
    } else {
        long_fft<0, 1>(u, sync, io, ts, z);
        long_finish<0, 1, false, false>(u, z, ov, ts, tg, io.fi, io.emit, io.scale, io.stage, false, o);
        short_fft<1, 1>(u, sync, io, ts, z);
        short_scatter<1>(u, z, ts->cs256, io.stage);
        sync.barrier();
        short_finish<1>(u, io.stage, ov, io.fi[1], ts->wshort, io.emit, io.scale, o);
        if (io.emit) out_stage<0, 2>(u, o, io.scratch, io.interleaved);
    }
    frame_tail(u, sync, io, io.scratch, false);
}

// One frame of the worker's (up to) two chains.
template <class Sync>
AACFB_HD void worker_frame(int u, Sync &sync, const FrameIO &io, const SynthTables *ts, const SynthTables *tg, Pts &z,
                           Ovl &ov) {
    const bool any_short = is_short(io.fi[0]) || (io.nch == 2 && is_short(io.fi[1]));
    if (any_short) frame_with_short(u, sync, io, ts, tg, z, ov);
    else if (io.nch == 2) frame_all_long<2>(u, sync, io, ts, tg, z, ov);
    else frame_all_long<1>(u, sync, io, ts, tg, z, ov);
}

}  // namespace aacfb
