// aacfb_worker.cuh -- the per-frame schedule of one worker thread.
//
// A worker (64 threads) owns one staging buffer per frame: two spectrum rows
// of 1024 floats (one per chain), filled by TMA in the kernel.  Once a row
// has been consumed into registers the same 4 KiB serve as that chain's FFT
// second exchange buffer (the first exchange goes through a per-worker scratch
// buffer); for EIGHT_SHORT frames stage and scratch serve as the 2048-sample
// window-product arrays of short_products (the windowed halves of `buf`, filter_bank.js:43).  Finished PCM never touches
// shared memory: it is paired up by a warp shuffle and stored directly.
//
// `Sync` provides what the schedule needs from the outside:
//   sync.barrier()      all 64 threads of the worker
//   sync.stage_free()   "this thread no longer reads this frame's stage": once all 64
//                       threads have said so the stage is refilled (TMA in the kernel)
//   sync.partner(u, v)  the value v of thread 63-u (a warp shuffle in the kernel)
// The kernel implements them with named barriers, a TMA issue and SHFL, the
// CPU emulation (tests) with a std::barrier and a mailbox.
#pragma once
#include "aacfb_core.cuh"
#include "aacfb_geometry.h"

// Tuning switches (tools/variant.sh builds variants; the defaults are the measured best).
// AACFB_ROT must be the same for every instantiation: a long frame has to come out with the same
// bits whichever instantiation (long-only / generic) happens to process its slice.
#ifndef AACFB_PK
#define AACFB_PK 1      // packed two-chain arithmetic (FFMA2) on frames of two long or two short chains
#endif
#ifndef AACFB_GENERIC_UNIFORM
#define AACFB_GENERIC_UNIFORM 0   // generic instantiations: also compile the uniform-ONLY_LONG finish
#endif
#ifndef AACFB_TW_SYM
#define AACFB_TW_SYM 0   // EXPERIMENT (off): derive 3 of the 7 FFT twiddles of a 3-stage pass as i * another one
                         // (roots[k + L/4] = i roots[k] ideally; the reference's recurrence-built table deviates
                         // by up to 1.8e-6 from that, which is why this is not free for parity -- see load_tw)
#endif
#ifndef AACFB_ROT
#define AACFB_ROT 1     // MDCT twiddles of the pre- and post-twiddle derived by rotation (cs_at)
#endif

namespace aacfb {

struct FrameIO {
    float *stage;                 // 2048 floats filled by TMA: row of chain 0, row of chain 1
    float *scratch;               // 2048 floats, alternates between two buffers from frame to frame
    FrameBits fi[2];              // packed aacfb_frame_info of each chain
    int nch;                      // 1 or 2 live chains
    const aacfb_stereo_ops *ops;  // stereo tools of this pair-frame (staged next to the rows) or nullptr
    const DqCtx *dq;              // != nullptr: the stage holds aacfb_qframe records (dequant_stage turns them into rows)
    OutDst dst;                   // where the frame's PCM goes
};

AACFB_HD bool is_short(FrameBits fi) { return fb_seq(fi) == AACFB_EIGHT_SHORT_SEQUENCE; }

// Buffer plan of a frame (S = stage, X = scratch).  Every arrow that crosses
// threads is separated by exactly one worker barrier; because the scratch
// buffer alternates between frames, nothing of frame f+1 can collide with the
// write-out of frame f.
//
//   long:   rows S -> regs -> X (exchange 1) -> regs -> S (exchange 2) -> regs -> global
//   short:  rows S -> regs -> X (exchange)   -> regs -> IMDCT buffers S / X -> regs -> global

// The seven twiddles of a 3-stage pass, tw[j] at p[j * stride]: tw[2] = roots[a + L/4], tw[5] = roots[b + L/4],
// tw[6] = roots[c + L/4] where tw[1] = roots[a], tw[3] = roots[b], tw[4] = roots[c] (tables.h).
AACFB_HD void load_tw(const float2 *p, int stride, float2 *tw) {
    tw[0] = p[0]; tw[1] = p[stride]; tw[3] = p[3 * stride]; tw[4] = p[4 * stride];
    tw[2].x = -tw[1].y; tw[2].y = tw[1].x;
    tw[5].x = -tw[3].y; tw[5].y = tw[3].x;
    tw[6].x = -tw[4].y; tw[6].y = tw[4].x;
}

// 512-point inverse FFT of chains C0..C0+NCH-1 (fft.js:105-192 on the
// pre-twiddled rows, mdct.js:73-79): two barriers.
template <int C0, int NCH, bool PK, bool ROT, bool STAG = false, class Sync>
AACFB_HD void long_fft(int u, Sync &sync, const FrameIO &io, const SynthTables *ts, Pts &z) {
    const float *row[2] = {io.stage, io.stage + kRowFloats};
    float2 *bufx[2] = {reinterpret_cast<float2 *>(io.scratch), reinterpret_cast<float2 *>(io.scratch + kRowFloats)};
    float2 *bufs[2] = {reinterpret_cast<float2 *>(io.stage), reinterpret_cast<float2 *>(io.stage + kRowFloats)};
    long_load<C0, NCH, PK, ROT, STAG>(u, row, ts->cs2048, z);
    pass_a<C0, NCH, PK>(z, ts->rootsA);
    ex1_write<C0, NCH, PK>(u, z, bufx);
    sync.barrier();  // exchange 1 complete; every thread has consumed its part of the rows
    ex1_read<C0, NCH, PK>(u, bufx, z);
#if AACFB_TW_SYM
    float2 twb[7];
    load_tw(ts->twB + 7 * passb_blo(u), 1, twb);
    pass_3stage<C0, NCH, PK>(z, twb);
#else
    pass_3stage<C0, NCH, PK>(z, ts->twB + 7 * passb_blo(u));
#endif
    ex2_write<C0, NCH, PK>(u, z, bufs);
    sync.barrier();  // exchange 2 complete; exchange-1 data is dead
    ex2_read<C0, NCH, PK>(u, bufs, z);
    float2 twc[7];
#if AACFB_TW_SYM
    load_tw(&ts->twC[0][u], 64, twc);
#else
#pragma unroll
    for (int j = 0; j < 7; ++j) twc[j] = ts->twC[j][u];
#endif
    pass_3stage<C0, NCH, PK>(z, twc);
}

// The 8 x 64-point FFTs of EIGHT_SHORT for chains C0..C0+NCH-1: one barrier.
template <int C0, int NCH, bool PK, class Sync>
AACFB_HD void short_fft(int u, Sync &sync, const FrameIO &io, const SynthTables *ts, Pts &z) {
    const float *row[2] = {io.stage, io.stage + kRowFloats};
    float2 *bufx[2] = {reinterpret_cast<float2 *>(io.scratch), reinterpret_cast<float2 *>(io.scratch + kRowFloats)};
    short_load<C0, NCH, PK>(u, sync, row, ts->cs256, z);
    pass_a<C0, NCH, PK>(z, ts->roots64A);
    exs_write<C0, NCH, PK>(u, z, bufx);
    sync.barrier();
    exs_read<C0, NCH, PK>(u, bufx, z);
#if AACFB_TW_SYM
    float2 tws[7];
    load_tw(ts->twS + 7 * (u & 7), 1, tws);
    pass_3stage<C0, NCH, PK>(z, tws);
#else
    pass_3stage<C0, NCH, PK>(z, ts->twS + 7 * (u & 7));
#endif
}

// A frame whose chains are all long transforms: two blocking barriers; the PCM
// leaves straight from the finishing arithmetic.
// UNIFORM_PATH: also instantiate the specialisation for frames whose chains are all ONLY_LONG with
// equal shapes.  The generic pass leaves it out: its code footprint (long + short paths running
// side by side on one SM) is what the instruction cache has to hold.
template <int NCH, bool UNIFORM_PATH, bool PK, bool ROTL, bool ROTF, bool STAG = false, class Sync>
AACFB_HD void frame_all_long(int u, Sync &sync, const FrameIO &io, const SynthTables *ts, const SynthTables *tg,
                             Pts &z, Ovl &ov) {
    long_fft<0, NCH, PK, ROTL, STAG>(u, sync, io, ts, z);
    sync.stage_free();  // exchange 2 has been read back: the stage may be refilled
    Out none;
    const bool uniform = UNIFORM_PATH && fb_seq(io.fi[0]) == AACFB_ONLY_LONG_SEQUENCE &&
                         (NCH == 1 || ((io.fi[0] ^ io.fi[1]) & 0x00ffffffu) == 0);
    if (UNIFORM_PATH && uniform) long_finish<0, NCH, true, true, PK, ROTF>(u, sync, z, ov, ts, tg, io.fi, io.dst, none);
    else long_finish<0, NCH, false, true, PK, ROTF>(u, sync, z, ov, ts, tg, io.fi, io.dst, none);
}

// A frame with at least one EIGHT_SHORT chain: results pass through registers.
template <bool PK, class Sync>
AACFB_HD void frame_with_short(int u, Sync &sync, const FrameIO &io, const SynthTables *ts, const SynthTables *tg,
                               Pts &z, Ovl &ov) {
    Out o;
    const OutDst &d = io.dst;
    const bool s0 = is_short(io.fi[0]);
    const bool s1 = io.nch == 2 && is_short(io.fi[1]);
    if (io.nch == 2 && s0 && s1) {
        short_fft<0, 2, PK>(u, sync, io, ts, z);
        short_products<0>(u, z, ts->cs256, ts->wsp, io.fi[0], io.stage);    // rows are dead since the exchange barrier
        sync.barrier();
        short_finish<0>(u, io.stage, ov, d.emit, o);
        short_products<1>(u, z, ts->cs256, ts->wsp, io.fi[1], io.scratch);  // exchange data dead since the last barrier
        sync.barrier();
        sync.stage_free();                                   // chain 0's IMDCT buffer has been consumed
        short_finish<1>(u, io.scratch, ov, d.emit, o);
        if (d.emit) out_store<0, 2>(u, sync, o, d);
        return;
    }
    if (io.nch == 1) {
        short_fft<0, 1, false>(u, sync, io, ts, z);
        short_products<0>(u, z, ts->cs256, ts->wsp, io.fi[0], io.stage);
        sync.barrier();
        short_finish<0>(u, io.stage, ov, d.emit, o);
        sync.stage_free();
        if (d.emit) out_store<0, 1>(u, sync, o, d);
        return;
    }
    if (s0) {  // chain 1 long first (it only touches its own halves), then chain 0 short
        long_fft<1, 1, false, AACFB_ROT != 0>(u, sync, io, ts, z);
        long_finish<1, 1, false, false, false, AACFB_ROT != 0>(u, sync, z, ov, ts, tg, io.fi, d, o);
        short_fft<0, 1, false>(u, sync, io, ts, z);
        short_products<0>(u, z, ts->cs256, ts->wsp, io.fi[0], io.stage);
        sync.barrier();
        short_finish<0>(u, io.stage, ov, d.emit, o);
    } else {
        long_fft<0, 1, false, AACFB_ROT != 0>(u, sync, io, ts, z);
        long_finish<0, 1, false, false, false, AACFB_ROT != 0>(u, sync, z, ov, ts, tg, io.fi, d, o);
        short_fft<1, 1, false>(u, sync, io, ts, z);
        short_products<1>(u, z, ts->cs256, ts->wsp, io.fi[1], io.stage);
        sync.barrier();
        short_finish<1>(u, io.stage, ov, d.emit, o);
    }
    sync.stage_free();
    if (d.emit) out_store<0, 2>(u, sync, o, d);
}

// One frame of the worker's (up to) two chains.  GENERIC = false is the
// instantiation for work items without any EIGHT_SHORT frame: it contains no
// short-window code at all, so its register allocation is that of the long
// path alone (the kernel is compiled twice, see synth_kernel).
// IOV: the instantiation that also takes quantised input (AACFB_IN_Q16) and / or emits int16 PCM; the
// float-only instantiations (IOV = false) contain none of that code.
template <bool GENERIC, bool STEREO, bool IOV, class Sync>
AACFB_HD void worker_frame(int u, Sync &sync, const FrameIO &io, const SynthTables *ts, const SynthTables *tg, Pts &z,
                           Ovl &ov) {
    // inverse quantisation first: ics.js:203-266 runs inside the bit parse, before the stereo tools
    if (IOV && io.dq) dequant_stage(u, sync, io.stage, io.nch, io.fi, *io.dq);
    // Packed two-chain arithmetic (FFMA2) halves the FP instruction count; on its own it is neutral
    // for the long-only instantiation (bounded by the shared-memory pipe, not by issue slots), but
    // it is what pays for deriving the MDCT twiddles in registers instead of loading them (cs_at):
    // together 0.2265 -> 0.198 ms on config 2 (profiles/README.md).
    constexpr bool PK = AACFB_PK != 0;
    constexpr bool ROTL = AACFB_ROT != 0, ROTF = AACFB_ROT != 0;
    // M/S and intensity stereo act on the spectra before anything else (decoder.js:294-301): in place
    // on the staged rows, one op per group of 4 coefficients.  (Applying them per element while the
    // rows are read into registers was measured 6 % slower: 16 byte loads + selects per thread
    // against 4 vector read-modify-writes.)
    if (STEREO && io.ops) {
        stereo_apply(u, io.stage, io.ops);
        sync.barrier();
    }
    if (GENERIC) {
        const bool any_short = is_short(io.fi[0]) || (io.nch == 2 && is_short(io.fi[1]));
        if (any_short) { frame_with_short<PK>(u, sync, io, ts, tg, z, ov); return; }
    }
    constexpr bool UNI = !GENERIC || AACFB_GENERIC_UNIFORM != 0;
    // staggered row reads (long_load): only where registers are to spare -- the plain long-only instantiation
    // (config 2 +0.7 %; the stereo one loses 0.5 %, the generic ones 5 %: profiles/r04_ab_experiments.txt)
    constexpr bool STAG = !GENERIC && !STEREO && !IOV && AACFB_LONG_STAGGER != 0;
    if (io.nch == 2) frame_all_long<2, UNI, PK, ROTL, ROTF, STAG>(u, sync, io, ts, tg, z, ov);
    else frame_all_long<1, UNI, false, ROTL, ROTF, STAG>(u, sync, io, ts, tg, z, ov);
}

}  // namespace aacfb
