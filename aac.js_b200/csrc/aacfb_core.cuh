// aacfb_core.cuh -- per-thread compute phases of the synthesis kernel.
//
// A *worker* is 64 threads that carry two (stream, channel) chains through
// time.  Each thread keeps 8 complex points per chain in registers; a
// 512-point inverse FFT is three register passes (radix-2 DIT stages 1-3,
// 4-6, 7-9) with two swizzled shared-memory exchanges in between.  The
// butterfly network, twiddle values and operation order are those of the
// reference's FFT.process (src/fft.js:105-192) so rounding tracks it; only
// the *placement* of butterflies on threads is new.
//
// Every phase is a pure function of (thread index, registers, staging
// buffer, tables), written __host__ __device__ so that tests can run the
// very same code on the CPU (csrc/aacfb_emul.cu) where no GPU exists.
//
// All arithmetic goes through f_mul/f_add/f_fma so that neither nvcc nor gcc
// contracts or reassociates anything: host emulation and device agree bit
// for bit.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/aacfb.h"
#include "aacfb_tables.h"

#include "aacfb_geometry.h"

namespace aacfb {

#if defined(__CUDA_ARCH__)
AACFB_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
AACFB_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
AACFB_HD float f_sub(float a, float b) { return __fsub_rn(a, b); }
AACFB_HD float f_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#else
AACFB_HD float f_mul(float a, float b) { volatile float r = a * b; return r; }
AACFB_HD float f_add(float a, float b) { volatile float r = a + b; return r; }
AACFB_HD float f_sub(float a, float b) { volatile float r = a - b; return r; }
AACFB_HD float f_fma(float a, float b, float c) { return fmaf(a, b, c); }
#endif

#if defined(__CUDA_ARCH__)
AACFB_HD uint32_t f_bits(float v) { return __float_as_uint(v); }
AACFB_HD float f_from_bits(uint32_t b) { return __uint_as_float(b); }
#else
AACFB_HD uint32_t f_bits(float v) { uint32_t b; memcpy(&b, &v, 4); return b; }
AACFB_HD float f_from_bits(uint32_t b) { float v; memcpy(&v, &b, 4); return v; }
#endif

// ---- packed pairs: the same quantity of chain 0 (.x) and chain 1 (.y) -----------------------
// A worker runs identical arithmetic on its two chains with shared twiddles and windows, which
// is the shape of sm_100's packed FP32 instructions (PTX fma/mul/add/sub.rn.f32x2 -> SASS FFMA2 /
// FMUL2 / FADD2: one issue slot for both chains, per-lane IEEE rounding identical to the scalar
// instruction, operand negation and scalar broadcast are free operand modifiers).  The host
// build (CPU emulation in tests) performs the two scalar operations instead: same bits.
struct F2 {
    float x, y;
};
#if defined(__CUDA_ARCH__)
AACFB_HD unsigned long long f2_bits(F2 v) {
    unsigned long long d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(v.x), "f"(v.y));
    return d;
}
AACFB_HD F2 f2_from(unsigned long long d) {
    F2 v;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(d));
    return v;
}
AACFB_HD F2 f_fma(F2 a, F2 b, F2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
    return f2_from(d);
}
AACFB_HD F2 f_mul(F2 a, F2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return f2_from(d);
}
AACFB_HD F2 f_add(F2 a, F2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return f2_from(d);
}
AACFB_HD F2 f_sub(F2 a, F2 b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return f2_from(d);
}
#else
AACFB_HD F2 f_fma(F2 a, F2 b, F2 c) { return F2{f_fma(a.x, b.x, c.x), f_fma(a.y, b.y, c.y)}; }
AACFB_HD F2 f_mul(F2 a, F2 b) { return F2{f_mul(a.x, b.x), f_mul(a.y, b.y)}; }
AACFB_HD F2 f_add(F2 a, F2 b) { return F2{f_add(a.x, b.x), f_add(a.y, b.y)}; }
AACFB_HD F2 f_sub(F2 a, F2 b) { return F2{f_sub(a.x, b.x), f_sub(a.y, b.y)}; }
#endif
AACFB_HD F2 f_neg(F2 a) { return F2{-a.x, -a.y}; }
AACFB_HD float f_neg(float a) { return -a; }
// a scalar (twiddle, window value) applied to both chains
AACFB_HD F2 f_fma(F2 a, float s, F2 c) { return f_fma(a, F2{s, s}, c); }
AACFB_HD F2 f_fma(float s, F2 a, F2 c) { return f_fma(F2{s, s}, a, c); }
AACFB_HD F2 f_mul(F2 a, float s) { return f_mul(a, F2{s, s}); }

#ifndef AACFB_SHORT_PAIRLOAD
#define AACFB_SHORT_PAIRLOAD 1   // tuning switch: EIGHT_SHORT rows read as 8-byte pairs + one shuffle
#endif
#ifndef AACFB_LONG_STAGGER
#define AACFB_LONG_STAGGER 1     // tuning switch: half of the lanes read x[2n] / x[1023 - 2n] in the opposite order (long_load, long-only instantiation)
#endif
#ifndef AACFB_SHORT_STAGGER
#define AACFB_SHORT_STAGGER 0    // tuning switch: odd windows read their row pairs in the opposite order (short_load); see profiles/r04_ab_experiments.txt
#endif

constexpr int kWorkerThreads = 64;
constexpr int kRowFloats = 1024;        // one channel-frame of spectrum
constexpr int kStageFloats = 2 * 1024;  // two chains per stage

AACFB_HD constexpr int brev3(int q) { return ((q & 1) << 2) | (q & 2) | ((q >> 2) & 1); }
AACFB_HD int brev6(int u) {
    return ((u & 1) << 5) | ((u & 2) << 3) | ((u & 4) << 1) | ((u & 8) >> 1) | ((u & 16) >> 3) | ((u & 32) >> 5);
}

// Logical thread index u of (warp-in-worker w, lane l): chosen so that the
// mirror partner 63-u of every thread is lane l^31 of the same warp, while
// each half-warp still covers 16 consecutive u (all bank-conflict analyses
// below are per half-warp).  warp 0 = {0..15, 48..63}, warp 1 = {16..47}.
AACFB_HD int worker_thread_index(int w, int l) { return w == 0 ? (l < 16 ? l : 32 + l) : 16 + l; }

// aacfb_frame_info packed into one 32-bit word (one 8-byte load per channel-frame):
// window_sequence | shape_prev << 8 | shape_cur << 16 | max_sfb << 24.
typedef uint32_t FrameBits;
AACFB_HD int fb_seq(FrameBits b) { return b & 3; }
AACFB_HD int fb_shape_prev(FrameBits b) { return (b >> 8) & 1; }
AACFB_HD int fb_shape_cur(FrameBits b) { return (b >> 16) & 1; }
AACFB_HD FrameBits fb_pack(const aacfb_frame_info &fi) {
    return (uint32_t)fi.window_sequence | ((uint32_t)fi.shape_prev << 8) | ((uint32_t)fi.shape_cur << 16) |
           ((uint32_t)fi.max_sfb << 24);
}

// Registers of one thread: 8 complex points for each of the worker's 2 chains.
struct Pts {
    float r[2][8], i[2][8];
};
// Overlap carried between frames: positions m(q) and 1023-m(q), q = 0..7.
struct Ovl {
    float a[2][8], b[2][8];
};
// Phases below are templated on <C0, NCH>: they act on chains C0 .. C0+NCH-1.

// ---------------------------------------------------------------- butterflies
// fft.js:180-188: t = x[hi]*w (kept in double there); x[hi] = x[lo]-t; x[lo] += t.
// lo = a + t takes two FMAs per component; hi = a - t is formed as 2a - lo (one FMA, 2a is
// exact): 6 instead of 8 FMAs per butterfly, at the price of lo's rounding error (<= 1 ulp of
// lo) reappearing in hi -- well inside the parity budget (tests measure ~3e-7 of 1e-5).
// V = float (one chain) or F2 (both chains of the worker at once).
template <class V>
AACFB_HD void bfly(V &ar, V &ai, V &br, V &bi, float wr, float wi) {
    const V lr = f_fma(br, wr, f_fma(f_neg(bi), wi, ar));
    const V li = f_fma(br, wi, f_fma(bi, wr, ai));
    const V hr = f_fma(2.0f, ar, f_neg(lr));
    const V hi = f_fma(2.0f, ai, f_neg(li));
    ar = lr; ai = li; br = hr; bi = hi;
}
template <class V>
AACFB_HD void bfly1(V &ar, V &ai, V &br, V &bi) {  // w = (1, 0): exact in the reference too
    const V lr = f_add(ar, br), li = f_add(ai, bi);
    const V hr = f_sub(ar, br), hi = f_sub(ai, bi);
    ar = lr; ai = li; br = hr; bi = hi;
}
// fft.js:140-170, inverse branch: the fused first two stages on 4 points.
template <class V>
AACFB_HD void base4(V *r, V *i) {
    const V a0 = f_add(r[0], r[1]), a1 = f_add(i[0], i[1]);
    const V b0 = f_add(r[2], r[3]), b1 = f_add(i[2], i[3]);
    const V c0 = f_sub(r[0], r[1]), c1 = f_sub(i[0], i[1]);
    const V d0 = f_sub(r[2], r[3]), d1 = f_sub(i[2], i[3]);
    r[0] = f_add(a0, b0); i[0] = f_add(a1, b1);
    r[2] = f_sub(a0, b0); i[2] = f_sub(a1, b1);
    r[1] = f_sub(c0, d1); i[1] = f_add(c1, d0);
    r[3] = f_add(c0, d1); i[3] = f_sub(c1, d0);
}

// Stages 1-3 on 8 consecutive array positions (regs indexed by position&7).
// wA = roots[k * L/8], k = 0..3 (fft.js:173-178 with i = 4).
// Both chains of a worker as packed pairs (registers only: pure renaming).
struct Pts2 {
    F2 r[8], i[8];
};
AACFB_HD Pts2 pts_pack(const Pts &z) {
    Pts2 p;
#pragma unroll
    for (int q = 0; q < 8; ++q) { p.r[q] = F2{z.r[0][q], z.r[1][q]}; p.i[q] = F2{z.i[0][q], z.i[1][q]}; }
    return p;
}
AACFB_HD void pts_unpack(const Pts2 &p, Pts &z) {
#pragma unroll
    for (int q = 0; q < 8; ++q) { z.r[0][q] = p.r[q].x; z.r[1][q] = p.r[q].y; z.i[0][q] = p.i[q].x; z.i[1][q] = p.i[q].y; }
}

// PK: run the two chains as packed pairs (NCH == 2 only).
template <int C0, int NCH, bool PK>
AACFB_HD void pass_a(Pts &z, const float2 *wA) {
    if constexpr (PK && NCH == 2) {
        Pts2 p = pts_pack(z);
        base4(&p.r[0], &p.i[0]);
        base4(&p.r[4], &p.i[4]);
        bfly1(p.r[0], p.i[0], p.r[4], p.i[4]);
#pragma unroll
        for (int k = 1; k < 4; ++k) bfly(p.r[k], p.i[k], p.r[4 + k], p.i[4 + k], wA[k].x, wA[k].y);
        pts_unpack(p, z);
    } else {
#pragma unroll
        for (int c = C0; c < C0 + NCH; ++c) {
            base4(&z.r[c][0], &z.i[c][0]);
            base4(&z.r[c][4], &z.i[c][4]);
            bfly1(z.r[c][0], z.i[c][0], z.r[c][4], z.i[c][4]);
#pragma unroll
            for (int k = 1; k < 4; ++k) bfly(z.r[c][k], z.i[c][k], z.r[c][4 + k], z.i[c][4 + k], wA[k].x, wA[k].y);
        }
    }
}

// Three stages on 8 points whose array positions differ in three consecutive
// bits (regs indexed by those bits).  tw[0] serves the first stage, tw[1..2]
// the second, tw[3..6] the third (see SynthTables::twB/twC/twS).
template <int C0, int NCH, bool PK>
AACFB_HD void pass_3stage(Pts &z, const float2 *tw) {
    if constexpr (PK && NCH == 2) {
        Pts2 p = pts_pack(z);
#pragma unroll
        for (int q = 0; q < 8; q += 2) bfly(p.r[q], p.i[q], p.r[q + 1], p.i[q + 1], tw[0].x, tw[0].y);
#pragma unroll
        for (int q = 0; q < 8; q += 4) {
            bfly(p.r[q], p.i[q], p.r[q + 2], p.i[q + 2], tw[1].x, tw[1].y);
            bfly(p.r[q + 1], p.i[q + 1], p.r[q + 3], p.i[q + 3], tw[2].x, tw[2].y);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) bfly(p.r[q], p.i[q], p.r[q + 4], p.i[q + 4], tw[3 + q].x, tw[3 + q].y);
        pts_unpack(p, z);
    } else {
#pragma unroll
        for (int c = C0; c < C0 + NCH; ++c) {
#pragma unroll
            for (int q = 0; q < 8; q += 2) bfly(z.r[c][q], z.i[c][q], z.r[c][q + 1], z.i[c][q + 1], tw[0].x, tw[0].y);
#pragma unroll
            for (int q = 0; q < 8; q += 4) {
                bfly(z.r[c][q], z.i[c][q], z.r[c][q + 2], z.i[c][q + 2], tw[1].x, tw[1].y);
                bfly(z.r[c][q + 1], z.i[c][q + 1], z.r[c][q + 3], z.i[c][q + 3], tw[2].x, tw[2].y);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) bfly(z.r[c][q], z.i[c][q], z.r[c][q + 4], z.i[c][q + 4], tw[3 + q].x, tw[3 + q].y);
        }
    }
}

// ------------------------------------------------------------ long transform
// Thread roles.  Pass A / pass C thread u: input indices n = u + 64*j and
// output bins k = 64*q + u.  Pass B thread v holds array positions
// p = 64*bhi + 8*q + blo with (b7 b6 b8 b2 b1 b0) = bits of v.
AACFB_HD int passb_bhi(int v) { return (((v >> 3) & 1) << 2) | (((v >> 5) & 1) << 1) | ((v >> 4) & 1); }
AACFB_HD int passb_blo(int v) { return v & 7; }

// The MDCT twiddles a thread needs, cs2048[u + 64 j] (j = 0..7, in the pre- and in the
// post-twiddle), have angles pi/16 apart: (c, s)[u + 64 j] = (c, s)[u] rotated by j pi/16.  Deriving
// seven of the eight from one table load costs 4 immediate-operand FMAs each and saves seven
// 8-byte shared-memory loads per phase -- the synthesis kernel is bound by the shared-memory
// pipe (time tracks wavefronts, profiles/), not by issue slots.  The derived values differ from
// the f32-rounded table by <= 2 ulp (tests: PCM parity unchanged at the 1e-7 level).
AACFB_HD constexpr float rot_cos(int j) {
    return j == 0 ? 1.0f : j == 1 ? 0.98078528040323043f : j == 2 ? 0.92387953251128674f : j == 3 ? 0.83146961230254524f
         : j == 4 ? 0.70710678118654757f : j == 5 ? 0.55557023301960229f : j == 6 ? 0.38268343236508984f
                                                                          : 0.19509032201612833f;
}
AACFB_HD constexpr float rot_sin(int j) { return j == 0 ? 0.0f : rot_cos(8 - j); }
template <bool ROT>
AACFB_HD float2 cs_at(const float2 *cs2048, float2 cs0, int u, int j) {
    if (!ROT) return cs2048[u + 64 * j];
    if (j == 0) return cs0;
    float2 r;
    r.x = f_fma(cs0.x, rot_cos(j), -f_mul(cs0.y, rot_sin(j)));
    r.y = f_fma(cs0.y, rot_cos(j), f_mul(cs0.x, rot_sin(j)));
    return r;
}

// Pre-twiddle (mdct.js:73-76) straight from the staged spectrum row into the
// bit-reversed register order pass A needs: reg q <- input n = u + 64*brev3(q).
template <int C0, int NCH, bool PK, bool ROT, bool STAG = false>
AACFB_HD void long_load(int u, const float *const *row, const float2 *cs2048, Pts &z) {
    float2 cs0 = {0.f, 0.f};
    if (ROT) cs0 = cs2048[u];
    // AACFB_LONG_STAGGER: x[2n] lives on the even banks and x[1023 - 2n] on the odd ones, and lanes u, u + 16 of a warp
    // share a bank in either: 2 wavefronts per scalar load.  Half of the lanes (bit 4 of u) fetch the two in the
    // opposite order -- 1023 - 2n = 2n ^ 1023 -- so that every load instruction covers 16 even and 16 odd banks (1
    // wavefront: 32 of the 312 wavefronts of a long channel-frame), and swap the results back with two selects.
    // STAG: only the long-only instantiation does it (config 2 +0.75 %); in the generic ones, at the register cap,
    // the extra live values cost the long path 5 % (config 5), see profiles/r04_ab_experiments.txt.
    const bool sw = STAG && (u & 16) != 0;
    const int flip = sw ? 1023 : 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int n = u + 64 * brev3(q);
        const float2 cs = cs_at<ROT>(cs2048, cs0, u, brev3(q));
        const int ia = (2 * n) ^ flip, ib = ia ^ 1023;
        if constexpr (PK && NCH == 2) {
            const float a0 = row[0][ia], a1 = row[1][ia], b0 = row[0][ib], b1 = row[1][ib];
            const F2 x0{sw ? b0 : a0, sw ? b1 : a1}, x1{sw ? a0 : b0, sw ? a1 : b1};
            const F2 zi = f_fma(x0, cs.x, f_mul(x1, cs.y));
            const F2 zr = f_fma(x1, cs.x, f_neg(f_mul(x0, cs.y)));
            z.i[0][q] = zi.x; z.i[1][q] = zi.y; z.r[0][q] = zr.x; z.r[1][q] = zr.y;
        } else {
#pragma unroll
            for (int c = C0; c < C0 + NCH; ++c) {
                const float a = row[c][ia], b = row[c][ib];
                const float x0 = sw ? b : a, x1 = sw ? a : b;
                z.i[c][q] = f_fma(x0, cs.x, f_mul(x1, cs.y));
                z.r[c][q] = f_fma(x1, cs.x, -f_mul(x0, cs.y));
            }
        }
    }
}

// Exchange 1 (pass A -> pass B).  Element index e = array position; stored at
// e ^ (e >> 5) so that both the scatter below and the gather in ex1_read hit
// 16 distinct 8-byte bank pairs per half-warp.
template <int C0, int NCH, bool PK>
AACFB_HD void ex1_write(int u, const Pts &z, float2 *const *buf) {
    const int a = brev6(u);
    const int base = (8 * a) ^ (a >> 2);
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int c = C0; c < C0 + NCH; ++c) {
            float2 t;
            // two chains: buffer 0 holds the real parts of both, buffer 1 the imaginary parts, so
            // that a packed (chain 0, chain 1) register pair is one 8-byte access
            if (PK && NCH == 2) { t.x = c == 0 ? z.r[0][q] : z.i[0][q]; t.y = c == 0 ? z.r[1][q] : z.i[1][q]; }
            else { t.x = z.r[c][q]; t.y = z.i[c][q]; }
            buf[c][base ^ q] = t;
        }
}
template <int C0, int NCH, bool PK>
AACFB_HD void ex1_read(int v, float2 *const *buf, Pts &z) {
    const int bhi = passb_bhi(v), blo = passb_blo(v);
    const int base = ((64 * bhi) | blo) ^ (bhi << 1);
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int c = C0; c < C0 + NCH; ++c) {
            const float2 t = buf[c][base ^ ((8 * q) ^ (q >> 2))];
            if (PK && NCH == 2) { if (c == 0) { z.r[0][q] = t.x; z.r[1][q] = t.y; } else { z.i[0][q] = t.x; z.i[1][q] = t.y; } }
            else { z.r[c][q] = t.x; z.i[c][q] = t.y; }
        }
}
// Exchange 2 (pass B -> pass C): stored at e ^ (bit8(e) << 3).
template <int C0, int NCH, bool PK>
AACFB_HD void ex2_write(int v, const Pts &z, float2 *const *buf) {
    const int bhi = passb_bhi(v), blo = passb_blo(v);
    const int base = ((64 * bhi) | blo) ^ ((bhi >> 2) << 3);
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int c = C0; c < C0 + NCH; ++c) {
            float2 t;
            // two chains: buffer 0 holds the real parts of both, buffer 1 the imaginary parts, so
            // that a packed (chain 0, chain 1) register pair is one 8-byte access
            if (PK && NCH == 2) { t.x = c == 0 ? z.r[0][q] : z.i[0][q]; t.y = c == 0 ? z.r[1][q] : z.i[1][q]; }
            else { t.x = z.r[c][q]; t.y = z.i[c][q]; }
            buf[c][base ^ (8 * q)] = t;
        }
}
template <int C0, int NCH, bool PK>
AACFB_HD void ex2_read(int u, float2 *const *buf, Pts &z) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int c = C0; c < C0 + NCH; ++c) {
            const float2 t = buf[c][(64 * q + u) ^ ((q >> 2) << 3)];
            if (PK && NCH == 2) { if (c == 0) { z.r[0][q] = t.x; z.r[1][q] = t.y; } else { z.i[0][q] = t.x; z.i[1][q] = t.y; } }
            else { z.r[c][q] = t.x; z.i[c][q] = t.y; }
        }
}

// Finished samples that cannot leave yet (frames with an EIGHT_SHORT chain
// finish chain by chain) wait in registers.
struct Out {
    float a[2][8], b[2][8];
};

// Where the PCM of the current frame goes.
struct OutDst {
    bool emit;          // false for the halo frame of a chunk
    bool interleaved;   // chains are channels 0, 1 of one stereo stream: pcm[n][2]
    float scale;        // 2^-15 (decoder.js:210) or 1 for the inner seam; a power of two
    float inv_scale;    // 1 / scale
    float *out0, *out1; // sample 0 of this frame for each chain (int16_t * behind the cast when s16)
    int ostride;        // distance between successive samples of one chain
    bool s16;           // AACFB_PCM_S16: samples leave as int16 (scale is 1 then); constant false in the
                        // float-only instantiations, where everything behind it folds away
};

// Float sample (un-normalised, decoder.js:210 before the division) -> int16 the way a JS sink does it:
// Int16Array[i] = max(-32768, min(32767, Math.round(x))): round half UP, saturate, NaN -> 0
// (include/aacfb.h, AACFB_PCM_S16), i.e. floor(x + 0.5) with the sum taken exactly.
AACFB_HD int pcm_s16(float x) {
#if defined(__CUDA_ARCH__)
    // floor(x + 0.5) exactly, in two instructions: the sum rounded DOWN is the largest float <= the exact
    // sum, and no integer lies between the two (integers of that size are floats), so their floors agree.
    // cvt.rmi saturates at the int32 range and turns NaN into 0.
    int r = __float2int_rd(__fadd_rd(x, 0.5f));
    r = r < -32768 ? -32768 : r;
    return r > 32767 ? 32767 : r;
#else
    if (!(x == x)) return 0;
    double r = floor((double)x + 0.5);
    r = r < -32768.0 ? -32768.0 : r;
    r = r > 32767.0 ? 32767.0 : r;
    return (int)r;
#endif
}

// Effective windows of a long-transform frame as (value at m, value at 1023-m):
//   first half : ONLY_LONG/LONG_START -> long window of shape_prev (filter_bank.js:109-111,124-126)
//                LONG_STOP            -> 0 | short asc | 1            (filter_bank.js:184-194)
//   second half: ONLY_LONG/LONG_STOP  -> reversed long window of shape_cur (filter_bank.js:114-116,198-200)
//                LONG_START           -> 1 | short desc | 0           (filter_bank.js:129-139)
// All four live in tables of the same layout, so a frame only picks two table pointers per
// chain and the arithmetic is the same for every sequence.  `wz` is the worker's shared-memory
// copy of the long windows, `g` the full table set in global memory (window switching is rare);
// both carry the output scale already (scale_windows, aacfb_tables.h).  A second-half entry w
// is used as (w.y, w.x).
struct LongWin {
    const float2 *first, *second;
};
AACFB_HD LongWin long_windows(FrameBits fi, const float2 (*wz)[512], const SynthTables *g) {
    LongWin w;
    w.first = fb_seq(fi) == AACFB_LONG_STOP_SEQUENCE ? g->fwz_stop[fb_shape_prev(fi)] : wz[fb_shape_prev(fi)];
    w.second = fb_seq(fi) == AACFB_LONG_START_SEQUENCE ? g->swz_start[fb_shape_cur(fi)] : wz[fb_shape_cur(fi)];
    return w;
}

// ------------------------------------------------------------- PCM write-out
// Thread u holds, for q = 0..7, the samples at m = long_pos_of_bin(64q+u)
// (even) and at 1023-m (odd).  The odd neighbour m+1 of its own q is the
// mirrored sample of q' = 7-q of thread 63-u.  Threads are numbered so that
// 63-u sits in the same warp at lane ^ 31 (see worker_thread_index), hence one
// shuffle per sample turns the scattered ownership into contiguous runs
//   interleaved stereo: (L[m], R[m], L[m+1], R[m+1]) -> one 16-byte store
//   planar / mono:      (x[m], x[m+1])               -> one  8-byte store
// and every warp-level store covers whole 32-byte sectors of the PCM row
// (decoder.js:204-213 interleave).  a[h][c] / b[h][c]: samples of chain c at
// m and 1023-m for q = qq (h = 0) and q = 7-qq (h = 1).
template <int C0, int NCH, class Sync>
AACFB_HD void emit_pair(int u, Sync &sync, int qq, const float (*a)[2], const float (*b)[2], const OutDst &d) {
    float r_lo[2], r_hi[2];
#pragma unroll
    for (int c = C0; c < C0 + NCH; ++c) {
        r_lo[c] = sync.partner(u, b[1][c]);  // partner's 1023-m' of q' = 7-qq  == my m(qq) + 1
        r_hi[c] = sync.partner(u, b[0][c]);  // partner's 1023-m' of q' = qq    == my m(7-qq) + 1
    }
    const int m_lo = long_pos_of_bin(64 * qq + u), m_hi = long_pos_of_bin(64 * (7 - qq) + u);
    if (d.s16) {   // same runs as below, 2 bytes per sample
        if (NCH == 2 && d.interleaved) {
            int16_t *o = reinterpret_cast<int16_t *>(d.out0);
            uint2 v;
            v.x = (uint32_t)(pcm_s16(a[0][0]) & 0xffff) | ((uint32_t)pcm_s16(a[0][1]) << 16);
            v.y = (uint32_t)(pcm_s16(r_lo[0]) & 0xffff) | ((uint32_t)pcm_s16(r_lo[1]) << 16);
            *reinterpret_cast<uint2 *>(o + 2 * m_lo) = v;
            v.x = (uint32_t)(pcm_s16(a[1][0]) & 0xffff) | ((uint32_t)pcm_s16(a[1][1]) << 16);
            v.y = (uint32_t)(pcm_s16(r_hi[0]) & 0xffff) | ((uint32_t)pcm_s16(r_hi[1]) << 16);
            *reinterpret_cast<uint2 *>(o + 2 * m_hi) = v;
        } else {
#pragma unroll
            for (int c = C0; c < C0 + NCH; ++c) {
                int16_t *o = reinterpret_cast<int16_t *>(c == 0 ? d.out0 : d.out1);
                if (d.ostride == 1) {
                    *reinterpret_cast<uint32_t *>(o + m_lo) = (uint32_t)(pcm_s16(a[0][c]) & 0xffff) | ((uint32_t)pcm_s16(r_lo[c]) << 16);
                    *reinterpret_cast<uint32_t *>(o + m_hi) = (uint32_t)(pcm_s16(a[1][c]) & 0xffff) | ((uint32_t)pcm_s16(r_hi[c]) << 16);
                } else {
                    o[(size_t)m_lo * d.ostride] = (int16_t)pcm_s16(a[0][c]);
                    o[(size_t)(m_lo + 1) * d.ostride] = (int16_t)pcm_s16(r_lo[c]);
                    o[(size_t)m_hi * d.ostride] = (int16_t)pcm_s16(a[1][c]);
                    o[(size_t)(m_hi + 1) * d.ostride] = (int16_t)pcm_s16(r_hi[c]);
                }
            }
        }
        return;
    }
    if (NCH == 2 && d.interleaved) {
        float4 v;
        v.x = a[0][0]; v.y = a[0][1]; v.z = r_lo[0]; v.w = r_lo[1];
        *reinterpret_cast<float4 *>(d.out0 + 2 * m_lo) = v;
        v.x = a[1][0]; v.y = a[1][1]; v.z = r_hi[0]; v.w = r_hi[1];
        *reinterpret_cast<float4 *>(d.out0 + 2 * m_hi) = v;
    } else {
#pragma unroll
        for (int c = C0; c < C0 + NCH; ++c) {
            float *o = c == 0 ? d.out0 : d.out1;
            if (d.ostride == 1) {
                float2 v;
                v.x = a[0][c]; v.y = r_lo[c];
                *reinterpret_cast<float2 *>(o + m_lo) = v;
                v.x = a[1][c]; v.y = r_hi[c];
                *reinterpret_cast<float2 *>(o + m_hi) = v;
            } else {
                o[(size_t)m_lo * d.ostride] = a[0][c];
                o[(size_t)(m_lo + 1) * d.ostride] = r_lo[c];
                o[(size_t)m_hi * d.ostride] = a[1][c];
                o[(size_t)(m_hi + 1) * d.ostride] = r_hi[c];
            }
        }
    }
}
// Write-out of samples parked in registers (frames with a short chain).
template <int C0, int NCH, class Sync>
AACFB_HD void out_store(int u, Sync &sync, const Out &o, const OutDst &d) {
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
        float a[2][2], b[2][2];
#pragma unroll
        for (int c = C0; c < C0 + NCH; ++c) {
            a[0][c] = o.a[c][qq]; b[0][c] = o.b[c][qq];
            a[1][c] = o.a[c][7 - qq]; b[1][c] = o.b[c][7 - qq];
        }
        emit_pair<C0, NCH>(u, sync, qq, a, b, d);
    }
}

// Post-twiddle (mdct.js:82-87), reorder (mdct.js:90-114), window and
// overlap-add (filter_bank.js:105-141,180-202); the scale of decoder.js:210
// rides in the window tables (scale_windows).  Thread u owns bins k = 64q+u, i.e. output positions m and 1023-m with
// m = long_pos_of_bin(k); the same thread owned them in every earlier frame,
// so the overlap lives in registers.
//   UNIFORM  : all chains are ONLY_LONG with the same shapes (the common case):
//              one shared-memory window load serves every chain and both halves.
//   TO_GLOBAL: store the PCM right away (else park it in `o`).
template <int C0, int NCH, bool UNIFORM, bool TO_GLOBAL, bool PK, bool ROT, class Sync>
AACFB_HD void long_finish(int u, Sync &sync, const Pts &z, Ovl &ov, const SynthTables *ts, const SynthTables *tg,
                          const FrameBits *fi, const OutDst &d, Out &o) {
    LongWin win[2];
#pragma unroll
    for (int c = C0; c < C0 + NCH; ++c) win[c] = long_windows(fi[UNIFORM ? C0 : c], ts->wz, tg);
    float2 cs0 = {0.f, 0.f};
    if (ROT) cs0 = ts->cs2048[u];
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
        float a[2][2], b[2][2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int q = h ? 7 - qq : qq;
            const int k = 64 * q + u;
            const float2 cs = cs_at<ROT>(ts->cs2048, cs0, u, q);
            float2 wf_u, ws_u;
            if (UNIFORM) {
                wf_u = ts->wz[fb_shape_prev(fi[C0])][k];
                ws_u = fb_shape_prev(fi[C0]) == fb_shape_cur(fi[C0]) ? wf_u : ts->wz[fb_shape_cur(fi[C0])][k];
            }
            if constexpr (PK && NCH == 2) {   // both chains at once (packed pairs)
                const F2 re{z.r[0][q], z.r[1][q]}, im{z.i[0][q], z.i[1][q]};
                const F2 pr = f_fma(re, cs.x, f_neg(f_mul(im, cs.y)));
                const F2 pi = f_fma(im, cs.x, f_mul(re, cs.y));
                const F2 F = (q < 4) ? pr : pi;
                const F2 S = (q < 4) ? f_neg(pi) : pr;
                if (d.emit) {
                    F2 wx, wy;
                    if (UNIFORM) { wx = F2{wf_u.x, wf_u.x}; wy = F2{wf_u.y, wf_u.y}; }
                    else {
                        const float2 w0 = win[0].first[k], w1 = win[1].first[k];
                        wx = F2{w0.x, w1.x}; wy = F2{w0.y, w1.y};
                    }
                    const F2 va = f_fma(F, wx, F2{ov.a[0][q], ov.a[1][q]});
                    const F2 vb = f_fma(f_neg(F), wy, F2{ov.b[0][q], ov.b[1][q]});
                    a[h][0] = va.x; a[h][1] = va.y; b[h][0] = vb.x; b[h][1] = vb.y;
                }
                F2 sx, sy;
                if (UNIFORM) { sx = F2{ws_u.x, ws_u.x}; sy = F2{ws_u.y, ws_u.y}; }
                else {
                    const float2 w0 = win[0].second[k], w1 = win[1].second[k];
                    sx = F2{w0.x, w1.x}; sy = F2{w0.y, w1.y};
                }
                const F2 oa = f_mul(S, sy), ob = f_mul(S, sx);
                ov.a[0][q] = oa.x; ov.a[1][q] = oa.y; ov.b[0][q] = ob.x; ov.b[1][q] = ob.y;
            } else {
#pragma unroll
                for (int c = C0; c < C0 + NCH; ++c) {
                    const float re = z.r[c][q], im = z.i[c][q];
                    const float pr = f_fma(re, cs.x, -f_mul(im, cs.y));
                    const float pi = f_fma(im, cs.x, f_mul(re, cs.y));
                    // first-half sample at m is F, at 1023-m is -F; second-half sample is S at both
                    const float F = (q < 4) ? pr : pi;
                    const float S = (q < 4) ? -pi : pr;
                    if (d.emit) {
                        const float2 wf = UNIFORM ? wf_u : win[c].first[k];
                        a[h][c] = f_fma(F, wf.x, ov.a[c][q]);
                        b[h][c] = f_fma(-F, wf.y, ov.b[c][q]);
                    }
                    const float2 ws = UNIFORM ? ws_u : win[c].second[k];
                    ov.a[c][q] = f_mul(S, ws.y);
                    ov.b[c][q] = f_mul(S, ws.x);
                }
            }
        }
        if (d.emit) {
            if (TO_GLOBAL) emit_pair<C0, NCH>(u, sync, qq, a, b, d);
            else {
#pragma unroll
                for (int c = C0; c < C0 + NCH; ++c) {
                    o.a[c][qq] = a[0][c]; o.b[c][qq] = b[0][c];
                    o.a[c][7 - qq] = a[1][c]; o.b[c][7 - qq] = b[1][c];
                }
            }
        }
    }
}

// ----------------------------------------------------------- short transform
// EIGHT_SHORT: thread u = 8*w + g works on window w.  Pass A' takes inputs
// n = g + 8*j of that window, pass B' produces bins k = 8*q + g.
//
// Packed two-chain path: like the long transform's stride-2 reads, scalar loads of x[2n] and
// x[127 - 2n] only ever touch every other bank, and here four windows of a warp land on the same
// eight banks (4-way conflicts).  Reading 8-byte pairs (x[2n], x[2n+1]) halves that: the odd
// element is the x[127 - 2n'] of n' = 63 - n = (7-g) + 8(7-j), i.e. what thread u ^ 7 (lane ^ 7)
// needs for its register 7-q, so one shuffle per value replaces the second load.
template <int C0, int NCH, bool PK, class Sync>
AACFB_HD void short_load(int u, Sync &sync, const float *const *row, const float2 *cs256, Pts &z) {
    const int w = u >> 3, g = u & 7;
    if constexpr (PK && NCH == 2 && AACFB_SHORT_PAIRLOAD != 0) {
        const float2 *r0 = reinterpret_cast<const float2 *>(row[0]) + 64 * w, *r1 = reinterpret_cast<const float2 *>(row[1]) + 64 * w;
        // AACFB_SHORT_STAGGER: pair n = g + 8 b sits on banks 2g, 2g + 1 (+ 16 if b is odd) whatever the window, so the
        // two windows of a half-warp collide (4 wavefronts per LDS.64 instead of 2: 52 of the 486 wavefronts of an
        // EIGHT_SHORT channel-frame).  b and 7 - b have opposite parity and are fetched as a pair anyway: odd windows
        // fetch them in the opposite order (n ^ 56) and swap the results back.
        const bool sw = AACFB_SHORT_STAGGER != 0 && (w & 1) != 0;
        const int flip = sw ? 56 : 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int qa = q, qb = 7 - q;   // brev3(7 - q) = 7 - brev3(q)
            const int na = g + 8 * brev3(qa), nb = g + 8 * brev3(qb);
            float2 a0 = r0[na ^ flip], a1 = r1[na ^ flip], b0 = r0[nb ^ flip], b1 = r1[nb ^ flip];
            if (AACFB_SHORT_STAGGER != 0) {
                const float2 t0 = a0, t1 = a1;
                a0.x = sw ? b0.x : t0.x; a0.y = sw ? b0.y : t0.y; a1.x = sw ? b1.x : t1.x; a1.y = sw ? b1.y : t1.y;
                b0.x = sw ? t0.x : b0.x; b0.y = sw ? t0.y : b0.y; b1.x = sw ? t1.x : b1.x; b1.y = sw ? t1.y : b1.y;
            }
            const F2 x1a{sync.partner7(u, b0.y), sync.partner7(u, b1.y)};
            const F2 x1b{sync.partner7(u, a0.y), sync.partner7(u, a1.y)};
            const F2 x0a{a0.x, a1.x}, x0b{b0.x, b1.x};
            const float2 ca = cs256[na], cb = cs256[nb];
            const F2 zia = f_fma(x0a, ca.x, f_mul(x1a, ca.y)), zra = f_fma(x1a, ca.x, f_neg(f_mul(x0a, ca.y)));
            const F2 zib = f_fma(x0b, cb.x, f_mul(x1b, cb.y)), zrb = f_fma(x1b, cb.x, f_neg(f_mul(x0b, cb.y)));
            z.i[0][qa] = zia.x; z.i[1][qa] = zia.y; z.r[0][qa] = zra.x; z.r[1][qa] = zra.y;
            z.i[0][qb] = zib.x; z.i[1][qb] = zib.y; z.r[0][qb] = zrb.x; z.r[1][qb] = zrb.y;
        }
    } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int n = g + 8 * brev3(q);
            const float2 cs = cs256[n];
            if constexpr (PK && NCH == 2) {
                const int i0 = 128 * w + 2 * n, i1 = 128 * w + 127 - 2 * n;
                const F2 x0{row[0][i0], row[1][i0]}, x1{row[0][i1], row[1][i1]};
                const F2 zi = f_fma(x0, cs.x, f_mul(x1, cs.y));
                const F2 zr = f_fma(x1, cs.x, f_neg(f_mul(x0, cs.y)));
                z.i[0][q] = zi.x; z.i[1][q] = zi.y; z.r[0][q] = zr.x; z.r[1][q] = zr.y;
            } else {
#pragma unroll
                for (int c = C0; c < C0 + NCH; ++c) {
                    const float x0 = row[c][128 * w + 2 * n], x1 = row[c][128 * w + 127 - 2 * n];
                    z.i[c][q] = f_fma(x0, cs.x, f_mul(x1, cs.y));
                    z.r[c][q] = f_fma(x1, cs.x, -f_mul(x0, cs.y));
                }
            }
        }
    }
}
// Exchange between the two passes of the 8 x 64-point FFTs.  Element
// e = 64w + position; stored at e ^ (12*bit6(e)) ^ ((e>>4)&3).
template <int C0, int NCH, bool PK>
AACFB_HD void exs_write(int u, const Pts &z, float2 *const *buf) {
    const int w = u >> 3, a = brev3(u & 7);
    const int base = ((64 * w) | (8 * a)) ^ (12 * (w & 1)) ^ (a >> 1);
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int c = C0; c < C0 + NCH; ++c) {
            float2 t;
            // two chains: buffer 0 holds the real parts of both, buffer 1 the imaginary parts, so
            // that a packed (chain 0, chain 1) register pair is one 8-byte access
            if (PK && NCH == 2) { t.x = c == 0 ? z.r[0][q] : z.i[0][q]; t.y = c == 0 ? z.r[1][q] : z.i[1][q]; }
            else { t.x = z.r[c][q]; t.y = z.i[c][q]; }
            buf[c][base ^ q] = t;
        }
}
template <int C0, int NCH, bool PK>
AACFB_HD void exs_read(int u, float2 *const *buf, Pts &z) {
    const int w = u >> 3, g = u & 7;
    const int base = ((64 * w) | g) ^ (12 * (w & 1));
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int c = C0; c < C0 + NCH; ++c) {
            const float2 t = buf[c][base ^ ((8 * q) ^ (q >> 1))];
            if (PK && NCH == 2) { if (c == 0) { z.r[0][q] = t.x; z.r[1][q] = t.y; } else { z.i[0][q] = t.x; z.i[1][q] = t.y; } }
            else { z.r[c][q] = t.x; z.i[c][q] = t.y; }
        }
}
// ---- EIGHT_SHORT window + overlap-add through two product arrays ----------------------
// filter_bank.js:148-176 adds, at position t = 128j + i of the 1152-sample short-window
// sequence (j = 0..8, i = 0..127),
//     Z[t] = y_{j-1}[128 + i] * W[127 - i]  +  y_j[i] * W'[i]
// (y_w = the 256 IMDCT outputs of window w, y_{-1} = y_8 = 0, W = short window of shape_cur,
// W' = shape_prev's for j = 0), with  out[n] = overlap[n] + Z[n - 448]  for n >= 448  and
// overlap'[n] = Z[n + 576]  for n < 576, 0 above.  The thread that holds bin k of window w after
// the FFT owns exactly four IMDCT outputs of that window (mdct.js:90-114 writes every post-twiddled
// value at two mirrored positions), so it multiplies them by their window values right away and
// stores the PRODUCTS:
//     P1[128w + i]  = y_w[i]       * W'[i]          (first halves)
//     P2[128w + i]  = y_w[128 + i] * W[127 - i]     (second halves, i.e. Z's first term at t = 128(w+1) + i)
// 2 x 1024 floats = one 8 KiB staging buffer per chain.  The consumer then needs two loads per
// Z value instead of four loads, two multiplies and the window-sequence branching.
// Index swizzle (both arrays): flipping bit 0 with bit 5 makes the consumers' stride-2 reads
// (threads own positions 2u + const) conflict-free; flipping bit 4 with bit 7 spreads the four
// windows a warp's producers write at once over both halves of the banks.
// AACFB_SWZ9: bit 0 additionally flips with bit 9 (bit 2 of the window index): the four windows a
// warp's producers write at once ({0,1,6,7} / {2,3,4,5}) then cover all 32 banks instead of 16
// (no 2-way store conflicts), at the price of 2-way conflicts in the quarter of the consumer
// loads whose two half-warps straddle a 512 boundary.
#ifndef AACFB_SWZ9
#define AACFB_SWZ9 1
#endif
AACFB_HD constexpr int short_swz(int t) {
    return AACFB_SWZ9 ? t ^ (((t >> 5) ^ (t >> 9)) & 1) ^ (((t >> 7) & 1) << 4) : t ^ ((t >> 5) & 1) ^ (((t >> 7) & 1) << 4);
}
// AACFB_SWZ_LINEAR: the swizzle is linear over GF(2) (XORs of bit extractions), so for an index that is a sum of
// parts with disjoint bits -- a per-thread part and a part known at compile time -- it splits into
// short_swz(thread part) ^ short_swz(constant part): one XOR per address instead of the whole bit fiddle
// (the address arithmetic of the product arrays was ~30 % of the instructions of an EIGHT_SHORT frame).
//   producer   128 w + pa = (128 w + 2 g) + c_q        128 w + pb = (128 w + 15 - 2 g) + c'_q      (c_q, c'_q: bits 4-6)
//   consumer   t = C + 2u: C = 128 j + 64 for every index the sums use -> 128 (j + [u >= 32]) + ((64 + 2u) & 127)
//              t = C - 2u: C = 128 j + 63                              -> 128 (j - [u >= 32]) + ((63 - 2u) & 127)
#ifndef AACFB_SWZ_LINEAR
#define AACFB_SWZ_LINEAR 1
#endif
struct ShortIdx {          // per-thread parts of the consumer's indices
    int lo_up, lo_dn;      // short_swz((64 + 2u) & 127), short_swz((63 - 2u) & 127)
    bool hi;               // u >= 32
};
AACFB_HD ShortIdx short_idx(int u) {
    ShortIdx s;
    s.lo_up = short_swz((64 + 2 * u) & 127); s.lo_dn = short_swz((63 - 2 * u) & 127); s.hi = u >= 32;
    return s;
}
// swizzled index of t = CT + 2u (NEG = false) or CT - 2u (NEG = true); only meaningful for 0 <= t < 1024
template <int CT, bool NEG>
AACFB_HD int short_swz_at(const ShortIdx &s) {
    constexpr int cl = NEG ? 63 : 64;
    static_assert((CT - cl) % 128 == 0, "index constant");
    constexpr int j0 = (CT - cl) / 128, j1 = NEG ? j0 - 1 : j0 + 1;
    constexpr int k0 = (j0 >= 0 && j0 < 8) ? short_swz(128 * j0) : 0, k1 = (j1 >= 0 && j1 < 8) ? short_swz(128 * j1) : 0;
    return (NEG ? s.lo_dn : s.lo_up) ^ (s.hi ? k1 : k0);
}

// Producer: thread u = 8w + g, bins k = 8q + g of window w.  `wsp[shape][k]` = (W[pa(k)], W[pb(k)]):
// the two window values a bin needs, as one 8-byte entry (same for every window w: a broadcast
// load); it carries the output scale (scale_windows), so the products are in output units like
// the overlap registers.
template <int C>
AACFB_HD void short_products(int u, const Pts &z, const float2 *cs256, const float2 (*wsp)[64], FrameBits fi,
                             float *buf) {
    const int w = u >> 3, g = u & 7;
    const float2 *wcur = wsp[fb_shape_cur(fi)];
    const bool first_differs = fb_shape_prev(fi) != fb_shape_cur(fi);   // (uniform over the worker)
    const float2 *wfirst = w == 0 ? wsp[fb_shape_prev(fi)] : wcur;      // filter_bank.js:153 vs :157-160
    float *p1 = buf, *p2 = buf + 1024;
    const int ta = short_swz(128 * w + 2 * g), tb = short_swz(128 * w + 15 - 2 * g);   // (AACFB_SWZ_LINEAR)
    (void)ta; (void)tb;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int k = 8 * q + g;
        const float2 cs = cs256[k];
        const float re = z.r[C][q], im = z.i[C][q];
        const float pr = f_fma(re, cs.x, -f_mul(im, cs.y));
        const float pi = f_fma(im, cs.x, f_mul(re, cs.y));
        // the two positions (one even, one odd) of this bin in either half of the window
        const int pa = q < 4 ? 64 + 2 * k : 2 * (k - 32);
        const int pb = q < 4 ? 63 - 2 * k : 191 - 2 * k;
        const float2 wc = wcur[k];
        float2 wf = wc;
        if (first_differs) wf = wfirst[k];
        const float wa = wc.x, wb = wc.y, fa = wf.x, fb = wf.y;
#if AACFB_SWZ_LINEAR
        const int ia = ta ^ short_swz(q < 4 ? 64 + 16 * q : 16 * (q - 4)), ib = tb ^ short_swz(q < 4 ? 48 - 16 * q : 176 - 16 * q);
        (void)pa; (void)pb;
#else
        const int ia = short_swz(128 * w + pa), ib = short_swz(128 * w + pb);
#endif
        if (q < 4) {  // k < 32: y[64+2k] = pr, y[63-2k] = -pr, y[192+2k] = y[191-2k] = -pi
            p1[ia] = f_mul(pr, fa);
            p1[ib] = f_mul(-pr, fb);
            p2[ia] = f_mul(-pi, wb);   // second-half index i = pa uses W[127 - pa] = W[pb]
            p2[ib] = f_mul(-pi, wa);
        } else {      // k >= 32: y[2(k-32)] = pi, y[191-2k] = -pi, y[128+2(k-32)] = y[319-2k] = pr
            p1[ia] = f_mul(pi, fa);
            p1[ib] = f_mul(-pi, fb);
            p2[ia] = f_mul(pr, wb);
            p2[ib] = f_mul(pr, wa);
        }
    }
}

// Consumer for output position n (range [LO, HI] known at compile time): returns the sample
// out[n] (if EMIT) and replaces the overlap.  Loads that cannot apply to the range fold away.
// NEG: n = HI - 2u (the mirrored positions), otherwise n = LO + 2u.
template <int LO, int HI, bool NEG>
AACFB_HD void short_ola(int n, const ShortIdx &sx, const float *buf, bool emit, float &ovl, float &out) {
    const float *p1 = buf, *p2 = buf + 1024;
    constexpr int C = NEG ? HI : LO;   // n = C -+ 2u
#if AACFB_SWZ_LINEAR
#define AACFB_SIDX(off) short_swz_at<C + (off), NEG>(sx)
#else
#define AACFB_SIDX(off) short_swz(n + (off))
    (void)sx;
#endif
    // out[n] = (overlap + y_{j-1} term) + y_j term, t = n - 448   (filter_bank.js:153-161)
    float o = ovl;
    if (HI >= 576) { const bool on = LO >= 576 || n >= 576; const float v = on ? p2[on ? AACFB_SIDX(-576) : 0] : 0.f; o = f_add(o, v); }
    if (HI >= 448) { const bool on = LO >= 448 || n >= 448; const float v = on ? p1[on ? AACFB_SIDX(-448) : 0] : 0.f; o = f_add(o, v); }
    if (emit) out = o;
    // overlap'[n] = y_{j-1} term + y_j term, t = n + 576             (filter_bank.js:164-176)
    float nv = 0.f;
    if (LO < 576) { const bool on = HI < 576 || n < 576; nv = on ? p2[on ? AACFB_SIDX(448) : 0] : 0.f; }
    if (LO < 448) { const bool on = HI < 448 || n < 448; const float v = on ? p1[on ? AACFB_SIDX(576) : 0] : 0.f; nv = f_add(nv, v); }
    ovl = nv;
#undef AACFB_SIDX
}
// Window + overlap-add of ONE chain of an EIGHT_SHORT frame from its product arrays.
template <int C>
AACFB_HD void short_finish(int u, const float *buf, Ovl &ov, bool emit, Out &o) {
    const ShortIdx sx = short_idx(u);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        // m = long_pos_of_bin(64q + u): 512 + 128q + 2u (q < 4), 128(q - 4) + 2u (q >= 4); mirror 1023 - m
        const int m = long_pos_of_bin(64 * q + u), mm = 1023 - m;
        switch (q) {  // compile-time ranges of m and 1023 - m (u = 0..63)
        case 0: short_ola<512, 638, false>(m, sx, buf, emit, ov.a[C][0], o.a[C][0]); short_ola<385, 511, true>(mm, sx, buf, emit, ov.b[C][0], o.b[C][0]); break;
        case 1: short_ola<640, 766, false>(m, sx, buf, emit, ov.a[C][1], o.a[C][1]); short_ola<257, 383, true>(mm, sx, buf, emit, ov.b[C][1], o.b[C][1]); break;
        case 2: short_ola<768, 894, false>(m, sx, buf, emit, ov.a[C][2], o.a[C][2]); short_ola<129, 255, true>(mm, sx, buf, emit, ov.b[C][2], o.b[C][2]); break;
        case 3: short_ola<896, 1022, false>(m, sx, buf, emit, ov.a[C][3], o.a[C][3]); short_ola<1, 127, true>(mm, sx, buf, emit, ov.b[C][3], o.b[C][3]); break;
        case 4: short_ola<0, 126, false>(m, sx, buf, emit, ov.a[C][4], o.a[C][4]); short_ola<897, 1023, true>(mm, sx, buf, emit, ov.b[C][4], o.b[C][4]); break;
        case 5: short_ola<128, 254, false>(m, sx, buf, emit, ov.a[C][5], o.a[C][5]); short_ola<769, 895, true>(mm, sx, buf, emit, ov.b[C][5], o.b[C][5]); break;
        case 6: short_ola<256, 382, false>(m, sx, buf, emit, ov.a[C][6], o.a[C][6]); short_ola<641, 767, true>(mm, sx, buf, emit, ov.b[C][6], o.b[C][6]); break;
        default: short_ola<384, 510, false>(m, sx, buf, emit, ov.a[C][7], o.a[C][7]); short_ola<513, 639, true>(mm, sx, buf, emit, ov.b[C][7], o.b[C][7]); break;
        }
    }
}

// ---------------------------------------------------------------- overlap I/O
// The state in memory is FilterBank.overlaps (unscaled); registers hold it times `scale`.
template <int C>
AACFB_HD void ovl_load(int u, const float *state, Ovl &ov, float scale) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int m = long_pos_of_bin(64 * q + u);
        ov.a[C][q] = f_mul(state[m], scale);
        ov.b[C][q] = f_mul(state[1023 - m], scale);
    }
}
template <int C>
AACFB_HD void ovl_store(int u, const Ovl &ov, float *state, float inv_scale) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int m = long_pos_of_bin(64 * q + u);
        state[m] = f_mul(ov.a[C][q], inv_scale);
        state[1023 - m] = f_mul(ov.b[C][q], inv_scale);
    }
}

// ------------------------------------------------------ inverse quantisation
// ICStream.decodeSpectralData's arithmetic (ics.js:203-266) on a staged aacfb_qframe record
// (include/aacfb.h): rec = group_len[8] | band[120] u16 | reserved[8] | q[1024] i16.
#if defined(__CUDA_ARCH__)
template <class T> AACFB_HD T dq_ld(const T *p) { return __ldg(p); }
#else
template <class T> AACFB_HD T dq_ld(const T *p) { return *p; }
#endif

// Perceptual noise substitution AS SHIPPED (ics.js:228-242) for the 4 coefficients 4 c4 .. 4 c4 + 3 of
// band (g, sfb), window wg of the group.  The generator restarts at 0x1F2E3D4C in every new ICStream
// (one per element per frame) and its output is a fixed sequence (DequantTables::noise) that is 0
// from index noise_len = 11 on; what a coefficient gets depends only on how many noise
// coefficients the reference generated before it, in its (group, band, window, k) order.
AACFB_HD_NOINLINE void dequant_noise4(const DequantTables *D, const uint8_t *rec, bool is_short, int max_sfb, int g, int wg,
                             int sfb, int c4, uint32_t code, float *out) {
    const uint16_t *band = reinterpret_cast<const uint16_t *>(rec + 8);
    const uint16_t *swb = is_short ? D->swb_short : D->swb_long;
    const int noise_len = dq_ld(&D->noise_len);
    int count = 0;
    bool done = false;
    for (int gg = 0; gg <= g && !done; ++gg) {
        const int len = rec[gg];
        for (int b = 0; b < max_sfb; ++b) {
            if (gg == g && b == sfb) { done = true; break; }
            if ((band[gg * max_sfb + b] & AACFB_BAND_KIND_MASK) == AACFB_BAND_NOISE) {
                count += ((int)dq_ld(swb + b + 1) - (int)dq_ld(swb + b)) * len;
                if (count >= noise_len) { done = true; break; }   // everything from here on is 0 * (sf / sqrt(0))
            }
        }
    }
    const int lo = dq_ld(swb + sfb), width = (int)dq_ld(swb + sfb + 1) - lo;
    const int n0 = count + wg * width;                       // generator outputs consumed before this band-window
    double energy = 0.0;                                      // ics.js:231-237 (zeros add nothing)
    for (int k = 0; k < width && n0 + k < noise_len; ++k) {
        const double v = (double)dq_ld(&D->noise[n0 + k]);
        energy += v * v;
    }
    const double sfv = -(double)dq_ld(&D->sf[code & AACFB_BAND_INDEX_MASK]);   // scaleFactors[idx] = -SCALEFACTOR_TABLE[..], ics.js:158
    const double scale = sfv / sqrt(energy);                  // ics.js:239: -Infinity once the generator is stuck at 0
    const int k0 = ((4 * c4) & (is_short ? 127 : 1023)) - lo;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = n0 + k0 + j;
        const double v = i < noise_len ? (double)dq_ld(&D->noise[i]) : 0.0;
        out[j] = (float)(v * scale);                          // ics.js:240-241 (0 * -Infinity = NaN)
    }
}

// The 4 coefficients of group c4 (scalefactor-band edges are multiples of 4, tables.js:34-124).
AACFB_HD_NOINLINE void dequant4(const DequantTables *D, const uint8_t *rec, FrameBits fi, int c4, const int16_t *q4, float *out) {
    const bool is_short = fb_seq(fi) == AACFB_EIGHT_SHORT_SEQUENCE;
    const int max_sfb = (int)(fi >> 24);
    int sfb, g = 0, wg = 0;
    bool valid = true;
    if (is_short) {   // window w of the frame belongs to group g: groupOff += groupLen << 7, ics.js:259
        const int w = c4 >> 5;
        sfb = dq_ld(&D->sfb_short[c4 & 31]);
        int acc = 0;
        valid = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int len = rec[k];
            if (!valid && len != 0 && w >= acc && w < acc + len) { g = k; wg = w - acc; valid = true; }
            acc += len;
        }
    } else {
        sfb = dq_ld(&D->sfb_long[c4]);
    }
    const int idx = g * max_sfb + sfb;
    uint32_t code = AACFB_BAND_ZERO;
    if (valid && sfb < max_sfb && idx < 120) code = reinterpret_cast<const uint16_t *>(rec + 8)[idx];
    const uint32_t kind = code & AACFB_BAND_KIND_MASK;
    if (kind == AACFB_BAND_SPECTRAL) {          // ics.js:243-256
        const float s = dq_ld(&D->sf[code & AACFB_BAND_INDEX_MASK]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int v = q4[j];
            int a = v < 0 ? -v : v;
            a = a > 8191 ? 8191 : a;              // outside IQ_TABLE: `undefined` -> NaN
            const float t = dq_ld(&D->iq[a]);
            out[j] = f_mul(v > 0 ? t : -t, s);    // (buf[j] > 0) ? IQ[buf[j]] : -IQ[-buf[j]]  (q = 0 gives -0)
        }
    } else if (kind == AACFB_BAND_NOISE) {
        dequant_noise4(D, rec, is_short, max_sfb, g, wg, sfb, c4, code, out);
    } else {                                    // ZERO_BT / intensity (ics.js:222-227) and everything above maxSFB
        out[0] = out[1] = out[2] = out[3] = 0.f;
    }
}

// One whole record -> 1024 floats (the pre-pass kernel and the tests).
AACFB_HD void dequant_row(const DequantTables *D, const uint8_t *rec, FrameBits fi, float *out) {
    const int16_t *q = reinterpret_cast<const int16_t *>(rec + 256);
    for (int c4 = 0; c4 < 256; ++c4) dequant4(D, rec, fi, c4, q + 4 * c4, out + 4 * c4);
}

// Worker phase: the aacfb_qframe records of the frame's rows have landed at byte kQLandOffset of each
// 4 KiB row slot; turn them into float rows in place.  Thread u owns groups u + 64 j: the 8-byte reads
// of q and the 16-byte writes of the result are conflict-free.  All reads happen before the first
// barrier (the floats overwrite the record), the second one publishes the rows.
//
// The lookups are arranged for throughput: which scalefactor band a thread's four groups fall into
// depends on the sample rate alone, so it is read once per kernel (DqCtx::sfb_long4 / sfb_short1); the
// scalefactor table and IQ_TABLE[0..1023] sit in shared memory (gathers through the L1 thrashed: it is
// all but carved out by the staging buffers); band kinds are resolved by selects, not branches, so the
// 4 + 16 independent table reads of a row are in flight together.  |q| >= 1024 and noise bands (rare)
// take the general path through global memory.
constexpr int kQFrameBytes = 2304;
constexpr int kQLandOffset = 4096 - kQFrameBytes;
constexpr int kDqIqLo = 1024;                 // entries of IQ_TABLE kept next to the rows
struct DqCtx {
    const DequantTables *D;                   // global memory: |q| >= kDqIqLo, PNS, band edges
    const float *iq_lo;                       // IQ_TABLE[0 .. kDqIqLo - 1]
    const float *sf;                          // SCALEFACTOR_TABLE padded to 512 entries (NaN from 428 on)
    uint32_t sfb_long4;                       // sfb_long[u + 64 j] in byte j
    uint32_t sfb_short1;                      // sfb_short[u & 31]
    float iq_lane;                            // IQ_TABLE[lane]: values below 32 -- nearly all of a real stream -- are
                                              // looked up with one warp shuffle instead of a bank-conflicted gather
};
// IQ_TABLE[a] for a < kDqIqLo (callers redo larger ones).  AACFB_DQ_SHFL: values below 32 through a warp
// shuffle of a per-lane copy instead of the shared-memory gather -- measured SLOWER on B200 (0.397 vs
// 0.345 ms on config 2, independent of the spread of the integers: the gather's bank conflicts are not
// what bounds this phase), so it is off.
#ifndef AACFB_DQ_SHFL
#define AACFB_DQ_SHFL 0
#endif
AACFB_HD float dq_iq(const DqCtx &dq, int a) {
#if defined(__CUDA_ARCH__) && AACFB_DQ_SHFL
    float t = __shfl_sync(0xffffffffu, dq.iq_lane, a & 31);
    if (a >= 32) t = dq.iq_lo[a & (kDqIqLo - 1)];
    return t;
#else
    return dq.iq_lo[a & (kDqIqLo - 1)];
#endif
}
AACFB_HD void dq_thread_consts(const DequantTables *D, int u, DqCtx &dq) {
    dq.sfb_long4 = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) dq.sfb_long4 |= (uint32_t)dq_ld(&D->sfb_long[u + 64 * j]) << (8 * j);
    dq.sfb_short1 = dq_ld(&D->sfb_short[u & 31]);
#if defined(__CUDA_ARCH__)
    dq.iq_lane = dq_ld(&D->iq[threadIdx.x & 31]);
#else
    dq.iq_lane = 0.f;
#endif
}

template <class Sync>
AACFB_HD void dequant_stage(int u, Sync &sync, float *stage, int nch, const FrameBits *fi, const DqCtx &dq) {
    float v[2][16];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (c < nch) {
            const uint8_t *rec = reinterpret_cast<const uint8_t *>(stage + c * kRowFloats) + kQLandOffset;
            const uint16_t *band = reinterpret_cast<const uint16_t *>(rec + 8);
            const bool is_short = fb_seq(fi[c]) == AACFB_EIGHT_SHORT_SEQUENCE;
            const int max_sfb = (int)(fi[c] >> 24);
            uint32_t code[4];
            if (!is_short) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int sfb = (int)((dq.sfb_long4 >> (8 * j)) & 0xffu);
                    code[j] = sfb < max_sfb ? (uint32_t)band[sfb < 119 ? sfb : 119] : (uint32_t)AACFB_BAND_ZERO;
                }
            } else {   // group j's window is 2 j + (u >> 5); its window group from the cumulative group lengths
                const uint2 gl = *reinterpret_cast<const uint2 *>(rec);
                const int sfb = (int)dq.sfb_short1;
                int gidx[4] = {-1, -1, -1, -1};
                int acc = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int len = (int)(((k < 4 ? gl.x : gl.y) >> (8 * (k & 3))) & 0xffu);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int w = 2 * j + (u >> 5);
                        if (gidx[j] < 0 && w < acc + len) gidx[j] = k;   // len == 0 never matches: w >= acc
                    }
                    acc += len;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int idx = gidx[j] * max_sfb + sfb;
                    const bool ok = gidx[j] >= 0 && sfb < max_sfb && idx < 120;
                    code[j] = ok ? (uint32_t)band[ok ? idx : 0] : (uint32_t)AACFB_BAND_ZERO;
                }
            }
            float s[4];
            uint2 raw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s[j] = dq.sf[code[j] & AACFB_BAND_INDEX_MASK];
                raw[j] = *reinterpret_cast<const uint2 *>(rec + 256 + 8 * (u + 64 * j));
            }
            bool any_noise = false;
            int big = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t kind = code[j] & AACFB_BAND_KIND_MASK;
                any_noise |= kind == AACFB_BAND_NOISE;
                // zero / intensity / noise bands: scale 0 and no sign -> +0 whatever q holds (ics.js:222-227)
                const uint32_t sign_mask = kind == AACFB_BAND_SPECTRAL ? 0x80000000u : 0u;
                const float sc = kind == AACFB_BAND_SPECTRAL ? s[j] : 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t word = k < 2 ? raw[j].x : raw[j].y;
                    const int q = (k & 1) ? ((int)word >> 16) : (int)(int16_t)(word & 0xffffu);
                    const int a = q < 0 ? -q : q;
                    big |= a;
                    // (buf[j] > 0) ? IQ[buf[j]] : -IQ[-buf[j]]  (ics.js:250-252): the table is >= +0, so the sign
                    // is OR-ed in; q - 1 is negative exactly when q <= 0 (q = 0 gives -0 like the reference)
                    const uint32_t t = f_bits(dq_iq(dq, a)) | ((uint32_t)(q - 1) & sign_mask);
                    v[c][4 * j + k] = f_mul(f_from_bits(t), sc);
                }
            }
            if (big >= kDqIqLo) {   // some |q| beyond the shared-memory part of IQ_TABLE (rare): redo those elements
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if ((code[j] & AACFB_BAND_KIND_MASK) != AACFB_BAND_SPECTRAL) continue;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t word = k < 2 ? raw[j].x : raw[j].y;
                        const int q = (k & 1) ? ((int)word >> 16) : (int)(int16_t)(word & 0xffffu);
                        const int a = q < 0 ? -q : q;
                        if (a < kDqIqLo) continue;
                        const float t = dq_ld(&dq.D->iq[a > 8191 ? 8191 : a]);   // outside IQ_TABLE: `undefined` -> NaN
                        v[c][4 * j + k] = f_mul(q > 0 ? t : -t, s[j]);
                    }
                }
            }
            if (any_noise) {   // perceptual noise substitution: the general (slow) path for those groups
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if ((code[j] & AACFB_BAND_KIND_MASK) != AACFB_BAND_NOISE) continue;
                    int16_t q4[4] = {0, 0, 0, 0};
                    float tmp[4];               // (not &v[..]: that would push the whole array into local memory)
                    dequant4(dq.D, rec, fi[c], u + 64 * j, q4, tmp);
                    v[c][4 * j] = tmp[0]; v[c][4 * j + 1] = tmp[1]; v[c][4 * j + 2] = tmp[2]; v[c][4 * j + 3] = tmp[3];
                }
            }
        }
    }
    sync.barrier();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (c < nch) {
            float4 *row = reinterpret_cast<float4 *>(stage + c * kRowFloats);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 t;
                t.x = v[c][4 * j]; t.y = v[c][4 * j + 1]; t.z = v[c][4 * j + 2]; t.w = v[c][4 * j + 3];
                row[u + 64 * j] = t;
            }
        }
    }
    sync.barrier();
}

// ------------------------------------------------------------ stereo tools
// processMS (decoder.js:379-404) and processIS (decoder.js:337-376) on the staged rows of a
// channel pair (chain 0 = left, chain 1 = right), driven by the host's per-4-coefficient op
// table (aacfb_stereo_ops, include/aacfb.h).  64 threads x 4 groups of 4 coefficients.
AACFB_HD void stereo_apply(int u, float *stage, const aacfb_stereo_ops *ops) {
    float4 *L = reinterpret_cast<float4 *>(stage), *R = reinterpret_cast<float4 *>(stage + kRowFloats);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int g = u + 64 * i;
        const int op = ops->op[g];
        if (op == AACFB_STEREO_NONE) continue;
        const float4 l = L[g];
        float4 r = R[g];
        if (op == AACFB_STEREO_MS) {  // t = l - r; l += r; r = t
            float4 m;
            m.x = f_add(l.x, r.x); m.y = f_add(l.y, r.y); m.z = f_add(l.z, r.z); m.w = f_add(l.w, r.w);
            r.x = f_sub(l.x, r.x); r.y = f_sub(l.y, r.y); r.z = f_sub(l.z, r.z); r.w = f_sub(l.w, r.w);
            L[g] = m;
        } else {                      // right = left * scale
            const float sc = ops->scale[(op - AACFB_STEREO_IS) & 127];
            r.x = f_mul(l.x, sc); r.y = f_mul(l.y, sc); r.z = f_mul(l.z, sc); r.w = f_mul(l.w, sc);
        }
        R[g] = r;
    }
}

// ------------------------------------------------------------------- TNS
// tns.js:105-177 with `tmp` -> `top` at :122.  Strictly serial per chain:
// each tap is a rounded f32 read-modify-write in the reference's order
// (i = 1..min(m,order), tns.js:158-161), so one thread runs one chain with
// the last ORD outputs (AR) / inputs (MA) in registers, and many chains run
// side by side.  Taps beyond min(m,order) multiply a zero history / a zero
// coefficient, which leaves the accumulator unchanged, so no per-tap
// predicate is needed.  x = unfiltered row (read only), y = filtered row.
constexpr int kTnsPrefetch = 8;  // float4s per block = one 128-byte line per thread, one block fetched ahead

// Four samples through the serial chain (history h, coefficients c in registers).
template <int ORD, bool AR>
AACFB_HD float4 tns_quad(float4 t, int inc, float (&h)[ORD], const float (&c)[ORD], bool nan_from, int m0) {
    float v[4];
    if (inc > 0) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else         { v[0] = t.w; v[1] = t.z; v[2] = t.y; v[3] = t.x; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float acc = v[j];
#pragma unroll
        for (int i = 0; i < ORD; ++i) acc = f_fma(AR ? -h[i] : h[i], c[i], acc);
        // MA branch, order 20: the reference's tmp has 20 slots, tmp[20] reads
        // undefined -> NaN once m >= 20 (tns.js:43,169)
        if (!AR && nan_from && m0 + j >= AACFB_TNS_MAX_ORDER) acc = NAN;
        const float push = AR ? acc : v[j];
#pragma unroll
        for (int i = ORD - 1; i > 0; --i) h[i] = h[i - 1];
        h[0] = push;
        v[j] = acc;
    }
    float4 o;
    if (inc > 0) { o.x = v[0]; o.y = v[1]; o.z = v[2]; o.w = v[3]; }
    else         { o.x = v[3]; o.y = v[2]; o.z = v[1]; o.w = v[0]; }
    return o;
}

// Not inlined on the device: each (ORD, AR) instantiation gets its own register allocation.
template <int ORD, bool AR>
AACFB_HD_NOINLINE void tns_run(const float *x, float *y, int start, int size, int inc, const float *lpc, int order) {
    float h[ORD], c[ORD];
#pragma unroll
    for (int i = 0; i < ORD; ++i) { h[i] = 0.f; c[i] = i < order ? lpc[i] : 0.f; }
    // Band edges are multiples of 4 (tables.js:34-124), so runs are whole float4s.  The chain
    // is latency-bound and its loads are address-independent: whole blocks of kTnsPrefetch
    // float4s (one 128-byte line) are fetched one block ahead of the arithmetic.
    const int first = inc > 0 ? start : start - 3;
    const int step = inc > 0 ? 4 : -4;
    const int n4 = size >> 2, nblk = n4 / kTnsPrefetch;
    const bool nan_from = order == AACFB_TNS_MAX_ORDER;
    float4 cur[kTnsPrefetch];
    if (nblk > 0) {
#pragma unroll
        for (int p = 0; p < kTnsPrefetch; ++p) cur[p] = *reinterpret_cast<const float4 *>(x + first + p * step);
    }
    for (int b = 0; b < nblk; ++b) {
        const int base = first + b * kTnsPrefetch * step;
        float4 nxt[kTnsPrefetch];
        if (b + 1 < nblk) {
#pragma unroll
            for (int p = 0; p < kTnsPrefetch; ++p)
                nxt[p] = *reinterpret_cast<const float4 *>(x + base + (kTnsPrefetch + p) * step);
        }
#pragma unroll
        for (int p = 0; p < kTnsPrefetch; ++p) {
            const float4 o = tns_quad<ORD, AR>(cur[p], inc, h, c, nan_from, 4 * (b * kTnsPrefetch + p));
            *reinterpret_cast<float4 *>(y + base + p * step) = o;
        }
        if (b + 1 < nblk) {
#pragma unroll
            for (int p = 0; p < kTnsPrefetch; ++p) cur[p] = nxt[p];
        }
    }
    for (int q = nblk * kTnsPrefetch; q < n4; ++q) {  // tail shorter than a block
        const float4 t = *reinterpret_cast<const float4 *>(x + first + q * step);
        *reinterpret_cast<float4 *>(y + first + q * step) = tns_quad<ORD, AR>(t, inc, h, c, nan_from, 4 * q);
    }
}

template <bool AR>
AACFB_HD void tns_dispatch(const float *x, float *y, int start, int size, int inc, const float *lpc, int order) {
    if (order <= 4) tns_run<4, AR>(x, y, start, size, inc, lpc, order);
    else if (order <= 8) tns_run<8, AR>(x, y, start, size, inc, lpc, order);
    else if (order <= 12) tns_run<12, AR>(x, y, start, size, inc, lpc, order);
    else if (order <= 16) tns_run<16, AR>(x, y, start, size, inc, lpc, order);
    else tns_run<20, AR>(x, y, start, size, inc, lpc, order);
}

// One filter of the block, located in coefficient space.
struct TnsFilter {
    int start, size, inc, order;
    const float *coef;
    bool valid;   // false once the block is exhausted
    bool active;  // has a non-empty run (tns.js:125,147)
};
// Iterates the filters of a block in (window, filter) order and maps each to
// its coefficient run: top/bottom stacking tns.js:113,121-122 (`tmp` read as
// `top`), band clamp :106,142-143, direction :149-152, window offset :154.
struct TnsWalker {
    const uint8_t *block;
    uint32_t bytes, pos;
    uint32_t nf[2];  // n_filt[0..7], fetched once (blocks are 4-byte aligned)
    const uint16_t *swb;
    int swb_count, window_count, mmm, w, f, bottom;
    AACFB_HD TnsWalker(FrameBits fi, const uint8_t *b, uint32_t n, int sample_index, const TnsBandTables &bt) {
        init(fi, b, n, bt.swb_long[sample_index], bt.swb_long_count[sample_index], bt.swb_short[sample_index],
             bt.swb_short_count[sample_index], bt.tns_max_bands[sample_index]);
    }
    // the same from one sample rate's tables (the kernel keeps them in shared memory)
    AACFB_HD TnsWalker(FrameBits fi, const uint8_t *b, uint32_t n, const uint16_t *swb_long, int n_long,
                       const uint16_t *swb_short, int n_short, int max_bands) {
        init(fi, b, n, swb_long, n_long, swb_short, n_short, max_bands);
    }
    AACFB_HD void init(FrameBits fi, const uint8_t *b, uint32_t n, const uint16_t *swb_long, int n_long,
                       const uint16_t *swb_short, int n_short, int max_bands /* tns.js:23: the long table, for short windows too */) {
        block = b; bytes = n; pos = 8; w = 0; f = 0;
        nf[0] = nf[1] = 0u;
        if (n >= 8) {
            nf[0] = reinterpret_cast<const uint32_t *>(b)[0];
            nf[1] = reinterpret_cast<const uint32_t *>(b)[1];
        }
        const bool is_short = fb_seq(fi) == AACFB_EIGHT_SHORT_SEQUENCE;
        swb = is_short ? swb_short : swb_long;
        swb_count = is_short ? n_short : n_long;
        window_count = is_short ? 8 : 1;
        const int max_sfb = (int)(fi >> 24);
        mmm = max_bands < max_sfb ? max_bands : max_sfb;
        bottom = swb_count;
    }
    AACFB_HD TnsFilter next() {
        TnsFilter r;
        r.valid = false; r.active = false; r.start = r.size = r.order = 0; r.inc = 1; r.coef = nullptr;
        if (bytes < 8) return r;
        while (w < 8 && f >= (int)((nf[w >> 2] >> (8 * (w & 3))) & 0xffu)) { ++w; f = 0; bottom = swb_count; }
        if (w >= 8 || pos + 4 > bytes) return r;
        const uint32_t hdr = *reinterpret_cast<const uint32_t *>(block + pos);  // aacfb_tns_filter
        const int length = hdr & 0xffu, order = (hdr >> 8) & 0xffu, direction = (hdr >> 16) & 0xffu;
        r.coef = reinterpret_cast<const float *>(block + pos + 4);
        pos += 4 + 4 * order;
        ++f;
        if (order > AACFB_TNS_MAX_ORDER || pos > bytes) return r;
        r.valid = true;
        r.order = order;
        if (w >= window_count) return r;  // tns.js:111 only visits w < windowCount
        const int top = bottom;
        bottom = top - length;
        if (bottom < 0) bottom = 0;
        if (order == 0) return r;
        const int lo = swb[bottom < mmm ? bottom : mmm], hi = swb[top < mmm ? top : mmm];
        if (hi - lo <= 0) return r;
        r.size = hi - lo;
        r.inc = direction ? -1 : 1;
        r.start = (direction ? hi - 1 : lo) + w * 128;
        r.active = true;
        return r;
    }
};

// One channel-frame: y = TNS(x).  Every float4 of the row is written exactly
// once: first the coefficients no filter touches are copied, then each filter
// (direct-form coefficients from the reflection coefficients, tns.js:128-140)
// runs over its band range.  Filters of one window cover disjoint, downward
// stacked ranges, so every run reads unfiltered input from x.
AACFB_HD void tns_apply(FrameBits fi, const uint8_t *block, uint32_t block_bytes, int sample_index, bool ar,
                        const TnsBandTables &bt, const float *x, float *y) {
    uint32_t covered[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) covered[i] = 0u;
    {
        TnsWalker wk(fi, block, block_bytes, sample_index, bt);
        for (TnsFilter ft = wk.next(); ft.valid; ft = wk.next()) {
            if (!ft.active) continue;
            const int lo = ft.inc > 0 ? ft.start : ft.start - ft.size + 1;
            for (int q = lo >> 2; q < (lo + ft.size) >> 2; ++q) covered[q >> 5] |= 1u << (q & 31);
        }
    }
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    float4 *y4 = reinterpret_cast<float4 *>(y);
    for (int g = 0; g < 8; ++g) {
        const uint32_t cv = covered[g];
        if (cv == 0xffffffffu) continue;
#pragma unroll 8
        for (int b = 0; b < 32; ++b)
            if (!((cv >> b) & 1u)) y4[32 * g + b] = x4[32 * g + b];
    }
    float lpc[AACFB_TNS_MAX_ORDER];
    TnsWalker wk(fi, block, block_bytes, sample_index, bt);
    for (TnsFilter ft = wk.next(); ft.valid; ft = wk.next()) {
        if (!ft.active) continue;
        for (int i = 0; i < ft.order; ++i) {
            const float r = -ft.coef[i];
            lpc[i] = r;
            for (int j = 0, len = (i + 1) >> 1; j < len; ++j) {
                const float fwd = lpc[j], bwd = lpc[i - 1 - j];
                lpc[j] = f_fma(r, bwd, fwd);
                lpc[i - 1 - j] = f_fma(r, fwd, bwd);
            }
        }
        if (ar) tns_dispatch<true>(x, y, ft.start, ft.size, ft.inc, lpc, ft.order);
        else tns_dispatch<false>(x, y, ft.start, ft.size, ft.inc, lpc, ft.order);
    }
}

}  // namespace aacfb
