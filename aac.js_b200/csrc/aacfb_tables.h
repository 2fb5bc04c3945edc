// aacfb_tables.h -- constant tables of the synthesis kernels.
//
// Values follow the reference exactly (same recurrences, same rounding):
//   FFT roots      reference src/fft.js:59-103 (f32-rounded rotation recurrence)
//   MDCT twiddles  reference src/mdct_tables.js:21-601 (closed form, 15 decimals)
//   sine / KBD     reference src/filter_bank.js:46-86
// but are stored in layouts chosen for the kernel's shared-memory access
// patterns (one conflict-free vector load per use), see DESIGN.md "Tables".
#pragma once
#include <stdint.h>

#if !defined(__CUDACC__) && !defined(__VECTOR_TYPES_H__)
struct alignas(8) float2 { float x, y; };
#endif

namespace aacfb {

// Index maps shared by the table builder and the kernels -------------------
// IMDCT output position pair owned by post-twiddled FFT bin k (N=2048):
// bin k feeds positions m(k) and 1023-m(k) of both halves (mdct.js:90-114).
static inline
#if defined(__CUDACC__)
__host__ __device__
#endif
int long_pos_of_bin(int k) { return k < 256 ? 512 + 2 * k : 2 * (k - 256); }

struct alignas(16) SynthTables {
    // radix-2 DIT twiddles roots512[k*m] regrouped per thread:
    //   twC[j][u]: pass C (stages i=64,128,256) of thread u
    //       j=0: roots[4u]   j=1: roots[2u]   j=2: roots[2(64+u)]   j=3+q: roots[64q+u]
    //   twB[b*7+j]: pass B (stages i=8,16,32) of a thread with low bits b
    //       j=0: roots[32b]  j=1: roots[16b]  j=2: roots[16(8+b)]   j=3+q: roots[8(8q+b)]
    float2 twC[7][64];
    float2 twB[64];          // 56 used
    float2 cs2048[512];      // (c_k, s_k) of MDCT_TABLE_2048 rounded to f32
    float2 wz[2][512];       // [shape][k] = (W[m(k)], W[1023-m(k)]) long windows
    // 64-point FFT (EIGHT_SHORT): twS[g*7+j] from roots64, same scheme as twB
    //       j=0: r64[4g]     j=1: r64[2g]     j=2: r64[2(8+g)]      j=3+q: r64[8q+g]
    float2 twS[64];          // 56 used
    float2 cs256[64];        // MDCT_TABLE_256 rounded to f32
    float2 wsp[2][64];       // [shape][k] = (W[pa(k)], W[pb(k)]): the two short-window values FFT bin k of a
                             // window needs (pa = 64+2k | 2(k-32), pb = 63-2k | 191-2k; short_products)
    float2 rootsA[4];        // roots512[64k], k=0..3  (pass A stage i=4)
    float2 roots64A[4];      // roots64[8k],  k=0..3
    // window-switching variants (global memory, read only by START/STOP frames)
    float2 fwz_stop[2][512];  // LONG_STOP first-half window:  0 | short asc | 1   (filter_bank.js:184-194)
    float2 swz_start[2][512]; // LONG_START second-half window: 1 | short desc | 0 (filter_bank.js:129-139),
                              // stored (at 1023-m, at m) like a wz entry that is read as a second half
};

// bytes of SynthTables that the kernel stages into shared memory (everything
// up to and including roots64A)
constexpr int kSmemTableBytes = (7 * 64 + 64 + 512 + 2 * 512 + 64 + 64) * 8 + 2 * 128 * 4 + 8 * 8;

struct HostTables {
    float roots512[512][2];   // (re, col 1)
    float roots64[64][2];
    double mdct2048[512][2];
    double mdct256[64][2];
    float sine1024[1024], kbd1024[1024], sine128[128], kbd128[128];
    SynthTables synth;
    // scalefactor band tables for TNS (reference src/tables.js:34-163, tns.js:65)
};

const HostTables &host_tables();   // built once, thread-safe

// Multiply every window table of S by `scale` (a power of two).  The kernels work on windows
// that carry the output scale of decoder.js:210: scaling a window by a power of two commutes
// with every rounding downstream, so (overlap + F*W) * 2^-15 becomes overlap' + F*(W * 2^-15)
// with the overlap kept in scaled units -- bit-identical, two multiplies per bin cheaper.
void scale_windows(SynthTables &S, float scale);

// Inverse quantisation (ICStream.decodeSpectralData, reference src/ics.js:203-266).  Lives in global
// memory and is read through the L1 (quantised values are small: the hot part of `iq` is a few lines).
struct alignas(16) DequantTables {
    float iq[8192];          // IQ_TABLE tables.js:181-191 = f32(i^(4/3)); [8191] = NaN (the reference reads `undefined`)
    float sf[512];           // SCALEFACTOR_TABLE tables.js:168-176 = f32(2^((i-200)/4)); [428..511] = NaN (`undefined`)
    float noise[32];         // PNS as shipped (ics.js:234-235): the generator's outputs as Float32Array elements;
                             // 0 from noise_len on (the state collapses)
    int noise_len, pad[3];
    uint16_t swb_long[52], swb_short[16];   // info.swbOffsets of the context's sample rate (tables.js:126-154)
    uint8_t sfb_long[256], sfb_short[32];   // scalefactor band that holds coefficients 4i .. 4i+3
};
void build_dequant_tables(int sample_index, DequantTables &D);

// TNS band tables (device + host share the same flat arrays)
struct TnsBandTables {
    uint16_t swb_long[12][52];
    uint16_t swb_short[12][16];
    uint8_t swb_long_count[12];
    uint8_t swb_short_count[12];
    uint8_t tns_max_bands[13];
};
const TnsBandTables &tns_band_tables();

}  // namespace aacfb
