// aacfb_tables.cpp -- builds the constant tables of the synthesis kernels.
// Product code: independent of oracle/ (tests compare the two bit for bit).
#include "aacfb_tables.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace aacfb {
namespace {

constexpr double kPi = 3.14159265358979323846;

// Rotation recurrences of the reference's root tables.  Every store into the
// reference's Float32Array rounds to float; its `var`s stay double.
// fft.js:82-103: all three columns are f32.
void make_roots_long(int len, float (*out)[2]) {
    const double cs = std::cos(2 * kPi / len), sn = std::sin(2 * kPi / len);
    float re = 1.f, im_neg = 0.f;  // columns 0 and 2
    out[0][0] = 1.f;
    out[0][1] = 0.f;
    for (int i = 1; i < len; ++i) {
        const float nre = (float)((double)re * cs + (double)im_neg * sn);
        const float nim = (float)((double)im_neg * cs - (double)re * sn);
        re = nre;
        im_neg = nim;
        out[i][0] = re;
        out[i][1] = -im_neg;
    }
}
// fft.js:59-80: the running imaginary part is a double, only the stored
// copies are f32.
void make_roots_short(int len, float (*out)[2]) {
    const double cs = std::cos(2 * kPi / len), sn = std::sin(2 * kPi / len);
    float re = 1.f;
    double im_neg = 0.0;
    out[0][0] = 1.f;
    out[0][1] = 0.f;
    for (int i = 1; i < len; ++i) {
        const float nre = (float)((double)re * cs + im_neg * sn);
        im_neg = im_neg * cs - (double)re * sn;  // uses the previous re
        re = nre;
        out[i][0] = re;
        out[i][1] = (float)(-im_neg);
    }
}

// mdct_tables.js holds sqrt(2/N)*(cos,sin)(2pi(k+1/8)/N) as 15-decimal
// literals; a JS literal denotes the nearest double.
double as_printed(double v) {
    char txt[48];
    std::snprintf(txt, sizeof txt, "%.15f", v);
    return std::strtod(txt, nullptr);
}
void make_mdct(int n, double (*out)[2]) {
    const double g = std::sqrt(2.0 / n);
    for (int k = 0; k < n / 4; ++k) {
        const double ph = 2 * kPi * (k + 0.125) / n;
        out[k][0] = as_printed(g * std::cos(ph));
        out[k][1] = as_printed(g * std::sin(ph));
    }
}

// filter_bank.js:46-52
void make_sine(int len, float *w) {
    for (int i = 0; i < len; ++i) w[i] = (float)std::sin((i + 0.5) * (kPi / (2.0 * len)));
}
// filter_bank.js:54-79: 50-term Horner series for I0, running sum kept in
// double but snapshotted into an f32 array, then sqrt(f[n]/(sum+1)).
void make_kbd(double alpha, int len, float *w) {
    const double step = kPi / len, a2 = (alpha * step) * (alpha * step);
    float *cum = (float *)std::malloc(sizeof(float) * len);
    double total = 0;
    for (int n = 0; n < len; ++n) {
        const double t = (double)n * (len - n) * a2;
        double i0 = 1;
        for (int j = 50; j > 0; --j) i0 = i0 * t / (j * j) + 1;
        total += i0;
        cum[n] = (float)total;
    }
    total += 1;
    for (int n = 0; n < len; ++n) w[n] = (float)std::sqrt((double)cum[n] / total);
    std::free(cum);
}

float2 f2(float a, float b) {
    float2 r;
    r.x = a;
    r.y = b;
    return r;
}

void build(HostTables &T) {
    std::memset(&T, 0, sizeof T);
    make_roots_long(512, T.roots512);
    make_roots_short(64, T.roots64);
    make_mdct(2048, T.mdct2048);
    make_mdct(256, T.mdct256);
    make_sine(1024, T.sine1024);
    make_sine(128, T.sine128);
    make_kbd(4, 1024, T.kbd1024);
    make_kbd(6, 128, T.kbd128);

    SynthTables &S = T.synth;
    auto r512 = [&](int i) { return f2(T.roots512[i][0], T.roots512[i][1]); };
    auto r64 = [&](int i) { return f2(T.roots64[i][0], T.roots64[i][1]); };
    for (int u = 0; u < 64; ++u) {
        S.twC[0][u] = r512(4 * u);
        S.twC[1][u] = r512(2 * u);
        S.twC[2][u] = r512(2 * (64 + u));
        for (int q = 0; q < 4; ++q) S.twC[3 + q][u] = r512(64 * q + u);
    }
    for (int b = 0; b < 8; ++b) {
        S.twB[b * 7 + 0] = r512(32 * b);
        S.twB[b * 7 + 1] = r512(16 * b);
        S.twB[b * 7 + 2] = r512(16 * (8 + b));
        for (int q = 0; q < 4; ++q) S.twB[b * 7 + 3 + q] = r512(8 * (8 * q + b));
        S.twS[b * 7 + 0] = r64(4 * b);
        S.twS[b * 7 + 1] = r64(2 * b);
        S.twS[b * 7 + 2] = r64(2 * (8 + b));
        for (int q = 0; q < 4; ++q) S.twS[b * 7 + 3 + q] = r64(8 * q + b);
    }
    for (int k = 0; k < 512; ++k) S.cs2048[k] = f2((float)T.mdct2048[k][0], (float)T.mdct2048[k][1]);
    for (int k = 0; k < 64; ++k) S.cs256[k] = f2((float)T.mdct256[k][0], (float)T.mdct256[k][1]);
    for (int k = 0; k < 4; ++k) {
        S.rootsA[k] = r512(64 * k);
        S.roots64A[k] = r64(8 * k);
    }
    const float *wl[2] = {T.sine1024, T.kbd1024};
    const float *ws[2] = {T.sine128, T.kbd128};
    for (int sh = 0; sh < 2; ++sh) {
        for (int k = 0; k < 64; ++k) {
            const int pa = k < 32 ? 64 + 2 * k : 2 * (k - 32), pb = k < 32 ? 63 - 2 * k : 191 - 2 * k;
            S.wsp[sh][k] = f2(ws[sh][pa], ws[sh][pb]);
        }
        // effective first-half window of LONG_STOP and second-half window of
        // LONG_START, as functions of the output position n in [0,1024)
        float stop_first[1024], start_second[1024];
        for (int n = 0; n < 1024; ++n) {
            stop_first[n] = n < 448 ? 0.f : (n < 576 ? ws[sh][n - 448] : 1.f);
            start_second[n] = n < 448 ? 1.f : (n < 576 ? ws[sh][127 - (n - 448)] : 0.f);
        }
        for (int k = 0; k < 512; ++k) {
            const int m = long_pos_of_bin(k), mm = 1023 - m;
            S.wz[sh][k] = f2(wl[sh][m], wl[sh][mm]);
            S.fwz_stop[sh][k] = f2(stop_first[m], stop_first[mm]);
            S.swz_start[sh][k] = f2(start_second[mm], start_second[m]);  // like wz read as a second half: (at 1023-m, at m)
        }
    }
}

// Scalefactor-window-band boundaries (ISO/IEC 14496-3 tables 4.110-4.128) as
// the reference groups them per sampleIndex (src/tables.js:34-163).  Stored
// as band widths; the offsets the reference lists are their running sums.
struct WidthRun { uint8_t width, count; };
#define RUNS(...) { __VA_ARGS__, {0, 0} }
const WidthRun L96[] = RUNS({4,14},{8,5},{12,5},{16,2},{24,1},{28,1},{36,1},{44,1},{64,11});
const WidthRun L64[] = RUNS({4,14},{8,4},{12,3},{16,3},{20,1},{24,2},{28,1},{36,1},{40,18});
const WidthRun L48[] = RUNS({4,10},{8,7},{12,4},{16,2},{20,2},{24,2},{28,2},{32,19},{96,1});
const WidthRun L32[] = RUNS({4,10},{8,7},{12,4},{16,2},{20,2},{24,2},{28,2},{32,22});
const WidthRun L24[] = RUNS({4,11},{8,10},{12,4},{16,3},{20,2},{24,2},{28,2},{32,1},{36,2},{40,1},{44,1},{48,1},{52,2},{64,5});
const WidthRun L16[] = RUNS({8,11},{12,9},{16,4},{20,3},{24,2},{28,2},{32,1},{36,1},{40,2},{44,1},{48,1},{52,1},{56,1},{60,1},{64,3});
const WidthRun L8[]  = RUNS({12,13},{16,7},{20,4},{24,3},{28,2},{32,1},{36,2},{40,1},{44,1},{48,1},{52,1},{56,1},{60,1},{64,1},{80,1});
const WidthRun S96[] = RUNS({4,6},{8,3},{16,1},{28,1},{36,1});
const WidthRun S48[] = RUNS({4,5},{8,3},{12,3},{16,3});
const WidthRun S24[] = RUNS({4,7},{8,3},{12,2},{16,2},{20,1});
const WidthRun S16[] = RUNS({4,8},{8,2},{12,2},{16,1},{20,2});
const WidthRun S8[]  = RUNS({4,7},{8,4},{12,1},{16,1},{20,2});

int expand(const WidthRun *r, uint16_t *dst, int cap) {
    int n = 0, pos = 0;
    dst[n++] = 0;
    for (; r->count; ++r)
        for (int i = 0; i < r->count && n < cap; ++i) dst[n++] = (uint16_t)(pos += r->width);
    for (int i = n; i < cap; ++i) dst[i] = (uint16_t)pos;  // clamp reads past the end
    return n - 1;  // number of bands
}

void build_bands(TnsBandTables &B) {
    std::memset(&B, 0, sizeof B);
    const WidthRun *lng[12] = {L96, L96, L64, L48, L48, L32, L24, L24, L16, L16, L16, L8};
    const WidthRun *sht[12] = {S96, S96, S96, S48, S48, S48, S24, S24, S16, S16, S16, S8};
    // the band counts equal SWB_LONG/SHORT_WINDOW_COUNT, tables.js:157-163
    for (int i = 0; i < 12; ++i) {
        B.swb_long_count[i] = (uint8_t)expand(lng[i], B.swb_long[i], 52);
        B.swb_short_count[i] = (uint8_t)expand(sht[i], B.swb_short[i], 16);
    }
    const uint8_t mb[13] = {31, 31, 34, 40, 42, 51, 46, 46, 42, 42, 42, 39, 39};  // tns.js:65
    std::memcpy(B.tns_max_bands, mb, sizeof mb);
}

}  // namespace

const HostTables &host_tables() {
    static HostTables *T = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        T = new HostTables;
        build(*T);
    });
    return *T;
}

void scale_windows(SynthTables &S, float scale) {
    for (int sh = 0; sh < 2; ++sh) {
        for (int k = 0; k < 512; ++k) {
            S.wz[sh][k].x *= scale; S.wz[sh][k].y *= scale;
            S.fwz_stop[sh][k].x *= scale; S.fwz_stop[sh][k].y *= scale;
            S.swz_start[sh][k].x *= scale; S.swz_start[sh][k].y *= scale;
        }
        for (int k = 0; k < 64; ++k) { S.wsp[sh][k].x *= scale; S.wsp[sh][k].y *= scale; }
    }
}

// ics.js:203-266 needs IQ_TABLE / SCALEFACTOR_TABLE (tables.js:168-191), the scalefactor bands
// of the context's sample rate and -- for perceptual noise substitution as shipped -- what the
// generator of ics.js:234 produces: randomState = (randomState * (1664525 + 1013904223))|0 from
// 0x1F2E3D4C, the product formed in double precision and wrapped by ToInt32.
void build_dequant_tables(int sample_index, DequantTables &D) {
    std::memset(&D, 0, sizeof D);
    const float nan = std::nanf("");
    const double four_thirds = 4.0 / 3.0;
    for (int i = 0; i < 8191; ++i) D.iq[i] = (float)std::pow((double)i, four_thirds);
    D.iq[8191] = nan;
    for (int i = 0; i < 512; ++i) D.sf[i] = i < 428 ? (float)std::pow(2.0, (i - 200) / 4.0) : nan;
    double state = (double)0x1F2E3D4C;
    D.noise_len = 0;
    for (int k = 0; k < 32; ++k) {
        const double prod = state * (double)(1664525 + 1013904223);
        // ToInt32: truncate (already integral), modulo 2^32, reinterpret as signed
        double m = std::fmod(prod, 4294967296.0);
        if (m < 0) m += 4294967296.0;
        if (m >= 2147483648.0) m -= 4294967296.0;
        state = m + 0.0;                    // ToInt32 never yields -0
        D.noise[k] = (float)state;          // data[off + k] = this.randomState (Float32Array store)
        if (state != 0.0) D.noise_len = k + 1;
    }
    const TnsBandTables &B = tns_band_tables();
    std::memcpy(D.swb_long, B.swb_long[sample_index], sizeof D.swb_long);
    std::memcpy(D.swb_short, B.swb_short[sample_index], sizeof D.swb_short);
    const int nl = B.swb_long_count[sample_index], ns = B.swb_short_count[sample_index];
    for (int i = 0; i < 256; ++i) {
        int b = 0;
        while (b + 1 < nl && D.swb_long[b + 1] <= 4 * i) ++b;
        D.sfb_long[i] = (uint8_t)b;
    }
    for (int i = 0; i < 32; ++i) {
        int b = 0;
        while (b + 1 < ns && D.swb_short[b + 1] <= 4 * i) ++b;
        D.sfb_short[i] = (uint8_t)b;
    }
}

const TnsBandTables &tns_band_tables() {
    static TnsBandTables *B = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        B = new TnsBandTables;
        build_bands(*B);
    });
    return *B;
}

}  // namespace aacfb
