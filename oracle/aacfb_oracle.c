/*
 * aacfb_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the filterbank-synthesis path of audiocogs/aac.js
 * @2d9bd01 under the JavaScript rounding model: every JS `var` is a double,
 * every element of a Float32Array is rounded to float at the store.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file's shared object.  The shipped library
 * (aac.js_b200/csrc) never links or calls it.
 *
 * PINNING.  The reference has no tests, fixtures or golden vectors for this
 * path and no JavaScript engine exists in this image, so
 * it cannot be run by a stock engine here.  Instead its own, unmodified source
 * files for the path are executed by tools/jsmini.py (a generic ES5-subset
 * interpreter written for this repository): tools/js_reference.py produced
 * tests/golden/jsref_*.npz, and tests/test_oracle_pin.py requires this oracle
 * to reproduce them BIT FOR BIT (all four window sequences, both window
 * shapes, TNS as shipped = identity, TNS with the one-token fix in both
 * branches, the decoder's interleave) -- live on fresh random frames too when
 * /root/reference is present.  Caveat, stated once: the interpreter, not V8,
 * ran the code; Math.sin/cos/sqrt come from libm (table generation only).
 * Reference-independent identities (direct O(N^2) IMDCT, TDAC, window power
 * complementarity) are in tests/test_oracle_identities.py.
 *
 * Build:  make -C oracle      (gcc -O2 -ffp-contract=off, never -ffast-math)
 *
 * Each function cites the reference file:line it restates.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/aacfb.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ tables */

/* src/tables.js:34-163 -- scalefactor-window-band offsets (ISO 14496-3
 * tables 4.110-4.128), indexed by sampleIndex as the reference indexes them. */
static const uint16_t SWB1024_96[] = {0,4,8,12,16,20,24,28,32,36,40,44,48,52,56,64,72,80,88,96,108,120,132,144,156,172,188,212,240,276,320,384,448,512,576,640,704,768,832,896,960,1024};
static const uint16_t SWB1024_64[] = {0,4,8,12,16,20,24,28,32,36,40,44,48,52,56,64,72,80,88,100,112,124,140,156,172,192,216,240,268,304,344,384,424,464,504,544,584,624,664,704,744,784,824,864,904,944,984,1024};
static const uint16_t SWB1024_48[] = {0,4,8,12,16,20,24,28,32,36,40,48,56,64,72,80,88,96,108,120,132,144,160,176,196,216,240,264,292,320,352,384,416,448,480,512,544,576,608,640,672,704,736,768,800,832,864,896,928,1024};
static const uint16_t SWB1024_32[] = {0,4,8,12,16,20,24,28,32,36,40,48,56,64,72,80,88,96,108,120,132,144,160,176,196,216,240,264,292,320,352,384,416,448,480,512,544,576,608,640,672,704,736,768,800,832,864,896,928,960,992,1024};
static const uint16_t SWB1024_24[] = {0,4,8,12,16,20,24,28,32,36,40,44,52,60,68,76,84,92,100,108,116,124,136,148,160,172,188,204,220,240,260,284,308,336,364,396,432,468,508,552,600,652,704,768,832,896,960,1024};
static const uint16_t SWB1024_16[] = {0,8,16,24,32,40,48,56,64,72,80,88,100,112,124,136,148,160,172,184,196,212,228,244,260,280,300,320,344,368,396,424,456,492,532,572,616,664,716,772,832,896,960,1024};
static const uint16_t SWB1024_8[]  = {0,12,24,36,48,60,72,84,96,108,120,132,144,156,172,188,204,220,236,252,268,288,308,328,348,372,396,420,448,476,508,544,580,620,664,712,764,820,880,944,1024};
static const uint16_t SWB128_96[] = {0,4,8,12,16,20,24,32,40,48,64,92,128};
static const uint16_t SWB128_64[] = {0,4,8,12,16,20,24,32,40,48,64,92,128};
static const uint16_t SWB128_48[] = {0,4,8,12,16,20,28,36,44,56,68,80,96,112,128};
static const uint16_t SWB128_24[] = {0,4,8,12,16,20,24,28,36,44,52,64,76,92,108,128};
static const uint16_t SWB128_16[] = {0,4,8,12,16,20,24,28,32,40,48,60,72,88,108,128};
static const uint16_t SWB128_8[]  = {0,4,8,12,16,20,24,28,36,44,52,60,72,88,108,128};

#define NLEN(a) ((int)(sizeof(a) / sizeof((a)[0])))
typedef struct { const uint16_t *off; int n; } swb_t;
#define SW(a) { a, NLEN(a) }
/* tables.js:126-154: both arrays have 12 entries (sampleIndex 0..11) */
static const swb_t SWB_OFFSET_1024[12] = { SW(SWB1024_96), SW(SWB1024_96), SW(SWB1024_64), SW(SWB1024_48), SW(SWB1024_48), SW(SWB1024_32), SW(SWB1024_24), SW(SWB1024_24), SW(SWB1024_16), SW(SWB1024_16), SW(SWB1024_16), SW(SWB1024_8) };
static const swb_t SWB_OFFSET_128[12]  = { SW(SWB128_96), SW(SWB128_96), SW(SWB128_64), SW(SWB128_48), SW(SWB128_48), SW(SWB128_48), SW(SWB128_24), SW(SWB128_24), SW(SWB128_16), SW(SWB128_16), SW(SWB128_16), SW(SWB128_8) };
static const uint8_t SWB_SHORT_WINDOW_COUNT[12] = {12,12,12,14,14,14,15,15,15,15,15,15}; /* tables.js:157-159 */
static const uint8_t SWB_LONG_WINDOW_COUNT[12]  = {41,41,47,49,49,51,47,47,43,43,43,40}; /* tables.js:161-163 */
static const int TNS_MAX_BANDS_1024[13] = {31,31,34,40,42,51,46,46,42,42,42,39,39};     /* tns.js:65 */

static float  g_roots_long[512][3];   /* fft.js:82-103  generateFFTTableLong(512)  */
static float  g_roots_short[64][2];   /* fft.js:59-80   generateFFTTableShort(64)  */
static double g_mdct_2048[512][2];    /* mdct_tables.js:21-534                     */
static double g_mdct_256[64][2];      /* mdct_tables.js:536-601                    */
static float  g_sine_1024[1024], g_sine_128[128], g_kbd_1024[1024], g_kbd_128[128];
static const float *g_long_windows[2], *g_short_windows[2]; /* filter_bank.js:85-86 */
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

/* fft.js:59-80 */
static void gen_fft_table_short(int len, float (*f)[2]) {
    double t = 2 * M_PI / len, cosT = cos(t), sinT = sin(t);
    f[0][0] = 1; f[0][1] = 0;
    double lastImag = 0;
    for (int i = 1; i < len; i++) {
        f[i][0] = (float)((double)f[i - 1][0] * cosT + lastImag * sinT);
        lastImag = lastImag * cosT - (double)f[i - 1][0] * sinT;
        f[i][1] = (float)(-lastImag);
    }
}

/* fft.js:82-103 */
static void gen_fft_table_long(int len, float (*f)[3]) {
    double t = 2 * M_PI / len, cosT = cos(t), sinT = sin(t);
    f[0][0] = 1; f[0][1] = 0; f[0][2] = 0;
    for (int i = 1; i < len; i++) {
        f[i][0] = (float)((double)f[i - 1][0] * cosT + (double)f[i - 1][2] * sinT);
        f[i][2] = (float)((double)f[i - 1][2] * cosT - (double)f[i - 1][0] * sinT);
        f[i][1] = -f[i][2];
    }
}

/* mdct_tables.js prints sqrt(2/N)*(cos,sin)(2*pi*(k+1/8)/N) to 15 decimals;
 * the JS value is the double nearest that decimal literal.  Reproduce by
 * formatting the closed form to 15 decimals and parsing it back
 * (tests/test_tables.py checks every entry against the reference file). */
static double round15(double v) {
    char b[64];
    snprintf(b, sizeof b, "%.15f", v);
    return strtod(b, NULL);
}
static void gen_mdct_table(int N, double (*tab)[2]) {
    double sc = sqrt(2.0 / N);
    for (int k = 0; k < N / 4; k++) {
        double a = 2 * M_PI * (k + 0.125) / N;
        tab[k][0] = round15(sc * cos(a));
        tab[k][1] = round15(sc * sin(a));
    }
}

/* filter_bank.js:46-52 */
static void gen_sine_window(int len, float *d) {
    for (int i = 0; i < len; i++) d[i] = (float)sin((i + 0.5) * (M_PI / (2.0 * len)));
}

/* filter_bank.js:54-79 */
static void gen_kbd_window(double alpha, int len, float *out) {
    double PIN = M_PI / len, sum = 0;
    float *f = (float *)malloc(sizeof(float) * len);
    double alpha2 = (alpha * PIN) * (alpha * PIN);
    for (int n = 0; n < len; n++) {
        double tmp = (double)n * (len - n) * alpha2, bessel = 1;
        for (int j = 50; j > 0; j--) bessel = bessel * tmp / (j * j) + 1;
        sum += bessel;
        f[n] = (float)sum;
    }
    sum++;
    for (int n = 0; n < len; n++) out[n] = (float)sqrt((double)f[n] / sum);
    free(f);
}

static void init_tables(void) {
    gen_fft_table_long(512, g_roots_long);
    gen_fft_table_short(64, g_roots_short);
    gen_mdct_table(2048, g_mdct_2048);
    gen_mdct_table(256, g_mdct_256);
    gen_sine_window(1024, g_sine_1024);   /* filter_bank.js:81-84 */
    gen_sine_window(128, g_sine_128);
    gen_kbd_window(4, 1024, g_kbd_1024);
    gen_kbd_window(6, 128, g_kbd_128);
    g_long_windows[0] = g_sine_1024;  g_long_windows[1] = g_kbd_1024;
    g_short_windows[0] = g_sine_128;  g_short_windows[1] = g_kbd_128;
}

/* -------------------------------------------------------------------- FFT */

/* fft.js:105-192, forward=false only (imOffset=1, scale=1).  `in` is the
 * AoS buffer `input[i][0..1]`; `rev` is this.rev. */
static void fft_process_inverse(int length, float (*in)[2], float (*rev)[2]) {
    /* bit-reversal copy, fft.js:112-125 */
    int ii = 0;
    for (int i = 0; i < length; i++) {
        rev[i][0] = in[ii][0];
        rev[i][1] = in[ii][1];
        int k = length >> 1;
        while (ii >= k && k > 0) { ii -= k; k >>= 1; }
        ii += k;
    }
    for (int i = 0; i < length; i++) { in[i][0] = rev[i][0]; in[i][1] = rev[i][1]; }

    /* bottom base-4 round, fft.js:139-170 (a,b,c,d,e1,e2 are Float32Arrays) */
    for (int i = 0; i < length; i += 4) {
        float a0 = (float)((double)in[i][0] + (double)in[i + 1][0]);
        float a1 = (float)((double)in[i][1] + (double)in[i + 1][1]);
        float b0 = (float)((double)in[i + 2][0] + (double)in[i + 3][0]);
        float b1 = (float)((double)in[i + 2][1] + (double)in[i + 3][1]);
        float c0 = (float)((double)in[i][0] - (double)in[i + 1][0]);
        float c1 = (float)((double)in[i][1] - (double)in[i + 1][1]);
        float d0 = (float)((double)in[i + 2][0] - (double)in[i + 3][0]);
        float d1 = (float)((double)in[i + 2][1] - (double)in[i + 3][1]);
        in[i][0] = (float)((double)a0 + (double)b0);
        in[i][1] = (float)((double)a1 + (double)b1);
        in[i + 2][0] = (float)((double)a0 - (double)b0);
        in[i + 2][1] = (float)((double)a1 - (double)b1);
        float e10 = (float)((double)c0 - (double)d1);
        float e11 = (float)((double)c1 + (double)d0);
        float e20 = (float)((double)c0 + (double)d1);
        float e21 = (float)((double)c1 - (double)d0);
        /* inverse branch, fft.js:163-168 */
        in[i + 1][0] = e10; in[i + 1][1] = e11;
        in[i + 3][0] = e20; in[i + 3][1] = e21;
    }

    /* iterations from bottom to top, fft.js:172-191 */
    for (int i = 4; i < length; i <<= 1) {
        int shift = i << 1, m = length / shift;
        for (int j = 0; j < length; j += shift) {
            for (int k = 0; k < i; k++) {
                int km = k * m;
                double rootRe, rootIm;
                if (length == 512) { rootRe = g_roots_long[km][0]; rootIm = g_roots_long[km][1]; }
                else               { rootRe = g_roots_short[km][0]; rootIm = g_roots_short[km][1]; }
                double xr = in[i + j + k][0], xi = in[i + j + k][1];
                double zRe = xr * rootRe - xi * rootIm;
                double zIm = xr * rootIm + xi * rootRe;
                double ur = in[j + k][0], ui = in[j + k][1];
                in[i + j + k][0] = (float)((ur - zRe) * 1);
                in[i + j + k][1] = (float)((ui - zIm) * 1);
                in[j + k][0] = (float)((ur + zRe) * 1);
                in[j + k][1] = (float)((ui + zIm) * 1);
            }
        }
    }
}

/* ------------------------------------------------------------------- MDCT */

typedef struct {
    float buf[512][2];  /* MDCT.buf, mdct.js:54-57 */
    float rev[512][2];  /* FFT.rev,  fft.js:45-48  */
} mdct_scratch;

/* mdct.js:62-115.  N = 2048 or 256. */
static void mdct_process(int N, const float *input, int inOffset, float *output, int outOffset,
                         mdct_scratch *sc) {
    int N2 = N >> 1, N4 = N >> 2, N8 = N >> 3;
    const double (*sincos)[2] = (N == 2048) ? g_mdct_2048 : g_mdct_256;
    float (*buf)[2] = sc->buf;

    /* pre-IFFT complex multiplication, mdct.js:73-76 */
    for (int k = 0; k < N4; k++) {
        double x0 = input[inOffset + 2 * k], x1 = input[inOffset + N2 - 1 - 2 * k];
        buf[k][1] = (float)((x0 * sincos[k][0]) + (x1 * sincos[k][1]));
        buf[k][0] = (float)((x1 * sincos[k][0]) - (x0 * sincos[k][1]));
    }

    /* complex IFFT, non-scaling, mdct.js:79 */
    fft_process_inverse(N4, buf, sc->rev);

    /* post-IFFT complex multiplication, mdct.js:82-87 (tmp is a Float32Array) */
    for (int k = 0; k < N4; k++) {
        double t0 = buf[k][0], t1 = buf[k][1];
        buf[k][1] = (float)((t1 * sincos[k][0]) + (t0 * sincos[k][1]));
        buf[k][0] = (float)((t0 * sincos[k][0]) - (t1 * sincos[k][1]));
    }

    /* reordering, mdct.js:90-114 */
    float *o = output + outOffset;
    for (int k = 0; k < N8; k += 2) {
        o[2 * k] = buf[N8 + k][1];
        o[2 + 2 * k] = buf[N8 + 1 + k][1];
        o[1 + 2 * k] = -buf[N8 - 1 - k][0];
        o[3 + 2 * k] = -buf[N8 - 2 - k][0];

        o[N4 + 2 * k] = buf[k][0];
        o[N4 + 2 + 2 * k] = buf[1 + k][0];
        o[N4 + 1 + 2 * k] = -buf[N4 - 1 - k][1];
        o[N4 + 3 + 2 * k] = -buf[N4 - 2 - k][1];

        o[N2 + 2 * k] = buf[N8 + k][0];
        o[N2 + 2 + 2 * k] = buf[N8 + 1 + k][0];
        o[N2 + 1 + 2 * k] = -buf[N8 - 1 - k][1];
        o[N2 + 3 + 2 * k] = -buf[N8 - 2 - k][1];

        o[N2 + N4 + 2 * k] = -buf[k][1];
        o[N2 + N4 + 2 + 2 * k] = -buf[1 + k][1];
        o[N2 + N4 + 1 + 2 * k] = buf[N4 - 1 - k][0];
        o[N2 + N4 + 3 + 2 * k] = buf[N4 - 2 - k][0];
    }
}

/* ------------------------------------------------------------- FilterBank */

/* filter_bank.js:88-204.  `overlap` = this.overlaps[channel]; `buf` =
 * this.buf (2048 floats).  Unknown window_sequence: the switch has no default
 * (filter_bank.js:104-203) so nothing happens and `output` keeps the zeros it
 * was allocated with (decoder.js:230). */
static void filterbank_process(const aacfb_frame_info *info, const float *input, float *output,
                               float *overlap, float *buf, mdct_scratch *sc) {
    const int length = 1024, shortLen = 128, mid = 448, trans = 64; /* filter_bank.js:29-33 */
    const float *longWindows = g_long_windows[info->shape_cur & 1];
    const float *shortWindows = g_short_windows[info->shape_cur & 1];
    const float *longWindowsPrev = g_long_windows[info->shape_prev & 1];
    const float *shortWindowsPrev = g_short_windows[info->shape_prev & 1];
#define D(x) ((double)(x))
    switch (info->window_sequence) {
    case AACFB_ONLY_LONG_SEQUENCE: /* filter_bank.js:105-118 */
        mdct_process(2048, input, 0, buf, 0, sc);
        for (int i = 0; i < length; i++) output[i] = (float)(D(overlap[i]) + (D(buf[i]) * D(longWindowsPrev[i])));
        for (int i = 0; i < length; i++) overlap[i] = (float)(D(buf[length + i]) * D(longWindows[length - 1 - i]));
        break;

    case AACFB_LONG_START_SEQUENCE: /* filter_bank.js:120-141 */
        mdct_process(2048, input, 0, buf, 0, sc);
        for (int i = 0; i < length; i++) output[i] = (float)(D(overlap[i]) + (D(buf[i]) * D(longWindowsPrev[i])));
        for (int i = 0; i < mid; i++) overlap[i] = buf[length + i];
        for (int i = 0; i < shortLen; i++) overlap[mid + i] = (float)(D(buf[length + mid + i]) * D(shortWindows[shortLen - i - 1]));
        for (int i = 0; i < mid; i++) overlap[mid + shortLen + i] = 0;
        break;

    case AACFB_EIGHT_SHORT_SEQUENCE: /* filter_bank.js:143-178 */
        for (int i = 0; i < 8; i++) mdct_process(256, input, i * shortLen, buf, 2 * i * shortLen, sc);
        for (int i = 0; i < mid; i++) output[i] = overlap[i];
        for (int i = 0; i < shortLen; i++) {
            output[mid + i] = (float)(D(overlap[mid + i]) + D(buf[i]) * D(shortWindowsPrev[i]));
            output[mid + 1 * shortLen + i] = (float)(D(overlap[mid + shortLen * 1 + i]) + (D(buf[shortLen * 1 + i]) * D(shortWindows[shortLen - 1 - i])) + (D(buf[shortLen * 2 + i]) * D(shortWindows[i])));
            output[mid + 2 * shortLen + i] = (float)(D(overlap[mid + shortLen * 2 + i]) + (D(buf[shortLen * 3 + i]) * D(shortWindows[shortLen - 1 - i])) + (D(buf[shortLen * 4 + i]) * D(shortWindows[i])));
            output[mid + 3 * shortLen + i] = (float)(D(overlap[mid + shortLen * 3 + i]) + (D(buf[shortLen * 5 + i]) * D(shortWindows[shortLen - 1 - i])) + (D(buf[shortLen * 6 + i]) * D(shortWindows[i])));
            if (i < trans)
                output[mid + 4 * shortLen + i] = (float)(D(overlap[mid + shortLen * 4 + i]) + (D(buf[shortLen * 7 + i]) * D(shortWindows[shortLen - 1 - i])) + (D(buf[shortLen * 8 + i]) * D(shortWindows[i])));
        }
        for (int i = 0; i < shortLen; i++) {
            if (i >= trans)
                overlap[mid + 4 * shortLen + i - length] = (float)((D(buf[shortLen * 7 + i]) * D(shortWindows[shortLen - 1 - i])) + (D(buf[shortLen * 8 + i]) * D(shortWindows[i])));
            overlap[mid + 5 * shortLen + i - length] = (float)((D(buf[shortLen * 9 + i]) * D(shortWindows[shortLen - 1 - i])) + (D(buf[shortLen * 10 + i]) * D(shortWindows[i])));
            overlap[mid + 6 * shortLen + i - length] = (float)((D(buf[shortLen * 11 + i]) * D(shortWindows[shortLen - 1 - i])) + (D(buf[shortLen * 12 + i]) * D(shortWindows[i])));
            overlap[mid + 7 * shortLen + i - length] = (float)((D(buf[shortLen * 13 + i]) * D(shortWindows[shortLen - 1 - i])) + (D(buf[shortLen * 14 + i]) * D(shortWindows[i])));
            overlap[mid + 8 * shortLen + i - length] = (float)((D(buf[shortLen * 15 + i]) * D(shortWindows[shortLen - 1 - i])));
        }
        for (int i = 0; i < mid; i++) overlap[mid + shortLen + i] = 0;
        break;

    case AACFB_LONG_STOP_SEQUENCE: /* filter_bank.js:180-202 */
        mdct_process(2048, input, 0, buf, 0, sc);
        for (int i = 0; i < mid; i++) output[i] = overlap[i];
        for (int i = 0; i < shortLen; i++) output[mid + i] = (float)(D(overlap[mid + i]) + (D(buf[mid + i]) * D(shortWindowsPrev[i])));
        for (int i = 0; i < mid; i++) output[mid + shortLen + i] = (float)(D(overlap[mid + shortLen + i]) + D(buf[mid + shortLen + i]));
        for (int i = 0; i < length; i++) overlap[i] = (float)(D(buf[length + i]) * D(longWindows[length - 1 - i]));
        break;

    default:
        break;
    }
#undef D
}

/* -------------------------------------------------------------------- TNS */

/* tns.js:105-177.  One channel-frame, in place.  `block` is one TNS block of
 * the aacfb.h blob (n_filt[8] + filters).  mode: AS_SHIPPED leaves data
 * untouched (tns.js:122 yields NaN bounds so neither filter loop runs);
 * FIXED_AR = tmp->top with decode=true; FIXED_MA = tmp->top with decode=false.
 * sample_index selects maxBands (tns.js:23) and swbOffsets/swbCount
 * (ics.js:297-305). */
static void tns_process(const aacfb_frame_info *info, const uint8_t *block, size_t block_bytes,
                        int sample_index, uint32_t mode, float *data) {
    if ((mode & AACFB_TNS_MODE_MASK) == AACFB_TNS_AS_SHIPPED) return;
    if (!block || block_bytes < 8) return;
    const int decode = ((mode & AACFB_TNS_MODE_MASK) == AACFB_TNS_FIXED_AR);
    const int is_short = (info->window_sequence == AACFB_EIGHT_SHORT_SEQUENCE);
    const int si = sample_index < 12 ? sample_index : 11;
    const uint16_t *swbOffsets = is_short ? SWB_OFFSET_128[si].off : SWB_OFFSET_1024[si].off;
    const int swbCount = is_short ? SWB_SHORT_WINDOW_COUNT[si] : SWB_LONG_WINDOW_COUNT[si];
    const int windowCount = is_short ? 8 : 1;
    const int maxBands = TNS_MAX_BANDS_1024[sample_index < 13 ? sample_index : 12]; /* tns.js:23 */
    const int mmm = maxBands < info->max_sfb ? maxBands : info->max_sfb;          /* tns.js:106 */
    float lpc[AACFB_TNS_MAX_ORDER];
    float tmp[AACFB_TNS_MAX_ORDER + 1]; /* JS tmp has 20 slots: slot 20 handled below */
    memset(tmp, 0, sizeof tmp);

    const uint8_t *n_filt = block;
    const uint8_t *p = block + 8, *end_p = block + block_bytes;

    for (int w = 0; w < 8; w++) {
        int bottom = swbCount; /* tns.js:113 */
        for (int filt = 0; filt < n_filt[w]; filt++) {
            if (p + 4 > end_p) return;
            aacfb_tns_filter h;
            memcpy(&h, p, 4);
            const float *autoc = (const float *)(p + 4);
            p += 4 + 4 * (size_t)h.order;
            if (w >= windowCount) continue; /* tns.js:111 loops w < windowCount only */
            int top = bottom;               /* tns.js:121 */
            bottom = top - h.length;        /* tns.js:122 with tmp -> top */
            if (bottom < 0) bottom = 0;
            int order = h.order;
            if (order == 0) continue;       /* tns.js:125 */

            /* calculate lpc coefficients, tns.js:128-140 */
            for (int i = 0; i < order; i++) {
                double r = -(double)autoc[i];
                lpc[i] = (float)r;
                for (int j = 0, len = (i + 1) >> 1; j < len; j++) {
                    double f = lpc[j], b = lpc[i - 1 - j];
                    lpc[j] = (float)(f + r * b);
                    lpc[i - 1 - j] = (float)(b + r * f);
                }
            }

            int start = swbOffsets[bottom < mmm ? bottom : mmm]; /* tns.js:142 */
            int end = swbOffsets[top < mmm ? top : mmm];         /* tns.js:143 */
            int size, inc = 1;
            if ((size = end - start) <= 0) continue;             /* tns.js:147 */
            if (h.direction) { inc = -1; start = end - 1; }      /* tns.js:149-152 */
            start += w * 128;                                    /* tns.js:154 */

            if (decode) {
                /* ar filter, tns.js:156-162 */
                for (int m = 0; m < size; m++, start += inc) {
                    int lim = m < order ? m : order;
                    for (int i = 1; i <= lim; i++)
                        data[start] = (float)((double)data[start] - (double)data[start - i * inc] * (double)lpc[i - 1]);
                }
            } else {
                /* ma filter, tns.js:163-174.  JS tmp = Float32Array(20): a read
                 * of tmp[20] is undefined -> NaN, a write to tmp[20] is dropped. */
                for (int m = 0; m < size; m++, start += inc) {
                    tmp[0] = data[start];
                    int lim = m < order ? m : order;
                    for (int i = 1; i <= lim; i++) {
                        double ti = (i < AACFB_TNS_MAX_ORDER) ? (double)tmp[i] : (double)NAN;
                        data[start] = (float)((double)data[start] + ti * (double)lpc[i - 1]);
                    }
                    for (int i = order; i > 0; i--)
                        if (i < AACFB_TNS_MAX_ORDER) tmp[i] = tmp[i - 1];
                }
            }
        }
    }
}

/* ------------------------------------------------------------ stereo tools */

/* What processMS / processIS read of a CPEElement and its two ICStreams
 * (cpe.js:24-75, ics.js:25-34,270-310).  Index 0 = left, 1 = right. */
typedef struct oracle_cpe {
    int32_t common_window, mask_present;     /* element.commonWindow, element.maskPresent */
    uint8_t ms_used[128];                    /* element.ms_used[idx]                      */
    int32_t window_sequence[2];              /* info.windowSequence (selects swbOffsets)  */
    int32_t group_count[2];                  /* info.groupCount                           */
    int32_t group_length[2][8];              /* info.groupLength[g]                       */
    int32_t max_sfb[2];                      /* info.maxSFB                               */
    int32_t band_types[2][120];              /* ics.bandTypes[idx]                        */
    int32_t sect_end[2][120];                /* ics.sectEnd[idx]                          */
    float   scale_factors[2][120];           /* ics.scaleFactors[idx] (Float32Array)      */
} oracle_cpe;

enum { NOISE_BT = 13, INTENSITY_BT2 = 14, INTENSITY_BT = 15 };  /* ics.js:39-41 */

static const uint16_t *cpe_offsets(const oracle_cpe *e, int ch, int sample_index) {
    /* ics.js:300-309: EIGHT_SHORT -> SWB_OFFSET_128, else SWB_OFFSET_1024 */
    return e->window_sequence[ch] == AACFB_EIGHT_SHORT_SEQUENCE ? SWB_OFFSET_128[sample_index].off
                                                                : SWB_OFFSET_1024[sample_index].off;
}

/* decoder.js:379-404 */
static void process_ms(const oracle_cpe *e, int sample_index, float *left, float *right) {
    const uint16_t *offsets = cpe_offsets(e, 0, sample_index);   /* ics = element.left */
    const int windowGroups = e->group_count[0], maxSFB = e->max_sfb[0];
    int groupOff = 0, idx = 0;
    for (int g = 0; g < windowGroups; g++) {
        for (int i = 0; i < maxSFB; i++, idx++) {
            if (e->ms_used[idx] && e->band_types[0][idx] < NOISE_BT && e->band_types[1][idx] < NOISE_BT) {
                for (int w = 0; w < e->group_length[0][g]; w++) {
                    const int off = groupOff + w * 128 + offsets[i];
                    for (int j = 0; j < offsets[i + 1] - offsets[i]; j++) {
                        const double t = (double)left[off + j] - (double)right[off + j];
                        left[off + j] = (float)((double)left[off + j] + (double)right[off + j]);
                        right[off + j] = (float)t;
                    }
                }
            }
        }
        groupOff += e->group_length[0][g] * 128;
    }
}

/* decoder.js:337-376 */
static void process_is(const oracle_cpe *e, int sample_index, const float *left, float *right) {
    const uint16_t *offsets = cpe_offsets(e, 1, sample_index);   /* ics = element.right */
    const int windowGroups = e->group_count[1], maxSFB = e->max_sfb[1];
    const int32_t *bandTypes = e->band_types[1], *sectEnd = e->sect_end[1];
    int idx = 0, groupOff = 0;
    for (int g = 0; g < windowGroups; g++) {
        for (int i = 0; i < maxSFB;) {
            const int end = sectEnd[idx];
            if (bandTypes[idx] == INTENSITY_BT || bandTypes[idx] == INTENSITY_BT2) {
                for (; i < end; i++, idx++) {
                    double c = bandTypes[idx] == INTENSITY_BT ? 1 : -1;
                    if (e->mask_present) c *= e->ms_used[idx] ? -1 : 1;
                    const double scale = c * (double)e->scale_factors[1][idx];
                    for (int w = 0; w < e->group_length[1][g]; w++) {
                        const int off = groupOff + w * 128 + offsets[i], len = offsets[i + 1] - offsets[i];
                        for (int j = 0; j < len; j++) right[off + j] = (float)((double)left[off + j] * scale);
                    }
                }
            } else {
                idx += end - i;
                i = end;
            }
        }
        groupOff += e->group_length[1][g] * 128;
    }
}

/* processPair's stereo part, decoder.js:300-307: M/S only with a common window and a mask, then IS. */
__attribute__((visibility("default")))
void aacfb_oracle_stereo(const oracle_cpe *e, int sample_index, float *left, float *right) {
    if (e->common_window && e->mask_present) process_ms(e, sample_index, left, right);
    process_is(e, sample_index, left, right);
}

/* ------------------------------------------------------------- ADTS header */

/* AV.Bitstream as readHeader uses it: read(n) MSB first, advance(n). */
typedef struct { const uint8_t *p; size_t bit, bits; } obits;
static uint32_t ob_read(obits *b, int n) {
    uint32_t v = 0;
    for (int i = 0; i < n; i++, b->bit++) v = (v << 1) | (b->bit < b->bits ? (b->p[b->bit >> 3] >> (7 - (b->bit & 7))) & 1u : 0u);
    return v;
}
static void ob_advance(obits *b, int n) { b->bit += (size_t)n; }

/* adts_demuxer.js:28-52, statement by statement.  out[0..5] = profile, samplingIndex, chanConfig,
 * frameLength, numFrames, bits consumed.  Returns 0, or -1 for 'Invalid ADTS header.' */
__attribute__((visibility("default")))
int aacfb_oracle_adts_header(const uint8_t *data, size_t size, uint32_t *out) {
    obits st = {data, 0, size * 8};
    if (ob_read(&st, 12) != 0xfff) return -1;                 /* :29-30 */
    ob_advance(&st, 3);                                       /* :33 mpeg version and layer */
    const int protectionAbsent = ob_read(&st, 1) != 0;        /* :34 */
    out[0] = ob_read(&st, 2) + 1;                             /* :36 profile */
    out[1] = ob_read(&st, 4);                                 /* :37 samplingIndex */
    ob_advance(&st, 1);                                       /* :39 private */
    out[2] = ob_read(&st, 3);                                 /* :40 chanConfig */
    ob_advance(&st, 4);                                       /* :41 */
    out[3] = ob_read(&st, 13);                                /* :43 frameLength */
    ob_advance(&st, 11);                                      /* :44 fullness */
    out[4] = ob_read(&st, 2) + 1;                             /* :46 numFrames */
    if (!protectionAbsent) ob_advance(&st, 16);               /* :48-49 */
    out[5] = (uint32_t)st.bit;
    return 0;
}

/* --------------------------------------------------------- exported entry */

#define API __attribute__((visibility("default")))

API void aacfb_oracle_init(void) { pthread_once(&g_once, init_tables); }

/* Copy a table as floats (same numbering as aacfb_get_table in aacfb.h).
 * MDCT tables are returned rounded to float; use _table_f64 for the doubles. */
API int aacfb_oracle_table(int which, float *dst, int cap) {
    aacfb_oracle_init();
    int n = 0;
    switch (which) {
    case 0: n = 1024; if (cap < n) return -1; for (int i = 0; i < 512; i++) { dst[2*i] = g_roots_long[i][0]; dst[2*i+1] = g_roots_long[i][1]; } break;
    case 1: n = 128;  if (cap < n) return -1; for (int i = 0; i < 64; i++)  { dst[2*i] = g_roots_short[i][0]; dst[2*i+1] = g_roots_short[i][1]; } break;
    case 2: n = 1024; if (cap < n) return -1; for (int i = 0; i < 512; i++) { dst[2*i] = (float)g_mdct_2048[i][0]; dst[2*i+1] = (float)g_mdct_2048[i][1]; } break;
    case 3: n = 128;  if (cap < n) return -1; for (int i = 0; i < 64; i++)  { dst[2*i] = (float)g_mdct_256[i][0]; dst[2*i+1] = (float)g_mdct_256[i][1]; } break;
    case 4: n = 1024; if (cap < n) return -1; memcpy(dst, g_sine_1024, 4096); break;
    case 5: n = 1024; if (cap < n) return -1; memcpy(dst, g_kbd_1024, 4096); break;
    case 6: n = 128;  if (cap < n) return -1; memcpy(dst, g_sine_128, 512); break;
    case 7: n = 128;  if (cap < n) return -1; memcpy(dst, g_kbd_128, 512); break;
    default: return -1;
    }
    return n;
}

API int aacfb_oracle_table_f64(int which, double *dst, int cap) {
    aacfb_oracle_init();
    if (which == 2) { if (cap < 1024) return -1; memcpy(dst, g_mdct_2048, sizeof g_mdct_2048); return 1024; }
    if (which == 3) { if (cap < 128) return -1;  memcpy(dst, g_mdct_256, sizeof g_mdct_256);  return 128; }
    return -1;
}

/* Inner seams, for unit tests. */
API void aacfb_oracle_fft(int length, float *aos /* [length][2] in place */) {
    aacfb_oracle_init();
    static __thread float rev[512][2];
    fft_process_inverse(length, (float (*)[2])aos, rev);
}
API void aacfb_oracle_mdct(int N, const float *input, float *output) {
    aacfb_oracle_init();
    static __thread mdct_scratch sc;
    mdct_process(N, input, 0, output, 0, &sc);
}
API void aacfb_oracle_filterbank(const aacfb_frame_info *info, const float *input, float *output, float *overlap) {
    aacfb_oracle_init();
    static __thread mdct_scratch sc;
    static __thread float buf[2048];
    memset(output, 0, 4096);
    filterbank_process(info, input, output, overlap, buf, &sc);
}
API void aacfb_oracle_tns(const aacfb_frame_info *info, const uint8_t *block, size_t block_bytes,
                          int sample_index, uint32_t mode, float *data) {
    aacfb_oracle_init();
    tns_process(info, block, block_bytes, sample_index, mode, data);
}

/* ------------------------------------------------------------ inverse quantisation
 * ICStream.decodeSpectralData, ics.js:203-266, with the Huffman decoder's output (buf[j],
 * ics.js:247) taken from aacfb_qframe.q and bandTypes / scaleFactors from aacfb_qframe.band
 * (include/aacfb.h).  tables.js:168-191 build the two lookup tables with Math.pow. */
static float g_iq_table[8191];        /* tables.js:181-191 */
static float g_sf_table[428];         /* tables.js:168-176 */
static pthread_once_t g_dq_once = PTHREAD_ONCE_INIT;
static void init_dequant(void) {
    double four_thirds = 4.0 / 3.0;
    for (int i = 0; i < 8191; i++) g_iq_table[i] = (float)pow((double)i, four_thirds);
    for (int i = 0; i < 428; i++) g_sf_table[i] = (float)pow(2.0, (i - 200) / 4.0);
}
/* JS ToInt32 of an integral double (the |0 of ics.js:234) */
static double to_int32(double x) {
    if (isnan(x) || isinf(x)) return 0;
    double m = fmod(trunc(x), 4294967296.0);
    if (m < 0) m += 4294967296.0;
    if (m >= 2147483648.0) m -= 4294967296.0;
    return m + 0.0;   /* ToInt32 never yields -0 (fmod keeps the sign of a negative multiple of 2^32) */
}
static void decode_spectral_data(const aacfb_qframe *qf, const aacfb_frame_info *info, int sample_index, float *data) {
    pthread_once(&g_dq_once, init_dequant);
    memset(data, 0, 4096);                       /* this.data = new Float32Array(frameLength), ics.js:28 */
    int is_short = info->window_sequence == AACFB_EIGHT_SHORT_SEQUENCE;
    const uint16_t *offsets = is_short ? SWB_OFFSET_128[sample_index].off : SWB_OFFSET_1024[sample_index].off; /* ics.js:301,307 */
    int maxSFB = info->max_sfb, windowGroups = 0;
    while (windowGroups < 8 && qf->group_len[windowGroups]) windowGroups++;
    double randomState = (double)0x1F2E3D4C;      /* ics.js:31; a new ICStream per element per frame, decoder.js:146,154 */
    int groupOff = 0, idx = 0;
    for (int g = 0; g < windowGroups; g++) {
        int groupLen = qf->group_len[g];
        for (int sfb = 0; sfb < maxSFB; sfb++, idx++) {
            unsigned code = qf->band[idx], kind = code & AACFB_BAND_KIND_MASK, ti = code & AACFB_BAND_INDEX_MASK;
            int off = groupOff + offsets[sfb], width = offsets[sfb + 1] - offsets[sfb];
            /* scaleFactors[idx] is a Float32Array element: +-SCALEFACTOR_TABLE[i], `undefined` -> NaN outside the table */
            float tab = ti < 428 ? g_sf_table[ti] : NAN;
            if (kind == AACFB_BAND_ZERO) {                         /* ics.js:222-227 */
                for (int group = 0; group < groupLen; group++, off += 128)
                    for (int i = off; i < off + width; i++) data[i] = 0;
            } else if (kind == AACFB_BAND_NOISE) {                 /* ics.js:228-242 */
                float sf = -tab;                                   /* ics.js:158 */
                for (int group = 0; group < groupLen; group++, off += 128) {
                    double energy = 0;
                    for (int k = 0; k < width; k++) {
                        randomState = to_int32(randomState * (double)(1664525 + 1013904223));
                        data[off + k] = (float)randomState;
                        energy += (double)data[off + k] * (double)data[off + k];
                    }
                    double scale = (double)sf / sqrt(energy);
                    for (int k = 0; k < width; k++) data[off + k] = (float)((double)data[off + k] * scale);
                }
            } else {                                               /* ics.js:243-256 */
                float sf = tab;                                    /* ics.js:171 */
                for (int group = 0; group < groupLen; group++, off += 128)
                    for (int k = 0; k < width; k++) {
                        int b = qf->q[off + k];
                        float iq;
                        if (b > 0) iq = b < 8191 ? g_iq_table[b] : NAN;
                        else iq = -b < 8191 ? -g_iq_table[-b] : NAN;
                        data[off + k] = iq;
                        data[off + k] = (float)((double)data[off + k] * (double)sf);
                    }
            }
        }
        groupOff += groupLen << 7;
    }
}
API void aacfb_oracle_dequant(const aacfb_qframe *qf, const aacfb_frame_info *info, int sample_index, float *data) {
    decode_spectral_data(qf, info, sample_index, data);
}
API int aacfb_oracle_dequant_table(int which, float *dst, int cap) {
    pthread_once(&g_dq_once, init_dequant);
    if (which == 0) { if (cap < 8191) return -1; memcpy(dst, g_iq_table, sizeof g_iq_table); return 8191; }
    if (which == 1) { if (cap < 428) return -1; memcpy(dst, g_sf_table, sizeof g_sf_table); return 428; }
    return -1;
}

/* AACFB_PCM_S16 (include/aacfb.h): Int16Array[i] = Math.max(-32768, Math.min(32767, Math.round(x))).
 * Math.round = floor(x + 0.5) evaluated exactly (x is a float, the sum is exact in double);
 * NaN survives min/max and is stored as 0. */
static int16_t pcm_s16(float x) {
    if (isnan(x)) return 0;
    double r = floor((double)x + 0.5);
    if (r > 32767) r = 32767;
    if (r < -32768) r = -32768;
    return (int16_t)r;
}
API void aacfb_oracle_pcm_s16(const float *x, int16_t *out, int n) { for (int i = 0; i < n; i++) out[i] = pcm_s16(x[i]); }

/* The whole path for a batch, same contract as aacfb_process (aacfb.h):
 *   spectra [S][T][C][1024], info [S][T][C], pcm [S][T][1024][C],
 *   overlap [S][C][1024] in/out.   decoder.js:263-269 + :204-213. */
typedef struct {
    const float *spectra; const aacfb_frame_info *info; const uint8_t *tns_blob;
    const uint32_t *tns_offsets; float *overlap; float *pcm;
    int S, T, C, sample_index; uint32_t flags; int s0, s1;
    const aacfb_qframe *q;   /* != NULL: input is quantised (decodeSpectralData runs first) */
    int16_t *pcm16;          /* != NULL: AACFB_PCM_S16 output instead of pcm */
} job_t;

static void *job_run(void *arg) {
    job_t *j = (job_t *)arg;
    mdct_scratch *sc = (mdct_scratch *)malloc(sizeof *sc);
    float *buf = (float *)malloc(2048 * sizeof(float));
    float *data = (float *)malloc(1024 * sizeof(float));  /* ics.data copy (caller's spectra stay const) */
    float *chan = (float *)malloc((size_t)j->C * 1024 * sizeof(float)); /* this.data[ch], decoder.js:229-231 */
    const int C = j->C, T = j->T;
    for (int s = j->s0; s < j->s1; s++) {
        for (int t = 0; t < T; t++) {
            size_t f = (size_t)s * T + t;
            for (int c = 0; c < C; c++) {
                size_t cf = f * C + c;
                const aacfb_frame_info *inf = &j->info[cf];
                if (j->q) decode_spectral_data(&j->q[cf], inf, j->sample_index, data);   /* ics.js:80 */
                else memcpy(data, j->spectra + cf * 1024, 4096);
                if (inf->tns_present && j->tns_offsets && j->tns_blob) { /* decoder.js:263-264 */
                    uint32_t o0 = j->tns_offsets[cf], o1 = j->tns_offsets[cf + 1];
                    if (o1 > o0) tns_process(inf, j->tns_blob + o0, o1 - o0, j->sample_index, j->flags, data);
                }
                float *out = chan + (size_t)c * 1024;
                memset(out, 0, 4096);
                filterbank_process(inf, data, out, j->overlap + ((size_t)s * C + c) * 1024, buf, sc); /* decoder.js:269 */
            }
            /* Interleave channels, decoder.js:204-213 */
            size_t jj = 0;
            if (j->pcm16) {
                int16_t *o = j->pcm16 + f * 1024 * C;
                for (int k = 0; k < 1024; k++)
                    for (int i = 0; i < C; i++) o[jj++] = pcm_s16(chan[(size_t)i * 1024 + k]);
            } else {
                float *o = j->pcm + f * 1024 * C;
                for (int k = 0; k < 1024; k++)
                    for (int i = 0; i < C; i++) o[jj++] = (float)((double)chan[(size_t)i * 1024 + k] / 32768);
            }
        }
    }
    free(sc); free(buf); free(data); free(chan);
    return NULL;
}

static int process_any(const float *spectra, const aacfb_qframe *q, const aacfb_frame_info *info, const uint8_t *tns_blob,
                       const uint32_t *tns_offsets, float *overlap, float *pcm, int16_t *pcm16, int S, int T, int C,
                       int sample_index, uint32_t flags, int n_threads);
API int aacfb_oracle_process(const float *spectra, const aacfb_frame_info *info, const uint8_t *tns_blob,
                             const uint32_t *tns_offsets, float *overlap, float *pcm, int S, int T, int C,
                             int sample_index, uint32_t flags, int n_threads) {
    return process_any(spectra, NULL, info, tns_blob, tns_offsets, overlap, pcm, NULL, S, T, C, sample_index, flags, n_threads);
}
/* same contract as aacfb_process_io (no stereo tools: callers apply aacfb_oracle_stereo first) */
API int aacfb_oracle_process_io(const void *input, uint32_t in_format, const aacfb_frame_info *info, const uint8_t *tns_blob,
                                const uint32_t *tns_offsets, float *overlap, void *pcm, uint32_t pcm_format, int S, int T,
                                int C, int sample_index, uint32_t flags, int n_threads) {
    return process_any(in_format == AACFB_IN_Q16 ? NULL : (const float *)input,
                       in_format == AACFB_IN_Q16 ? (const aacfb_qframe *)input : NULL, info, tns_blob, tns_offsets, overlap,
                       pcm_format == AACFB_PCM_S16 ? NULL : (float *)pcm, pcm_format == AACFB_PCM_S16 ? (int16_t *)pcm : NULL,
                       S, T, C, sample_index, flags, n_threads);
}
static int process_any(const float *spectra, const aacfb_qframe *q, const aacfb_frame_info *info, const uint8_t *tns_blob,
                       const uint32_t *tns_offsets, float *overlap, float *pcm, int16_t *pcm16, int S, int T, int C,
                       int sample_index, uint32_t flags, int n_threads) {
    aacfb_oracle_init();
    if (S <= 0 || T < 0 || C <= 0) return -1;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > S) n_threads = S;
    job_t *jobs = (job_t *)calloc(n_threads, sizeof *jobs);
    pthread_t *th = (pthread_t *)calloc(n_threads, sizeof *th);
    for (int i = 0; i < n_threads; i++) {
        jobs[i] = (job_t){spectra, info, tns_blob, tns_offsets, overlap, pcm, S, T, C, sample_index, flags,
                          (int)((long)S * i / n_threads), (int)((long)S * (i + 1) / n_threads), q, pcm16};
        if (n_threads == 1) job_run(&jobs[i]);
        else pthread_create(&th[i], NULL, job_run, &jobs[i]);
    }
    if (n_threads > 1) for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
    free(jobs); free(th);
    return 0;
}
