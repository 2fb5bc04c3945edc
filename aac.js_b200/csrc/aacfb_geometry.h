// aacfb_geometry.h -- how a batch is cut into work items (host + device).
#pragma once
#include <stddef.h>

#if defined(__CUDACC__)
#define AACFB_HD __host__ __device__ __forceinline__
#define AACFB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define AACFB_HD inline
#define AACFB_HD_NOINLINE inline
#endif

namespace aacfb {

// ------------------------------------------------------------ work geometry
// The batch is S streams x T frames x nc channels.  Chain h = s*nc + j is one
// (stream, channel) sequence through time; chains are paired (2i, 2i+1) and
// time is cut into chunks of L frames.  One work item = (pair, chunk).  A
// chunk that does not start at t = 0 first runs frame t0-1 as a *halo* (no
// output) to rebuild the overlap that frame t0 needs: overlap[t] depends on
// frame t alone (filter_bank.js:114-116), so no state crosses items.
struct Geometry {
    int S, T, nc;        // batch shape: spectra [S][T][nc][1024], pcm [S][T][1024][nc]
    int c_state, c0;     // overlap state is [*][c_state][1024]; chain j is state channel c0+j
    int s_base;          // first stream of this batch in the overlap state
    int L, n_chunks, n_pairs;
};
AACFB_HD Geometry make_geometry(int S, int T, int nc, int c_state, int c0, int s_base, int L) {
    Geometry g;
    g.S = S; g.T = T; g.nc = nc; g.c_state = c_state; g.c0 = c0; g.s_base = s_base;
    g.L = L < 1 ? 1 : L;
    g.n_chunks = (T + g.L - 1) / g.L;
    g.n_pairs = (S * nc + 1) / 2;
    return g;
}
struct Item {
    int nch;             // live chains
    int s[2], j[2];      // stream / channel-in-batch of each chain
    int t0, t1;          // frames [t0, t1) are emitted
    bool interleaved;
};
AACFB_HD Item make_item(const Geometry &g, int item) {
    Item it;
    // chunk-major order: the (shorter) last chunks of all pairs are handed out last
    const int chunk = item / g.n_pairs, pair = item % g.n_pairs;
    const int h0 = 2 * pair, total = g.S * g.nc;
    it.nch = (h0 + 1 < total) ? 2 : 1;
    for (int c = 0; c < 2; ++c) {
        const int h = (h0 + c < total) ? h0 + c : h0;
        it.s[c] = h / g.nc;
        it.j[c] = h % g.nc;
    }
    it.t0 = chunk * g.L;
    it.t1 = it.t0 + g.L < g.T ? it.t0 + g.L : g.T;
    it.interleaved = it.nch == 2 && g.nc == 2 && it.s[0] == it.s[1] && it.j[0] == 0;
    return it;
}
AACFB_HD size_t cf_index(const Geometry &g, int s, int t, int j) { return ((size_t)s * g.T + t) * g.nc + j; }
AACFB_HD size_t state_index(const Geometry &g, int s, int j) {
    return ((size_t)(g.s_base + s) * g.c_state + g.c0 + j) * 1024;
}

}  // namespace aacfb
