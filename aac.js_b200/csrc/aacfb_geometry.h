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
// (stream, channel) sequence through time; chains are paired (2i, 2i+1).  All
// pairs' frames, laid end to end, form one sequence of n_pairs*T *pair-frames*
// (flat index f = pair*T + t); a work item is a contiguous slice of Q of them.
// Frames cost the same, so equal slices balance the workers with no scheduling
// slack.  A slice that does not start at t = 0 first runs frame f0-1 as a
// *halo* (no output) to rebuild the overlap frame f0 needs -- overlap[t]
// depends on frame t alone (filter_bank.js:114-116) -- so no state crosses
// items; where a slice runs into the next pair it stores the finished pair's
// overlap and loads the next one's.
struct Geometry {
    int S, T, nc;        // batch shape: spectra [S][T][nc][1024], pcm [S][T][1024][nc]
    int c_state, c0;     // overlap state is [*][c_state][1024]; chain j is state channel c0+j
    int s_base;          // first stream of this batch in the overlap state
    int n_pairs, Q, n_items;
    int stride;          // items are handed out in the order slice = item * stride mod n_items
};
AACFB_HD Geometry make_geometry(int S, int T, int nc, int c_state, int c0, int s_base, int frames_per_item) {
    Geometry g;
    g.S = S; g.T = T; g.nc = nc; g.c_state = c_state; g.c0 = c0; g.s_base = s_base;
    g.n_pairs = (S * nc + 1) / 2;
    g.Q = frames_per_item < 1 ? 1 : frames_per_item;
    const long total = (long)g.n_pairs * T;
    g.n_items = (int)((total + g.Q - 1) / g.Q);
    // Neighbouring slices are neighbours in memory; workers that run side by side should not
    // be (DRAM channel / page spread), so the hand-out order strides through the slices with a
    // step near n_items * 0.38 that is coprime to n_items.
    int k = (int)(g.n_items * 0.381966f) | 1;
    auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
    while (k > 1 && gcd(k, g.n_items) != 1) k -= 2;
    g.stride = k < 1 ? 1 : k;
    return g;
}
// The two chains of a pair.
struct Pair {
    int nch;             // live chains
    int s[2], j[2];      // stream / channel-in-batch of each chain
    bool interleaved;    // channels 0,1 of one stereo stream: PCM is one [1024][2] row
};
AACFB_HD Pair make_pair(const Geometry &g, int pair) {
    Pair p;
    const int h0 = 2 * pair, total = g.S * g.nc;
    p.nch = (h0 + 1 < total) ? 2 : 1;
    for (int c = 0; c < 2; ++c) {
        const int h = (h0 + c < total) ? h0 + c : h0;
        p.s[c] = h / g.nc;
        p.j[c] = h % g.nc;
    }
    p.interleaved = p.nch == 2 && g.nc == 2 && p.s[0] == p.s[1] && p.j[0] == 0;
    return p;
}
// Flat pair-frame range [f0, f1) of an item; the halo frame (if any) is f0 - 1.
// (n_pairs * T fits in 31 bits: the C-ABI layer checks.)
AACFB_HD int item_slice(const Geometry &g, int item) {
    return (int)(((unsigned long long)(unsigned)item * (unsigned)g.stride) % (unsigned)g.n_items);
}
AACFB_HD int item_begin(const Geometry &g, int item) { return item_slice(g, item) * g.Q; }
AACFB_HD int item_end(const Geometry &g, int item) {
    const int total = g.n_pairs * g.T, e = (item_slice(g, item) + 1) * g.Q;
    return e < total ? e : total;
}
AACFB_HD size_t cf_index(const Geometry &g, int s, int t, int j) { return ((size_t)s * g.T + t) * g.nc + j; }
AACFB_HD size_t state_index(const Geometry &g, int s, int j) {
    return ((size_t)(g.s_base + s) * g.c_state + g.c0 + j) * 1024;
}

}  // namespace aacfb
