// csrc/conv_sched.h -- host-side scheduling arithmetic of CONV mode: how a level is cut into (strip, segment) CTAs,
// which levels are chained, how many builds are in flight.  Free of CUDA types so that the CPU test suite can
// compile it with g++ (tests/test_host.py); the kernels and launchers include it for the same constants.
#pragma once
#include <algorithm>
#include <vector>

namespace sspyr {

constexpr int CONV_TW = 128;             // strip / tile width (outputs)
constexpr int STRIP_TH = 32;             // output rows per step of the marching strip kernel
constexpr int STRIP_CTAS_PER_SM = 4;     // 256-thread CTAs at <= 64 registers and <= 55 KB of shared memory (every radius <= 12)

// Vertical segmentation of a level for the strip kernel: about `waves` co-resident waves of CTAs, segments of a
// multiple of 32 rows, at least seg_min.  Depends on the plane geometry only, so all levels of an octave get the
// same (strip, segment) grid -- which the level chaining relies on.
// waves <= 0: automatic -- 3 waves, but segments of at least `long_rows` rows (8 steps: the 2R warm-up rows and the
// exposed first load are paid once per segment) as long as that still leaves 1.5 waves; measured on 8K: 5-step
// segments x 2.7 waves 0.683 ms, 8-step x 1.7 waves 0.668 ms, 4-step x 3.4 waves 0.734 ms per pyramid.
inline int march_seg_rows(int H, int W, int frames, int sms, int waves, int seg_min, int long_rows = 8 * STRIP_TH) {
    const long long strips = (long long)((W + CONV_TW - 1) / CONV_TW) * frames;
    auto rows_for = [&](long long ctas) {
        long long segs = ctas / strips;
        if (segs < 1) segs = 1;
        const int r = (int)((H + segs - 1) / segs);
        return (r + STRIP_TH - 1) / STRIP_TH * STRIP_TH;
    };
    int seg_rows = rows_for((long long)sms * STRIP_CTAS_PER_SM * (waves > 0 ? waves : 3));
    if (waves <= 0 && seg_rows < long_rows) {
        const long long ctas_long = strips * ((H + long_rows - 1) / long_rows);
        if (2 * ctas_long >= 3LL * sms * STRIP_CTAS_PER_SM) seg_rows = long_rows;
    }
    return seg_rows < seg_min ? seg_min : seg_rows;
}

inline long long march_ctas(int H, int W, int frames, int seg_rows) {
    return (long long)((W + CONV_TW - 1) / CONV_TW) * ((H + seg_rows - 1) / seg_rows) * frames;
}

// Level chaining (conv_march.cuh) pays off only when a level is more than one wave of CTAs: a smaller level has no
// idle tail to fill -- all of its CTAs run at once and finish together -- and the counter handshake then only adds
// latency (1080p: 5 % slower).  conv_chain: 0 off, 1 automatic, 2 always.
inline bool level_chained(int conv_chain, long long ctas, int sms) {
    return conv_chain > 1 || (conv_chain == 1 && ctas > (long long)STRIP_CTAS_PER_SM * sms);
}

// Cascade (conv_cascade.cuh): segment height of an octave.  Consumers trail producers by about two diagonals of the
// block order = 2 x levels x strips items, so the rows in flight between a plane's write and its read-back grow with
// the segment height: keep that window (Gaussian planes only, DoG stores are streaming) near a third of the 126 MB
// L2 -- 8192 / strips rows, a multiple of 32 in [32, 256].  The 2R warm-up rows of a segment are the price of short
// segments (8K: 128 rows, +4 % row-pass work; 16K: 64 rows).
inline int cascade_seg_rows(int W, int tuned) {
    if (tuned > 0) return (tuned + STRIP_TH - 1) / STRIP_TH * STRIP_TH;
    const int strips = (W + CONV_TW - 1) / CONV_TW;
    int r = 8192 / (strips > 0 ? strips : 1) / STRIP_TH * STRIP_TH;
    return r < STRIP_TH ? STRIP_TH : (r > 8 * STRIP_TH ? 8 * STRIP_TH : r);
}

// Radius class a cascade level runs in (zero-padded taps): few distinct hot loops per SM, see conv_cascade.cuh.
constexpr int CASC_RADIUS_PAD = 12;
inline int cascade_radius_class(int R) { return R <= 6 ? 6 : (R <= 10 ? 10 : 12); }

struct CascItemGeom { int seg_rows, nsegs, nstrips, first_level, H, W; };

// Source rows [ya, yb) of segment j of a level with (class) radius R in a plane of H rows.
inline void cascade_source_rows(const CascItemGeom& G, int j, int R, int* ya, int* yb) {
    *ya = std::max(j * G.seg_rows - R, 0);
    *yb = std::min((j + 1) * G.seg_rows + R, G.H);
}

// Is `order` (items as octave << 28 | level << 24 | strip << 14 | segment, strips of a group consecutive) safe, i.e.
// does every item come after everything it reads?  Level s-1 segments that hold its source rows (same octave), or,
// for the first level of a lower octave, the level-S segments of the octave above that hold rows 2 ya .. 2 (yb - 1).
inline bool cascade_order_is_safe(const std::vector<unsigned>& order, const CascItemGeom* g, int octaves, int nl, int S,
                                  const int* radius) {
    std::vector<std::vector<long long>> first(octaves), last(octaves);      // first / last position of a group's strips
    for (int o = 0; o < octaves; ++o) { first[o].assign((size_t)nl * g[o].nsegs, -1); last[o].assign((size_t)nl * g[o].nsegs, -1); }
    for (size_t i = 0; i < order.size(); ++i) {
        const int o = (int)(order[i] >> 28), s = (int)((order[i] >> 24) & 15u), j = (int)(order[i] & 16383u);
        if (o >= octaves || s >= nl || j >= g[o].nsegs) return false;
        const size_t k = (size_t)s * g[o].nsegs + j;
        if (first[o][k] < 0) first[o][k] = (long long)i;
        last[o][k] = (long long)i;
    }
    for (int o = 0; o < octaves; ++o)
        for (int s = g[o].first_level; s < nl; ++s)
            for (int j = 0; j < g[o].nsegs; ++j) {
                const long long me = first[o][(size_t)s * g[o].nsegs + j];
                if (me < 0) return false;                                   // an item is missing
                int ya, yb;
                cascade_source_rows(g[o], j, radius[s], &ya, &yb);
                if (s > g[o].first_level) {
                    for (int jj = ya / g[o].seg_rows; jj <= std::min((yb - 1) / g[o].seg_rows, g[o].nsegs - 1); ++jj)
                        if (last[o][(size_t)(s - 1) * g[o].nsegs + jj] >= me) return false;
                } else if (o > 0) {
                    const CascItemGeom& U = g[o - 1];
                    for (int jj = (2 * ya) / U.seg_rows; jj <= std::min((2 * (yb - 1)) / U.seg_rows, U.nsegs - 1); ++jj)
                        if (last[o - 1][(size_t)S * U.nsegs + jj] >= me) return false;
                }
            }
    return true;
}

// The block order of one frame of the cascade.  Sort key = the octave-0 row at which an item can run: inside an octave
// one segment per segment index plus two per level; octave o+1 starts about one segment (plus the blur radius) after
// level S of octave o has passed the same place in the frame, so the octaves advance down the frame TOGETHER instead of
// one after the other (octave-major order measured 2.2x slower on 8K: each small octave is a long chain of few items).
// If the keyed order fails cascade_order_is_safe, the plain octave-major diagonal order (always safe) is returned.
inline std::vector<unsigned> cascade_item_table(const CascItemGeom* g, int octaves, int nl, int S, const int* radius,
                                                bool* keyed_out) {
    struct Item { long long key; int o, s, j; };
    std::vector<Item> items;
    long long off = 0;
    for (int o = 0; o < octaves; ++o) {
        const long long unit = (long long)g[o].seg_rows << o;
        for (int s = g[o].first_level; s < nl; ++s)
            for (int j = 0; j < g[o].nsegs; ++j)
                items.push_back({off + unit * (j + 1 + 2 * (s - g[o].first_level)), o, s, j});
        off += unit * (2 * (S - g[o].first_level) + 2) + ((long long)(2 * CASC_RADIUS_PAD + STRIP_TH) << (o + 1));
    }
    auto emit = [&](const std::vector<Item>& v) {
        std::vector<unsigned> tab;
        for (const Item& it : v)
            for (int x = 0; x < g[it.o].nstrips; ++x)
                tab.push_back((unsigned)it.o << 28 | (unsigned)it.s << 24 | (unsigned)x << 14 | (unsigned)it.j);
        return tab;
    };
    std::vector<Item> keyed = items;
    std::stable_sort(keyed.begin(), keyed.end(), [](const Item& a, const Item& b) {
        return a.key != b.key ? a.key < b.key : (a.o != b.o ? a.o < b.o : (a.s != b.s ? a.s < b.s : a.j < b.j));
    });
    std::vector<unsigned> tab = emit(keyed);
    const bool ok = cascade_order_is_safe(tab, g, octaves, nl, S, radius);
    if (keyed_out) *keyed_out = ok;
    if (ok) return tab;
    std::vector<Item> plain = items;                       // octave-major, diagonal t = j + 2 (s - first), level, segment
    std::stable_sort(plain.begin(), plain.end(), [&](const Item& a, const Item& b) {
        if (a.o != b.o) return a.o < b.o;
        const int ta = a.j + 2 * (a.s - g[a.o].first_level), tb = b.j + 2 * (b.s - g[b.o].first_level);
        return ta != tb ? ta < tb : (a.s != b.s ? a.s < b.s : a.j < b.j);
    });
    return emit(plain);
}

// Frame lanes: builds of different frame slots in flight at once.  Row bands reading their neighbours' planes in
// place: the CTAs at a band edge spin until the neighbour GPU has published the level they read.  A spinning CTA waits
// for a kernel that PRECEDES the neighbour's own spinners of that build in its stream, and whose own waits are for
// levels this GPU has already finished, so progress never depends on a free slot here; the round-1 cap of 3 builds in
// flight was conservative and is now only the fallback when conv_band_lanes is not set (default 6, sspyr_internal.h).
inline int frame_lanes(int conv_lanes, int frames, bool banded) {
    int n = conv_lanes < frames ? conv_lanes : frames;
    if (banded && n > 3) n = 3;
    return n < 1 ? 1 : (n > 16 ? 16 : n);
}

}  // namespace sspyr
