// csrc/conv_sched.h -- host-side scheduling arithmetic of CONV mode: how a level is cut into (strip, segment) CTAs,
// which levels are chained, how many builds are in flight.  Free of CUDA types so that the CPU test suite can
// compile it with g++ (tests/test_host.py); the kernels and launchers include it for the same constants.
#pragma once

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

// Frame lanes: builds of different frame slots in flight at once.  Row bands reading their neighbours' planes in
// place keep at most 3: the CTAs at a band edge spin until the neighbour GPU has published the level they read, so
// the builds in flight must stay few enough that waiting CTAs can never fill a GPU.
inline int frame_lanes(int conv_lanes, int frames, bool banded) {
    int n = conv_lanes < frames ? conv_lanes : frames;
    if (banded && n > 3) n = 3;
    return n < 1 ? 1 : (n > 16 ? 16 : n);
}

}  // namespace sspyr
