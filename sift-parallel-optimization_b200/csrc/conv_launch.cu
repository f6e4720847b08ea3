// csrc/conv_launch.cu -- host side of CONV mode: per-level launches of conv_kernel.cuh, the decimation chain
// between octaves, row-band halo staging, and the DoG extremum scan.  Also the halo part of the C ABI.
#include <algorithm>
#include <cstring>

#include <cudaTypedefs.h>

#include "conv_march.cuh"

namespace sspyr {

#define SSPYR_DECL(n) cudaError_t launch_conv_r##n(const ConvParams&, int, int, cudaStream_t, int, int, int);
SSPYR_DECL(1) SSPYR_DECL(2) SSPYR_DECL(3) SSPYR_DECL(4) SSPYR_DECL(5) SSPYR_DECL(6) SSPYR_DECL(7) SSPYR_DECL(8)
SSPYR_DECL(9) SSPYR_DECL(10) SSPYR_DECL(11) SSPYR_DECL(12) SSPYR_DECL(13) SSPYR_DECL(14) SSPYR_DECL(15) SSPYR_DECL(16)
SSPYR_DECL(20) SSPYR_DECL(24) SSPYR_DECL(28) SSPYR_DECL(32)
#undef SSPYR_DECL
#define SSPYR_DECL(n)                                                                                  \
    cudaError_t launch_march_r##n(const ConvParams&, int, cudaStream_t, int, int, const CUtensorMap*, int, bool); \
    int march_box_cols_r##n();
SSPYR_DECL(1) SSPYR_DECL(2) SSPYR_DECL(3) SSPYR_DECL(4) SSPYR_DECL(5) SSPYR_DECL(6) SSPYR_DECL(7) SSPYR_DECL(8)
SSPYR_DECL(9) SSPYR_DECL(10) SSPYR_DECL(11) SSPYR_DECL(12)
#undef SSPYR_DECL

namespace {

// One thread publishes "this band has finished level `value` of the running build" to its neighbours (system-scope
// release: everything the preceding kernels in the stream wrote is visible to a peer GPU that acquires the counter).
// Counter values are (build - 1) * CONV_FLAG_STRIDE + level + 1 with the build number read from the slot's epoch word.
__global__ void conv_signal_kernel(unsigned* flag, unsigned value, const unsigned* epoch) {
    __threadfence_system();
    const unsigned v = (*epoch - 1u) * CONV_FLAG_STRIDE + value;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(v) : "memory");
}

// One thread waits until the neighbour's counter reaches `need` (relative to the running build).  The neighbour runs
// on another GPU, so it makes progress on its own; a bounded spin (about 2 s) marks the handle instead of hanging.
__global__ void conv_wait_kernel(const unsigned* peer_flag, unsigned need, const unsigned* epoch, unsigned* my_flags) {
    const unsigned want = (*epoch - 1u) * CONV_FLAG_STRIDE + need;
    const long long t0 = clock64();
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(peer_flag) : "memory");
        if (v >= want) break;
        if (clock64() - t0 > 4000000000LL) { my_flags[CONV_FLAG_TIMEOUT] = want; return; }
        __nanosleep(200);
    }
}

// First kernel of a banded build: next build number of the slot (on the DEVICE -- the launch sequence has no per-build
// parameters and replays as a CUDA graph), then "nothing of this build may be written before both neighbours have
// finished reading the previous one": every octave counter of theirs at its end-of-build value.
__global__ void conv_begin_kernel(unsigned* epoch, const unsigned* peer_up, const unsigned* peer_dn, int octaves, int nl,
                                  unsigned* my_flags) {
    const unsigned b = *epoch + 1u;
    if (b >= 2u) {
        const unsigned prev_done = (b - 2u) * CONV_FLAG_STRIDE + (unsigned)nl;
        const long long t0 = clock64();
        for (int side = 0; side < 2; ++side) {
            const unsigned* f = side == 0 ? peer_up : peer_dn;
            if (!f) continue;
            for (int o = 0; o < octaves; ++o)
                for (;;) {
                    unsigned v;
                    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f + o) : "memory");
                    if (v >= prev_done) break;
                    if (clock64() - t0 > 4000000000LL) { my_flags[CONV_FLAG_TIMEOUT] = prev_done; o = octaves; break; }
                    __nanosleep(200);
                }
        }
    }
    *epoch = b;
}

// compiled radius that serves a requested one (taps are zero-padded up to it)
int compiled_radius(int r) { return r <= 16 ? r : r <= 20 ? 20 : r <= 24 ? 24 : r <= 28 ? 28 : 32; }

cudaError_t dispatch(int rt, const ConvParams& P, int src_kind, int variant, cudaStream_t st, int device, int frames, int sms) {
    switch (rt) {
#define SSPYR_CASE(n) case n: return launch_conv_r##n(P, src_kind, variant, st, device, frames, sms);
        SSPYR_CASE(1) SSPYR_CASE(2) SSPYR_CASE(3) SSPYR_CASE(4) SSPYR_CASE(5) SSPYR_CASE(6) SSPYR_CASE(7) SSPYR_CASE(8)
        SSPYR_CASE(9) SSPYR_CASE(10) SSPYR_CASE(11) SSPYR_CASE(12) SSPYR_CASE(13) SSPYR_CASE(14) SSPYR_CASE(15)
        SSPYR_CASE(16) SSPYR_CASE(20) SSPYR_CASE(24) SSPYR_CASE(28) SSPYR_CASE(32)
#undef SSPYR_CASE
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t dispatch_march(int r, const ConvParams& P, int src_kind, cudaStream_t st, int device, int frames,
                           const CUtensorMap* tmap, int seg_rows, bool pdl) {
    switch (r) {
#define SSPYR_CASE(n) case n: return launch_march_r##n(P, src_kind, st, device, frames, tmap, seg_rows, pdl);
        SSPYR_CASE(1) SSPYR_CASE(2) SSPYR_CASE(3) SSPYR_CASE(4) SSPYR_CASE(5) SSPYR_CASE(6) SSPYR_CASE(7) SSPYR_CASE(8)
        SSPYR_CASE(9) SSPYR_CASE(10) SSPYR_CASE(11) SSPYR_CASE(12)
#undef SSPYR_CASE
        default: return cudaErrorInvalidValue;
    }
}

int march_box_cols(int r) {
    switch (r) {
#define SSPYR_CASE(n) case n: return march_box_cols_r##n();
        SSPYR_CASE(1) SSPYR_CASE(2) SSPYR_CASE(3) SSPYR_CASE(4) SSPYR_CASE(5) SSPYR_CASE(6) SSPYR_CASE(7) SSPYR_CASE(8)
        SSPYR_CASE(9) SSPYR_CASE(10) SSPYR_CASE(11) SSPYR_CASE(12)
#undef SSPYR_CASE
        default: return 0;
    }
}

// Tensor map of a float plane for the strip kernel's TMA staging: (columns = row pitch, rows, frame slots),
// box = box_cols x 32 x 1, no swizzle, zero fill (never relied on: edge steps do not use TMA).
bool make_plane_tensor_map(CUtensorMap* map, const float* plane0, int pitch, int rows, int frames, size_t frame_floats,
                           int box_cols) {
    static PFN_cuTensorMapEncodeTiled encode = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
        cudaGetLastError();
    }
    if (!encode || box_cols <= 0 || box_cols > 256 || pitch < box_cols) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)(frames > 0 ? frames : 1)};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * sizeof(float), (cuuint64_t)frame_floats * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)STRIP_TH, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (rows < STRIP_TH) return false;
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(plane0), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// Minimum segment height of the strip kernel.  A handle with several frame slots streams frames: builds of different
// slots are in flight at once (frame lanes), the GPU is filled by them and what counts is the work per pyramid -- every
// segment pays 2R warm-up rows and one exposed load, so 96-row segments (3 steps) beat 32-row ones by 1.4x on band-sized
// and 1080p planes (profiles/r2_bands.md).  A single-slot handle builds one pyramid at a time: 32-row segments, the
// shortest critical path.  An explicit conv_seg_min always wins.
static int conv_seg_min_rows(const sspyr_ctx* h, int count) {
    if (h->tune.conv_seg_min > 0) return h->tune.conv_seg_min;
    if (count > 1) return 32;                             // a batched launch fills the GPU by itself (1080p x 8 frames: 96-row
                                                          // segments measured 14 % slower than 32-row ones)
    const bool banded = h->cfg.full_height != h->cfg.height;
    const int lanes = banded ? h->tune.conv_band_lanes : h->tune.conv_lanes;
    return (h->cfg.frames > 1 && lanes > 1 && !h->tune.timing) ? 96 : 32;
}

bool conv_has_up(const sspyr_ctx* h) { return h->cfg.band_row0 > 0; }
bool conv_has_down(const sspyr_ctx* h) { return h->cfg.band_row0 + h->cfg.height < h->cfg.full_height; }

float* conv_halo_plane(const sspyr_ctx* h, int octave, int down) {
    return h->d_halo + h->halo_off[octave] + (size_t)down * h->halo_rmax * h->oct[octave].pitch;
}

unsigned char* conv_halo_raw(const sspyr_ctx* h, int down) {
    return h->d_halo_raw + (size_t)down * h->halo_rmax * h->in_pitch_bytes;
}

// The blur that PRODUCES (octave, level) for frame slots first..first+count-1 (contiguous, own input slots).
// Does (level) of this handle run on the marching strip kernel?  (radius <= 12 and not disabled)
static bool level_marches(const sspyr_ctx* h, int level) {
    return h->conv[level].radius <= 12 && h->tune.conv_march != 0;
}

cudaError_t launch_conv_step(const sspyr_ctx* h, int first, int count, int octave, int level, cudaStream_t st,
                             int* launches, bool chain) {
    if (level == 0 && octave > 0) return cudaSuccess;      // written by the decimating epilogue of (octave-1, S)
    const int nl = h->nl, S = h->cfg.S;
    const OctGeom& g = h->oct[octave];
    ConvParams P{};
    int src_kind;
    if (level == 0) {                                        // octave 0 from the raw frame
        size_t pitch_bytes = 0;
        P.src = frame_input(h, first, &pitch_bytes);
        P.src_pitch = (int)(pitch_bytes / h->elem_bytes);
        P.src_frame_stride = h->in_frame_bytes / h->elem_bytes;
        src_kind = h->cfg.pixel_type;
        P.top_halo = conv_has_up(h) ? conv_halo_raw(h, 0) : nullptr;
        P.bot_halo = conv_has_down(h) ? conv_halo_raw(h, 1) : nullptr;
        const int R0 = h->conv[0].radius;
        if (h->peer[0].attached)      // last R rows of the band above / first rows of the band below, in place
            P.top_halo = h->peer[0].in + (size_t)first * h->peer[0].in_frame_bytes + (size_t)(h->peer[0].height - R0) * h->in_pitch_bytes;
        if (h->peer[1].attached) P.bot_halo = h->peer[1].in + (size_t)first * h->peer[1].in_frame_bytes;
    } else {
        P.src = frame_out(h, first) + g.off + (size_t)plane_index(nl, SSPYR_KIND_GAUSS, level - 1) * g.plane;
        P.src_pitch = g.pitch;
        P.src_frame_stride = h->frame_floats;
        src_kind = CONV_SRC_PLANE;
        P.top_halo = conv_has_up(h) ? conv_halo_plane(h, octave, 0) : nullptr;
        P.bot_halo = conv_has_down(h) ? conv_halo_plane(h, octave, 1) : nullptr;
        const int Rl = h->conv[level].radius;
        const int pi = plane_index(nl, SSPYR_KIND_GAUSS, level - 1);
        for (int side = 0; side < 2; ++side) {
            const sspyr_ctx::Peer& q = h->peer[side];
            if (!q.attached) continue;
            const float* pl = q.out + (size_t)first * q.frame_floats + q.off[octave] + (size_t)pi * q.plane[octave];
            if (side == 0) P.top_halo = pl + (size_t)(q.H[octave] - Rl) * g.pitch;
            else P.bot_halo = pl;
        }
    }
    const int R = h->conv[level].radius;
    const int RT = compiled_radius(R);
    P.halo_rows = R;
    float* obase = frame_out(h, first) + g.off;
    P.dst_g = obase + (size_t)plane_index(nl, SSPYR_KIND_GAUSS, level) * g.plane;
    P.dst_d = (level >= 1 && (h->cfg.outputs & SSPYR_OUT_DOG))
                  ? obase + (size_t)plane_index(nl, SSPYR_KIND_DOG, level - 1) * g.plane : nullptr;
    P.dst_frame_stride = h->frame_floats;
    P.dst_pitch = g.pitch;
    P.H = g.H;
    P.W = g.W;
    if (level == S && octave + 1 < h->octaves) {
        const OctGeom& n = h->oct[octave + 1];
        P.dst_dec = frame_out(h, first) + n.off;             // G_0 of the next octave
        P.dec_pitch = n.pitch;
        P.dec_H = n.H;
        P.dec_W = n.W;
    }
    P.timeout_mark = h->d_flag + CONV_FLAG_TIMEOUT;         // any bounded wait that gives up is reported by sspyr_sync
    std::memset(P.taps, 0, sizeof(P.taps));
    std::memcpy(P.taps + (RT - R), h->h_tables.data() + h->conv[level].taps_off, sizeof(float) * (2 * R + 1));
    const int variant = 0;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    // The marching strip kernel (conv_march.cuh) row-filters every input row once and measured faster at every size;
    // radii above 12 and conv_march = 0 use the one-tile-per-CTA kernel.
    const bool march = level_marches(h, level);
    const bool peered_any = h->peer[0].attached || h->peer[1].attached;
    const int seg_rows = march_seg_rows(g.H, g.W, count, sms, h->tune.conv_waves, conv_seg_min_rows(h, count));
    const long long ctas = march_ctas(g.H, g.W, count, seg_rows);
    // Level chaining (whole-pyramid builds): the strip kernels of one octave count finished segments; a level whose
    // source plane was produced by a chained strip kernel of the same octave (same grid) waits per segment instead of
    // per grid.  The first level of a chain -- octave 0 level 0 from the raw frame, or level 1 of a later octave, whose
    // base comes from the octave above through an event -- keeps the grid-wide dependency.  Only grids of more than one
    // wave are chained (level_chained, conv_sched.h).  Row bands chain too once BOTH their neighbours are attached over
    // peer memory with counters of their own: the seam is one more segment boundary (peer_seg_up / peer_seg_dn), the
    // whole-level progress flags then only order a level that starts a chain and the builds of a slot among each other.
    const bool band = conv_has_up(h) || conv_has_down(h);
    const bool band_chains = band && (!conv_has_up(h) || (h->peer[0].attached && h->peer[0].seg)) &&
                             (!conv_has_down(h) || (h->peer[1].attached && h->peer[1].seg)) && h->tune.conv_band_chain != 0;
    bool chained_dep = false;
    if (chain && march && h->d_seg && level_chained(h->tune.conv_chain, ctas, sms) && (!band || band_chains)) {
        unsigned* lv0 = h->d_seg + (size_t)first * h->seg_frame_stride + h->seg_off[octave];
        P.seg_pub = lv0 + (size_t)level * h->seg_cap[octave];
        P.seg_frame_stride = (unsigned)h->seg_frame_stride;
        P.seg_sys = (conv_has_up(h) ? 1 : 0) | (conv_has_down(h) ? 2 : 0);
        const bool src_same_octave = level >= 2 || (level == 1 && octave == 0);
        if (src_same_octave && level_marches(h, level - 1) && h->tune.pdl != 0) {
            P.seg_dep = lv0 + (size_t)(level - 1) * h->seg_cap[octave];
            chained_dep = true;
            const int Rl = h->conv[level].radius;
            if (h->peer[0].attached) {                        // the band above: the segment rows that hold its last R rows
                const sspyr_ctx::Peer& q = h->peer[0];
                const int nrows = march_seg_rows(q.H[octave], g.W, count, sms, h->tune.conv_waves, conv_seg_min_rows(h, count));
                P.peer_seg_up = q.seg + (size_t)first * q.seg_frame_stride + q.seg_off[octave] + (size_t)(level - 1) * q.seg_cap[octave];
                P.peer_up_nsegs = (q.H[octave] + nrows - 1) / nrows;
                P.peer_up_first = std::max(0, (q.H[octave] - Rl) / nrows);
                if (P.peer_up_nsegs - P.peer_up_first > 2) P.peer_up_first = P.peer_up_nsegs - 2;   // (32-row segments, R <= 12: two rows at most)
            }
            if (h->peer[1].attached) {                        // the band below: its first segment row holds its first R rows
                const sspyr_ctx::Peer& q = h->peer[1];
                P.peer_seg_dn = q.seg + (size_t)first * q.seg_frame_stride + q.seg_off[octave] + (size_t)(level - 1) * q.seg_cap[octave];
            }
        }
        P.timeout_mark = h->d_flag + CONV_FLAG_TIMEOUT;
        P.src_evict_first = h->tune.conv_l2hint != 0 && level >= 1;
    }
    // TMA staging for float-plane sources: the map covers the frames of this launch (frame = 3rd coordinate)
    CUtensorMap tmap;
    const CUtensorMap* tm = nullptr;
    if (march && src_kind == CONV_SRC_PLANE && h->tune.conv_tma != 0 &&
        make_plane_tensor_map(&tmap, static_cast<const float*>(P.src), g.pitch, g.H, count, h->frame_floats, march_box_cols(R)))
        tm = &tmap;
    // Peer-memory halos: per-octave progress counters.  After level s of build b octave o publishes
    // (b-1)*CONV_FLAG_STRIDE + s + 1; a level first waits until both neighbours have published the level whose
    // rows it is about to read (level s-1 of its octave, or level S of the octave above for the decimated base).
    // The strip kernel does both itself (edge CTAs wait, the last CTA signals); the tile kernel (R > 12) gets
    // one-thread wait / signal kernels around it.
    const bool peered = peered_any;
    // (a banded build covers one frame slot: `first`; every slot has its own counters, so builds of different
    //  slots can be in flight at once)
    unsigned* my_flags = h->d_flag + (size_t)CONV_FLAG_BLOCK * first;
    const size_t peer_block = (size_t)CONV_FLAG_BLOCK * first;
    const unsigned* epoch = my_flags + CONV_FLAG_EPOCH;      // bumped by conv_begin_kernel at the start of every build
    const bool first_level = octave == 0 && level == 0;
    const int wo = (level == 1 && octave > 0) ? octave - 1 : octave;
    const unsigned need = (unsigned)((level == 1 && octave > 0) ? S : level - 1) + 1;   // relative to the build
    const bool fused_sync = peered && march && h->tune.conv_fused_sync != 0;
    if (fused_sync) {
        if (!first_level && !chained_dep) {                  // (a chained level waits per segment, also across the seam)
            P.wait_up = h->peer[0].attached ? h->peer[0].flag + peer_block + wo : nullptr;
            P.wait_dn = h->peer[1].attached ? h->peer[1].flag + peer_block + wo : nullptr;
            P.wait_need = need;
        }
        P.signal_flag = my_flags + octave;
        P.signal_value = (unsigned)level + 1;
        P.done_count = my_flags + CONV_FLAG_DONE + 16 * octave + level;
        P.epoch = epoch;
        P.timeout_mark = h->d_flag + CONV_FLAG_TIMEOUT;
    } else if (peered && !first_level) {
        for (int side = 0; side < 2; ++side)
            if (h->peer[side].attached) {
                conv_wait_kernel<<<1, 1, 0, st>>>(h->peer[side].flag + peer_block + wo, need, epoch, h->d_flag);
                ++*launches;
            }
    }
    // (PDL on peered launches measured slightly slower -- except along a chain, which only exists through it)
    cudaError_t e = march ? dispatch_march(R, P, src_kind, st, h->device, count, tm, seg_rows, h->tune.pdl != 0 && (!peered_any || P.seg_pub != nullptr))
                          : dispatch(RT, P, src_kind, variant, st, h->device, count, sms);
    if (e == cudaSuccess) ++*launches;
    if (e == cudaSuccess && peered && !fused_sync) {
        conv_signal_kernel<<<1, 1, 0, st>>>(my_flags + octave, (unsigned)level + 1, epoch);
        ++*launches;
        e = cudaGetLastError();
    }
    return e;
}

// Start of a build on a band with attached neighbours: one single-thread kernel (conv_begin_kernel).
cudaError_t conv_begin_build(sspyr_ctx* h, int slot, cudaStream_t st, int* launches) {
    if (!(h->peer[0].attached || h->peer[1].attached)) return cudaSuccess;
    ++h->build_seq[slot];
    unsigned* my_flags = h->d_flag + (size_t)CONV_FLAG_BLOCK * slot;
    const size_t blk = (size_t)CONV_FLAG_BLOCK * slot;
    conv_begin_kernel<<<1, 1, 0, st>>>(my_flags + CONV_FLAG_EPOCH, h->peer[0].attached ? h->peer[0].flag + blk : nullptr,
                                       h->peer[1].attached ? h->peer[1].flag + blk : nullptr, h->octaves, h->nl, h->d_flag);
    ++*launches;
    return cudaGetLastError();
}

// Whole pyramid of one or more frame slots (no row bands: a banded handle is driven level by level so that the
// host can exchange halos between steps).  Octaves run concurrently: octave o+1 only depends on level S of octave o
// (its decimated base), so each octave gets its own stream, forked from the handle's stream by events and joined
// back at the end -- the small octaves' short kernels hide behind the large ones instead of queueing after them.
cudaError_t launch_conv(sspyr_ctx* h, int first, int count, int* launches, const ConvStreams& cs) {
    const int S = h->cfg.S;
    if (conv_begin_build(h, first, cs.main, launches) != cudaSuccess) return cudaGetLastError();
    const bool fork = h->tune.conv_streams != 0 && h->octaves > 1 && cs.aux != nullptr;
    cudaError_t e;
    if (!fork) {
        for (int o = 0; o < h->octaves; ++o)
            for (int s = 0; s < h->nl; ++s)
                if ((e = launch_conv_step(h, first, count, o, s, cs.main, launches, true)) != cudaSuccess) return e;
        return cudaSuccess;
    }
    for (int o = 0; o < h->octaves; ++o) {
        cudaStream_t st = o == 0 ? cs.main : cs.aux[o - 1];
        if (o > 0 && (e = cudaStreamWaitEvent(st, cs.ev_base[o - 1], 0)) != cudaSuccess) return e;   // base of octave o ready
        for (int s = 0; s < h->nl; ++s) {
            if ((e = launch_conv_step(h, first, count, o, s, st, launches, true)) != cudaSuccess) return e;
            if (s == S && o + 1 < h->octaves && (e = cudaEventRecord(cs.ev_base[o], st)) != cudaSuccess) return e;
        }
        if (o > 0) {                                            // join
            if ((e = cudaEventRecord(cs.ev_done[o - 1], st)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(cs.main, cs.ev_done[o - 1], 0)) != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

void conv_drop_graphs(sspyr_ctx* h) {
    for (auto& g : h->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
}

// launch_conv, replayed as a CUDA graph from the second build of the same slots on: 26+ small launches and
// their cross-stream events cost more host time than the small levels take on the GPU.
cudaError_t launch_conv_graphed(sspyr_ctx* h, int first, int count, int* launches, const ConvStreams& cs) {
    if (h->seg_dirty && h->d_seg) {          // segment counters possibly out of step: restart all of them from zero
        conv_drop_graphs(h);                 // (rare: after sspyr_set_tuning or a failed build; every lane has been
        cudaError_t me = cudaStreamSynchronize(h->stream);   //  joined into the handle's stream, so this drains them all)
        if (me == cudaSuccess) me = cudaMemset(h->d_seg, 0, sizeof(unsigned) * h->seg_frame_stride * h->cfg.frames);
        if (me != cudaSuccess) return me;
    }
    h->seg_dirty = false;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    // (row bands over peer memory included: their build number lives on the device, conv_begin_kernel)
    // Not when a neighbour band lives on the SAME GPU (single-GPU tests of the band protocol): kernels that wait on one
    // another must be able to run side by side, and two graphs launched on one device are not guaranteed to.
    const bool neighbour_here = (h->peer[0].attached && h->peer[0].same_device) || (h->peer[1].attached && h->peer[1].same_device);
    if (h->tune.conv_graph == 0 || neighbour_here || cudaStreamIsCapturing(cs.main, &st) != cudaSuccess ||
        st != cudaStreamCaptureStatusNone)
        return launch_conv(h, first, count, launches, cs);
    sspyr_ctx::GraphEntry* ge = nullptr;
    for (auto& g : h->graphs)
        if (g.first == first && g.count == count) ge = &g;
    if (!ge) {
        h->graphs.push_back({first, count, 0, 0, nullptr});
        ge = &h->graphs.back();
    }
    if (ge->exec) {
        *launches += ge->launches;
        return cudaGraphLaunch(ge->exec, cs.main);
    }
    if (ge->seen++ == 0) return launch_conv(h, first, count, launches, cs);      // first use: eager (sets kernel attributes)
    cudaError_t e = cudaStreamBeginCapture(cs.main, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { cudaGetLastError(); return launch_conv(h, first, count, launches, cs); }
    int n = 0;
    const cudaError_t le = launch_conv(h, first, count, &n, cs);
    cudaGraph_t graph = nullptr;
    e = cudaStreamEndCapture(cs.main, &graph);
    if (le != cudaSuccess || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        h->tune.conv_graph = 0;                                              // do not try again on this handle
        return launch_conv(h, first, count, launches, cs);
    }
    e = cudaGraphInstantiate(&ge->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        ge->exec = nullptr;
        cudaGetLastError();
        h->tune.conv_graph = 0;
        return launch_conv(h, first, count, launches, cs);
    }
    ge->launches = n;
    *launches += n;
    return cudaGraphLaunch(ge->exec, cs.main);
}

}  // namespace sspyr
