// csrc/sspyr_internal.h -- private to libsspyr.so (host state of one handle + kernel launch ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/sspyr.h"

// internal bits or-ed into RefParams::outputs (above the public SSPYR_OUT_* mask)
#define SSPYR_INT_INIT_ONLY (1 << 16)   // K0 only: G_s = p

namespace sspyr {

struct CascMaps;                          // conv_cascade.cuh: per-octave tensor maps of the cascade kernel

constexpr int CONV_FLAG_BLOCK = 512;      // CONV counters per frame slot: slot f uses d_flag[512 f ..): [0..15] per-octave progress,
                                          // [16] of slot 0: wait-timeout marker, [48..51] build epochs (cascade / bands),
                                          // [256 + 16 octave + level] finished-CTA counts of a level (levels of an octave overlap
                                          // when they are chained, so every level counts its own CTAs)
constexpr int CONV_FLAG_DONE = 256;

// ---- per-octave geometry of one frame slot -------------------------------------------------------
struct OctGeom {
    int H = 0, W = 0;          // H>>o, W>>o                      (GuassDePyramid.h:66)
    int pitch = 0;             // floats, multiple of 32 (128-byte rows)
    size_t plane = 0;          // H * pitch floats
    size_t off = 0;            // float offset of this octave inside a frame slot
    size_t fw_off = 0;         // REF: float offset of the column-window table [NL][pitch] in d_tables
    size_t fh_off = 0;         // REF: float offset of the row-window table    [H][8] (transposed)
    size_t ext_off = 0;        // byte offset of the extrema planes [S][H][pitch] inside an extrema slot
};

// ---- kernel parameter blocks (passed __grid_constant__) ---------------------------------------------
struct RefOct {
    float* base;               // [G_0..G_{S+1} | DoG_0..DoG_{S+1} | G_{S+2}] of frame 0 of the launch
    const float* fw;           // [NL][pitch]
    const float* fh;           // [H][8] ([H][16] when S+3 > 8): row window, transposed
    int H, W, pitch;
    unsigned long long plane;
};

struct RefParams {
    const void* img;           // frame 0 of the launch
    unsigned long long img_frame_stride;   // bytes between consecutive frames (batched launch)
    unsigned long long out_frame_stride;   // floats between consecutive frame slots
    int img_pitch;             // elements between input rows
    int H, W;                  // input rows (this band) / columns
    int octaves;
    int outputs;               // SSPYR_OUT_* mask
    RefOct oct[SSPYR_MAX_OCTAVES];
};

struct Tuning {
    int rows_per_thread = 0;   // 0 = default
    int block = 0;             // threads per CTA
    int bx = 0;                // CTA width in quads (threads along a row); 0 = pick the least-padding width
    int occ = 0;               // REF: > 0 = persistent grid of this many CTAs per SM (row-group prefetch); 0 = one group per thread
    int prefetch_next = -1;    // REF: after a build, pull the next slot's input toward L2: 1 on, 0 off, -1 = when the
                               // handle has several slots and a frame is small next to the 126 MB L2 (measured: +10 % on
                               // 1080p, +4 % on 4K, -4 % on 8K where the frame itself would flush L2)
    int pdl = -1;              // programmatic dependent launch between consecutive builds; -1 = default (on)
    int conv_march = 1;        // CONV: marching strip kernel for radii <= 12 (0 = one-tile-per-CTA kernel everywhere)
    int conv_tma = 1;          // CONV strip kernel: TMA (cp.async.bulk.tensor) staging of interior steps
    int conv_waves = 0;        // CONV strip kernel: CTA waves to aim for (0 = automatic: 3, or 8-step segments, see march_seg_rows)
    int conv_l2hint = 0;       // CONV chained levels: TMA-load the (dead after this level) source plane with L2 evict-first
                               // (measured SLOWER -- 8K 0.679 vs 0.667 ms, 16K 5.02 vs 4.68 ms: the halo re-reads of the
                               // neighbouring CTAs then miss -- so it is off)
    int conv_seg_min = 0;      // CONV strip kernel: minimum segment height in rows (0 = 32)
    int conv_fused_sync = 1;   // CONV peer bands: wait/signal inside the strip kernel (0 = one-thread kernels around it)
    int conv_graph = 1;        // CONV: replay the per-frame launch sequence as a CUDA graph from its 2nd use on
    int conv_streams = 1;      // CONV: run octaves on concurrent streams
    int conv_lanes = 8;        // CONV: builds of different frame slots in flight at once (stream sets, <= 16), 1 = one at a
                               // time (measured, 1080p: 1 lane 0.163, 4 lanes 0.080 (5 slots) / 0.062, 8 lanes 0.056 ms per frame)
    int conv_cascade = 0;      // CONV, whole frames: ONE launch per build, levels pipelined through L2 (conv_cascade.cuh):
                               // 1 = for frames of >= 4 Mpixel, 2 = always, 0 = never (one launch per level, conv_march.cuh,
                               // which is also the only path for row bands)
    int conv_casc_seg = 0;     // cascade: segment height in rows (0 = automatic, cascade_seg_rows)
    int conv_band_lanes = 6;   // CONV row bands over peer memory: builds of different slots in flight (<= frame slots; measured on
                               // 8 GPUs, 8K / 16K: 3 -> 2.9x / 6.1x, 6 -> 3.6x / 6.8x, 8 -> 3.7x / 6.7x of one GPU, profiles/r2_bands.md)
    int conv_band_chain = 0;   // CONV row bands over peer memory: chain levels across the band seam through the neighbours'
                               // segment counters (0, default = whole-level progress flags between all levels: measured
                               // 5-7 % faster on 2 GPUs, profiles/r2_bands.md)
    int conv_chain = 1;        // CONV strip kernel: consecutive levels of an octave overlap -- a level's CTA starts as soon
                               // as the segments of the previous level it reads are published (per-segment counters),
                               // instead of after the whole previous grid (0 = grid-wide dependency only)
    int timing = 0;            // bracket every build with CUDA events (sspyr_elapsed_ms); events between two
                               // launches stop them from overlapping, so this is off unless asked for
};

// CONV mode: one level step (see conv_kernels.cu)
struct ConvLevel {
    int radius = 0;
    size_t taps_off = 0;       // float offset in d_tables of taps[2R+1]
};

}  // namespace sspyr

// ---- the handle -----------------------------------------------------------------------------------
struct sspyr_ctx {
    sspyr_config cfg{};
    int octaves = 0, nl = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    sspyr::OctGeom oct[SSPYR_MAX_OCTAVES];
    size_t frame_floats = 0;                 // floats per output frame slot
    float* d_out = nullptr;                  // cfg.frames slots
    unsigned char* d_ext = nullptr;          // extrema flags (optional)
    size_t ext_frame_bytes = 0;
    unsigned char* d_kp = nullptr;           // keypoint lists (optional): per slot [count, capacity, 0, 0][int4 records]
    size_t kp_frame_bytes = 0;
    int kp_capacity = 0;
    unsigned char* d_in = nullptr;           // cfg.frames input slots
    size_t in_pitch_bytes = 0, in_frame_bytes = 0, elem_bytes = 4;
    std::vector<const void*> ext_in;         // per-slot external device input (nullptr = own slot)
    std::vector<size_t> ext_pitch;
    std::vector<char> built;
    float* d_tables = nullptr;
    std::vector<float> h_tables;
    std::vector<sspyr::ConvLevel> conv;      // per level
    float* d_halo = nullptr;                 // CONV row-band halo receive buffers: per octave [up|down][rmax][pitch]
    size_t halo_off[SSPYR_MAX_OCTAVES] = {0};
    unsigned char* d_halo_raw = nullptr;     // same for the raw frame (octave 0, level 0): [up|down][rmax][in_pitch]
    int halo_rmax = 0;
    // CONV row bands over peer memory (NVLink): the neighbours' planes are read IN the blur kernel instead of
    // being copied into d_halo; progress counters in peer memory order the steps (conv_launch.cu).
    struct Peer {
        bool attached = false, local = false;
        bool same_device = false;                // neighbour band lives on THIS GPU (tests): its kernels must be able to run
                                                 // beside ours, so the launch sequence is not replayed as a graph
        const float* out = nullptr;              // neighbour's d_out
        const unsigned char* in = nullptr;       // neighbour's d_in
        const unsigned* flag = nullptr;          // neighbour's progress counter
        const unsigned* seg = nullptr;           // neighbour's segment build counters (level chaining across the band seam)
        size_t seg_frame_stride = 0, seg_off[SSPYR_MAX_OCTAVES] = {0}, seg_cap[SSPYR_MAX_OCTAVES] = {0};
        void* ipc_out = nullptr;                 // cudaIpcOpenMemHandle results (closed in destroy)
        void* ipc_in = nullptr;
        int height = 0, H[SSPYR_MAX_OCTAVES] = {0};
        size_t off[SSPYR_MAX_OCTAVES] = {0}, plane[SSPYR_MAX_OCTAVES] = {0};
        size_t frame_floats = 0, in_frame_bytes = 0;
    } peer[2];                                   // [0] = band above, [1] = band below
    unsigned* d_flag = nullptr;                  // per-octave progress counters + timeout marker (inside d_out's allocation)
    // CONV level chaining: one build counter per (frame slot, octave, level, segment of a strip), see conv_march.cuh
    unsigned* d_seg = nullptr;                   // (inside d_out's allocation, behind d_flag)
    size_t seg_frame_stride = 0;                 // counters per frame slot
    size_t seg_off[SSPYR_MAX_OCTAVES] = {0};     // first counter of an octave inside a slot
    size_t seg_cap[SSPYR_MAX_OCTAVES] = {0};     // counters per level of that octave (strips x ceil(H/32))
    bool seg_dirty = false;                      // counters may be out of step (tuning changed, failed build): zero them first
    std::vector<unsigned> build_seq;             // per frame slot: builds started so far (all bands issue the same sequence)
    sspyr::CascMaps* casc_maps = nullptr;        // cascade kernel: tensor maps (created on first use) and which octaves have one
    bool casc_tma[SSPYR_MAX_OCTAVES] = {false};
    unsigned* d_casc_tab = nullptr;              // cascade kernel: block index -> work item (rebuilt when the segmentation changes)
    unsigned casc_items = 0;
    bool casc_keyed = false;                     // the row-keyed order passed its dependency check (else: octave-major)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    struct GraphEntry { int first, count, seen, launches; cudaGraphExec_t exec; };
    std::vector<GraphEntry> graphs;          // CONV: captured whole-pyramid launch sequences, by (first slot, count)
    std::vector<cudaStream_t> aux;           // CONV: one extra stream per octave >= 1
    std::vector<cudaEvent_t> ev_base, ev_done;
    // CONV frame lanes (unbanded handles with several frame slots): a build runs on the stream set of lane
    // (first slot % lanes) and is joined back into the handle's stream at once, so everything the caller enqueues
    // afterwards still sees it complete -- but the NEXT build, on another lane, only waits for the library's own
    // earlier operations on the handle's stream (uploads, downloads, extrema: ev_tail), not for this one.  A pyramid
    // is a chain of 18+ dependent level kernels; several frames in flight hide that latency (1080p: 3-4x).
    struct Lane {
        cudaStream_t main = nullptr;
        std::vector<cudaStream_t> aux;
        std::vector<cudaEvent_t> ev_base, ev_done;
        cudaEvent_t done = nullptr;          // end of the lane's latest build
        unsigned seen_tail = 0;              // tail_seq the lane has already waited for
    };
    std::vector<Lane> lanes;
    std::vector<int> slot_lane;              // lane of the latest build that wrote a slot (-1: none)
    cudaEvent_t ev_tail = nullptr;           // latest library operation on the handle's stream that a build must follow
    unsigned tail_seq = 0;
    bool strict_order = false;               // raw device pointers were handed out / taken in: every build follows the
                                             // whole stream (the caller's own kernels may read or write the slots)
    bool timed = false;
    int last_launches = 0;
    int last_first = 0, last_count = 0;      // frame slots the previous kernel wrote (PDL overlap guard)
    sspyr::Tuning tune;
    std::string err;
};

namespace sspyr {

// Launchers (defined in the kernel translation units).  Return cudaError_t; *launches += kernels enqueued.
cudaError_t launch_ref(sspyr_ctx* h, int first_frame, int count, int outputs, int* launches);
// the stream set a whole-pyramid CONV build runs on: the handle's own, or one of its lanes
struct ConvStreams {
    cudaStream_t main;
    cudaStream_t* aux;           // octaves - 1 streams (or null)
    cudaEvent_t* ev_base;
    cudaEvent_t* ev_done;
};
cudaError_t launch_conv(sspyr_ctx* h, int first_frame, int count, int* launches, const ConvStreams& cs);
cudaError_t conv_begin_build(sspyr_ctx* h, int slot, cudaStream_t st, int* launches);
cudaError_t launch_conv_graphed(sspyr_ctx* h, int first_frame, int count, int* launches, const ConvStreams& cs);
void conv_drop_graphs(sspyr_ctx* h);
cudaError_t launch_conv_step(const sspyr_ctx* h, int first_frame, int count, int octave, int level, cudaStream_t st,
                             int* launches, bool chain = false);
bool conv_has_up(const sspyr_ctx* h);
bool conv_has_down(const sspyr_ctx* h);
float* conv_halo_plane(const sspyr_ctx* h, int octave, int down);
unsigned char* conv_halo_raw(const sspyr_ctx* h, int down);
cudaError_t launch_extrema(const sspyr_ctx* h, int first_frame, int count, int* launches);
bool conv_cascade_ok(const sspyr_ctx* h);
cudaError_t launch_conv_cascade(sspyr_ctx* h, int first_frame, int count, int* launches);
void conv_cascade_free(sspyr_ctx* h);

inline const unsigned char* frame_input(const sspyr_ctx* h, int frame, size_t* pitch_bytes) {
    if (h->ext_in[frame]) {
        *pitch_bytes = h->ext_pitch[frame];
        return static_cast<const unsigned char*>(h->ext_in[frame]);
    }
    *pitch_bytes = h->in_pitch_bytes;
    return h->d_in + (size_t)frame * h->in_frame_bytes;
}

inline float* frame_out(const sspyr_ctx* h, int frame) { return h->d_out + (size_t)frame * h->frame_floats; }

// index of a plane inside an octave block [G_0..G_{S+1} | DoG_0..DoG_{S+1} | G_{S+2}]
inline int plane_index(int nl, int kind, int level) {
    if (kind == SSPYR_KIND_GAUSS) return level == nl - 1 ? 2 * nl - 2 : level;
    if (kind == SSPYR_KIND_DOG) return nl - 1 + level;
    /* INPLACE */ return nl - 1 + level;   // slots 0..S+1 = DoG, slot S+2 = G_{S+2} (contiguous tail)
}

}  // namespace sspyr
