// csrc/conv_kernel.cuh -- CONV mode: one fused sm_100a kernel per scale-space level.  This file holds the
// shared definitions and the one-tile-per-CTA kernel (used for radii > 12 and on request); the default kernel for
// radii <= 12 is the marching strip kernel in conv_march.cuh, which does the same arithmetic in a better schedule.
//
// NOT in the reference (its "GaussFilter" is a pointwise window, GuassDePyramid.h:122-131); this is the
// separable Gaussian blur BASELINE.json's north_star describes, specified in DESIGN.md "CONV mode" and
// restated on the CPU in oracle/sspyr_oracle.c (orc_conv_build).  What it keeps from the reference: level
// and DoG counts (S+3 / S+2, :64,:140), DoG sign and slot order G_s - G_{s+1} (:143), octave sides H>>o, W>>o
// (:66), even-phase decimation (:80).
//
// One launch produces level s of one octave from level s-1 (or from the raw frame for octave 0, level 0):
//     tile + halo  --cp.async / ld.global-->  smem  --row pass-->  smem  --column pass-->  registers
//     epilogue:  G_s (kept in L2 for the next level),  DoG_{s-1} = G_{s-1} - G_s (streaming store),
//                and, for s == S, the even-phase 2x decimation into level 0 of the next octave
// so every level is read from HBM/L2 exactly once and written once; DoG and the next octave's base never
// cost a separate pass.
//
//   * CTA tile TW x TH = 128 x 32 outputs, 256 threads, 3-4 CTAs resident per SM so that one CTA's cp.async
//     staging overlaps its neighbours' filtering.  (Also built, selectable and measured slower: 64-row tiles;
//     persistent CTAs that double-buffer the staging tile.)
//   * Row pass: a thread owns 16 consecutive outputs of one row; lanes of a warp take consecutive ROWS and
//     the smem pitch is 4*odd floats, so its 128-bit smem loads/stores are bank-conflict free.
//   * Column pass: a thread owns 4 adjacent columns x (TH/8) rows and slides down the rows, holding the
//     accumulators in registers; lanes take consecutive column quads (conflict-free 128-bit loads).
//   * Taps are kernel parameters (constant bank), loops are fully unrolled on the compile-time radius, so the
//     inner loop is FFMA with a constant-bank operand.  Border: clamp to edge; for a row band the rows beyond
//     the band come from the neighbour's halo rows (top_halo / bot_halo) instead of the clamp.
//   * Tensor cores are deliberately not used: 2R+1 <= 21 taps on fp32 data with a 1e-4 budget is a stencil,
//     not a contraction (a banded-Toeplitz GEMM would waste > 80 % of the MMA and need 3xTF32 splitting).
#pragma once
#include <cuda_pipeline_primitives.h>

#include "conv_sched.h"
#include "sspyr_internal.h"

namespace sspyr {

constexpr int CONV_PX = 16;             // outputs per thread in the row pass
constexpr int CONV_THREADS = 256;

// input kinds of a level kernel
constexpr int CONV_FLAG_STRIDE = 32;      // progress-counter values per build (>= levels)
constexpr int CONV_FLAG_EPOCH = 51;       // d_flag[64 slot + 51]: builds of the slot started (row bands over peer memory)
constexpr int CONV_FLAG_TIMEOUT = 16;     // d_flag[0..15]: per-octave counters; d_flag[16]: wait-timeout marker (slot 0's block)
constexpr int CONV_SRC_PLANE = 3;       // float plane of the previous level (SSPYR_PIXEL_* = 0,1,2 are raw frames)

struct ConvParams {
    const void* src;                    // frame 0 of the launch: previous level plane, or the raw frame
    const void* top_halo;               // halo_rows rows just above this band (same pitch/type as src) or null
    const void* bot_halo;               // halo_rows rows just below this band, or null
    float* dst_g;                       // G_s
    float* dst_d;                       // DoG_{s-1} (null for s == 0)
    float* dst_dec;                     // level 0 of the next octave (null unless s == S and a next octave exists)
    unsigned long long src_frame_stride;   // elements between frames of a batched launch
    unsigned long long dst_frame_stride;   // floats
    int src_pitch, dst_pitch, dec_pitch;   // elements / floats
    int H, W, dec_H, dec_W;
    int halo_rows;
    // Peer-memory row bands, synchronisation fused into the strip kernel (null = not used):
    const unsigned* wait_up;            // neighbour-above's progress counter: CTAs that stage its rows wait for `wait_need`
    const unsigned* wait_dn;            // neighbour-below's
    unsigned wait_need;                 // (relative to the build: the kernel adds (*epoch - 1) * CONV_FLAG_STRIDE)
    unsigned* signal_flag;              // this band's counter for this octave: the last CTA to finish publishes signal_value
    unsigned signal_value;              // (relative, like wait_need)
    const unsigned* epoch;              // builds of this frame slot started so far, bumped on the device at the start of a build:
                                        // no launch parameter changes from build to build, so the sequence replays as a CUDA graph
    unsigned* done_count;               // CTAs finished so far (self-resetting)
    unsigned* timeout_mark;
    // Level chaining inside one octave (strip kernel, null = not used): every (strip, segment) CTA counts its builds
    // in seg_pub; with seg_dep set, a CTA waits for the up to 3x3 segments of the previous level it reads instead
    // of for the whole previous grid.
    unsigned* seg_pub;                  // this level's counters, frame 0 of the launch: [segment][strip]
    const unsigned* seg_dep;            // the previous level's counters (same geometry)
    unsigned seg_frame_stride;          // counters between consecutive frame slots
    // ... across a band seam (row bands over peer memory, chained levels): the previous level's counters of the NEIGHBOUR
    // bands, read through the peer mapping.  The first segment row of this band also waits for the neighbour-above's
    // segment rows [peer_up_first, peer_up_nsegs) (they hold its last R rows), the last one for the neighbour-below's row 0.
    const unsigned* peer_seg_up;
    const unsigned* peer_seg_dn;
    int peer_up_first, peer_up_nsegs;
    int seg_sys;                        // bit 0 / 1: a band above / below acquires the counters of this band's first / last segment rows
    int src_evict_first;                // strip kernel, TMA staging: load the source plane with an L2 evict-first policy
    float taps[2 * 32 + 1];             // taps[k + R], k = -R..R
};

template <int R> __host__ __device__ constexpr int conv_ra() { return (R + 3) / 4 * 4; }          // aligned halo
template <int R> __host__ __device__ constexpr int conv_pitch_in() {                            // 4 * odd
    int q = (CONV_TW + 2 * conv_ra<R>()) / 4;
    return 4 * (q | 1);
}
__host__ __device__ constexpr int conv_pitch_t() { return 4 * ((CONV_TW / 4) | 1); }             // 132
template <int R, int TH, int NBUF> __host__ __device__ constexpr size_t conv_smem_bytes() {   // staging buffer(s) + sT
    return sizeof(float) * (size_t)(TH + 2 * R) * (NBUF * conv_pitch_in<R>() + conv_pitch_t());
}

namespace {

// ---- packed fp32 (Blackwell FFMA2): two independent IEEE fma.rn per instruction -----------------------------
// d = {a.lo*b.lo + c.lo, a.hi*b.hi + c.hi}.  Each half rounds exactly like fmaf, so results are bit-identical to
// the scalar chains; a {w, w} multiplier built from one scalar is folded by ptxas into a broadcast operand
// (SASS: FFMA2 Rd, Ra.F32x2.HI_LO, URw.F32, Rc.F32x2.HI_LO), i.e. the taps cost no extra registers or moves.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int SRC>
__device__ __forceinline__ float load_src(const void* __restrict__ base, size_t idx) {
    if constexpr (SRC == SSPYR_PIXEL_I32) return (float)__ldg(static_cast<const int*>(base) + idx);
    else if constexpr (SRC == SSPYR_PIXEL_U8) return (float)__ldg(static_cast<const unsigned char*>(base) + idx);
    else if constexpr (SRC == CONV_SRC_PLANE) return __ldcg(static_cast<const float*>(base) + idx);   // written by a grid that may
    else return __ldg(static_cast<const float*>(base) + idx);                                         // still run (chained levels): no .nc
}

// Stage one tile (centre + halo) of frame fz into sIn: clamp to edge, or the neighbour's halo rows for a band.
// Interior tiles of a float plane go through cp.async (no registers, completion tracked by the pipeline);
// everything else is loaded, converted and stored synchronously.
template <int R, int SRC, int TH>
__device__ __forceinline__ void conv_stage_tile(const ConvParams& P, float* __restrict__ sIn, int x0, int y0, size_t fz,
                                                int tid) {
    constexpr int RA = conv_ra<R>();
    constexpr int ROWS = TH + 2 * R;
    constexpr int PIN = conv_pitch_in<R>();
    constexpr int COLS = CONV_TW + 2 * RA;
    constexpr int elem = SRC == SSPYR_PIXEL_U8 ? 1 : 4;
    const unsigned char* src = static_cast<const unsigned char*>(P.src) + fz * P.src_frame_stride * elem;
    const int warp = tid >> 5, lane = tid & 31;
    const bool fast_x = x0 - RA >= 0 && x0 + CONV_TW + RA <= P.W;   // whole staged span inside the row: 128-bit path
    for (int ly = warp; ly < ROWS; ly += CONV_THREADS / 32) {
        const int gy = y0 - R + ly;
        const unsigned char* row;
        if (gy < 0) {
            row = P.top_halo ? static_cast<const unsigned char*>(P.top_halo) + (size_t)max(P.halo_rows + gy, 0) * P.src_pitch * elem
                             : src;
        } else if (gy >= P.H) {
            row = P.bot_halo ? static_cast<const unsigned char*>(P.bot_halo) + (size_t)min(gy - P.H, P.halo_rows - 1) * P.src_pitch * elem
                             : src + (size_t)(P.H - 1) * P.src_pitch * elem;
        } else {
            row = src + (size_t)gy * P.src_pitch * elem;
        }
        float* srow = sIn + (size_t)ly * PIN;
        if (fast_x) {
            if constexpr (SRC == CONV_SRC_PLANE) {
                const float* grow = reinterpret_cast<const float*>(row) + (x0 - RA);
                for (int q = lane; q < COLS / 4; q += 32) __pipeline_memcpy_async(srow + 4 * q, grow + 4 * q, 16);
            } else {
                for (int q = lane; q < COLS / 4; q += 32) {
                    float4 v;
                    if constexpr (SRC == SSPYR_PIXEL_I32) {
                        const int4 t = __ldg(reinterpret_cast<const int4*>(reinterpret_cast<const int*>(row) + (x0 - RA)) + q);
                        v = make_float4((float)t.x, (float)t.y, (float)t.z, (float)t.w);
                    } else if constexpr (SRC == SSPYR_PIXEL_U8) {
                        const uchar4 t = __ldg(reinterpret_cast<const uchar4*>(row + (x0 - RA)) + q);
                        v = make_float4((float)t.x, (float)t.y, (float)t.z, (float)t.w);
                    } else {
                        v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(row) + (x0 - RA)) + q);
                    }
                    *reinterpret_cast<float4*>(srow + 4 * q) = v;
                }
            }
        } else {
            for (int lx = lane; lx < COLS; lx += 32) {
                const int gx = min(max(x0 - RA + lx, 0), P.W - 1);
                srow[lx] = load_src<SRC>(row, (size_t)gx);
            }
        }
    }
}

// CTAs walk the tile list (tile = blockIdx.x, += gridDim.x; x fastest, then y, then frame).
//   NBUF == 1: one tile per CTA (grid = ntiles), single staging buffer -- more CTAs resident per SM.
//   NBUF == 2: persistent CTAs with a two-deep software pipeline -- while tile i is filtered out of one staging
//              buffer the cp.async loads of tile i+1 land in the other (measured slower on B200: the second
//              buffer costs a resident CTA per SM, see profiles/).
template <int R, int SRC, int TH, int NBUF>
__global__ void __launch_bounds__(CONV_THREADS, 2)
conv_level_kernel(const __grid_constant__ ConvParams P, int tiles_x, int tiles_y, int ntiles) {
    constexpr int RA = conv_ra<R>();
    constexpr int ROWS = TH + 2 * R;              // rows of the staged tile
    constexpr int PIN = conv_pitch_in<R>();
    constexpr int PT = conv_pitch_t();
    constexpr int PY = TH / 8;                    // rows per thread in the column pass
    extern __shared__ __align__(16) float smem[];
    constexpr int BUF = ROWS * PIN;                         // [ROWS][PIN] x 2: input tile + halo (centre at column RA)
    float* sT = smem + (size_t)NBUF * ROWS * PIN;           // [ROWS][PT]: row-pass result

    const int tid = threadIdx.x;
    int tile = blockIdx.x;
    if (tile >= ntiles) return;
    auto coords = [&](int t, int& x0, int& y0, size_t& fz) {
        const int tx = t % tiles_x, r = t / tiles_x;
        x0 = tx * CONV_TW;
        y0 = (r % tiles_y) * TH;
        fz = (size_t)(r / tiles_y);
    };
    int x0, y0;
    size_t fz;
    coords(tile, x0, y0, fz);
    conv_stage_tile<R, SRC, TH>(P, smem, x0, y0, fz, tid);
    __pipeline_commit();
    int cur = 0;
    for (; tile < ntiles; tile += gridDim.x, cur ^= 1) {
        const int next = tile + gridDim.x;
        if constexpr (NBUF == 2) {
            if (next < ntiles) {                               // prefetch the next tile into the other buffer
                int nx, ny;
                size_t nf;
                coords(next, nx, ny, nf);
                conv_stage_tile<R, SRC, TH>(P, smem + (cur ^ 1) * BUF, nx, ny, nf, tid);
            }
            __pipeline_commit();
            __pipeline_wait_prior(1);                          // everything but the newest group has landed
        } else {
            __pipeline_wait_prior(0);
        }
        __syncthreads();
        const float* sIn = smem + (NBUF == 2 ? cur : 0) * BUF;
        coords(tile, x0, y0, fz);

        // ---- row pass: sT[row][c] = sum_k taps[k] * sIn[row][RA + c + k] ---------------------------------
        {
            constexpr int NB = CONV_TW / CONV_PX;                 // column blocks per row
            constexpr int NIN = CONV_PX + 2 * RA;                 // aligned input window per task
            for (int task = tid; task < ROWS * NB; task += CONV_THREADS) {
                const int row = task % ROWS, cb = task / ROWS;    // consecutive lanes -> consecutive rows
                const float4* in4 = reinterpret_cast<const float4*>(sIn + (size_t)row * PIN + cb * CONV_PX);
                float in[NIN];
#pragma unroll
                for (int q = 0; q < NIN / 4; ++q) {
                    const float4 v = in4[q];
                    in[4 * q] = v.x; in[4 * q + 1] = v.y; in[4 * q + 2] = v.z; in[4 * q + 3] = v.w;
                }
                float acc[CONV_PX];
#pragma unroll
                for (int i = 0; i < CONV_PX; ++i) acc[i] = 0.0f;
#pragma unroll
                for (int k = 0; k <= 2 * R; ++k) {                 // taps outer: 16 independent FMA chains in flight
                    const float w = P.taps[k];
#pragma unroll
                    for (int i = 0; i < CONV_PX; ++i) acc[i] = fmaf(w, in[i + k + (RA - R)], acc[i]);
                }
                float4* out4 = reinterpret_cast<float4*>(sT + (size_t)row * PT + cb * CONV_PX);
#pragma unroll
                for (int q = 0; q < CONV_PX / 4; ++q) out4[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
            }
        }
        __syncthreads();

        // ---- column pass + epilogue -------------------------------------------------------------------------
        {
            const int cq = tid & 31, rb = tid >> 5;              // column quad, row block
            const float* tcol = sT + (size_t)(rb * PY) * PT + cq * 4;
            float acc[PY][4];
#pragma unroll
            for (int j = 0; j < PY; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f; }
#pragma unroll
            for (int i = 0; i < PY + 2 * R; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(tcol + (size_t)i * PT);
#pragma unroll
                for (int j = 0; j < PY; ++j) {
                    if (i - j >= 0 && i - j <= 2 * R) {            // compile-time after unrolling
                        const float w = P.taps[i - j];
                        acc[j][0] = fmaf(w, v.x, acc[j][0]);
                        acc[j][1] = fmaf(w, v.y, acc[j][1]);
                        acc[j][2] = fmaf(w, v.z, acc[j][2]);
                        acc[j][3] = fmaf(w, v.w, acc[j][3]);
                    }
                }
            }
            const int x = x0 + cq * 4;
            if (x < P.W) {
                const int nvalid = P.W - x;
                float* g = P.dst_g + fz * P.dst_frame_stride;
                float* d = P.dst_d ? P.dst_d + fz * P.dst_frame_stride : nullptr;
                float* dec = P.dst_dec ? P.dst_dec + fz * P.dst_frame_stride : nullptr;
#pragma unroll
                for (int j = 0; j < PY; ++j) {
                    const int y = y0 + rb * PY + j;
                    if (y >= P.H) break;
                    const size_t o = (size_t)y * P.dst_pitch + x;
                    if (nvalid >= 4) {
                        *reinterpret_cast<float4*>(g + o) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) if (i < nvalid) g[o + i] = acc[j][i];
                    }
                    if (d) {                                        // DoG_{s-1} = G_{s-1} - G_s  (GuassDePyramid.h:143)
                        const float4 c = *reinterpret_cast<const float4*>(sIn + (size_t)(R + rb * PY + j) * PIN + RA + cq * 4);
                        const float dv[4] = {c.x - acc[j][0], c.y - acc[j][1], c.z - acc[j][2], c.w - acc[j][3]};
                        if (nvalid >= 4) {
                            __stcs(reinterpret_cast<float4*>(d + o), make_float4(dv[0], dv[1], dv[2], dv[3]));
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i) if (i < nvalid) __stcs(d + o + i, dv[i]);
                        }
                    }
                    if (dec && (y & 1) == 0) {                      // even-phase decimation (GuassDePyramid.h:80)
                        const int dy = y >> 1, dx = x >> 1;
                        if (dy < P.dec_H && dx < P.dec_W) {
                            float* q = dec + (size_t)dy * P.dec_pitch + dx;
                            if (dx + 1 < P.dec_W) *reinterpret_cast<float2*>(q) = make_float2(acc[j][0], acc[j][2]);
                            else q[0] = acc[j][0];
                        }
                    }
                }
            }
        }
        __syncthreads();        // sT and this staging buffer are free for the tile after next
        if constexpr (NBUF == 1) {
            if (next < ntiles) {
                coords(next, x0, y0, fz);
                conv_stage_tile<R, SRC, TH>(P, smem, x0, y0, fz, tid);
                __pipeline_commit();
            }
        }
    }
}

template <int R, int SRC, int TH, int NBUF>
cudaError_t launch_conv_one(const ConvParams& P, cudaStream_t st, int device, int frames, int max_ctas) {
    constexpr size_t smem = conv_smem_bytes<R, TH, NBUF>();
    static bool configured[64] = {false};         // the attribute is per device
    if (device < 0 || device >= 64 || !configured[device]) {
        cudaError_t e = cudaFuncSetAttribute(conv_level_kernel<R, SRC, TH, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (device >= 0 && device < 64) configured[device] = true;
    }
    const int tiles_x = (P.W + CONV_TW - 1) / CONV_TW, tiles_y = (P.H + TH - 1) / TH;
    const long long ntiles = (long long)tiles_x * tiles_y * frames;
    if (ntiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    long long ctas = ntiles;
    if (NBUF == 2) {    // persistent grid: as many CTAs as fit on the GPU at once (smem-limited), never more than tiles
        int per_sm = (int)((227 * 1024) / (smem + 1024));
        per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
        ctas = (long long)max_ctas * per_sm;
        if (ctas > ntiles) ctas = ntiles;
    }
    conv_level_kernel<R, SRC, TH, NBUF><<<(unsigned)ctas, CONV_THREADS, smem, st>>>(P, tiles_x, tiles_y, (int)ntiles);
    return cudaGetLastError();
}

// (64-row tiles and persistent double-buffered CTAs were built and measured slower in round 1 and are gone; the
//  NBUF template parameter of the kernel stays at 1.)
template <int R, int SRC>
cudaError_t launch_conv_th(const ConvParams& P, int /*variant*/, cudaStream_t st, int device, int frames, int sms) {
    return launch_conv_one<R, SRC, 32, 1>(P, st, device, frames, sms);
}

template <int R>
cudaError_t launch_conv_src(const ConvParams& P, int src_kind, int variant, cudaStream_t st, int device, int frames, int sms) {
    switch (src_kind) {
        case SSPYR_PIXEL_I32: return launch_conv_th<R, SSPYR_PIXEL_I32>(P, variant, st, device, frames, sms);
        case SSPYR_PIXEL_F32: return launch_conv_th<R, SSPYR_PIXEL_F32>(P, variant, st, device, frames, sms);
        case SSPYR_PIXEL_U8: return launch_conv_th<R, SSPYR_PIXEL_U8>(P, variant, st, device, frames, sms);
        default: return launch_conv_th<R, CONV_SRC_PLANE>(P, variant, st, device, frames, sms);
    }
}

}  // namespace

}  // namespace sspyr
