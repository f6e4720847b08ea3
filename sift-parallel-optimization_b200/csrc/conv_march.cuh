// csrc/conv_march.cuh -- CONV mode, large levels: column-strip "marching" kernel.
//
// Same arithmetic as conv_kernel.cuh (same taps, same row-then-column order, same fp32 FMA chains, so the
// results are bit-identical), different schedule.  The tile kernel recomputes the row pass for 2R halo rows of
// every 32-row tile and moves every intermediate through shared memory twice; for big levels that makes it
// shared-memory- and FMA-bound below the HBM roofline.  Here a thread owns 4 adjacent output columns and
// marches DOWN a vertical segment of the image:
//
//   per input row:  128-bit loads of its 4+2R inputs from a shared-memory ring   (row pass, registers only)
//                   4 x (2R+1) FMAs scatter the row-pass value into a ring of 2R+1 partial output rows held in
//                   REGISTERS (column pass, no shared memory, no recompute)
//                   the output row that just completed is stored: G_s, DoG_{s-1} = centre - G_s, decimated base
//
// so shared memory carries each input value once, FMAs are the minimum 2(2R+1) per pixel, and the only
// redundancy is the 2R warm-up rows at the top of each segment.  The row loop is unrolled 2R+1 times so that
// the register ring is indexed statically.  Input rows arrive through a cp.async ring of DEPTH batches of 8
// rows (two batches in flight ahead of the one being filtered); the ring also keeps the BACK batches that
// still hold the centre rows DoG needs.  A CTA is 4 warps = 512 columns; segments are sized so that the whole
// grid is co-resident (one wave).  Used for R <= 12 when the level is large enough to fill the GPU this way;
// everything else goes to the tile kernel.
#pragma once
#include "conv_kernel.cuh"

namespace sspyr {

constexpr int MARCH_SW = 512;            // strip width per CTA (4 warps x 32 quads x 4 columns)
constexpr int MARCH_THREADS = 128;
constexpr int MARCH_NB = 8;              // rows per staging batch

template <int R> __host__ __device__ constexpr int march_depth() { return (R + MARCH_NB - 1) / MARCH_NB + 3; }
template <int R> __host__ __device__ constexpr int march_pitch() { return MARCH_SW + 2 * conv_ra<R>(); }
template <int R> __host__ __device__ constexpr size_t march_smem_bytes() {
    return sizeof(float) * (size_t)march_depth<R>() * MARCH_NB * march_pitch<R>();
}

namespace {

template <int R, int SRC>
__device__ __forceinline__ void march_stage_batch(const ConvParams& P, float* __restrict__ ring, int b, int total,
                                                  int y_begin, int x0, size_t fz, int tid) {
    constexpr int RA = conv_ra<R>();
    constexpr int PITCH = march_pitch<R>();
    constexpr int CH = PITCH / 4;                            // 16-byte chunks per staged row
    constexpr int DEPTH = march_depth<R>();
    constexpr int elem = SRC == SSPYR_PIXEL_U8 ? 1 : 4;
    const unsigned char* src = static_cast<const unsigned char*>(P.src) + fz * P.src_frame_stride * elem;
    const int n0 = b * MARCH_NB;
    float* dst = ring + (size_t)((b % DEPTH) * MARCH_NB) * PITCH;
    for (int c = tid; c < MARCH_NB * CH; c += MARCH_THREADS) {
        const int rr = c / CH, q = c - rr * CH;
        const int n = n0 + rr;
        if (n >= total) break;
        const int gy = y_begin - R + n;
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
        float* s = dst + (size_t)rr * PITCH + 4 * q;
        const int gx = x0 - RA + 4 * q;
        if (gx >= 0 && gx + 4 <= P.W) {
            if constexpr (SRC == CONV_SRC_PLANE) {
                __pipeline_memcpy_async(s, reinterpret_cast<const float*>(row) + gx, 16);
            } else if constexpr (SRC == SSPYR_PIXEL_I32) {
                const int4 t = __ldg(reinterpret_cast<const int4*>(reinterpret_cast<const int*>(row) + gx));
                *reinterpret_cast<float4*>(s) = make_float4((float)t.x, (float)t.y, (float)t.z, (float)t.w);
            } else if constexpr (SRC == SSPYR_PIXEL_U8) {
                const uchar4 t = __ldg(reinterpret_cast<const uchar4*>(row + gx));
                *reinterpret_cast<float4*>(s) = make_float4((float)t.x, (float)t.y, (float)t.z, (float)t.w);
            } else {
                *reinterpret_cast<float4*>(s) = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(row) + gx));
            }
        } else {                                             // chunk straddles the image edge: clamp per element
#pragma unroll
            for (int i = 0; i < 4; ++i)
                s[i] = load_src<SRC == CONV_SRC_PLANE ? SSPYR_PIXEL_F32 : SRC>(row, (size_t)min(max(gx + i, 0), P.W - 1));
        }
    }
}

// grid: x = 512-column strips, y = vertical segments of seg_rows output rows, z = frame
template <int R, int SRC>
__global__ void __launch_bounds__(MARCH_THREADS)
conv_march_kernel(const __grid_constant__ ConvParams P, int seg_rows) {
    constexpr int U = 2 * R + 1;
    constexpr int RA = conv_ra<R>();
    constexpr int PITCH = march_pitch<R>();
    constexpr int DEPTH = march_depth<R>();
    constexpr int NIN = 4 + 2 * RA;
    extern __shared__ __align__(16) float ring[];

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * MARCH_SW;
    const int y_begin = blockIdx.y * seg_rows;
    const size_t fz = blockIdx.z;
    if (y_begin >= P.H) return;
    const int y_end = min(P.H, y_begin + seg_rows);
    const int total = (y_end - y_begin) + 2 * R;             // input rows n = 0..total-1  <->  gy = y_begin - R + n
    const int nbatches = (total + MARCH_NB - 1) / MARCH_NB;

    march_stage_batch<R, SRC>(P, ring, 0, total, y_begin, x0, fz, tid);
    __pipeline_commit();
    if (nbatches > 1) march_stage_batch<R, SRC>(P, ring, 1, total, y_begin, x0, fz, tid);
    __pipeline_commit();

    const int x = x0 + 4 * tid;                              // this thread's 4 output columns
    const int nvalid = P.W - x;                              // <= 0: staging helper only
    float* g = P.dst_g + fz * P.dst_frame_stride;
    float* d = P.dst_d ? P.dst_d + fz * P.dst_frame_stride : nullptr;
    float* dec = P.dst_dec ? P.dst_dec + fz * P.dst_frame_stride : nullptr;

    float acc[U][4];
#pragma unroll
    for (int j = 0; j < U; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;

    for (int base = 0; base < total; base += U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int n = base + u;
            if (n >= total) break;
            if ((n % MARCH_NB) == 0) {                       // batch boundary (uniform)
                const int b = n / MARCH_NB;
                __pipeline_wait_prior(1);                    // batches 0..b have landed (b+1 may be in flight)
                __syncthreads();                             // ... for every thread, and batch b-1 is fully consumed
                if (b + 2 < nbatches) march_stage_batch<R, SRC>(P, ring, b + 2, total, y_begin, x0, fz, tid);
                __pipeline_commit();
            }
            // ---- row pass: t[i] = sum_k taps[k] * in[x + i + k - R] ----
            const float* srow = ring + (size_t)(((n / MARCH_NB) % DEPTH) * MARCH_NB + (n % MARCH_NB)) * PITCH + 4 * tid;
            float in[NIN];
#pragma unroll
            for (int q = 0; q < NIN / 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(srow + 4 * q);
                in[4 * q] = v.x; in[4 * q + 1] = v.y; in[4 * q + 2] = v.z; in[4 * q + 3] = v.w;
            }
            float t[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int k = 0; k <= 2 * R; ++k) {
                const float w = P.taps[k];
#pragma unroll
                for (int i = 0; i < 4; ++i) t[i] = fmaf(w, in[i + k + (RA - R)], t[i]);
            }
            // ---- column pass: out(j) += taps[k] * T(n) for j = n - k; slot(j) = j mod U is static here ----
            // (k descending so that each output accumulates its rows top to bottom, as the tile kernel does)
#pragma unroll
            for (int k = 2 * R; k >= 0; --k) {
                constexpr int dummy = 0; (void)dummy;
                const int slot = ((u - k) % U + U) % U;
                const float w = P.taps[k];
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[slot][i] = fmaf(w, t[i], acc[slot][i]);
            }
            // ---- output row j = n - 2R is complete: slot (u + 1) % U ----
            {
                constexpr int dummy2 = 0; (void)dummy2;
                const int slot = (u + 1) % U;
                const int j = n - 2 * R;
                if (j >= 0 && nvalid > 0) {
                    const int y = y_begin + j;
                    const size_t o = (size_t)y * P.dst_pitch + x;
                    if (nvalid >= 4) {
                        *reinterpret_cast<float4*>(g + o) = make_float4(acc[slot][0], acc[slot][1], acc[slot][2], acc[slot][3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) if (i < nvalid) g[o + i] = acc[slot][i];
                    }
                    if (d) {                                 // DoG_{s-1} = G_{s-1} - G_s; centre row = input row n - R
                        const int nc = n - R;
                        const float4 c = *reinterpret_cast<const float4*>(
                            ring + (size_t)(((nc / MARCH_NB) % DEPTH) * MARCH_NB + (nc % MARCH_NB)) * PITCH + RA + 4 * tid);
                        const float dv[4] = {c.x - acc[slot][0], c.y - acc[slot][1], c.z - acc[slot][2], c.w - acc[slot][3]};
                        if (nvalid >= 4) {
                            __stcs(reinterpret_cast<float4*>(d + o), make_float4(dv[0], dv[1], dv[2], dv[3]));
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i) if (i < nvalid) __stcs(d + o + i, dv[i]);
                        }
                    }
                    if (dec && (y & 1) == 0) {               // even-phase decimation (GuassDePyramid.h:80)
                        const int dy = y >> 1, dx = x >> 1;
                        if (dy < P.dec_H && dx < P.dec_W) {
                            float* q = dec + (size_t)dy * P.dec_pitch + dx;
                            if (dx + 1 < P.dec_W) *reinterpret_cast<float2*>(q) = make_float2(acc[slot][0], acc[slot][2]);
                            else q[0] = acc[slot][0];
                        }
                    }
                }
                acc[slot][0] = acc[slot][1] = acc[slot][2] = acc[slot][3] = 0.0f;
            }
        }
    }
}

template <int R, int SRC>
cudaError_t launch_march_one(const ConvParams& P, cudaStream_t st, int device, int frames, int sms) {
    constexpr size_t smem = march_smem_bytes<R>();
    static bool configured[64] = {false};
    if (device < 0 || device >= 64 || !configured[device]) {
        cudaError_t e = cudaFuncSetAttribute(conv_march_kernel<R, SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (device >= 0 && device < 64) configured[device] = true;
    }
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
    const int strips = (P.W + MARCH_SW - 1) / MARCH_SW;
    // one co-resident wave: as many vertical segments as fit, each at least 32 rows
    long long segs = (long long)sms * per_sm / ((long long)strips * frames);
    if (segs < 1) segs = 1;
    int seg_rows = (int)((P.H + segs - 1) / segs);
    if (seg_rows < 32) seg_rows = 32;
    seg_rows = (seg_rows + 1) & ~1;                           // even: decimation rows stay aligned with segments
    const int nseg = (P.H + seg_rows - 1) / seg_rows;
    const dim3 grid(strips, nseg, frames);
    conv_march_kernel<R, SRC><<<grid, MARCH_THREADS, smem, st>>>(P, seg_rows);
    return cudaGetLastError();
}

template <int R>
cudaError_t launch_march_src(const ConvParams& P, int src_kind, cudaStream_t st, int device, int frames, int sms) {
    switch (src_kind) {
        case SSPYR_PIXEL_I32: return launch_march_one<R, SSPYR_PIXEL_I32>(P, st, device, frames, sms);
        case SSPYR_PIXEL_F32: return launch_march_one<R, SSPYR_PIXEL_F32>(P, st, device, frames, sms);
        case SSPYR_PIXEL_U8: return launch_march_one<R, SSPYR_PIXEL_U8>(P, st, device, frames, sms);
        default: return launch_march_one<R, CONV_SRC_PLANE>(P, st, device, frames, sms);
    }
}

}  // namespace

}  // namespace sspyr
