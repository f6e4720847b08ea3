// csrc/extrema.cu -- 26-neighbour DoG extremum scan with keypoint compaction (the next SIFT step
// after the pyramid; the reference stops at DoG, GuassDePyramid.h:136-149).  Both modes.
//
// One launch per frame covers every octave.  A CTA owns a 64 x 16 pixel tile and walks up the S+2 DoG planes of its
// octave with a four-plane ring in shared memory filled by cp.async: the tile (+ 1-pixel halo) of plane s+2 is in flight
// while level s is scanned out of planes s-1, s, s+1, so every DoG plane is read from HBM/L2 ONCE (x1.16 halo, 16-byte
// copies on the aligned interior) and the loads hide behind the comparisons (the first version loaded, synchronised and
// scanned in turn: long-scoreboard and barrier stalls, 1.7 TB/s; profiles/r2_extrema_c2_ncu_full.txt).
// Results, either or both:
//   * SSPYR_OUT_KEYPOINTS: a compacted list of (x, y, octave, level, value) records per frame slot -- warp-aggregated
//     reservation (one atomicAdd per warp and row) behind a per-slot cursor.  A 1080p pyramid is 66 MB of planes; its
//     keypoints are kilobytes: this is what takes the end-to-end path off the PCIe link.
//   * SSPYR_OUT_EXTREMA: one flag byte per pixel and level (the round-1 output, kept for the tests' exact comparison
//     with the oracle scan).
#include <cuda_pipeline_primitives.h>

#include "sspyr_internal.h"

namespace sspyr {

namespace {

constexpr int EXT_TW = 64, EXT_TH = 16, EXT_THREADS = 256;
constexpr int EXT_PITCH = EXT_TW + 8;        // smem row: [3 unused | left halo | 64 | right halo | 3 unused], interior 16-byte aligned
constexpr int EXT_ROWS = EXT_TH + 2;

struct ExtOct {
    const float* dog;                         // DoG_0 of frame slot 0 of the launch
    unsigned char* flags;                     // [S][H][pitch] or null
    unsigned long long plane;
    int H, W, pitch;
    int tiles_x;
    unsigned tile_base;                       // first block of this octave
};

struct ExtParams {
    ExtOct oct[SSPYR_MAX_OCTAVES];
    unsigned long long dog_frame_stride;      // floats
    unsigned long long flag_frame_stride;     // bytes
    unsigned long long kp_frame_stride;       // bytes between the keypoint buffers of consecutive slots
    unsigned char* kp;                        // slot 0 of the launch: [count, capacity, 0, 0][records...] or null
    unsigned tiles_per_frame;
    int octaves, S, capacity;
    float thresh;
};

// plane tile (+ halo, coordinates clamped into the plane: border pixels are never tested, their halo is never used),
// asynchronously: 16-byte cp.async on the aligned interior of full-width tiles, 4-byte cp.async for the rest
__device__ __forceinline__ void load_tile(float* __restrict__ s, const float* __restrict__ plane, int H, int W, int pitch,
                                          int x0, int y0, int tid) {
    const bool full = x0 + EXT_TW <= W;
    for (int i = tid; i < EXT_ROWS * (EXT_TW / 4); i += EXT_THREADS) {
        const int r = i / (EXT_TW / 4), q = i - r * (EXT_TW / 4);
        const int gy = min(max(y0 - 1 + r, 0), H - 1);
        float* d = s + r * EXT_PITCH + 4 + 4 * q;
        if (full) {
            __pipeline_memcpy_async(d, plane + (size_t)gy * pitch + x0 + 4 * q, 16);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) __pipeline_memcpy_async(d + k, plane + (size_t)gy * pitch + min(x0 + 4 * q + k, W - 1), 4);
        }
    }
    for (int i = tid; i < EXT_ROWS * 2; i += EXT_THREADS) {       // the two halo columns
        const int r = i >> 1, side = i & 1;
        const int gy = min(max(y0 - 1 + r, 0), H - 1);
        const int gx = side ? min(x0 + EXT_TW, W - 1) : max(x0 - 1, 0);
        __pipeline_memcpy_async(s + r * EXT_PITCH + (side ? 4 + EXT_TW : 3), plane + (size_t)gy * pitch + gx, 4);
    }
}

__global__ void __launch_bounds__(EXT_THREADS)
extrema_tile_kernel(const __grid_constant__ ExtParams P) {
    __shared__ __align__(16) float ring[4][EXT_ROWS * EXT_PITCH];
    const unsigned fz = blockIdx.x / P.tiles_per_frame;
    unsigned t = blockIdx.x - fz * P.tiles_per_frame;
    int o = 0;
    while (o + 1 < P.octaves && t >= P.oct[o + 1].tile_base) ++o;
    const ExtOct& O = P.oct[o];
    t -= O.tile_base;
    const int x0 = (int)(t % (unsigned)O.tiles_x) * EXT_TW, y0 = (int)(t / (unsigned)O.tiles_x) * EXT_TH;
    const int tid = threadIdx.x, lane = tid & 31;
    const int cx = tid & (EXT_TW - 1), ry = (tid >> 6) * 4;      // this thread: column cx, rows ry .. ry+3 of the tile
    const float* dog = O.dog + (size_t)fz * P.dog_frame_stride;
    unsigned char* flags = O.flags ? O.flags + (size_t)fz * P.flag_frame_stride : nullptr;
    unsigned* kp_head = P.kp ? reinterpret_cast<unsigned*>(P.kp + (size_t)fz * P.kp_frame_stride) : nullptr;
    int4* kp_rec = kp_head ? reinterpret_cast<int4*>(kp_head + 4) : nullptr;

    // planes 0, 1, 2 start moving now (one cp.async group each); inside the loop plane s+2 is issued before level s is
    // scanned and the wait leaves that newest group in flight
    for (int q = 0; q < 3 && q <= P.S + 1; ++q) {
        load_tile(ring[q], dog + (size_t)q * O.plane, O.H, O.W, O.pitch, x0, y0, tid);
        __pipeline_commit();
    }
    for (int s = 1; s <= P.S; ++s) {
        if (s + 2 <= P.S + 1) load_tile(ring[(s + 2) & 3], dog + (size_t)(s + 2) * O.plane, O.H, O.W, O.pitch, x0, y0, tid);
        __pipeline_commit();                                      // (an empty group when there is no plane s+2: keeps the count)
        __pipeline_wait_prior(1);                                 // planes <= s+1 have landed
        __syncthreads();
        const float* lo = ring[(s - 1) & 3];
        const float* mid = ring[s & 3];
        const float* hi = ring[(s + 1) & 3];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = ry + j, gy = y0 + r, gx = x0 + cx;
            bool found = false;
            float v = 0.0f;
            if (gy >= 1 && gy < O.H - 1 && gx >= 1 && gx < O.W - 1) {
                const int c = (r + 1) * EXT_PITCH + 4 + cx;       // centre in the tile
                v = mid[c];
                if (fabsf(v) > P.thresh) {
                    bool is_max = true, is_min = true;
#pragma unroll
                    for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
                        for (int dc = -1; dc <= 1; ++dc) {
                            const int k = c + dr * EXT_PITCH + dc;
                            if (dr != 0 || dc != 0) { const float n = mid[k]; is_max &= v > n; is_min &= v < n; }
                        }
                    if (is_max || is_min) {
#pragma unroll
                        for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
                            for (int dc = -1; dc <= 1; ++dc) {
                                const int k = c + dr * EXT_PITCH + dc;
                                const float a = lo[k], b = hi[k];
                                is_max &= v > a && v > b;
                                is_min &= v < a && v < b;
                            }
                        found = is_max || is_min;
                    }
                }
            }
            if (flags && gy < O.H && gx < O.W) flags[(size_t)(s - 1) * O.plane + (size_t)gy * O.pitch + gx] = found ? 1 : 0;
            if (kp_head) {                                        // warp-aggregated append
                const unsigned m = __ballot_sync(0xffffffffu, found);
                if (m) {
                    unsigned base = 0;
                    if (lane == __ffs(m) - 1) base = atomicAdd(kp_head, (unsigned)__popc(m));
                    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                    if (found) {
                        const unsigned idx = base + (unsigned)__popc(m & ((1u << lane) - 1u));
                        if (idx < (unsigned)P.capacity) kp_rec[idx] = make_int4(gx, gy, (o << 16) | s, __float_as_int(v));
                    }
                }
            }
        }
        __syncthreads();                                          // ring slot (s - 1) & 3 receives plane s + 3 in the next iteration
    }
}

}  // namespace

// Scan frame slots first .. first+count-1 (contiguous).  Clears the keypoint cursors first.
cudaError_t launch_extrema(const sspyr_ctx* h, int first, int count, int* launches) {
    if (h->cfg.S < 1 || count < 1) return cudaSuccess;
    ExtParams P{};
    unsigned tiles = 0;
    for (int o = 0; o < h->octaves; ++o) {
        const OctGeom& g = h->oct[o];
        ExtOct& O = P.oct[o];
        O.dog = frame_out(h, first) + g.off + (size_t)(h->nl - 1) * g.plane;
        O.flags = h->d_ext ? h->d_ext + (size_t)first * h->ext_frame_bytes + g.ext_off : nullptr;
        O.plane = g.plane;
        O.H = g.H; O.W = g.W; O.pitch = g.pitch;
        O.tiles_x = (g.W + EXT_TW - 1) / EXT_TW;
        O.tile_base = tiles;
        tiles += (unsigned)O.tiles_x * (unsigned)((g.H + EXT_TH - 1) / EXT_TH);
    }
    P.dog_frame_stride = h->frame_floats;
    P.flag_frame_stride = h->ext_frame_bytes;
    P.kp_frame_stride = h->kp_frame_bytes;
    P.kp = h->d_kp ? h->d_kp + (size_t)first * h->kp_frame_bytes : nullptr;
    P.tiles_per_frame = tiles;
    P.octaves = h->octaves;
    P.S = h->cfg.S;
    P.capacity = h->kp_capacity;
    P.thresh = h->cfg.extrema_thresh;
    if ((unsigned long long)tiles * count >= 0x7fffffffULL) return cudaErrorInvalidValue;
    if (h->d_kp)
        for (int f = 0; f < count; ++f) {
            const cudaError_t e = cudaMemsetAsync(h->d_kp + (size_t)(first + f) * h->kp_frame_bytes, 0, 4, h->stream);
            if (e != cudaSuccess) return e;
        }
    extrema_tile_kernel<<<tiles * (unsigned)count, EXT_THREADS, 0, h->stream>>>(P);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace sspyr
