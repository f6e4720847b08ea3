// csrc/ref_kernel.cuh -- REF mode: the reference's whole pipeline in ONE fused sm_100a kernel.
//
// Reference stages replaced (ZhangShuui/SIFT-parallel-optimization, GuassDePyramid.h):
//   K0  GaussPyInit      :76-86    level[o][s][r][c] = (float) data[r<<o][c<<o]     (decimate from the ORIGINAL)
//   K2  GaussFilter rows :122-126  L[r][c] *= f_s,o[c]
//   K3  GaussFilter cols :127-131  L[r][c] *= f_s,o[r]
//   K4  GenerateDoG      :140-146  L[s] -= L[s+1]
// Closed form per octave o, level s:  G_s(r,c) = ((p * fW[c]) * fH[r]),  DoG_s = G_s - G_{s+1},
// with p = (float) img[r<<o][c<<o].  Two separately rounded multiplies then one subtract, in the
// reference's order -- __fmul_rn/__fsub_rn keep nvcc from contracting them into FMAs, so the result is
// bit-identical to the CPU header (window tables are computed on the host with the header's own libm
// expression, K1 :118-121, see sspyr_api.cu).
//
// Design (HBM-bound, write-dominated: 1 plane read, 2S+5 planes written per octave):
//   * Tile-owner mapping: the thread that loads input quad (r, 4j..4j+3) emits EVERY octave that pixel
//     quad feeds -- octave 0 always; octave o when r % 2^o == 0 (columns 4j and 4j+2 for o=1, column 4j
//     for o>=2 when j % 2^(o-2) == 0).  Each input pixel is therefore read from HBM exactly once and no
//     level is ever materialised before its final value (the reference writes every level 3 times).
//   * 128-bit coalesced loads (ld.global.nc) and streaming stores (st.global.cs): a warp writes 512
//     contiguous bytes per plane per row; rows are 128-byte aligned (pitch % 32 == 0).
//   * Column-window values for octave 0 live in registers across the rows a thread walks; row-window
//     values are warp-uniform broadcast loads.
//   * No shared memory, no tensor cores: there is no reuse and 0.3 flop/byte.
#pragma once
#include "sspyr_internal.h"

namespace sspyr {

namespace {

template <int N> struct Vec;
template <> struct Vec<1> { using T = float; };
template <> struct Vec<2> { using T = float2; };
template <> struct Vec<4> { using T = float4; };

template <int N>
__device__ __forceinline__ void load_tab(float (&w)[N], const float* __restrict__ p) {
    if constexpr (N == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else if constexpr (N == 2) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(p));
        w[0] = v.x; w[1] = v.y;
    } else {
        w[0] = __ldg(p);
    }
}

// Streaming (evict-first) store of N contiguous floats, the first `nvalid` of which exist.
template <int N>
__device__ __forceinline__ void store_out(float* __restrict__ p, const float (&v)[N], int nvalid) {
    if (nvalid >= N) {
        if constexpr (N == 4) __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
        else if constexpr (N == 2) __stcs(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
        else __stcs(p, v[0]);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (i < nvalid) __stcs(p + i, v[i]);
    }
}

// All S+3 levels and S+2 DoGs of N horizontally adjacent pixels of one octave.
//   w[s][i] : column window of level s at output column ocol+i   (registers)
template <int NL, int N>
__device__ __forceinline__ void emit_levels(const RefOct& oc, int outputs, int orow, int ocol,
                                            const float (&p)[N], const float (&w)[NL][N]) {
    const int nvalid = oc.W - ocol;
    float* __restrict__ out = oc.base + (size_t)orow * oc.pitch + ocol;
    const float* __restrict__ fh = oc.fh + orow;
    const bool want_g = outputs & SSPYR_OUT_GAUSS;
    const bool want_top = outputs & (SSPYR_OUT_GAUSS | SSPYR_OUT_GAUSS_TOP);
    const bool want_d = outputs & SSPYR_OUT_DOG;
    const bool init_only = outputs & SSPYR_INT_INIT_ONLY;   // K0 state: level = decimated pixel
    float prev[N];
#pragma unroll
    for (int s = 0; s < NL; ++s) {
        const float f = __ldg(fh + (size_t)s * oc.H);
        float g[N];
#pragma unroll
        for (int i = 0; i < N; ++i)
            g[i] = init_only ? p[i] : __fmul_rn(__fmul_rn(p[i], w[s][i]), f);         // K2 then K3
        if (s < NL - 1) {
            if (want_g) store_out<N>(out + (size_t)s * oc.plane, g, nvalid);
        } else {
            if (want_top) store_out<N>(out + (size_t)(2 * NL - 2) * oc.plane, g, nvalid);
        }
        if (s > 0 && want_d) {
            float d[N];
#pragma unroll
            for (int i = 0; i < N; ++i) d[i] = __fsub_rn(prev[i], g[i]);               // K4
            store_out<N>(out + (size_t)(NL - 1 + s - 1) * oc.plane, d, nvalid);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) prev[i] = g[i];
    }
}

template <int NL, int N>
__device__ __forceinline__ void load_windows(float (&w)[NL][N], const RefOct& oc, int ocol) {
#pragma unroll
    for (int s = 0; s < NL; ++s) load_tab<N>(w[s], oc.fw + (size_t)s * oc.pitch + ocol);
}

// One input quad -> float pixels (K0's int->float cast, GuassDePyramid.h:80).
template <int PIX>
__device__ __forceinline__ void load_quad(float (&p)[4], const unsigned char* __restrict__ row, int c, int W) {
    if (c + 4 <= W) {
        if constexpr (PIX == SSPYR_PIXEL_I32) {
            int4 v;
            asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "l"(reinterpret_cast<const int4*>(row) + (c >> 2)));
            p[0] = (float)v.x; p[1] = (float)v.y; p[2] = (float)v.z; p[3] = (float)v.w;
        } else if constexpr (PIX == SSPYR_PIXEL_F32) {
            float4 v;
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "l"(reinterpret_cast<const float4*>(row) + (c >> 2)));
            p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
        } else {
            const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(row) + (c >> 2));
            p[0] = (float)v.x; p[1] = (float)v.y; p[2] = (float)v.z; p[3] = (float)v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = 0.0f;
            if (c + i < W) {
                if constexpr (PIX == SSPYR_PIXEL_I32) v = (float)__ldg(reinterpret_cast<const int*>(row) + c + i);
                else if constexpr (PIX == SSPYR_PIXEL_F32) v = __ldg(reinterpret_cast<const float*>(row) + c + i);
                else v = (float)__ldg(row + c + i);
            }
            p[i] = v;
        }
    }
}

template <int PIX> __host__ __device__ constexpr int elem_bytes() { return PIX == SSPYR_PIXEL_U8 ? 1 : 4; }

// grid.x : chunks of (row group, quad) work items, quads fastest;  grid.y : frame within the batch.
template <int NL, int PIX, int RPT>
__global__ void __launch_bounds__(256)
ref_fused_kernel(const __grid_constant__ RefParams P) {
    const int W4 = (P.W + 3) >> 2;
    const int groups = (P.H + RPT - 1) / RPT;
    const long long items = (long long)W4 * groups;
    const unsigned char* __restrict__ img =
        static_cast<const unsigned char*>(P.img) + (size_t)blockIdx.y * P.img_frame_stride;
    const size_t fofs = (size_t)blockIdx.y * P.out_frame_stride;

    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < items;
         t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t % W4);
        const int rg = (int)(t / W4);
        const int c = j << 2;

        RefOct o0 = P.oct[0];
        o0.base += fofs;
        float w0[NL][4];
        load_windows<NL, 4>(w0, o0, c);

        const int r_end = min(P.H, (rg + 1) * RPT);
#pragma unroll
        for (int rr = 0; rr < RPT; ++rr) {
            const int r = rg * RPT + rr;
            if (r >= r_end) break;
            float p[4];
            load_quad<PIX>(p, img + (size_t)r * P.img_pitch * elem_bytes<PIX>(), c, P.W);
            emit_levels<NL, 4>(o0, P.outputs, r, c, p, w0);

            // higher octaves fed by this quad: r % 2^o == 0 and the column phase matches
            if ((r & 1) == 0 && P.octaves > 1) {
                {   // octave 1: input columns c, c+2 -> output columns c/2, c/2+1
                    RefOct o1 = P.oct[1];
                    const int orow = r >> 1, ocol = c >> 1;
                    if (orow < o1.H && ocol < o1.W) {
                        o1.base += fofs;
                        const float p2[2] = {p[0], p[2]};
                        float w1[NL][2];
                        load_windows<NL, 2>(w1, o1, ocol);
                        emit_levels<NL, 2>(o1, P.outputs, orow, ocol, p2, w1);
                    }
                }
                for (int o = 2; o < P.octaves; ++o) {
                    if ((r & ((1 << o) - 1)) != 0) break;
                    if ((j & ((1 << (o - 2)) - 1)) != 0) break;   // column 4j must be a multiple of 2^o
                    RefOct oc = P.oct[o];
                    const int orow = r >> o, ocol = c >> o;
                    if (orow < oc.H && ocol < oc.W) {
                        oc.base += fofs;
                        const float p1[1] = {p[0]};
                        float w1[NL][1];
                        load_windows<NL, 1>(w1, oc, ocol);
                        emit_levels<NL, 1>(oc, P.outputs, orow, ocol, p1, w1);
                    }
                }
            }
        }
    }
}

template <int NL, int PIX>
cudaError_t launch_rpt(const RefParams& P, int rpt, dim3 grid, int block, cudaStream_t st) {
    switch (rpt) {
        case 1: ref_fused_kernel<NL, PIX, 1><<<grid, block, 0, st>>>(P); break;
        case 2: ref_fused_kernel<NL, PIX, 2><<<grid, block, 0, st>>>(P); break;
        case 4: ref_fused_kernel<NL, PIX, 4><<<grid, block, 0, st>>>(P); break;
        default: ref_fused_kernel<NL, PIX, 8><<<grid, block, 0, st>>>(P); break;
    }
    return cudaGetLastError();
}

template <int NL>
cudaError_t launch_pix(const RefParams& P, int pix, int rpt, dim3 grid, int block, cudaStream_t st) {
    switch (pix) {
        case SSPYR_PIXEL_I32: return launch_rpt<NL, SSPYR_PIXEL_I32>(P, rpt, grid, block, st);
        case SSPYR_PIXEL_F32: return launch_rpt<NL, SSPYR_PIXEL_F32>(P, rpt, grid, block, st);
        default: return launch_rpt<NL, SSPYR_PIXEL_U8>(P, rpt, grid, block, st);
    }
}

}  // namespace

}  // namespace sspyr
