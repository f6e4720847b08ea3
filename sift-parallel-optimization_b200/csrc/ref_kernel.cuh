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
//   * Every load a thread needs (pixel quads of its RPT rows, column windows, row windows) is issued
//     BEFORE its first store, so a thread pays one DRAM latency, not one per level (stores may alias the
//     tables as far as the compiler knows, so it cannot hoist them itself).  Threads are persistent down the
//     frame: column windows are loaded once and the next row group is prefetched before the current one is
//     stored, so the store stream never stalls behind a DRAM read.
//   * 128-bit coalesced loads (ld.global.nc) and streaming stores (st.global.cs): a warp writes 512
//     contiguous bytes per plane per row; rows are 128-byte aligned (pitch % 32 == 0).  Row windows are
//     stored transposed, [row][8], so one row's S+3 values are two warp-uniform 128-bit loads.
//   * Launched with programmatic stream serialization and an immediate launch_dependents trigger: the
//     next frame's grid fills SMs as this one drains (frames are independent), hiding ramp and tail.
//   * No shared memory, no tensor cores: there is no reuse and 0.3 flop/byte.
#pragma once
#include "sspyr_internal.h"

namespace sspyr {

namespace {

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

// Row window of one output row: fh_t[row][0..FT-1] (levels padded to FT = 8, or 16 when S+3 > 8) -> warp-uniform
// 128-bit loads.
template <int NL> __host__ __device__ constexpr int row_table_stride() { return NL <= 8 ? 8 : 16; }

template <int NL>
__device__ __forceinline__ void load_row_window(float (&f)[NL], const float* __restrict__ fh_t, int orow) {
    constexpr int FT = row_table_stride<NL>();
    const float4* q = reinterpret_cast<const float4*>(fh_t + (size_t)orow * FT);
#pragma unroll
    for (int k = 0; k < (NL + 3) / 4; ++k) {
        const float4 a = __ldg(q + k);
        if (4 * k + 0 < NL) f[4 * k + 0] = a.x;
        if (4 * k + 1 < NL) f[4 * k + 1] = a.y;
        if (4 * k + 2 < NL) f[4 * k + 2] = a.z;
        if (4 * k + 3 < NL) f[4 * k + 3] = a.w;
    }
}

// Streaming (evict-first) store of N contiguous floats; FULL = all N exist (vector store).
template <int N, bool FULL>
__device__ __forceinline__ void store_out(float* __restrict__ p, const float (&v)[N], int nvalid) {
    if constexpr (FULL) {
        if constexpr (N == 4) __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
        else if constexpr (N == 2) __stcs(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
        else __stcs(p, v[0]);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (i < nvalid) __stcs(p + i, v[i]);
    }
}

// All S+3 levels and S+2 DoGs of N horizontally adjacent pixels of one octave; every operand is already
// in registers.   out = &plane0[orow][ocol];  planes are `plane` floats apart:
//   [G_0..G_{S+1} | DoG_0..DoG_{S+1} | G_{S+2}]
template <int NL, int N, bool FULL>
__device__ __forceinline__ void emit_levels(float* __restrict__ out, unsigned plane, int outputs, int nvalid,
                                            const float (&p)[N], const float (&w)[NL][N], const float (&fh)[NL]) {
    const bool want_g = outputs & SSPYR_OUT_GAUSS;
    const bool want_top = outputs & (SSPYR_OUT_GAUSS | SSPYR_OUT_GAUSS_TOP);
    const bool want_d = outputs & SSPYR_OUT_DOG;
    const bool init_only = outputs & SSPYR_INT_INIT_ONLY;   // K0 state: level = decimated pixel
    float prev[N];
#pragma unroll
    for (int s = 0; s < NL; ++s) {
        float g[N];
#pragma unroll
        for (int i = 0; i < N; ++i)
            g[i] = init_only ? p[i] : __fmul_rn(__fmul_rn(p[i], w[s][i]), fh[s]);      // K2 then K3
        if (s < NL - 1) {
            if (want_g) store_out<N, FULL>(out + (size_t)s * plane, g, nvalid);
        } else {
            if (want_top) store_out<N, FULL>(out + (size_t)(2 * NL - 2) * plane, g, nvalid);
        }
        if (s > 0 && want_d) {
            float d[N];
#pragma unroll
            for (int i = 0; i < N; ++i) d[i] = __fsub_rn(prev[i], g[i]);               // K4
            store_out<N, FULL>(out + (size_t)(NL - 1 + s - 1) * plane, d, nvalid);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) prev[i] = g[i];
    }
}

template <int NL, int N>
__device__ __forceinline__ void load_windows(float (&w)[NL][N], const float* __restrict__ fw, int pitch, int ocol) {
#pragma unroll
    for (int s = 0; s < NL; ++s) load_tab<N>(w[s], fw + (size_t)s * pitch + ocol);
}

template <int NL, int N>
__device__ __forceinline__ void emit_dispatch(float* __restrict__ out, unsigned plane, int outputs, int nvalid,
                                              const float (&p)[N], const float (&w)[NL][N], const float (&fh)[NL]) {
    if (nvalid >= N) emit_levels<NL, N, true>(out, plane, outputs, nvalid, p, w, fh);
    else emit_levels<NL, N, false>(out, plane, outputs, nvalid, p, w, fh);
}

// One input quad -> float pixels (K0's int->float cast, GuassDePyramid.h:80).
template <int PIX>
__device__ __forceinline__ void load_quad(float (&p)[4], const unsigned char* __restrict__ row, int c, int W) {
    if (c + 4 <= W) {
        if constexpr (PIX == SSPYR_PIXEL_I32) {
            int4 v;
            asm("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                : "l"(reinterpret_cast<const int4*>(row) + (c >> 2)));
            p[0] = (float)v.x; p[1] = (float)v.y; p[2] = (float)v.z; p[3] = (float)v.w;
        } else if constexpr (PIX == SSPYR_PIXEL_F32) {
            float4 v;
            asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
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

// grid.x: blocks of BX quads along a row;  grid.y: blocks of BY row groups (RPT rows each);  grid.z: frame.
// Everything one thread needs for RPT consecutive rows of its column quad.
template <int NL, int RPT>
struct RefRows {
    float p[RPT][4];                 // pixel quads (K0's int->float cast already applied)
    float f0[RPT][NL];               // octave-0 row windows
    float f1[(RPT + 1) / 2][NL];     // octave-1 row windows of the even rows
};

template <int NL, int PIX, int RPT>
__device__ __forceinline__ void ref_load_rows(RefRows<NL, RPT>& q, const RefParams& P, const unsigned char* __restrict__ img,
                                              int r0, int c, bool col1) {
#pragma unroll
    for (int rr = 0; rr < RPT; ++rr) {
        const int r = min(r0 + rr, P.H - 1);                        // clamp: rows past the end are not stored
        load_quad<PIX>(q.p[rr], img + (size_t)r * P.img_pitch * elem_bytes<PIX>(), c, P.W);
    }
#pragma unroll
    for (int rr = 0; rr < RPT; ++rr) load_row_window<NL>(q.f0[rr], P.oct[0].fh, min(r0 + rr, P.H - 1));
    if (col1 && (RPT > 1 || (r0 & 1) == 0)) {
#pragma unroll
        for (int k = 0; k < (RPT + 1) / 2; ++k) load_row_window<NL>(q.f1[k], P.oct[1].fh, min((r0 >> 1) + k, P.oct[1].H - 1));
    }
}

// grid.x: blocks of bx quads along a row;  grid.y: blocks of by row groups (RPT rows each);  grid.z: frame.
// A thread keeps its column quad and walks down the frame in steps of gridDim.y*blockDim.y*RPT rows: the column
// windows are loaded once, and the loads of the NEXT row group are issued before the stores of the current one
// (software prefetch), so after the first group no thread ever waits on DRAM with its stores unissued.  With a
// grid that covers every row group the loop runs once.
// WALK = false: one row group per thread (no prefetch registers); WALK = true: the persistent row walk.
// (S+3 > 8 levels is the rarely used wide build: no register cap, so nothing spills)
template <int NL, int PIX, int RPT, bool WALK>
__global__ void __launch_bounds__(128, NL <= 8 ? 4 : 1)
ref_fused_kernel(const __grid_constant__ RefParams P) {
    // Frames are independent: let the next launch in the stream start filling SMs right away (PDL).
    asm volatile("griddepcontrol.launch_dependents;");

    const int j = blockIdx.x * blockDim.x + threadIdx.x;            // quad index along the row
    int r0 = (blockIdx.y * blockDim.y + threadIdx.y) * RPT;         // first of this thread's RPT rows
    const int stride = gridDim.y * blockDim.y * RPT;
    const int c = j << 2;
    if (c < P.W && r0 < P.H) {
        const unsigned char* __restrict__ img =
            static_cast<const unsigned char*>(P.img) + (size_t)blockIdx.z * P.img_frame_stride;
        const size_t fofs = (size_t)blockIdx.z * P.out_frame_stride;
        const int outputs = P.outputs;
        const RefOct& o0 = P.oct[0];
        const RefOct& o1 = P.oct[1];
        const int ocol1 = c >> 1;
        const bool col1 = P.octaves > 1 && ocol1 < o1.W;            // this quad feeds octave 1 (input columns c, c+2)
        constexpr int R1 = (RPT + 1) / 2;

        // ---- loads that do not depend on the row: column windows of octaves 0 and 1 ----------------------
        float w0[NL][4];
        load_windows<NL, 4>(w0, o0.fw, o0.pitch, c);
        float w1[NL][2];
        if (col1) load_windows<NL, 2>(w1, o1.fw, o1.pitch, ocol1);
        RefRows<NL, RPT> cur;
        ref_load_rows<NL, PIX, RPT>(cur, P, img, r0, c, col1);
        const int nvalid0 = o0.W - c, nvalid1 = o1.W - ocol1;

        for (;;) {
            const int rn = r0 + stride;
            const bool more = WALK && rn < P.H;
            RefRows<NL, WALK ? RPT : 1> nxt;
            if constexpr (WALK) {
                if (more) ref_load_rows<NL, PIX, RPT>(nxt, P, img, rn, c, col1);   // prefetch before this group's stores
            }

            // ---- octave 0 ----------------------------------------------------------------------------------
            float* out0 = o0.base + fofs + (size_t)r0 * o0.pitch + c;
#pragma unroll
            for (int rr = 0; rr < RPT; ++rr)
                if (r0 + rr < P.H)
                    emit_dispatch<NL, 4>(out0 + (size_t)rr * o0.pitch, (unsigned)o0.plane, outputs, nvalid0, cur.p[rr], w0, cur.f0[rr]);

            // ---- octave 1: even rows of the group (r0 is a multiple of RPT) ------------------------------------
            if (col1 && (RPT > 1 || (r0 & 1) == 0)) {
                float* out1 = o1.base + fofs + (size_t)(r0 >> 1) * o1.pitch + ocol1;
#pragma unroll
                for (int k = 0; k < R1; ++k) {
                    const int orow = (r0 >> 1) + k;
                    if (orow < o1.H && r0 + 2 * k < P.H) {
                        const float p2[2] = {cur.p[2 * k < RPT ? 2 * k : 0][0], cur.p[2 * k < RPT ? 2 * k : 0][2]};
                        emit_dispatch<NL, 2>(out1 + (size_t)k * o1.pitch, (unsigned)o1.plane, outputs, nvalid1, p2, w1, cur.f1[k]);
                    }
                }
            }

            // ---- octaves >= 2: one pixel per participating quad (1/16 of the octave-0 work and falling) ------
#pragma unroll
            for (int rr = 0; rr < RPT; rr += 4) {
                const int r = r0 + rr;
                if (r >= P.H) break;
                for (int o = 2; o < P.octaves; ++o) {
                    if ((r & ((1 << o) - 1)) != 0) break;
                    if ((j & ((1 << (o - 2)) - 1)) != 0) break;     // column 4j must be a multiple of 2^o
                    const RefOct& oc = P.oct[o];
                    const int orow = r >> o, ocol = c >> o;
                    if (orow < oc.H && ocol < oc.W) {
                        const float p1[1] = {cur.p[rr][0]};
                        float wv[NL][1];
                        float fv[NL];
                        load_windows<NL, 1>(wv, oc.fw, oc.pitch, ocol);
                        load_row_window<NL>(fv, oc.fh, orow);
                        emit_levels<NL, 1, true>(oc.base + fofs + (size_t)orow * oc.pitch + ocol, (unsigned)oc.plane, outputs,
                                                 1, p1, wv, fv);
                    }
                }
            }
            if (!more) break;
            if constexpr (WALK) cur = nxt;
            r0 = rn;
        }
    }
    // Chain completion: this grid may not finish before the grid it overlapped with has finished and
    // flushed, so "kernel N+1 done" always implies "kernel N done" for whatever follows in the stream.
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Pull the rows of a frame toward L2 (one 128-byte line per thread, evict-last): issued for the NEXT frame slot
// while the current one is being built, so that the build kernel's input loads are L2 hits and the DRAM sees the
// reads as one dense burst instead of a trickle between its writes.
__global__ void __launch_bounds__(256)
ref_prefetch_kernel(const unsigned char* __restrict__ img, unsigned long long pitch_bytes, int row_bytes, int rows) {
    asm volatile("griddepcontrol.launch_dependents;");
    const int lines = (row_bytes + 127) >> 7;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (long long)lines * rows) {
        const int r = (int)(t / lines), l = (int)(t - (long long)r * lines);
        asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(img + (size_t)r * pitch_bytes + ((size_t)l << 7)));
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

inline cudaError_t launch_prefetch(const void* img, size_t pitch_bytes, int row_bytes, int rows, cudaStream_t st) {
    const long long n = (long long)((row_bytes + 127) >> 7) * rows;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((n + 255) / 256));
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, ref_prefetch_kernel, static_cast<const unsigned char*>(img),
                              (unsigned long long)pitch_bytes, row_bytes, rows);
}

template <int NL, int PIX, int RPT, bool WALK>
cudaError_t launch_one(const RefParams& P, dim3 grid, dim3 block, cudaStream_t st, bool pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, ref_fused_kernel<NL, PIX, RPT, WALK>, P);
}

template <int NL, int PIX>
cudaError_t launch_rpt(const RefParams& P, int rpt, bool walk, dim3 grid, dim3 block, cudaStream_t st, bool pdl) {
    if (walk) return rpt == 1 ? launch_one<NL, PIX, 1, true>(P, grid, block, st, pdl) : launch_one<NL, PIX, 2, true>(P, grid, block, st, pdl);
    return rpt == 1 ? launch_one<NL, PIX, 1, false>(P, grid, block, st, pdl) : launch_one<NL, PIX, 2, false>(P, grid, block, st, pdl);
}

template <int NL>
cudaError_t launch_pix(const RefParams& P, int pix, int rpt, bool walk, dim3 grid, dim3 block, cudaStream_t st, bool pdl) {
    switch (pix) {
        case SSPYR_PIXEL_I32: return launch_rpt<NL, SSPYR_PIXEL_I32>(P, rpt, walk, grid, block, st, pdl);
        case SSPYR_PIXEL_F32: return launch_rpt<NL, SSPYR_PIXEL_F32>(P, rpt, walk, grid, block, st, pdl);
        default: return launch_rpt<NL, SSPYR_PIXEL_U8>(P, rpt, walk, grid, block, st, pdl);
    }
}

}  // namespace

}  // namespace sspyr
