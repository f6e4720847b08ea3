// csrc/conv_march.cuh -- CONV mode, large levels: the tile kernel made to MARCH down a column strip.
//
// Same arithmetic as conv_kernel.cuh (same taps, same row-then-column order, same fp32 FMA chains: results are
// bit-identical), different schedule.  The plain tile kernel row-filters TH+2R rows to produce TH output rows
// (a 1.4-1.6x FMA overhead at TH = 32), its row pass is unbalanced (44 rows x 8 blocks over 256 threads) and
// every CTA first waits for its own loads.  ncu shows it issue-bound (IPC 0.66, top stall "not selected"), so
// the only way forward is fewer instructions.  Here a CTA owns a 128-column strip and walks down a vertical
// segment in steps of 32 rows:
//
//     step k:   wait for the 32 new input rows (cp.async, issued during step k-1's column pass)
//               row pass of exactly those 32 rows (32 x 8 tasks = one per thread) into sT rows [2R, 2R+32)
//               issue the cp.async loads of step k+1 (the staging buffer is free again)
//               column pass + epilogue for 32 output rows out of sT rows [0, 32+2R)
//               carry the last 2R rows of sT to the top for the next step
//
// so the row pass runs once per input row (only the 2R warm-up rows of a segment are extra), every thread has
// the same amount of work in every phase, the loads of the next step overlap the column pass of this one
// without a second staging buffer, and shared memory per CTA drops to ~44 KB (4 CTAs per SM).  The DoG centre
// values are read out of the staged tile before it is handed to the next step's loads (the first R rows of a step,
// staged one step earlier, are re-read from L2).  Segments are 8 steps long where the level is large enough.
//
// Schedule across launches: consecutive levels of an octave are CHAINED -- every (strip, segment)
// CTA counts its builds in a per-slot counter and the next level's CTAs wait for the 3x3 segments they read instead
// of for the whole grid; row bands over peer memory wait for / signal the neighbour GPU's per-slot progress
// counters from their two edge segments, which are dispatched last.
#pragma once
#include <cuda.h>            // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint)

#include "conv_kernel.cuh"

namespace sspyr {

// ---- TMA (cp.async.bulk.tensor) + mbarrier helpers --------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// Bounded wait (a lost TMA must not hang the GPU): returns false on time-out.
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
#pragma unroll 1
    for (int spin = 0; spin < (1 << 22); ++spin) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

// One 3-D box (columns, rows, frame) of the source plane -> shared memory, completion on `bar`.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2),
                   "r"(smem_u32(bar)) : "memory");
}

// Same, with an L2 evict-first policy: the source plane of a level is dead once the level has read it (only the
// halo re-reads of neighbouring CTAs follow shortly), so it should not displace the G_s rows being written, which
// the chained next level reads back from L2.
__device__ __forceinline__ void tma_load_3d_evict_first(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2,
                                                        unsigned long long* bar) {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2),
                   "r"(smem_u32(bar)), "l"(pol) : "memory");
}

template <int R> __host__ __device__ constexpr size_t strip_smem_bytes() {
    return sizeof(float) * ((size_t)STRIP_TH * conv_pitch_in<R>() + (size_t)(STRIP_TH + 2 * R) * conv_pitch_t()) + 16;   // + mbarrier
}
// (Compiling the kernel for 5 resident CTAs per SM -- 48 registers, a few spilled words, possible for radii <= 8 -- was
//  measured 3-6 % SLOWER than 4 CTAs at 64 registers on every workload, profiles/r2_conv_experiments.md; it stays at 4.)

namespace {

// Stage input rows gy0 .. gy0+NROWS-1 (clamped to the frame, or taken from the neighbour band's halo rows) of the
// strip starting at column x0 into sIn rows 0..NROWS-1.  16-byte chunks that lie inside the row go through
// cp.async (float planes) or a 128-bit load + convert (raw frames); chunks that straddle the edge are clamped.
// (A warp-per-row variant with the row pointer hoisted was measured 15 % slower: fewer copies in flight.)
template <int R, int SRC, int NROWS>
__device__ __forceinline__ void strip_stage_rows(const ConvParams& P, float* __restrict__ sIn, int gy0, int x0,
                                                 const unsigned char* __restrict__ src, int tid) {
    constexpr int RA = conv_ra<R>();
    constexpr int PIN = conv_pitch_in<R>();
    constexpr int CH = (CONV_TW + 2 * RA) / 4;               // 16-byte chunks per staged row (34..40)
    constexpr int elem = SRC == SSPYR_PIXEL_U8 ? 1 : 4;
#pragma unroll 1
    for (int c = tid; c < NROWS * CH; c += CONV_THREADS) {       // chunks dealt linearly: every thread has 4-5 in flight
        const int rr = c / CH, q = c - rr * CH;
        const int gy = gy0 + rr;
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
        float* s = sIn + rr * PIN + 4 * q;
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
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                s[i] = load_src<SRC>(row, (size_t)min(max(gx + i, 0), P.W - 1));
        }
    }
}

// Row pass of NROWS staged rows: sT[T0 + row][c] = sum_k taps[k] * sIn[row][RA + c + k - R]
// (PIN = pitch of the staged rows; taps = 2R+1 values in the constant bank, centre at R)
template <int R, int NROWS, int T0, int PIN = conv_pitch_in<R>()>
__device__ __forceinline__ void strip_row_pass(const float* __restrict__ taps, const float* __restrict__ sIn, float* __restrict__ sT, int tid) {
    constexpr int RA = conv_ra<R>();
    constexpr int PT = conv_pitch_t();
    constexpr int NB = CONV_TW / CONV_PX;
    constexpr int NIN = CONV_PX + 2 * RA;
#pragma unroll 1
    for (int task = tid; task < NROWS * NB; task += CONV_THREADS) {
        const int row = task % NROWS, cb = task / NROWS;          // consecutive lanes -> consecutive rows (compile-time NROWS)
        const float4* in4 = reinterpret_cast<const float4*>(sIn + row * PIN + cb * CONV_PX);
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
        for (int k = 0; k <= 2 * R; ++k) {                     // taps outer: 16 independent FMA chains in flight
            const float w = taps[k];                           // (packed FFMA2 here needs shifted operand pairs: the
#pragma unroll                                                 //  extra moves and registers made it slower, measured)
            for (int i = 0; i < CONV_PX; ++i) acc[i] = fmaf(w, in[i + k + (RA - R)], acc[i]);
        }
        float4* out4 = reinterpret_cast<float4*>(sT + (T0 + row) * PT + cb * CONV_PX);
#pragma unroll
        for (int q = 0; q < CONV_PX / 4; ++q) out4[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    }
}

// grid: x = 128-column strips, y = vertical segments of seg_rows (multiple of 32) output rows, z = frame
// TMA = true (float-plane sources): the 32-row x PIN-column box of every interior step is fetched by ONE
// cp.async.bulk.tensor issued by one thread and tracked by an mbarrier, instead of ~1200 cp.async from all threads.
// Steps that touch the frame edge (clamp-to-edge is not a TMA fill mode) or a neighbour band's halo rows keep
// the cp.async path.
template <int R, int SRC, bool TMA>
__global__ void __launch_bounds__(CONV_THREADS, STRIP_CTAS_PER_SM)
conv_strip_kernel(const __grid_constant__ ConvParams P, int seg_rows, const __grid_constant__ CUtensorMap tmap) {
    static_assert(2 * R <= STRIP_TH, "the carried rows must fit above the new ones");
    constexpr int TH = STRIP_TH;
    constexpr int PIN = conv_pitch_in<R>();
    constexpr int PT = conv_pitch_t();
    // Column-pass mapping: a thread owns PX = 4 adjacent columns x PY = 4 rows.  (A 2 x 8 mapping -- 36 % fewer
    // column-pass shared-memory wavefronts -- and a running-pointer epilogue -- ~30 fewer instructions per step --
    // were built and measured in round 2: both neutral to slower, profiles/r2_conv_eval_builds_ab.txt, and deleted.)
    constexpr int PX = 4;
    constexpr int PY = 16 / PX;
    constexpr int TPB = CONV_TW / PX;                    // threads per row block of the column pass (32 or 64)
    static_assert(TPB * (TH / PY) == CONV_THREADS, "column-pass mapping must cover the step");
    constexpr int elem = SRC == SSPYR_PIXEL_U8 ? 1 : 4;
    extern __shared__ __align__(128) float strip_smem[];   // (own name: the tile kernel's array is declared 16-byte aligned)
    float* sIn = strip_smem;                            // [TH][PIN]     staged input rows (centre at column RA)
    float* sT = strip_smem + (size_t)TH * PIN;          // [TH+2R][PT]   row-pass results: 2R carried rows + TH new ones
    // mbarrier of the TMA staging: kept in the dynamic allocation so that sIn stays at its 128-byte aligned base
    unsigned long long& bar = *reinterpret_cast<unsigned long long*>(sT + (size_t)(TH + 2 * R) * PT);
    constexpr int RA_ = conv_ra<R>();

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * CONV_TW;
    // Row bands over peer memory: CTAs are dispatched in blockIdx order, so the two segments at the band edges --
    // the only ones that wait for a neighbour GPU -- are mapped to the LAST y indices: they start when the interior
    // is already under way (the neighbour has usually published by then) and never keep interior CTAs off the SMs.
    const bool edge_last = P.wait_up || P.wait_dn || P.peer_seg_up || P.peer_seg_dn;
    const int seg = edge_last ? (int)((blockIdx.y + 1) % gridDim.y) : (int)blockIdx.y;
    const int y_begin = seg * seg_rows;
    const size_t fz = blockIdx.z;
    // Programmatic dependent launch along the level chain.
    //   * no segment counters (tile-kernel neighbours, stepwise builds): the next level may be scheduled while this one
    //     drains; it reads nothing this kernel writes before its own griddepcontrol.wait returns, i.e. before this
    //     grid has completed and flushed.
    //   * level chaining (seg_pub set): every CTA counts the builds of its (strip, segment) in seg_pub when its rows
    //     are written.  A level with seg_dep does NOT wait for the previous grid: each CTA waits only for the up to
    //     3x3 segments of the previous level that its tile + halo reads, so consecutive levels overlap (no idle
    //     tail between them, and the rows just written are still in L2).  On a row band whose neighbours are attached,
    //     the first / last segment row additionally waits for the neighbour band's last / first segment rows of the
    //     previous level, through the peer mapping (system scope): the seam is just one more segment boundary.
    //     The first level of a chain keeps the grid-wide wait and lets its dependents go only AFTER it: a chained CTA can
    //     then never run before the previous build of these planes has completed (own counter final, nobody still
    //     reading what it overwrites).  Dependents are scheduled only once every CTA of this grid has started, so a
    //     waiting CTA's producers are always resident or done (or run on another GPU): no deadlock.
    unsigned seg_next = 0;
    const size_t seg_idx = fz * P.seg_frame_stride + (size_t)seg * gridDim.x + blockIdx.x;
    if (P.seg_dep) {
        asm volatile("griddepcontrol.launch_dependents;");
        if (tid < 32) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seg_next) : "l"(P.seg_pub + seg_idx) : "memory");
            seg_next += 1;
            // lanes 0..8: the 3x3 segments of this band; 9..14: neighbour-above (up to 2 segment rows x 3 strips);
            // 15..17: neighbour-below (its first segment row x 3 strips)
            const unsigned* f = nullptr;
            bool sys = false;
            if (tid < 9) {
                const int sx = (int)blockIdx.x + tid % 3 - 1, sy = seg + tid / 3 - 1;
                if (sx >= 0 && sx < (int)gridDim.x && sy >= 0 && sy < (int)gridDim.y)
                    f = P.seg_dep + fz * P.seg_frame_stride + (size_t)sy * gridDim.x + sx;
            } else if (tid < 15) {
                const int q = tid - 9, sx = (int)blockIdx.x + q % 3 - 1, sy = P.peer_up_first + q / 3;
                if (P.peer_seg_up && seg == 0 && sx >= 0 && sx < (int)gridDim.x && sy < P.peer_up_nsegs) {
                    f = P.peer_seg_up + (size_t)sy * gridDim.x + sx;
                    sys = true;
                }
            } else if (tid < 18) {
                const int sx = (int)blockIdx.x + (tid - 15) - 1;
                if (P.peer_seg_dn && (seg + 1) * seg_rows + R > P.H && sx >= 0 && sx < (int)gridDim.x) {   // halo reaches below the band
                    f = P.peer_seg_dn + sx;
                    sys = true;
                }
            }
            if (f) {
                const long long t0 = clock64();
                for (;;) {
                    unsigned v;
                    if (sys) asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                    else asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                    if ((int)(v - seg_next) >= 0) break;
                    if (*reinterpret_cast<volatile unsigned*>(P.timeout_mark) != 0) break;     // someone gave up: do not pile up waits
                    if (clock64() - t0 > 4000000000LL) { *P.timeout_mark = seg_next | 0x80000000u; break; }
                    __nanosleep(64);
                }
            }
            __syncwarp();
            if (tid == 0) asm volatile("fence.proxy.async.global;" ::: "memory");   // the TMA loads below read those rows
        }
        __syncthreads();
    } else if (P.seg_pub) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;");
        if (tid == 0) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seg_next) : "l"(P.seg_pub + seg_idx) : "memory");
            seg_next += 1;
        }
    } else {
        asm volatile("griddepcontrol.launch_dependents;");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    if (y_begin >= P.H) return;                          // (never true: the grid covers exactly the segments)
    const int y_end = min(P.H, y_begin + seg_rows);
    const int nsteps = (y_end - y_begin + TH - 1) / TH;

    // Fused halo synchronisation (row bands over peer memory): only the CTAs whose rows reach into a neighbour
    // band wait for that neighbour to have published the level they read; interior CTAs start at once, so the
    // halo latency hides behind the band's interior.
    const unsigned band_epoch = P.epoch ? (*P.epoch - 1u) * CONV_FLAG_STRIDE : 0u;   // (written by the build's first kernel)
    if (tid == 0) {
        const unsigned wait_need = band_epoch + P.wait_need;
        const bool need_up = P.wait_up && y_begin - R < 0;
        const bool need_dn = P.wait_dn && y_begin + nsteps * TH + R > P.H;
        for (int side = 0; side < 2; ++side) {
            const unsigned* f = side == 0 ? (need_up ? P.wait_up : nullptr) : (need_dn ? P.wait_dn : nullptr);
            if (!f) continue;
            const long long t0 = clock64();
            for (;;) {
                unsigned v;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                if (v >= wait_need) break;
                if (clock64() - t0 > 4000000000LL) { *P.timeout_mark = wait_need; break; }
                __nanosleep(100);
            }
        }
    }
    if (P.wait_up || P.wait_dn) __syncthreads();

    const unsigned char* src = static_cast<const unsigned char*>(P.src) + fz * P.src_frame_stride * elem;
    // a step may use TMA when its whole box lies inside the plane: needed columns inside [0, W), the 4-column
    // overshoot of the padded box inside the row pitch, rows inside [0, H)
    const bool tma_cols = TMA && x0 - RA_ >= 0 && x0 + CONV_TW + RA_ <= P.W && x0 - RA_ + PIN <= P.src_pitch;
    unsigned phase = 0;
    bool pending_tma = false;
    if (TMA) {
        if (tid == 0) mbar_init(&bar, 1);
        __syncthreads();
    }
    auto stage_step = [&](int gy0) {                     // rows gy0 .. gy0+TH-1 -> sIn
        if (TMA && tma_cols && gy0 >= 0 && gy0 + TH <= P.H) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of sIn are done
                mbar_expect_tx(&bar, (unsigned)(TH * PIN * sizeof(float)));
                if (P.src_evict_first) tma_load_3d_evict_first(sIn, &tmap, x0 - RA_, gy0, (int)fz, &bar);
                else tma_load_3d(sIn, &tmap, x0 - RA_, gy0, (int)fz, &bar);
            }
            pending_tma = true;
        } else {
            strip_stage_rows<R, SRC, TH>(P, sIn, gy0, x0, src, tid);
            pending_tma = false;
        }
        __pipeline_commit();
    };

    // prologue: the 2R warm-up rows above the segment go to a scratch area inside sT (the rows the first step's row
    // pass will overwrite later) while the first step's 32 rows go to sIn -- both in flight at once, one wait
    float* scratch = sT + (size_t)(2 * R) * PT;
    static_assert((size_t)2 * R * conv_pitch_in<R>() <= (size_t)STRIP_TH * conv_pitch_t(), "warm-up scratch must fit in sT");
    strip_stage_rows<R, SRC, 2 * R>(P, scratch, y_begin - R, x0, src, tid);
    stage_step(y_begin + R);
    __pipeline_wait_prior(0);
    bool lost = false;                                   // a TMA load never completed (bounded wait): flag it, skip the work,
    if (pending_tma) {                                   // but still publish the counters so that nobody waits on this CTA
        if (!mbar_wait(&bar, phase)) lost = true;
        phase ^= 1;
    }
    lost = __syncthreads_or(lost);                       // (CTA-uniform: threads time out individually)
    if (!lost) strip_row_pass<R, 2 * R, 0>(P.taps, scratch, sT, tid);      // scratch (sT rows >= 2R) -> sT rows [0, 2R)

    const int cq = tid % TPB, rb = tid / TPB;           // column group / row block of the column pass
    const int x = x0 + cq * PX;
    const int nvalid = P.W - x;
    float* g = P.dst_g + fz * P.dst_frame_stride;
    float* d = P.dst_d ? P.dst_d + fz * P.dst_frame_stride : nullptr;
    float* dec = P.dst_dec ? P.dst_dec + fz * P.dst_frame_stride : nullptr;

#pragma unroll 1
    for (int k = 0; k < nsteps && !lost; ++k) {
        bool miss = false;
        if (k > 0) {                                     // (step 0 was awaited in the prologue)
            if (pending_tma) {
                miss = !mbar_wait(&bar, phase);          // (bounded; never observed)
                phase ^= 1;
            } else {
                __pipeline_wait_prior(0);
            }
        }
        if (__syncthreads_or(miss)) { lost = true; break; }   // new rows landed; carried rows are in place; scratch is free
        strip_row_pass<R, TH, 2 * R>(P.taps, sIn, sT, tid);

        // ---- centre values for DoG_{s-1} = G_{s-1} - G_s: input rows yr .. yr+PY-1 of this thread's quad ----
        // Output row y0 + m is staged row m - R of this step, so all but the first R rows of the step are still in
        // sIn: they are read from shared memory here, BEFORE the buffer is handed to the next step's loads (a global
        // re-read was an L2 hit whose latency the column pass could not cover: 11 % of all stall samples).  Only the
        // warps that own the first R output rows re-read those rows from global (they were staged one step ago).
        const int y0 = y_begin + k * TH;
        const int yr = y0 + rb * PY;                     // first output row of this thread
        float cen[PY][PX];
        if (d && nvalid >= PX) {
#pragma unroll
            for (int j = 0; j < PY; ++j) {
                const int m = rb * PY + j - R;           // staged row of this step (warp-uniform)
                if (m >= 0) {
                    const float* c = sIn + (size_t)m * PIN + RA_ + cq * PX;
                    if constexpr (PX == 4) {
                        const float4 t = *reinterpret_cast<const float4*>(c);
                        cen[j][0] = t.x; cen[j][1] = t.y; cen[j][2] = t.z; cen[j][3] = t.w;
                    } else {
                        const float2 t = *reinterpret_cast<const float2*>(c);
                        cen[j][0] = t.x; cen[j][1] = t.y;
                    }
                } else {
                    const unsigned char* crow = src + (size_t)min(yr + j, P.H - 1) * P.src_pitch * elem;
                    if constexpr (PX == 4) {
                        if constexpr (SRC == SSPYR_PIXEL_I32) {
                            const int4 t = __ldg(reinterpret_cast<const int4*>(reinterpret_cast<const int*>(crow) + x));
                            cen[j][0] = (float)t.x; cen[j][1] = (float)t.y; cen[j][2] = (float)t.z; cen[j][3] = (float)t.w;
                        } else if constexpr (SRC == SSPYR_PIXEL_U8) {
                            const uchar4 t = __ldg(reinterpret_cast<const uchar4*>(crow + x));
                            cen[j][0] = (float)t.x; cen[j][1] = (float)t.y; cen[j][2] = (float)t.z; cen[j][3] = (float)t.w;
                        } else if constexpr (SRC == CONV_SRC_PLANE) {   // the producing grid may still be running: coherent load
                            const float4 t = __ldcg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(crow) + x));
                            cen[j][0] = t.x; cen[j][1] = t.y; cen[j][2] = t.z; cen[j][3] = t.w;
                        } else {
                            const float4 t = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(crow) + x));
                            cen[j][0] = t.x; cen[j][1] = t.y; cen[j][2] = t.z; cen[j][3] = t.w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < PX; ++i)
                            cen[j][i] = load_src<SRC>(crow, (size_t)(x + i));
                    }
                }
            }
        }
        __syncthreads();
        if (k + 1 < nsteps) stage_step(y_begin + R + (k + 1) * TH);   // in flight during the column pass below

        // ---- column pass: output rows y0 + rb*PY + j from sT rows rb*PY + j .. + 2R -------------------------
        const float* tcol = sT + (size_t)(rb * PY) * PT + cq * PX;
        f32x2 a01[PY], a23[PX == 4 ? PY : 1];            // packed accumulators: columns (0,1) [and (2,3)] of each row
#pragma unroll
        for (int j = 0; j < PY; ++j) a01[j] = pk2(0.0f, 0.0f);
        if constexpr (PX == 4) {
#pragma unroll
            for (int j = 0; j < PY; ++j) a23[j] = pk2(0.0f, 0.0f);
        }
#pragma unroll
        for (int i = 0; i < PY + 2 * R; ++i) {
            f32x2 v01, v23 = 0;
            if constexpr (PX == 4) {
                const float4 v = *reinterpret_cast<const float4*>(tcol + (size_t)i * PT);
                v01 = pk2(v.x, v.y);
                v23 = pk2(v.z, v.w);
            } else {
                const float2 v = *reinterpret_cast<const float2*>(tcol + (size_t)i * PT);
                v01 = pk2(v.x, v.y);
            }
#pragma unroll
            for (int j = 0; j < PY; ++j) {
                if (i - j >= 0 && i - j <= 2 * R) {            // compile-time after unrolling
                    const f32x2 w = pk2(P.taps[i - j], P.taps[i - j]);
                    a01[j] = fma2(w, v01, a01[j]);
                    if constexpr (PX == 4) a23[j] = fma2(w, v23, a23[j]);
                }
            }
        }
        float acc[PY][PX];
#pragma unroll
        for (int j = 0; j < PY; ++j) {
            unpk2(a01[j], acc[j][0], acc[j][1]);
            if constexpr (PX == 4) unpk2(a23[j], acc[j][2], acc[j][3]);
        }
        if (nvalid >= PX) {                              // full group: vector stores, one running offset
            unsigned o = (unsigned)yr * (unsigned)P.dst_pitch + (unsigned)x;   // a plane has < 2^32 floats
#pragma unroll
            for (int j = 0; j < PY; ++j, o += (unsigned)P.dst_pitch) {
                const int y = yr + j;
                if (y >= y_end) break;
                if constexpr (PX == 4) {
                    *reinterpret_cast<float4*>(g + o) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
                    if (d)                               // DoG_{s-1} = G_{s-1} - G_s  (GuassDePyramid.h:143)
                        __stcs(reinterpret_cast<float4*>(d + o), make_float4(cen[j][0] - acc[j][0], cen[j][1] - acc[j][1],
                                                                             cen[j][2] - acc[j][2], cen[j][3] - acc[j][3]));
                } else {
                    *reinterpret_cast<float2*>(g + o) = make_float2(acc[j][0], acc[j][1]);
                    if (d) __stcs(reinterpret_cast<float2*>(d + o), make_float2(cen[j][0] - acc[j][0], cen[j][1] - acc[j][1]));
                }
                if (dec && (y & 1) == 0) {               // even-phase decimation (GuassDePyramid.h:80)
                    const int dy = y >> 1, dx = x >> 1;
                    if (dy < P.dec_H && dx < P.dec_W) {
                        float* q = dec + (size_t)dy * P.dec_pitch + dx;
                        if constexpr (PX == 4) {
                            if (dx + 1 < P.dec_W) *reinterpret_cast<float2*>(q) = make_float2(acc[j][0], acc[j][2]);
                            else q[0] = acc[j][0];
                        } else {
                            q[0] = acc[j][0];
                        }
                    }
                }
            }
        } else if (nvalid > 0) {                         // ragged right edge: element by element
#pragma unroll
            for (int j = 0; j < PY; ++j) {
                const int y = yr + j;
                if (y >= y_end) break;
                const size_t o = (size_t)y * P.dst_pitch + x;
                const unsigned char* crow = src + (size_t)y * P.src_pitch * elem;
#pragma unroll
                for (int i = 0; i < PX; ++i)
                    if (i < nvalid) {
                        g[o + i] = acc[j][i];
                        if (d) __stcs(d + o + i, load_src<SRC>(crow, (size_t)(x + i)) - acc[j][i]);
                    }
                if (dec && (y & 1) == 0) {
                    const int dy = y >> 1, dx = x >> 1;
                    if (dy < P.dec_H && dx < P.dec_W) {
                        float* q = dec + (size_t)dy * P.dec_pitch + dx;
                        q[0] = acc[j][0];
                        if constexpr (PX == 4) {
                            if (dx + 1 < P.dec_W && nvalid > 2) q[1] = acc[j][2];
                        }
                    }
                }
            }
        }
        // carry the last 2R row-pass rows to the top: rows [TH, TH+2R) -> [0, 2R)   (disjoint since 2R <= TH).
        // Only the first CW row blocks' column passes read the destination rows (rb*PY < 2R), so only their warps meet
        // at a named barrier and do the copy; the other warps go straight on to wait for the next step's rows.  The
        // source rows are not written before the next row pass, which every warp enters through the full barrier
        // at the top of the loop.
        constexpr int CW = (2 * R + PY - 1) / PY;        // row blocks whose column pass reads rows [0, 2R)
        static_assert(CW * TPB <= CONV_THREADS && (CW * TPB) % 32 == 0, "carry warps");
        if (k + 1 < nsteps && rb < CW) {
            asm volatile("bar.sync 1, %0;" ::"n"(CW * TPB) : "memory");
            for (int c = tid; c < 2 * R * (CONV_TW / 4); c += CW * TPB) {
                const int rr = c / (CONV_TW / 4), q = c - rr * (CONV_TW / 4);
                *reinterpret_cast<float4*>(sT + (size_t)rr * PT + 4 * q) = *reinterpret_cast<const float4*>(sT + (size_t)(TH + rr) * PT + 4 * q);
            }
        }
    }
    if (lost && P.timeout_mark && tid == 0) *P.timeout_mark = 0xC0000000u | (unsigned)blockIdx.x;   // surfaces in sspyr_sync
    // Level chaining: this (strip, segment) is written -- count the build for the next level's CTAs.
    if (P.seg_pub) {
        __syncthreads();
        if (tid == 0) {
            asm volatile("fence.proxy.async.global;" ::: "memory");
            __threadfence();
            // a neighbour GPU acquires the counters of the band's first segment row (seg_sys & 1: a band above exists) and
            // of the rows that hold its last 12 rows (seg_sys & 2: a band below exists): those are released at system scope
            if (((P.seg_sys & 1) && seg == 0) || ((P.seg_sys & 2) && (seg + 1) * seg_rows + 12 > P.H)) {
                __threadfence_system();
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(P.seg_pub + seg_idx), "r"(seg_next) : "memory");
            } else {
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(P.seg_pub + seg_idx), "r"(seg_next) : "memory");
            }
        }
    }
    // Fused completion signal: the last CTA of the grid publishes "this level of this octave is written"
    // (system-scope release after a device-scope count of finished CTAs) for the neighbours to acquire.
    if (P.signal_flag) {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            const unsigned total = gridDim.x * gridDim.y * gridDim.z;
            if (atomicAdd(P.done_count, 1u) == total - 1) {
                *P.done_count = 0;                       // ready for the next launch on this octave's stream
                __threadfence_system();
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(P.signal_flag), "r"(band_epoch + P.signal_value) : "memory");
            }
        }
    }
}

template <int R, int SRC, bool TMA>
cudaError_t launch_march_one(const ConvParams& P, cudaStream_t st, int device, int frames, const CUtensorMap& tmap,
                             int seg_rows, bool pdl) {
    constexpr size_t smem = strip_smem_bytes<R>();
    static_assert(smem + 1024 <= (228 * 1024) / STRIP_CTAS_PER_SM, "the segmentation counts on at least 4 CTAs per SM");
    static bool configured[64] = {false};
    if (device < 0 || device >= 64 || !configured[device]) {
        cudaError_t e = cudaFuncSetAttribute(conv_strip_kernel<R, SRC, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (device >= 0 && device < 64) configured[device] = true;
    }
    const int strips = (P.W + CONV_TW - 1) / CONV_TW;
    const int nseg = (P.H + seg_rows - 1) / seg_rows;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(strips, nseg, frames);
    cfg.blockDim = dim3(CONV_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, conv_strip_kernel<R, SRC, TMA>, P, seg_rows, tmap);
}

// tmap: tensor map of the source plane (box = PIN columns x 32 rows x 1 frame) or nullptr -> cp.async staging
template <int R>
cudaError_t launch_march_src(const ConvParams& P, int src_kind, cudaStream_t st, int device, int frames,
                             const CUtensorMap* tmap, int seg_rows, bool pdl) {
    static const CUtensorMap none{};
    switch (src_kind) {
        case SSPYR_PIXEL_I32: return launch_march_one<R, SSPYR_PIXEL_I32, false>(P, st, device, frames, none, seg_rows, pdl);
        case SSPYR_PIXEL_F32: return launch_march_one<R, SSPYR_PIXEL_F32, false>(P, st, device, frames, none, seg_rows, pdl);
        case SSPYR_PIXEL_U8: return launch_march_one<R, SSPYR_PIXEL_U8, false>(P, st, device, frames, none, seg_rows, pdl);
        default:
            return tmap ? launch_march_one<R, CONV_SRC_PLANE, true>(P, st, device, frames, *tmap, seg_rows, pdl)
                        : launch_march_one<R, CONV_SRC_PLANE, false>(P, st, device, frames, none, seg_rows, pdl);
    }
}

}  // namespace

}  // namespace sspyr
