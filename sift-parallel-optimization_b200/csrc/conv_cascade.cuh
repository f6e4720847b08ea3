// csrc/conv_cascade.cuh -- CONV mode, whole pyramid in ONE launch: every (octave, level, segment, strip) work item
// of a build is a CTA of the same grid, ordered so that a level's source rows are read back out of L2 instead of
// DRAM.
//
// Why (profiles/r2_conv_eval_builds_ab.txt): the one-launch-per-level schedule of conv_march.cuh moves 12 bytes per
// level-pixel through DRAM -- every Gaussian plane is written, evicted (an 8K plane is 133 MB, the L2 126 MB) and read
// back by the next level -- 3.0 GB for an 8K pyramid whose compulsory traffic (B_full) is 2.08 GB, and it
// runs at 73 % of the copy rate whatever the instruction count.  Here the levels of an octave run CONCURRENTLY, a few
// dozen rows behind one another:
//
//   * work item = the marching strip of conv_march.cuh: a 128-column strip, one vertical segment, one level; same
//     arithmetic, same order -> bit-identical planes (tests compare the two paths bitwise);
//   * block order = frame, then a host-built item table (conv_cascade.cu): items sorted by the image row at which
//     they can run -- segment + 2 per level inside an octave, the octaves interleaved so that octave o+1 follows level
//     S of octave o down the frame instead of waiting for it to finish.  Everything an item reads is produced by items
//     EARLIER in the table (the host verifies this for every item and falls back to the plain octave-major order
//     otherwise), and CTAs are dispatched in block-index order, so a waiting CTA's producers are always resident or
//     finished: no deadlock, whatever fits on the GPU;
//   * ordering is carried by one 32-bit counter per item, (build << 16) | steps finished, published after every
//     32-row step (st.release.gpu) and acquired by the consumer's first warp before it stages the rows
//     (ld.acquire.gpu + fence.proxy.async, then one TMA box).  A consumer therefore trails its producer by two steps,
//     not by a kernel boundary, and what it reads was written microseconds ago;
//   * builds of the same frame slot are ordered by a per-slot epoch (the last CTA of a build bumps it; a new build's
//     items wait for it), so consecutive launches may overlap under programmatic dependent launch whatever slots they
//     touch.  No CUDA graph, no side streams, no events: one launch per build.
//
// Taps are read from the kernel parameter block through a level index that is only known at run time (uniform loads
// into uniform registers).  Radii are padded to three classes (6, 10, 12; zero taps at both ends leave every FMA chain
// bit-identical): CTAs of different levels share an SM, and five different fully unrolled hot loops (10-14 KB each)
// thrash its instruction cache (ncu: 23 % of all stall samples were no_instruction with one body per radius).
#pragma once
#include "conv_march.cuh"

namespace sspyr {

constexpr int CASC_PIN = 156;                 // staged row pitch for every radius (4 * 39): one TMA box shape per octave
constexpr int CASC_MAX_TMA_OCT = 8;           // octaves with a tensor map in the parameter block (smaller ones use cp.async)
constexpr int CASC_MAX_R = 12;
constexpr int CASC_MAX_FRAMES = 64;           // frame slots per launch
constexpr unsigned CASC_SLOT_EPOCH = 48;      // d_flag[64 slot + 48]: builds of the slot completed
constexpr unsigned CASC_SLOT_FIN = 49;        // d_flag[64 slot + 49]: items of the running build finished

struct CascOct {
    float* base;                              // frame slot 0 of the launch: [G_0..G_{S+1} | DoG_0..DoG_{S+1} | G_{S+2}]
    unsigned* ctr;                            // frame slot 0 of the launch: item counters [level][segment][strip]
    unsigned long long plane;                 // floats per plane
    int H, W, pitch;
    int seg_rows, nsegs, nstrips;
    int first_level;                          // 0 for octave 0 (blurs the raw frame), 1 below (level 0 is the decimated base)
    unsigned seg_cap;                         // counters per level
    int tma;                                  // a tensor map for this octave exists (index = octave)
};

struct CascLevel {
    int radius;
    int pad[3];
    float taps[2 * CASC_MAX_R + 1 + 3];       // taps[k + R], k = -R..R
};

struct CascParams {
    const void* raw;                          // input frame of slot 0 of the launch
    unsigned long long raw_frame_stride;      // elements
    unsigned long long out_frame_stride;      // floats
    unsigned* slot_flags;                     // d_flag block of slot 0 of the launch (CONV_FLAG_BLOCK apart)
    unsigned* timeout_mark;
    unsigned ctr_frame_stride;
    const unsigned* item_tab;                 // block index inside a frame -> octave << 28 | level << 24 | strip << 14 | segment
    unsigned items_per_frame;                 // work items (= grid blocks) per frame
    int raw_pitch, raw_kind;
    int slot0;                                // absolute index of the launch's first slot (TMA frame coordinate)
    int octaves, nl, S, want_dog;
    unsigned short bseq[CASC_MAX_FRAMES];     // per frame of the launch: builds of that slot started before this one (mod 2^16)
    CascOct oct[SSPYR_MAX_OCTAVES];
    CascLevel lev[SSPYR_MAX_LEVELS];
};

struct CascMaps {
    CUtensorMap m[CASC_MAX_TMA_OCT];          // per octave: (pitch, H, planes, frame slots), box 156 x 32 x 1 x 1
};

constexpr size_t casc_smem_bytes() {
    return sizeof(float) * ((size_t)STRIP_TH * CASC_PIN + (size_t)(STRIP_TH + 2 * CASC_MAX_R) * conv_pitch_t()) + 16;
}

namespace {

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
                   "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Bounded spin until *p has reached `need` (wrap-safe).  A time-out marks the handle and lets the CTA go on: the
// planes are then garbage, sspyr_sync reports it, nothing hangs.
__device__ __forceinline__ void spin_until(const unsigned* p, unsigned need, unsigned* timeout_mark) {
    if ((int)(ld_acquire(p) - need) >= 0) return;
    const long long t0 = clock64();
    for (;;) {
        __nanosleep(40);
        if ((int)(ld_acquire(p) - need) >= 0) return;
        if (*reinterpret_cast<volatile unsigned*>(timeout_mark) != 0) return;          // someone gave up: do not pile up waits
        if (clock64() - t0 > 4000000000LL) { *timeout_mark = 0xA0000000u | (need & 0xffffu); return; }
    }
}

// Where an item's source rows come from: the counters of the producing level and its segment / strip geometry.
struct CascDep {
    const unsigned* ctr;                      // null: the raw frame (nothing to wait for)
    int seg_rows, nsegs, nstrips, H;
    int scale;                                // 1: same octave; 2: the source is the decimated base written by the octave above
};

// Warp 0: wait until rows [ya, yb) x columns [xa, xb) of the source (this octave's coordinates, already clamped to the
// plane) have been written in build `b16`.
__device__ __forceinline__ void casc_wait_rows(const CascDep& D, unsigned b16, int ya, int yb, int xa, int xb,
                                               unsigned* timeout_mark, int lane) {
    const int ra = D.scale * ya, rb = D.scale * (yb - 1);          // first / last source row in the producer's coordinates
    const int sa = ra / D.seg_rows, sb = min(rb / D.seg_rows, D.nsegs - 1);
    const int ca = (D.scale * xa) / CONV_TW, cb = min((D.scale * (xb - 1)) / CONV_TW, D.nstrips - 1);
    const int nc = cb - ca + 1, n = (sb - sa + 1) * nc;
    for (int i = lane; i < n; i += 32) {
        const int sg = sa + i / nc, st = ca + i % nc;
        const int seg_lo = sg * D.seg_rows, seg_hi = min(seg_lo + D.seg_rows, D.H);
        const int last = min(rb, seg_hi - 1) - seg_lo;                // last needed row inside that segment
        const unsigned steps = (unsigned)(last / STRIP_TH + 1);
        spin_until(D.ctr + (size_t)sg * D.nstrips + st, b16 + steps, timeout_mark);
    }
    __syncwarp();
    if (lane == 0) asm volatile("fence.proxy.async.global;" ::: "memory");     // the TMA loads that follow read those rows
}

// Stage rows gy0 .. gy0+NROWS-1 (clamped to the plane) of the strip at x0 into sIn: the cp.async / load+convert path
// for the raw frame, for plane edges and for octaves without a tensor map.  `kind` is a run-time value here.
template <int R, int NROWS>
__device__ __forceinline__ void casc_stage_rows(float* __restrict__ sIn, int gy0, int x0, const unsigned char* __restrict__ src,
                                                int src_pitch, int H, int W, int kind, int tid) {
    constexpr int RA = conv_ra<R>();
    constexpr int CH = (CONV_TW + 2 * RA) / 4;
    const int elem = kind == SSPYR_PIXEL_U8 ? 1 : 4;
#pragma unroll 1
    for (int c = tid; c < NROWS * CH; c += CONV_THREADS) {
        const int rr = c / CH, q = c - rr * CH;
        const int gy = min(max(gy0 + rr, 0), H - 1);
        const unsigned char* row = src + (size_t)gy * src_pitch * elem;
        float* s = sIn + rr * CASC_PIN + 4 * q;
        const int gx = x0 - RA + 4 * q;
        if (gx >= 0 && gx + 4 <= W) {
            if (kind == CONV_SRC_PLANE) {
                __pipeline_memcpy_async(s, reinterpret_cast<const float*>(row) + gx, 16);
            } else if (kind == SSPYR_PIXEL_I32) {
                const int4 t = __ldg(reinterpret_cast<const int4*>(reinterpret_cast<const int*>(row) + gx));
                *reinterpret_cast<float4*>(s) = make_float4((float)t.x, (float)t.y, (float)t.z, (float)t.w);
            } else if (kind == SSPYR_PIXEL_U8) {
                const uchar4 t = __ldg(reinterpret_cast<const uchar4*>(row + gx));
                *reinterpret_cast<float4*>(s) = make_float4((float)t.x, (float)t.y, (float)t.z, (float)t.w);
            } else {
                *reinterpret_cast<float4*>(s) = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(row) + gx));
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const size_t idx = (size_t)min(max(gx + i, 0), W - 1);
                float v;
                if (kind == CONV_SRC_PLANE) v = __ldcg(reinterpret_cast<const float*>(row) + idx);
                else if (kind == SSPYR_PIXEL_I32) v = (float)__ldg(reinterpret_cast<const int*>(row) + idx);
                else if (kind == SSPYR_PIXEL_U8) v = (float)__ldg(row + idx);
                else v = __ldg(reinterpret_cast<const float*>(row) + idx);
                s[i] = v;
            }
        }
    }
}

// One work item: level s of octave o, segment `seg`, strip `strip`, frame fz of the launch.
template <int R>
__device__ __forceinline__ void cascade_item(const CascParams& C, const CascMaps& M, int o, int s, int seg, int strip,
                                             unsigned fz, float* __restrict__ smem) {
    constexpr int TH = STRIP_TH, PIN = CASC_PIN, PT = conv_pitch_t(), RA = conv_ra<R>();
    constexpr int PX = 4, PY = 4, TPB = CONV_TW / PX;
    const CascOct& O = C.oct[o];
    const float* __restrict__ taps = C.lev[s].taps;
    float* sIn = smem;                                    // [TH][PIN]
    float* sT = smem + (size_t)TH * PIN;                  // [TH + 2R][PT]
    unsigned long long& bar = *reinterpret_cast<unsigned long long*>(smem + (size_t)TH * PIN + (size_t)(TH + 2 * CASC_MAX_R) * PT);

    const int tid = threadIdx.x, lane = tid & 31;
    const int x0 = strip * CONV_TW;
    const int y_begin = seg * O.seg_rows;
    const int y_end = min(O.H, y_begin + O.seg_rows);
    const int nsteps = (y_end - y_begin + TH - 1) / TH;
    const int H = O.H, W = O.W;

    // ---- planes --------------------------------------------------------------------------------------------------
    float* obase = O.base + (size_t)fz * C.out_frame_stride;
    float* g = obase + (size_t)(s == C.nl - 1 ? 2 * C.nl - 2 : s) * O.plane;
    float* d = (s >= 1 && C.want_dog) ? obase + (size_t)(C.nl - 1 + s - 1) * O.plane : nullptr;   // DoG_{s-1}
    float* dec = nullptr;
    int dec_H = 0, dec_W = 0, dec_pitch = 0;
    if (s == C.S && o + 1 < C.octaves) {
        const CascOct& N = C.oct[o + 1];
        dec = N.base + (size_t)fz * C.out_frame_stride;   // G_0 of the next octave
        dec_H = N.H; dec_W = N.W; dec_pitch = N.pitch;
    }
    const bool raw = o == 0 && s == 0;
    const int kind = raw ? C.raw_kind : CONV_SRC_PLANE;
    const int elem = kind == SSPYR_PIXEL_U8 ? 1 : 4;
    const int src_pitch = raw ? C.raw_pitch : O.pitch;
    const int src_plane = s - 1;                          // G_{s-1} (for s == 1 of a lower octave: the decimated base, plane 0)
    const unsigned char* src = raw ? static_cast<const unsigned char*>(C.raw) + (size_t)fz * C.raw_frame_stride * elem
                                   : reinterpret_cast<const unsigned char*>(obase + (size_t)src_plane * O.plane);
    const int dst_pitch = O.pitch;

    // ---- counters ----------------------------------------------------------------------------------------------------
    unsigned* ctr_o = O.ctr + (size_t)fz * C.ctr_frame_stride;
    unsigned* own = ctr_o + (size_t)s * O.seg_cap + (size_t)seg * O.nstrips + strip;
    unsigned* flags = C.slot_flags + (size_t)fz * CONV_FLAG_BLOCK;
    CascDep D;
    if (raw) {
        D.ctr = nullptr; D.seg_rows = 1; D.nsegs = D.nstrips = D.H = 0; D.scale = 1;
    } else if (s > O.first_level) {
        D.ctr = ctr_o + (size_t)(s - 1) * O.seg_cap; D.seg_rows = O.seg_rows; D.nsegs = O.nsegs; D.nstrips = O.nstrips; D.H = H; D.scale = 1;
    } else {                                              // decimated base: written by level S of the octave above
        const CascOct& U = C.oct[o - 1];
        D.ctr = U.ctr + (size_t)fz * C.ctr_frame_stride + (size_t)C.S * U.seg_cap;
        D.seg_rows = U.seg_rows; D.nsegs = U.nsegs; D.nstrips = U.nstrips; D.H = U.H; D.scale = 2;
    }
    const int xa = max(x0 - RA, 0), xb = min(x0 + CONV_TW + RA, W);   // source columns this strip stages

    // The build number comes from the host (C.bseq).  First: "the previous build of this frame slot is complete" --
    // nobody still reads what this item overwrites, and the item's own counter is final -- i.e. the slot's epoch has
    // reached this build's number.  All CTAs of that build were dispatched before any of this one (stream order,
    // block order), so the wait cannot deadlock.
    const unsigned b16 = (unsigned)C.bseq[fz] << 16;
    if (tid < 32) {
        if (lane == 0) {
            const unsigned want = b16 >> 16;
            if ((ld_acquire(flags + CASC_SLOT_EPOCH) & 0xffffu) != want) {
                const long long t0 = clock64();
                for (;;) {
                    __nanosleep(100);
                    if ((ld_acquire(flags + CASC_SLOT_EPOCH) & 0xffffu) == want) break;
                    if (*reinterpret_cast<volatile unsigned*>(C.timeout_mark) != 0) break;
                    if (clock64() - t0 > 4000000000LL) { *C.timeout_mark = 0xB0000000u | want; break; }
                }
            }
        }
        __syncwarp();
        if (D.ctr)                                        // warm-up rows + the first step's rows
            casc_wait_rows(D, b16, max(y_begin - R, 0), min(y_begin + R + TH, H), xa, xb, C.timeout_mark, lane);
    }
    __syncthreads();

    // a step may use TMA when its whole box lies inside the plane (clamp-to-edge is not a TMA fill mode)
    const bool tma_cols = !raw && O.tma && x0 - RA >= 0 && x0 + CONV_TW + RA <= W && x0 - RA + PIN <= src_pitch;
    unsigned phase = 0;
    bool pending_tma = false;
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    auto step_uses_tma = [&](int gy0) { return tma_cols && gy0 >= 0 && gy0 + TH <= H; };
    auto stage_step = [&](int gy0) {                     // rows gy0 .. gy0+TH-1 -> sIn
        if (step_uses_tma(gy0)) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of sIn are done
                mbar_expect_tx(&bar, (unsigned)(TH * PIN * sizeof(float)));
                tma_load_4d(sIn, &M.m[o], x0 - RA, gy0, src_plane, C.slot0 + (int)fz, &bar);
            }
            pending_tma = true;
        } else {
            casc_stage_rows<R, TH>(sIn, gy0, x0, src, src_pitch, H, W, kind, tid);
            pending_tma = false;
        }
        __pipeline_commit();
    };

    // prologue: the 2R warm-up rows go to a scratch area inside sT, the first step's rows to sIn; one wait for both
    float* scratch = sT + (size_t)(2 * R) * PT;
    static_assert((size_t)2 * R * CASC_PIN <= (size_t)STRIP_TH * conv_pitch_t(), "warm-up scratch must fit in sT");
    casc_stage_rows<R, 2 * R>(scratch, y_begin - R, x0, src, src_pitch, H, W, kind, tid);
    stage_step(y_begin + R);
    __pipeline_wait_prior(0);
    bool lost = false;
    if (pending_tma) {
        if (!mbar_wait(&bar, phase)) lost = true;
        phase ^= 1;
    }
    lost = __syncthreads_or(lost);
    if (!lost) strip_row_pass<R, 2 * R, 0, PIN>(taps, scratch, sT, tid);

    const int cq = tid % TPB, rb = tid / TPB;
    const int x = x0 + cq * PX;
    const int nvalid = W - x;
#pragma unroll 1
    for (int k = 0; k < nsteps && !lost; ++k) {
        bool miss = false;
        if (k > 0) {
            if (pending_tma) {
                miss = !mbar_wait(&bar, phase);
                phase ^= 1;
            } else {
                __pipeline_wait_prior(0);
            }
        }
        if (__syncthreads_or(miss)) { lost = true; break; }   // new rows landed; carried rows in place; every thread has stored step k-1
        if (k > 0 && tid == 0) {                         // steps 0..k-1 of this item are written: publish (release is cumulative
            st_release(own, b16 + (unsigned)k);          // over the other threads' stores, ordered before it by the barrier)
        }
        strip_row_pass<R, TH, 2 * R, PIN>(taps, sIn, sT, tid);

        // centre values for DoG_{s-1} = G_{s-1} - G_s (see conv_march.cuh): out of the staged tile, except the first R rows
        const int y0 = y_begin + k * TH;
        const int yr = y0 + rb * PY;
        float cen[PY][PX];
        if (d && nvalid >= PX) {
#pragma unroll
            for (int j = 0; j < PY; ++j) {
                const int m = rb * PY + j - R;
                if (m >= 0) {
                    const float4 t = *reinterpret_cast<const float4*>(sIn + (size_t)m * PIN + RA + cq * PX);
                    cen[j][0] = t.x; cen[j][1] = t.y; cen[j][2] = t.z; cen[j][3] = t.w;
                } else {
                    const float* crow = reinterpret_cast<const float*>(src) + (size_t)min(yr + j, H - 1) * src_pitch;
                    const float4 t = __ldcg(reinterpret_cast<const float4*>(crow + x));
                    cen[j][0] = t.x; cen[j][1] = t.y; cen[j][2] = t.z; cen[j][3] = t.w;
                }
            }
        }
        __syncthreads();
        if (k + 1 < nsteps) {                            // next step's rows: in flight during the column pass below
            const int gy0 = y_begin + R + (k + 1) * TH;
            if (D.ctr) {
                if (tid < 32) casc_wait_rows(D, b16, min(gy0, H - 1), min(gy0 + TH, H), xa, xb, C.timeout_mark, lane);
                if (!step_uses_tma(gy0)) __syncthreads();  // every thread stages: all of them wait for the first warp
            }
            stage_step(gy0);
        }

        // ---- column pass: output rows y0 + rb*PY + j from sT rows rb*PY + j .. + 2R -------------------------------
        const float* tcol = sT + (size_t)(rb * PY) * PT + cq * PX;
        f32x2 a01[PY], a23[PY];
#pragma unroll
        for (int j = 0; j < PY; ++j) { a01[j] = pk2(0.0f, 0.0f); a23[j] = pk2(0.0f, 0.0f); }
#pragma unroll
        for (int i = 0; i < PY + 2 * R; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(tcol + (size_t)i * PT);
            const f32x2 v01 = pk2(v.x, v.y), v23 = pk2(v.z, v.w);
#pragma unroll
            for (int j = 0; j < PY; ++j) {
                if (i - j >= 0 && i - j <= 2 * R) {
                    const f32x2 w = pk2(taps[i - j], taps[i - j]);
                    a01[j] = fma2(w, v01, a01[j]);
                    a23[j] = fma2(w, v23, a23[j]);
                }
            }
        }
        float acc[PY][PX];
#pragma unroll
        for (int j = 0; j < PY; ++j) {
            unpk2(a01[j], acc[j][0], acc[j][1]);
            unpk2(a23[j], acc[j][2], acc[j][3]);
        }
        if (nvalid >= PX) {                               // full group: vector stores, one running offset
            unsigned off = (unsigned)yr * (unsigned)dst_pitch + (unsigned)x;   // a plane has < 2^32 floats
#pragma unroll
            for (int j = 0; j < PY; ++j, off += (unsigned)dst_pitch) {
                const int y = yr + j;
                if (y >= y_end) break;
                *reinterpret_cast<float4*>(g + off) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
                if (d)                                    // DoG_{s-1} = G_{s-1} - G_s  (GuassDePyramid.h:143)
                    __stcs(reinterpret_cast<float4*>(d + off), make_float4(cen[j][0] - acc[j][0], cen[j][1] - acc[j][1],
                                                                          cen[j][2] - acc[j][2], cen[j][3] - acc[j][3]));
                if (dec && (y & 1) == 0) {                // even-phase decimation (GuassDePyramid.h:80)
                    const int dy = y >> 1, dx = x >> 1;
                    if (dy < dec_H && dx < dec_W) {
                        float* q = dec + (size_t)dy * dec_pitch + dx;
                        if (dx + 1 < dec_W) *reinterpret_cast<float2*>(q) = make_float2(acc[j][0], acc[j][2]);
                        else q[0] = acc[j][0];
                    }
                }
            }
        } else if (nvalid > 0) {                          // ragged right edge: element by element
#pragma unroll
            for (int j = 0; j < PY; ++j) {
                const int y = yr + j;
                if (y >= y_end) break;
                const size_t off = (size_t)y * dst_pitch + x;
                const float* crow = reinterpret_cast<const float*>(src) + (size_t)y * src_pitch;
#pragma unroll
                for (int i = 0; i < PX; ++i)
                    if (i < nvalid) {
                        g[off + i] = acc[j][i];
                        if (d) __stcs(d + off + i, __ldcg(crow + x + i) - acc[j][i]);
                    }
                if (dec && (y & 1) == 0) {
                    const int dy = y >> 1, dx = x >> 1;
                    if (dy < dec_H && dx < dec_W) {
                        float* q = dec + (size_t)dy * dec_pitch + dx;
                        q[0] = acc[j][0];
                        if (dx + 1 < dec_W && nvalid > 2) q[1] = acc[j][2];
                    }
                }
            }
        }
        // carry the last 2R row-pass rows to the top (only the warps whose column pass reads them, named barrier)
        constexpr int CW = (2 * R + PY - 1) / PY;
        static_assert(CW * TPB <= CONV_THREADS && (CW * TPB) % 32 == 0, "carry warps");
        if (k + 1 < nsteps && rb < CW) {
            asm volatile("bar.sync 1, %0;" ::"n"(CW * TPB) : "memory");
            for (int c = tid; c < 2 * R * (CONV_TW / 4); c += CW * TPB) {
                const int rr = c / (CONV_TW / 4), q = c - rr * (CONV_TW / 4);
                *reinterpret_cast<float4*>(sT + (size_t)rr * PT + 4 * q) = *reinterpret_cast<const float4*>(sT + (size_t)(TH + rr) * PT + 4 * q);
            }
        }
    }
    // ---- the item is written: next build number in the counter, one more finished item of the slot's build ------
    __syncthreads();
    if (tid == 0) {
        if (lost) *C.timeout_mark = 0xC0000000u | (unsigned)strip;
        __threadfence();
        st_release(own, b16 + 0x10000u);
        if (atomicAdd(flags + CASC_SLOT_FIN, 1u) == C.items_per_frame - 1) {     // last item of this build of the slot
            flags[CASC_SLOT_FIN] = 0;
            __threadfence();
            st_release(flags + CASC_SLOT_EPOCH, ((b16 >> 16) + 1u) & 0xffffu);
        }
    }
}

// grid.x = frames x items_per_frame; block = 256 threads; dynamic smem = casc_smem_bytes()
__global__ void __launch_bounds__(CONV_THREADS, 4)
conv_cascade_kernel(const __grid_constant__ CascParams C, const __grid_constant__ CascMaps M) {
    extern __shared__ __align__(128) float casc_smem[];
    asm volatile("griddepcontrol.launch_dependents;");    // the next build may fill SMs as this one drains (ordered by counters)
    const unsigned fz = blockIdx.x / C.items_per_frame;
    const unsigned it = __ldg(C.item_tab + (blockIdx.x - fz * C.items_per_frame));
    const int o = (int)(it >> 28), s = (int)((it >> 24) & 15u), strip = (int)((it >> 14) & 1023u), seg = (int)(it & 16383u);
    switch (C.lev[s].radius) {                            // (the padded class radius)
        case 6: cascade_item<6>(C, M, o, s, seg, strip, fz, casc_smem); break;
        case 10: cascade_item<10>(C, M, o, s, seg, strip, fz, casc_smem); break;
        case 12: cascade_item<12>(C, M, o, s, seg, strip, fz, casc_smem); break;
        default: break;
    }
    // chain completion: "this grid done" implies "the grid it overlapped with done" for whatever follows in the stream
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace

}  // namespace sspyr
