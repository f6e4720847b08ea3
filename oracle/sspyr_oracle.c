/* oracle/sspyr_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's scale-space hot path
 * (ZhangShuui/SIFT-parallel-optimization, serial header GuassDePyramid.h), generalised from the
 * header's square / all-octaves / sigma=2 case to rectangular H x W images, a chosen octave count,
 * any S and sigma0, float or int pixels and row bands.  Every function cites the reference lines it
 * follows.  Only tests/, oracle/make_golden.py, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may load this; the product (sspyr C-ABI + CUDA kernels) never
 * does and has no CPU fallback.
 *
 * PARITY PIN: REF mode is pinned -- tests/test_oracle.py proves orc_ref_* == the unmodified header
 * (oracle/_ref/libsiftref.so) BIT FOR BIT on square inputs, and tests/golden/ holds vectors the
 * header itself produced (oracle/make_golden.py).  CONV mode ("parity unpinned"): the reference
 * contains no convolution at all (its GaussFilter is a pointwise window, GuassDePyramid.h:122-131),
 * so orc_conv_* is the specification itself, frozen in DESIGN.md, with no upstream vector to pin it.
 *
 * Dense layouts (planes back to back, row-major, no padding), H_o = H>>o, W_o = W>>o:
 *     gauss  [o][s < S+3][H_o][W_o]       Gaussian (REF: window-multiplied) levels
 *     dog    [o][s < S+2][H_o][W_o]       DoG_s = G_s - G_{s+1}
 *     inplace[o][s < S+3][H_o][W_o]       the reference's own result: slots 0..S+1 DoG, slot S+2 = G_{S+2}
 *
 * Build: gcc -O2 -std=gnu11 -fopenmp -ffp-contract=off -fPIC -shared  (NO -ffast-math, NO FMA
 * contraction: the header's arithmetic is two separately rounded multiplies and one subtract).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* GuassDePyramid.h:7-8 -- `const float sigma = 2.0; const float PI = 3.1414926;` (sic).
 * Written as double literals narrowed to float, exactly as the header does. */
static const float ORC_SIGMA_REF = 2.0;
static const float ORC_PI_REF = 3.1414926;

float orc_sigma_ref(void) { return ORC_SIGMA_REF; }
float orc_pi_ref(void) { return ORC_PI_REF; }

/* ------------------------------------------------------------------------------------------------
 * Integer geometry.  GuassDePyramid.h:48-53 (layer = floor(log2 len)+1), :66 (side = length/step).
 * Rectangular rule: count halvings of min(H,W) so that every octave keeps >= 1 pixel on both axes.
 * ---------------------------------------------------------------------------------------------- */
int orc_octaves_all(int h, int w) {
    int len = h < w ? h : w, x = 0;
    while (len) { x++; len /= 2; }
    return x;
}

void orc_level_dims(int h, int w, int o, int* ho, int* wo) {
    *ho = h >> o;
    *wo = w >> o;
}

long long orc_plane_floats(int h, int w, int octaves) {
    long long n = 0;
    for (int o = 0; o < octaves; ++o) n += (long long)(h >> o) * (w >> o);
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * K1 -- the window table of one (octave, level, axis).  GuassDePyramid.h:107-121:
 *     float len=length; while(theLayer--) len/=2;  int MyLen=len;  len=(len-1)/2;
 *     float sig=sigma/(i+1);
 *     filter[i] = exp(-(i-len)*(i-len)/(2*sig*sig))/(sig*sqrt(2*PI));
 * `exp`/`sqrt` on float arguments resolve to the float overloads (<math.h> under g++), i.e.
 * expf/sqrtf.  Note the FLOAT halving: for axis 1080, octave 4: len = 67.5, MyLen = 67,
 * centre = 33.25 (every non-serial variant of the reference would use 33.0).
 * Writes f[0 .. axis_len>>o) and returns that count.
 * ---------------------------------------------------------------------------------------------- */
int orc_window(int axis_len, int o, int s, float sigma0, float* f) {
    float len = (float)axis_len;
    for (int t = o; t != 0; --t) len /= 2;
    int mylen = (int)len;
    len = (len - 1) / 2;
    float sig = sigma0 / (s + 1);
    for (int i = 0; i < mylen; ++i)
        f[i] = expf(-(i - len) * (i - len) / (2 * sig * sig)) / (sig * sqrtf(2 * ORC_PI_REF));
    return mylen;
}

/* ------------------------------------------------------------------------------------------------
 * REF mode, line-by-line mirror of the serial header (slow, cache-hostile like the original; used
 * to pin the restatement and as the single-core "port" CPU baseline).
 *   K0  GaussPyInit   GuassDePyramid.h:76-86    every level := (float) data[r<<o][c<<o]
 *   K2  row sweep     :122-126                  L[r][c] *= f[c]
 *   K3  column sweep  :127-131                  L[k][j] *= f[k]   (walks DOWN the rows, as written)
 *   K4  DoG           :140-146                  L[s] -= L[s+1], s ascending
 * out = inplace layout.  do_dog = 0 stops after K3 (what a second instance on which only the public
 * GaussFilter(o) was called holds): then out = gauss layout.
 * Pixels: img_i32 (the header's int**) or img_f32 (superset: already-float pixels); exactly one is
 * non-NULL.  pitch in elements.  Rows [row0, row0+h) of a full_h-row image are processed; the row
 * window is taken from the FULL axis (row band support; row0 must be a multiple of 2^(octaves-1)).
 * ---------------------------------------------------------------------------------------------- */
int orc_ref_mirror(const int32_t* img_i32, const float* img_f32, size_t pitch, int h, int w,
                   int octaves, int S, float sigma0, int row0, int full_h, int do_dog, float* out) {
    if ((img_i32 == NULL) == (img_f32 == NULL)) return -1;
    if (full_h <= 0) full_h = h;
    const int nl = S + 3;
    float* fw = (float*)malloc(sizeof(float) * (size_t)(w > 0 ? w : 1));
    float* fh = (float*)malloc(sizeof(float) * (size_t)(full_h > 0 ? full_h : 1));
    if (!fw || !fh) { free(fw); free(fh); return -2; }
    float* oct = out;
    for (int o = 0; o < octaves; ++o) {
        const int ho = h >> o, wo = w >> o, step = 1 << o, r0 = row0 >> o;
        const size_t plane = (size_t)ho * wo;
        /* K0 */
        for (int s = 0; s < nl; ++s) {
            float* L = oct + s * plane;
            for (int k = 0; k < ho; ++k)
                for (int l = 0; l < wo; ++l)
                    L[(size_t)k * wo + l] = img_i32 ? (float)img_i32[(size_t)(k * step) * pitch + (size_t)l * step]
                                                    : img_f32[(size_t)(k * step) * pitch + (size_t)l * step];
        }
        /* GaussFilter(o) */
        for (int s = 0; s < nl; ++s) {
            float* L = oct + s * plane;
            orc_window(w, o, s, sigma0, fw);
            orc_window(full_h, o, s, sigma0, fh);
            for (int j = 0; j < ho; ++j)
                for (int k = 0; k < wo; ++k) L[(size_t)j * wo + k] *= fw[k];
            for (int j = 0; j < wo; ++j)
                for (int k = 0; k < ho; ++k) L[(size_t)k * wo + j] *= fh[r0 + k];
        }
        /* K4 */
        if (do_dog)
            for (int s = 0; s < S + 2; ++s) {
                float* L = oct + s * plane;
                const float* N = oct + (s + 1) * plane;
                for (size_t i = 0; i < plane; ++i) L[i] -= N[i];
            }
        oct += (size_t)nl * plane;
    }
    free(fw);
    free(fh);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * REF mode, closed form:  G_s(r,c) = ((p * fW_s,o[c]) * fH_s,o[r]),
 * DoG_s = G_s - G_{s+1}.  Same two roundings in the same order as K2 then K3, so it equals the
 * mirror bit for bit; OpenMP over rows.  Any of gauss / dog / inplace may be NULL.
 * ---------------------------------------------------------------------------------------------- */
int orc_ref_build(const int32_t* img_i32, const float* img_f32, size_t pitch, int h, int w,
                  int octaves, int S, float sigma0, int row0, int full_h, float* gauss, float* dog,
                  float* inplace, int threads) {
    if ((img_i32 == NULL) == (img_f32 == NULL)) return -1;
    if (full_h <= 0) full_h = h;
    const int nl = S + 3;
    float* fw = (float*)malloc(sizeof(float) * (size_t)nl * (size_t)(w > 0 ? w : 1));
    float* fh = (float*)malloc(sizeof(float) * (size_t)nl * (size_t)(full_h > 0 ? full_h : 1));
    if (!fw || !fh) { free(fw); free(fh); return -2; }
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif
    size_t g_off = 0, d_off = 0, i_off = 0;
    for (int o = 0; o < octaves; ++o) {
        const int ho = h >> o, wo = w >> o, step = 1 << o, r0 = row0 >> o;
        const int fho = full_h >> o;
        const size_t plane = (size_t)ho * wo;
        for (int s = 0; s < nl; ++s) {
            orc_window(w, o, s, sigma0, fw + (size_t)s * wo);
            orc_window(full_h, o, s, sigma0, fh + (size_t)s * fho);
        }
#pragma omp parallel for num_threads(threads) schedule(static)
        for (int r = 0; r < ho; ++r) {
            for (int c = 0; c < wo; ++c) {
                const size_t src = (size_t)(r * step) * pitch + (size_t)c * step;
                const float p = img_i32 ? (float)img_i32[src] : img_f32[src];
                float prev = 0.0f;
                for (int s = 0; s < nl; ++s) {
                    float g = p * fw[(size_t)s * wo + c];
                    g = g * fh[(size_t)s * fho + r0 + r];
                    const size_t px = (size_t)r * wo + c;
                    if (gauss) gauss[g_off + s * plane + px] = g;
                    if (s > 0) {
                        const float d = prev - g;
                        if (dog) dog[d_off + (s - 1) * plane + px] = d;
                        if (inplace) inplace[i_off + (s - 1) * plane + px] = d;
                    }
                    if (s == nl - 1 && inplace) inplace[i_off + s * plane + px] = g;
                    prev = g;
                }
            }
        }
        g_off += (size_t)nl * plane;
        d_off += (size_t)(S + 2) * plane;
        i_off += (size_t)nl * plane;
    }
    free(fw);
    free(fh);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * REF mode, threaded three-sweep "port" for CPU timing with all host threads: the reference's own
 * stage structure (K0 materialise -> K2 row sweep -> K3 per-row broadcast multiply as in
 * GaussDePyramid-AVX512xPTHREAD.h:234-241 -> K4 subtract), rows split over OpenMP threads as in
 * GaussDePyramid-OpenMP.h:232-247.  work = caller scratch of (S+3)*sum(H_o*W_o) floats, receives
 * the inplace layout.  include_init: run K0 inside (1) or assume `work` already holds K0 (0).
 * ---------------------------------------------------------------------------------------------- */
int orc_ref_sweeps_mt(const int32_t* img, size_t pitch, int h, int w, int octaves, int S,
                      float sigma0, float* work, int include_init, int threads) {
    const int nl = S + 3;
    float* fw = (float*)malloc(sizeof(float) * (size_t)(w > 0 ? w : 1));
    float* fh = (float*)malloc(sizeof(float) * (size_t)(h > 0 ? h : 1));
    if (!fw || !fh) { free(fw); free(fh); return -2; }
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif
    float* oct = work;
    for (int o = 0; o < octaves; ++o) {
        const int ho = h >> o, wo = w >> o, step = 1 << o;
        const size_t plane = (size_t)ho * wo;
        if (include_init)
            for (int s = 0; s < nl; ++s) {
                float* L = oct + s * plane;
#pragma omp parallel for num_threads(threads) schedule(static)
                for (int r = 0; r < ho; ++r)
                    for (int c = 0; c < wo; ++c)
                        L[(size_t)r * wo + c] = (float)img[(size_t)(r * step) * pitch + (size_t)c * step];
            }
        for (int s = 0; s < nl; ++s) {
            float* L = oct + s * plane;
            orc_window(w, o, s, sigma0, fw);
            orc_window(h, o, s, sigma0, fh);
#pragma omp parallel for num_threads(threads) schedule(static)
            for (int r = 0; r < ho; ++r) {
                float* row = L + (size_t)r * wo;
                for (int c = 0; c < wo; ++c) row[c] *= fw[c];
            }
#pragma omp parallel for num_threads(threads) schedule(static)
            for (int r = 0; r < ho; ++r) {
                float* row = L + (size_t)r * wo;
                const float f = fh[r];
                for (int c = 0; c < wo; ++c) row[c] *= f;
            }
        }
        for (int s = 0; s < S + 2; ++s) {
            float* L = oct + s * plane;
            const float* N = oct + (s + 1) * plane;
#pragma omp parallel for num_threads(threads) schedule(static)
            for (int r = 0; r < ho; ++r)
                for (int c = 0; c < wo; ++c) L[(size_t)r * wo + c] -= N[(size_t)r * wo + c];
        }
        oct += (size_t)nl * plane;
    }
    free(fw);
    free(fh);
    return 0;
}

/* ================================================================================================
 * CONV mode -- the true separable Gaussian scale space the north_star describes.  NOT in the
 * reference ("parity unpinned"); this function IS the specification ("CONV mode").
 * It keeps every convention of the reference that still applies: level count S+3 and DoG count S+2
 * (GuassDePyramid.h:64,140), DoG sign and slot order G_s - G_{s+1} (:143), octave sides H>>o, W>>o
 * (:66), top-left / even-phase decimation (:80).
 *
 *   sigma_s      = sigma0 * 2^(s/S)                      s = 0..S+2     (absolute, per-octave pixel units)
 *   octave 0     : G_0 = I (*) g( sqrt(max(sigma0^2 - sigma_in^2, 0.01)) )
 *   s >= 1       : G_s = G_{s-1} (*) g( sqrt(sigma_s^2 - sigma_{s-1}^2) )          (incremental)
 *   octave o+1   : G_0(r,c) = G_S of octave o at (2r, 2c)                           (no extra blur)
 *   taps         : radius R = ceil(radius_sigmas * sigma_inc), w[k] = exp(-k^2/(2 sigma_inc^2)) in
 *                  double, normalised to sum 1 in double, rounded to float (orc_conv_taps)
 *   border       : clamp to edge (replicate) on both axes
 *   pass order   : row pass (along c) into a float intermediate, then column pass (along r)
 *   accumulation : double in this oracle (the GPU accumulates in fp32 FMA; tolerance 1e-4 of full scale)
 * ============================================================================================== */
double orc_conv_sigma_inc(int s, int S, float sigma0, float sigma_in) {
    if (s == 0) {
        double d = (double)sigma0 * sigma0 - (double)sigma_in * sigma_in;
        if (d < 0.01) d = 0.01;
        return sqrt(d);
    }
    const double k = pow(2.0, 1.0 / (double)S);
    const double prev = (double)sigma0 * pow(k, (double)(s - 1));
    const double tot = prev * k;
    return sqrt(tot * tot - prev * prev);
}

int orc_conv_radius(double sigma_inc, float radius_sigmas) {
    int r = (int)ceil((double)radius_sigmas * sigma_inc);
    return r < 1 ? 1 : r;
}

/* taps[0..2R] (centre at R); returns R. */
int orc_conv_taps(double sigma_inc, float radius_sigmas, float* taps) {
    const int R = orc_conv_radius(sigma_inc, radius_sigmas);
    double sum = 0.0;
    for (int k = -R; k <= R; ++k) sum += exp(-(double)k * k / (2.0 * sigma_inc * sigma_inc));
    for (int k = -R; k <= R; ++k)
        taps[k + R] = (float)(exp(-(double)k * k / (2.0 * sigma_inc * sigma_inc)) / sum);
    return R;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void conv_blur(const float* src, float* dst, float* tmp, int h, int w, const float* taps,
                      int R, int threads) {
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int r = 0; r < h; ++r)
        for (int c = 0; c < w; ++c) {
            double acc = 0.0;
            for (int k = -R; k <= R; ++k)
                acc += (double)taps[k + R] * (double)src[(size_t)r * w + clampi(c + k, 0, w - 1)];
            tmp[(size_t)r * w + c] = (float)acc;
        }
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int r = 0; r < h; ++r)
        for (int c = 0; c < w; ++c) {
            double acc = 0.0;
            for (int k = -R; k <= R; ++k)
                acc += (double)taps[k + R] * (double)tmp[(size_t)clampi(r + k, 0, h - 1) * w + c];
            dst[(size_t)r * w + c] = (float)acc;
        }
}

/* gauss (required) and dog (optional) in the dense layouts above. */
int orc_conv_build(const int32_t* img_i32, const float* img_f32, size_t pitch, int h, int w,
                   int octaves, int S, float sigma0, float sigma_in, float radius_sigmas,
                   float* gauss, float* dog, int threads) {
    if ((img_i32 == NULL) == (img_f32 == NULL) || !gauss || S < 1) return -1;
    const int nl = S + 3;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif
    float* base = (float*)malloc(sizeof(float) * (size_t)h * w);
    float* tmp = (float*)malloc(sizeof(float) * (size_t)h * w);
    float taps[2 * 64 + 1];
    if (!base || !tmp) { free(base); free(tmp); return -2; }
    for (int r = 0; r < h; ++r)
        for (int c = 0; c < w; ++c)
            base[(size_t)r * w + c] = img_i32 ? (float)img_i32[(size_t)r * pitch + c] : img_f32[(size_t)r * pitch + c];
    size_t g_off = 0, d_off = 0, prev_g_off = 0;
    for (int o = 0; o < octaves; ++o) {
        const int ho = h >> o, wo = w >> o;
        const size_t plane = (size_t)ho * wo;
        float* G = gauss + g_off;
        if (o == 0) {
            const double si = orc_conv_sigma_inc(0, S, sigma0, sigma_in);
            if (orc_conv_radius(si, radius_sigmas) > 64) { free(base); free(tmp); return -3; }
            const int R = orc_conv_taps(si, radius_sigmas, taps);
            conv_blur(base, G, tmp, ho, wo, taps, R, threads);
        } else {
            const int hp = h >> (o - 1), wp = w >> (o - 1);
            const float* prevS = gauss + prev_g_off + (size_t)S * ((size_t)hp * wp);
            for (int r = 0; r < ho; ++r)
                for (int c = 0; c < wo; ++c) G[(size_t)r * wo + c] = prevS[(size_t)(2 * r) * wp + 2 * c];
        }
        for (int s = 1; s < nl; ++s) {
            const double si = orc_conv_sigma_inc(s, S, sigma0, sigma_in);
            if (orc_conv_radius(si, radius_sigmas) > 64) { free(base); free(tmp); return -3; }
            const int R = orc_conv_taps(si, radius_sigmas, taps);
            conv_blur(G + (size_t)(s - 1) * plane, G + (size_t)s * plane, tmp, ho, wo, taps, R, threads);
        }
        if (dog)
            for (int s = 0; s < S + 2; ++s)
                for (size_t i = 0; i < plane; ++i)
                    dog[d_off + s * plane + i] = G[s * plane + i] - G[(s + 1) * plane + i];
        prev_g_off = g_off;
        g_off += (size_t)nl * plane;
        d_off += (size_t)(S + 2) * plane;
    }
    free(base);
    free(tmp);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * 26-neighbour DoG extremum flags (not in the reference, which stops at
 * DoG, GuassDePyramid.h:136-149).  For one octave: dog = [S+2][ho][wo]; flags = [S][ho][wo] uint8,
 * flags[s-1][r][c] = 1 if |v| > thresh and v is strictly greater (or strictly smaller) than all 26
 * neighbours in levels s-1, s, s+1 (s = 1..S), interior pixels only (1-pixel border excluded).
 * ---------------------------------------------------------------------------------------------- */
int orc_extrema_octave(const float* dog, int S, int ho, int wo, float thresh, uint8_t* flags) {
    const size_t plane = (size_t)ho * wo;
    memset(flags, 0, (size_t)S * plane);
    for (int s = 1; s <= S; ++s)
        for (int r = 1; r < ho - 1; ++r)
            for (int c = 1; c < wo - 1; ++c) {
                const float v = dog[s * plane + (size_t)r * wo + c];
                if (!(fabsf(v) > thresh)) continue;
                int is_max = 1, is_min = 1;
                for (int ds = -1; ds <= 1; ++ds)
                    for (int dr = -1; dr <= 1; ++dr)
                        for (int dc = -1; dc <= 1; ++dc) {
                            if (!ds && !dr && !dc) continue;
                            const float n = dog[(s + ds) * plane + (size_t)(r + dr) * wo + (c + dc)];
                            if (!(v > n)) is_max = 0;
                            if (!(v < n)) is_min = 0;
                        }
                flags[(s - 1) * plane + (size_t)r * wo + c] = (uint8_t)(is_max | is_min);
            }
    return 0;
}

/* FNV-1a 64 over raw bytes: the fingerprint the golden fixtures use for planes too big to commit. */
uint64_t orc_fnv1a64(const void* data, size_t nbytes) {
    const uint8_t* p = (const uint8_t*)data;
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < nbytes; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}
