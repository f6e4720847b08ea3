// csrc/sspyr_api.cu -- the extern "C" boundary declared in include/sspyr.h: host state, window/tap
// tables, buffer layout, copies.  All compute is in ref_kernels.cu / conv_kernels.cu; there is no CPU
// implementation of the hot path anywhere in this library.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>

#include <nvtx3/nvToolsExt.h>     // header-only NVTX 3: ranges cost nothing unless a profiler injects its library

#include "conv_sched.h"
#include "sspyr_internal.h"

using namespace sspyr;

namespace {
// NVTX range around one library call (tracing): upload / build / download show up by name on the
// host timeline of Nsight Systems / Compute next to the kernels they enqueue.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace

namespace {

// rows x cols floats from a pitched device plane block to a dense host block; one linear copy when the rows are
// already contiguous (pitch == cols), which the copy engine moves a little faster than a strided 2-D copy
cudaError_t copy_planes_to_host(float* dst, const float* src, int cols, int pitch, size_t rows, cudaStream_t st) {
    if (pitch == cols) return cudaMemcpyAsync(dst, src, sizeof(float) * rows * cols, cudaMemcpyDeviceToHost, st);
    return cudaMemcpy2DAsync(dst, sizeof(float) * cols, src, sizeof(float) * pitch, sizeof(float) * cols, rows,
                             cudaMemcpyDeviceToHost, st);
}

thread_local std::string g_create_err;

int fail(sspyr_ctx* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_err = msg;
    return code;
}

int fail_cuda(sspyr_ctx* h, cudaError_t e, const char* what) {
    cudaGetLastError();   // clear sticky-less error state
    return fail(h, SSPYR_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CU(h, call)                                            \
    do {                                                       \
        cudaError_t e_ = (call);                               \
        if (e_ != cudaSuccess) return fail_cuda(h, e_, #call); \
    } while (0)

inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// The reference's mistyped pi and its sigma, narrowed from double literals exactly as the header does
// (GuassDePyramid.h:7-8).
const float kPiRef = 3.1414926;
const float kSigmaRef = 2.0;

// K1 -- window of one (axis, octave, level): GuassDePyramid.h:107-121.  Same expression, same float
// types, same libm (expf/sqrtf are what `exp`/`sqrt` resolve to there), so the table is bit-identical
// to the one the header builds; nvcc's host pass is told not to contract (-ffp-contract=off).
void ref_window(int axis_len, int o, int s, float sigma0, float* f) {
    float len = (float)axis_len;                 // :107
    for (int t = o; t != 0; --t) len /= 2;       // :109-112  (float halving: 1080 -> 67.5 at o=4)
    const int mylen = (int)len;                  // :114
    len = (len - 1) / 2;                         // :115
    const float sig = sigma0 / (s + 1);          // :118
    for (int i = 0; i < mylen; ++i)              // :119-121
        f[i] = expf(-(i - len) * (i - len) / (2 * sig * sig)) / (sig * sqrtf(2 * kPiRef));
}

int octaves_all(int h, int w) {                  // GuassDePyramid.h:48-53 on the short side
    int len = h < w ? h : w, x = 0;
    while (len) { x++; len /= 2; }
    return x;
}

// CONV tap schedule (DESIGN.md "CONV mode"): incremental sigma of level s and its normalised taps.
double conv_sigma_inc(int s, int S, float sigma0, float sigma_in) {
    if (s == 0) {
        double d = (double)sigma0 * sigma0 - (double)sigma_in * sigma_in;
        if (d < 0.01) d = 0.01;
        return std::sqrt(d);
    }
    const double k = std::pow(2.0, 1.0 / (double)S);
    const double prev = (double)sigma0 * std::pow(k, (double)(s - 1));
    const double tot = prev * k;
    return std::sqrt(tot * tot - prev * prev);
}

int conv_make_taps(double si, float radius_sigmas, std::vector<float>& taps) {
    int R = (int)std::ceil((double)radius_sigmas * si);
    if (R < 1) R = 1;
    double sum = 0.0;
    for (int k = -R; k <= R; ++k) sum += std::exp(-(double)k * k / (2.0 * si * si));
    taps.resize(2 * R + 1);
    for (int k = -R; k <= R; ++k) taps[k + R] = (float)(std::exp(-(double)k * k / (2.0 * si * si)) / sum);
    return R;
}

bool valid_frame(const sspyr_ctx* h, int frame) { return frame >= 0 && frame < h->cfg.frames; }

int plane_lookup(sspyr_ctx* h, int octave, int level, int kind, int* index) {
    if (octave < 0 || octave >= h->octaves) return fail(h, SSPYR_ERR_ARG, "octave out of range");
    const int nl = h->nl;
    int lim = kind == SSPYR_KIND_DOG ? nl - 1 : nl;
    if (kind != SSPYR_KIND_GAUSS && kind != SSPYR_KIND_DOG && kind != SSPYR_KIND_INPLACE)
        return fail(h, SSPYR_ERR_ARG, "bad plane kind");
    if (level < 0 || level >= lim) return fail(h, SSPYR_ERR_ARG, "level out of range");
    const int out = h->cfg.outputs;
    bool have;
    if (kind == SSPYR_KIND_GAUSS) have = (out & SSPYR_OUT_GAUSS) || (level == nl - 1 && (out & SSPYR_OUT_GAUSS_TOP));
    else if (kind == SSPYR_KIND_DOG) have = out & SSPYR_OUT_DOG;
    else have = level == nl - 1 ? (out & (SSPYR_OUT_GAUSS | SSPYR_OUT_GAUSS_TOP)) : (out & SSPYR_OUT_DOG);
    if (!have) return fail(h, SSPYR_ERR_STATE, "that plane is not among the configured outputs");
    *index = plane_index(nl, kind, level);
    return SSPYR_OK;
}

// ---- CONV frame lanes (sspyr_internal.h) ---------------------------------------------------------------
ConvStreams own_streams(sspyr_ctx* h) {
    return ConvStreams{h->stream, h->aux.empty() ? nullptr : h->aux.data(), h->ev_base.data(), h->ev_done.data()};
}

// A library operation that later builds must follow was enqueued on the handle's stream.
cudaError_t mark_tail(sspyr_ctx* h) {
    if (!h->ev_tail) return cudaSuccess;
    ++h->tail_seq;
    return cudaEventRecord(h->ev_tail, h->stream);
}

int lane_count(const sspyr_ctx* h) {
    if (h->cfg.mode != SSPYR_MODE_CONV || h->tune.timing) return 1;
    // (bands driven level by level from the host -- no peers attached -- never come here)
    const bool banded = h->cfg.full_height != h->cfg.height;
    if (banded && h->tune.conv_band_lanes > 3)               // explicit request for more band builds in flight than the default 3
        return std::min(std::min(h->tune.conv_band_lanes, h->cfg.frames), 16);
    return frame_lanes(h->tune.conv_lanes, h->cfg.frames, banded);
}

cudaError_t ensure_lanes(sspyr_ctx* h, int n) {
    cudaError_t e;
    if (!h->ev_tail && (e = cudaEventCreateWithFlags(&h->ev_tail, cudaEventDisableTiming)) != cudaSuccess) return e;
    if (h->slot_lane.empty()) h->slot_lane.assign(h->cfg.frames, -1);
    while ((int)h->lanes.size() < n) {
        h->lanes.emplace_back();
        sspyr_ctx::Lane& L = h->lanes.back();
        if ((e = cudaStreamCreateWithFlags(&L.main, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming)) != cudaSuccess)
            return e;
        const int na = h->octaves - 1;
        L.aux.assign(na, nullptr);
        L.ev_base.assign(na, nullptr);
        L.ev_done.assign(na, nullptr);
        for (int o = 0; o < na; ++o)
            if ((e = cudaStreamCreateWithFlags(&L.aux[o], cudaStreamNonBlocking)) != cudaSuccess ||
                (e = cudaEventCreateWithFlags(&L.ev_base[o], cudaEventDisableTiming)) != cudaSuccess ||
                (e = cudaEventCreateWithFlags(&L.ev_done[o], cudaEventDisableTiming)) != cudaSuccess)
                return e;
    }
    return cudaSuccess;
}

void destroy_lanes(sspyr_ctx* h) {
    for (auto& L : h->lanes) {
        if (L.main) cudaStreamDestroy(L.main);
        if (L.done) cudaEventDestroy(L.done);
        for (cudaStream_t st : L.aux) if (st) cudaStreamDestroy(st);
        for (cudaEvent_t ev : L.ev_base) if (ev) cudaEventDestroy(ev);
        for (cudaEvent_t ev : L.ev_done) if (ev) cudaEventDestroy(ev);
    }
    h->lanes.clear();
    if (h->ev_tail) cudaEventDestroy(h->ev_tail);
    h->ev_tail = nullptr;
}

// Whole-pyramid CONV build of slots first..first+count-1 on lane (first % lanes), joined into the handle's stream.
cudaError_t build_on_lane(sspyr_ctx* h, int first, int count, int nlanes, int* launches) {
    cudaError_t e = ensure_lanes(h, nlanes);
    if (e != cudaSuccess) return e;
    const int li = first % nlanes;
    sspyr_ctx::Lane& L = h->lanes[li];
    if (h->strict_order && (e = mark_tail(h)) != cudaSuccess) return e;        // follow everything on the stream
    if (L.seen_tail != h->tail_seq) {
        if ((e = cudaStreamWaitEvent(L.main, h->ev_tail, 0)) != cudaSuccess) return e;
        L.seen_tail = h->tail_seq;
    }
    for (int f = first; f < first + count; ++f) {                               // slots last written through another lane
        const int k = h->slot_lane[f];
        if (k >= 0 && k != li && k < (int)h->lanes.size() &&
            (e = cudaStreamWaitEvent(L.main, h->lanes[k].done, 0)) != cudaSuccess)
            return e;
        h->slot_lane[f] = li;
    }
    const ConvStreams cs{L.main, L.aux.empty() ? nullptr : L.aux.data(), L.ev_base.data(), L.ev_done.data()};
    if ((e = launch_conv_graphed(h, first, count, launches, cs)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(L.done, L.main)) != cudaSuccess) return e;
    return cudaStreamWaitEvent(h->stream, L.done, 0);
}

// DoG extremum scan of slots first .. first+count-1 (modulo the slot count), one launch per contiguous run.
cudaError_t scan_extrema(sspyr_ctx* h, int first, int count, int* launches) {
    if (!(h->cfg.outputs & (SSPYR_OUT_EXTREMA | SSPYR_OUT_KEYPOINTS))) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    int done = 0;
    while (done < count && e == cudaSuccess) {
        const int f0 = (first + done) % h->cfg.frames;
        const int n = std::min(count - done, h->cfg.frames - f0);
        e = launch_extrema(h, f0, n, launches);
        done += n;
    }
    return e;
}

}  // namespace

extern "C" {

int sspyr_version(void) { return SSPYR_VERSION; }

int sspyr_default_config(sspyr_config* cfg) {
    if (!cfg) return SSPYR_ERR_ARG;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->S = 3;
    cfg->mode = SSPYR_MODE_REF;
    cfg->outputs = SSPYR_OUT_ALL;
    cfg->pixel_type = SSPYR_PIXEL_I32;
    cfg->frames = 1;
    cfg->device = -1;
    cfg->sigma_in = 0.5f;
    cfg->radius_sigmas = 3.0f;
    cfg->extrema_thresh = 0.0f;
    return SSPYR_OK;
}

int sspyr_create(const sspyr_config* cfg_in, sspyr_handle* out) {
    if (!cfg_in || !out) return fail(nullptr, SSPYR_ERR_ARG, "null argument");
    *out = nullptr;
    sspyr_config cfg = *cfg_in;
    if (cfg.height < 1 || cfg.width < 1) return fail(nullptr, SSPYR_ERR_ARG, "height and width must be >= 1");
    if (cfg.S < 0 || cfg.S + 3 > SSPYR_MAX_LEVELS) return fail(nullptr, SSPYR_ERR_ARG, "S out of range");
    if (cfg.mode != SSPYR_MODE_REF && cfg.mode != SSPYR_MODE_CONV) return fail(nullptr, SSPYR_ERR_ARG, "bad mode");
    if (cfg.mode == SSPYR_MODE_CONV && cfg.S < 1) return fail(nullptr, SSPYR_ERR_ARG, "CONV mode needs S >= 1");
    if (cfg.pixel_type < 0 || cfg.pixel_type > SSPYR_PIXEL_U8) return fail(nullptr, SSPYR_ERR_ARG, "bad pixel type");
    if (cfg.frames < 1) cfg.frames = 1;
    if (cfg.outputs == 0) cfg.outputs = SSPYR_OUT_ALL;
    if (cfg.full_height <= 0) cfg.full_height = cfg.height;
    if (cfg.band_row0 < 0 || cfg.band_row0 + cfg.height > cfg.full_height)
        return fail(nullptr, SSPYR_ERR_ARG, "row band outside the full image");
    if (cfg.sigma0 <= 0.0f) cfg.sigma0 = cfg.mode == SSPYR_MODE_REF ? kSigmaRef : 1.6f;
    if (cfg.sigma_in < 0.0f) cfg.sigma_in = 0.5f;
    if (cfg.radius_sigmas <= 0.0f) cfg.radius_sigmas = 3.0f;
    const int all = octaves_all(cfg.full_height, cfg.width);
    if (cfg.octaves <= 0) cfg.octaves = all;
    if (cfg.octaves > all || cfg.octaves > SSPYR_MAX_OCTAVES)
        return fail(nullptr, SSPYR_ERR_ARG, "more octaves than floor(log2(min(H,W)))+1");
    const bool banded = cfg.full_height != cfg.height;
    if (banded) {
        const int align = 1 << (cfg.octaves - 1);
        if (cfg.band_row0 % align) return fail(nullptr, SSPYR_ERR_ARG, "band_row0 must be a multiple of 2^(octaves-1)");
        if (cfg.band_row0 + cfg.height != cfg.full_height && cfg.height % align)
            return fail(nullptr, SSPYR_ERR_ARG, "interior band heights must be multiples of 2^(octaves-1)");
    }
    const int scan = cfg.outputs & (SSPYR_OUT_EXTREMA | SSPYR_OUT_KEYPOINTS);
    if (scan && !(cfg.outputs & SSPYR_OUT_DOG))
        return fail(nullptr, SSPYR_ERR_ARG, "SSPYR_OUT_EXTREMA / SSPYR_OUT_KEYPOINTS need SSPYR_OUT_DOG");
    if (cfg.max_keypoints < 0) return fail(nullptr, SSPYR_ERR_ARG, "max_keypoints must be >= 0");
    if (scan && banded)         // the scan treats the handle's first/last row as the image border
        return fail(nullptr, SSPYR_ERR_UNSUPPORTED, "SSPYR_OUT_EXTREMA is not available on a row-band handle (band seams would be "
                                                    "scanned as image borders); scan the bands' DoG planes with a whole-frame handle");
    if (cfg.mode == SSPYR_MODE_CONV) cfg.outputs |= SSPYR_OUT_GAUSS;   // the blur chain reads its own levels

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, SSPYR_ERR_CUDA,
                    std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e));
    }
    sspyr_ctx* h = new (std::nothrow) sspyr_ctx();
    if (!h) return fail(nullptr, SSPYR_ERR_NOMEM, "host allocation failed");
    auto bail = [&](int code, const std::string& msg) {
        g_create_err = msg;
        sspyr_destroy(h);
        return code;
    };
    if (cfg.device < 0) {
        if ((e = cudaGetDevice(&cfg.device)) != cudaSuccess) return bail(SSPYR_ERR_CUDA, cudaGetErrorString(e));
    } else if (cfg.device >= ndev) {
        return bail(SSPYR_ERR_ARG, "device ordinal out of range");
    }
    if ((e = cudaSetDevice(cfg.device)) != cudaSuccess) return bail(SSPYR_ERR_CUDA, cudaGetErrorString(e));
    h->cfg = cfg;
    h->device = cfg.device;
    h->octaves = cfg.octaves;
    h->nl = cfg.S + 3;
    h->elem_bytes = cfg.pixel_type == SSPYR_PIXEL_U8 ? 1 : 4;
    h->ext_in.assign(cfg.frames, nullptr);
    h->ext_pitch.assign(cfg.frames, 0);
    h->built.assign(cfg.frames, 0);

    // ---- output layout ----
    const int nl = h->nl, S = cfg.S;
    size_t off = 0, tab = 0, ext = 0;
    for (int o = 0; o < h->octaves; ++o) {
        OctGeom& g = h->oct[o];
        g.H = cfg.height >> o;
        g.W = cfg.width >> o;
        if (g.H < 1 || g.W < 1) {
            // a short last band can run out of rows before the full image does
            return bail(SSPYR_ERR_ARG, "band too short for the requested octave count");
        }
        g.pitch = (int)round_up((size_t)g.W, 32);
        g.plane = (size_t)g.H * g.pitch;
        if (g.plane >= (1ull << 32)) return bail(SSPYR_ERR_UNSUPPORTED, "a level plane exceeds 2^32 floats");
        g.off = off;
        off += (size_t)(2 * nl - 1) * g.plane;
        g.ext_off = ext;
        ext += (size_t)(S > 0 ? S : 0) * g.plane;
        if (cfg.mode == SSPYR_MODE_REF) {
            g.fw_off = tab;
            tab += (size_t)nl * g.pitch;
            g.fh_off = tab;
            tab += (size_t)(nl <= 8 ? 8 : 16) * g.H;      // transposed [H][8] ([H][16] for more than 8 levels)
        }
    }
    h->frame_floats = round_up(off, 64);
    h->ext_frame_bytes = round_up(ext, 256);

    // ---- tables ----
    if (cfg.mode == SSPYR_MODE_REF) {
        h->h_tables.assign(tab, 0.0f);
        std::vector<float> full((size_t)(cfg.full_height > cfg.width ? cfg.full_height : cfg.width) + 1);
        for (int o = 0; o < h->octaves; ++o) {
            const OctGeom& g = h->oct[o];
            const int r0 = cfg.band_row0 >> o;
            for (int s = 0; s < nl; ++s) {
                ref_window(cfg.width, o, s, cfg.sigma0, full.data());
                std::memcpy(&h->h_tables[g.fw_off + (size_t)s * g.pitch], full.data(), sizeof(float) * g.W);
                ref_window(cfg.full_height, o, s, cfg.sigma0, full.data());
                for (int r = 0; r < g.H; ++r) h->h_tables[g.fh_off + (size_t)r * (nl <= 8 ? 8 : 16) + s] = full[r0 + r];
            }
        }
    } else {
        h->conv.resize(nl);
        std::vector<float> taps;
        for (int s = 0; s < nl; ++s) {
            const int R = conv_make_taps(conv_sigma_inc(s, S, cfg.sigma0, cfg.sigma_in), cfg.radius_sigmas, taps);
            if (R > 32) return bail(SSPYR_ERR_UNSUPPORTED, "CONV tap radius > 32 (lower sigma0, S or radius_sigmas)");
            h->conv[s].radius = R;
            h->conv[s].taps_off = h->h_tables.size();
            h->h_tables.insert(h->h_tables.end(), taps.begin(), taps.end());
            h->h_tables.resize(round_up(h->h_tables.size(), 32), 0.0f);
        }
    }

    // ---- device memory ----
    const size_t in_row = (size_t)cfg.width * h->elem_bytes;
    h->in_pitch_bytes = round_up(in_row, 128);
    h->in_frame_bytes = round_up(h->in_pitch_bytes * cfg.height + 128, 256);
    auto dmalloc = [&](void** p, size_t bytes) { return cudaMalloc(p, bytes ? bytes : 256); };
    // CONV: one build counter per (frame slot, octave, level, 128-column strip, 32-row block) -- level chaining
    // (conv_march.cuh), cascade items (conv_cascade.cuh).  They live INSIDE the output allocation, behind the progress
    // flags, so that a neighbour band reads them through the same CUDA-IPC mapping as the planes.
    size_t seg_words = 0;
    if (cfg.mode == SSPYR_MODE_CONV) {
        size_t n = 0;
        for (int o = 0; o < h->octaves; ++o) {
            h->seg_off[o] = n;
            h->seg_cap[o] = (size_t)((h->oct[o].W + 127) / 128) * (size_t)((h->oct[o].H + 31) / 32);
            n += h->seg_cap[o] * nl;
        }
        h->seg_frame_stride = n;
        if (n * cfg.frames < (1ull << 31)) seg_words = n * cfg.frames;
    }
    if ((e = dmalloc((void**)&h->d_out, sizeof(float) * (h->frame_floats * cfg.frames + (size_t)CONV_FLAG_BLOCK * cfg.frames + seg_words))) != cudaSuccess ||
        (e = dmalloc((void**)&h->d_in, h->in_frame_bytes * cfg.frames)) != cudaSuccess ||
        (e = dmalloc((void**)&h->d_tables, sizeof(float) * h->h_tables.size())) != cudaSuccess) {
        cudaGetLastError();
        return bail(e == cudaErrorMemoryAllocation ? SSPYR_ERR_NOMEM : SSPYR_ERR_CUDA,
                    std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
    if (cfg.mode == SSPYR_MODE_CONV && banded) {
        int rmax = 0;
        for (int s2 = 0; s2 < nl; ++s2) rmax = h->conv[s2].radius > rmax ? h->conv[s2].radius : rmax;
        h->halo_rmax = rmax;
        size_t hf = 0;
        for (int o = 0; o < h->octaves; ++o) {
            if (h->oct[o].H < rmax)
                return bail(SSPYR_ERR_ARG, "row band shorter than the blur radius at some octave (use fewer bands or octaves)");
            h->halo_off[o] = hf;
            hf += (size_t)2 * rmax * h->oct[o].pitch;
        }
        if ((e = dmalloc((void**)&h->d_halo, sizeof(float) * hf)) != cudaSuccess ||
            (e = dmalloc((void**)&h->d_halo_raw, (size_t)2 * rmax * h->in_pitch_bytes)) != cudaSuccess ||
            (e = cudaMemset(h->d_halo, 0, sizeof(float) * hf)) != cudaSuccess ||
            (e = cudaMemset(h->d_halo_raw, 0, (size_t)2 * rmax * h->in_pitch_bytes)) != cudaSuccess) {
            cudaGetLastError();
            return bail(SSPYR_ERR_NOMEM, std::string("cudaMalloc (halo): ") + cudaGetErrorString(e));
        }
    }
    if (cfg.outputs & SSPYR_OUT_EXTREMA) {
        if ((e = dmalloc((void**)&h->d_ext, h->ext_frame_bytes * cfg.frames)) != cudaSuccess) {
            cudaGetLastError();
            return bail(SSPYR_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        }
    }
    if (cfg.outputs & SSPYR_OUT_KEYPOINTS) {
        h->kp_capacity = cfg.max_keypoints > 0 ? cfg.max_keypoints : (1 << 20);
        h->kp_frame_bytes = round_up(16 + sizeof(sspyr_keypoint) * (size_t)h->kp_capacity, 256);
        if ((e = dmalloc((void**)&h->d_kp, h->kp_frame_bytes * cfg.frames)) != cudaSuccess ||
            (e = cudaMemset(h->d_kp, 0, h->kp_frame_bytes * cfg.frames)) != cudaSuccess) {
            cudaGetLastError();
            return bail(SSPYR_ERR_NOMEM, std::string("cudaMalloc (keypoints): ") + cudaGetErrorString(e));
        }
        const unsigned cap = (unsigned)h->kp_capacity;       // header word 1 of every slot: the capacity
        for (int f = 0; f < cfg.frames; ++f)
            if ((e = cudaMemcpy(h->d_kp + (size_t)f * h->kp_frame_bytes + 4, &cap, 4, cudaMemcpyHostToDevice)) != cudaSuccess)
                return bail(SSPYR_ERR_CUDA, std::string("device setup: ") + cudaGetErrorString(e));
    }
    h->d_flag = reinterpret_cast<unsigned*>(h->d_out + h->frame_floats * cfg.frames);   // inside d_out: one IPC handle covers it
    if (seg_words) {
        h->d_seg = h->d_flag + (size_t)CONV_FLAG_BLOCK * cfg.frames;
        if ((e = cudaMemset(h->d_seg, 0, sizeof(unsigned) * seg_words)) != cudaSuccess)
            return bail(SSPYR_ERR_CUDA, std::string("device setup: ") + cudaGetErrorString(e));
    }
    h->build_seq.assign(cfg.frames, 0);
    if ((e = cudaMemset(h->d_flag, 0, (size_t)CONV_FLAG_BLOCK * cfg.frames * sizeof(float))) != cudaSuccess)
        return bail(SSPYR_ERR_CUDA, std::string("device setup: ") + cudaGetErrorString(e));
    if ((e = cudaMemcpy(h->d_tables, h->h_tables.data(), sizeof(float) * h->h_tables.size(),
                        cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemset(h->d_in, 0, h->in_frame_bytes * cfg.frames)) != cudaSuccess ||
        (e = cudaEventCreate(&h->ev0)) != cudaSuccess || (e = cudaEventCreate(&h->ev1)) != cudaSuccess)
        return bail(SSPYR_ERR_CUDA, std::string("device setup: ") + cudaGetErrorString(e));
    if (cfg.mode == SSPYR_MODE_CONV && h->octaves > 1) {
        h->aux.assign(h->octaves - 1, nullptr);
        h->ev_base.assign(h->octaves - 1, nullptr);
        h->ev_done.assign(h->octaves - 1, nullptr);
        for (int o = 0; o + 1 < h->octaves; ++o)
            if ((e = cudaStreamCreateWithFlags(&h->aux[o], cudaStreamNonBlocking)) != cudaSuccess ||
                (e = cudaEventCreateWithFlags(&h->ev_base[o], cudaEventDisableTiming)) != cudaSuccess ||
                (e = cudaEventCreateWithFlags(&h->ev_done[o], cudaEventDisableTiming)) != cudaSuccess)
                return bail(SSPYR_ERR_CUDA, std::string("stream setup: ") + cudaGetErrorString(e));
    }
    *out = h;
    return SSPYR_OK;
}

int sspyr_destroy(sspyr_handle h) {
    if (!h) return SSPYR_OK;
    cudaSetDevice(h->device);
    for (int side = 0; side < 2; ++side) {
        if (h->peer[side].ipc_out) cudaIpcCloseMemHandle(h->peer[side].ipc_out);
        if (h->peer[side].ipc_in) cudaIpcCloseMemHandle(h->peer[side].ipc_in);
    }
    if (h->d_out) cudaFree(h->d_out);
    if (h->d_ext) cudaFree(h->d_ext);
    if (h->d_kp) cudaFree(h->d_kp);
    if (h->d_in) cudaFree(h->d_in);
    if (h->d_tables) cudaFree(h->d_tables);
    if (h->d_halo) cudaFree(h->d_halo);
    if (h->d_halo_raw) cudaFree(h->d_halo_raw);
    conv_drop_graphs(h);
    conv_cascade_free(h);
    destroy_lanes(h);
    for (cudaStream_t st : h->aux) if (st) cudaStreamDestroy(st);
    for (cudaEvent_t ev : h->ev_base) if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : h->ev_done) if (ev) cudaEventDestroy(ev);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    cudaGetLastError();
    delete h;
    return SSPYR_OK;
}

const char* sspyr_last_error(sspyr_handle h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int sspyr_num_octaves(sspyr_handle h) { return h ? h->octaves : SSPYR_ERR_ARG; }
int sspyr_num_levels(sspyr_handle h) { return h ? h->nl : SSPYR_ERR_ARG; }
int sspyr_num_dogs(sspyr_handle h) { return h ? h->nl - 1 : SSPYR_ERR_ARG; }

int sspyr_level_dims(sspyr_handle h, int octave, int* rows, int* cols, size_t* pitch_floats) {
    if (!h) return SSPYR_ERR_ARG;
    if (octave < 0 || octave >= h->octaves) return fail(h, SSPYR_ERR_ARG, "octave out of range");
    if (rows) *rows = h->oct[octave].H;
    if (cols) *cols = h->oct[octave].W;
    if (pitch_floats) *pitch_floats = (size_t)h->oct[octave].pitch;
    return SSPYR_OK;
}

int sspyr_algorithmic_bytes(sspyr_handle h, uint64_t* bytes) {
    if (!h || !bytes) return SSPYR_ERR_ARG;
    const int out = h->cfg.outputs, nl = h->nl;
    int planes = 0;
    if (out & SSPYR_OUT_GAUSS) planes += nl;
    else if (out & SSPYR_OUT_GAUSS_TOP) planes += 1;
    if (out & SSPYR_OUT_DOG) planes += nl - 1;
    uint64_t px = 0, ext = 0;
    for (int o = 0; o < h->octaves; ++o) px += (uint64_t)h->oct[o].H * h->oct[o].W;
    if (out & SSPYR_OUT_EXTREMA) ext = px * (uint64_t)h->cfg.S;
    *bytes = (uint64_t)h->cfg.height * h->cfg.width * h->elem_bytes + 4ull * planes * px + ext;
    return SSPYR_OK;
}

int sspyr_set_stream(sspyr_handle h, void* cuda_stream) {
    if (!h) return SSPYR_ERR_ARG;
    conv_drop_graphs(h);
    h->stream = static_cast<cudaStream_t>(cuda_stream);
    return SSPYR_OK;
}

int sspyr_upload(sspyr_handle h, int frame, const void* host, size_t pitch_bytes) {
    NvtxRange nvtx_("sspyr_upload");
    if (!h || !host) return SSPYR_ERR_ARG;
    if (!valid_frame(h, frame)) return fail(h, SSPYR_ERR_ARG, "frame slot out of range");
    const size_t row = (size_t)h->cfg.width * h->elem_bytes;
    if (pitch_bytes == 0) pitch_bytes = row;
    if (pitch_bytes < row) return fail(h, SSPYR_ERR_ARG, "pitch smaller than a row");
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaMemcpy2DAsync(h->d_in + (size_t)frame * h->in_frame_bytes, h->in_pitch_bytes, host, pitch_bytes,
                            row, h->cfg.height, cudaMemcpyHostToDevice, h->stream));
    CU(h, mark_tail(h));
    if (h->ext_in[frame]) conv_drop_graphs(h);
    h->ext_in[frame] = nullptr;
    h->built[frame] = 0;
    return SSPYR_OK;
}

int sspyr_set_input_device(sspyr_handle h, int frame, const void* dev, size_t pitch_bytes) {
    if (!h) return SSPYR_ERR_ARG;
    if (!valid_frame(h, frame)) return fail(h, SSPYR_ERR_ARG, "frame slot out of range");
    if (dev) {
        const size_t row = (size_t)h->cfg.width * h->elem_bytes;
        if (pitch_bytes == 0) pitch_bytes = row;
        if (pitch_bytes < row || pitch_bytes % 16 || reinterpret_cast<uintptr_t>(dev) % 16)
            return fail(h, SSPYR_ERR_ARG, "device image must be 16-byte aligned with a pitch that is a multiple of 16 bytes");
    }
    if (h->ext_in[frame] != dev || h->ext_pitch[frame] != pitch_bytes) conv_drop_graphs(h);   // pointers are baked in
    h->ext_in[frame] = dev;
    if (dev) h->strict_order = true;     // produced by the caller's own work on the stream: builds follow the whole stream
    h->ext_pitch[frame] = pitch_bytes;
    h->built[frame] = 0;
    return SSPYR_OK;
}

int sspyr_build_batch(sspyr_handle h, int first, int count) {
    NvtxRange nvtx_("sspyr_build");
    if (!h) return SSPYR_ERR_ARG;
    if (!valid_frame(h, first) || count < 1) return fail(h, SSPYR_ERR_ARG, "bad frame range");
    CU(h, cudaSetDevice(h->device));
    int launches = 0;
    if (h->tune.timing) CU(h, cudaEventRecord(h->ev0, h->stream));
    cudaError_t e = cudaSuccess;
    if (h->cfg.mode == SSPYR_MODE_REF) {
        e = launch_ref(h, first, count, h->cfg.outputs, &launches);
        if (e == cudaSuccess) e = scan_extrema(h, first, count, &launches);
    } else {
        const bool banded_conv = h->cfg.full_height != h->cfg.height;
        const bool peers_ok = (!conv_has_up(h) || h->peer[0].attached) && (!conv_has_down(h) || h->peer[1].attached);
        if (banded_conv && !peers_ok)
            return fail(h, SSPYR_ERR_STATE, "a row-band CONV handle without attached neighbours is driven level by level: "
                                            "sspyr_conv_step + halo exchange");
        if (banded_conv && count != 1) return fail(h, SSPYR_ERR_ARG, "row-band CONV builds one frame slot at a time");
        int done = 0;
        while (done < count && e == cudaSuccess) {           // contiguous own slots share one launch per level
            const int f0 = (first + done) % h->cfg.frames;
            int n = 1;
            if (!h->ext_in[f0])
                while (done + n < count && f0 + n < h->cfg.frames && !h->ext_in[f0 + n]) ++n;
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            const bool capturing = cudaStreamIsCapturing(h->stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone;
            const int nlanes = lane_count(h);
            if (!capturing && conv_cascade_ok(h)) {          // one launch per build on the handle's stream (build numbers are
                if (n > 64) n = 64;                          // launch parameters, so not while the caller captures a graph)
                e = launch_conv_cascade(h, f0, n, &launches);
            } else if (nlanes > 1 && !capturing)
                e = build_on_lane(h, f0, n, nlanes, &launches);
            else
                e = launch_conv_graphed(h, f0, n, &launches, own_streams(h));
            done += n;
        }
        if (e == cudaSuccess) e = scan_extrema(h, first, count, &launches);
        if (e == cudaSuccess && (h->cfg.outputs & (SSPYR_OUT_EXTREMA | SSPYR_OUT_KEYPOINTS))) e = mark_tail(h);
    }
    if (e != cudaSuccess) {
        h->seg_dirty = true;                                 // some levels may have counted this build, others not
        return fail_cuda(h, e, "kernel launch");
    }
    if (h->tune.timing) CU(h, cudaEventRecord(h->ev1, h->stream));
    h->timed = h->tune.timing != 0;
    h->last_launches = launches;
    for (int i = 0; i < count; ++i) h->built[(first + i) % h->cfg.frames] = 1;
    return SSPYR_OK;
}

int sspyr_build(sspyr_handle h, int frame) { return sspyr_build_batch(h, frame, 1); }

int sspyr_build_stage(sspyr_handle h, int frame, int stage) {
    NvtxRange nvtx_("sspyr_build_stage");
    if (!h) return SSPYR_ERR_ARG;
    if (stage == SSPYR_STAGE_DOG) return sspyr_build_batch(h, frame, 1);
    if (stage != SSPYR_STAGE_INIT && stage != SSPYR_STAGE_FILTER) return fail(h, SSPYR_ERR_ARG, "bad stage");
    if (!valid_frame(h, frame)) return fail(h, SSPYR_ERR_ARG, "frame slot out of range");
    if (h->cfg.mode != SSPYR_MODE_REF) return fail(h, SSPYR_ERR_UNSUPPORTED, "partial stages exist in REF mode only");
    if (!(h->cfg.outputs & SSPYR_OUT_GAUSS)) return fail(h, SSPYR_ERR_STATE, "partial stages need SSPYR_OUT_GAUSS");
    CU(h, cudaSetDevice(h->device));
    int launches = 0;
    if (h->tune.timing) CU(h, cudaEventRecord(h->ev0, h->stream));
    const int outputs = SSPYR_OUT_GAUSS | (stage == SSPYR_STAGE_INIT ? SSPYR_INT_INIT_ONLY : 0);
    const cudaError_t e = launch_ref(h, frame, 1, outputs, &launches);
    if (e != cudaSuccess) return fail_cuda(h, e, "kernel launch");
    if (h->tune.timing) CU(h, cudaEventRecord(h->ev1, h->stream));
    h->timed = h->tune.timing != 0;
    h->last_launches = launches;
    h->built[frame] = 1;
    return SSPYR_OK;
}

int sspyr_sync(sspyr_handle h) {
    NvtxRange nvtx_("sspyr_sync");
    if (!h) return SSPYR_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    if (h->cfg.mode == SSPYR_MODE_CONV) {                    // a bounded wait (neighbour band, previous level, TMA) gave up
        unsigned mark = 0;
        CU(h, cudaMemcpy(&mark, h->d_flag + 16, sizeof(mark), cudaMemcpyDeviceToHost));
        if (mark) {
            h->seg_dirty = true;
            CU(h, cudaMemset(h->d_flag + 16, 0, sizeof(mark)));
            return fail(h, SSPYR_ERR_STATE, "timed out waiting for a neighbour band or a previous level (counter value " + std::to_string(mark & 0x7fffffffu) + ")");
        }
    }
    return SSPYR_OK;
}

int sspyr_elapsed_ms(sspyr_handle h, float* ms) {
    if (!h || !ms) return SSPYR_ERR_ARG;
    if (!h->timed) return fail(h, SSPYR_ERR_STATE, "no timed build yet (enable with sspyr_set_tuning(h, \"timing\", 1))");
    CU(h, cudaEventSynchronize(h->ev1));
    CU(h, cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return SSPYR_OK;
}

int sspyr_last_launches(sspyr_handle h) { return h ? h->last_launches : SSPYR_ERR_ARG; }

int sspyr_device_ptr(sspyr_handle h, int frame, int octave, int level, int kind, void** ptr) {
    if (!h || !ptr) return SSPYR_ERR_ARG;
    h->strict_order = true;       // the caller's own kernels may now touch the slots: builds follow the whole stream
    if (!valid_frame(h, frame)) return fail(h, SSPYR_ERR_ARG, "frame slot out of range");
    if (kind == SSPYR_KIND_KEYPOINTS) {
        if (!h->d_kp) return fail(h, SSPYR_ERR_STATE, "keypoint output not configured");
        *ptr = h->d_kp + (size_t)frame * h->kp_frame_bytes;
        return SSPYR_OK;
    }
    if (kind == SSPYR_KIND_EXTREMA) {
        if (!h->d_ext) return fail(h, SSPYR_ERR_STATE, "extrema output not configured");
        if (octave < 0 || octave >= h->octaves || level < 0 || level >= h->cfg.S)
            return fail(h, SSPYR_ERR_ARG, "extrema plane out of range");
        *ptr = h->d_ext + (size_t)frame * h->ext_frame_bytes + h->oct[octave].ext_off + (size_t)level * h->oct[octave].plane;
        return SSPYR_OK;
    }
    int idx = 0;
    const int rc = plane_lookup(h, octave, level, kind, &idx);
    if (rc) return rc;
    *ptr = frame_out(h, frame) + h->oct[octave].off + (size_t)idx * h->oct[octave].plane;
    return SSPYR_OK;
}

int sspyr_download(sspyr_handle h, int frame, int octave, int level, int kind, void* dst, size_t pitch_bytes) {
    NvtxRange nvtx_("sspyr_download");
    if (!h || !dst) return SSPYR_ERR_ARG;
    if (!valid_frame(h, frame)) return fail(h, SSPYR_ERR_ARG, "frame slot out of range");
    if (!h->built[frame]) return fail(h, SSPYR_ERR_STATE, "frame slot has not been built since its last upload");
    void* src = nullptr;
    const bool strict = h->strict_order;                     // (the pointer does not leave the library here)
    const int rc = sspyr_device_ptr(h, frame, octave, level, kind, &src);
    h->strict_order = strict;
    if (rc) return rc;
    const OctGeom& g = h->oct[octave];
    const size_t es = kind == SSPYR_KIND_EXTREMA ? 1 : sizeof(float);
    const size_t row = (size_t)g.W * es;
    if (pitch_bytes == 0) pitch_bytes = row;
    if (pitch_bytes < row) return fail(h, SSPYR_ERR_ARG, "pitch smaller than a row");
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaMemcpy2DAsync(dst, pitch_bytes, src, (size_t)g.pitch * es, row, g.H, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return SSPYR_OK;
}

int sspyr_download_inplace(sspyr_handle h, int frame, float* dst) {
    NvtxRange nvtx_("sspyr_download_inplace");
    if (!h || !dst) return SSPYR_ERR_ARG;
    if (!valid_frame(h, frame)) return fail(h, SSPYR_ERR_ARG, "frame slot out of range");
    if (!h->built[frame]) return fail(h, SSPYR_ERR_STATE, "frame slot has not been built since its last upload");
    const int out = h->cfg.outputs;
    if (!(out & SSPYR_OUT_DOG) || !(out & (SSPYR_OUT_GAUSS | SSPYR_OUT_GAUSS_TOP)))
        return fail(h, SSPYR_ERR_STATE, "in-place layout needs DOG and GAUSS(_TOP) outputs");
    CU(h, cudaSetDevice(h->device));
    const int nl = h->nl;
    for (int o = 0; o < h->octaves; ++o) {
        const OctGeom& g = h->oct[o];
        const float* src = frame_out(h, frame) + g.off + (size_t)(nl - 1) * g.plane;   // DoG_0 .. G_{S+2}
        CU(h, copy_planes_to_host(dst, src, g.W, g.pitch, (size_t)nl * g.H, h->stream));
        dst += (size_t)nl * g.H * g.W;
    }
    CU(h, mark_tail(h));
    return SSPYR_OK;
}

int sspyr_download_gauss(sspyr_handle h, int frame, float* dst) {
    NvtxRange nvtx_("sspyr_download_gauss");
    if (!h || !dst) return SSPYR_ERR_ARG;
    if (!valid_frame(h, frame)) return fail(h, SSPYR_ERR_ARG, "frame slot out of range");
    if (!h->built[frame]) return fail(h, SSPYR_ERR_STATE, "frame slot has not been built since its last upload");
    if (!(h->cfg.outputs & SSPYR_OUT_GAUSS)) return fail(h, SSPYR_ERR_STATE, "GAUSS output not configured");
    CU(h, cudaSetDevice(h->device));
    const int nl = h->nl;
    for (int o = 0; o < h->octaves; ++o) {
        const OctGeom& g = h->oct[o];
        const float* base = frame_out(h, frame) + g.off;
        CU(h, copy_planes_to_host(dst, base, g.W, g.pitch, (size_t)(nl - 1) * g.H, h->stream));
        dst += (size_t)(nl - 1) * g.H * g.W;
        CU(h, copy_planes_to_host(dst, base + (size_t)(2 * nl - 2) * g.plane, g.W, g.pitch, g.H, h->stream));
        dst += (size_t)g.H * g.W;
    }
    CU(h, mark_tail(h));
    return SSPYR_OK;
}

int sspyr_download_keypoints(sspyr_handle h, int frame, sspyr_keypoint* dst, int capacity, int* count) {
    NvtxRange nvtx_("sspyr_download_keypoints");
    if (!h || !count || capacity < 0 || (capacity > 0 && !dst)) return SSPYR_ERR_ARG;
    if (!valid_frame(h, frame)) return fail(h, SSPYR_ERR_ARG, "frame slot out of range");
    if (!h->d_kp) return fail(h, SSPYR_ERR_STATE, "keypoint output not configured (SSPYR_OUT_KEYPOINTS)");
    if (!h->built[frame]) return fail(h, SSPYR_ERR_STATE, "frame slot has not been built since its last upload");
    CU(h, cudaSetDevice(h->device));
    const unsigned char* slot = h->d_kp + (size_t)frame * h->kp_frame_bytes;
    CU(h, cudaMemcpyAsync(count, slot, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    const int n = capacity < h->kp_capacity ? capacity : h->kp_capacity;
    if (n > 0) CU(h, cudaMemcpyAsync(dst, slot + 16, sizeof(sspyr_keypoint) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CU(h, mark_tail(h));
    return SSPYR_OK;
}

int sspyr_host_alloc(size_t bytes, void** ptr) {
    if (!ptr) return SSPYR_ERR_ARG;
    *ptr = nullptr;
    const cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, e == cudaErrorMemoryAllocation ? SSPYR_ERR_NOMEM : SSPYR_ERR_CUDA,
                    std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
    }
    return SSPYR_OK;
}

int sspyr_host_free(void* ptr) {
    if (!ptr) return SSPYR_OK;
    const cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, SSPYR_ERR_CUDA, std::string("cudaFreeHost: ") + cudaGetErrorString(e));
    }
    return SSPYR_OK;
}

int sspyr_window_table(sspyr_handle h, int octave, int level, int axis, float* dst, int capacity) {
    if (!h || !dst) return SSPYR_ERR_ARG;
    if (h->cfg.mode != SSPYR_MODE_REF) return fail(h, SSPYR_ERR_STATE, "window tables exist in REF mode only");
    if (octave < 0 || octave >= h->octaves || level < 0 || level >= h->nl || (axis != 0 && axis != 1))
        return fail(h, SSPYR_ERR_ARG, "bad table index");
    const OctGeom& g = h->oct[octave];
    const int n = axis == 0 ? g.H : g.W;
    if (capacity < n) return fail(h, SSPYR_ERR_ARG, "destination too small");
    // read back from the DEVICE copy: this is what the kernel actually multiplies by
    CU(h, cudaSetDevice(h->device));
    if (axis == 1) {
        CU(h, cudaMemcpy(dst, h->d_tables + g.fw_off + (size_t)level * g.pitch, sizeof(float) * n, cudaMemcpyDeviceToHost));
    } else {   // row window is stored transposed, [row][8]
        CU(h, cudaMemcpy2D(dst, sizeof(float), h->d_tables + g.fh_off + level, (h->nl <= 8 ? 8 : 16) * sizeof(float), sizeof(float), n,
                           cudaMemcpyDeviceToHost));
    }
    return n;
}

int sspyr_conv_taps(sspyr_handle h, int level, float* dst, int capacity, int* radius) {
    if (!h || !dst) return SSPYR_ERR_ARG;
    if (h->cfg.mode != SSPYR_MODE_CONV) return fail(h, SSPYR_ERR_STATE, "taps exist in CONV mode only");
    if (level < 0 || level >= h->nl) return fail(h, SSPYR_ERR_ARG, "level out of range");
    const int R = h->conv[level].radius;
    if (capacity < 2 * R + 1) return fail(h, SSPYR_ERR_ARG, "destination too small");
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaMemcpy(dst, h->d_tables + h->conv[level].taps_off, sizeof(float) * (2 * R + 1), cudaMemcpyDeviceToHost));
    if (radius) *radius = R;
    return SSPYR_OK;
}

int sspyr_halo_rows(sspyr_handle h, int octave, int level, int* rows) {
    if (!h || !rows) return SSPYR_ERR_ARG;
    if (h->cfg.mode != SSPYR_MODE_CONV) { *rows = 0; return SSPYR_OK; }   // REF mode is pointwise: radius 0
    if (octave < 0 || octave >= h->octaves || level < 0 || level >= h->nl) return fail(h, SSPYR_ERR_ARG, "bad (octave, level)");
    *rows = (level == 0 && octave > 0) ? 0 : h->conv[level].radius;
    return SSPYR_OK;
}

int sspyr_halo_ptrs(sspyr_handle h, int frame, int octave, int level, void** send_up, void** send_down,
                    void** recv_up, void** recv_down, size_t* bytes) {
    if (!h) return SSPYR_ERR_ARG;
    if (h->cfg.mode != SSPYR_MODE_CONV || !h->d_halo) return fail(h, SSPYR_ERR_STATE, "not a row-band CONV handle");
    if (!valid_frame(h, frame)) return fail(h, SSPYR_ERR_ARG, "frame slot out of range");
    if (octave < 0 || octave >= h->octaves || level < 0 || level >= h->nl) return fail(h, SSPYR_ERR_ARG, "bad (octave, level)");
    if (level == 0 && octave > 0) return fail(h, SSPYR_ERR_ARG, "level 0 of octave > 0 is a decimation: no halo");
    const int R = h->conv[level].radius;
    const OctGeom& g = h->oct[octave];
    if (level == 0) {                                        // the raw frame feeds (0, 0)
        if (h->ext_in[frame]) return fail(h, SSPYR_ERR_UNSUPPORTED, "row-band CONV needs the frame in the handle's own slot (sspyr_upload)");
        unsigned char* in = h->d_in + (size_t)frame * h->in_frame_bytes;
        if (send_up) *send_up = in;
        if (send_down) *send_down = in + (size_t)(h->cfg.height - R) * h->in_pitch_bytes;
        if (recv_up) *recv_up = conv_halo_raw(h, 0);
        if (recv_down) *recv_down = conv_halo_raw(h, 1);
        if (bytes) *bytes = (size_t)R * h->in_pitch_bytes;
    } else {
        float* src = frame_out(h, frame) + g.off + (size_t)plane_index(h->nl, SSPYR_KIND_GAUSS, level - 1) * g.plane;
        if (send_up) *send_up = src;
        if (send_down) *send_down = src + (size_t)(g.H - R) * g.pitch;
        if (recv_up) *recv_up = conv_halo_plane(h, octave, 0);
        if (recv_down) *recv_down = conv_halo_plane(h, octave, 1);
        if (bytes) *bytes = sizeof(float) * (size_t)R * g.pitch;
    }
    return SSPYR_OK;
}

int sspyr_conv_step(sspyr_handle h, int frame, int octave, int level) {
    NvtxRange nvtx_("sspyr_conv_step");
    if (!h) return SSPYR_ERR_ARG;
    if (h->cfg.mode != SSPYR_MODE_CONV) return fail(h, SSPYR_ERR_STATE, "sspyr_conv_step is for CONV mode");
    if (!valid_frame(h, frame)) return fail(h, SSPYR_ERR_ARG, "frame slot out of range");
    if (octave < 0 || octave >= h->octaves || level < 0 || level >= h->nl) return fail(h, SSPYR_ERR_ARG, "bad (octave, level)");
    CU(h, cudaSetDevice(h->device));
    int launches = 0;
    if (octave == 0 && level == 0) {
        const cudaError_t e0 = conv_begin_build(h, frame, h->stream, &launches);
        if (e0 != cudaSuccess) return fail_cuda(h, e0, "kernel launch");
    }
    const cudaError_t e = launch_conv_step(h, frame, 1, octave, level, h->stream, &launches);
    if (e != cudaSuccess) return fail_cuda(h, e, "kernel launch");
    h->last_launches = launches;
    if (octave == h->octaves - 1 && level == h->nl - 1) {
        if (h->cfg.outputs & (SSPYR_OUT_EXTREMA | SSPYR_OUT_KEYPOINTS)) {
            const cudaError_t e2 = launch_extrema(h, frame, 1, &launches);
            if (e2 != cudaSuccess) return fail_cuda(h, e2, "kernel launch");
        }
        h->built[frame] = 1;
        CU(h, mark_tail(h));                                 // a later whole-pyramid build on a frame lane follows this
    }
    return SSPYR_OK;
}

namespace {

struct IpcBlob {                              // what a neighbour needs to read this band's planes in place
    uint32_t magic, bytes;
    int32_t height, width, octaves, nl, frames, pixel_type, mode, pad;
    uint64_t frame_floats, in_frame_bytes, in_pitch_bytes, flag_off_floats;
    uint64_t seg_off_words, seg_frame_stride;          // segment counters: offset from the flags, counters per slot (0: none)
    struct { int32_t H, pitch; uint64_t off, plane, seg_off, seg_cap; } oct[SSPYR_MAX_OCTAVES];
    cudaIpcMemHandle_t out, in;
};
static_assert(sizeof(IpcBlob) <= SSPYR_IPC_BLOB_BYTES, "blob too large");
const uint32_t kBlobMagic = 0x53505952u;      // 'SPYR'

int peer_check(sspyr_ctx* h, int side, int width, int octaves, int nl, int frames, int pixel_type, int mode) {
    if (side != SSPYR_SIDE_ABOVE && side != SSPYR_SIDE_BELOW) return fail(h, SSPYR_ERR_ARG, "bad side");
    if (h->cfg.mode != SSPYR_MODE_CONV || h->cfg.full_height == h->cfg.height)
        return fail(h, SSPYR_ERR_STATE, "peer halos are for row-band CONV handles");
    if ((side == SSPYR_SIDE_ABOVE && !conv_has_up(h)) || (side == SSPYR_SIDE_BELOW && !conv_has_down(h)))
        return fail(h, SSPYR_ERR_ARG, "this band has no neighbour on that side");
    if (width != h->cfg.width || octaves != h->octaves || nl != h->nl || frames != h->cfg.frames ||
        pixel_type != h->cfg.pixel_type || mode != h->cfg.mode)
        return fail(h, SSPYR_ERR_ARG, "neighbour band has a different configuration");
    return SSPYR_OK;
}

}  // namespace

int sspyr_ipc_export(sspyr_handle h, void* blob, size_t capacity, size_t* bytes) {
    if (!h || !blob) return SSPYR_ERR_ARG;
    if (capacity < sizeof(IpcBlob)) return fail(h, SSPYR_ERR_ARG, "blob buffer smaller than SSPYR_IPC_BLOB_BYTES");
    CU(h, cudaSetDevice(h->device));
    IpcBlob b;
    std::memset(&b, 0, sizeof(b));
    b.magic = kBlobMagic;
    b.bytes = sizeof(IpcBlob);
    b.height = h->cfg.height; b.width = h->cfg.width; b.octaves = h->octaves; b.nl = h->nl;
    b.frames = h->cfg.frames; b.pixel_type = h->cfg.pixel_type; b.mode = h->cfg.mode;
    b.frame_floats = h->frame_floats; b.in_frame_bytes = h->in_frame_bytes; b.in_pitch_bytes = h->in_pitch_bytes;
    b.flag_off_floats = h->frame_floats * h->cfg.frames;
    b.seg_off_words = (uint64_t)CONV_FLAG_BLOCK * h->cfg.frames;
    b.seg_frame_stride = h->d_seg ? h->seg_frame_stride : 0;
    for (int o = 0; o < h->octaves; ++o) {
        b.oct[o].H = h->oct[o].H; b.oct[o].pitch = h->oct[o].pitch; b.oct[o].off = h->oct[o].off; b.oct[o].plane = h->oct[o].plane;
        b.oct[o].seg_off = h->seg_off[o]; b.oct[o].seg_cap = h->seg_cap[o];
    }
    CU(h, cudaIpcGetMemHandle(&b.out, h->d_out));
    CU(h, cudaIpcGetMemHandle(&b.in, h->d_in));
    std::memcpy(blob, &b, sizeof(b));
    if (bytes) *bytes = sizeof(b);
    return SSPYR_OK;
}

int sspyr_ipc_attach(sspyr_handle h, int side, const void* blob, size_t bytes) {
    if (!h || !blob) return SSPYR_ERR_ARG;
    if (bytes < sizeof(IpcBlob)) return fail(h, SSPYR_ERR_ARG, "short blob");
    IpcBlob b;
    std::memcpy(&b, blob, sizeof(b));
    if (b.magic != kBlobMagic || b.bytes != sizeof(IpcBlob)) return fail(h, SSPYR_ERR_ARG, "not an sspyr IPC blob");
    const int rc = peer_check(h, side, b.width, b.octaves, b.nl, b.frames, b.pixel_type, b.mode);
    if (rc) return rc;
    CU(h, cudaSetDevice(h->device));
    sspyr_ctx::Peer& q = h->peer[side];
    if (q.attached) return fail(h, SSPYR_ERR_STATE, "side already attached");
    CU(h, cudaIpcOpenMemHandle(&q.ipc_out, b.out, cudaIpcMemLazyEnablePeerAccess));
    CU(h, cudaIpcOpenMemHandle(&q.ipc_in, b.in, cudaIpcMemLazyEnablePeerAccess));
    q.out = static_cast<const float*>(q.ipc_out);
    q.in = static_cast<const unsigned char*>(q.ipc_in);
    q.flag = reinterpret_cast<const unsigned*>(q.out + b.flag_off_floats);
    q.seg = b.seg_frame_stride ? q.flag + b.seg_off_words : nullptr;
    q.seg_frame_stride = b.seg_frame_stride;
    q.height = b.height;
    q.frame_floats = b.frame_floats;
    q.in_frame_bytes = b.in_frame_bytes;
    for (int o = 0; o < h->octaves; ++o) {
        q.H[o] = b.oct[o].H; q.off[o] = b.oct[o].off; q.plane[o] = b.oct[o].plane;
        q.seg_off[o] = b.oct[o].seg_off; q.seg_cap[o] = b.oct[o].seg_cap;
    }
    q.attached = true;
    return SSPYR_OK;
}

int sspyr_peer_attach_local(sspyr_handle h, int side, sspyr_handle n) {
    if (!h || !n || h == n) return SSPYR_ERR_ARG;
    const int rc = peer_check(h, side, n->cfg.width, n->octaves, n->nl, n->cfg.frames, n->cfg.pixel_type, n->cfg.mode);
    if (rc) return rc;
    sspyr_ctx::Peer& q = h->peer[side];
    if (q.attached) return fail(h, SSPYR_ERR_STATE, "side already attached");
    if (n->device != h->device) {               // bands of one process on two GPUs: plain peer access
        CU(h, cudaSetDevice(h->device));
        const cudaError_t e = cudaDeviceEnablePeerAccess(n->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail_cuda(h, e, "cudaDeviceEnablePeerAccess");
        cudaGetLastError();
    }
    q.out = n->d_out;
    q.in = n->d_in;
    q.flag = n->d_flag;
    q.seg = n->d_seg;
    q.seg_frame_stride = n->seg_frame_stride;
    q.height = n->cfg.height;
    q.frame_floats = n->frame_floats;
    q.in_frame_bytes = n->in_frame_bytes;
    for (int o = 0; o < h->octaves; ++o) {
        q.H[o] = n->oct[o].H; q.off[o] = n->oct[o].off; q.plane[o] = n->oct[o].plane;
        q.seg_off[o] = n->seg_off[o]; q.seg_cap[o] = n->seg_cap[o];
    }
    q.local = true;
    q.same_device = n->device == h->device;
    q.attached = true;
    return SSPYR_OK;
}

int sspyr_set_tuning(sspyr_handle h, const char* key, int value) {
    if (!h || !key) return SSPYR_ERR_ARG;
    // Retuning is rare and changes how the next build is issued (other streams, other segment grid, no captured
    // launch sequences): let everything in flight finish first -- every lane has been joined into the stream --
    // and make the next lane build follow this point.
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    conv_drop_graphs(h);
    mark_tail(h);
    cudaGetLastError();
    h->seg_dirty = true;                                     // the segment geometry of the next build may differ
    if (!std::strcmp(key, "rows_per_thread")) h->tune.rows_per_thread = value;
    else if (!std::strcmp(key, "block")) h->tune.block = value;
    else if (!std::strcmp(key, "bx")) h->tune.bx = value;
    else if (!std::strcmp(key, "pdl")) h->tune.pdl = value;
    else if (!std::strcmp(key, "occ")) h->tune.occ = value;
    else if (!std::strcmp(key, "prefetch_next")) h->tune.prefetch_next = value;
    else if (!std::strcmp(key, "timing")) h->tune.timing = value;
    else if (!std::strcmp(key, "conv_streams")) h->tune.conv_streams = value;
    else if (!std::strcmp(key, "conv_march")) h->tune.conv_march = value;
    else if (!std::strcmp(key, "conv_graph")) h->tune.conv_graph = value;
    else if (!std::strcmp(key, "conv_tma")) h->tune.conv_tma = value;
    else if (!std::strcmp(key, "conv_fused_sync")) h->tune.conv_fused_sync = value;
    else if (!std::strcmp(key, "conv_waves")) h->tune.conv_waves = value;
    else if (!std::strcmp(key, "conv_seg_min")) h->tune.conv_seg_min = value;
    else if (!std::strcmp(key, "conv_chain")) h->tune.conv_chain = value;
    else if (!std::strcmp(key, "conv_band_chain")) h->tune.conv_band_chain = value;
    else if (!std::strcmp(key, "conv_band_lanes")) h->tune.conv_band_lanes = value;
    else if (!std::strcmp(key, "conv_cascade")) h->tune.conv_cascade = value;
    else if (!std::strcmp(key, "conv_casc_seg")) h->tune.conv_casc_seg = value;
    else if (!std::strcmp(key, "conv_l2hint")) h->tune.conv_l2hint = value;
    else if (!std::strcmp(key, "conv_lanes")) h->tune.conv_lanes = value;
    else return fail(h, SSPYR_ERR_ARG, std::string("unknown tuning key ") + key);
    return SSPYR_OK;
}

}  // extern "C"
