// csrc/conv_cascade.cu -- host side of the cascade build (conv_cascade.cuh): the parameter block of one launch, the
// per-octave tensor maps, and the launch itself.  Its own translation unit: the kernel holds all twelve radii.
#include <algorithm>
#include <cstring>
#include <vector>

#include <cudaTypedefs.h>

#include "conv_cascade.cuh"

namespace sspyr {

namespace {

PFN_cuTensorMapEncodeTiled tensor_map_encoder() {
    static PFN_cuTensorMapEncodeTiled encode = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
        cudaGetLastError();
    }
    return encode;
}

// All planes of one octave of every frame slot of the handle as one 4-D tensor (columns = row pitch, rows, planes,
// slots); box = 156 columns x 32 rows of one plane of one slot.
bool make_octave_map(CUtensorMap* map, float* base, const OctGeom& g, int planes, int frames, size_t frame_floats) {
    PFN_cuTensorMapEncodeTiled encode = tensor_map_encoder();
    if (!encode || g.H < STRIP_TH || g.pitch < CASC_PIN) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)g.pitch, (cuuint64_t)g.H, (cuuint64_t)planes, (cuuint64_t)(frames > 0 ? frames : 1)};
    const cuuint64_t strides[3] = {(cuuint64_t)g.pitch * sizeof(float), (cuuint64_t)g.plane * sizeof(float),
                                   (cuuint64_t)frame_floats * sizeof(float)};
    const cuuint32_t box[4] = {(cuuint32_t)CASC_PIN, (cuuint32_t)STRIP_TH, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (strides[2] >= (1ull << 40) || strides[1] >= (1ull << 40)) return false;
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// Can this handle build through the cascade kernel?  (whole frames, every radius compiled in, counters allocated)
bool conv_cascade_ok(const sspyr_ctx* h) {
    if (h->cfg.mode != SSPYR_MODE_CONV || h->cfg.full_height != h->cfg.height || !h->d_seg || h->tune.conv_cascade == 0)
        return false;
    // Automatic (conv_cascade = 1): frames of at least 4 Mpixel.  Below that every plane of a frame stays in L2 under the
    // per-level schedule anyway, and the cascade's longer dependency chain per frame (levels trail one another by three
    // steps, octaves follow one another) costs more than its single launch saves.  2 = always.
    if (h->tune.conv_cascade == 1 && (long long)h->cfg.height * h->cfg.width < (4ll << 20)) return false;
    for (int s = 0; s < h->nl; ++s)
        if (h->conv[s].radius > CASC_MAX_R) return false;
    return true;
}

cudaError_t launch_conv_cascade(sspyr_ctx* h, int first, int count, int* launches) {
    if (count < 1 || count > CASC_MAX_FRAMES || first + count > h->cfg.frames) return cudaErrorInvalidValue;
    static_assert(sizeof(CascParams) + sizeof(CascMaps) < 32000, "kernel parameter block");
    if (!h->casc_maps) {                                   // once per handle: the maps cover every frame slot
        h->casc_maps = new (std::nothrow) CascMaps();
        if (!h->casc_maps) return cudaErrorMemoryAllocation;
        std::memset(h->casc_maps, 0, sizeof(CascMaps));
        for (int o = 0; o < h->octaves && o < CASC_MAX_TMA_OCT; ++o)
            h->casc_tma[o] = h->tune.conv_tma != 0 &&
                             make_octave_map(&h->casc_maps->m[o], h->d_out + h->oct[o].off, h->oct[o], 2 * h->nl - 1,
                                             h->cfg.frames, h->frame_floats);
    }
    if (h->seg_dirty) {                                    // counters possibly out of step (retuned, failed build): restart
        cudaError_t me = cudaStreamSynchronize(h->stream);
        if (me == cudaSuccess) me = cudaMemset(h->d_seg, 0, sizeof(unsigned) * h->seg_frame_stride * h->cfg.frames);
        if (me == cudaSuccess) me = cudaMemset(h->d_flag, 0, sizeof(unsigned) * (size_t)CONV_FLAG_BLOCK * h->cfg.frames);
        if (me != cudaSuccess) return me;
        std::fill(h->build_seq.begin(), h->build_seq.end(), 0u);
        if (h->d_casc_tab) cudaFree(h->d_casc_tab);
        h->d_casc_tab = nullptr;
        h->seg_dirty = false;
    }
    CascParams C;
    std::memset(&C, 0, sizeof(C));
    size_t pitch_bytes = 0;
    C.raw = frame_input(h, first, &pitch_bytes);
    C.raw_pitch = (int)(pitch_bytes / h->elem_bytes);
    C.raw_frame_stride = h->in_frame_bytes / h->elem_bytes;
    C.raw_kind = h->cfg.pixel_type;
    C.out_frame_stride = h->frame_floats;
    C.slot_flags = h->d_flag + (size_t)CONV_FLAG_BLOCK * first;
    C.timeout_mark = h->d_flag + CONV_FLAG_TIMEOUT;
    C.ctr_frame_stride = (unsigned)h->seg_frame_stride;
    C.slot0 = first;
    C.octaves = h->octaves;
    C.nl = h->nl;
    C.S = h->cfg.S;
    C.want_dog = (h->cfg.outputs & SSPYR_OUT_DOG) ? 1 : 0;
    CascItemGeom geom[SSPYR_MAX_OCTAVES];
    int radius[SSPYR_MAX_LEVELS];
    for (int s = 0; s < h->nl; ++s) radius[s] = cascade_radius_class(h->conv[s].radius);
    for (int o = 0; o < h->octaves; ++o) {
        const OctGeom& g = h->oct[o];
        CascOct& O = C.oct[o];
        O.base = frame_out(h, first) + g.off;
        O.ctr = h->d_seg + (size_t)first * h->seg_frame_stride + h->seg_off[o];
        O.plane = g.plane;
        O.H = g.H; O.W = g.W; O.pitch = g.pitch;
        O.seg_rows = cascade_seg_rows(g.W, h->tune.conv_casc_seg);
        O.nsegs = (g.H + O.seg_rows - 1) / O.seg_rows;
        O.nstrips = (g.W + CONV_TW - 1) / CONV_TW;
        O.first_level = o == 0 ? 0 : 1;
        O.seg_cap = (unsigned)h->seg_cap[o];
        O.tma = o < CASC_MAX_TMA_OCT && h->casc_tma[o];
        geom[o] = CascItemGeom{O.seg_rows, O.nsegs, O.nstrips, O.first_level, g.H, g.W};
        if (O.nstrips > 1023 || O.nsegs > 16383) return cudaErrorInvalidValue;
    }
    if (!h->d_casc_tab) {                                  // (dropped whenever the segmentation may change: seg_dirty)
        bool keyed = false;
        const std::vector<unsigned> tab = cascade_item_table(geom, h->octaves, h->nl, h->cfg.S, radius, &keyed);
        cudaError_t me = cudaMalloc((void**)&h->d_casc_tab, sizeof(unsigned) * tab.size());
        if (me == cudaSuccess) me = cudaMemcpy(h->d_casc_tab, tab.data(), sizeof(unsigned) * tab.size(), cudaMemcpyHostToDevice);
        if (me != cudaSuccess) return me;
        h->casc_items = (unsigned)tab.size();
        h->casc_keyed = keyed;
    }
    const unsigned long long items = h->casc_items;
    if (items * (unsigned long long)count >= 0x7fffffffULL) return cudaErrorInvalidValue;
    C.item_tab = h->d_casc_tab;
    C.items_per_frame = (unsigned)items;
    for (int s = 0; s < h->nl; ++s) {
        const int R = h->conv[s].radius, RC = radius[s];
        C.lev[s].radius = RC;                              // taps zero-padded to the class radius (bit-identical chains)
        std::memcpy(C.lev[s].taps + (RC - R), h->h_tables.data() + h->conv[s].taps_off, sizeof(float) * (2 * R + 1));
    }
    for (int f = 0; f < count; ++f) C.bseq[f] = (unsigned short)(h->build_seq[first + f]++ & 0xffffu);

    static bool configured[64] = {false};
    if (h->device < 0 || h->device >= 64 || !configured[h->device]) {
        cudaError_t e = cudaFuncSetAttribute(conv_cascade_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)casc_smem_bytes());
        if (e != cudaSuccess) return e;
        if (h->device >= 0 && h->device < 64) configured[h->device] = true;
    }
    // Programmatic dependent launch: the grid may start while its predecessor in the stream drains -- the counters
    // order everything it touches.  Not when the input is the caller's own device image (its producer's stores are
    // only guaranteed visible at the kernel boundary) or raw device pointers are in the caller's hands.
    const bool pdl = h->tune.pdl != 0 && !h->ext_in[first] && !h->strict_order;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(items * (unsigned long long)count));
    cfg.blockDim = dim3(CONV_THREADS);
    cfg.dynamicSmemBytes = casc_smem_bytes();
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, conv_cascade_kernel, C, *h->casc_maps);
    if (e == cudaSuccess) ++*launches;
    return e;
}

void conv_cascade_free(sspyr_ctx* h) {
    delete h->casc_maps;
    h->casc_maps = nullptr;
    if (h->d_casc_tab) cudaFree(h->d_casc_tab);
    h->d_casc_tab = nullptr;
}

}  // namespace sspyr
