// csrc/ref_launch.cu -- host side of the fused REF build: fills the kernel parameter block from the
// handle and picks the instantiation (ref_inst.cu) for the handle's level count.
#include "sspyr_internal.h"

namespace sspyr {

#define SSPYR_DECL(n) cudaError_t launch_ref_nl##n(const RefParams&, int, int, bool, dim3, dim3, cudaStream_t, bool);
SSPYR_DECL(3) SSPYR_DECL(4) SSPYR_DECL(5) SSPYR_DECL(6) SSPYR_DECL(7) SSPYR_DECL(8) SSPYR_DECL(9) SSPYR_DECL(10)
SSPYR_DECL(11) SSPYR_DECL(12) SSPYR_DECL(13) SSPYR_DECL(14) SSPYR_DECL(15) SSPYR_DECL(16)
#undef SSPYR_DECL
cudaError_t launch_ref_prefetch(const void* img, size_t pitch_bytes, int row_bytes, int rows, cudaStream_t st);

namespace {

int sm_count(int device) {
    static int cached[64] = {0};
    if (device >= 0 && device < 64 && cached[device]) return cached[device];
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    if (device >= 0 && device < 64) cached[device] = n;
    return n;
}

}  // namespace

// Enqueue the fused REF build of frame slots first .. first+count-1 (one launch when the slots are the
// handle's own contiguous buffers; one launch per frame when a slot reads an external device image).
cudaError_t launch_ref(sspyr_ctx* h, int first, int count, int outputs, int* launches) {
    const int frames = h->cfg.frames;
    int done = 0;
    while (done < count) {
        const int f0 = (first + done) % frames;
        int n = 1;
        if (!h->ext_in[f0])
            while (done + n < count && f0 + n < frames && !h->ext_in[f0 + n]) ++n;

        RefParams P{};
        size_t pitch_bytes = 0;
        P.img = frame_input(h, f0, &pitch_bytes);
        P.img_frame_stride = h->in_frame_bytes;
        P.out_frame_stride = h->frame_floats;
        P.img_pitch = (int)(pitch_bytes / h->elem_bytes);
        P.H = h->cfg.height;
        P.W = h->cfg.width;
        P.octaves = h->octaves;
        P.outputs = outputs;
        for (int o = 0; o < h->octaves; ++o) {
            const OctGeom& g = h->oct[o];
            P.oct[o].base = frame_out(h, f0) + g.off;
            P.oct[o].fw = h->d_tables + g.fw_off;
            P.oct[o].fh = h->d_tables + g.fh_off;
            P.oct[o].H = g.H;
            P.oct[o].W = g.W;
            P.oct[o].pitch = g.pitch;
            P.oct[o].plane = g.plane;
        }

        int rpt = h->tune.rows_per_thread;
        if (rpt <= 0) rpt = 2;                       // measured best on C2..C4 (profiles/, sweep_ref.py)
        rpt = rpt >= 2 ? 2 : 1;
        int threads = h->tune.block > 0 ? h->tune.block : 128;
        threads = threads > 128 ? 128 : (threads < 32 ? 32 : (threads / 32) * 32);   // kernels are built for <= 128
        const int W4 = (P.W + 3) >> 2;
        int bx = h->tune.bx;
        if (bx <= 0) {   // widest CTA row among {128,96,64,32} that wastes the fewest padded quads
            int best_waste = 1 << 30;
            for (int cand : {128, 96, 64, 32}) {
                if (cand > threads) continue;
                const int waste = (W4 + cand - 1) / cand * cand - W4;
                if (waste < best_waste) { best_waste = waste; bx = cand; }
            }
        }
        bx = bx > threads ? threads : (bx < 32 ? 32 : (bx / 32) * 32);
        const int by = threads / bx > 0 ? threads / bx : 1;
        const int row_groups = (P.H + rpt - 1) / rpt;
        const dim3 block(bx, by, 1);
        int gy = (row_groups + by - 1) / by;
        // persistent rows: cap the grid at `occ` CTAs per SM so that every thread walks several row groups and
        // prefetches the next one while it stores the current one
        if (h->tune.occ > 0) {
            const long long gx = (W4 + bx - 1) / bx;
            long long cap = ((long long)sm_count(h->device) * h->tune.occ + gx * n - 1) / (gx * n);
            if (cap < 1) cap = 1;
            if (gy > cap) gy = (int)cap;
        }
        const dim3 grid((W4 + bx - 1) / bx, gy, n);
        // Overlap with the previous launch only when it wrote other frame slots: two builds of the SAME slot
        // (e.g. a K0-only stage followed by the full build) must stay ordered or their stores would race.
        bool clash = false;
        for (int k = 0; k < n; ++k)
            for (int q = 0; q < h->last_count; ++q)
                clash |= ((f0 + k) % frames) == ((h->last_first + q) % frames);
        // A slot that reads a caller-produced device image (sspyr_set_input_device), or any build once raw device
        // pointers are in the caller's hands (strict_order), is launched WITHOUT the PDL attribute: the kernel reads
        // its whole input before its griddepcontrol.wait, and under PDL it could be scheduled at the producer's
        // implicit trigger, before the producer's stores are guaranteed visible.
        const bool pdl = h->tune.pdl != 0 && !clash && !h->ext_in[f0] && !h->strict_order;
        h->last_first = f0;
        h->last_count = n;

        cudaError_t e;
        switch (h->nl) {
            case 3: e = launch_ref_nl3(P, h->cfg.pixel_type, rpt, h->tune.occ > 0, grid, block, h->stream, pdl); break;
            case 4: e = launch_ref_nl4(P, h->cfg.pixel_type, rpt, h->tune.occ > 0, grid, block, h->stream, pdl); break;
            case 5: e = launch_ref_nl5(P, h->cfg.pixel_type, rpt, h->tune.occ > 0, grid, block, h->stream, pdl); break;
            case 6: e = launch_ref_nl6(P, h->cfg.pixel_type, rpt, h->tune.occ > 0, grid, block, h->stream, pdl); break;
            case 7: e = launch_ref_nl7(P, h->cfg.pixel_type, rpt, h->tune.occ > 0, grid, block, h->stream, pdl); break;
            case 8: e = launch_ref_nl8(P, h->cfg.pixel_type, rpt, h->tune.occ > 0, grid, block, h->stream, pdl); break;
            case 9: e = launch_ref_nl9(P, h->cfg.pixel_type, rpt, h->tune.occ > 0, grid, block, h->stream, pdl); break;
            case 10: e = launch_ref_nl10(P, h->cfg.pixel_type, rpt, h->tune.occ > 0, grid, block, h->stream, pdl); break;
            case 11: e = launch_ref_nl11(P, h->cfg.pixel_type, rpt, h->tune.occ > 0, grid, block, h->stream, pdl); break;
            case 12: e = launch_ref_nl12(P, h->cfg.pixel_type, rpt, h->tune.occ > 0, grid, block, h->stream, pdl); break;
            case 13: e = launch_ref_nl13(P, h->cfg.pixel_type, rpt, h->tune.occ > 0, grid, block, h->stream, pdl); break;
            case 14: e = launch_ref_nl14(P, h->cfg.pixel_type, rpt, h->tune.occ > 0, grid, block, h->stream, pdl); break;
            case 15: e = launch_ref_nl15(P, h->cfg.pixel_type, rpt, h->tune.occ > 0, grid, block, h->stream, pdl); break;
            case 16: e = launch_ref_nl16(P, h->cfg.pixel_type, rpt, h->tune.occ > 0, grid, block, h->stream, pdl); break;
            default: return cudaErrorInvalidValue;   // S+3 in 3..16 (create() rejects the rest)
        }
        if (e != cudaSuccess) return e;
        ++*launches;
        done += n;
    }
    const bool auto_pf = h->tune.prefetch_next < 0 && frames > 1 && h->in_frame_bytes <= (48u << 20) &&
                         h->in_frame_bytes >= (4u << 20);   // tiny frames are launch-bound: an extra launch costs more
    if (h->tune.prefetch_next > 0 || auto_pf) {  // warm L2 with the slot that is most likely built next
        const int nf = (first + count) % frames;
        size_t pitch_bytes = 0;
        const void* img = frame_input(h, nf, &pitch_bytes);
        const cudaError_t e = launch_ref_prefetch(img, pitch_bytes, (int)((size_t)h->cfg.width * h->elem_bytes), h->cfg.height, h->stream);
        if (e != cudaSuccess) return e;
        ++*launches;
    }
    return cudaSuccess;
}

}  // namespace sspyr
