// csrc/ref_inst.cu -- one translation unit per level count (compiled with -DSSPYR_NL=3..8, in
// parallel): instantiates the fused REF kernel of ref_kernel.cuh for every pixel type / rows-per-thread.
#include "ref_kernel.cuh"

#ifndef SSPYR_NL
#error "compile with -DSSPYR_NL=<levels>"
#endif
#define SSPYR_CAT2(a, b) a##b
#define SSPYR_CAT(a, b) SSPYR_CAT2(a, b)

namespace sspyr {
#if SSPYR_NL == 6
cudaError_t launch_ref_prefetch(const void* img, size_t pitch_bytes, int row_bytes, int rows, cudaStream_t st) {
    return launch_prefetch(img, pitch_bytes, row_bytes, rows, st);
}
#endif
cudaError_t SSPYR_CAT(launch_ref_nl, SSPYR_NL)(const RefParams& P, int pix, int rpt, bool walk, dim3 grid, dim3 block,
                                               cudaStream_t st, bool pdl) {
    return launch_pix<SSPYR_NL>(P, pix, rpt, walk, grid, block, st, pdl);
}
}  // namespace sspyr
