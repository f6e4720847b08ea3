// csrc/conv_inst.cu -- one translation unit per tap radius (compiled with -DSSPYR_R=<radius>, in parallel):
// instantiates the fused CONV level kernels (conv_kernel.cuh, conv_march.cuh) for every input kind.
#include "conv_march.cuh"

#ifndef SSPYR_R
#error "compile with -DSSPYR_R=<radius>"
#endif
#define SSPYR_CAT2(a, b) a##b
#define SSPYR_CAT(a, b) SSPYR_CAT2(a, b)

namespace sspyr {
cudaError_t SSPYR_CAT(launch_conv_r, SSPYR_R)(const ConvParams& P, int src_kind, int variant, cudaStream_t st, int device,
                                              int frames, int sms) {
    return launch_conv_src<SSPYR_R>(P, src_kind, variant, st, device, frames, sms);
}
#if SSPYR_R <= 12
cudaError_t SSPYR_CAT(launch_march_r, SSPYR_R)(const ConvParams& P, int src_kind, cudaStream_t st, int device, int frames,
                                               const CUtensorMap* tmap, int seg_rows, bool pdl) {
    return launch_march_src<SSPYR_R>(P, src_kind, st, device, frames, tmap, seg_rows, pdl);
}
int SSPYR_CAT(march_box_cols_r, SSPYR_R)() { return conv_pitch_in<SSPYR_R>(); }
#endif
}  // namespace sspyr
