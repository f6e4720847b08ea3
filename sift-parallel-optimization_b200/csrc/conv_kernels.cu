// csrc/conv_kernels.cu -- CONV mode (true separable Gaussian blur chain) and the DoG extremum scan.
// Placeholder until the REF path is measured: every entry point reports "unsupported".
#include "sspyr_internal.h"

namespace sspyr {

cudaError_t launch_conv(const sspyr_ctx*, int, int*) { return cudaErrorNotSupported; }
cudaError_t launch_conv_step(const sspyr_ctx*, int, int, int, int*) { return cudaErrorNotSupported; }
cudaError_t launch_extrema(const sspyr_ctx*, int, int*) { return cudaErrorNotSupported; }

}  // namespace sspyr

extern "C" {

int sspyr_halo_rows(sspyr_handle h, int, int, int*) { return h ? SSPYR_ERR_UNSUPPORTED : SSPYR_ERR_ARG; }
int sspyr_halo_ptrs(sspyr_handle h, int, int, int, void**, void**, void**, void**, size_t*) {
    return h ? SSPYR_ERR_UNSUPPORTED : SSPYR_ERR_ARG;
}
int sspyr_conv_step(sspyr_handle h, int, int, int) { return h ? SSPYR_ERR_UNSUPPORTED : SSPYR_ERR_ARG; }

}  // extern "C"
