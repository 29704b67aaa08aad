// tcgen05 (3xTF32) path of the gathered implicit-GEMM convolution — placeholder until the
// tensor-core kernel lands; the entry point reports UNSUPPORTED so that callers fail loudly.
#include "common.cuh"
int sassd_gconv_tc(const sassd_gconv_desc*, const float*, const float*, const float*, const float*, const int32_t*,
                   const int32_t*, float*, cudaStream_t) {
    return SASSD_ERR_UNSUPPORTED;
}
