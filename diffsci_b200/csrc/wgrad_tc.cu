// wgrad_tc.cu -- K2 (bf16 throughput mode): convolution weight gradient on the tcgen05 tensor cores.
// Placeholder dispatch: shapes not taken here fall back to the CUDA-core split-K kernel in gemm_ffma.cu
// (same arithmetic, fp32 accumulation), which is the fp32-parity path anyway.
#include "tc_common.cuh"

using namespace dsk;

extern "C" int dsk_conv_wgrad_tc(const dsk_conv_desc* d, const void* x, const void* dy, float* dw, void* ws, int accumulate,
                                 void* stream) {
  (void)d; (void)x; (void)dy; (void)dw; (void)ws; (void)accumulate; (void)stream;
  return DSK_ERR_UNSUPPORTED;
}
