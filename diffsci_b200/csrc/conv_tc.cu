// conv_tc.cu -- K1 (bf16 throughput mode): tcgen05/TMEM implicit-GEMM convolution fed by TMA.
// (placeholder until the tcgen05 kernel lands: reports UNSUPPORTED, never falls back silently)
#include "common.cuh"
using namespace dsk;
extern "C" int dsk_conv_fwd_tc(const dsk_conv_desc* d, const void* in, const void* w, const float* bias,
                               const float* chan_bias, const void* residual, void* out, void* stream) {
  (void)d; (void)in; (void)w; (void)bias; (void)chan_bias; (void)residual; (void)out; (void)stream;
  set_error("dsk_conv_fwd: the tcgen05 (bf16-weight) path does not support this shape yet");
  return DSK_ERR_UNSUPPORTED;
}
