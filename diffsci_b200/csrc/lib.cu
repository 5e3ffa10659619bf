// lib.cu -- library-level entry points: version, thread-local error text, launch counter,
// device check, and the dsk_conv_fwd dispatcher (FFMA fp32-parity kernel vs tcgen05 bf16 kernel).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace dsk {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace dsk

using namespace dsk;

extern "C" int dsk_conv_fwd_ffma(const dsk_conv_desc*, const void*, const void*, const float*, const float*,
                                 const void*, void*, void*);
extern "C" int dsk_conv_fwd_tc(const dsk_conv_desc*, const void*, const void*, const float*, const float*, const void*,
                               void*, void*);

extern "C" int dsk_version(void) { return 100; }
extern "C" const char* dsk_last_error(void) { return g_err; }
extern "C" uint64_t dsk_launch_count(void) { return g_launches.load(); }

extern "C" int dsk_check_device(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    set_error("dsk_check_device: %s", cudaGetErrorString(e));
    return DSK_ERR_CUDA;
  }
  if (prop.major != 10) {
    set_error("dsk_check_device: device %d is sm_%d%d; libdiffsci_b200 is built for sm_100a only (no fallback)", device,
              prop.major, prop.minor);
    return DSK_ERR_UNSUPPORTED;
  }
  return DSK_OK;
}

extern "C" int dsk_conv_fwd(const dsk_conv_desc* d, const void* in, const void* w, const float* bias,
                            const float* chan_bias, const void* residual, void* out, void* stream) {
  DSK_REQUIRE(d != nullptr, "dsk_conv_fwd: null descriptor");
  if (d->w_dtype == DSK_F32) return dsk_conv_fwd_ffma(d, in, w, bias, chan_bias, residual, out, stream);
  if (d->w_dtype == DSK_BF16 || d->w_dtype == DSK_F16 || d->w_dtype == DSK_SPLIT_F16)
    return dsk_conv_fwd_tc(d, in, w, bias, chan_bias, residual, out, stream);
  set_error("dsk_conv_fwd: bad w_dtype %d", d->w_dtype);
  return DSK_ERR_ARG;
}
