// common.cuh -- shared helpers for libdiffsci_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/diffsci_b200.h"

#define DSK_NUM_SMS 148

namespace dsk {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define DSK_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      dsk::set_error(__VA_ARGS__);             \
      return DSK_ERR_ARG;                      \
    }                                          \
  } while (0)

// every kernel launch goes through this macro: counts launches (gpu_launches evidence) and
// turns launch-configuration errors into an error return instead of a silent no-op.
#define DSK_LAUNCH(kernel, grid, block, smem, stream, ...)                                       \
  do {                                                                                           \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                  \
    dsk::g_launches.fetch_add(1, std::memory_order_relaxed);                                     \
    cudaError_t _e = cudaPeekAtLastError();                                                      \
    if (_e != cudaSuccess) {                                                                     \
      dsk::set_error("%s launch failed: %s (%s:%d)", #kernel, cudaGetErrorString(_e), __FILE__,  \
                     __LINE__);                                                                  \
      return DSK_ERR_CUDA;                                                                       \
    }                                                                                            \
  } while (0)

// ---- dtype helpers ---------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// ---- the two 16-bit formats behind one interface ----------------------------------------------------------------------
// Templated kernels use H16<T> (T = __nv_bfloat16 | __half); the tensor-core kernels move 16-bit data as raw words and
// take the format as a warp-uniform run-time flag (f16 != 0: IEEE half, else bfloat16).
inline bool is_h16(int dtype) { return dtype == DSK_BF16 || dtype == DSK_F16; }
template <typename T> struct H16;
template <> struct H16<__nv_bfloat16> {
  typedef __nv_bfloat162 T2;
  static __device__ __forceinline__ T2 pack(float a, float b) { return __floats2bfloat162_rn(a, b); }
  static __device__ __forceinline__ float lo(T2 v) { return __low2float(v); }
  static __device__ __forceinline__ float hi(T2 v) { return __high2float(v); }
};
template <> struct H16<__half> {
  typedef __half2 T2;
  static __device__ __forceinline__ T2 pack(float a, float b) { return __floats2half2_rn(a, b); }
  static __device__ __forceinline__ float lo(T2 v) { return __low2float(v); }
  static __device__ __forceinline__ float hi(T2 v) { return __high2float(v); }
};
__device__ __forceinline__ uint32_t pack_h2(float a, float b, int f16) {
  if (f16) { const __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v, int f16) {
  if (f16) return __half22float2(*reinterpret_cast<const __half2*>(&v));
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__device__ __forceinline__ uint16_t pack_h1(float a, int f16) {
  if (f16) { const __half h = __float2half_rn(a); return *reinterpret_cast<const uint16_t*>(&h); }
  const __nv_bfloat16 h = __float2bfloat16_rn(a);
  return *reinterpret_cast<const uint16_t*>(&h);
}
__device__ __forceinline__ float unpack_h1(uint16_t v, int f16) {
  if (f16) return __half2float(*reinterpret_cast<const __half*>(&v));
  return __uint_as_float((uint32_t)v << 16);
}

// fmt: format of the 16-bit copy -- DSK_BF16, DSK_F16, or DSK_SPLIT_F16 (rows of 2 * K: [hi | lo], the B operand of the split
// tensor-core kernels)
__device__ __forceinline__ void put16(uint16_t* o16, int64_t row, int K, int k, float v, int fmt) {
  if (fmt == DSK_SPLIT_F16) {
    const __half h = __float2half_rn(v), l = __float2half_rn((v - __half2float(h)) * 2048.0f);   // lo stored times 2^11
    o16[row * 2 * K + k] = *reinterpret_cast<const uint16_t*>(&h);
    o16[row * 2 * K + K + k] = *reinterpret_cast<const uint16_t*>(&l);
  } else {
    o16[row * K + k] = pack_h1(v, fmt == DSK_F16);
  }
}

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// grid sized as a multiple of the SM count for grid-stride bandwidth kernels
inline int grid_for(int64_t work_items, int per_block, int max_waves = 8) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)DSK_NUM_SMS * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// y[B, D + 2*(ndim==3), H+2, W+2, C] = x[B, D, H, W, C] wrapped by one pixel per spatial axis (elementwise.cu)
int pad_circular_launch(const void* x, void* y, int B, int D, int H, int W, int C, int ndim, int dtype, cudaStream_t st);
}  // namespace dsk
