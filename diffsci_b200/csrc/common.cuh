// common.cuh -- shared helpers for libdiffsci_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
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
