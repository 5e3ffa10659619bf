// tc_common.cuh -- tcgen05 / TMEM / TMA / mbarrier primitives shared by the tensor-core kernels (sm_100a).
// Bit layouts follow cute/arch/mma_sm100_desc.hpp; semantics verified on B200 by tests/cuda/umma_probe.cu.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace dsk {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
#ifdef DSK_DEBUG_TIMEOUT
  for (long spin = 0; spin < (1L << 28) && !ok; ++spin)
#else
  while (!ok)
#endif
  {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  }
#ifdef DSK_DEBUG_TIMEOUT
  if (!ok) { printf("conv_tc: mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x); __trap(); }
#endif
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp bit layout)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}


#define DSK_TMEM_LD_X32(v, taddr)                                                                                          \
  asm volatile(                                                                                                            \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                            \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),      \
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),        \
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),        \
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                                           \
      : "r"(taddr));                                                                                                       \
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

// One lane of a fully-active warp, chosen by the hardware: the form the compiler turns into a plain predicated
// UTCHMMA / UTMALDG (an `if (lane == 0)` region makes it wrap every uniform-datapath op in a waterfall loop).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
// Descriptor words for K-major SWIZZLE_128B operands.  lo = start address >> 4 | LBO(=1) << 16 ; hi = SBO >> 4 |
// version(1) << 14 | SWIZZLE_128B(2) << 29.  Shared memory is < 256 KB so (addr >> 4) never carries into the LBO field
// and a view is advanced by adding (bytes >> 4) to the low word.
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return (saddr >> 4) | 0x10000u; }
__device__ __forceinline__ constexpr uint32_t umma_desc_hi(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ uint64_t umma_desc64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

#define DSK_TMEM_LD_X16(v, taddr)                                                                                          \
  asm volatile(                                                                                                            \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"              \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),      \
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                                                 \
      : "r"(taddr));                                                                                                       \
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

// bf16 x bf16 -> fp32, both operands K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// kind::f16 instruction descriptor for either 16-bit format (a_format bits 7-9, b_format bits 10-12: 0 = F16, 1 = BF16),
// fp32 accumulate, both operands K-major
__device__ __forceinline__ uint32_t umma_idesc_h16(int n, int m, int f16) {
  const uint32_t fmt = f16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
inline CUtensorMapDataType tmap_h16(int dtype) { return dtype == DSK_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    cudaDriverEntryPointQueryResult q;
    void* ptr = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFn)ptr;
  }
  return fn;
}


}  // namespace dsk
