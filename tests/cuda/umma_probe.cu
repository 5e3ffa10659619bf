// umma_probe.cu -- standalone experiments that pin down the tcgen05 / TMA semantics the implicit-GEMM
// convolution relies on (run on a B200: `nvcc -gencode arch=compute_100a,code=sm_100a ... && ./umma_probe`).
//
//  P1  UMMA K-major SWIZZLE_128B operand whose start address is shifted by r0 rows of 128 B
//      (not 1024-B aligned), with base_offset = 0 and base_offset = (r0 & 7), and with SBO = 1024 / 1280:
//      which combination reads "row r of the 128B-swizzled buffer" correctly?
//  P2  TMA tiled loads with SWIZZLE_128B: byte layout in smem, signed / out-of-bounds coordinates (zero
//      fill), 4-D boxes whose rows are 128 B.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x)                                                                            \
  do {                                                                                   \
    cudaError_t e_ = (x);                                                                \
    if (e_ != cudaSuccess) {                                                             \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);    \
      exit(2);                                                                           \
    }                                                                                    \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  for (int spin = 0; spin < (1 << 26) && !ok; ++spin) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  }
  if (!ok) { printf("mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                          // LBO (ignored for swizzled K-major), CUTLASS sets 1
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}

constexpr int ROWS = 176;   // rows of 128 B in the A buffer
constexpr int N = 64;

// ---------------------------------------------------------------------------------- P1
__global__ void __launch_bounds__(128) probe_mma(const __nv_bfloat16* __restrict__ Ag, const __nv_bfloat16* __restrict__ Bg,
                                                  float* __restrict__ Dg, int r0, int sbo_bytes, int base_off) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* As = smem;                       // ROWS x 128 B, address-swizzled
  uint8_t* Bs = smem + 24 * 1024;           // 64 x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  // software 128B swizzle keyed on the absolute smem address: chunk' = chunk ^ ((addr >> 7) & 7)
  for (int i = tid; i < ROWS * 8; i += 128) {
    int r = i >> 3, c = i & 7;
    uint4 v = reinterpret_cast<const uint4*>(Ag)[r * 8 + c];
    uint32_t off = r * 128;
    uint32_t phase = ((smem_u32(As) + off) >> 7) & 7;
    *reinterpret_cast<uint4*>(As + off + ((c ^ phase) << 4)) = v;
  }
  for (int i = tid; i < N * 8; i += 128) {
    int r = i >> 3, c = i & 7;
    uint4 v = reinterpret_cast<const uint4*>(Bg)[r * 8 + c];
    uint32_t off = r * 128;
    uint32_t phase = ((smem_u32(Bs) + off) >> 7) & 7;
    *reinterpret_cast<uint4*>(Bs + off + ((c ^ phase) << 4)) = v;
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      uint64_t da = make_desc(smem_u32(As) + r0 * 128 + k * 32, sbo_bytes, base_off);
      uint64_t db = make_desc(smem_u32(Bs) + k * 32, 1024, 0);
      uint32_t acc = k > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // each warp reads its 32 lanes x 64 columns
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) Dg[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

// ---------------------------------------------------------------------------------- P2
// TMA: 4-D tensor (C=64 bf16, W, H, Dd); box (64, bw, bh, 1) at signed coordinates -> dump raw smem
__global__ void __launch_bounds__(32) probe_tma(const __grid_constant__ CUtensorMap tmap, uint8_t* __restrict__ dump, int bytes,
                                                 int cw, int ch, int cd) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&bar, bytes);
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            smem_u32(smem)),
        "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(0), "r"(cw), "r"(ch), "r"(cd), "r"(smem_u32(&bar))
        : "memory");
  }
  __syncwarp();
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < bytes; i += 32) dump[i] = smem[i];
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float bf(__nv_bfloat16 v) { return __bfloat162float(v); }

int main() {
  // ------------------------------------------------ P1
  std::vector<__nv_bfloat16> A(ROWS * 64), B(N * 64);
  srand(1);
  for (auto& v : A) v = __float2bfloat16((float)(rand() % 17 - 8) / 8.0f);
  for (auto& v : B) v = __float2bfloat16((float)(rand() % 13 - 6) / 4.0f);
  __nv_bfloat16 *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, A.size() * 2));
  CK(cudaMalloc(&dB, B.size() * 2));
  CK(cudaMalloc(&dD, 128 * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
  std::vector<float> D(128 * N);
  int r0s[] = {0, 8, 1, 3, 10, 21};
  int sbos[] = {1024, 1280};
  for (int sbo : sbos)
    for (int r0 : r0s)
      for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
        int base_off = bo_mode ? (r0 & 7) : 0;
        if (bo_mode && base_off == 0) continue;
        CK(cudaMemset(dD, 0, 128 * N * 4));
        probe_mma<<<1, 128, 40 * 1024>>>(dA, dB, dD, r0, sbo, base_off);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("P1 r0=%d sbo=%d base_off=%d: %s\n", r0, sbo, base_off, cudaGetErrorString(e)); return 3; }
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        // hypothesis H_addr: row i of the operand is buffer row r0 + (i/8)*(sbo/128) + i%8 ("address based swizzle")
        double err = 0;
        for (int i = 0; i < 128; ++i) {
          int r = r0 + (i / 8) * (sbo / 128) + (i % 8);
          for (int n = 0; n < N; ++n) {
            float ref = 0;
            for (int c = 0; c < 64; ++c) ref += bf(A[r * 64 + c]) * bf(B[n * 64 + c]);
            err = fmax(err, fabs(ref - D[i * N + n]));
          }
        }
        printf("P1 r0=%2d sbo=%4d base_off=%d : max|err| vs address-based rows = %g  %s\n", r0, sbo, base_off, err,
               err < 1e-3 ? "OK" : "MISMATCH");
      }
  // ------------------------------------------------ P2
  EncodeTiled encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 4; }
  const int C = 64, W = 12, H = 9, Dd = 3;
  std::vector<__nv_bfloat16> T((size_t)C * W * H * Dd);
  for (int d = 0; d < Dd; ++d)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w)
        for (int c = 0; c < C; ++c) T[(((size_t)d * H + h) * W + w) * C + c] = __float2bfloat16((float)(d * 100 + h * 10 + w) + c / 64.0f);
  __nv_bfloat16* dT;
  CK(cudaMalloc(&dT, T.size() * 2));
  CK(cudaMemcpy(dT, T.data(), T.size() * 2, cudaMemcpyHostToDevice));
  const int bw = 10, bh = 4;
  CUtensorMap tmap;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Dd};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, bw, bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dT, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)cr); return 5; }
  const int bytes = 128 * bw * bh;
  uint8_t* ddump;
  CK(cudaMalloc(&ddump, bytes));
  CK(cudaFuncSetAttribute(probe_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024));
  int coords[][3] = {{0, 0, 0}, {-1, -1, 1}, {5, 7, 2}};
  for (auto& co : coords) {
    probe_tma<<<1, 32, 16 * 1024>>>(tmap, ddump, bytes, co[0], co[1], co[2]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("P2 coords (%d,%d,%d): %s\n", co[0], co[1], co[2], cudaGetErrorString(e)); return 6; }
    std::vector<uint8_t> dump(bytes);
    CK(cudaMemcpy(dump.data(), ddump, bytes, cudaMemcpyDeviceToHost));
    // hypothesis: box row q = (hb*bw + wb) lives at q*128, 16B chunk c stored at chunk (c ^ (q & 7)); OOB -> zeros
    int bad = 0;
    for (int hb = 0; hb < bh; ++hb)
      for (int wb = 0; wb < bw; ++wb) {
        int q = hb * bw + wb, w = co[0] + wb, h = co[1] + hb, d = co[2];
        bool in = w >= 0 && w < W && h >= 0 && h < H && d >= 0 && d < Dd;
        for (int c = 0; c < 64; ++c) {
          int chunk = (c / 8) ^ (q & 7);
          __nv_bfloat16 got = *reinterpret_cast<__nv_bfloat16*>(&dump[q * 128 + chunk * 16 + (c % 8) * 2]);
          float want = in ? bf(T[(((size_t)d * H + h) * W + w) * C + c]) : 0.0f;
          if (bf(got) != want) ++bad;
        }
      }
    printf("P2 TMA box(64,%d,%d,1) at (w=%d,h=%d,d=%d): %d mismatches vs row-major rows + address swizzle + zero OOB  %s\n", bw, bh,
           co[0], co[1], co[2], bad, bad == 0 ? "OK" : "MISMATCH");
  }
  printf("probe done\n");
  return 0;
}
