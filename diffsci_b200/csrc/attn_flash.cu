// attn_flash.cu -- softmax(alpha Q K^T) V for one head as ONE tcgen05 kernel: the scores live in TMEM, the probabilities are
// written back into TMEM as the 16-bit A operand of the P V product, the output accumulates in TMEM; neither the L x L scores
// nor the probabilities ever reach shared memory or HBM (reference nets/attention.py:42-44, 93-102:
// torch.nn.MultiheadAttention(C, 1 head) -> SDPA).
//
// One CTA = 128 query rows of one sample; keys / values stream through in tiles of 128 keys.  C = 64 * NC channels (NC = 2, 4).
//   warp 0     TMA producer: Q once (NC K-major SWIZZLE_128B boxes of 128 rows x 64 channels), then a ring of stages in the
//              order the tensor pipe consumes them:  K_0, K_1, V_0, K_2, V_1, ...   (a K tile = 2 stages of NC/2 channel
//              chunks [128 keys x 64 ch]; a V tile = 2 stages of 64 keys x C as it lies = MN-major B operand)
//   warp 1     MMA issuer + TMEM owner:  S_{j+1} = Q K_{j+1}^T (SS form, N = 128) is issued BEFORE O += P_j V_j (TS form: A = P_j
//              from TMEM, N = C), so the tensor pipe works on the next scores while the softmax warps turn S_j into P_j
//   warps 2-5  softmax / correction / epilogue, one thread per query row (TMEM lane), no cross-thread reduction at all:
//              row max of the tile, p = exp2(c s - m) with a LAZY reference maximum m (it only moves when a tile's maximum
//              exceeds it by more than 2^8: the un-normalised p stays <= 256, exact in both 16-bit formats' range, and the O
//              accumulator is rescaled -- a TMEM round trip of the thread's own row -- only then), row sum in fp32,
//              P packed to 16 bits and stored over the first 64 columns of its own score buffer.
// TMEM (512 columns): O [0, C) | S_0 / P_0 [256, 384) | S_1 / P_1 [384, 512).
// Shared memory at C = 256: Q 64 KB + 5 stages x 32 KB = 224 KB.
#include "tc_common.cuh"

namespace dsk {

constexpr int AF_THREADS = 192;
constexpr uint32_t AF_S_COL0 = 256, AF_S_COLS = 128;

struct AttnFlashParams {
  int L, kv_tiles, q_tiles;
  float scale_log2;         // alpha * log2(e)
  void* out;                // [batch][L][ldo]
  int out_mode;             // 0: 16-bit (the operand format), 1: fp32, 2: split fp16 (hi at [0, C), 2^11 * lo at [C, 2C))
  int64_t ldo, strideO;
};

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

#define DSK_TMEM_LD_X32_NOWAIT(v, taddr)                                                                                   \
  asm volatile(                                                                                                            \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                            \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),      \
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),        \
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),        \
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                                           \
      : "r"(taddr))
#define DSK_TMEM_WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#define DSK_TMEM_ST_X32(taddr, v)                                                                                          \
  asm volatile(                                                                                                            \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                                      \
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr), \
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),      \
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),        \
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),        \
      "r"(v[31])                                                                                                           \
      : "memory")
#define DSK_TMEM_WAIT_ST() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool F16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (F16) { const __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <int NC, bool F16>
__global__ void __launch_bounds__(AF_THREADS, 1)
attn_flash_kernel(const __grid_constant__ CUtensorMap tmapQ, const __grid_constant__ CUtensorMap tmapK,
                  const __grid_constant__ CUtensorMap tmapV, const AttnFlashParams p) {
  constexpr int C = 64 * NC;
  constexpr int STAGE = NC * 8192;                         // half a K tile (NC/2 chunks of 16 KB) or half a V tile (64 keys x C)
  constexpr int NSTAGE = NC == 4 ? 5 : 8;
  constexpr int Q_BYTES = NC * 16384;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* q_smem = smem;
  uint8_t* ring = smem + Q_BYTES;
  __shared__ uint64_t full[NSTAGE], empty[NSTAGE], q_full, s_full[2], p_full[2], pv_done, o_done;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.q_tiles, b = blockIdx.x / p.q_tiles;
  const int T = p.kv_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); }
    mbar_init(&pv_done, 1);
    mbar_init(&o_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (elect_one_sync()) {
      auto tma3 = [&](const CUtensorMap* tm_, uint8_t* dst, uint64_t* bar, int c0, int c1, int c2) {
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                smem_u32(dst)),
            "l"(reinterpret_cast<uint64_t>(tm_)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
            : "memory");
      };
      mbar_expect_tx(&q_full, Q_BYTES);
#pragma unroll
      for (int c = 0; c < NC; ++c) tma3(&tmapQ, q_smem + c * 16384, &q_full, c * 64, qt * 128, b);
      uint32_t seq = 0;
      auto load_k = [&](int j) {
#pragma unroll
        for (int h = 0; h < 2; ++h, ++seq) {
          const uint32_t slot = seq % NSTAGE, ph = (seq / NSTAGE) & 1;
          mbar_wait(&empty[slot], ph ^ 1);
          mbar_expect_tx(&full[slot], STAGE);
#pragma unroll
          for (int cc = 0; cc < NC / 2; ++cc)
            tma3(&tmapK, ring + (size_t)slot * STAGE + cc * 16384, &full[slot], (h * (NC / 2) + cc) * 64, j * 128, b);
        }
      };
      auto load_v = [&](int j) {
#pragma unroll
        for (int h = 0; h < 2; ++h, ++seq) {
          const uint32_t slot = seq % NSTAGE, ph = (seq / NSTAGE) & 1;
          mbar_wait(&empty[slot], ph ^ 1);
          mbar_expect_tx(&full[slot], STAGE);
#pragma unroll
          for (int c = 0; c < NC; ++c) tma3(&tmapV, ring + (size_t)slot * STAGE + c * 8192, &full[slot], c * 64, j * 128 + h * 64, b);
        }
      };
      load_k(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) load_k(j + 1);
        load_v(j);
      }
    }
  } else if (warp == 1) {
    // S = Q K^T: both operands K-major, N = 128 keys.  O += P V: A = P from TMEM, B = V as it lies ([keys][channels], MN-major:
    // LBO = 8 KB between the 64-channel atoms, SBO = 1 KB between 8-key groups, 16 keys = 2 KB per k-step), N = C.
    const uint32_t idesc_s = umma_idesc_h16(128, 128, F16 ? 1 : 0);
    const uint32_t idesc_pv = umma_idesc_h16(C, 128, F16 ? 1 : 0) | (1u << 16);
    constexpr uint32_t HI = umma_desc_hi(1024);
    const uint32_t q16 = smem_u32(q_smem) >> 4, ring16 = smem_u32(ring) >> 4;
    uint32_t seq = 0;
    mbar_wait(&q_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    auto issue_s = [&](int j) {
      const uint32_t tacc = tmem_base + AF_S_COL0 + (uint32_t)(j & 1) * AF_S_COLS;
#pragma unroll
      for (int h = 0; h < 2; ++h, ++seq) {
        const uint32_t slot = seq % NSTAGE;
        mbar_wait(&full[slot], (seq / NSTAGE) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one_sync()) {
#pragma unroll
          for (int cc = 0; cc < NC / 2; ++cc) {
            const uint32_t a_lo = (q16 + (uint32_t)((h * (NC / 2) + cc) * (16384 >> 4))) | 0x10000u;
            const uint32_t b_lo = (ring16 + slot * (uint32_t)(STAGE >> 4) + (uint32_t)(cc * (16384 >> 4))) | 0x10000u;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_bf16(tacc, umma_desc64(a_lo + k4 * 2u, HI), umma_desc64(b_lo + k4 * 2u, HI), idesc_s, (h | cc | k4) ? 1u : 0u);
          }
          umma_commit(&empty[slot]);
        }
        __syncwarp();
      }
      if (elect_one_sync()) umma_commit(&s_full[j & 1]);
      __syncwarp();
    };
    auto issue_pv = [&](int j) {
      const uint32_t tp = tmem_base + AF_S_COL0 + (uint32_t)(j & 1) * AF_S_COLS;
      mbar_wait(&p_full[j & 1], (j >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int h = 0; h < 2; ++h, ++seq) {
        const uint32_t slot = seq % NSTAGE;
        mbar_wait(&full[slot], (seq / NSTAGE) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one_sync()) {
          const uint32_t b_lo = (ring16 + slot * (uint32_t)(STAGE >> 4)) | ((8192u >> 4) << 16);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_f16_ts(tmem_base, tp + (uint32_t)((h * 4 + k4) * 8), umma_desc64(b_lo + k4 * (2048u >> 4), HI), idesc_pv,
                        (j | h | k4) ? 1u : 0u);
          umma_commit(&empty[slot]);
        }
        __syncwarp();
      }
      if (elect_one_sync()) umma_commit(&pv_done);
      __syncwarp();
    };
    issue_s(0);
    for (int j = 0; j < T; ++j) {
      if (j + 1 < T) issue_s(j + 1);
      issue_pv(j);
    }
    if (elect_one_sync()) umma_commit(&o_done);
    __syncwarp();
  } else {
    const int q = warp & 3;                                  // TMEM lane quadrant of this warp
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float c = p.scale_log2;
    float ms = -INFINITY, l = 0.0f;                          // reference maximum (scaled, log2 units), row sum of p
    for (int j = 0; j < T; ++j) {
      const uint32_t ts = lane_addr + AF_S_COL0 + (uint32_t)(j & 1) * AF_S_COLS;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t s0[32], s1[32], s2[32], s3[32];
      DSK_TMEM_LD_X32_NOWAIT(s0, ts);
      DSK_TMEM_LD_X32_NOWAIT(s1, ts + 32);
      DSK_TMEM_LD_X32_NOWAIT(s2, ts + 64);
      DSK_TMEM_LD_X32_NOWAIT(s3, ts + 96);
      DSK_TMEM_WAIT_LD();
      const int nvalid = p.L - j * 128;                      // keys of this tile that exist (TMA zero-fills the rest)
      if (nvalid < 128) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          if (e >= nvalid) s0[e] = 0xff800000u;
          if (e + 32 >= nvalid) s1[e] = 0xff800000u;
          if (e + 64 >= nvalid) s2[e] = 0xff800000u;
          if (e + 96 >= nvalid) s3[e] = 0xff800000u;
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        mx0 = fmaxf(mx0, __uint_as_float(s0[e]));
        mx1 = fmaxf(mx1, __uint_as_float(s1[e]));
        mx2 = fmaxf(mx2, __uint_as_float(s2[e]));
        mx3 = fmaxf(mx3, __uint_as_float(s3[e]));
      }
      const float tmax = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c;
      if (j == 0) {
        ms = tmax;
      } else {
        const bool need = tmax > ms + 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // rare: move this row's reference maximum and rescale what has been accumulated under the old one.  P V_{j-1} must
          // have finished (pv_done phase j-1; S_j complete => every MMA up to P V_{j-2} is, so the parity is unambiguous) and
          // P V_j cannot start before this warp arrives on p_full below.
          mbar_wait(&pv_done, (j - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const float f = need ? ex2_approx(ms - tmax) : 1.0f;
          if (need) { ms = tmax; l *= f; }
#pragma unroll 1
          for (int c0 = 0; c0 < C; c0 += 32) {
            uint32_t o[32];
            DSK_TMEM_LD_X32_NOWAIT(o, lane_addr + c0);
            DSK_TMEM_WAIT_LD();
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
            DSK_TMEM_ST_X32(lane_addr + c0, o);
          }
        }
      }
      uint32_t pk[32];
      float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f, l3 = 0.0f;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float a0 = ex2_approx(fmaf(__uint_as_float(s0[2 * e]), c, -ms)), a1 = ex2_approx(fmaf(__uint_as_float(s0[2 * e + 1]), c, -ms));
        const float b0 = ex2_approx(fmaf(__uint_as_float(s1[2 * e]), c, -ms)), b1 = ex2_approx(fmaf(__uint_as_float(s1[2 * e + 1]), c, -ms));
        l0 += a0; l1 += a1; l2 += b0; l3 += b1;
        pk[e] = pack2<F16>(a0, a1);
        pk[16 + e] = pack2<F16>(b0, b1);
      }
      DSK_TMEM_ST_X32(ts, pk);
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float a0 = ex2_approx(fmaf(__uint_as_float(s2[2 * e]), c, -ms)), a1 = ex2_approx(fmaf(__uint_as_float(s2[2 * e + 1]), c, -ms));
        const float b0 = ex2_approx(fmaf(__uint_as_float(s3[2 * e]), c, -ms)), b1 = ex2_approx(fmaf(__uint_as_float(s3[2 * e + 1]), c, -ms));
        l0 += a0; l1 += a1; l2 += b0; l3 += b1;
        pk[e] = pack2<F16>(a0, a1);
        pk[16 + e] = pack2<F16>(b0, b1);
      }
      DSK_TMEM_ST_X32(ts + 32, pk);
      l += (l0 + l1) + (l2 + l3);
      DSK_TMEM_WAIT_ST();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j & 1]);
    }
    // epilogue: O / l -> global.  16-bit rows go through a per-warp staging area (the Q tile's shared memory, free now) so that
    // 4 lanes write 64 contiguous bytes of a row and a store instruction covers 8 rows (as gemm_tc's epilogue).
    mbar_wait(&o_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const float inv = 1.0f / l;
    const int m = qt * 128 + row, m_warp = qt * 128 + q * 32;
    const bool mvalid = m < p.L;
    const int64_t obase = (int64_t)b * p.strideO + (int64_t)m * p.ldo;
    uint4* stg = reinterpret_cast<uint4*>(q_smem) + q * 128;
#pragma unroll 1
    for (int c0 = 0; c0 < C; c0 += 32) {
      uint32_t o[32];
      DSK_TMEM_LD_X32_NOWAIT(o, lane_addr + c0);
      DSK_TMEM_WAIT_LD();
      float f[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(o[e]) * inv;
      if (p.out_mode == 1) {
        if (mvalid) {
          float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + obase + c0);
#pragma unroll
          for (int e4 = 0; e4 < 8; ++e4) op[e4] = make_float4(f[4 * e4], f[4 * e4 + 1], f[4 * e4 + 2], f[4 * e4 + 3]);
        }
      } else {
#pragma unroll 1
        for (int part = 0; part < (p.out_mode == 2 ? 2 : 1); ++part) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 pk;
            uint32_t* oh = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float v0 = f[g * 8 + 2 * e], v1 = f[g * 8 + 2 * e + 1];
              if (part) {                                    // lo half of the split format: 2^11 * (v - fp16(v))
                v0 = (v0 - __half2float(__float2half_rn(v0))) * 2048.0f;
                v1 = (v1 - __half2float(__float2half_rn(v1))) * 2048.0f;
              }
              oh[e] = pack2<F16>(v0, v1);
            }
            stg[lane * 4 + (g ^ ((lane >> 1) & 3))] = pk;
          }
          __syncwarp();
          const int jj = lane & 3;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int r = 8 * k + (lane >> 2);
            const uint4 val = stg[r * 4 + (jj ^ ((r >> 1) & 3))];
            if (m_warp + r < p.L)
              *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + (int64_t)b * p.strideO + (int64_t)(m_warp + r) * p.ldo +
                                        part * C + c0 + jj * 8) = val;
          }
          __syncwarp();
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

template <int NC, bool F16>
static int launch_attn_flash(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFlashParams& p, int batch,
                             cudaStream_t st) {
  constexpr int NSTAGE = NC == 4 ? 5 : 8;
  const size_t smem = (size_t)NC * 16384 + (size_t)NSTAGE * NC * 8192 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_flash_kernel<NC, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("attn_flash: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return DSK_ERR_CUDA; }
    configured = true;
  }
  DSK_LAUNCH((attn_flash_kernel<NC, F16>), p.q_tiles * batch, AF_THREADS, smem, st, tq, tk, tv, p);
  return DSK_OK;
}

}  // namespace dsk

using namespace dsk;

extern "C" int dsk_attn_flash(const void* Q, const void* K, const void* V, void* O, int L, int C, int64_t ldq, int64_t ldk, int64_t ldv,
                              int64_t ldo, int64_t strideQ, int64_t strideK, int64_t strideV, int64_t strideO, int batch, float alpha,
                              int dtype, int out_mode, void* stream) {
  DSK_REQUIRE(is_h16(dtype), "dsk_attn_flash: dtype must be DSK_BF16 or DSK_F16, got %d", dtype);
  DSK_REQUIRE(Q && K && V && O && L > 0 && batch > 0, "dsk_attn_flash: bad arguments");
  DSK_REQUIRE(C == 128 || C == 256, "dsk_attn_flash: C = %d (128 and 256 are built: the output accumulator takes C TMEM columns)", C);
  DSK_REQUIRE(out_mode >= 0 && out_mode <= 2 && (out_mode != 2 || dtype == DSK_F16), "dsk_attn_flash: bad out_mode %d", out_mode);
  DSK_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && strideQ % 8 == 0 && strideK % 8 == 0 && strideV % 8 == 0 &&
                  strideO % 8 == 0,
              "dsk_attn_flash: leading dimensions and batch strides must be multiples of 8 elements");
  DSK_REQUIRE((((uintptr_t)Q | (uintptr_t)K | (uintptr_t)V | (uintptr_t)O) & 15) == 0, "dsk_attn_flash: 16-byte alignment");
  EncodeTiledFn encode = get_encode();
  DSK_REQUIRE(encode != nullptr, "dsk_attn_flash: cuTensorMapEncodeTiled is unavailable");
  auto make = [&](CUtensorMap* tm, const void* base, int64_t ld, int64_t stride, int box_rows) -> bool {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? stride : (int64_t)L * ld) * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return encode(tm, tmap_h16(dtype), 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  CUtensorMap tq, tk, tv;
  DSK_REQUIRE(make(&tq, Q, ldq, strideQ, 128), "dsk_attn_flash: tensor map for Q failed");
  DSK_REQUIRE(make(&tk, K, ldk, strideK, 128), "dsk_attn_flash: tensor map for K failed");
  DSK_REQUIRE(make(&tv, V, ldv, strideV, 64), "dsk_attn_flash: tensor map for V failed");
  AttnFlashParams p;
  p.L = L;
  p.kv_tiles = (L + 127) / 128;
  p.q_tiles = (L + 127) / 128;
  p.scale_log2 = alpha * 1.4426950408889634f;
  p.out = O; p.out_mode = out_mode; p.ldo = ldo; p.strideO = strideO;
  cudaStream_t st = as_stream(stream);
  const bool f16 = dtype == DSK_F16;
  if (C == 256) return f16 ? launch_attn_flash<4, true>(tq, tk, tv, p, batch, st) : launch_attn_flash<4, false>(tq, tk, tv, p, batch, st);
  return f16 ? launch_attn_flash<2, true>(tq, tk, tv, p, batch, st) : launch_attn_flash<2, false>(tq, tk, tv, p, batch, st);
}
