// gemm_tc.cu -- batched bf16 GEMM on tcgen05/TMEM fed by TMA:  C[b] = alpha * A[b] * B[b]^T (+ bias) (+ residual)
// A: [M, K] row-major (K contiguous, leading dimension lda), B: [N, K] row-major (ldb): both K-major operands.
// transA / transB: the operand is stored [K, M] / [K, N] (M / N contiguous) and is consumed as it lies, as an MN-MAJOR
// UMMA operand (64-element x 64-row SWIZZLE_128B boxes, LBO = 8 KB between the 64-wide atoms, SBO = 1 KB between
// 8-row K groups) -- the A^T B products of the backward passes (weight gradients, dV = P^T dO, dK = dS^T Q) need no
// transposition pass.
//
// Used for everything GEMM-shaped around the attention block (nn.MultiheadAttention(C, 1 head), reference
// nets/attention.py:42-44,68): packed Q|K projection, V^T projection (row bias), Q K^T, P V and the output
// projection; and for 1x1 convolutions.  Same skeleton as conv_tc.cu: persistent CTAs, warp-specialised
// TMA producer / MMA issuer / 4 epilogue warps, SWIZZLE_128B K-major tiles, double-buffered TMEM accumulators.
#include "tc_common.cuh"

namespace dsk {

constexpr int GT_THREADS = 192;   // warp 0: TMA, warp 1: MMA + TMEM owner, warps 2-5: epilogue
constexpr int GT_STAGES = 4;

struct GemmTcParams {
  int M, N, K, batch;
  int a_bmul, b_bmul;       // 0: operand shared by every batch (stride 0), 1: per-batch
  int tiles_m, tiles_n, total_tiles;
  float alpha;
  const float* bias;        // [N] (bias_rows == 0) or [M] (bias_rows == 1), or null
  int bias_rows;
  const uint16_t* residual; // [batch][M][ldc] in the 16-bit format (fp32 when res_f32) or null
  void* out;                // 16-bit or fp32 [batch][M][ldc]
  int out_f32;
  int f16;                  // 16-bit format of operands / 16-bit outputs: 0 bfloat16, 1 IEEE half
  int res_f32;              // residual is fp32 (split parity mode)
  // split-fp16 operands (DSK_SPLIT_F16): `vparts` virtual K chunks per real 64-wide chunk accumulate (A hi, B hi), (A hi, B lo),
  // (A lo, B hi) [vparts = 3] or (A hi, B), (A lo, B) [vparts = 2]; a_lo / b_lo: offset of the lo half along the operand's
  // CONTIGUOUS coordinate (k for K-major operands, m / n for MN-major ones)
  int vparts, a_lo, b_lo;
  int64_t ldc, strideC;
  int transA, transB;       // operand stored [K][M] / [K][N] (MN-major)
  int epi;                  // 0: C = alpha A B^T (+bias, +residual);  1: row statistics only;  2: softmax probabilities
  float2* rowstat;          // epi 1: [batch][M][tiles_n] (max, sum exp(s - max)) of s = alpha * acc over this tile's columns
  const float2* rowfinal;   // epi 2: [batch][M] (row max, 1 / row sum): out = exp(s - max) / sum as bf16
};
constexpr int EPI_NORMAL = 0, EPI_ROWSTAT = 1, EPI_SOFTMAX = 2;

template <int N_TILE>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const GemmTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int A_BYTES = 128 * 128, B_BYTES = N_TILE * 128, STAGE = A_BYTES + B_BYTES;
  __shared__ uint64_t full[GT_STAGES], empty[GT_STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint4 stage_s[4][32 * 8];                   // per epilogue warp: 32 rows x 128 B (coalescing stage of the stores)
  // split operands: the hi*lo + lo*hi products (lo halves are stored times 2^11) go to a second accumulator set; the epilogue
  // combines acc = hh + 2^-11 lo (host: N_TILE <= 128 in split mode)
  const uint32_t nsets = p.vparts > 1 ? 2u : 1u;
  const uint32_t BUF_COLS = nsets * N_TILE;
  const uint32_t TMEM_COLS = 2 * BUF_COLS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kchunks = ((p.K + 63) / 64) * p.vparts;      // virtual K chunks

  if (threadIdx.x == 0) {
    for (int i = 0; i < GT_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (elect_one_sync()) {
      uint32_t seq = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int tn = t % p.tiles_n, tm = (t / p.tiles_n) % p.tiles_m, b = t / (p.tiles_n * p.tiles_m);
        int kc = 0, part = -1;
        for (int vkc = 0; vkc < kchunks; ++vkc, ++seq) {
          if (++part == p.vparts) { part = 0; ++kc; }              // (real chunk, part) without a division per chunk
          const int ao = (p.vparts > 1 && part == p.vparts - 1) ? p.a_lo : 0, bo = (p.vparts == 3 && part == 1) ? p.b_lo : 0;
          const uint32_t slot = seq % GT_STAGES, ph = (seq / GT_STAGES) & 1;
          mbar_wait(&empty[slot], ph ^ 1);
          mbar_expect_tx(&full[slot], STAGE);
          uint8_t* sa = smem + (size_t)slot * STAGE;
          auto tma3 = [&](const CUtensorMap* tm_, uint8_t* dst, int c0, int c1, int c2) {
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                    smem_u32(dst)),
                "l"(reinterpret_cast<uint64_t>(tm_)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&full[slot]))
                : "memory");
          };
          if (!p.transA) {
            tma3(&tmapA, sa, kc * 64 + ao, tm * 128, b * p.a_bmul);
          } else {
            tma3(&tmapA, sa, tm * 128 + ao, kc * 64, b * p.a_bmul);
            tma3(&tmapA, sa + 8192, tm * 128 + 64 + ao, kc * 64, b * p.a_bmul);
          }
          if (!p.transB) {
            tma3(&tmapB, sa + A_BYTES, kc * 64 + bo, tn * N_TILE, b * p.b_bmul);
          } else {
#pragma unroll
            for (int h = 0; h < N_TILE / 64; ++h) tma3(&tmapB, sa + A_BYTES + h * 8192, tn * N_TILE + h * 64 + bo, kc * 64, b * p.b_bmul);
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_h16(N_TILE, 128, p.f16) | (p.transA ? 1u << 15 : 0u) | (p.transB ? 1u << 16 : 0u);
    constexpr uint32_t HI = umma_desc_hi(1024);
    // descriptor low word: K-major: LBO field = 1 (unused), k-step of 16 elements = 32 B; MN-major: LBO = 8 KB, k-step of
    // 16 rows = 2 KB
    const uint32_t a_lbo = p.transA ? (8192u >> 4) << 16 : 0x10000u, a_step = p.transA ? 2048u >> 4 : 2u;
    const uint32_t b_lbo = p.transB ? (8192u >> 4) << 16 : 0x10000u, b_step = p.transB ? 2048u >> 4 : 2u;
    uint32_t seq = 0, it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      mbar_wait(&acc_empty[as], aph ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_acc = tmem_base + as * BUF_COLS;
      uint32_t used = 0;
      int part = -1;
      for (int kc = 0; kc < kchunks; ++kc, ++seq) {
        if (++part == p.vparts) part = 0;
        const uint32_t slot = seq % GT_STAGES;
        mbar_wait(&full[slot], (seq / GT_STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa16 = smem_u32(smem + (size_t)slot * STAGE) >> 4;
        const uint32_t a_lo = sa16 | a_lbo, b_lo = (sa16 + (A_BYTES >> 4)) | b_lbo;
        const uint32_t set = part ? 1u : 0u;                      // virtual chunk part > 0: a lo operand is involved
        const uint32_t accum = (used >> set) & 1u;
        used |= 1u << set;
        if (elect_one_sync()) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_bf16(tmem_acc + set * N_TILE, umma_desc64(a_lo + k4 * a_step, HI), umma_desc64(b_lo + k4 * b_step, HI), idesc,
                      k4 == 0 ? accum : 1u);
          umma_commit(&empty[slot]);
        }
        __syncwarp();
      }
      if (elect_one_sync()) umma_commit(&acc_full[as]);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const int tn = t % p.tiles_n, tm = (t / p.tiles_n) % p.tiles_m, b = t / (p.tiles_n * p.tiles_m);
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      mbar_wait(&acc_full[as], aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int m = tm * 128 + row;
      const bool mvalid = m < p.M;
      const float rb = (p.bias != nullptr && p.bias_rows && mvalid) ? p.bias[m] : 0.0f;
      const int64_t obase = (int64_t)b * p.strideC + (int64_t)m * p.ldc;
      const uint32_t taddr = tmem_base + as * BUF_COLS + ((uint32_t)(q * 32) << 16);
      // The 32-column groups are NOT unrolled and every data-dependent choice (output dtype, residual, bias kind, ragged
      // edge) is a warp-uniform branch around a small straight-line body: the first version unrolled everything into
      // 11.6 K SASS instructions and ran at the speed of the instruction cache (ncu: tensor pipe 4-6 % active).
      const bool full_n = (tn + 1) * N_TILE <= p.N;
      const bool col_bias = p.bias != nullptr && !p.bias_rows;
      if (p.epi == EPI_ROWSTAT) {
        // online (max, sum of exp) of this row over the tile's columns; nothing else is written: the scores never reach HBM
        float mx = -INFINITY, sm = 0.0f;
#pragma unroll 1
        for (int c0 = 0; c0 < N_TILE; c0 += 32) {
          uint32_t v[32];
          DSK_TMEM_LD_X32(v, taddr + c0);
          const int n = tn * N_TILE + c0;
          if (n >= p.N) continue;
          const int nv = p.N - n;
          float sv[32], g = -INFINITY;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            sv[e] = (full_n || e < nv) ? p.alpha * __uint_as_float(v[e]) : -INFINITY;
            g = fmaxf(g, sv[e]);
          }
          const float nm = fmaxf(mx, g);
          float acc = 0.0f;
#pragma unroll
          for (int e = 0; e < 32; ++e) acc += __expf(sv[e] - nm);
          sm = sm * __expf(mx - nm) + acc;
          mx = nm;
        }
        if (mvalid) p.rowstat[((int64_t)b * p.M + m) * p.tiles_n + tn] = make_float2(mx, sm);
      } else if (p.epi == EPI_SOFTMAX) {
        // 64 columns = one 128-byte line per row per iteration.  A thread holds one ROW of the tile; storing it directly
        // makes every warp-wide st.global.v4 touch 32 different lines.  Instead the warp stages its 32 x 128 B in shared
        // memory (XOR-swizzled 16-byte slots, conflict-free both ways) and writes it back transposed: 8 lanes per row,
        // 4 complete lines per store instruction.
        const float2 rf = mvalid ? p.rowfinal[(int64_t)b * p.M + m] : make_float2(0.0f, 0.0f);
        uint4* stg = stage_s[q];
        const int m_warp = tm * 128 + q * 32;
#pragma unroll 1
        for (int c0 = 0; c0 < N_TILE; c0 += 64) {
          const int n = tn * N_TILE + c0;
          if (n >= p.N) continue;                               // warp-uniform
          uint4 pk[8];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t v[32];
            DSK_TMEM_LD_X32(v, taddr + c0 + hh * 32);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t* oh = reinterpret_cast<uint32_t*>(&pk[hh * 4 + g]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float f0 = __expf(fmaf(p.alpha, __uint_as_float(v[g * 8 + 2 * e]), -rf.x)) * rf.y;
                const float f1 = __expf(fmaf(p.alpha, __uint_as_float(v[g * 8 + 2 * e + 1]), -rf.x)) * rf.y;
                oh[e] = pack_h2(f0, f1, p.f16);
              }
            }
          }
          if (n + 64 <= p.N) {                                  // warp-uniform
#pragma unroll
            for (int j = 0; j < 8; ++j) stg[lane * 8 + (j ^ (lane & 7))] = pk[j];
            __syncwarp();
            const int j = lane & 7;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int r = 4 * k + (lane >> 3);
              const uint4 val = stg[r * 8 + (j ^ (r & 7))];
              if (m_warp + r < p.M)
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t)b * p.strideC + (int64_t)(m_warp + r) * p.ldc + n +
                                          j * 8) = val;
            }
            __syncwarp();
          } else if (mvalid) {                                  // ragged right edge: per-row scalar stores
            const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(pk);
            __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + obase + n;
#pragma unroll 1
            for (int e = 0; e < 64 && n + e < p.N; ++e) orow[e] = src[e];
          }
        }
      } else {
      const int m_warp = tm * 128 + q * 32;
#pragma unroll 1
      for (int c0 = 0; c0 < N_TILE; c0 += 32) {
        uint32_t v[32];
        DSK_TMEM_LD_X32(v, taddr + c0);
        if (nsets > 1) {
          uint32_t u[32];
          DSK_TMEM_LD_X32(u, taddr + N_TILE + c0);
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(fmaf(__uint_as_float(u[e]), 1.0f / 2048.0f, __uint_as_float(v[e])));
        }
        const int n = tn * N_TILE + c0;
        if (n >= p.N) continue;                                 // warp-uniform
        float f[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) f[e] = fmaf(p.alpha, __uint_as_float(v[e]), rb);
        if (full_n || n + 32 <= p.N) {                          // warp-uniform
          if (col_bias) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
            for (int e4 = 0; e4 < 8; ++e4) {
              const float4 bb = __ldg(bp + e4);
              f[4 * e4] += bb.x; f[4 * e4 + 1] += bb.y; f[4 * e4 + 2] += bb.z; f[4 * e4 + 3] += bb.w;
            }
          }
          if (p.residual != nullptr && mvalid) {
            if (p.res_f32) {
              const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + obase + n);
#pragma unroll
              for (int e4 = 0; e4 < 8; ++e4) {
                const float4 r = rp[e4];
                f[4 * e4] += r.x; f[4 * e4 + 1] += r.y; f[4 * e4 + 2] += r.z; f[4 * e4 + 3] += r.w;
              }
            } else {
              const uint4* rp = reinterpret_cast<const uint4*>(p.residual + obase + n);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint4 rr = rp[g];
                const uint32_t* rw = reinterpret_cast<const uint32_t*>(&rr);
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float2 t = unpack_h2(rw[e], p.f16); f[g * 8 + 2 * e] += t.x; f[g * 8 + 2 * e + 1] += t.y; }
              }
            }
          }
          if (p.out_f32) {
            if (mvalid) {
              float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + obase + n);
#pragma unroll
              for (int e4 = 0; e4 < 8; ++e4) o[e4] = make_float4(f[4 * e4], f[4 * e4 + 1], f[4 * e4 + 2], f[4 * e4 + 3]);
            }
          } else {
            // staged, transposed store (see the softmax epilogue): 4 lanes per row write 64 contiguous bytes, 8 rows per
            // instruction; every lane takes part, row validity is checked where the row is written
            uint4* stg = stage_s[q];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 pk;
              uint32_t* oh = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
              for (int e = 0; e < 4; ++e) oh[e] = pack_h2(f[g * 8 + 2 * e], f[g * 8 + 2 * e + 1], p.f16);
              stg[lane * 4 + (g ^ ((lane >> 1) & 3))] = pk;
            }
            __syncwarp();
            const int j = lane & 3;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int r = 8 * k + (lane >> 2);
              const uint4 val = stg[r * 4 + (j ^ ((r >> 1) & 3))];
              if (m_warp + r < p.M)
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t)b * p.strideC + (int64_t)(m_warp + r) * p.ldc + n +
                                          j * 8) = val;
            }
            __syncwarp();
          }
        } else if (mvalid) {
          // ragged N edge: scalar, rolled (f is indexed dynamically -> local memory; rare and tiny)
#pragma unroll 1
          for (int e = 0; e < 32 && n + e < p.N; ++e) {
            float x = f[e];
            if (col_bias) x += __ldg(p.bias + n + e);
            if (p.residual != nullptr)
              x += p.res_f32 ? reinterpret_cast<const float*>(p.residual)[obase + n + e] : unpack_h1(p.residual[obase + n + e], p.f16);
            if (p.out_f32) reinterpret_cast<float*>(p.out)[obase + n + e] = x;
            else reinterpret_cast<uint16_t*>(p.out)[obase + n + e] = pack_h1(x, p.f16);
          }
        }
      }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

// fp32 scores -> bf16 probabilities, one block per row (row cached in registers: cols <= 256 * 32)
// fmt: DSK_BF16 | DSK_F16 (P rows of `cols`), DSK_SPLIT_F16 (P rows of 2 * cols: hi | lo)
__global__ void __launch_bounds__(256) softmax_rows_bf16_kernel(const float* __restrict__ S, uint16_t* __restrict__ P,
                                                                 int64_t rows, int cols, int fmt) {
  __shared__ float red[8];
  constexpr int MAXV = 8;   // float4 vectors per thread (cols <= 8192)
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float4* src = reinterpret_cast<const float4*>(S + r * cols);
    const int nv = cols >> 2;
    float4 v[MAXV];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = threadIdx.x + i * 256;
      if (idx < nv) {
        v[i] = src[idx];
        m = fmaxf(m, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    __syncthreads();
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = threadIdx.x + i * 256;
      if (idx < nv) {
        v[i].x = expf(v[i].x - m); v[i].y = expf(v[i].y - m); v[i].z = expf(v[i].z - m); v[i].w = expf(v[i].w - m);
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    sum = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w];
    __syncthreads();
    const float inv = 1.0f / sum;
    const bool split = fmt == DSK_SPLIT_F16;
    uint2* dst = reinterpret_cast<uint2*>(P + r * cols * (split ? 2 : 1));
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = threadIdx.x + i * 256;
      if (idx < nv) {
        const float p0 = v[i].x * inv, p1 = v[i].y * inv, p2 = v[i].z * inv, p3 = v[i].w * inv;
        uint2 o;
        o.x = pack_h2(p0, p1, fmt != DSK_BF16);
        o.y = pack_h2(p2, p3, fmt != DSK_BF16);
        dst[idx] = o;
        if (split) {
          const float2 h0 = unpack_h2(o.x, 1), h1 = unpack_h2(o.y, 1);
          uint2 l;
          l.x = pack_h2((p0 - h0.x) * 2048.0f, (p1 - h0.y) * 2048.0f, 1);      // lo stored times 2^11 (DSK_SPLIT_F16)
          l.y = pack_h2((p2 - h1.x) * 2048.0f, (p3 - h1.y) * 2048.0f, 1);
          dst[nv + idx] = l;
        }
      }
    }
  }
}

// dS = P * (dP - rowsum(dP * P)) for softmax rows: P bf16 probabilities, dP fp32 (GEMM output), dS bf16 (next GEMM operand)
__global__ void __launch_bounds__(256) softmax_bwd_rows_bf16_kernel(const __nv_bfloat16* __restrict__ P, const float* __restrict__ dP,
                                                                     __nv_bfloat16* __restrict__ dS, int64_t rows, int cols) {
  __shared__ float red[8];
  constexpr int MAXV = 8;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float4* g = reinterpret_cast<const float4*>(dP + r * cols);
    const uint2* pp = reinterpret_cast<const uint2*>(P + r * cols);
    const int nv = cols >> 2;
    float4 pv[MAXV], gv[MAXV];
    float dot = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = threadIdx.x + i * 256;
      if (idx < nv) {
        gv[i] = g[idx];
        const uint2 u = pp[idx];
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
        pv[i] = make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
        dot += (gv[i].x * pv[i].x + gv[i].y * pv[i].y) + (gv[i].z * pv[i].z + gv[i].w * pv[i].w);
      }
    }
    dot = warp_sum(dot);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    dot = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) dot += red[w];
    __syncthreads();
    uint2* dst = reinterpret_cast<uint2*>(dS + r * cols);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = threadIdx.x + i * 256;
      if (idx < nv) {
        __nv_bfloat162 a = __floats2bfloat162_rn(pv[i].x * (gv[i].x - dot), pv[i].y * (gv[i].y - dot));
        __nv_bfloat162 b = __floats2bfloat162_rn(pv[i].z * (gv[i].z - dot), pv[i].w * (gv[i].w - dot));
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&a);
        o.y = *reinterpret_cast<uint32_t*>(&b);
        dst[idx] = o;
      }
    }
  }
}

// (max, sum) partials of the score tiles of a row -> (row max, 1 / row sum)
__global__ void __launch_bounds__(256) rowstat_combine_kernel(const float2* __restrict__ part, float2* __restrict__ fin, int64_t rows, int tiles) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float2* p = part + r * tiles;
  float mx = -INFINITY;
  for (int t = 0; t < tiles; ++t) mx = fmaxf(mx, p[t].x);
  float sm = 0.0f;
  for (int t = 0; t < tiles; ++t) sm += p[t].y * __expf(p[t].x - mx);
  fin[r] = make_float2(mx, 1.0f / sm);
}

template <int N_TILE>
static int launch_gemm_tc(const CUtensorMap& ta, const CUtensorMap& tb, const GemmTcParams& p, cudaStream_t st) {
  const size_t smem = (size_t)GT_STAGES * (128 * 128 + N_TILE * 128) + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<N_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return DSK_ERR_CUDA; }
    configured = true;
  }
  const int grid = p.total_tiles < DSK_NUM_SMS ? p.total_tiles : DSK_NUM_SMS;
  DSK_LAUNCH((gemm_tc_kernel<N_TILE>), grid, GT_THREADS, smem, st, ta, tb, p);
  return DSK_OK;
}

}  // namespace dsk

using namespace dsk;

// dtype: DSK_BF16 | DSK_F16: plain 16-bit operands.  DSK_SPLIT_F16: A and B are split-fp16 (hi | lo along their contiguous
// coordinate, lo halves at element offsets a_lo / b_lo), 3 MMAs per k-step; b_lo < 0: B is plain fp16, 2 MMAs per k-step.
static int gemm_tc_run(const void* A, const void* Bm, void* C, const float* bias, int bias_rows, const void* residual, int M, int N, int K,
                       int64_t lda, int64_t ldb, int64_t ldc, int64_t strideA, int64_t strideB, int64_t strideC, int batch, float alpha,
                       int out_f32, int transA, int transB, int epi, float2* rowstat, const float2* rowfinal, int force_tile,
                       void* stream, int dtype = DSK_BF16, int a_lo = 0, int b_lo = 0, int res_f32 = 0) {
  DSK_REQUIRE(is_h16(dtype) || dtype == DSK_SPLIT_F16, "dsk_gemm_tc: bad dtype %d", dtype);
  const bool split = dtype == DSK_SPLIT_F16;
  DSK_REQUIRE(!split || (K % 64 == 0 && a_lo > 0 && a_lo % 8 == 0 && (b_lo < 0 || (b_lo > 0 && b_lo % 8 == 0))),
              "dsk_gemm_tc: split operands need K %% 64 == 0 and lo offsets that are multiples of 8 elements");
  DSK_REQUIRE(A && Bm && (C || epi == EPI_ROWSTAT), "dsk_gemm_bf16_tc: null pointer");
  DSK_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0, "dsk_gemm_bf16_tc: bad shape");
  DSK_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0 && strideA % 8 == 0 && strideB % 8 == 0,
              "dsk_gemm_bf16_tc: K, lda, ldb, ldc and batch strides must be multiples of 8 elements (16-byte TMA alignment)");
  DSK_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)Bm & 15) == 0 && ((uintptr_t)C & 15) == 0, "dsk_gemm_bf16_tc: 16-byte alignment");
  EncodeTiledFn encode = get_encode();
  DSK_REQUIRE(encode != nullptr, "dsk_gemm_bf16_tc: cuTensorMapEncodeTiled is unavailable");
  DSK_REQUIRE(!split || epi == EPI_NORMAL, "dsk_gemm_tc: split operands take the plain epilogue only");
  const int n_tile = force_tile ? force_tile : ((N > 128 && !split) ? 256 : (N > 64 ? 128 : 64));
  CUtensorMap ta, tb;
  // K-major operand [rows][K]: box 64 (k) x box_rows;  MN-major operand [K][rows]: box 64 (rows) x 64 (k)
  auto make = [&](CUtensorMap* tm, const void* base, int rows, int64_t ld, int64_t stride, int box_rows, int trans, int lo_off) -> bool {
    const bool shared = batch == 1 || stride == 0;
    const int inner = (trans ? rows : K) + lo_off, outer = trans ? K : rows;
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)(shared ? 1 : batch)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(shared ? (int64_t)outer * ld : stride) * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)(trans ? 64 : box_rows), 1};
    cuuint32_t es[3] = {1, 1, 1};
    return encode(tm, tmap_h16(dtype == DSK_BF16 ? DSK_BF16 : DSK_F16), 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  DSK_REQUIRE(make(&ta, A, M, lda, strideA, 128, transA, split ? a_lo : 0), "dsk_gemm_bf16_tc: tensor map for A failed");
  DSK_REQUIRE(make(&tb, Bm, N, ldb, strideB, n_tile, transB, (split && b_lo > 0) ? b_lo : 0), "dsk_gemm_bf16_tc: tensor map for B failed");
  GemmTcParams p;
  p.M = M; p.N = N; p.K = K; p.batch = batch;
  p.a_bmul = (batch > 1 && strideA != 0) ? 1 : 0;
  p.b_bmul = (batch > 1 && strideB != 0) ? 1 : 0;
  p.tiles_m = (M + 127) / 128; p.tiles_n = (N + n_tile - 1) / n_tile;
  p.total_tiles = p.tiles_m * p.tiles_n * batch;
  p.alpha = alpha; p.bias = bias; p.bias_rows = bias_rows;
  p.residual = (const uint16_t*)residual; p.out = C; p.out_f32 = out_f32; p.ldc = ldc; p.strideC = strideC;
  p.f16 = dtype == DSK_BF16 ? 0 : 1; p.res_f32 = res_f32;
  p.vparts = split ? (b_lo > 0 ? 3 : 2) : 1; p.a_lo = a_lo; p.b_lo = b_lo > 0 ? b_lo : 0;
  p.transA = transA ? 1 : 0; p.transB = transB ? 1 : 0;
  p.epi = epi; p.rowstat = rowstat; p.rowfinal = rowfinal;
  cudaStream_t st = as_stream(stream);
  if (n_tile == 64) return launch_gemm_tc<64>(ta, tb, p, st);
  if (n_tile == 128) return launch_gemm_tc<128>(ta, tb, p, st);
  return launch_gemm_tc<256>(ta, tb, p, st);
}

extern "C" int dsk_gemm_bf16_tc(const void* A, const void* Bm, void* C, const float* bias, int bias_rows, const void* residual,
                                int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc, int64_t strideA, int64_t strideB,
                                int64_t strideC, int batch, float alpha, int out_f32, int transA, int transB, void* stream) {
  return gemm_tc_run(A, Bm, C, bias, bias_rows, residual, M, N, K, lda, ldb, ldc, strideA, strideB, strideC, batch, alpha, out_f32, transA,
                     transB, EPI_NORMAL, nullptr, nullptr, 0, stream);
}

extern "C" int dsk_gemm_tc(const void* A, const void* Bm, void* C, const float* bias, int bias_rows, const void* residual, int res_f32,
                           int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc, int64_t strideA, int64_t strideB, int64_t strideC,
                           int batch, float alpha, int out_f32, int transA, int transB, int dtype, int a_lo, int b_lo, void* stream) {
  return gemm_tc_run(A, Bm, C, bias, bias_rows, residual, M, N, K, lda, ldb, ldc, strideA, strideB, strideC, batch, alpha, out_f32, transA,
                     transB, EPI_NORMAL, nullptr, nullptr, 0, stream, dtype, a_lo, b_lo, res_f32);
}

extern "C" int64_t dsk_attn_softmax_ws_bytes(int batch, int L) {
  if (batch <= 0 || L <= 0) return 0;
  const int tiles = (L + 255) / 256;
  return (int64_t)batch * L * (tiles + 1) * (int64_t)sizeof(float2);
}

// P[b] = softmax(alpha Q[b] K[b]^T) as bf16, without the scores ever reaching HBM: pass 1 = QK^T with a row-statistics
// epilogue (per 256-column tile), a tiny combine, pass 2 = QK^T again with the exp / normalise / bf16 epilogue.
extern "C" int dsk_attn_softmax_qk_h16(const void* Q, const void* Kmat, void* P, void* ws, int L, int C, int64_t ldq, int64_t ldk,
                                       int64_t strideQ, int64_t strideK, int batch, float alpha, int dtype, void* stream);
extern "C" int dsk_attn_softmax_qk(const void* Q, const void* Kmat, void* P, void* ws, int L, int C, int64_t ldq, int64_t ldk,
                                   int64_t strideQ, int64_t strideK, int batch, float alpha, void* stream) {
  return dsk_attn_softmax_qk_h16(Q, Kmat, P, ws, L, C, ldq, ldk, strideQ, strideK, batch, alpha, DSK_BF16, stream);
}
extern "C" int dsk_attn_softmax_qk_h16(const void* Q, const void* Kmat, void* P, void* ws, int L, int C, int64_t ldq, int64_t ldk,
                                       int64_t strideQ, int64_t strideK, int batch, float alpha, int dtype, void* stream) {
  DSK_REQUIRE(is_h16(dtype), "dsk_attn_softmax_qk: bad dtype %d", dtype);
  DSK_REQUIRE(Q && Kmat && P && ws && L > 0 && C > 0 && batch > 0 && L % 8 == 0, "dsk_attn_softmax_qk: bad arguments (L %% 8 == 0)");
  const int tiles = (L + 255) / 256;
  float2* part = reinterpret_cast<float2*>(ws);
  float2* fin = part + (int64_t)batch * L * tiles;
  int rc = gemm_tc_run(Q, Kmat, P, nullptr, 0, nullptr, L, L, C, ldq, ldk, L, strideQ, strideK, (int64_t)L * L, batch, alpha, 0, 0, 0,
                       EPI_ROWSTAT, part, nullptr, 256, stream, dtype);
  if (rc != DSK_OK) return rc;
  const int64_t rows = (int64_t)batch * L;
  DSK_LAUNCH(rowstat_combine_kernel, (int)((rows + 255) / 256), 256, 0, as_stream(stream), part, fin, rows, tiles);
  return gemm_tc_run(Q, Kmat, P, nullptr, 0, nullptr, L, L, C, ldq, ldk, L, strideQ, strideK, (int64_t)L * L, batch, alpha, 0, 0, 0,
                     EPI_SOFTMAX, nullptr, fin, 256, stream, dtype);
}

extern "C" int dsk_softmax_rows_h16(const float* S, void* P, int64_t rows, int cols, int dtype, void* stream) {
  DSK_REQUIRE(S && P && rows > 0 && cols > 0, "dsk_softmax_rows_h16: bad arguments");
  DSK_REQUIRE(is_h16(dtype) || dtype == DSK_SPLIT_F16, "dsk_softmax_rows_h16: bad dtype %d", dtype);
  DSK_REQUIRE(cols % 4 == 0 && cols <= 8192, "dsk_softmax_rows_h16: cols=%d must be a multiple of 4 and <= 8192", cols);
  int64_t grid = rows < (int64_t)DSK_NUM_SMS * 16 ? rows : (int64_t)DSK_NUM_SMS * 16;
  DSK_LAUNCH(softmax_rows_bf16_kernel, (int)grid, 256, 0, as_stream(stream), S, (uint16_t*)P, rows, cols, dtype);
  return DSK_OK;
}
extern "C" int dsk_softmax_rows_bf16(const float* S, void* P, int64_t rows, int cols, void* stream) {
  return dsk_softmax_rows_h16(S, P, rows, cols, DSK_BF16, stream);
}

extern "C" int dsk_softmax_bwd_rows_bf16(const void* P, const float* dP, void* dS, int64_t rows, int cols, void* stream) {
  DSK_REQUIRE(P && dP && dS && rows > 0 && cols > 0, "dsk_softmax_bwd_rows_bf16: bad arguments");
  DSK_REQUIRE(cols % 4 == 0 && cols <= 8192, "dsk_softmax_bwd_rows_bf16: cols=%d must be a multiple of 4 and <= 8192", cols);
  int64_t grid = rows < (int64_t)DSK_NUM_SMS * 16 ? rows : (int64_t)DSK_NUM_SMS * 16;
  DSK_LAUNCH(softmax_bwd_rows_bf16_kernel, (int)grid, 256, 0, as_stream(stream), (const __nv_bfloat16*)P, dP, (__nv_bfloat16*)dS, rows, cols);
  return DSK_OK;
}
