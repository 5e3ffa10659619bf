// gemm_ffma.cu -- K1 (fp32-parity mode): implicit-GEMM convolution and batched GEMM on the
// CUDA cores (FFMA, fp32 accumulate).  This is the path that meets the 1e-5 fp32 parity bar
// of the north star; the bf16 throughput path is the tcgen05 kernel in conv_tc.cu.
//
// Reference call sites: torch.nn.Conv2d/Conv3d(padding='same') (nets/commonlayers.py:777-833,
// 53-58, 123-128; nets/punetg.py:203-214; nets/adm.py:268-284), torch.nn.Linear and the
// projections / QK^T / PV products inside nn.MultiheadAttention (nets/attention.py:42-44).
//
// Tiling: 128 (rows = output pixels) x 64 (cols = output channels) x 16 (k) per CTA of 256
// threads, 8x4 accumulators per thread, register prefetch of the next k-slab.  The im2col gather
// (zero padding, optional nearest x2 upsample of the input) happens in the A loader.
#include "common.cuh"

namespace dsk {

constexpr int BM = 128, BN = 64, BK = 16, GT = 256;

// ------------------------------------------------------------------------------------------------
template <typename TI, typename TO>
struct ConvProb {
  static constexpr bool kTransA = false;
  static constexpr bool kTransB = false;
  __device__ __forceinline__ int k_begin() const { return 0; }
  __device__ __forceinline__ int k_end() const { return K; }
  __device__ __forceinline__ bool tile_is_empty(int, int) const { return false; }
  const TI* in;
  const float* w;         // [taps*Cin][Cout]
  const float* bias;      // [Cout] or null
  const float* chan_bias; // [B][Cout] or null
  const TO* residual;     // channels-last like out, or null
  TO* out;
  float* out_nchw;        // fp32 NC(D)HW output (if set, `out` is unused)
  int B, D, H, W, Cin, Cout, ks, ndim, up2;
  int Di, Hi, Wi;         // input spatial size
  int M, N, K;
  float alpha;
  int act;
  int circ;               // circular padding: wrap the gather coordinates (CircularConv, commonlayers.py:918-1032)

  struct RowCtx {
    int b, d, h, w;
    bool valid;
  };
  __device__ __forceinline__ RowCtx row_ctx(int m) const {
    RowCtx c;
    c.valid = m < M;
    int p = c.valid ? m : 0;
    c.w = p % W; p /= W;
    c.h = p % H; p /= H;
    c.d = p % D;
    c.b = p / D;
    return c;
  }
  // 8 consecutive k (k % 8 == 0) of im2col row `c`
  __device__ __forceinline__ void load_a(const RowCtx& c, int k, float* o) const {
    if ((Cin & 7) == 0) {
      const int tap = k / Cin, ci = k - tap * Cin;
      const TI* p = tap_ptr(c, tap);
      if (p == nullptr || k >= K) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.0f;
      } else {
        ld8(p + ci, o);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kk = k + j;
        float v = 0.0f;
        if (kk < K) {
          const int tap = kk / Cin, ci = kk - tap * Cin;
          const TI* p = tap_ptr(c, tap);
          if (p != nullptr) v = to_f32<TI>(p[ci]);
        }
        o[j] = v;
      }
    }
  }
  __device__ __forceinline__ const TI* tap_ptr(const RowCtx& c, int tap) const {
    if (!c.valid) return nullptr;
    const int r = ks >> 1;
    int kw = tap % ks, t2 = tap / ks;
    int kh = t2 % ks, kd = t2 / ks;
    int zw = c.w + kw - r, zh = c.h + kh - r, zd = ndim == 3 ? c.d + kd - r : 0;
    if (circ) {
      zw = zw < 0 ? zw + W : (zw >= W ? zw - W : zw);
      zh = zh < 0 ? zh + H : (zh >= H ? zh - H : zh);
      zd = zd < 0 ? zd + D : (zd >= D ? zd - D : zd);
    } else if ((unsigned)zw >= (unsigned)W || (unsigned)zh >= (unsigned)H || (unsigned)zd >= (unsigned)D) return nullptr;
    if (up2) {  // conv over F.interpolate(x, 2, 'nearest'): source voxel = floor(coord / 2)
      zw >>= 1; zh >>= 1;
      if (ndim == 3) zd >>= 1;
    }
    return in + ((((int64_t)c.b * Di + zd) * Hi + zh) * Wi + zw) * Cin;
  }
  static __device__ __forceinline__ void ld8(const float* p, float* o) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
  static __device__ __forceinline__ void ld8(const __nv_bfloat16* p, float* o) {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[2 * j] = __low2float(h[j]); o[2 * j + 1] = __high2float(h[j]); }
  }
  static __device__ __forceinline__ void ld8(const __half* p, float* o) {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[2 * j] = __low2float(h[j]); o[2 * j + 1] = __high2float(h[j]); }
  }
  // 4 consecutive n of weight row k
  __device__ __forceinline__ void load_b(int k, int n, float* o) const {
    if (k < K && (Cout & 3) == 0 && n + 3 < N) {
      float4 v = *reinterpret_cast<const float4*>(w + (int64_t)k * Cout + n);
      o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = (k < K && n + j < N) ? w[(int64_t)k * Cout + n + j] : 0.0f;
    }
  }
  __device__ __forceinline__ void store(int m, int n, const float* acc) const {
    if (m >= M) return;
    RowCtx c = row_ctx(m);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n + j;
      if (nn >= N) break;
      float v = acc[j];
      if (bias != nullptr) v += bias[nn];
      if (chan_bias != nullptr) v += chan_bias[(int64_t)c.b * Cout + nn];
      if (residual != nullptr) v += to_f32<TO>(residual[(int64_t)m * Cout + nn]);
      if (out_nchw != nullptr) {
        const int64_t S = (int64_t)D * H * W;
        out_nchw[((int64_t)c.b * Cout + nn) * S + (m - (int64_t)c.b * S)] = v;
      } else {
        out[(int64_t)m * Cout + nn] = from_f32<TO>(v);
      }
    }
  }
  __device__ __forceinline__ void select_batch(int) {}
};

// ------------------------------------------------------------------------------------------------
template <bool TRANSA, bool TRANSB>
struct GemmProb {
  static constexpr bool kTransA = TRANSA;
  static constexpr bool kTransB = TRANSB;
  const float* A;
  const float* Bm;
  float* C;
  const float* bias;
  int M, N, K, lda, ldb, ldc;
  int64_t sA, sB, sC;
  float alpha;
  int act;
  float beta;   // C = act(alpha * op(A) op(B) + bias) + beta * C
  __device__ __forceinline__ int k_begin() const { return 0; }
  __device__ __forceinline__ int k_end() const { return K; }
  __device__ __forceinline__ bool tile_is_empty(int m0, int n0) const { return m0 >= M || n0 >= N; }
  // TRANSA: A is stored [K][M] (lda >= M); 8 consecutive m of row k
  __device__ __forceinline__ void load_a_t(int k, int m, float* o) const {
    const float* p = A + (int64_t)k * lda + m;
    if (k < K && m + 7 < M && ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
      float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
      o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (k < K && m + j < M) ? p[j] : 0.0f;
    }
  }
  struct RowCtx { int m; bool valid; };
  __device__ __forceinline__ RowCtx row_ctx(int m) const { return RowCtx{m, m < M}; }
  __device__ __forceinline__ void load_a(const RowCtx& c, int k, float* o) const {
    const float* p = A + (int64_t)c.m * lda + k;
    if (c.valid && k + 7 < K && ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
      float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
      o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (c.valid && k + j < K) ? p[j] : 0.0f;
    }
  }
  // TRANSB: B is [N][K]; returns 4 consecutive k of row n.  else: B is [K][N]; 4 consecutive n of row k.
  __device__ __forceinline__ void load_b(int k, int n, float* o) const {
    if (TRANSB) {
      const float* p = Bm + (int64_t)n * ldb + k;
      if (n < N && k + 3 < K && ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        float4 v = *reinterpret_cast<const float4*>(p);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = (n < N && k + j < K) ? p[j] : 0.0f;
      }
    } else {
      const float* p = Bm + (int64_t)k * ldb + n;
      if (k < K && n + 3 < N && ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        float4 v = *reinterpret_cast<const float4*>(p);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = (k < K && n + j < N) ? p[j] : 0.0f;
      }
    }
  }
  __device__ __forceinline__ void store(int m, int n, const float* acc) const {
    if (m >= M) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n + j >= N) break;
      float v = alpha * acc[j];
      if (bias != nullptr) v += bias[n + j];
      if (act == 1) v = silu_f(v);
      else if (act == 2) v = fmaxf(v, 0.0f);
      if (beta != 0.0f) v += beta * C[(int64_t)m * ldc + n + j];
      C[(int64_t)m * ldc + n + j] = v;
    }
  }
  __device__ __forceinline__ void select_batch(int z) {
    A += z * sA; Bm += z * sB; C += z * sC;
  }
};

// Grouped GEMM: blockIdx.z selects one of many small independent problems described by a device table (the per-block
// time MLPs: one launch per layer for every ResNet block).  `Z` (optional) receives the pre-activation.
struct GroupedGemmDesc {
  const float* A;
  const float* Bm;
  float* C;
  const float* bias;
  float* Z;
  int M, N, K, lda, ldb, ldc;
};
template <bool TRANSA, bool TRANSB>
struct GroupedGemmProb : GemmProb<TRANSA, TRANSB> {
  using Base = GemmProb<TRANSA, TRANSB>;
  const GroupedGemmDesc* table;
  float* Z;
  __device__ __forceinline__ void select_batch(int z) {
    const GroupedGemmDesc d = table[z];
    this->A = d.A; this->Bm = d.Bm; this->C = d.C; this->bias = d.bias; Z = d.Z;
    this->M = d.M; this->N = d.N; this->K = d.K; this->lda = d.lda; this->ldb = d.ldb; this->ldc = d.ldc;
  }
  __device__ __forceinline__ void store(int m, int n, const float* acc) const {
    if (m >= this->M) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n + j >= this->N) break;
      float v = this->alpha * acc[j];
      if (this->bias != nullptr) v += this->bias[n + j];
      if (Z != nullptr) Z[(int64_t)m * this->ldc + n + j] = v;
      if (this->act == 1) v = silu_f(v);
      else if (this->act == 2) v = fmaxf(v, 0.0f);
      this->C[(int64_t)m * this->ldc + n + j] = v;
    }
  }
};

// For transB GEMMs (B stored [N][K]) the B tile is loaded k-contiguous (4 consecutive k of one n)
// so that global reads stay coalesced; the transposing smem store is bank-conflict free.
// kTransA problems (A stored [K][M]: weight-gradient and A^T B products) load 8 consecutive m of one k.
// blockIdx.z selects a batch entry or, for the split-K weight gradient, a slice [k_begin, k_end) of the reduction.
template <class Prob>
__global__ void __launch_bounds__(GT) ffma_gemm_kernel(Prob p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  p.select_batch(blockIdx.z);
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (p.tile_is_empty(m0, n0)) return;     // grouped launches size the grid for the largest problem (block-uniform exit)
  // A loader: thread -> (row, 8 consecutive k); kTransA: thread -> (k, 8 consecutive rows)
  const int a_row = Prob::kTransA ? (tid & 15) * 8 : tid & (BM - 1);
  const int a_k = Prob::kTransA ? tid >> 4 : (tid >> 7) * 8;
  const typename Prob::RowCtx actx = p.row_ctx(m0 + a_row);
  // B loader: thread -> (k, 4 consecutive n), or for kTransB (n, 4 consecutive k)
  const int b_k = Prob::kTransB ? (tid & 3) * 4 : tid >> 4;
  const int b_n = Prob::kTransB ? tid >> 2 : (tid & 15) * 4;
  const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 thread grid, 8 rows x 4 cols each

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  float ra[8], rb[4];
  const int kbeg = p.k_begin(), kend = p.k_end();
  if constexpr (Prob::kTransA) p.load_a_t(kbeg + a_k, m0 + a_row, ra);
  else p.load_a(actx, kbeg + a_k, ra);
  p.load_b(kbeg + b_k, n0 + b_n, rb);
  const int nk = (kend - kbeg + BK - 1) / BK;
  for (int kt = 0; kt < nk; ++kt) {
    if constexpr (Prob::kTransA) {
      *reinterpret_cast<float4*>(&As[a_k][a_row]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
      *reinterpret_cast<float4*>(&As[a_k][a_row + 4]) = make_float4(ra[4], ra[5], ra[6], ra[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) As[a_k + j][a_row] = ra[j];
    }
    if (Prob::kTransB) {
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[b_k + j][b_n] = rb[j];
    } else {
      *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    }
    __syncthreads();
    if (kt + 1 < nk) {
      const int kn = kbeg + (kt + 1) * BK;
      if constexpr (Prob::kTransA) p.load_a_t(kn + a_k, m0 + a_row, ra);
      else p.load_a(actx, kn + a_k, ra);
      p.load_b(kn + b_k, n0 + b_n, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) p.store(m0 + ty * 8 + i, n0 + tx * 4, acc[i]);
}

// ------------------------------------------------------------------------------------------------
// Weight gradient of conv_same as a split-K "A^T B" product (backward of torch.nn.Conv2d/3d, reached by
// loss.backward() in KarrasModule.training_step, karrasmodule.py:1146-1155):
//   dW[tap*Cin + ci][co] = sum_pix Xcol[pix][tap*Cin + ci] * dY[pix][co]
// rows m = (tap, ci), columns n = co, reduction k = output pixel.  Both operands are read k-major exactly as they lie in
// HBM (channels-last), the im2col gather (zero padding, optional nearest x2 upsample of x) happens in the A loader.
// Each blockIdx.z reduces one slice of the pixels into ws[z][m][n]; wgrad_reduce_kernel sums the slices in a fixed
// order (deterministic, no atomics) and writes the reference layout [Cout][Cin][taps].
template <typename TI, typename TG>
struct WgradProb {
  static constexpr bool kTransA = true;
  static constexpr bool kTransB = false;
  const TI* x;
  const TG* dy;
  float* ws;
  int B, D, H, W, Cin, Cout, ks, ndim, up2;
  int Di, Hi, Wi;
  int M, N, K;           // taps*Cin, Cout, pixels
  int kper;              // pixels per split (multiple of BK)
  int kb, ke;
  int circ;
  struct RowCtx { int dummy; };
  __device__ __forceinline__ RowCtx row_ctx(int) const { return RowCtx{0}; }
  __device__ __forceinline__ int k_begin() const { return kb; }
  __device__ __forceinline__ int k_end() const { return ke; }
  __device__ __forceinline__ bool tile_is_empty(int, int) const { return false; }
  __device__ __forceinline__ void select_batch(int z) {
    kb = z * kper;
    ke = min(K, kb + kper);
    ws += (int64_t)z * M * N;
  }
  __device__ __forceinline__ const TI* tap_ptr(int pb, int pd, int ph, int pw, int tap) const {
    const int r = ks >> 1;
    int kw = tap % ks, t2 = tap / ks;
    int kh = t2 % ks, kd = t2 / ks;
    int zw = pw + kw - r, zh = ph + kh - r, zd = ndim == 3 ? pd + kd - r : 0;
    if (circ) {
      zw = zw < 0 ? zw + W : (zw >= W ? zw - W : zw);
      zh = zh < 0 ? zh + H : (zh >= H ? zh - H : zh);
      zd = zd < 0 ? zd + D : (zd >= D ? zd - D : zd);
    } else if ((unsigned)zw >= (unsigned)W || (unsigned)zh >= (unsigned)H || (unsigned)zd >= (unsigned)D) return nullptr;
    if (up2) {
      zw >>= 1; zh >>= 1;
      if (ndim == 3) zd >>= 1;
    }
    return x + ((((int64_t)pb * Di + zd) * Hi + zh) * Wi + zw) * Cin;
  }
  // 8 consecutive rows m (m % 8 == 0) of im2col column `k` (= output pixel)
  __device__ __forceinline__ void load_a_t(int k, int m, float* o) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.0f;
    if (k >= ke || m >= M) return;
    int p = k;
    const int pw = p % W; p /= W;
    const int ph = p % H; p /= H;
    const int pd = p % D;
    const int pb = p / D;
    if ((Cin & 7) == 0) {
      const int tap = m / Cin, ci = m - tap * Cin;
      const TI* q = tap_ptr(pb, pd, ph, pw, tap);
      if (q != nullptr) ConvProb<TI, float>::ld8(q + ci, o);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int mm = m + j;
        if (mm < M) {
          const int tap = mm / Cin, ci = mm - tap * Cin;
          const TI* q = tap_ptr(pb, pd, ph, pw, tap);
          if (q != nullptr) o[j] = to_f32<TI>(q[ci]);
        }
      }
    }
  }
  __device__ __forceinline__ void load_b(int k, int n, float* o) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = (k < ke && n + j < N) ? to_f32<TG>(dy[(int64_t)k * Cout + n + j]) : 0.0f;
  }
  __device__ __forceinline__ void store(int m, int n, const float* acc) const {
    if (m >= M) return;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n + j < N) ws[(int64_t)m * N + n + j] = acc[j];
  }
};

// dw[co][ci][tap] (reference layout) (+)= sum_z ws[z][tap*Cin + ci][co]
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int Cout, int Cin,
                                                            int taps, int nsplit, int accumulate) {
  const int64_t MN = (int64_t)taps * Cin * Cout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < MN; i += (int64_t)gridDim.x * blockDim.x) {
    // i indexes ws row-major [m][co] so that the reads are coalesced
    const int co = (int)(i % Cout);
    const int m = (int)(i / Cout);
    const int tap = m / Cin, ci = m - tap * Cin;
    float s = 0.0f;
    int z = 0;
    for (; z + 8 <= nsplit; z += 8) {       // 8 loads in flight, summed in the same (sequential) order
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = ws[(int64_t)(z + k) * MN + i];
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[k];
    }
    for (; z < nsplit; ++z) s += ws[(int64_t)z * MN + i];
    float* o = dw + ((int64_t)co * Cin + ci) * taps + tap;
    *o = accumulate ? *o + s : s;
  }
}

// bf16 layout: [tap][rows][Cin] with rows = max(Cout, pad_rows) (zero rows beyond Cout: the N_TILE = 16 convout path)
// dgrad = 1: the packed weights of the DATA-GRADIENT convolution dX = conv_same(dY, W') with W'[tap'][co -> in][ci -> out],
// tap' = taps-1-tap (all axes flipped): the same two layouts with the roles of Cin and Cout exchanged.
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ w, float* __restrict__ o32,
                                                           uint16_t* __restrict__ o16, int Cout, int Cin, int taps, int rows,
                                                           int dgrad, int fmt) {
  const int64_t total = (int64_t)Cout * Cin * taps;
  if (o16 != nullptr && rows > Cout) {
    const int64_t padded = (int64_t)taps * rows * Cin * (fmt == DSK_SPLIT_F16 ? 2 : 1);
    const int rl = Cin * (fmt == DSK_SPLIT_F16 ? 2 : 1);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += (int64_t)gridDim.x * blockDim.x)
      if ((i / rl) % rows >= Cout) o16[i] = 0;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // reference layout: w[co][ci][tap]
    int tap = (int)(i % taps);
    int64_t r = i / taps;
    int ci = (int)(r % Cin);
    int co = (int)(r / Cin);
    float v = w[i];
    if (dgrad) {
      const int tf = taps - 1 - tap;
      if (o32 != nullptr) o32[((int64_t)tf * Cout + co) * Cin + ci] = v;                      // [tap'][in = co][out = ci]
      if (o16 != nullptr) put16(o16, (int64_t)tf * Cin + ci, Cout, co, v, fmt);             // [tap'][out = ci][in = co]
      continue;
    }
    if (o32 != nullptr) o32[((int64_t)tap * Cin + ci) * Cout + co] = v;          // [tap][ci][co]
    if (o16 != nullptr) put16(o16, (int64_t)tap * rows + co, Cin, ci, v, fmt);   // [tap][co][ci]
  }
}

// bf16 packing through a shared-memory transposition (the element-per-thread kernel above writes 2-byte elements with a
// stride of Cout*Cin: 12 us per 256x256x27 weight, 65 launches per training iteration).  Block = one fixed channel of the
// outer operand side x 64 channels of the inner (contiguous-in-output) side x all taps:
//   dgrad = 0: out[tap][co][ci]  -- block (co, 64 ci): reads 64*taps CONTIGUOUS floats, writes `taps` rows of 128 bytes
//   dgrad = 1: out[taps-1-tap][ci][co] -- block (ci, 64 co): reads 64 runs of `taps` floats, writes `taps` rows of 128 bytes
__global__ void __launch_bounds__(256) pack_weight_bf16_tiled_kernel(const float* __restrict__ w, uint16_t* __restrict__ o16, int Cout,
                                                                      int Cin, int taps, int dgrad, int fmt) {
  __shared__ float tile[27][65];
  const int inner = dgrad ? Cout : Cin;                    // contiguous dimension of the output rows
  const int chunks = (inner + 63) / 64;
  const int fixed = blockIdx.x / chunks, c0 = (blockIdx.x % chunks) * 64;
  const int n = min(64, inner - c0);
  for (int idx = threadIdx.x; idx < n * taps; idx += blockDim.x) {
    const int cl = idx / taps, tap = idx - cl * taps;
    const int co = dgrad ? c0 + cl : fixed, ci = dgrad ? fixed : c0 + cl;
    tile[tap][cl] = w[((int64_t)co * Cin + ci) * taps + tap];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < n * taps; idx += blockDim.x) {
    const int tap = idx / n, cl = idx - tap * n;
    const int64_t row = dgrad ? (int64_t)(taps - 1 - tap) * Cin + fixed : (int64_t)tap * Cout + fixed;
    put16(o16, row, inner, c0 + cl, tile[tap][cl], fmt);
  }
}

// The same tiled packing for MANY weights in one launch (a training iteration re-packs every convolution's forward and
// data-gradient layouts after the optimizer step: 62 launches of ~8 us each on C4, 4-6 % of an iteration, latency-bound).
// jobs[j].block0 = first block of job j (ascending); a block finds its job by binary search.
struct PackJob {
  const float* w;
  uint16_t* out;
  int Cout, Cin, taps, dgrad, fmt, block0;
};
__global__ void __launch_bounds__(256) pack_weight_multi_kernel(const PackJob* __restrict__ jobs, int njobs) {
  __shared__ float tile[27][65];
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {                                   // last job with block0 <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].block0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const PackJob jb = jobs[lo];
  const int blk = blockIdx.x - jb.block0;
  const int Cout = jb.Cout, Cin = jb.Cin, taps = jb.taps, dgrad = jb.dgrad;
  const int inner = dgrad ? Cout : Cin;
  const int chunks = (inner + 63) / 64;
  const int fixed = blk / chunks, c0 = (blk % chunks) * 64;
  const int n = min(64, inner - c0);
  for (int idx = threadIdx.x; idx < n * taps; idx += blockDim.x) {
    const int cl = idx / taps, tap = idx - cl * taps;
    const int co = dgrad ? c0 + cl : fixed, ci = dgrad ? fixed : c0 + cl;
    tile[tap][cl] = jb.w[((int64_t)co * Cin + ci) * taps + tap];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < n * taps; idx += blockDim.x) {
    const int tap = idx / n, cl = idx - tap * n;
    const int64_t row = dgrad ? (int64_t)(taps - 1 - tap) * Cin + fixed : (int64_t)tap * Cout + fixed;
    put16(jb.out, row, inner, c0 + cl, tile[tap][cl], jb.fmt);
  }
}

static bool pack_tiled_ok(int Cout, int Cin, int taps, int rows, int dtype) {
  return dtype != DSK_F32 && rows == Cout && taps <= 27 && (int64_t)Cout * Cin * taps >= 16384;
}

template <typename TI, typename TO>
static int launch_conv(const dsk_conv_desc* d, const void* in, const void* w, const float* bias, const float* chan_bias,
                       const void* residual, void* out, cudaStream_t st) {
  ConvProb<TI, TO> p;
  p.in = (const TI*)in; p.w = (const float*)w; p.bias = bias; p.chan_bias = chan_bias;
  p.residual = (const TO*)residual;
  p.out = d->out_nchw_f32 ? nullptr : (TO*)out;
  p.out_nchw = d->out_nchw_f32 ? (float*)out : nullptr;
  p.B = d->B; p.D = d->D; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout;
  p.ks = d->ksize; p.ndim = d->ndim; p.up2 = d->up2; p.circ = d->circular;
  p.Di = (d->up2 && d->ndim == 3) ? d->D / 2 : d->D;
  p.Hi = d->up2 ? d->H / 2 : d->H;
  p.Wi = d->up2 ? d->W / 2 : d->W;
  const int taps = d->ndim == 3 ? d->ksize * d->ksize * d->ksize : d->ksize * d->ksize;
  const int64_t M = (int64_t)d->B * d->D * d->H * d->W;
  if (M > 0x7fffffff) { set_error("dsk_conv_fwd: too many output pixels"); return DSK_ERR_ARG; }
  p.M = (int)M; p.N = d->Cout; p.K = taps * d->Cin; p.alpha = 1.0f; p.act = 0;
  dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN, 1);
  DSK_LAUNCH((ffma_gemm_kernel<ConvProb<TI, TO>>), grid, GT, 0, st, p);
  return DSK_OK;
}

}  // namespace dsk

namespace dsk {
int conv_small_dispatch(const dsk_conv_desc* d, const void* in, const void* w, const float* bias, const float* chan_bias,
                        const void* residual, void* out, cudaStream_t st);
}
using namespace dsk;

// implemented in conv_tc.cu (tcgen05 path); returns DSK_ERR_UNSUPPORTED for shapes it does not take
extern "C" int dsk_conv_fwd_tc(const dsk_conv_desc* d, const void* in, const void* w, const float* bias,
                               const float* chan_bias, const void* residual, void* out, void* stream);

extern "C" int dsk_conv_fwd_ffma(const dsk_conv_desc* d, const void* in, const void* w, const float* bias,
                                 const float* chan_bias, const void* residual, void* out, void* stream) {
  DSK_REQUIRE(d && in && w && out, "dsk_conv_fwd: null pointer");
  DSK_REQUIRE(d->circular != 2, "dsk_conv_fwd: a pre-padded input (circular = 2) is a tcgen05-path layout");
  DSK_REQUIRE(residual == nullptr || d->res_dtype == DSK_RES_SAME || (d->res_dtype == DSK_RES_F32 && d->out_dtype == DSK_F32 && !d->out_nchw_f32),
              "dsk_conv_fwd: an fp32 residual with a 16-bit output (res_dtype) is a tcgen05-path combination");
  DSK_REQUIRE(d->B > 0 && d->D > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "dsk_conv_fwd: bad shape");
  DSK_REQUIRE((d->ksize == 1 || d->ksize == 3) && (d->ndim == 2 || d->ndim == 3), "dsk_conv_fwd: ksize=%d ndim=%d unsupported", d->ksize, d->ndim);
  DSK_REQUIRE(d->ndim == 3 || d->D == 1, "dsk_conv_fwd: ndim=2 needs D=1");
  DSK_REQUIRE(!d->up2 || (d->H % 2 == 0 && d->W % 2 == 0 && (d->ndim == 2 || d->D % 2 == 0)), "dsk_conv_fwd: up2 needs even output size");
  cudaStream_t st = as_stream(stream);
  {  // few-channel first/last convs are bandwidth kernels (conv_small.cu), not GEMM tiles
    const int handled = conv_small_dispatch(d, in, w, bias, chan_bias, residual, out, st);
    if (handled != 0) return handled < 0 ? handled : DSK_OK;
  }
  const int ti = d->in_dtype, to = d->out_nchw_f32 ? (residual ? d->out_dtype : DSK_F32) : d->out_dtype;
  if (ti == DSK_F32 && to == DSK_F32) return launch_conv<float, float>(d, in, w, bias, chan_bias, residual, out, st);
  if (ti == DSK_BF16 && to == DSK_BF16) return launch_conv<__nv_bfloat16, __nv_bfloat16>(d, in, w, bias, chan_bias, residual, out, st);
  if (ti == DSK_BF16 && to == DSK_F32) return launch_conv<__nv_bfloat16, float>(d, in, w, bias, chan_bias, residual, out, st);
  if (ti == DSK_F32 && to == DSK_BF16) return launch_conv<float, __nv_bfloat16>(d, in, w, bias, chan_bias, residual, out, st);
  if (ti == DSK_F16 && to == DSK_F16) return launch_conv<__half, __half>(d, in, w, bias, chan_bias, residual, out, st);
  if (ti == DSK_F16 && to == DSK_F32) return launch_conv<__half, float>(d, in, w, bias, chan_bias, residual, out, st);
  if (ti == DSK_F32 && to == DSK_F16) return launch_conv<float, __half>(d, in, w, bias, chan_bias, residual, out, st);
  DSK_REQUIRE(false, "dsk_conv_fwd: bad dtypes %d -> %d", ti, to);
  return DSK_OK;
}

extern "C" int dsk_pack_conv_weight(const float* w_ref, void* w_packed, int Cout, int Cin, int taps, int dtype, void* stream) {
  DSK_REQUIRE(w_ref && w_packed && Cout > 0 && Cin > 0 && taps > 0, "dsk_pack_conv_weight: bad arguments");
  DSK_REQUIRE(dtype >= DSK_F32 && dtype <= DSK_SPLIT_F16, "dsk_pack_conv_weight: bad dtype %d", dtype);
  const int rows = (dtype != DSK_F32 && Cout <= 16) ? 16 : Cout;
  if (pack_tiled_ok(Cout, Cin, taps, rows, dtype)) {
    DSK_LAUNCH(pack_weight_bf16_tiled_kernel, Cout * ((Cin + 63) / 64), 256, 0, as_stream(stream), w_ref, (uint16_t*)w_packed, Cout, Cin,
               taps, 0, dtype);
    return DSK_OK;
  }
  const int grid = grid_for((int64_t)rows * Cin * taps * (dtype == DSK_SPLIT_F16 ? 2 : 1), 256, 8);
  DSK_LAUNCH(pack_weight_kernel, grid, 256, 0, as_stream(stream), w_ref, dtype == DSK_F32 ? (float*)w_packed : nullptr,
             dtype != DSK_F32 ? (uint16_t*)w_packed : nullptr, Cout, Cin, taps, rows, 0, dtype);
  return DSK_OK;
}

extern "C" int dsk_pack_conv_weight_dgrad(const float* w_ref, void* w_packed, int Cout, int Cin, int taps, int dtype, void* stream) {
  DSK_REQUIRE(w_ref && w_packed && Cout > 0 && Cin > 0 && taps > 0, "dsk_pack_conv_weight_dgrad: bad arguments");
  DSK_REQUIRE(dtype == DSK_F32 || is_h16(dtype), "dsk_pack_conv_weight_dgrad: bad dtype %d", dtype);
  if (pack_tiled_ok(Cout, Cin, taps, Cout, dtype)) {
    DSK_LAUNCH(pack_weight_bf16_tiled_kernel, Cin * ((Cout + 63) / 64), 256, 0, as_stream(stream), w_ref, (uint16_t*)w_packed, Cout, Cin,
               taps, 1, dtype);
    return DSK_OK;
  }
  const int grid = grid_for((int64_t)Cout * Cin * taps, 256, 8);
  DSK_LAUNCH(pack_weight_kernel, grid, 256, 0, as_stream(stream), w_ref, dtype == DSK_F32 ? (float*)w_packed : nullptr,
             dtype != DSK_F32 ? (uint16_t*)w_packed : nullptr, Cout, Cin, taps, Cout, 1, dtype);
  return DSK_OK;
}

extern "C" int dsk_pack_conv_weights_multi(const void* jobs, int njobs, int total_blocks, void* stream) {
  DSK_REQUIRE(jobs && njobs > 0 && total_blocks > 0, "dsk_pack_conv_weights_multi: bad arguments");
  DSK_LAUNCH(pack_weight_multi_kernel, total_blocks, 256, 0, as_stream(stream), (const PackJob*)jobs, njobs);
  return DSK_OK;
}

namespace dsk {
int wgrad_reduce_launch(const float* ws, float* dw, int Cout, int Cin, int taps, int nsplit, int accumulate, cudaStream_t st) {
  DSK_LAUNCH(wgrad_reduce_kernel, grid_for((int64_t)taps * Cin * Cout, 256, 8), 256, 0, st, ws, dw, Cout, Cin, taps, nsplit, accumulate);
  return DSK_OK;
}
}  // namespace dsk
extern "C" int64_t dsk_conv_wgrad_tc_ws_bytes(const dsk_conv_desc* d);
namespace dsk {
// conv_small.cu: weight gradient of the few-channel first / last convolutions
int64_t wgrad_few_ws_bytes(const dsk_conv_desc* d);
int wgrad_few_dispatch(const dsk_conv_desc* d, const void* x, const void* dy, float* dw, void* ws, int accumulate, cudaStream_t st);
}

// ---- weight gradient (CUDA-core path) ---------------------------------------------------------------------------
static int wgrad_splits(int M, int N, int K) {
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  int want = (4 * DSK_NUM_SMS + tiles - 1) / tiles;          // ~4 waves of CTAs
  int maxs = (K + 4 * BK - 1) / (4 * BK);                    // at least 64 pixels per slice
  if (want > maxs) want = maxs;
  if (want > 512) want = 512;
  if (want < 1) want = 1;
  return want;
}

extern "C" int64_t dsk_conv_wgrad_ws_bytes(const dsk_conv_desc* d) {
  if (!d || d->B <= 0 || d->Cin <= 0 || d->Cout <= 0) return 0;
  const int taps = d->ndim == 3 ? d->ksize * d->ksize * d->ksize : d->ksize * d->ksize;
  const int64_t K = (int64_t)d->B * d->D * d->H * d->W;
  if (K > 0x7fffffff) return 0;
  const int M = taps * d->Cin, N = d->Cout;
  const int64_t ffma = (int64_t)wgrad_splits(M, N, (int)K) * M * N * (int64_t)sizeof(float);
  const int64_t tc = dsk_conv_wgrad_tc_ws_bytes(d);
  const int64_t few = wgrad_few_ws_bytes(d);
  const int64_t m = ffma > tc ? ffma : tc;
  return m > few ? m : few;
}

template <typename TI, typename TG>
static int launch_wgrad(const dsk_conv_desc* d, const void* x, const void* dy, float* dw, void* ws, int accumulate, cudaStream_t st) {
  const int taps = d->ndim == 3 ? d->ksize * d->ksize * d->ksize : d->ksize * d->ksize;
  WgradProb<TI, TG> p;
  p.x = (const TI*)x; p.dy = (const TG*)dy; p.ws = (float*)ws;
  p.B = d->B; p.D = d->D; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout; p.ks = d->ksize; p.ndim = d->ndim; p.up2 = d->up2;
  p.circ = d->circular;
  p.Di = (d->up2 && d->ndim == 3) ? d->D / 2 : d->D;
  p.Hi = d->up2 ? d->H / 2 : d->H;
  p.Wi = d->up2 ? d->W / 2 : d->W;
  p.M = taps * d->Cin; p.N = d->Cout; p.K = (int)((int64_t)d->B * d->D * d->H * d->W);
  const int nsplit = wgrad_splits(p.M, p.N, p.K);
  int kper = (p.K + nsplit - 1) / nsplit;
  kper = (kper + BK - 1) / BK * BK;
  p.kper = kper; p.kb = 0; p.ke = 0;
  const int used = (p.K + kper - 1) / kper;                  // slices that own at least one pixel
  dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN, used);
  DSK_LAUNCH((ffma_gemm_kernel<WgradProb<TI, TG>>), grid, GT, 0, st, p);
  return wgrad_reduce_launch((const float*)ws, dw, d->Cout, d->Cin, taps, used, accumulate, st);
}

// implemented in wgrad_tc.cu (tcgen05 path); returns DSK_ERR_UNSUPPORTED for shapes it does not take
extern "C" int dsk_conv_wgrad_tc(const dsk_conv_desc* d, const void* x, const void* dy, float* dw, void* ws, int accumulate,
                                 void* stream);

extern "C" int dsk_conv_wgrad(const dsk_conv_desc* d, const void* x, const void* dy, float* dw, void* ws, int accumulate,
                              void* stream) {
  DSK_REQUIRE(d && x && dy && dw && ws, "dsk_conv_wgrad: null pointer");
  DSK_REQUIRE(d->B > 0 && d->D > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "dsk_conv_wgrad: bad shape");
  DSK_REQUIRE((d->ksize == 1 || d->ksize == 3) && (d->ndim == 2 || d->ndim == 3), "dsk_conv_wgrad: ksize=%d ndim=%d unsupported", d->ksize, d->ndim);
  DSK_REQUIRE(d->ndim == 3 || d->D == 1, "dsk_conv_wgrad: ndim=2 needs D=1");
  DSK_REQUIRE((int64_t)d->B * d->D * d->H * d->W <= 0x7fffffff, "dsk_conv_wgrad: too many pixels");
  if (d->w_dtype == DSK_BF16) {
    const int rc = dsk_conv_wgrad_tc(d, x, dy, dw, ws, accumulate, stream);
    if (rc != DSK_ERR_UNSUPPORTED) return rc;
  }
  cudaStream_t st = as_stream(stream);
  {
    const int rc = wgrad_few_dispatch(d, x, dy, dw, ws, accumulate, st);
    if (rc != 0) return rc < 0 ? rc : DSK_OK;
  }
  if (d->in_dtype == DSK_F32 && d->out_dtype == DSK_F32) return launch_wgrad<float, float>(d, x, dy, dw, ws, accumulate, st);
  if (d->in_dtype == DSK_BF16 && d->out_dtype == DSK_BF16) return launch_wgrad<__nv_bfloat16, __nv_bfloat16>(d, x, dy, dw, ws, accumulate, st);
  if (d->in_dtype == DSK_F32 && d->out_dtype == DSK_BF16) return launch_wgrad<float, __nv_bfloat16>(d, x, dy, dw, ws, accumulate, st);
  if (d->in_dtype == DSK_BF16 && d->out_dtype == DSK_F32) return launch_wgrad<__nv_bfloat16, float>(d, x, dy, dw, ws, accumulate, st);
  DSK_REQUIRE(false, "dsk_conv_wgrad: bad dtypes %d / %d", d->in_dtype, d->out_dtype);
  return DSK_OK;
}

extern "C" int dsk_gemm_f32_ex(const float* A, const float* Bm, float* Cm, const float* bias, int M, int N, int K, int lda,
                               int ldb, int ldc, int64_t strideA, int64_t strideB, int64_t strideC, int batch, int transA,
                               int transB, float alpha, float beta, int act, void* stream) {
  DSK_REQUIRE(A && Bm && Cm, "dsk_gemm_f32: null pointer");
  DSK_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0 && batch <= 65535, "dsk_gemm_f32: bad shape M=%d N=%d K=%d batch=%d", M, N, K, batch);
  DSK_REQUIRE(lda >= (transA ? M : K) && ldc >= N && ldb >= (transB ? K : N), "dsk_gemm_f32: bad leading dimensions");
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, batch);
  DSK_REQUIRE(grid.y <= 65535, "dsk_gemm_f32: N too large");
#define GO(TA, TB)                                                                                                    \
  {                                                                                                                   \
    GemmProb<TA, TB> p{A, Bm, Cm, bias, M, N, K, lda, ldb, ldc, strideA, strideB, strideC, alpha, act, beta};         \
    DSK_LAUNCH((ffma_gemm_kernel<GemmProb<TA, TB>>), grid, GT, 0, as_stream(stream), p);                              \
  }
  if (transA && transB) GO(true, true)
  else if (transA) GO(true, false)
  else if (transB) GO(false, true)
  else GO(false, false)
#undef GO
  return DSK_OK;
}

extern "C" int dsk_gemm_f32(const float* A, const float* Bm, float* Cm, const float* bias, int M, int N, int K, int lda,
                            int ldb, int ldc, int64_t strideA, int64_t strideB, int64_t strideC, int batch, int transB,
                            float alpha, int act, void* stream) {
  return dsk_gemm_f32_ex(A, Bm, Cm, bias, M, N, K, lda, ldb, ldc, strideA, strideB, strideC, batch, 0, transB, alpha, 0.0f, act,
                         stream);
}

// Grouped fp32 GEMM over a device table of problems (see GroupedGemmDesc): C_g = act(op(A_g) op(B_g) + bias_g), Z_g = pre-activation.
extern "C" int dsk_grouped_gemm_f32(const void* table, int ngroups, int max_m, int max_n, int transA, int transB, int act, void* stream) {
  DSK_REQUIRE(table && ngroups > 0 && ngroups <= 65535 && max_m > 0 && max_n > 0 && act >= 0 && act <= 2, "dsk_grouped_gemm_f32: bad arguments");
  dim3 grid((max_m + BM - 1) / BM, (max_n + BN - 1) / BN, ngroups);
  DSK_REQUIRE(grid.y <= 65535, "dsk_grouped_gemm_f32: N too large");
#define GO(TA, TB)                                                                                       \
  {                                                                                                      \
    GroupedGemmProb<TA, TB> p;                                                                           \
    p.A = nullptr; p.Bm = nullptr; p.C = nullptr; p.bias = nullptr; p.M = p.N = p.K = 0;                 \
    p.lda = p.ldb = p.ldc = 0; p.sA = p.sB = p.sC = 0; p.alpha = 1.0f; p.act = act; p.beta = 0.0f;        \
    p.table = (const GroupedGemmDesc*)table; p.Z = nullptr;                                              \
    DSK_LAUNCH((ffma_gemm_kernel<GroupedGemmProb<TA, TB>>), grid, GT, 0, as_stream(stream), p);          \
  }
  if (transA && transB) GO(true, true)
  else if (transA) GO(true, false)
  else if (transB) GO(false, true)
  else GO(false, false)
#undef GO
  return DSK_OK;
}
