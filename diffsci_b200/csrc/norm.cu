// norm.cu -- K4: group LayerNorm / RMS norm + affine (+FiLM) + SiLU on channels-last tensors.
//
// Reference: torch.nn.GroupNorm(G, C) and GroupRMSNorm(G, C) followed by SiLU in ResnetBlockC
// (nets/commonlayers.py:362-384, 824-831); ADM GroupNorm(1, C) + FiLM (nets/adm.py:305-329).
// GroupRMSNorm alone costs the reference 7 full HBM passes; here a norm is one read for the
// statistics and one read + one write for the fused normalise/affine/FiLM/SiLU.
//
// Three launches: (1) per-(sample, spatial chunk) partial sums in fp32 per thread, combined in
// fp64 across the block, written to the workspace (deterministic, no atomics);
// (2) per-(sample, group) finalisation -> (mean, rstd); (3) the apply pass.
#include "common.cuh"

namespace dsk {

constexpr int NORM_THREADS = 256;

template <typename T> struct NormVec;                 // 16-byte vectors: 4 fp32 or 8 bf16
template <> struct NormVec<float> {
  static constexpr int V = 4;
  static __device__ __forceinline__ void ld(const float* p, float* o) {
    float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};
template <> struct NormVec<__nv_bfloat16> {
  static constexpr int V = 8;
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* o) {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[2 * j] = __low2float(h[j]); o[2 * j + 1] = __high2float(h[j]); }
  }
};
template <> struct NormVec<__half> {
  static constexpr int V = 8;
  static __device__ __forceinline__ void ld(const __half* p, float* o) {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[2 * j] = __low2float(h[j]); o[2 * j + 1] = __high2float(h[j]); }
  }
};
template <typename T, int V> __device__ __forceinline__ void st_vec(T* p, const float* v);
template <> __device__ __forceinline__ void st_vec<float, 4>(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void st_vec<float, 8>(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void st_vec<__nv_bfloat16, 4>(__nv_bfloat16* p, const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 raw;
  raw.x = *reinterpret_cast<uint32_t*>(&a);
  raw.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = raw;
}
template <> __device__ __forceinline__ void st_vec<__nv_bfloat16, 8>(__nv_bfloat16* p, const float* v) {
  uint4 o;
  __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) oh[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
  *reinterpret_cast<uint4*>(p) = o;
}

template <> __device__ __forceinline__ void st_vec<__half, 4>(__half* p, const float* v) {
  __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  uint2 raw;
  raw.x = *reinterpret_cast<uint32_t*>(&a);
  raw.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = raw;
}
template <> __device__ __forceinline__ void st_vec<__half, 8>(__half* p, const float* v) {
  uint4 o;
  __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
  *reinterpret_cast<uint4*>(p) = o;
}
// split-fp16 output (the A operand of the split tensor-core convolutions, conv_tc.cu): hi = fp16(v) at p, lo = fp16(2^11 (v - hi))
// at p + lo_off; hi + 2^-11 lo carries 22 mantissa bits of v (the scaling keeps lo out of fp16's subnormal range)
template <int V> __device__ __forceinline__ void st_split(__half* p, int64_t lo_off, const float* v) {
  float lo[V];
  float hi[V];
#pragma unroll
  for (int e = 0; e < V; ++e) { hi[e] = __half2float(__float2half_rn(v[e])); lo[e] = (v[e] - hi[e]) * 2048.0f; }
  st_vec<__half, V>(p, hi);
  st_vec<__half, V>(p + lo_off, lo);
}

// block size of the per-(sample, group) finalize kernels: the power of two covering `items` partials, 128 .. 1024
static inline int finalize_threads(int64_t items) {
  int t = 128;
  while (t < 1024 && t < items) t <<= 1;
  return t;
}
static inline int norm_chunks(int B, int64_t S, int C, int V) {
  int cv = C / V;
  int pl = NORM_THREADS / cv;
  if (pl < 1) pl = 1;
  // each thread accumulates at most 64 pixels in fp32 (then everything is combined in fp64)
  int64_t need = (S + (int64_t)pl * 64 - 1) / ((int64_t)pl * 64);
  int64_t fill = (2 * DSK_NUM_SMS + B - 1) / B;                       // enough blocks to occupy the chip
  int64_t by_work = (S + (int64_t)pl * 8 - 1) / ((int64_t)pl * 8);    // but >= 8 pixels per thread
  if (fill > by_work) fill = by_work;
  int64_t n = need > fill ? need : fill;
  if (n < 1) n = 1;
  return (int)n;
}

// partial[b][chunk][c] = (sum, sumsq) in fp64.  Thread = fixed V-channel vector, strided over the chunk's pixels.
template <typename T>
__global__ void __launch_bounds__(NORM_THREADS) norm_partial_kernel(const T* __restrict__ x, double2* __restrict__ partial,
                                                                     int64_t S, int C, int nchunks) {
  constexpr int V = NormVec<T>::V;
  extern __shared__ float2 sm[];   // [pl][C] per-thread fp32 partials (<= 64 pixels each)
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int cv = C / V;
  const int pl = max(1, NORM_THREADS / cv);
  const int64_t per = (S + nchunks - 1) / nchunks;
  const int64_t s0 = (int64_t)chunk * per;
  const int64_t s1 = min(S, s0 + per);
  const T* xb = x + (int64_t)b * S * C;
  for (int v = threadIdx.x; v < pl * cv; v += NORM_THREADS) {
    const int lane = v / cv, c0 = (v - lane * cv) * V;
    float sum[V], sq[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { sum[k] = 0; sq[k] = 0; }
    int64_t s = s0 + lane;
    // 8 independent 16-byte loads in flight per thread before any dependent arithmetic
    for (; s + 7 * (int64_t)pl < s1; s += 8 * (int64_t)pl) {
      float e[8][V];
#pragma unroll
      for (int u = 0; u < 8; ++u) NormVec<T>::ld(xb + (s + u * (int64_t)pl) * C + c0, e[u]);
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int k = 0; k < V; ++k) {
          sum[k] += e[u][k];
          sq[k] = fmaf(e[u][k], e[u][k], sq[k]);
        }
    }
    for (; s < s1; s += pl) {
      float e[V];
      NormVec<T>::ld(xb + s * C + c0, e);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        sum[k] += e[k];
        sq[k] = fmaf(e[k], e[k], sq[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) sm[lane * C + c0 + k] = make_float2(sum[k], sq[k]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += NORM_THREADS) {
    double a = 0, q = 0;
    for (int l = 0; l < pl; ++l) {
      float2 v = sm[l * C + c];
      a += (double)v.x;
      q += (double)v.y;
    }
    partial[((int64_t)b * nchunks + chunk) * C + c] = make_double2(a, q);
  }
}

// One block per (sample, group): combine the chunk partials, then fold mean / rstd / gamma / beta / FiLM into a
// per-(b, c) scale-shift pair:  y = act(x * a + s).
__global__ void __launch_bounds__(1024) norm_finalize_kernel(const double2* __restrict__ partial, float2* __restrict__ table,
                                                             float2* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const float* __restrict__ fsc, const float* __restrict__ fsh,
                                                             int64_t S, int C, int G, int nchunks, int mode, float eps) {
  const int b = blockIdx.x / G, g = blockIdx.x % G;
  const int cg = C / G;
  double a = 0, q = 0;
  // the (sample, group) slice of `partial` is contiguous when G == 1 (chunk-major rows of C) -- and in general the loads are
  // independent: unrolled so that 8 are in flight per thread (one at a time: 13 us per launch at 32 blocks x 128 threads,
  // latency-bound; 28 launches per ADM iteration)
#pragma unroll 8
  for (int i = threadIdx.x; i < nchunks * cg; i += blockDim.x) {
    const int chunk = i / cg, c = g * cg + (i - chunk * cg);
    double2 v = partial[((int64_t)b * nchunks + chunk) * C + c];
    a += v.x;
    q += v.y;
  }
  __shared__ double sa[1024], sq[1024];                // block = 128 .. 1024 threads (a power of two; norm_finalize_threads)
  __shared__ float2 mr;
  sa[threadIdx.x] = a;
  sq[threadIdx.x] = q;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sa[threadIdx.x] += sa[threadIdx.x + o];
      sq[threadIdx.x] += sq[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double n = (double)S * cg;
    double mean = sa[0] / n, ex2 = sq[0] / n;
    if (mode == 0) {
      double var = ex2 - mean * mean;
      if (var < 0) var = 0;
      mr = make_float2((float)mean, 1.0f / sqrtf((float)var + eps));
    } else {
      mr = make_float2(0.0f, 1.0f / sqrtf((float)ex2 + eps));  // x / sqrt(mean(x^2) + eps), commonlayers.py:377-378
    }
    stats[blockIdx.x] = mr;                                      // (mean, rstd) per (b, g): kept for dsk_norm_act_bwd
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cg; i += blockDim.x) {
    const int c = g * cg + i;
    float sc = mr.y, sh = -mr.x * mr.y;
    if (gamma != nullptr) { sc *= gamma[c]; sh = sh * gamma[c] + beta[c]; }
    if (fsc != nullptr) { const float f = fsc[(int64_t)b * C + c]; sc *= f; sh = sh * f + fsh[(int64_t)b * C + c]; }
    table[(int64_t)b * C + c] = make_float2(sc, sh);
  }
}

// G == C (one channel per group: the PUNetG norms).  Block = 32 channels x 32 chunk lanes of one sample: the chunk partials
// are summed by 32 lanes per channel (coalesced 512-byte rows, 32-way memory parallelism) and combined through shared memory
// in a fixed order -- one serial thread per (b, c) took 24 us per launch at B = 2 (56 launches per training iteration).
__global__ void __launch_bounds__(1024) norm_finalize_pc_kernel(const double2* __restrict__ partial, float2* __restrict__ table,
                                                                 float2* __restrict__ stats, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, const float* __restrict__ fsc,
                                                                 const float* __restrict__ fsh, int64_t S, int B, int C, int nchunks, int mode,
                                                                 float eps) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  double a = 0, q = 0;
  if (c < C)
    for (int ch = rl; ch < nchunks; ch += 32) {
      const double2 v = partial[((int64_t)b * nchunks + ch) * C + c];
      a += v.x;
      q += v.y;
    }
  __shared__ double ra[32][33], rq[32][33];
  ra[rl][threadIdx.x & 31] = a;
  rq[rl][threadIdx.x & 31] = q;
  __syncthreads();
  if (rl != 0 || c >= C) return;
  a = 0; q = 0;
  for (int k = 0; k < 32; ++k) { a += ra[k][threadIdx.x & 31]; q += rq[k][threadIdx.x & 31]; }
  const int i = b * C + c;
  const double n = (double)S;
  const double mean = a / n, ex2 = q / n;
  float2 mr;
  if (mode == 0) {
    double var = ex2 - mean * mean;
    if (var < 0) var = 0;
    mr = make_float2((float)mean, 1.0f / sqrtf((float)var + eps));
  } else {
    mr = make_float2(0.0f, 1.0f / sqrtf((float)ex2 + eps));
  }
  stats[i] = mr;
  float sc = mr.y, sh = -mr.x * mr.y;
  if (gamma != nullptr) { sc *= gamma[c]; sh = sh * gamma[c] + beta[c]; }
  if (fsc != nullptr) { const float f = fsc[i]; sc *= f; sh = sh * f + fsh[i]; }
  table[i] = make_float2(sc, sh);
}

// G == C (one channel per group: the PUNetG norms): one THREAD per (b, c): for large batches with few chunks per sample (B*C threads fill the machine).
__global__ void __launch_bounds__(128) norm_finalize_pc_serial_kernel(const double2* __restrict__ partial, float2* __restrict__ table,
                                                                float2* __restrict__ stats, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, const float* __restrict__ fsc,
                                                                const float* __restrict__ fsh, int64_t S, int B, int C, int nchunks, int mode,
                                                                float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  double a = 0, q = 0;
  for (int ch = 0; ch < nchunks; ++ch) {
    const double2 v = partial[((int64_t)b * nchunks + ch) * C + c];
    a += v.x;
    q += v.y;
  }
  const double n = (double)S;
  const double mean = a / n, ex2 = q / n;
  float2 mr;
  if (mode == 0) {
    double var = ex2 - mean * mean;
    if (var < 0) var = 0;
    mr = make_float2((float)mean, 1.0f / sqrtf((float)var + eps));
  } else {
    mr = make_float2(0.0f, 1.0f / sqrtf((float)ex2 + eps));
  }
  stats[i] = mr;
  float sc = mr.y, sh = -mr.x * mr.y;
  if (gamma != nullptr) { sc *= gamma[c]; sh = sh * gamma[c] + beta[c]; }
  if (fsc != nullptr) { const float f = fsc[i]; sc *= f; sh = sh * f + fsh[i]; }
  table[i] = make_float2(sc, sh);
}

// Finalize from the statistics a convolution epilogue left behind (conv_tc.cu): stats[b][slot][c] = (sum, sum of squares)
// over the pixels one epilogue warp stored.  G == C only.  Block = 32 channels x 8 slot lanes of one sample.
__global__ void __launch_bounds__(1024) norm_finalize_stats_kernel(const float2* __restrict__ st, float2* __restrict__ table,
                                                                    float2* __restrict__ stats, const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, const float* __restrict__ fsc,
                                                                    const float* __restrict__ fsh, int64_t S, int C, int nslots, int mode,
                                                                    float eps) {
  // block = 32 channels x 32 slot lanes of one sample; 4 loads in flight per thread
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  double a = 0, q = 0;
  if (c < C) {
    const float2* base = st + (int64_t)b * nslots * C + c;
    int s = rl;
    for (; s + 96 < nslots; s += 128) {
      const float2 v0 = base[(int64_t)s * C], v1 = base[(int64_t)(s + 32) * C], v2 = base[(int64_t)(s + 64) * C],
                   v3 = base[(int64_t)(s + 96) * C];
      a += ((double)v0.x + (double)v1.x) + ((double)v2.x + (double)v3.x);
      q += ((double)v0.y + (double)v1.y) + ((double)v2.y + (double)v3.y);
    }
    for (; s < nslots; s += 32) {
      const float2 v = base[(int64_t)s * C];
      a += (double)v.x;
      q += (double)v.y;
    }
  }
  __shared__ double ra[32][33], rq[32][33];
  ra[rl][threadIdx.x & 31] = a;
  rq[rl][threadIdx.x & 31] = q;
  __syncthreads();
  if (rl != 0 || c >= C) return;
  a = 0; q = 0;
  for (int k = 0; k < 32; ++k) { a += ra[k][threadIdx.x & 31]; q += rq[k][threadIdx.x & 31]; }
  const double n = (double)S;
  const double mean = a / n, ex2 = q / n;
  float2 mr;
  if (mode == 0) {
    double var = ex2 - mean * mean;
    if (var < 0) var = 0;
    mr = make_float2((float)mean, 1.0f / sqrtf((float)var + eps));
  } else {
    mr = make_float2(0.0f, 1.0f / sqrtf((float)ex2 + eps));
  }
  const int64_t i = (int64_t)b * C + c;
  stats[i] = mr;
  float sc = mr.y, sh = -mr.x * mr.y;
  if (gamma != nullptr) { sc *= gamma[c]; sh = sh * gamma[c] + beta[c]; }
  if (fsc != nullptr) { const float f = fsc[i]; sc *= f; sh = sh * f + fsh[i]; }
  table[i] = make_float2(sc, sh);
}

template <typename TO> __device__ __forceinline__ float silu_out(float v) { return silu_f(v); }
// bf16 output keeps 8 mantissa bits: the approximate exp / divide (rel. error ~1e-6) is invisible after rounding and
// takes the kernel from instruction-bound back to HBM-bound
template <> __device__ __forceinline__ float silu_out<__nv_bfloat16>(float v) { return __fdividef(v, 1.0f + __expf(-v)); }
template <> __device__ __forceinline__ float silu_out<__half>(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

// Thread = fixed V-channel vector (scale/shift held in registers), strided over the pixels of one sample.
// SPLIT (TO = __half): y is a split-fp16 tensor [B, S, 2C] -- channels [0, C) the fp16 value, [C, 2C) the fp16 remainder.
template <typename TI, typename TO, bool SPLIT = false>
__global__ void __launch_bounds__(NORM_THREADS) norm_apply_kernel(const TI* __restrict__ x, TO* __restrict__ y,
                                                                   const float2* __restrict__ table, int64_t S, int C, int silu) {
  constexpr int V = NormVec<TI>::V;
  const int b = blockIdx.y;
  const int cv = C / V;
  const int pl = max(1, NORM_THREADS / cv);
  const int CO = SPLIT ? 2 * C : C;                     // output row length
  const TI* xb = x + (int64_t)b * S * C;
  TO* yb = y + (int64_t)b * S * CO;
  for (int v = threadIdx.x; v < pl * cv; v += NORM_THREADS) {
    const int lane = v / cv, c0 = (v - lane * cv) * V;
    float sc[V], sh[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float2 t = table[(int64_t)b * C + c0 + k];
      sc[k] = t.x; sh[k] = t.y;
    }
    const int64_t step = (int64_t)gridDim.x * pl;
    int64_t s = (int64_t)blockIdx.x * pl + lane;
    for (; s + step < S; s += 2 * step) {     // two independent load->store chains per iteration
      float e0[V], e1[V], o0[V], o1[V];
      NormVec<TI>::ld(xb + s * C + c0, e0);
      NormVec<TI>::ld(xb + (s + step) * C + c0, e1);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float t0 = fmaf(e0[k], sc[k], sh[k]), t1 = fmaf(e1[k], sc[k], sh[k]);
        o0[k] = silu ? (SPLIT ? silu_f(t0) : silu_out<TO>(t0)) : t0;
        o1[k] = silu ? (SPLIT ? silu_f(t1) : silu_out<TO>(t1)) : t1;
      }
      if constexpr (SPLIT) {
        st_split<V>(yb + s * CO + c0, C, o0);
        st_split<V>(yb + (s + step) * CO + c0, C, o1);
      } else {
        st_vec<TO, V>(yb + s * C + c0, o0);
        st_vec<TO, V>(yb + (s + step) * C + c0, o1);
      }
    }
    for (; s < S; s += step) {
      float e[V], o[V];
      NormVec<TI>::ld(xb + s * C + c0, e);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float t = fmaf(e[k], sc[k], sh[k]);
        o[k] = silu ? (SPLIT ? silu_f(t) : silu_out<TO>(t)) : t;
      }
      if constexpr (SPLIT) st_split<V>(yb + s * CO + c0, C, o);
      else st_vec<TO, V>(yb + s * C + c0, o);
    }
  }
}

// The same apply pass writing a HALO-PADDED tensor y[B, D+2*pd, H+2, W+2, C] (pd = 1 for 3-D): every pixel goes to its interior
// position and, on a face / edge / corner, to the wrapped halo positions as well -- the input layout of a circular convolution on
// the tcgen05 path (conv_tc.cu), so periodic networks need no separate padding pass over the conv inputs.
template <typename TI, typename TO>
__global__ void __launch_bounds__(NORM_THREADS) norm_apply_pad_kernel(const TI* __restrict__ x, TO* __restrict__ y,
                                                                       const float2* __restrict__ table, int D, int H, int W, int C,
                                                                       int pd, int silu) {
  constexpr int V = NormVec<TI>::V;
  const int b = blockIdx.y;
  const int cv = C / V;
  const int pl = max(1, NORM_THREADS / cv);
  const unsigned S = (unsigned)D * H * W;
  const int Dp = D + 2 * pd, Hp = H + 2, Wp = W + 2;
  const TI* xb = x + (int64_t)b * S * C;
  TO* yb = y + (int64_t)b * Dp * Hp * Wp * C;
  for (int v = threadIdx.x; v < pl * cv; v += NORM_THREADS) {
    const int lane = v / cv, c0 = (v - lane * cv) * V;
    float sc[V], sh[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float2 t = table[(int64_t)b * C + c0 + k];
      sc[k] = t.x; sh[k] = t.y;
    }
    const unsigned step = gridDim.x * pl;
    for (unsigned s = blockIdx.x * pl + lane; s < S; s += step) {
      float e[V], o[V];
      NormVec<TI>::ld(xb + (int64_t)s * C + c0, e);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float t = fmaf(e[k], sc[k], sh[k]);
        o[k] = silu ? silu_out<TO>(t) : t;
      }
      const unsigned w = s % W, t2 = s / W, h = t2 % H, d = t2 / H;
      // target coordinates per axis: the interior position, plus the opposite halo when the pixel lies on a face
      int tw[3], th[3], td[3], nw = 1, nh = 1, nd = 1;
      tw[0] = w + 1; th[0] = h + 1; td[0] = d + pd;
      if (w == 0) tw[nw++] = W + 1;
      if (w == (unsigned)W - 1) tw[nw++] = 0;
      if (h == 0) th[nh++] = H + 1;
      if (h == (unsigned)H - 1) th[nh++] = 0;
      if (pd) {
        if (d == 0) td[nd++] = D + 1;
        if (d == (unsigned)D - 1) td[nd++] = 0;
      }
      for (int a = 0; a < nd; ++a)
        for (int bb = 0; bb < nh; ++bb)
          for (int cc = 0; cc < nw; ++cc)
            st_vec<TO, V>(yb + (((int64_t)td[a] * Hp + th[bb]) * Wp + tw[cc]) * C + c0, o);
    }
  }
}

// ---- small samples, per-channel norms (G == C), bf16: ONE kernel, ONE read.  A CTA owns (sample b, 32 channels): it loads its
// S x 32-channel slab into shared memory (64 B per pixel), reduces (sum, sum of squares) per channel -- fp32 per thread over at
// most S/64 pixels, fp64 across threads, fixed order -- folds the affine / FiLM parameters and applies scale/shift + SiLU from
// shared memory.  Replaces the statistics pass + finalize + apply launches (31 us -> ~12 us per norm on MNIST-size tensors,
// where each of the three is launch-latency bound); writes the same (scale, shift) / (mean, rstd) tables for the backward.
constexpr int SLAB_THREADS = 256, SLAB_CH = 32;
template <typename T>
__global__ void __launch_bounds__(SLAB_THREADS) norm_slab_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                                  float2* __restrict__ table, float2* __restrict__ stats,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ fsc, const float* __restrict__ fsh, int S, int C,
                                                                  int mode, int silu, float eps) {
  extern __shared__ __align__(16) uint8_t slab_raw[];
  uint4* slab = reinterpret_cast<uint4*>(slab_raw);                 // [S][4] 16-byte pieces (8 channels each)
  __shared__ double red_s[SLAB_THREADS / 32][SLAB_CH], red_q[SLAB_THREADS / 32][SLAB_CH];
  __shared__ float2 coef[SLAB_CH];
  const int b = blockIdx.y, c0 = blockIdx.x * SLAB_CH;
  const int j = threadIdx.x & 3, pl = threadIdx.x >> 2;             // piece within the pixel row, pixel lane (64 lanes)
  const T* xb = x + ((int64_t)b * S) * C + c0 + j * 8;
  float s8[8], q8[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s8[k] = 0.0f; q8[k] = 0.0f; }
  constexpr int PSTEP = SLAB_THREADS / 4, UNR = 4;            // 4 independent 16-byte loads in flight per thread
  for (int p0 = pl; p0 < S; p0 += PSTEP * UNR) {
    uint4 raw[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int p = p0 + u * PSTEP;
      raw[u] = p < S ? *reinterpret_cast<const uint4*>(xb + (int64_t)p * C) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int p = p0 + u * PSTEP;
      if (p < S) slab[p * 4 + j] = raw[u];
      const typename H16<T>::T2* h = reinterpret_cast<const typename H16<T>::T2*>(&raw[u]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {                            // zeros past the end add nothing (same order as a rolled loop)
        const float a = __low2float(h[k]), c = __high2float(h[k]);
        s8[2 * k] += a; q8[2 * k] = fmaf(a, a, q8[2 * k]);
        s8[2 * k + 1] += c; q8[2 * k + 1] = fmaf(c, c, q8[2 * k + 1]);
      }
    }
  }
  // lanes of a warp that share j (lane bits 2..4) meet by shuffle; the 8 warps meet in shared memory (fp64, fixed order)
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    double ds = (double)s8[k], dq = (double)q8[k];
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) { ds += __shfl_xor_sync(0xffffffffu, ds, o); dq += __shfl_xor_sync(0xffffffffu, dq, o); }
    if ((threadIdx.x & 31) < 4) { red_s[threadIdx.x >> 5][j * 8 + k] = ds; red_q[threadIdx.x >> 5][j * 8 + k] = dq; }
  }
  __syncthreads();
  if (threadIdx.x < SLAB_CH) {
    const int c = c0 + threadIdx.x;
    double a = 0, q = 0;
    for (int w = 0; w < SLAB_THREADS / 32; ++w) { a += red_s[w][threadIdx.x]; q += red_q[w][threadIdx.x]; }
    const double n = (double)S, mean = a / n, ex2 = q / n;
    float2 mr;
    if (mode == 0) {
      double var = ex2 - mean * mean;
      if (var < 0) var = 0;
      mr = make_float2((float)mean, 1.0f / sqrtf((float)var + eps));
    } else {
      mr = make_float2(0.0f, 1.0f / sqrtf((float)ex2 + eps));
    }
    const int64_t i = (int64_t)b * C + c;
    float sc = mr.y, sh = -mr.x * mr.y;
    if (gamma != nullptr) { sc *= gamma[c]; sh = sh * gamma[c] + beta[c]; }
    if (fsc != nullptr) { const float f = fsc[i]; sc *= f; sh = sh * f + fsh[i]; }
    stats[i] = mr;
    table[i] = make_float2(sc, sh);
    coef[threadIdx.x] = make_float2(sc, sh);
  }
  __syncthreads();
  if (y == nullptr) return;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { const float2 t = coef[j * 8 + k]; sc[k] = t.x; sh[k] = t.y; }
  T* yb = y + ((int64_t)b * S) * C + c0 + j * 8;
#pragma unroll 4
  for (int p = pl; p < S; p += SLAB_THREADS / 4) {
    const uint4 raw = slab[p * 4 + j];
    const typename H16<T>::T2* h = reinterpret_cast<const typename H16<T>::T2*>(&raw);
    float o[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float t0 = fmaf(__low2float(h[k]), sc[2 * k], sh[2 * k]), t1 = fmaf(__high2float(h[k]), sc[2 * k + 1], sh[2 * k + 1]);
      o[2 * k] = silu ? silu_out<T>(t0) : t0;
      o[2 * k + 1] = silu ? silu_out<T>(t1) : t1;
    }
    st_vec<T, 8>(yb + (int64_t)p * C, o);
  }
}

// shapes the slab kernel takes: per-channel bf16 norms whose 32-channel slab of one sample fits in shared memory, with
// enough (sample, channel-group) CTAs to fill the machine
static bool norm_slab_ok(int B, int64_t S, int C, int G, int in_dtype, int out_dtype) {
  static const int off = [] { const char* e = getenv("DSK_NORM_SLAB_OFF"); return e ? atoi(e) : 0; }();   // A/B measurements
  return !off && G == C && is_h16(in_dtype) && out_dtype == in_dtype && C % SLAB_CH == 0 && S * 64 <= 96 * 1024 &&
         (int64_t)B * (C / SLAB_CH) >= DSK_NUM_SMS;
}

}  // namespace dsk

using namespace dsk;

static inline int norm_v(int in_dtype) { return is_h16(in_dtype) ? 8 : 4; }

extern "C" int64_t dsk_norm_ws_bytes(int B, int64_t S, int C) {
  if (B <= 0 || S <= 0 || C <= 0 || C % 4) return 0;
  int nchunks = norm_chunks(B, S, C, 4);   // upper bound over both vector widths
  int n8 = (C % 8 == 0) ? norm_chunks(B, S, C, 8) : 0;
  if (n8 > nchunks) nchunks = n8;
  // [table: B*C float2 (folded scale, shift)] [stats: B*C float2 slots, B*G used (mean, rstd)] [partials]
  return (int64_t)B * nchunks * C * (int64_t)sizeof(double2) + 2 * (int64_t)B * C * (int64_t)sizeof(float2);
}

static int norm_apply_launch(const void* x, void* y, const float2* table, int B, int64_t S, int C, int silu, int in_dtype,
                             int out_dtype, cudaStream_t st) {
  const int V = norm_v(in_dtype);
  const int cv = C / V;
  const int pl = NORM_THREADS / cv > 0 ? NORM_THREADS / cv : 1;
  int64_t gx = (S + (int64_t)pl * 4 - 1) / ((int64_t)pl * 4);          // >= 4 pixels per thread
  const int64_t cap = (8LL * DSK_NUM_SMS + B - 1) / B;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 ag((unsigned)gx, B);
#define APPLY(TI, TO) DSK_LAUNCH((norm_apply_kernel<TI, TO>), ag, NORM_THREADS, 0, st, (const TI*)x, (TO*)y, table, S, C, silu)
  if (in_dtype == DSK_F32 && out_dtype == DSK_F32) APPLY(float, float);
  else if (in_dtype == DSK_F32 && out_dtype == DSK_BF16) APPLY(float, __nv_bfloat16);
  else if (in_dtype == DSK_BF16 && out_dtype == DSK_BF16) APPLY(__nv_bfloat16, __nv_bfloat16);
  else if (in_dtype == DSK_BF16 && out_dtype == DSK_F32) APPLY(__nv_bfloat16, float);
  else if (in_dtype == DSK_F32 && out_dtype == DSK_F16) APPLY(float, __half);
  else if (in_dtype == DSK_F16 && out_dtype == DSK_F16) APPLY(__half, __half);
  else if (in_dtype == DSK_F16 && out_dtype == DSK_F32) APPLY(__half, float);
  else if (in_dtype == DSK_F32 && out_dtype == DSK_SPLIT_F16)
    DSK_LAUNCH((norm_apply_kernel<float, __half, true>), ag, NORM_THREADS, 0, st, (const float*)x, (__half*)y, table, S, C, silu);
  else DSK_REQUIRE(false, "dsk_norm_act: bad dtype combination %d -> %d", in_dtype, out_dtype);
#undef APPLY
  return DSK_OK;
}

extern "C" int dsk_norm_act(const void* x, void* y, const float* gamma, const float* beta, const float* film_scale,
                            const float* film_shift, void* ws, int B, int64_t S, int C, int G, int mode, int silu,
                            int in_dtype, int out_dtype, void* stream) {
  DSK_REQUIRE(x && ws, "dsk_norm_act: null pointer");
  DSK_REQUIRE(B > 0 && B <= 65535 && S > 0 && C > 0 && G > 0 && C % G == 0, "dsk_norm_act: bad shape B=%d S=%lld C=%d G=%d", B, (long long)S, C, G);
  DSK_REQUIRE(in_dtype == DSK_F32 || is_h16(in_dtype), "dsk_norm_act: bad in_dtype %d", in_dtype);
  const int V = norm_v(in_dtype);
  DSK_REQUIRE(C % V == 0, "dsk_norm_act: C=%d must be a multiple of %d for this dtype", C, V);
  DSK_REQUIRE(mode == 0 || mode == 1, "dsk_norm_act: bad mode %d", mode);
  DSK_REQUIRE((gamma == nullptr) == (beta == nullptr), "dsk_norm_act: gamma/beta must both be set or both null");
  DSK_REQUIRE((film_scale == nullptr) == (film_shift == nullptr), "dsk_norm_act: FiLM scale/shift mismatch");
  cudaStream_t st = as_stream(stream);
  const int nchunks = norm_chunks(B, S, C, V);
  float2* table = reinterpret_cast<float2*>(ws);
  float2* stats = table + (int64_t)B * C;
  double2* partial = reinterpret_cast<double2*>(stats + (int64_t)B * C);
  if (norm_slab_ok(B, S, C, G, in_dtype, y == nullptr ? in_dtype : out_dtype)) {
    const size_t smem = (size_t)S * 64;
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(norm_slab_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(norm_slab_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      if (e != cudaSuccess) { set_error("dsk_norm_act: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return DSK_ERR_CUDA; }
      configured = true;
    }
    if (in_dtype == DSK_F16)
      DSK_LAUNCH(norm_slab_kernel<__half>, dim3(C / SLAB_CH, B), SLAB_THREADS, smem, st, (const __half*)x, (__half*)y, table, stats, gamma,
                 beta, film_scale, film_shift, (int)S, C, mode, silu, 1e-5f);
    else
      DSK_LAUNCH(norm_slab_kernel<__nv_bfloat16>, dim3(C / SLAB_CH, B), SLAB_THREADS, smem, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y,
                 table, stats, gamma, beta, film_scale, film_shift, (int)S, C, mode, silu, 1e-5f);
    return DSK_OK;
  }
  const int cv = C / V;
  const int pl = NORM_THREADS / cv > 0 ? NORM_THREADS / cv : 1;
  const size_t smem = (size_t)pl * C * sizeof(float2);
  DSK_REQUIRE(smem <= 48 * 1024, "dsk_norm_act: C=%d too large for the stats kernel", C);
  dim3 pg(nchunks, B);
  if (in_dtype == DSK_F32)
    DSK_LAUNCH(norm_partial_kernel<float>, pg, NORM_THREADS, smem, st, (const float*)x, partial, S, C, nchunks);
  else if (in_dtype == DSK_F16)
    DSK_LAUNCH(norm_partial_kernel<__half>, pg, NORM_THREADS, smem, st, (const __half*)x, partial, S, C, nchunks);
  else
    DSK_LAUNCH(norm_partial_kernel<__nv_bfloat16>, pg, NORM_THREADS, smem, st, (const __nv_bfloat16*)x, partial, S, C, nchunks);
  if (G == C)
  {
    if (nchunks >= 8)
      DSK_LAUNCH(norm_finalize_pc_kernel, dim3((C + 31) / 32, B), 1024, 0, st, partial, table, stats, gamma, beta, film_scale, film_shift, S,
                 B, C, nchunks, mode, 1e-5f);
    else
      DSK_LAUNCH(norm_finalize_pc_serial_kernel, (B * C + 127) / 128, 128, 0, st, partial, table, stats, gamma, beta, film_scale, film_shift,
                 S, B, C, nchunks, mode, 1e-5f);
  }
  else
    // few groups (ADM: G = 1): a block reduces nchunks * C / G partials -- one load per thread instead of a latency-bound loop
    DSK_LAUNCH(norm_finalize_kernel, B * G, finalize_threads((int64_t)nchunks * (C / G)), 0, st, partial, table, stats, gamma, beta, film_scale, film_shift, S, C, G, nchunks, mode,
               1e-5f);
  if (y == nullptr) return DSK_OK;      // statistics + folded table only (dsk_norm_apply_padded follows)
  return norm_apply_launch(x, y, table, B, S, C, silu, in_dtype, out_dtype, st);
}


extern "C" int dsk_norm_act_prestat(const void* x, void* y, const float* gamma, const float* beta, const float* film_scale,
                                    const float* film_shift, const void* conv_stats, int nslots, void* ws, int B, int64_t S, int C,
                                    int G, int mode, int silu, int in_dtype, int out_dtype, void* stream) {
  DSK_REQUIRE(x && ws && conv_stats, "dsk_norm_act_prestat: null pointer");
  DSK_REQUIRE(B > 0 && B <= 65535 && S > 0 && C > 0 && G == C && nslots > 0, "dsk_norm_act_prestat: needs one channel per group (G == C)");
  DSK_REQUIRE(in_dtype == DSK_F32 || is_h16(in_dtype), "dsk_norm_act_prestat: bad in_dtype %d", in_dtype);
  DSK_REQUIRE(C % norm_v(in_dtype) == 0 && (mode == 0 || mode == 1), "dsk_norm_act_prestat: bad C / mode");
  DSK_REQUIRE((gamma == nullptr) == (beta == nullptr) && (film_scale == nullptr) == (film_shift == nullptr), "dsk_norm_act_prestat: affine / FiLM mismatch");
  cudaStream_t st = as_stream(stream);
  float2* table = reinterpret_cast<float2*>(ws);
  float2* stats = table + (int64_t)B * C;
  dim3 fg((C + 31) / 32, B);
  DSK_LAUNCH(norm_finalize_stats_kernel, fg, 1024, 0, st, (const float2*)conv_stats, table, stats, gamma, beta, film_scale, film_shift, S, C,
             nslots, mode, 1e-5f);
  if (y == nullptr) return DSK_OK;      // table only
  return norm_apply_launch(x, y, table, B, S, C, silu, in_dtype, out_dtype, st);
}

extern "C" int dsk_norm_apply_padded(const void* x, void* y_padded, const void* ws, int B, int D, int H, int W, int C, int ndim, int silu,
                                     int in_dtype, int out_dtype, void* stream) {
  DSK_REQUIRE(x && y_padded && ws, "dsk_norm_apply_padded: null pointer");
  DSK_REQUIRE(B > 0 && B <= 65535 && D > 0 && H > 0 && W > 0 && C > 0 && ((ndim == 2 && D == 1) || ndim == 3), "dsk_norm_apply_padded: bad shape");
  DSK_REQUIRE((int64_t)D * H * W < 0x7fffffff, "dsk_norm_apply_padded: sample too large");
  DSK_REQUIRE(is_h16(in_dtype) && out_dtype == in_dtype && C % 8 == 0, "dsk_norm_apply_padded: 16-bit tensors with C %% 8 == 0 only");
  const float2* table = reinterpret_cast<const float2*>(ws);
  const int64_t S = (int64_t)D * H * W;
  const int cv = C / 8;
  const int pl = NORM_THREADS / cv > 0 ? NORM_THREADS / cv : 1;
  int64_t gx = (S + (int64_t)pl * 4 - 1) / ((int64_t)pl * 4);
  const int64_t cap = (8LL * DSK_NUM_SMS + B - 1) / B;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 ag((unsigned)gx, B);
  if (in_dtype == DSK_F16)
    DSK_LAUNCH((norm_apply_pad_kernel<__half, __half>), ag, NORM_THREADS, 0, as_stream(stream), (const __half*)x, (__half*)y_padded, table, D,
               H, W, C, ndim == 3 ? 1 : 0, silu);
  else
    DSK_LAUNCH((norm_apply_pad_kernel<__nv_bfloat16, __nv_bfloat16>), ag, NORM_THREADS, 0, as_stream(stream), (const __nv_bfloat16*)x,
               (__nv_bfloat16*)y_padded, table, D, H, W, C, ndim == 3 ? 1 : 0, silu);
  return DSK_OK;
}
