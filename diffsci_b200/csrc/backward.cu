// backward.cu -- K2 (bandwidth side): the backward passes of the norm / pooling / softmax / activation kernels and the
// channel reductions that give the bias and time-embedding gradients.  These are what loss.backward() reaches in
// KarrasModule.training_step (karras/karrasmodule.py:1146-1155) below the convolutions: autograd of
// torch.nn.GroupNorm / GroupRMSNorm + SiLU (nets/commonlayers.py:362-384, 824-831; nets/adm.py:305-329),
// MaxPool/AvgPool/Upsample (commonlayers.py:60-63,129; adm.py:361-381), softmax inside nn.MultiheadAttention
// (nets/attention.py:42-44) and the SiLU of the time MLPs (commonlayers.py:516-550).
//
// All kernels are HBM-bound streaming passes over channels-last tensors: 16-byte vectors, a thread owns a fixed channel
// vector so per-channel coefficients live in registers, reductions are two-stage (fp32 per thread, fp64 across chunks),
// deterministic, no atomics.
#include "common.cuh"

namespace dsk {

constexpr int BW_THREADS = 256;

template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int V = 4;
  static __device__ __forceinline__ void ld(const float* p, float* o) {
    float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  static __device__ __forceinline__ void st(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int V = 8;
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* o) {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[2 * j] = __low2float(h[j]); o[2 * j + 1] = __high2float(h[j]); }
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float* v) {
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) oh[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>(p) = o;
  }
};

// d/dz [z * sigmoid(z)]
__device__ __forceinline__ float dsilu_f(float z) {
  const float s = 1.0f / (1.0f + expf(-z));
  return s * (1.0f + z * (1.0f - s));
}
// bf16 tensors keep 8 mantissa bits: the approximate exp / divide (rel. error ~1e-6) is invisible after the rounding of the
// stored gradient and takes the norm-backward passes from instruction-bound back to HBM-bound (as silu_out in norm.cu).
template <typename T> __device__ __forceinline__ float dsilu_t(float z) { return dsilu_f(z); }
template <> __device__ __forceinline__ float dsilu_t<__nv_bfloat16>(float z) {
  const float s = __fdividef(1.0f, 1.0f + __expf(-z));
  return s * (1.0f + z * (1.0f - s));
}

// block size of norm_bwd_finalize_kernel: (channels of a group) x (chunk lanes), a power of two in 128 .. 1024
static inline int bwd_finalize_threads(int cg, int nchunks) {
  int t = 128;
  while (t < 1024 && t < (int64_t)cg * nchunks) t <<= 1;
  return t;
}
static inline int bw_chunks(int B, int64_t S, int C, int V) {
  int cv = C / V;
  int pl = BW_THREADS / cv;
  if (pl < 1) pl = 1;
  int64_t need = (S + (int64_t)pl * 64 - 1) / ((int64_t)pl * 64);     // <= 64 pixels per thread in fp32
  int64_t fill = (2 * DSK_NUM_SMS + B - 1) / B;
  int64_t by_work = (S + (int64_t)pl * 8 - 1) / ((int64_t)pl * 8);
  if (fill > by_work) fill = by_work;
  int64_t n = need > fill ? need : fill;
  if (n < 1) n = 1;
  return (int)n;
}

// ------------------------------------------------------------------------------------------------------------------
// Norm backward.  Forward (norm.cu): z = x*a + s with the folded per-(b,c) pair (a, s); y = silu(z) or z.
//   dz = dy * silu'(z);  per (b,c):  s1 = sum_S dz,  sx = sum_S dz*x.
// Pass 1 (this kernel) writes (s1, sx) per (b, chunk, c).  MODE_SUM: plain channel sums of dy (sx unused) -- the same
// reduction gives conv bias / time-embedding gradients.
template <typename T, bool NORM>
__global__ void __launch_bounds__(BW_THREADS) bwd_partial_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                  const float2* __restrict__ table, double2* __restrict__ partial,
                                                                  int64_t S, int C, int nchunks, int silu) {
  constexpr int V = Vec<T>::V;
  extern __shared__ float2 sm[];   // [pl][C]
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int cv = C / V;
  const int pl = max(1, BW_THREADS / cv);
  const int64_t per = (S + nchunks - 1) / nchunks;
  const int64_t s0 = (int64_t)chunk * per;
  const int64_t s1 = min(S, s0 + per);
  const T* xb = NORM ? x + (int64_t)b * S * C : nullptr;
  const T* gb = dy + (int64_t)b * S * C;
  for (int v = threadIdx.x; v < pl * cv; v += BW_THREADS) {
    const int lane = v / cv, c0 = (v - lane * cv) * V;
    float a[V], sh[V], acc1[V], accx[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      acc1[k] = 0; accx[k] = 0; a[k] = 1.0f; sh[k] = 0.0f;
      if (NORM) { const float2 t = table[(int64_t)b * C + c0 + k]; a[k] = t.x; sh[k] = t.y; }
    }
#pragma unroll 4
    for (int64_t s = s0 + lane; s < s1; s += pl) {
      float g[V], e[V];
      Vec<T>::ld(gb + s * C + c0, g);
      if (NORM) {
        Vec<T>::ld(xb + s * C + c0, e);
#pragma unroll
        for (int k = 0; k < V; ++k) {
          const float dz = silu ? g[k] * dsilu_t<T>(fmaf(e[k], a[k], sh[k])) : g[k];
          acc1[k] += dz;
          accx[k] = fmaf(dz, e[k], accx[k]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < V; ++k) acc1[k] += g[k];
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) sm[lane * C + c0 + k] = make_float2(acc1[k], accx[k]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += BW_THREADS) {
    double p = 0, q = 0;
    for (int l = 0; l < pl; ++l) {
      float2 v = sm[l * C + c];
      p += (double)v.x;
      q += (double)v.y;
    }
    partial[((int64_t)b * nchunks + chunk) * C + c] = make_double2(p, q);
  }
}

// One block per (sample, group): combine chunk partials; per-(b,c) sums S1 = sum dz, S2 = sum dz*xhat; group means of
// dxhat and dxhat*xhat (dxhat = dz*gamma*film); the apply-pass coefficients dx = dz*a + x*cb + cc with
//   cb = -rstd^2*m2,  cc = -rstd*m1 + mean*rstd^2*m2   (m1 = 0 for RMS);  FiLM gradients written directly.
__global__ void __launch_bounds__(1024) norm_bwd_finalize_kernel(const double2* __restrict__ partial, const float2* __restrict__ stats,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ fsc, float2* __restrict__ coef,
                                                                  float2* __restrict__ sums, float* __restrict__ dfsc,
                                                                  float* __restrict__ dfsh, int64_t S, int C, int G, int nchunks,
                                                                  int mode) {
  // block = cgp channel lanes (all channels of the group when they fit) x blockDim / cgp chunk lanes: the chunk partials of a
  // channel are summed by several lanes with every load of the block in flight at once (one thread per channel walking all chunks
  // was latency-bound: 8 - 15 us per launch with ADM's single group), combined through shared memory in a fixed order
  const int b = blockIdx.x / G, g = blockIdx.x % G;
  const int cg = C / G;
  const float2 mr = stats[blockIdx.x];
  const double mean = mr.x, rstd = mr.y;
  __shared__ double r1[1024], r2[1024];
  const int cgp = cg < (int)blockDim.x ? cg : (int)blockDim.x;   // channel lanes
  const int nrl = blockDim.x / cgp;                               // chunk lanes (>= 1)
  const int cl = threadIdx.x % cgp, rl = threadIdx.x / cgp;
  double a1 = 0, a2 = 0;
  for (int i0 = 0; i0 < cg; i0 += cgp) {
    const int i = i0 + cl;
    double s1 = 0, sx = 0;
    if (i < cg && rl < nrl) {
      const int c = g * cg + i;
#pragma unroll 4
      for (int ch = rl; ch < nchunks; ch += nrl) {
        const double2 v = partial[((int64_t)b * nchunks + ch) * C + c];
        s1 += v.x;
        sx += v.y;
      }
    }
    __syncthreads();                                    // r1 / r2 free again (previous channel block consumed)
    r1[threadIdx.x] = s1;
    r2[threadIdx.x] = sx;
    __syncthreads();
    if (rl == 0 && i < cg) {
      const int c = g * cg + i;
      s1 = 0; sx = 0;
      for (int k = 0; k < nrl; ++k) { s1 += r1[k * cgp + cl]; sx += r2[k * cgp + cl]; }
      const double s2 = rstd * (sx - mean * s1);
      sums[(int64_t)b * C + c] = make_float2((float)s1, (float)s2);
      const double gm = gamma != nullptr ? (double)gamma[c] : 1.0;
      const double f = fsc != nullptr ? (double)fsc[(int64_t)b * C + c] : 1.0;
      a1 += gm * f * s1;
      a2 += gm * f * s2;
      if (dfsc != nullptr) {
        const double bt = beta != nullptr ? (double)beta[c] : 0.0;
        dfsc[(int64_t)b * C + c] = (float)(gm * s2 + bt * s1);     // d/dfilm_scale of (xhat*gamma + beta)*fsc + fsh
        dfsh[(int64_t)b * C + c] = (float)s1;
      }
    }
  }
  __syncthreads();
  r1[threadIdx.x] = a1;                                 // non-zero in the first cgp threads only
  r2[threadIdx.x] = a2;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { r1[threadIdx.x] += r1[threadIdx.x + o]; r2[threadIdx.x] += r2[threadIdx.x + o]; }
    __syncthreads();
  }
  const double n = (double)S * cg;
  const double m1 = mode == 0 ? r1[0] / n : 0.0, m2 = r2[0] / n;
  const float cb = (float)(-rstd * rstd * m2);
  const float cc = (float)(-rstd * m1 + mean * rstd * rstd * m2);
  for (int i = threadIdx.x; i < cg; i += blockDim.x) coef[(int64_t)b * C + g * cg + i] = make_float2(cb, cc);
}

// G == C: block = 32 channels x 32 chunk lanes of one sample (see norm_finalize_pc_kernel), fixed-order combine.
__global__ void __launch_bounds__(1024) norm_bwd_finalize_pc_kernel(const double2* __restrict__ partial, const float2* __restrict__ stats,
                                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                     const float* __restrict__ fsc, float2* __restrict__ coef,
                                                                     float2* __restrict__ sums, float* __restrict__ dfsc,
                                                                     float* __restrict__ dfsh, int64_t S, int B, int C, int nchunks, int mode) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  double s1 = 0, sx = 0;
  if (c < C)
    for (int ch = rl; ch < nchunks; ch += 32) {
      const double2 v = partial[((int64_t)b * nchunks + ch) * C + c];
      s1 += v.x;
      sx += v.y;
    }
  __shared__ double r1[32][33], rx[32][33];
  r1[rl][threadIdx.x & 31] = s1;
  rx[rl][threadIdx.x & 31] = sx;
  __syncthreads();
  if (rl != 0 || c >= C) return;
  s1 = 0; sx = 0;
  for (int k = 0; k < 32; ++k) { s1 += r1[k][threadIdx.x & 31]; sx += rx[k][threadIdx.x & 31]; }
  const int i = b * C + c;
  const float2 mr = stats[i];
  const double mean = mr.x, rstd = mr.y;
  const double s2 = rstd * (sx - mean * s1);
  sums[i] = make_float2((float)s1, (float)s2);
  const double gm = gamma != nullptr ? (double)gamma[c] : 1.0;
  const double f = fsc != nullptr ? (double)fsc[i] : 1.0;
  if (dfsc != nullptr) {
    const double bt = beta != nullptr ? (double)beta[c] : 0.0;
    dfsc[i] = (float)(gm * s2 + bt * s1);
    dfsh[i] = (float)s1;
  }
  const double n = (double)S;
  const double m1 = mode == 0 ? gm * f * s1 / n : 0.0, m2 = gm * f * s2 / n;
  coef[i] = make_float2((float)(-rstd * rstd * m2), (float)(-rstd * m1 + mean * rstd * rstd * m2));
}

// G == C: one thread per (b, c) (the group reductions are over a single channel).
__global__ void __launch_bounds__(128) norm_bwd_finalize_pc_serial_kernel(const double2* __restrict__ partial, const float2* __restrict__ stats,
                                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                    const float* __restrict__ fsc, float2* __restrict__ coef,
                                                                    float2* __restrict__ sums, float* __restrict__ dfsc,
                                                                    float* __restrict__ dfsh, int64_t S, int B, int C, int nchunks, int mode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  const float2 mr = stats[i];
  const double mean = mr.x, rstd = mr.y;
  double s1 = 0, sx = 0;
  for (int ch = 0; ch < nchunks; ++ch) {
    const double2 v = partial[((int64_t)b * nchunks + ch) * C + c];
    s1 += v.x;
    sx += v.y;
  }
  const double s2 = rstd * (sx - mean * s1);
  sums[i] = make_float2((float)s1, (float)s2);
  const double gm = gamma != nullptr ? (double)gamma[c] : 1.0;
  const double f = fsc != nullptr ? (double)fsc[i] : 1.0;
  if (dfsc != nullptr) {
    const double bt = beta != nullptr ? (double)beta[c] : 0.0;
    dfsc[i] = (float)(gm * s2 + bt * s1);
    dfsh[i] = (float)s1;
  }
  const double n = (double)S;
  const double m1 = mode == 0 ? gm * f * s1 / n : 0.0, m2 = gm * f * s2 / n;
  coef[i] = make_float2((float)(-rstd * rstd * m2), (float)(-rstd * m1 + mean * rstd * rstd * m2));
}

// dgamma[c] = sum_b film[b,c]*S2[b,c] ; dbeta[c] = sum_b film[b,c]*S1[b,c]
__global__ void __launch_bounds__(256) norm_bwd_param_kernel(const float2* __restrict__ sums, const float* __restrict__ fsc,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int C) {
  // block = 32 channels x 8 sample lanes
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  double g = 0, bt = 0;
  if (c < C)
    for (int b = rl; b < B; b += 8) {
      const float2 v = sums[(int64_t)b * C + c];
      const double f = fsc != nullptr ? (double)fsc[(int64_t)b * C + c] : 1.0;
      g += f * v.y;
      bt += f * v.x;
    }
  __shared__ double rg[8][33], rb[8][33];
  rg[rl][threadIdx.x & 31] = g;
  rb[rl][threadIdx.x & 31] = bt;
  __syncthreads();
  if (rl == 0 && c < C) {
    double sg = 0, sb = 0;
    for (int k = 0; k < 8; ++k) { sg += rg[k][threadIdx.x & 31]; sb += rb[k][threadIdx.x & 31]; }
    dgamma[c] = (float)sg;
    dbeta[c] = (float)sb;
  }
}

// dx = dz*a + x*cb + cc (+ dres);  dres may alias dx.
template <typename T>
__global__ void __launch_bounds__(BW_THREADS) norm_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy, const T* dres,
                                                                     T* dx, const float2* __restrict__ table,
                                                                     const float2* __restrict__ coef, int64_t S, int C, int silu) {
  constexpr int V = Vec<T>::V;
  const int b = blockIdx.y;
  const int cv = C / V;
  const int pl = max(1, BW_THREADS / cv);
  const T* xb = x + (int64_t)b * S * C;
  const T* gb = dy + (int64_t)b * S * C;
  const T* rb = dres != nullptr ? dres + (int64_t)b * S * C : nullptr;
  T* ob = dx + (int64_t)b * S * C;
  for (int v = threadIdx.x; v < pl * cv; v += BW_THREADS) {
    const int lane = v / cv, c0 = (v - lane * cv) * V;
    float a[V], sh[V], cb[V], cc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float2 t = table[(int64_t)b * C + c0 + k], q = coef[(int64_t)b * C + c0 + k];
      a[k] = t.x; sh[k] = t.y; cb[k] = q.x; cc[k] = q.y;
    }
    const int64_t step = (int64_t)gridDim.x * pl;
#pragma unroll 4
    for (int64_t s = (int64_t)blockIdx.x * pl + lane; s < S; s += step) {
      float e[V], g[V], r[V], o[V];
      Vec<T>::ld(xb + s * C + c0, e);
      Vec<T>::ld(gb + s * C + c0, g);
      if (rb != nullptr) Vec<T>::ld(rb + s * C + c0, r);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float dz = silu ? g[k] * dsilu_t<T>(fmaf(e[k], a[k], sh[k])) : g[k];
        float t = fmaf(dz, a[k], fmaf(e[k], cb[k], cc[k]));
        if (rb != nullptr) t += r[k];
        o[k] = t;
      }
      Vec<T>::st(ob + s * C + c0, o);
    }
  }
}

// channel sums: out[b][c] (per_sample) or out[c] = sum over chunk partials (and samples)
__global__ void __launch_bounds__(1024) chansum_finalize_kernel(const double2* __restrict__ partial, float* __restrict__ out, int B,
                                                                int C, int nchunks, int per_sample) {
  // block = 8 channels x 128 row lanes over the (sample, chunk) partial rows: the kernel is a chain of dependent L2 loads on a
  // handful of blocks (32 channels x 32 row lanes: 9 us per launch, 45 launches per ADM iteration), so rows are spread over as
  // many lanes and blocks as the shape has; fixed-order combine through shared memory
  const int c = blockIdx.x * 8 + (threadIdx.x & 7), rl = threadIdx.x >> 3;
  const int64_t r0 = per_sample ? (int64_t)blockIdx.y * nchunks : 0;
  const int64_t rows = per_sample ? nchunks : (int64_t)B * nchunks;
  double s = 0;
  if (c < C) {
#pragma unroll 4
    for (int64_t r = rl; r < rows; r += 128) s += partial[(r0 + r) * C + c].x;
  }
  __shared__ double red[128][9];
  red[rl][threadIdx.x & 7] = s;
  __syncthreads();
  const int ch = threadIdx.x & 7;
  double t = 0;
  if (rl < 8)                                            // 8 lanes per channel add 16 rows each ...
    for (int k = 0; k < 16; ++k) t += red[rl * 16 + k][ch];
  __syncthreads();
  if (rl < 8) red[rl][ch] = t;                           // ... and lane 0 adds the 8 partial sums
  __syncthreads();
  if (rl == 0 && c < C) {
    t = 0;
    for (int k = 0; k < 8; ++k) t += red[k][ch];
    out[per_sample ? (int64_t)blockIdx.y * C + c : c] = (float)t;
  }
}

// generic scalar fallback for channel counts that are not a multiple of the vector width (convout: C = 1..4)
constexpr int CS_SPLITS = 64;
template <typename T>
__global__ void __launch_bounds__(256) chansum_small_kernel(const T* __restrict__ dy, float* __restrict__ out, int B, int64_t S, int C,
                                                             int per_sample, double* __restrict__ part) {
  // one block per (b or all, c[, slice of the pixels]): with `part` the pixel range is cut into gridDim.z slices whose sums
  // chansum_small_finalize_kernel adds in a fixed order (a single block per channel took 267 us on a 2 x 64^3 volume)
  // the slices cut the pixel axis of ONE sample (per_sample) or the flattened (sample, pixel) axis (sum over the batch too: many
  // small samples, e.g. 32 x 128^2, are then spread over all slices instead of 3 blocks walking the whole tensor: 289 -> ~10 us)
  const int c = blockIdx.x;
  const int64_t base = per_sample ? (int64_t)blockIdx.y * S : 0, n = per_sample ? S : (int64_t)B * S;
  const int64_t per = (n + gridDim.z - 1) / gridDim.z;
  const int64_t sa = (int64_t)blockIdx.z * per, sb = min(n, sa + per);
  double acc = 0;
  for (int64_t s0 = sa; s0 < sb; s0 += 16 * 256) {          // fp32 within a run of 16 elements per thread, fp64 across runs
    float local = 0.0f;
    for (int64_t s = s0 + threadIdx.x; s < min(sb, s0 + 16 * 256); s += 256) local += to_f32<T>(dy[(base + s) * C + c]);
    acc += (double)local;
  }
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  const int64_t o = per_sample ? (int64_t)blockIdx.y * C + c : c;
  if (part != nullptr) part[o * gridDim.z + blockIdx.z] = red[0];
  else out[o] = (float)red[0];
}

__global__ void __launch_bounds__(128) chansum_small_finalize_kernel(const double* __restrict__ part, float* __restrict__ out, int n, int splits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0;
  for (int z = 0; z < splits; ++z) s += part[(int64_t)i * splits + z];
  out[i] = (float)s;
}

// out[c] = sum_r in[r][c]  (fp32, small matrices: bias gradients of the time-MLP linears and attention projections)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t rows, int cols, int ld) {
  // block = 32 columns x 8 row-lanes
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  double acc = 0;
  if (c < cols)
    for (int64_t r = rl; r < rows; r += 8) acc += (double)in[r * ld + c];
  __shared__ double red[8][33];
  red[rl][threadIdx.x & 31] = acc;
  __syncthreads();
  if (rl == 0 && c < cols) {
    double s = 0;
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x & 31];
    out[c] = (float)s;
  }
}

// ---- pooling / upsampling backward ---------------------------------------------------------------------------------
// MaxPool(2) backward: the gradient of a window goes to its FIRST maximal element in (d, h, w) scan order (ATen's rule);
// AvgPool(2): dy / 2^ndim to every element.  One thread per pooled element; writes the whole window (+ dres).
template <typename T, bool IS_MAX>
__global__ void __launch_bounds__(256) pool2x_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, const T* dres, T* dx, int B,
                                                          int D, int H, int W, int C, int ndim) {
  const int Do = ndim == 3 ? D / 2 : 1, Ho = H / 2, Wo = W / 2;
  const int kd = ndim == 3 ? 2 : 1;
  const int64_t total = (int64_t)B * Do * Ho * Wo * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t p = i / C;
    int wo = (int)(p % Wo); p /= Wo;
    int ho = (int)(p % Ho); p /= Ho;
    int dz = (int)(p % Do);
    int b = (int)(p / Do);
    const float g = to_f32<T>(dy[i]);
    int best = 0;
    if (IS_MAX) {
      float m = -INFINITY;
      int idx = 0;
      for (int a = 0; a < kd; ++a)
        for (int bb = 0; bb < 2; ++bb)
          for (int cc = 0; cc < 2; ++cc, ++idx) {
            const int64_t src = ((((int64_t)b * D + (dz * kd + a)) * H + (ho * 2 + bb)) * W + (wo * 2 + cc)) * C + c;
            const float v = to_f32<T>(x[src]);
            if (v > m || (v != v && !(m != m))) { m = v; best = idx; }   // strict '>' keeps the first maximum; NaN propagates like ATen
          }
    }
    const float share = ndim == 3 ? 0.125f : 0.25f;
    int idx = 0;
    for (int a = 0; a < kd; ++a)
      for (int bb = 0; bb < 2; ++bb)
        for (int cc = 0; cc < 2; ++cc, ++idx) {
          const int64_t dst = ((((int64_t)b * D + (dz * kd + a)) * H + (ho * 2 + bb)) * W + (wo * 2 + cc)) * C + c;
          float v = IS_MAX ? (idx == best ? g : 0.0f) : g * share;
          if (dres != nullptr) v += to_f32<T>(dres[dst]);
          dx[dst] = from_f32<T>(v);
        }
  }
}

// odd trailing rows / columns that no pooling window covers (floor semantics) get zero gradient (+ dres)
template <typename T>
__global__ void __launch_bounds__(256) pool2x_bwd_tail_kernel(const T* dres, T* dx, int B, int D, int H, int W, int C, int ndim) {
  const int De = ndim == 3 ? (D / 2) * 2 : D, He = (H / 2) * 2, We = (W / 2) * 2;
  const int64_t total = (int64_t)B * D * H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / C;
    int w = (int)(p % W); p /= W;
    int h = (int)(p % H); p /= H;
    int d = (int)(p % D);
    if (w < We && h < He && d < De) continue;
    dx[i] = dres != nullptr ? dres[i] : from_f32<T>(0.0f);
  }
}

// nearest x2 upsample backward: dx[p] = sum of the 2^ndim children of dy (+ dres)
template <typename T>
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const T* __restrict__ dy, const T* dres, T* dx, int B, int D, int H, int W,
                                                              int C, int ndim) {
  // D, H, W: INPUT (low-resolution) size
  const int kd = ndim == 3 ? 2 : 1;
  const int Do = D * kd, Ho = H * 2, Wo = W * 2;
  const int64_t total = (int64_t)B * D * H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t p = i / C;
    int w = (int)(p % W); p /= W;
    int h = (int)(p % H); p /= H;
    int d = (int)(p % D);
    int b = (int)(p / D);
    float acc = 0.0f;
    for (int a = 0; a < kd; ++a)
      for (int bb = 0; bb < 2; ++bb)
        for (int cc = 0; cc < 2; ++cc)
          acc += to_f32<T>(dy[((((int64_t)b * Do + (d * kd + a)) * Ho + (h * 2 + bb)) * Wo + (w * 2 + cc)) * C + c]);
    if (dres != nullptr) acc += to_f32<T>(dres[i]);
    dx[i] = from_f32<T>(acc);
  }
}

// ---- 16-byte (8 x bf16) forms of the pooling / upsampling backward passes: one thread per (pixel, 8-channel chunk) ----
__device__ __forceinline__ void bf8_unpack(const uint4& r, float* o) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int j = 0; j < 4; ++j) { o[2 * j] = __low2float(h[j]); o[2 * j + 1] = __high2float(h[j]); }
}
__device__ __forceinline__ uint4 bf8_pack(const float* v) {
  uint4 o;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
  return o;
}

__global__ void __launch_bounds__(256) upsample2x_bwd_vec8_kernel(const uint4* __restrict__ dy, const uint4* dres, uint4* dx, int B, int D,
                                                                   int H, int W, int Cv, int ndim) {
  const int kd = ndim == 3 ? 2 : 1;
  const int Do = D * kd, Ho = H * 2, Wo = W * 2;
  const int64_t total = (int64_t)B * D * H * W * Cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cv);
    int64_t p = i / Cv;
    const int w = (int)(p % W); p /= W;
    const int h = (int)(p % H); p /= H;
    const int d = (int)(p % D);
    const int b = (int)(p / D);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t[8];
    for (int a = 0; a < kd; ++a)
#pragma unroll
      for (int bb = 0; bb < 2; ++bb)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {   // same summation order as the scalar kernel
          bf8_unpack(dy[((((int64_t)b * Do + (d * kd + a)) * Ho + (h * 2 + bb)) * Wo + (w * 2 + cc)) * Cv + c], t);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += t[k];
        }
    if (dres != nullptr) {
      bf8_unpack(dres[i], t);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += t[k];
    }
    dx[i] = bf8_pack(acc);
  }
}

template <bool IS_MAX>
__global__ void __launch_bounds__(256) pool2x_bwd_vec8_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, const uint4* dres,
                                                               uint4* dx, int B, int D, int H, int W, int Cv, int ndim) {
  const int Do = ndim == 3 ? D / 2 : 1, Ho = H / 2, Wo = W / 2;
  const int kd = ndim == 3 ? 2 : 1;
  const int64_t total = (int64_t)B * Do * Ho * Wo * Cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cv);
    int64_t p = i / Cv;
    const int wo = (int)(p % Wo); p /= Wo;
    const int ho = (int)(p % Ho); p /= Ho;
    const int dz = (int)(p % Do);
    const int b = (int)(p / Do);
    float g[8];
    bf8_unpack(dy[i], g);
    int best[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (IS_MAX) {
      float m[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
      int idx = 0;
      for (int a = 0; a < kd; ++a)
        for (int bb = 0; bb < 2; ++bb)
          for (int cc = 0; cc < 2; ++cc, ++idx) {
            float v[8];
            bf8_unpack(x[((((int64_t)b * D + (dz * kd + a)) * H + (ho * 2 + bb)) * W + (wo * 2 + cc)) * Cv + c], v);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (v[k] > m[k] || (v[k] != v[k] && !(m[k] != m[k]))) { m[k] = v[k]; best[k] = idx; }
          }
    }
    const float share = ndim == 3 ? 0.125f : 0.25f;
    int idx = 0;
    for (int a = 0; a < kd; ++a)
      for (int bb = 0; bb < 2; ++bb)
        for (int cc = 0; cc < 2; ++cc, ++idx) {
          const int64_t dst = ((((int64_t)b * D + (dz * kd + a)) * H + (ho * 2 + bb)) * W + (wo * 2 + cc)) * Cv + c;
          float v[8], r[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = IS_MAX ? (idx == best[k] ? g[k] : 0.0f) : g[k] * share;
          if (dres != nullptr) {
            bf8_unpack(dres[dst], r);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] += r[k];
          }
          dx[dst] = bf8_pack(v);
        }
  }
}

// ---- softmax backward (rows): dS = P * (dP - sum_j dP_j P_j), in place on dP ------------------------------------------
__global__ void __launch_bounds__(256) softmax_bwd_rows_kernel(const float* __restrict__ P, float* __restrict__ dP, int64_t rows,
                                                                int cols) {
  __shared__ float red[8];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* p = P + r * cols;
    float* g = dP + r * cols;
    float dot = 0.0f;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) dot = fmaf(p[c], g[c], dot);
    dot = warp_sum(dot);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    dot = 0.0f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) dot += red[w];
    __syncthreads();
    for (int c = threadIdx.x; c < cols; c += blockDim.x) g[c] = p[c] * (g[c] - dot);
  }
}

// ---- SiLU forward / backward on fp32 vectors (time MLPs) --------------------------------------------------------------
__global__ void __launch_bounds__(256) silu_fwd_kernel(const float* __restrict__ z, float* __restrict__ a, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] = silu_f(z[i]);
}
__global__ void __launch_bounds__(256) silu_bwd_kernel(const float* __restrict__ z, const float* __restrict__ da, float* __restrict__ dz,
                                                        int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dz[i] = da[i] * dsilu_f(z[i]);
}

__global__ void __launch_bounds__(256) relu_fwd_kernel(const float* __restrict__ z, float* __restrict__ a, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] = fmaxf(z[i], 0.0f);
}
__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ z, const float* __restrict__ da, float* __restrict__ dz,
                                                        int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dz[i] = z[i] > 0.0f ? da[i] : 0.0f;      // torch.relu: zero gradient at z == 0
}

// ---- y = a (+ b), each operand with its own dtype (gradient accumulation / casts between fp32 and bf16 buffers) --------
template <typename TA, typename TB, typename TO>
__global__ void __launch_bounds__(256) add_ex_kernel(const TA* __restrict__ a, const TB* b, TO* y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = to_f32<TA>(a[i]);
    if (b != nullptr) v += to_f32<TB>(b[i]);
    y[i] = from_f32<TO>(v);
  }
}

// ---- channel split (backward of the ADM concat skip, adm.py:296-297): da = dy[:, :Ca] (+ ra), db = dy[:, Ca:] (+ rb) -----
template <typename T>
__global__ void __launch_bounds__(256) split_channels_kernel(const T* __restrict__ dy, const T* ra, const T* rb, T* da, T* db,
                                                              int64_t rows, int Ca, int Cb) {
  const int Ct = Ca + Cb;
  const int64_t total = rows * Ct;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % Ct);
    int64_t r = i / Ct;
    float v = to_f32<T>(dy[i]);
    if (c < Ca) {
      if (da == nullptr) continue;
      const int64_t o = r * Ca + c;
      if (ra != nullptr) v += to_f32<T>(ra[o]);
      da[o] = from_f32<T>(v);
    } else {
      if (db == nullptr) continue;
      const int64_t o = r * Cb + (c - Ca);
      if (rb != nullptr) v += to_f32<T>(rb[o]);
      db[o] = from_f32<T>(v);
    }
  }
}

}  // namespace dsk

using namespace dsk;

static inline int vec_of(int dtype) { return dtype == DSK_BF16 ? 8 : 4; }

extern "C" int64_t dsk_bwd_ws_bytes(int B, int64_t S, int C) {
  if (B <= 0 || S <= 0 || C <= 0) return 0;
  int nchunks = 1;
  if (C % 4 == 0) nchunks = bw_chunks(B, S, C, 4);
  if (C % 8 == 0) { const int n8 = bw_chunks(B, S, C, 8); if (n8 > nchunks) nchunks = n8; }
  // [coef B*C float2][sums B*C float2][partials B*nchunks*C double2]   (few-channel sums: B*C*CS_SPLITS doubles)
  const int64_t part = (int64_t)B * nchunks * C * (int64_t)sizeof(double2), small = (int64_t)B * C * CS_SPLITS * (int64_t)sizeof(double);
  return 2 * (int64_t)B * C * (int64_t)sizeof(float2) + (part > small ? part : small);
}

extern "C" int dsk_norm_act_bwd(const void* x, const void* dy, const void* dres, void* dx, const float* gamma, const float* beta,
                                const float* film_scale, const void* fwd_ws, float* dgamma, float* dbeta, float* dfilm_scale,
                                float* dfilm_shift, void* ws, int B, int64_t S, int C, int G, int mode, int silu, int dtype,
                                void* stream) {
  DSK_REQUIRE(x && dy && dx && fwd_ws && ws, "dsk_norm_act_bwd: null pointer");
  DSK_REQUIRE(B > 0 && B <= 65535 && S > 0 && C > 0 && G > 0 && C % G == 0, "dsk_norm_act_bwd: bad shape");
  DSK_REQUIRE(dtype == DSK_F32 || dtype == DSK_BF16, "dsk_norm_act_bwd: bad dtype %d", dtype);
  const int V = vec_of(dtype);
  DSK_REQUIRE(C % V == 0, "dsk_norm_act_bwd: C=%d must be a multiple of %d for this dtype", C, V);
  DSK_REQUIRE((gamma == nullptr) == (dgamma == nullptr) && (gamma == nullptr) == (dbeta == nullptr) && (gamma == nullptr) == (beta == nullptr),
              "dsk_norm_act_bwd: gamma/beta/dgamma/dbeta must be all set or all null");
  DSK_REQUIRE((film_scale == nullptr) == (dfilm_scale == nullptr) && (film_scale == nullptr) == (dfilm_shift == nullptr),
              "dsk_norm_act_bwd: FiLM pointers must be all set or all null");
  cudaStream_t st = as_stream(stream);
  const float2* table = reinterpret_cast<const float2*>(fwd_ws);
  const float2* stats = table + (int64_t)B * C;
  float2* coef = reinterpret_cast<float2*>(ws);
  float2* sums = coef + (int64_t)B * C;
  double2* partial = reinterpret_cast<double2*>(sums + (int64_t)B * C);
  const int nchunks = bw_chunks(B, S, C, V);
  const int cv = C / V;
  const int pl = BW_THREADS / cv > 0 ? BW_THREADS / cv : 1;
  const size_t smem = (size_t)pl * C * sizeof(float2);
  DSK_REQUIRE(smem <= 48 * 1024, "dsk_norm_act_bwd: C=%d too large", C);
  dim3 pg(nchunks, B);
  if (dtype == DSK_F32)
    DSK_LAUNCH((bwd_partial_kernel<float, true>), pg, BW_THREADS, smem, st, (const float*)x, (const float*)dy, table, partial, S, C, nchunks, silu);
  else
    DSK_LAUNCH((bwd_partial_kernel<__nv_bfloat16, true>), pg, BW_THREADS, smem, st, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, table,
               partial, S, C, nchunks, silu);
  if (G == C)
  {
    if (nchunks >= 8)
      DSK_LAUNCH(norm_bwd_finalize_pc_kernel, dim3((C + 31) / 32, B), 1024, 0, st, partial, stats, gamma, beta, film_scale, coef, sums,
                 dfilm_scale, dfilm_shift, S, B, C, nchunks, mode);
    else
      DSK_LAUNCH(norm_bwd_finalize_pc_serial_kernel, (B * C + 127) / 128, 128, 0, st, partial, stats, gamma, beta, film_scale, coef, sums,
                 dfilm_scale, dfilm_shift, S, B, C, nchunks, mode);
  }
  else
    DSK_LAUNCH(norm_bwd_finalize_kernel, B * G, bwd_finalize_threads(C / G, nchunks), 0, st, partial, stats, gamma, beta, film_scale, coef, sums, dfilm_scale, dfilm_shift, S,
               C, G, nchunks, mode);
  if (dgamma != nullptr) DSK_LAUNCH(norm_bwd_param_kernel, (C + 31) / 32, 256, 0, st, sums, film_scale, dgamma, dbeta, B, C);
  int64_t gx = (S + (int64_t)pl * 4 - 1) / ((int64_t)pl * 4);
  const int64_t cap = (8LL * DSK_NUM_SMS + B - 1) / B;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 ag((unsigned)gx, B);
  if (dtype == DSK_F32)
    DSK_LAUNCH(norm_bwd_apply_kernel<float>, ag, BW_THREADS, 0, st, (const float*)x, (const float*)dy, (const float*)dres, (float*)dx, table,
               coef, S, C, silu);
  else
    DSK_LAUNCH(norm_bwd_apply_kernel<__nv_bfloat16>, ag, BW_THREADS, 0, st, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy,
               (const __nv_bfloat16*)dres, (__nv_bfloat16*)dx, table, coef, S, C, silu);
  return DSK_OK;
}

extern "C" int dsk_channel_sum(const void* dy, float* out, void* ws, int B, int64_t S, int C, int dtype, int per_sample, void* stream) {
  DSK_REQUIRE(dy && out && B > 0 && B <= 65535 && S > 0 && C > 0, "dsk_channel_sum: bad arguments");
  DSK_REQUIRE(dtype == DSK_F32 || dtype == DSK_BF16, "dsk_channel_sum: bad dtype %d", dtype);
  cudaStream_t st = as_stream(stream);
  const int V = vec_of(dtype);
  if (C % V != 0) {
    DSK_REQUIRE(C <= 65535, "dsk_channel_sum: C too large for the scalar path");
    const int rows = per_sample ? B : 1;
    const bool split = ws != nullptr && (per_sample ? S : (int64_t)B * S) >= 65536 && (int64_t)rows * C <= 4096;
    double* part = split ? reinterpret_cast<double*>(reinterpret_cast<float2*>(ws) + 2 * (int64_t)B * C) : nullptr;
    dim3 g(C, rows, split ? CS_SPLITS : 1);
    if (dtype == DSK_F32) DSK_LAUNCH(chansum_small_kernel<float>, g, 256, 0, st, (const float*)dy, out, B, S, C, per_sample, part);
    else DSK_LAUNCH(chansum_small_kernel<__nv_bfloat16>, g, 256, 0, st, (const __nv_bfloat16*)dy, out, B, S, C, per_sample, part);
    if (split) DSK_LAUNCH(chansum_small_finalize_kernel, (rows * C + 127) / 128, 128, 0, st, part, out, rows * C, CS_SPLITS);
    return DSK_OK;
  }
  DSK_REQUIRE(ws != nullptr, "dsk_channel_sum: workspace required");
  double2* partial = reinterpret_cast<double2*>(reinterpret_cast<float2*>(ws) + 2 * (int64_t)B * C);
  const int nchunks = bw_chunks(B, S, C, V);
  const int cv = C / V;
  const int pl = BW_THREADS / cv > 0 ? BW_THREADS / cv : 1;
  const size_t smem = (size_t)pl * C * sizeof(float2);
  DSK_REQUIRE(smem <= 48 * 1024, "dsk_channel_sum: C=%d too large", C);
  dim3 pg(nchunks, B);
  if (dtype == DSK_F32)
    DSK_LAUNCH((bwd_partial_kernel<float, false>), pg, BW_THREADS, smem, st, nullptr, (const float*)dy, nullptr, partial, S, C, nchunks, 0);
  else
    DSK_LAUNCH((bwd_partial_kernel<__nv_bfloat16, false>), pg, BW_THREADS, smem, st, nullptr, (const __nv_bfloat16*)dy, nullptr, partial, S, C,
               nchunks, 0);
  dim3 fg((C + 7) / 8, per_sample ? B : 1);
  DSK_LAUNCH(chansum_finalize_kernel, fg, 1024, 0, st, partial, out, B, C, nchunks, per_sample);
  return DSK_OK;
}

extern "C" int dsk_colsum_f32(const float* in, float* out, int64_t rows, int cols, int ld, void* stream) {
  DSK_REQUIRE(in && out && rows > 0 && cols > 0 && ld >= cols, "dsk_colsum_f32: bad arguments");
  DSK_LAUNCH(colsum_kernel, (cols + 31) / 32, 256, 0, as_stream(stream), in, out, rows, cols, ld);
  return DSK_OK;
}

extern "C" int dsk_pool2x_bwd(const void* x, const void* dy, const void* dres, void* dx, int B, int D, int H, int W, int C, int ndim,
                              int is_max, int dtype, void* stream) {
  DSK_REQUIRE(dy && dx && (x || !is_max), "dsk_pool2x_bwd: null pointer");
  DSK_REQUIRE(B > 0 && D > 0 && H > 1 && W > 1 && C > 0 && (ndim == 2 || ndim == 3), "dsk_pool2x_bwd: bad shape");
  DSK_REQUIRE(ndim == 2 ? D == 1 : D > 1, "dsk_pool2x_bwd: D=%d inconsistent with ndim=%d", D, ndim);
  const int64_t total = (int64_t)B * (ndim == 3 ? D / 2 : 1) * (H / 2) * (W / 2) * C;
  const int grid = grid_for(total, 256, 16);
  cudaStream_t st = as_stream(stream);
#define PB(T, M) DSK_LAUNCH((pool2x_bwd_kernel<T, M>), grid, 256, 0, st, (const T*)x, (const T*)dy, (const T*)dres, (T*)dx, B, D, H, W, C, ndim)
  const bool al16 = (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)dres) & 15) == 0;
  if (dtype == DSK_BF16 && C % 8 == 0 && al16) {
    const int vg = grid_for((int64_t)B * (ndim == 3 ? D / 2 : 1) * (H / 2) * (W / 2) * (C / 8), 256, 16);
    if (is_max) DSK_LAUNCH(pool2x_bwd_vec8_kernel<true>, vg, 256, 0, st, (const uint4*)x, (const uint4*)dy, (const uint4*)dres, (uint4*)dx, B, D,
                           H, W, C / 8, ndim);
    else DSK_LAUNCH(pool2x_bwd_vec8_kernel<false>, vg, 256, 0, st, (const uint4*)x, (const uint4*)dy, (const uint4*)dres, (uint4*)dx, B, D, H, W,
                    C / 8, ndim);
  } else if (dtype == DSK_F32) { if (is_max) PB(float, true); else PB(float, false); }
  else if (dtype == DSK_BF16) { if (is_max) PB(__nv_bfloat16, true); else PB(__nv_bfloat16, false); }
  else DSK_REQUIRE(false, "dsk_pool2x_bwd: bad dtype %d", dtype);
#undef PB
  if ((H & 1) || (W & 1) || (ndim == 3 && (D & 1))) {
    const int g2 = grid_for((int64_t)B * D * H * W * C, 256, 16);
    if (dtype == DSK_F32) DSK_LAUNCH(pool2x_bwd_tail_kernel<float>, g2, 256, 0, st, (const float*)dres, (float*)dx, B, D, H, W, C, ndim);
    else DSK_LAUNCH(pool2x_bwd_tail_kernel<__nv_bfloat16>, g2, 256, 0, st, (const __nv_bfloat16*)dres, (__nv_bfloat16*)dx, B, D, H, W, C, ndim);
  }
  return DSK_OK;
}

extern "C" int dsk_upsample2x_bwd(const void* dy, const void* dres, void* dx, int B, int D, int H, int W, int C, int ndim, int dtype,
                                  void* stream) {
  DSK_REQUIRE(dy && dx && B > 0 && D > 0 && H > 0 && W > 0 && C > 0 && (ndim == 2 || ndim == 3), "dsk_upsample2x_bwd: bad arguments");
  const int grid = grid_for((int64_t)B * D * H * W * C, 256, 16);
  cudaStream_t st = as_stream(stream);
  if (dtype == DSK_BF16 && C % 8 == 0 && ((((uintptr_t)dy | (uintptr_t)dx | (uintptr_t)dres) & 15) == 0))
    DSK_LAUNCH(upsample2x_bwd_vec8_kernel, grid_for((int64_t)B * D * H * W * (C / 8), 256, 16), 256, 0, st, (const uint4*)dy, (const uint4*)dres,
               (uint4*)dx, B, D, H, W, C / 8, ndim);
  else if (dtype == DSK_F32) DSK_LAUNCH(upsample2x_bwd_kernel<float>, grid, 256, 0, st, (const float*)dy, (const float*)dres, (float*)dx, B, D, H, W, C, ndim);
  else if (dtype == DSK_BF16)
    DSK_LAUNCH(upsample2x_bwd_kernel<__nv_bfloat16>, grid, 256, 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)dres, (__nv_bfloat16*)dx, B,
               D, H, W, C, ndim);
  else DSK_REQUIRE(false, "dsk_upsample2x_bwd: bad dtype %d", dtype);
  return DSK_OK;
}

extern "C" int dsk_softmax_bwd_rows(const float* P, float* dP, int64_t rows, int cols, void* stream) {
  DSK_REQUIRE(P && dP && rows > 0 && cols > 0, "dsk_softmax_bwd_rows: bad arguments");
  int64_t grid = rows < (int64_t)DSK_NUM_SMS * 16 ? rows : (int64_t)DSK_NUM_SMS * 16;
  DSK_LAUNCH(softmax_bwd_rows_kernel, (int)grid, 256, 0, as_stream(stream), P, dP, rows, cols);
  return DSK_OK;
}

extern "C" int dsk_silu_fwd(const float* z, float* a, int64_t n, void* stream) {
  DSK_REQUIRE(z && a && n > 0, "dsk_silu_fwd: bad arguments");
  DSK_LAUNCH(silu_fwd_kernel, grid_for(n, 256, 8), 256, 0, as_stream(stream), z, a, n);
  return DSK_OK;
}

extern "C" int dsk_silu_bwd(const float* z, const float* da, float* dz, int64_t n, void* stream) {
  DSK_REQUIRE(z && da && dz && n > 0, "dsk_silu_bwd: bad arguments");
  DSK_LAUNCH(silu_bwd_kernel, grid_for(n, 256, 8), 256, 0, as_stream(stream), z, da, dz, n);
  return DSK_OK;
}

extern "C" int dsk_relu_fwd(const float* z, float* a, int64_t n, void* stream) {
  DSK_REQUIRE(z && a && n > 0, "dsk_relu_fwd: bad arguments");
  DSK_LAUNCH(relu_fwd_kernel, grid_for(n, 256, 8), 256, 0, as_stream(stream), z, a, n);
  return DSK_OK;
}

extern "C" int dsk_relu_bwd(const float* z, const float* da, float* dz, int64_t n, void* stream) {
  DSK_REQUIRE(z && da && dz && n > 0, "dsk_relu_bwd: bad arguments");
  DSK_LAUNCH(relu_bwd_kernel, grid_for(n, 256, 8), 256, 0, as_stream(stream), z, da, dz, n);
  return DSK_OK;
}

extern "C" int dsk_add_ex(const void* a, int a_dtype, const void* b, int b_dtype, void* y, int y_dtype, int64_t n, void* stream) {
  DSK_REQUIRE(a && y && n > 0, "dsk_add_ex: bad arguments");
  const int grid = grid_for(n, 256, 16);
  cudaStream_t st = as_stream(stream);
  if (b == nullptr) b_dtype = a_dtype;
  const int key = a_dtype * 4 + b_dtype * 2 + y_dtype;
  typedef __nv_bfloat16 bf;
#define AX(TA, TB, TO) DSK_LAUNCH((add_ex_kernel<TA, TB, TO>), grid, 256, 0, st, (const TA*)a, (const TB*)b, (TO*)y, n)
  switch (key) {
    case 0: AX(float, float, float); break;
    case 1: AX(float, float, bf); break;
    case 2: AX(float, bf, float); break;
    case 3: AX(float, bf, bf); break;
    case 4: AX(bf, float, float); break;
    case 5: AX(bf, float, bf); break;
    case 6: AX(bf, bf, float); break;
    case 7: AX(bf, bf, bf); break;
    default: DSK_REQUIRE(false, "dsk_add_ex: bad dtypes %d %d %d", a_dtype, b_dtype, y_dtype);
  }
#undef AX
  return DSK_OK;
}

extern "C" int dsk_split_channels(const void* dy, const void* ra, const void* rb, void* da, void* db, int64_t rows, int Ca, int Cb,
                                  int dtype, void* stream) {
  DSK_REQUIRE(dy && (da || db) && rows > 0 && Ca > 0 && Cb > 0, "dsk_split_channels: bad arguments");
  const int grid = grid_for(rows * (Ca + Cb), 256, 16);
  cudaStream_t st = as_stream(stream);
  if (dtype == DSK_F32)
    DSK_LAUNCH(split_channels_kernel<float>, grid, 256, 0, st, (const float*)dy, (const float*)ra, (const float*)rb, (float*)da, (float*)db, rows,
               Ca, Cb);
  else if (dtype == DSK_BF16)
    DSK_LAUNCH(split_channels_kernel<__nv_bfloat16>, grid, 256, 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)ra,
               (const __nv_bfloat16*)rb, (__nv_bfloat16*)da, (__nv_bfloat16*)db, rows, Ca, Cb);
  else DSK_REQUIRE(false, "dsk_split_channels: bad dtype %d", dtype);
  return DSK_OK;
}
