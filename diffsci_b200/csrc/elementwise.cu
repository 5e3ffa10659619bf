// elementwise.cu -- K5/K6 bandwidth kernels: 2x pooling, skip add, layout conversion, channel
// concat, Fourier features, grouped small-batch linear (time MLPs), row softmax.
#include "common.cuh"
#include "philox.cuh"

namespace dsk {

// ---- 2x pooling on channels-last [B, D, H, W, C] (commonlayers.py:60-63,81; adm.py:361-371) ----
template <typename T, bool IS_MAX>
__global__ void __launch_bounds__(256) pool2x_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int D, int H, int W,
                                                      int C, int ndim) {
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
    float acc = IS_MAX ? -INFINITY : 0.0f;
    for (int a = 0; a < kd; ++a)
      for (int bb = 0; bb < 2; ++bb)
        for (int cc = 0; cc < 2; ++cc) {
          int64_t src = ((((int64_t)b * D + (dz * kd + a)) * H + (ho * 2 + bb)) * W + (wo * 2 + cc)) * C + c;
          float v = to_f32<T>(x[src]);
          acc = IS_MAX ? fmaxf(acc, v) : acc + v;
        }
    if (!IS_MAX) acc *= (ndim == 3 ? 0.125f : 0.25f);
    y[i] = from_f32<T>(acc);
  }
}

// bf16, C % 8 == 0: a thread pools 8 channels (one 16-byte vector) of one output pixel -- 8 (3-D) / 4 (2-D) vector loads,
// one vector store; max is taken on the bf16 values directly (exact), the mean in fp32.
template <typename T, bool IS_MAX>
__global__ void __launch_bounds__(256) pool2x_vec8_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int D, int H, int W,
                                                           int C8, int ndim) {
  typedef typename H16<T>::T2 T2;
  const int Do = ndim == 3 ? D / 2 : 1, Ho = H / 2, Wo = W / 2;
  const int kd = ndim == 3 ? 2 : 1;
  const int64_t total = (int64_t)B * Do * Ho * Wo * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    int64_t p = i / C8;
    const int wo = (int)(p % Wo); p /= Wo;
    const int ho = (int)(p % Ho); p /= Ho;
    const int dz = (int)(p % Do);
    const int b = (int)(p / Do);
    const int64_t base = ((((int64_t)b * D + dz * kd) * H + ho * 2) * W + wo * 2) * C8 + c;
    uint4 v[8];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int bb = 0; bb < 2; ++bb)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
          if (a < kd) v[(a * 2 + bb) * 2 + cc] = x[base + (((int64_t)a * H + bb) * W + cc) * C8];
    uint4 o;
    if (IS_MAX) {
      T2* oh = reinterpret_cast<T2*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        T2 m = reinterpret_cast<const T2*>(&v[0])[e];
#pragma unroll
        for (int k = 1; k < 8; ++k)
          if (k < 4 * kd) m = __hmax2(m, reinterpret_cast<const T2*>(&v[k])[e]);
        oh[e] = m;
      }
    } else {
      T2* oh = reinterpret_cast<T2*>(&o);
      const float sc = ndim == 3 ? 0.125f : 0.25f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float lo = 0.0f, hi = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < 4 * kd) {
            const T2 t = reinterpret_cast<const T2*>(&v[k])[e];
            lo += __low2float(t); hi += __high2float(t);
          }
        oh[e] = H16<T>::pack(lo * sc, hi * sc);
      }
    }
    y[i] = o;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = from_f32<T>(to_f32<T>(a[i]) + to_f32<T>(b[i]));
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = from_f32<TO>(to_f32<TI>(x[i]));
}

// fp32 [rows, C] -> split-fp16 [rows, 2C] (hi | lo, DSK_SPLIT_F16): the operand format of the split tensor-core kernels, for
// inputs that no fused producer (norm apply) writes in that form.  One thread = 4 channels of one row.
__global__ void __launch_bounds__(256) split_f16_kernel(const float4* __restrict__ x, __half* __restrict__ y, int64_t rows, int C4) {
  const int64_t total = rows * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C4;
    const int c = (int)(i - r * C4) * 4;
    const float4 v = x[i];
    const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn((v.x - f0.x) * 2048.0f, (v.y - f0.y) * 2048.0f);      // lo is stored times 2^11
    const __half2 l1 = __floats2half2_rn((v.z - f1.x) * 2048.0f, (v.w - f1.y) * 2048.0f);
    uint2 hi, lo;
    hi.x = *reinterpret_cast<const uint32_t*>(&h0); hi.y = *reinterpret_cast<const uint32_t*>(&h1);
    lo.x = *reinterpret_cast<const uint32_t*>(&l0); lo.y = *reinterpret_cast<const uint32_t*>(&l1);
    __half* row = y + r * (int64_t)(8 * C4);
    *reinterpret_cast<uint2*>(row + c) = hi;
    *reinterpret_cast<uint2*>(row + 4 * C4 + c) = lo;
  }
}

// y[b, s, c] = x[b, c, s]  (C is small at the module boundary: 1..4 channels)
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_cl_kernel(const float* __restrict__ x, T* __restrict__ y, int B, int C, int64_t S) {
  const int64_t total = (int64_t)B * C * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t p = i / C;
    int64_t s = p % S;
    int64_t b = p / S;
    y[i] = from_f32<T>(x[(b * C + c) * S + s]);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) cl_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ y, int B, int C, int64_t S) {
  const int64_t total = (int64_t)B * C * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t s = i % S;
    int64_t p = i / S;
    int c = (int)(p % C);
    int64_t b = p / C;
    y[i] = to_f32<T>(x[(b * S + s) * C + c]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) concat_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y,
                                                      int64_t rows, int Ca, int Cb) {
  const int Ct = Ca + Cb;
  const int64_t total = rows * Ct;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % Ct);
    int64_t r = i / Ct;
    y[i] = c < Ca ? a[r * Ca + c] : b[r * Cb + (c - Ca)];
  }
}

// ---- Fourier features (commonlayers.py:185-190) ---------------------------------------------
// x_proj = 2*pi*t*W in the reference's operation order ((2*pi*t)*W in fp32), then accurate
// sinf/cosf: |arg| reaches ~1e3 rad, fast-math intrinsics would destroy parity.
__global__ void fourier_kernel(const float* __restrict__ t, const float* __restrict__ W, float* __restrict__ out, int B, int half) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  int b = i / half, j = i - b * half;
  float p = (6.283185307179586f * t[b]) * W[j];
  out[(int64_t)b * 2 * half + j] = sinf(p);
  out[(int64_t)b * 2 * half + half + j] = cosf(p);
}

// ---- grouped small-batch linear ----------------------------------------------------------------
// One warp per output feature: the weight row is read once (coalesced) and reused for every batch
// row; B is small (1..256) on this path, the weights dominate the traffic.
constexpr int GL_MAXB = 8;
__global__ void __launch_bounds__(256) grouped_linear_kernel(const float* const* __restrict__ X, const float* const* __restrict__ Wt,
                                                              const float* const* __restrict__ bias, float* const* __restrict__ Y,
                                                              float* const* __restrict__ Z, const int* __restrict__ in_dim,
                                                              const int* __restrict__ out_dim, int B, int act) {
  const int g = blockIdx.y;
  const int K = in_dim[g], N = out_dim[g];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N) return;
  const float* w = Wt[g] + (int64_t)warp * K;
  const float* x = X[g];
  float* y = Y[g];
  float* z = Z != nullptr ? Z[g] : nullptr;           // pre-activation, kept for the backward pass
  const float bv = bias[g] != nullptr ? bias[g][warp] : 0.0f;
  for (int b0 = 0; b0 < B; b0 += GL_MAXB) {
    float acc[GL_MAXB];
#pragma unroll
    for (int r = 0; r < GL_MAXB; ++r) acc[r] = 0.0f;
    for (int k = lane; k < K; k += 32) {
      const float wv = w[k];
#pragma unroll
      for (int r = 0; r < GL_MAXB; ++r)
        if (b0 + r < B) acc[r] += wv * x[(int64_t)(b0 + r) * K + k];
    }
#pragma unroll
    for (int r = 0; r < GL_MAXB; ++r) {
      float v = warp_sum(acc[r]);
      if (lane == 0 && b0 + r < B) {
        v += bv;
        if (z != nullptr) z[(int64_t)(b0 + r) * N + warp] = v;
        if (act == 1) v = silu_f(v);
        else if (act == 2) v = fmaxf(v, 0.0f);
        y[(int64_t)(b0 + r) * N + warp] = v;
      }
    }
  }
}

// ---- grouped small-batch linear, backward ------------------------------------------------------------------------
// (a) one warp per (group, output feature n):  dz[b] = dy[b,n] * act'(z[b,n]) (written to dZ for pass (b)),
//     db[n] = sum_b dz[b],  dW[n,:] = sum_b dz[b] * x[b,:]   (lanes stride k; 16 accumulators = 512 k per sweep).
__global__ void __launch_bounds__(256) grouped_linear_wgrad_kernel(const float* const* __restrict__ dY, const float* const* __restrict__ Z,
                                                                    const float* const* __restrict__ X, float* const* __restrict__ dZ,
                                                                    float* const* __restrict__ dW, float* const* __restrict__ db,
                                                                    const int* __restrict__ in_dim, const int* __restrict__ out_dim,
                                                                    int B, int act) {
  const int g = blockIdx.y;
  const int K = in_dim[g], N = out_dim[g];
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= N) return;
  const float* dy = dY[g];
  const float* z = (act != 0 && Z != nullptr) ? Z[g] : nullptr;
  const float* x = X[g];
  float* dz = dZ[g];
  float* dw = dW[g] + (int64_t)n * K;
  float bsum = 0.0f;
  for (int kc = 0; kc < K; kc += 512) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
    for (int b = 0; b < B; ++b) {
      float d = dy[(int64_t)b * N + n];
      if (z != nullptr) {
        const float zz = z[(int64_t)b * N + n];
        if (act == 1) { const float sg = 1.0f / (1.0f + expf(-zz)); d *= sg * (1.0f + zz * (1.0f - sg)); }
        else d = zz > 0.0f ? d : 0.0f;
      }
      if (kc == 0) {
        bsum += d;
        if (lane == 0 && dz != dy) dz[(int64_t)b * N + n] = d;
      }
      const float* xr = x + (int64_t)b * K + kc + lane;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (kc + lane + 32 * j < K) acc[j] = fmaf(d, xr[32 * j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (kc + lane + 32 * j < K) dw[kc + lane + 32 * j] = acc[j];
  }
  if (lane == 0 && db != nullptr && db[g] != nullptr) db[g][n] = bsum;
}

// (b) dX[b,k] = sum_n dZ[b,n] W[n,k]; `shared`: every group consumed the same X, so ONE dX = sum over groups too
//     (ADM's per-block FiLM projections of the shared embedding).  Block = 128 k-columns x GLB_ROWS batch rows.
constexpr int GLB_ROWS = 8;
constexpr int GLB_NL = 8;        // lanes over the reduction (n) axis: a `shared` dX over 28 groups x 256 outputs was ONE block of 128
                                 // threads walking 7 168 terms (100 us per launch in C4 training); 8 lanes + a fixed-order combine
__global__ void __launch_bounds__(128 * GLB_NL) grouped_linear_dgrad_kernel(const float* const* __restrict__ dZ, const float* const* __restrict__ Wt,
                                                                             float* const* __restrict__ dX, const int* __restrict__ in_dim,
                                                                             const int* __restrict__ out_dim, int ngroups, int B, int shared,
                                                                             int accumulate) {
  const int g0 = shared ? 0 : blockIdx.z, g1 = shared ? ngroups : blockIdx.z + 1;
  const int K = in_dim[g0];
  const int kx = threadIdx.x & 127, nl = threadIdx.x >> 7;
  const int k = blockIdx.x * 128 + kx;
  const int b0 = blockIdx.y * GLB_ROWS;
  if (dX[g0] == nullptr) return;
  float acc[GLB_ROWS];
#pragma unroll
  for (int r = 0; r < GLB_ROWS; ++r) acc[r] = 0.0f;
  for (int g = g0; g < g1; ++g) {
    const int N = out_dim[g];
    const float* w = Wt[g];
    const float* dz = dZ[g];
    if (k < K)
      for (int n = nl; n < N; n += GLB_NL) {
        const float wv = w[(int64_t)n * K + k];
#pragma unroll
        for (int r = 0; r < GLB_ROWS; ++r)
          if (b0 + r < B) acc[r] = fmaf(dz[(int64_t)(b0 + r) * N + n], wv, acc[r]);
      }
  }
  __shared__ float red[GLB_NL][GLB_ROWS][128];
#pragma unroll
  for (int r = 0; r < GLB_ROWS; ++r) red[nl][r][kx] = acc[r];
  __syncthreads();
  if (nl == 0 && k < K) {
    float* o = dX[g0];
#pragma unroll
    for (int r = 0; r < GLB_ROWS; ++r)
      if (b0 + r < B) {
        float t = 0.0f;
#pragma unroll
        for (int l = 0; l < GLB_NL; ++l) t += red[l][r][kx];
        const int64_t i = (int64_t)(b0 + r) * K + k;
        o[i] = accumulate ? o[i] + t : t;
      }
  }
}

// (c) dZ_g = dY_g * act'(Z_g) and db_g = colsum(dZ_g) for every group: block = 32 columns x 8 row lanes.
__global__ void __launch_bounds__(256) grouped_dz_bias_kernel(const float* const* __restrict__ dY, const float* const* __restrict__ Z,
                                                               float* const* __restrict__ dZ, float* const* __restrict__ db,
                                                               const int* __restrict__ out_dim, int B, int act) {
  const int g = blockIdx.y;
  const int N = out_dim[g];
  const int n = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  if (blockIdx.x * 32 >= N) return;
  const float* dy = dY[g];
  const float* z = (act != 0 && Z != nullptr) ? Z[g] : nullptr;
  float* dz = dZ[g];
  double acc = 0;
  if (n < N)
    for (int b = rl; b < B; b += 8) {
      float d = dy[(int64_t)b * N + n];
      if (z != nullptr) {
        const float zz = z[(int64_t)b * N + n];
        if (act == 1) { const float sg = 1.0f / (1.0f + expf(-zz)); d *= sg * (1.0f + zz * (1.0f - sg)); }
        else d = zz > 0.0f ? d : 0.0f;
        dz[(int64_t)b * N + n] = d;
      }
      acc += (double)d;
    }
  __shared__ double red[8][33];
  red[rl][threadIdx.x & 31] = acc;
  __syncthreads();
  if (rl == 0 && n < N && db != nullptr && db[g] != nullptr) {
    double s = 0;
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x & 31];
    db[g][n] = (float)s;
  }
}

// out = x * (1 - m) + y * m, the known-region blend of Scheduler.inpaint / repaint (schedulers.py:111-116,151,162);
// the mask is broadcast over the leading dimensions (m index = i mod mask_n).  out may alias x.
__global__ void __launch_bounds__(256) mask_blend_kernel(float* __restrict__ out, const float* __restrict__ x, const float* __restrict__ y,
                                                          const float* __restrict__ m, int64_t n, int64_t mask_n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float mv = m[i % mask_n];
    out[i] = x[i] * (1.0f - mv) + y[i] * mv;
  }
}

// ---- row softmax (in place, fp32) ------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ S, int64_t rows, int cols) {
  __shared__ float red[8];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    float* p = S + r * cols;
    float m = -INFINITY;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) m = fmaxf(m, p[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
    __syncthreads();
    float sum = 0.0f;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
      float e = expf(p[c] - m);
      p[c] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    sum = 0.0f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) sum += red[w];
    __syncthreads();
    const float inv = 1.0f / sum;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) p[c] *= inv;
  }
}

// out = a0*x + a1*r1 + a2*r2 + a3*z  (null pointers are skipped): the arithmetic of the
// Integrator.step seam (integrators.py:29-113) and Scheduler.rhs scaling for foreign score functions.
__global__ void __launch_bounds__(256) lincomb_kernel(float* __restrict__ out, int64_t n, const float* __restrict__ x, float a0,
                                                       const float* __restrict__ r1, float a1, const float* __restrict__ r2,
                                                       float a2, const float* __restrict__ z, float a3) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = 0.0f;
    if (x != nullptr) v = a0 * x[i];
    if (r1 != nullptr) v += a1 * r1[i];
    if (r2 != nullptr) v += a2 * r2[i];
    if (z != nullptr) v += a3 * z[i];
    out[i] = v;
  }
}

// ---- dropout (training): y = x * keep / (1 - p) (+ dres), keep(i) = [u_i >= p] with u_i the i-th uniform of Philox stream
// (seed, stream_id) -- counter-based, so the backward pass regenerates the mask of its forward site instead of storing it.
// Replaces torch.nn.Dropout in ResnetBlockC / ADMBaseBlock (commonlayers.py:792, 830; adm.py:312-313).
template <typename T>
__global__ void __launch_bounds__(256) dropout_kernel(const T* __restrict__ x, const T* dres, T* y, int64_t n, float p, float scale,
                                                       uint64_t seed, uint32_t stream_id) {
  const int64_t quads = (n + 3) / 4;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (int64_t)gridDim.x * blockDim.x) {
    const uint4 ctr = make_uint4((uint32_t)q, (uint32_t)(q >> 32), stream_id, 0xD509u);
    const uint4 r = philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t i = q * 4 + k;
      if (i < n) {
        const float u = (float)(rr[k] >> 8) * 5.9604644775390625e-08f;   // 24-bit uniform in [0, 1)
        float v = u >= p ? to_f32<T>(x[i]) * scale : 0.0f;
        if (dres != nullptr) v += to_f32<T>(dres[i]);
        y[i] = from_f32<T>(v);
      }
    }
  }
}

__global__ void __launch_bounds__(256) philox_normal_kernel(float* __restrict__ out, int64_t n, uint64_t seed, uint32_t stream_id) {
  const int64_t quads = (n + 3) / 4;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (int64_t)gridDim.x * blockDim.x) {
    float v[4];
    philox_normal4(seed, stream_id, (uint64_t)q, v);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (q * 4 + k < n) out[q * 4 + k] = v[k];
  }
}


// ---- circular padding (CircularConv2d/3d, commonlayers.py:918-1032): y[b, d', h', w', :] = x[b, (d'-pd) mod D, (h'-1) mod H,
// (w'-1) mod W, :] with pd = 1 for 3-D, 0 for 2-D.  The halo-padded copy is what the tcgen05 convolution kernels read through
// TMA (a box load cannot wrap).  V = elements moved per thread (16-byte vectors when the channel run allows).
template <typename VT>
__global__ void __launch_bounds__(256) pad_circular_kernel(const VT* __restrict__ x, VT* __restrict__ y, int B, int D, int H, int W,
                                                            int Cv, int pd) {
  const int Dp = D + 2 * pd, Hp = H + 2, Wp = W + 2;
  const int64_t total = (int64_t)B * Dp * Hp * Wp * Cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int c = (int)(t % Cv); t /= Cv;
    int w = (int)(t % Wp) - 1; t /= Wp;
    int h = (int)(t % Hp) - 1; t /= Hp;
    int d = (int)(t % Dp) - pd;
    const int b = (int)(t / Dp);
    w = w < 0 ? w + W : (w >= W ? w - W : w);
    h = h < 0 ? h + H : (h >= H ? h - H : h);
    d = d < 0 ? d + D : (d >= D ? d - D : d);
    y[i] = x[((((int64_t)b * D + d) * H + h) * W + w) * Cv + c];
  }
}

int pad_circular_launch(const void* x, void* y, int B, int D, int H, int W, int C, int ndim, int dtype, cudaStream_t st) {
  const int pd = ndim == 3 ? 1 : 0;
  const int es = is_h16(dtype) ? 2 : 4;
  const int64_t row = (int64_t)C * es;                      // bytes of one pixel's channel run
  const int64_t npix = (int64_t)B * (D + 2 * pd) * (H + 2) * (W + 2);
  const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  if (row % 16 == 0 && al) {
    const int Cv = (int)(row / 16);
    DSK_LAUNCH(pad_circular_kernel<uint4>, grid_for(npix * Cv, 256, 16), 256, 0, st, (const uint4*)x, (uint4*)y, B, D, H, W, Cv, pd);
  } else if (es == 4) {
    DSK_LAUNCH(pad_circular_kernel<float>, grid_for(npix * C, 256, 16), 256, 0, st, (const float*)x, (float*)y, B, D, H, W, C, pd);
  } else {
    DSK_LAUNCH(pad_circular_kernel<uint16_t>, grid_for(npix * C, 256, 16), 256, 0, st, (const uint16_t*)x, (uint16_t*)y, B, D, H, W, C,
               pd);
  }
  return DSK_OK;
}
}  // namespace dsk

using namespace dsk;

extern "C" int dsk_lincomb(float* out, int64_t n, const float* x, float a0, const float* r1, float a1, const float* r2,
                           float a2, const float* z, float a3, void* stream) {
  DSK_REQUIRE(out && n > 0, "dsk_lincomb: bad arguments");
  DSK_LAUNCH(lincomb_kernel, grid_for(n, 256, 16), 256, 0, as_stream(stream), out, n, x, a0, r1, a1, r2, a2, z, a3);
  return DSK_OK;
}

extern "C" int dsk_dropout(const void* x, const void* dres, void* y, int64_t n, float p, uint64_t seed, uint32_t stream_id,
                           int dtype, void* stream) {
  DSK_REQUIRE(x && y && n > 0, "dsk_dropout: bad arguments");
  DSK_REQUIRE(p >= 0.0f && p < 1.0f, "dsk_dropout: p = %f outside [0, 1)", (double)p);
  const int grid = grid_for((n + 3) / 4, 256, 16);
  const float scale = 1.0f / (1.0f - p);
  if (dtype == DSK_F32)
    DSK_LAUNCH(dropout_kernel<float>, grid, 256, 0, as_stream(stream), (const float*)x, (const float*)dres, (float*)y, n, p, scale, seed,
               stream_id);
  else if (dtype == DSK_BF16)
    DSK_LAUNCH(dropout_kernel<__nv_bfloat16>, grid, 256, 0, as_stream(stream), (const __nv_bfloat16*)x, (const __nv_bfloat16*)dres,
               (__nv_bfloat16*)y, n, p, scale, seed, stream_id);
  else DSK_REQUIRE(false, "dsk_dropout: bad dtype %d", dtype);
  return DSK_OK;
}

extern "C" int dsk_philox_normal(float* out, int64_t n, uint64_t seed, uint32_t stream_id, void* stream) {
  DSK_REQUIRE(out && n > 0, "dsk_philox_normal: bad arguments");
  DSK_LAUNCH(philox_normal_kernel, grid_for((n + 3) / 4, 256, 16), 256, 0, as_stream(stream), out, n, seed, stream_id);
  return DSK_OK;
}

// fp32 input, C % 4 == 0: a thread pools 4 channels (one float4) of one output pixel; the result is stored as fp32 or rounded
// once to a 16-bit format (OUT16: 1 bf16, 2 fp16) -- the fp32-storage modes pool straight into the operand copy of the
// DownSampler convolution instead of writing fp32 and casting it in a second pass.
template <int OUT16, bool IS_MAX>
__global__ void __launch_bounds__(256) pool2x_f32v4_kernel(const float4* __restrict__ x, void* __restrict__ y, int B, int D, int H, int W,
                                                            int C4, int ndim) {
  const int Do = ndim == 3 ? D / 2 : 1, Ho = H / 2, Wo = W / 2;
  const int kd = ndim == 3 ? 2 : 1;
  const int64_t total = (int64_t)B * Do * Ho * Wo * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    int64_t p = i / C4;
    const int wo = (int)(p % Wo); p /= Wo;
    const int ho = (int)(p % Ho); p /= Ho;
    const int dz = (int)(p % Do);
    const int b = (int)(p / Do);
    const int64_t base = ((((int64_t)b * D + dz * kd) * H + ho * 2) * W + wo * 2) * C4 + c;
    float4 v[8];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int bb = 0; bb < 2; ++bb)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
          if (a < kd) v[(a * 2 + bb) * 2 + cc] = x[base + (((int64_t)a * H + bb) * W + cc) * C4];
    float4 o = v[0];
#pragma unroll
    for (int k = 1; k < 8; ++k)
      if (k < 4 * kd) {
        if (IS_MAX) { o.x = fmaxf(o.x, v[k].x); o.y = fmaxf(o.y, v[k].y); o.z = fmaxf(o.z, v[k].z); o.w = fmaxf(o.w, v[k].w); }
        else { o.x += v[k].x; o.y += v[k].y; o.z += v[k].z; o.w += v[k].w; }
      }
    if (!IS_MAX) { const float sc = ndim == 3 ? 0.125f : 0.25f; o.x *= sc; o.y *= sc; o.z *= sc; o.w *= sc; }
    if (OUT16 == 0) reinterpret_cast<float4*>(y)[i] = o;
    else reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_h2(o.x, o.y, OUT16 == 2), pack_h2(o.z, o.w, OUT16 == 2));
  }
}

extern "C" int dsk_pool2x_f32(const float* x, void* y, int B, int D, int H, int W, int C, int ndim, int is_max, int out_dtype,
                              void* stream) {
  DSK_REQUIRE(x && y, "dsk_pool2x_f32: null pointer");
  DSK_REQUIRE(B > 0 && D > 0 && H > 1 && W > 1 && C > 0 && C % 4 == 0 && (ndim == 2 || ndim == 3), "dsk_pool2x_f32: bad shape (C %% 4 == 0)");
  DSK_REQUIRE(ndim == 2 ? D == 1 : D > 1, "dsk_pool2x_f32: D=%d inconsistent with ndim=%d", D, ndim);
  DSK_REQUIRE(out_dtype == DSK_F32 || is_h16(out_dtype), "dsk_pool2x_f32: bad out_dtype %d", out_dtype);
  const int64_t total = (int64_t)B * (ndim == 3 ? D / 2 : 1) * (H / 2) * (W / 2) * (C / 4);
  const int grid = grid_for(total, 256, 16);
  cudaStream_t st = as_stream(stream);
#define POOLF(O, M) DSK_LAUNCH((pool2x_f32v4_kernel<O, M>), grid, 256, 0, st, (const float4*)x, y, B, D, H, W, C / 4, ndim)
  if (out_dtype == DSK_F32) { if (is_max) POOLF(0, true); else POOLF(0, false); }
  else if (out_dtype == DSK_BF16) { if (is_max) POOLF(1, true); else POOLF(1, false); }
  else { if (is_max) POOLF(2, true); else POOLF(2, false); }
#undef POOLF
  return DSK_OK;
}

extern "C" int dsk_pool2x(const void* x, void* y, int B, int D, int H, int W, int C, int ndim, int is_max, int dtype,
                          void* stream) {
  DSK_REQUIRE(x && y, "dsk_pool2x: null pointer");
  DSK_REQUIRE(B > 0 && D > 0 && H > 1 && W > 1 && C > 0 && (ndim == 2 || ndim == 3), "dsk_pool2x: bad shape");
  DSK_REQUIRE(ndim == 2 ? D == 1 : D > 1, "dsk_pool2x: D=%d inconsistent with ndim=%d", D, ndim);
  const int64_t total = (int64_t)B * (ndim == 3 ? D / 2 : 1) * (H / 2) * (W / 2) * C;
  const int grid = grid_for(total, 256, 16);
  cudaStream_t st = as_stream(stream);
#define POOL(T, M) DSK_LAUNCH((pool2x_kernel<T, M>), grid, 256, 0, st, (const T*)x, (T*)y, B, D, H, W, C, ndim)
  if (is_h16(dtype) && C % 8 == 0) {
    const int vgrid = grid_for(total / 8, 256, 16);
#define POOLV(T, M) DSK_LAUNCH((pool2x_vec8_kernel<T, M>), vgrid, 256, 0, st, (const uint4*)x, (uint4*)y, B, D, H, W, C / 8, ndim)
    if (dtype == DSK_F16) { if (is_max) POOLV(__half, true); else POOLV(__half, false); }
    else { if (is_max) POOLV(__nv_bfloat16, true); else POOLV(__nv_bfloat16, false); }
#undef POOLV
    return DSK_OK;
  }
  if (dtype == DSK_F32) { if (is_max) POOL(float, true); else POOL(float, false); }
  else if (dtype == DSK_BF16) { if (is_max) POOL(__nv_bfloat16, true); else POOL(__nv_bfloat16, false); }
  else if (dtype == DSK_F16) { if (is_max) POOL(__half, true); else POOL(__half, false); }
  else DSK_REQUIRE(false, "dsk_pool2x: bad dtype %d", dtype);
#undef POOL
  return DSK_OK;
}

extern "C" int dsk_add(const void* a, const void* b, void* y, int64_t n, int dtype, void* stream) {
  DSK_REQUIRE(a && b && y && n > 0, "dsk_add: bad arguments");
  const int grid = grid_for(n, 256, 16);
  if (dtype == DSK_F32) DSK_LAUNCH(add_kernel<float>, grid, 256, 0, as_stream(stream), (const float*)a, (const float*)b, (float*)y, n);
  else if (dtype == DSK_BF16) DSK_LAUNCH(add_kernel<__nv_bfloat16>, grid, 256, 0, as_stream(stream), (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)y, n);
  else if (dtype == DSK_F16) DSK_LAUNCH(add_kernel<__half>, grid, 256, 0, as_stream(stream), (const __half*)a, (const __half*)b, (__half*)y, n);
  else DSK_REQUIRE(false, "dsk_add: bad dtype %d", dtype);
  return DSK_OK;
}

extern "C" int dsk_cast(const void* x, void* y, int64_t n, int in_dtype, int out_dtype, void* stream) {
  DSK_REQUIRE(x && y && n > 0, "dsk_cast: bad arguments");
  const int grid = grid_for(n, 256, 16);
  cudaStream_t st = as_stream(stream);
  if (in_dtype == DSK_F32 && out_dtype == DSK_BF16) DSK_LAUNCH((cast_kernel<float, __nv_bfloat16>), grid, 256, 0, st, (const float*)x, (__nv_bfloat16*)y, n);
  else if (in_dtype == DSK_BF16 && out_dtype == DSK_F32) DSK_LAUNCH((cast_kernel<__nv_bfloat16, float>), grid, 256, 0, st, (const __nv_bfloat16*)x, (float*)y, n);
  else if (in_dtype == DSK_F32 && out_dtype == DSK_F32) DSK_LAUNCH((cast_kernel<float, float>), grid, 256, 0, st, (const float*)x, (float*)y, n);
  else if (in_dtype == DSK_F32 && out_dtype == DSK_F16) DSK_LAUNCH((cast_kernel<float, __half>), grid, 256, 0, st, (const float*)x, (__half*)y, n);
  else if (in_dtype == DSK_F16 && out_dtype == DSK_F32) DSK_LAUNCH((cast_kernel<__half, float>), grid, 256, 0, st, (const __half*)x, (float*)y, n);
  else DSK_REQUIRE(false, "dsk_cast: bad dtypes %d -> %d", in_dtype, out_dtype);
  return DSK_OK;
}

extern "C" int dsk_split_f16(const float* x, void* y, int64_t rows, int C, void* stream) {
  DSK_REQUIRE(x && y && rows > 0 && C > 0 && C % 4 == 0, "dsk_split_f16: bad arguments (C %% 4 == 0)");
  DSK_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, "dsk_split_f16: 16-byte alignment");
  DSK_LAUNCH(split_f16_kernel, grid_for(rows * (C / 4), 256, 16), 256, 0, as_stream(stream), (const float4*)x, (__half*)y, rows, C / 4);
  return DSK_OK;
}

extern "C" int dsk_nchw_to_cl(const float* x, void* y, int B, int C, int64_t S, int dtype, void* stream) {
  DSK_REQUIRE(x && y && B > 0 && C > 0 && S > 0, "dsk_nchw_to_cl: bad arguments");
  const int grid = grid_for((int64_t)B * C * S, 256, 16);
  if (dtype == DSK_F32) DSK_LAUNCH(nchw_to_cl_kernel<float>, grid, 256, 0, as_stream(stream), x, (float*)y, B, C, S);
  else if (dtype == DSK_BF16) DSK_LAUNCH(nchw_to_cl_kernel<__nv_bfloat16>, grid, 256, 0, as_stream(stream), x, (__nv_bfloat16*)y, B, C, S);
  else if (dtype == DSK_F16) DSK_LAUNCH(nchw_to_cl_kernel<__half>, grid, 256, 0, as_stream(stream), x, (__half*)y, B, C, S);
  else DSK_REQUIRE(false, "dsk_nchw_to_cl: bad dtype %d", dtype);
  return DSK_OK;
}

extern "C" int dsk_cl_to_nchw(const void* x, float* y, int B, int C, int64_t S, int dtype, void* stream) {
  DSK_REQUIRE(x && y && B > 0 && C > 0 && S > 0, "dsk_cl_to_nchw: bad arguments");
  const int grid = grid_for((int64_t)B * C * S, 256, 16);
  if (dtype == DSK_F32) DSK_LAUNCH(cl_to_nchw_kernel<float>, grid, 256, 0, as_stream(stream), (const float*)x, y, B, C, S);
  else if (dtype == DSK_BF16) DSK_LAUNCH(cl_to_nchw_kernel<__nv_bfloat16>, grid, 256, 0, as_stream(stream), (const __nv_bfloat16*)x, y, B, C, S);
  else if (dtype == DSK_F16) DSK_LAUNCH(cl_to_nchw_kernel<__half>, grid, 256, 0, as_stream(stream), (const __half*)x, y, B, C, S);
  else DSK_REQUIRE(false, "dsk_cl_to_nchw: bad dtype %d", dtype);
  return DSK_OK;
}

extern "C" int dsk_concat_channels(const void* a, const void* b, void* y, int64_t rows, int Ca, int Cb, int dtype,
                                   void* stream) {
  DSK_REQUIRE(a && b && y && rows > 0 && Ca > 0 && Cb > 0, "dsk_concat_channels: bad arguments");
  const int grid = grid_for(rows * (Ca + Cb), 256, 16);
  if (dtype == DSK_F32) DSK_LAUNCH(concat_kernel<float>, grid, 256, 0, as_stream(stream), (const float*)a, (const float*)b, (float*)y, rows, Ca, Cb);
  else if (is_h16(dtype)) DSK_LAUNCH(concat_kernel<uint16_t>, grid, 256, 0, as_stream(stream), (const uint16_t*)a, (const uint16_t*)b, (uint16_t*)y, rows, Ca, Cb);
  else DSK_REQUIRE(false, "dsk_concat_channels: bad dtype %d", dtype);
  return DSK_OK;
}

extern "C" int dsk_fourier(const float* t, const float* W, float* out, int B, int half, void* stream) {
  DSK_REQUIRE(t && W && out && B > 0 && half > 0, "dsk_fourier: bad arguments");
  DSK_LAUNCH(fourier_kernel, (B * half + 127) / 128, 128, 0, as_stream(stream), t, W, out, B, half);
  return DSK_OK;
}

extern "C" int dsk_grouped_linear(const float* const* X, const float* const* W, const float* const* bias, float* const* Y,
                                  float* const* Z, const int* in_dim, const int* out_dim, int ngroups, int max_out, int B, int act,
                                  void* stream) {
  DSK_REQUIRE(X && W && bias && Y && in_dim && out_dim, "dsk_grouped_linear: null pointer");
  DSK_REQUIRE(ngroups > 0 && ngroups <= 65535 && max_out > 0 && B > 0 && act >= 0 && act <= 2, "dsk_grouped_linear: bad arguments");
  dim3 grid((max_out + 7) / 8, ngroups);
  DSK_LAUNCH(grouped_linear_kernel, grid, 256, 0, as_stream(stream), X, W, bias, Y, Z, in_dim, out_dim, B, act);
  return DSK_OK;
}

extern "C" int dsk_grouped_linear_bwd(const float* const* dY, const float* const* Z, const float* const* X, const float* const* W,
                                      float* const* dZ, float* const* dW, float* const* db, float* const* dX, const int* in_dim,
                                      const int* out_dim, int ngroups, int max_out, int max_in, int B, int act, int shared_dx,
                                      int accumulate_dx, void* stream) {
  DSK_REQUIRE(dY && X && W && dZ && dW && in_dim && out_dim, "dsk_grouped_linear_bwd: null pointer");
  DSK_REQUIRE(ngroups > 0 && ngroups <= 65535 && max_out > 0 && max_in > 0 && B > 0 && act >= 0 && act <= 2 && (act == 0 || Z != nullptr),
              "dsk_grouped_linear_bwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  dim3 wg((max_out + 7) / 8, ngroups);
  DSK_LAUNCH(grouped_linear_wgrad_kernel, wg, 256, 0, st, dY, Z, X, dZ, dW, db, in_dim, out_dim, B, act);
  if (dX != nullptr) {
    DSK_REQUIRE((B + GLB_ROWS - 1) / GLB_ROWS <= 65535, "dsk_grouped_linear_bwd: B too large");
    dim3 dg((max_in + 127) / 128, (B + GLB_ROWS - 1) / GLB_ROWS, shared_dx ? 1 : ngroups);
    DSK_LAUNCH(grouped_linear_dgrad_kernel, dg, 128 * GLB_NL, 0, st, (const float* const*)dZ, W, dX, in_dim, out_dim, ngroups, B, shared_dx,
               accumulate_dx);
  }
  return DSK_OK;
}

extern "C" int dsk_softmax_rows(float* S, int64_t rows, int cols, void* stream) {
  DSK_REQUIRE(S && rows > 0 && cols > 0, "dsk_softmax_rows: bad arguments");
  int64_t grid = rows < (int64_t)DSK_NUM_SMS * 16 ? rows : (int64_t)DSK_NUM_SMS * 16;
  DSK_LAUNCH(softmax_rows_kernel, (int)grid, 256, 0, as_stream(stream), S, rows, cols);
  return DSK_OK;
}

extern "C" int dsk_grouped_dz_bias(const float* const* dY, const float* const* Z, float* const* dZ, float* const* db, const int* out_dim,
                                   int ngroups, int max_out, int B, int act, void* stream) {
  DSK_REQUIRE(dY && dZ && out_dim && ngroups > 0 && ngroups <= 65535 && max_out > 0 && B > 0 && act >= 0 && act <= 2 && (act == 0 || Z),
              "dsk_grouped_dz_bias: bad arguments");
  dim3 grid((max_out + 31) / 32, ngroups);
  DSK_LAUNCH(grouped_dz_bias_kernel, grid, 256, 0, as_stream(stream), dY, Z, dZ, db, out_dim, B, act);
  return DSK_OK;
}

extern "C" int dsk_mask_blend(float* out, const float* x, const float* y, const float* mask, int64_t n, int64_t mask_n, void* stream) {
  DSK_REQUIRE(out && x && y && mask && n > 0 && mask_n > 0 && n % mask_n == 0, "dsk_mask_blend: bad arguments");
  DSK_LAUNCH(mask_blend_kernel, grid_for(n, 256, 8), 256, 0, as_stream(stream), out, x, y, mask, n, mask_n);
  return DSK_OK;
}

extern "C" int dsk_pad_circular(const void* x, void* y, int B, int D, int H, int W, int C, int ndim, int dtype, void* stream) {
  DSK_REQUIRE(x && y && B > 0 && D > 0 && H > 0 && W > 0 && C > 0, "dsk_pad_circular: bad arguments");
  DSK_REQUIRE((ndim == 2 && D == 1) || ndim == 3, "dsk_pad_circular: bad ndim/D");
  DSK_REQUIRE(dtype == DSK_F32 || is_h16(dtype), "dsk_pad_circular: bad dtype %d", dtype);
  return pad_circular_launch(x, y, B, D, H, W, C, ndim, dtype, as_stream(stream));
}
