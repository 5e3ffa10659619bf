// train.cu -- K8/K9: fused EDM loss forward+backward, multi-tensor EMA and AdamW(+EMA).
//
// Reference: KarrasModule.loss_fn (karras/karrasmodule.py:590-650) with
// EDMNoiseSampler.loss_weighting (karras/noisesamplers.py:30-33); ModelEMA.update
// (karras/ema.py:139-147, one lerp_ launch per tensor there); torch.optim.AdamW defaults
// (karrasmodule.py:497-500).
#include "common.cuh"

namespace dsk {

// F, dF: fp32 NC(D)HW like x (the training path keeps fp32 at the loss boundary).
__global__ void __launch_bounds__(256) edm_loss_kernel(const float* __restrict__ F, const float* __restrict__ x,
                                                        const float* __restrict__ noise, const float* __restrict__ sigma,
                                                        const float* __restrict__ mask, float* __restrict__ loss_out,
                                                        float* __restrict__ dF, int B, int64_t CS, float sd, int kind,
                                                        const float* __restrict__ c_out_v, const float* __restrict__ c_skip_v,
                                                        const float* __restrict__ lam_v, float delta) {
  const int64_t N = (int64_t)B * CS;
  const float invN = 1.0f / (float)N;
  float local = 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / CS);
    const float sg = sigma[b];
    float c_skip, c_out, lam;
    if (c_out_v != nullptr) {   // any preconditioner / noise sampler: per-sample coefficients from the host objects
      c_out = c_out_v[b]; c_skip = c_skip_v[b]; lam = lam_v[b];
    } else {
      const float sum = sg * sg + sd * sd, rt = sqrtf(sum);
      c_skip = (sd * sd) / sum; c_out = (sg * sd) / rt;
      lam = sum / ((sg * sd) * (sg * sd));
    }
    const float xv = x[i];
    const float xn = xv + sg * noise[i];
    const float D = c_out * F[i] + c_skip * xn;
    const float r = D - xv;
    float l, g;
    if (kind == 0) {  // torch.nn.HuberLoss(delta): default 1 (karrasmodule.py:541-542), configurable (:560-562)
      const float a = fabsf(r);
      l = a <= delta ? 0.5f * r * r : delta * (a - 0.5f * delta);
      g = fminf(fmaxf(r, -delta), delta);
    } else {          // MSE
      l = r * r;
      g = 2.0f * r;
    }
    const float keep = mask != nullptr ? 1.0f - mask[i] : 1.0f;
    local += lam * (l * keep);
    dF[i] = (c_out * lam) * (g * keep) * invN;
  }
  local = warp_sum(local);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(loss_out, s * invN);
  }
}

// The same loss with one partial sum PER SAMPLE (grid.y = b): loss_b[b] = sum_n lam_b l(D, x) keep / (B C S).  Needed when a
// per-sample factor multiplies the loss on the host side and has its own gradient -- the learned uncertainty weighting
// exp(-u(c_noise)) of KarrasModule.loss_fn with has_dynamic_loss_weight (karrasmodule.py:594-602, DynamicLossWeight :1256-1278).
__global__ void __launch_bounds__(256) precond_loss_rows_kernel(const float* __restrict__ F, const float* __restrict__ x,
                                                                 const float* __restrict__ noise, const float* __restrict__ sigma,
                                                                 const float* __restrict__ mask, float* __restrict__ loss_b,
                                                                 float* __restrict__ dF, int B, int64_t CS, int kind,
                                                                 const float* __restrict__ c_out_v, const float* __restrict__ c_skip_v,
                                                                 const float* __restrict__ lam_v, float delta) {
  const int b = blockIdx.y;
  const float invN = 1.0f / (float)((int64_t)B * CS);
  const float sg = sigma[b], c_out = c_out_v[b], c_skip = c_skip_v[b], lam = lam_v[b];
  const int64_t base = (int64_t)b * CS;
  float local = 0.0f;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < CS; n += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = base + n;
    const float xv = x[i];
    const float r = c_out * F[i] + c_skip * (xv + sg * noise[i]) - xv;
    float l, g;
    if (kind == 0) {
      const float a = fabsf(r);
      l = a <= delta ? 0.5f * r * r : delta * (a - 0.5f * delta);
      g = fminf(fmaxf(r, -delta), delta);
    } else {
      l = r * r;
      g = 2.0f * r;
    }
    const float keep = mask != nullptr ? 1.0f - mask[i] : 1.0f;
    local += lam * (l * keep);
    dF[i] = (c_out * lam) * (g * keep) * invN;
  }
  local = warp_sum(local);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(loss_b + b, s * invN);
  }
}

// Multi-tensor kernels.  The tensors differ in size by four orders of magnitude (a bias vs a 512x512x3x3 weight), so
// the work is cut into fixed chunks of MT_CHUNK elements over ALL tensors: every block builds the prefix of chunk counts in
// shared memory (a few hundred entries) and grid-strides over the global chunk index, binary-searching its tensor.
constexpr int MT_CHUNK = 2048, MT_THREADS = 256, MT_MAX_TENSORS = 4096;

struct MtChunk { int t; int64_t off; int n; };

__device__ __forceinline__ int64_t mt_prefix(const int64_t* __restrict__ numel, int nt, int64_t* pre) {
  for (int i = threadIdx.x; i < nt; i += blockDim.x) pre[i + 1] = (numel[i] + MT_CHUNK - 1) / MT_CHUNK;
  __syncthreads();
  if (threadIdx.x == 0) {
    pre[0] = 0;
    for (int i = 0; i < nt; ++i) pre[i + 1] += pre[i];
  }
  __syncthreads();
  return pre[nt];
}

__device__ __forceinline__ MtChunk mt_find(const int64_t* __restrict__ numel, int nt, const int64_t* pre, int64_t chunk) {
  int lo = 0, hi = nt;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (pre[mid] <= chunk) lo = mid; else hi = mid;
  }
  MtChunk c;
  c.t = lo;
  c.off = (chunk - pre[lo]) * MT_CHUNK;
  const int64_t left = numel[lo] - c.off;
  c.n = left < MT_CHUNK ? (int)left : MT_CHUNK;
  return c;
}

__global__ void __launch_bounds__(MT_THREADS) ema_kernel(float* const* __restrict__ shadow, const float* const* __restrict__ param,
                                                          const int64_t* __restrict__ numel, int nt, float w) {
  extern __shared__ int64_t mt_pre[];
  const int64_t total = mt_prefix(numel, nt, mt_pre);
  for (int64_t chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
    const MtChunk c = mt_find(numel, nt, mt_pre, chunk);
    float* s = shadow[c.t] + c.off;
    const float* p = param[c.t] + c.off;
#pragma unroll
    for (int k = 0; k < MT_CHUNK / MT_THREADS; ++k) {
      const int i = threadIdx.x + k * MT_THREADS;
      if (i < c.n) {
        // torch.lerp(s, p, w) = s + w*(p - s) for w < 0.5, else p - (p - s)*(1 - w)
        const float sv = s[i], pv = p[i];
        s[i] = w < 0.5f ? sv + w * (pv - sv) : pv - (pv - sv) * (1.0f - w);
      }
    }
  }
}

__global__ void __launch_bounds__(MT_THREADS) adamw_ema_kernel(float* const* __restrict__ P, const float* const* __restrict__ G,
                                                                float* const* __restrict__ Mo, float* const* __restrict__ Vo,
                                                                float* const* __restrict__ Sh, const int64_t* __restrict__ numel, int nt,
                                                                float lr, float b1, float b2, float eps, float wd, float bc1,
                                                                float bc2_sqrt, float ema_w, float gscale) {
  extern __shared__ int64_t mt_pre[];
  const int64_t total = mt_prefix(numel, nt, mt_pre);
  for (int64_t chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
    const MtChunk c = mt_find(numel, nt, mt_pre, chunk);
    float* p = P[c.t] + c.off;
    const float* g = G[c.t] + c.off;
    float* m = Mo[c.t] + c.off;
    float* v = Vo[c.t] + c.off;
    float* s = Sh != nullptr ? Sh[c.t] + c.off : nullptr;
    constexpr int U = MT_CHUNK / MT_THREADS;
    float gv[U], pv[U], mv[U], vv[U], sv[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int i = threadIdx.x + k * MT_THREADS;
      if (i < c.n) {
        gv[k] = g[i]; pv[k] = p[i]; mv[k] = m[i]; vv[k] = v[i];
        sv[k] = s != nullptr ? s[i] : 0.0f;
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int i = threadIdx.x + k * MT_THREADS;
      if (i < c.n) {
        const float gk = gv[k] * gscale;
        float pk = pv[k] * (1.0f - lr * wd);
        const float mk = mv[k] + (gk - mv[k]) * (1.0f - b1);
        const float vk = vv[k] * b2 + (1.0f - b2) * gk * gk;
        const float denom = sqrtf(vk) / bc2_sqrt + eps;
        pk = pk - (lr / bc1) * (mk / denom);
        p[i] = pk; m[i] = mk; v[i] = vk;
        if (s != nullptr) s[i] = ema_w < 0.5f ? sv[k] + ema_w * (pk - sv[k]) : pk - (pk - sv[k]) * (1.0f - ema_w);
      }
    }
  }
}


// ---- ensemble losses --------------------------------------------------------------------------------
// EnsembleKarrasModule.loss_fn (karras/karrasmodule_new.py:963-1149): E noisy copies per sample, ONE network call on
// B*E rows, then an ensemble-aware loss (custom_losses.py:536-865) reduced to a scalar that is multiplied by
// mean_b(lambda(sigma_b)).  x_noised[b][e] = x[b] + sigma[b] * noise[b][e]   (karrasmodule_new.py:1014-1040)
__global__ void __launch_bounds__(256) ens_noise_add_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                                             const float* __restrict__ sigma, float* __restrict__ out, int E,
                                                             int64_t CS, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t be = i / CS, n = i - be * CS;
    const int b = (int)(be / E);
    out[i] = x[(int64_t)b * CS + n] + sigma[b] * noise[i];
  }
}

__device__ __forceinline__ float sgn_f(float v) { return (float)(v > 0.0f) - (float)(v < 0.0f); }

// One thread owns V consecutive elements n of one sample b and walks its E members: D_e = c_out F[b][e][n] + c_skip x_noised,
//   kind 0 / 1: sum_e l(D_e - x) * keep * s1[b]            (Huber delta=1 / MSE, EnsembleAwareHuberLoss / ...MSELoss)
//   kind 2    : s1[b] sum_e |D_e - x|  -  s2[b] sum_{i<j} |D_i - D_j|      (EnsembleAwareCRPSLoss, custom_losses.py:765-865)
// s1 / s2 carry every constant of the reference's reductions (1/(B E N), valid-pixel counts, mean lambda), computed per
// sample on the host side from sigma and the mask; dF = c_out * dL/dD in the same pass.
template <int E, int V>
__global__ void __launch_bounds__(256) ens_loss_kernel(const float* __restrict__ F, const float* __restrict__ x,
                                                        const float* __restrict__ noise, const float* __restrict__ sigma,
                                                        const float* __restrict__ c_out_v, const float* __restrict__ c_skip_v,
                                                        const float* __restrict__ s1_v, const float* __restrict__ s2_v,
                                                        const float* __restrict__ mask, float* __restrict__ loss_out,
                                                        float* __restrict__ dF, int B, int64_t CS, int64_t S, int mask_C, int kind) {
  const int64_t per_b = CS / V, total = (int64_t)B * per_b;
  float local = 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_b);
    const int64_t n = (i - (int64_t)b * per_b) * V;
    const float sg = sigma[b], c_out = c_out_v[b], c_skip = c_skip_v[b], s1 = s1_v[b];
    float xv[V], keep[V], D[E][V], g[E][V];
    if (V == 4) {
      const float4 t = *reinterpret_cast<const float4*>(x + (int64_t)b * CS + n);
      xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
    } else {
      xv[0] = x[(int64_t)b * CS + n];
    }
#pragma unroll
    for (int v = 0; v < V; ++v) keep[v] = 1.0f;
    if (mask != nullptr) {
      const int64_t mo = mask_C == 1 ? (int64_t)b * S + (n % S) : (int64_t)b * CS + n;
#pragma unroll
      for (int v = 0; v < V; ++v) keep[v] = 1.0f - mask[mo + v];
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int64_t o = ((int64_t)b * E + e) * CS + n;
      float f[V], z[V];
      if (V == 4) {
        const float4 a = *reinterpret_cast<const float4*>(F + o), c = *reinterpret_cast<const float4*>(noise + o);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
        z[0] = c.x; z[1] = c.y; z[2] = c.z; z[3] = c.w;
      } else {
        f[0] = F[o]; z[0] = noise[o];
      }
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float d = c_out * f[v] + c_skip * (xv[v] + sg * z[v]);
        D[e][v] = d;
        const float r = d - xv[v], a = fabsf(r);
        float l, gr;
        if (kind == 0) { l = a <= 1.0f ? 0.5f * r * r : a - 0.5f; gr = fminf(fmaxf(r, -1.0f), 1.0f); }
        else if (kind == 1) { l = r * r; gr = 2.0f * r; }
        else { l = a; gr = sgn_f(r); }
        local += s1 * (l * keep[v]);
        g[e][v] = s1 * (gr * keep[v]);
      }
    }
    if (kind == 2 && E > 1) {
      const float s2 = s2_v[b];
#pragma unroll
      for (int e = 0; e < E; ++e) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float sg_sum = 0.0f, pair = 0.0f;
#pragma unroll
          for (int j = 0; j < E; ++j) {
            const float d = D[e][v] - D[j][v];
            sg_sum += sgn_f(d);
            if (j > e) pair += fabsf(d);
          }
          local -= s2 * pair;
          g[e][v] -= s2 * sg_sum;
        }
      }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int64_t o = ((int64_t)b * E + e) * CS + n;
      if (V == 4) *reinterpret_cast<float4*>(dF + o) = make_float4(c_out * g[e][0], c_out * g[e][1], c_out * g[e][2], c_out * g[e][3]);
      else dF[o] = c_out * g[e][0];
    }
  }
  local = warp_sum(local);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(loss_out, s);
  }
}

template <int E>
static int ens_loss_launch(const float* F, const float* x, const float* noise, const float* sigma, const float* c_out,
                           const float* c_skip, const float* s1, const float* s2, const float* mask, float* loss_out, float* dF,
                           int B, int64_t CS, int64_t S, int mask_C, int kind, cudaStream_t st) {
  if constexpr (E <= 8) {   // 16-byte accesses; E > 8 would need > 128 registers per thread
    if (CS % 4 == 0 && S % 4 == 0) {
      DSK_LAUNCH((ens_loss_kernel<E, 4>), grid_for((int64_t)B * CS / 4, 256, 8), 256, 0, st, F, x, noise, sigma, c_out, c_skip, s1,
                 s2, mask, loss_out, dF, B, CS, S, mask_C, kind);
      return DSK_OK;
    }
  }
  DSK_LAUNCH((ens_loss_kernel<E, 1>), grid_for((int64_t)B * CS, 256, 8), 256, 0, st, F, x, noise, sigma, c_out, c_skip, s1, s2,
             mask, loss_out, dF, B, CS, S, mask_C, kind);
  return DSK_OK;
}

}  // namespace dsk

using namespace dsk;

extern "C" int dsk_edm_loss_fwd_bwd(const float* F, const float* x, const float* noise, const float* sigma,
                                    const float* mask, float* loss_out, float* dF, int B, int C, int64_t S,
                                    float sigma_data, int loss_kind, void* stream) {
  DSK_REQUIRE(F && x && noise && sigma && loss_out && dF, "dsk_edm_loss_fwd_bwd: null pointer");
  DSK_REQUIRE(B > 0 && C > 0 && S > 0 && (loss_kind == 0 || loss_kind == 1), "dsk_edm_loss_fwd_bwd: bad arguments");
  const int grid = grid_for((int64_t)B * C * S, 256, 8);
  DSK_LAUNCH(edm_loss_kernel, grid, 256, 0, as_stream(stream), F, x, noise, sigma, mask, loss_out, dF, B, (int64_t)C * S,
             sigma_data, loss_kind, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, 1.0f);
  return DSK_OK;
}

extern "C" int dsk_precond_loss_fwd_bwd_huber(const float* F, const float* x, const float* noise, const float* sigma,
                                              const float* c_out, const float* c_skip, const float* weight, const float* mask,
                                              float* loss_out, float* dF, int B, int C, int64_t S, int loss_kind, float delta,
                                              void* stream) {
  DSK_REQUIRE(F && x && noise && sigma && c_out && c_skip && weight && loss_out && dF, "dsk_precond_loss_fwd_bwd: null pointer");
  DSK_REQUIRE(B > 0 && C > 0 && S > 0 && (loss_kind == 0 || loss_kind == 1) && delta > 0.0f, "dsk_precond_loss_fwd_bwd: bad arguments");
  const int grid = grid_for((int64_t)B * C * S, 256, 8);
  DSK_LAUNCH(edm_loss_kernel, grid, 256, 0, as_stream(stream), F, x, noise, sigma, mask, loss_out, dF, B, (int64_t)C * S, 0.0f,
             loss_kind, c_out, c_skip, weight, delta);
  return DSK_OK;
}
extern "C" int dsk_precond_loss_fwd_bwd(const float* F, const float* x, const float* noise, const float* sigma,
                                        const float* c_out, const float* c_skip, const float* weight, const float* mask,
                                        float* loss_out, float* dF, int B, int C, int64_t S, int loss_kind, void* stream) {
  return dsk_precond_loss_fwd_bwd_huber(F, x, noise, sigma, c_out, c_skip, weight, mask, loss_out, dF, B, C, S, loss_kind, 1.0f, stream);
}


extern "C" int dsk_precond_loss_rows_huber(const float* F, const float* x, const float* noise, const float* sigma, const float* c_out,
                                           const float* c_skip, const float* weight, const float* mask, float* loss_b, float* dF,
                                           int B, int C, int64_t S, int loss_kind, float delta, void* stream);
extern "C" int dsk_precond_loss_rows(const float* F, const float* x, const float* noise, const float* sigma, const float* c_out,
                                     const float* c_skip, const float* weight, const float* mask, float* loss_b, float* dF,
                                     int B, int C, int64_t S, int loss_kind, void* stream) {
  return dsk_precond_loss_rows_huber(F, x, noise, sigma, c_out, c_skip, weight, mask, loss_b, dF, B, C, S, loss_kind, 1.0f, stream);
}
extern "C" int dsk_precond_loss_rows_huber(const float* F, const float* x, const float* noise, const float* sigma, const float* c_out,
                                           const float* c_skip, const float* weight, const float* mask, float* loss_b, float* dF,
                                           int B, int C, int64_t S, int loss_kind, float delta, void* stream) {
  DSK_REQUIRE(F && x && noise && sigma && c_out && c_skip && weight && loss_b && dF, "dsk_precond_loss_rows: null pointer");
  DSK_REQUIRE(B > 0 && B <= 65535 && C > 0 && S > 0 && (loss_kind == 0 || loss_kind == 1) && delta > 0.0f, "dsk_precond_loss_rows: bad arguments");
  const int64_t CS = (int64_t)C * S;
  int per_b = (int)((CS + 255) / 256);
  const int cap = (8 * DSK_NUM_SMS + B - 1) / B;
  if (per_b > cap) per_b = cap;
  if (per_b < 1) per_b = 1;
  DSK_LAUNCH(precond_loss_rows_kernel, dim3(per_b, B), 256, 0, as_stream(stream), F, x, noise, sigma, mask, loss_b, dF, B, CS,
             loss_kind, c_out, c_skip, weight, delta);
  return DSK_OK;
}

extern "C" int dsk_ensemble_noise_add(const float* x, const float* noise, const float* sigma, float* out, int B, int E,
                                      int64_t CS, void* stream) {
  DSK_REQUIRE(x && noise && sigma && out, "dsk_ensemble_noise_add: null pointer");
  DSK_REQUIRE(B > 0 && E > 0 && CS > 0, "dsk_ensemble_noise_add: bad arguments");
  const int64_t total = (int64_t)B * E * CS;
  DSK_LAUNCH(ens_noise_add_kernel, grid_for(total, 256, 8), 256, 0, as_stream(stream), x, noise, sigma, out, E, CS, total);
  return DSK_OK;
}

extern "C" int dsk_ensemble_loss_fwd_bwd(const float* F, const float* x, const float* noise, const float* sigma,
                                         const float* c_out, const float* c_skip, const float* s1, const float* s2,
                                         const float* mask, int mask_C, float* loss_out, float* dF, int B, int E, int C,
                                         int64_t S, int loss_kind, void* stream) {
  DSK_REQUIRE(F && x && noise && sigma && c_out && c_skip && s1 && loss_out && dF, "dsk_ensemble_loss_fwd_bwd: null pointer");
  DSK_REQUIRE(B > 0 && C > 0 && S > 0 && loss_kind >= 0 && loss_kind <= 2, "dsk_ensemble_loss_fwd_bwd: bad arguments");
  DSK_REQUIRE(E >= 1 && E <= 16, "dsk_ensemble_loss_fwd_bwd: ensemble size %d outside 1..16", E);
  DSK_REQUIRE(loss_kind != 2 || s2 != nullptr, "dsk_ensemble_loss_fwd_bwd: CRPS needs the pair-term scale s2");
  DSK_REQUIRE(mask == nullptr || mask_C == 1 || mask_C == C, "dsk_ensemble_loss_fwd_bwd: mask channels must be 1 or C");
  DSK_REQUIRE(mask == nullptr || loss_kind != 2, "dsk_ensemble_loss_fwd_bwd: the CRPS mask acts through s1 / s2 only");
  const int64_t CS = (int64_t)C * S;
  cudaStream_t st = as_stream(stream);
#define DSK_ENS_CASE(e) \
  case e: return ens_loss_launch<e>(F, x, noise, sigma, c_out, c_skip, s1, s2, mask, loss_out, dF, B, CS, S, mask_C, loss_kind, st);
  switch (E) {
    DSK_ENS_CASE(1) DSK_ENS_CASE(2) DSK_ENS_CASE(3) DSK_ENS_CASE(4) DSK_ENS_CASE(5) DSK_ENS_CASE(6) DSK_ENS_CASE(7) DSK_ENS_CASE(8)
    DSK_ENS_CASE(9) DSK_ENS_CASE(10) DSK_ENS_CASE(11) DSK_ENS_CASE(12) DSK_ENS_CASE(13) DSK_ENS_CASE(14) DSK_ENS_CASE(15)
    DSK_ENS_CASE(16)
  }
#undef DSK_ENS_CASE
  return DSK_ERR_ARG;
}

static inline int multi_grid(int ntensors, int64_t max_numel) {
  // upper bound of the chunk count, capped at a few waves of the machine
  int64_t chunks = (int64_t)ntensors * ((max_numel + MT_CHUNK - 1) / MT_CHUNK);
  const int64_t cap = 8LL * DSK_NUM_SMS;
  if (chunks > cap) chunks = cap;
  return chunks < 1 ? 1 : (int)chunks;
}

extern "C" int dsk_ema_update(float* const* shadow, const float* const* param, const int64_t* numel, int ntensors,
                              int64_t max_numel, float beta, void* stream) {
  DSK_REQUIRE(shadow && param && numel && ntensors > 0 && ntensors <= MT_MAX_TENSORS && max_numel > 0, "dsk_ema_update: bad arguments");
  DSK_LAUNCH(ema_kernel, multi_grid(ntensors, max_numel), MT_THREADS, (size_t)(ntensors + 1) * sizeof(int64_t), as_stream(stream), shadow,
             param, numel, ntensors, 1.0f - beta);
  return DSK_OK;
}

extern "C" int dsk_adamw_ema_step(float* const* p, const float* const* g, float* const* m, float* const* v,
                                  float* const* shadow, const int64_t* numel, int ntensors, int64_t max_numel, float lr,
                                  float beta1, float beta2, float eps, float wd, int step, float ema_beta,
                                  float grad_scale, void* stream) {
  DSK_REQUIRE(p && g && m && v && numel && ntensors > 0 && ntensors <= MT_MAX_TENSORS && max_numel > 0 && step > 0,
              "dsk_adamw_ema_step: bad arguments");
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.0f - powf(beta2, (float)step));
  DSK_LAUNCH(adamw_ema_kernel, multi_grid(ntensors, max_numel), MT_THREADS, (size_t)(ntensors + 1) * sizeof(int64_t), as_stream(stream),
             p, g, m, v, shadow, numel, ntensors, lr, beta1, beta2, eps, wd, bc1, bc2s, 1.0f - ema_beta, grad_scale);
  return DSK_OK;
}
