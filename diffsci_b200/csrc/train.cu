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
                                                        float* __restrict__ dF, int B, int64_t CS, float sd, int kind) {
  const int64_t N = (int64_t)B * CS;
  const float invN = 1.0f / (float)N;
  float local = 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / CS);
    const float sg = sigma[b];
    const float sum = sg * sg + sd * sd, rt = sqrtf(sum);
    const float c_skip = (sd * sd) / sum, c_out = (sg * sd) / rt;
    const float lam = sum / ((sg * sd) * (sg * sd));
    const float xv = x[i];
    const float xn = xv + sg * noise[i];
    const float D = c_out * F[i] + c_skip * xn;
    const float r = D - xv;
    float l, g;
    if (kind == 0) {  // Huber, delta = 1 (torch.nn.HuberLoss default, karrasmodule.py:541-542)
      const float a = fabsf(r);
      l = a <= 1.0f ? 0.5f * r * r : a - 0.5f;
      g = fminf(fmaxf(r, -1.0f), 1.0f);
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

// One launch for every parameter tensor: blockIdx.y = tensor, blockIdx.x strides its elements.
__global__ void __launch_bounds__(256) ema_kernel(float* const* __restrict__ shadow, const float* const* __restrict__ param,
                                                   const int64_t* __restrict__ numel, float w) {
  const int t = blockIdx.y;
  float* s = shadow[t];
  const float* p = param[t];
  const int64_t n = numel[t];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    // torch.lerp(s, p, w) = s + w*(p - s) for w < 0.5, else p - (p - s)*(1 - w)
    const float sv = s[i], pv = p[i];
    s[i] = w < 0.5f ? sv + w * (pv - sv) : pv - (pv - sv) * (1.0f - w);
  }
}

__global__ void __launch_bounds__(256) adamw_ema_kernel(float* const* __restrict__ P, const float* const* __restrict__ G,
                                                         float* const* __restrict__ Mo, float* const* __restrict__ Vo,
                                                         float* const* __restrict__ Sh, const int64_t* __restrict__ numel,
                                                         float lr, float b1, float b2, float eps, float wd, float bc1,
                                                         float bc2_sqrt, float ema_w, float gscale) {
  const int t = blockIdx.y;
  float* p = P[t];
  const float* g = G[t];
  float* m = Mo[t];
  float* v = Vo[t];
  float* s = Sh != nullptr ? Sh[t] : nullptr;
  const int64_t n = numel[t];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gv = g[i] * gscale;
    float pv = p[i] * (1.0f - lr * wd);
    const float mv = m[i] + (gv - m[i]) * (1.0f - b1);
    const float vv = v[i] * b2 + (1.0f - b2) * gv * gv;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pv = pv - (lr / bc1) * (mv / denom);
    p[i] = pv; m[i] = mv; v[i] = vv;
    if (s != nullptr) {
      const float sv = s[i];
      s[i] = ema_w < 0.5f ? sv + ema_w * (pv - sv) : pv - (pv - sv) * (1.0f - ema_w);
    }
  }
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
             sigma_data, loss_kind);
  return DSK_OK;
}

static inline dim3 multi_grid(int ntensors, int64_t max_numel) {
  int64_t bx = (max_numel + 255) / 256;
  int64_t cap = (4LL * DSK_NUM_SMS + ntensors - 1) / ntensors;
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3((unsigned)bx, (unsigned)ntensors);
}

extern "C" int dsk_ema_update(float* const* shadow, const float* const* param, const int64_t* numel, int ntensors,
                              int64_t max_numel, float beta, void* stream) {
  DSK_REQUIRE(shadow && param && numel && ntensors > 0 && ntensors <= 65535 && max_numel > 0, "dsk_ema_update: bad arguments");
  DSK_LAUNCH(ema_kernel, multi_grid(ntensors, max_numel), 256, 0, as_stream(stream), shadow, param, numel, 1.0f - beta);
  return DSK_OK;
}

extern "C" int dsk_adamw_ema_step(float* const* p, const float* const* g, float* const* m, float* const* v,
                                  float* const* shadow, const int64_t* numel, int ntensors, int64_t max_numel, float lr,
                                  float beta1, float beta2, float eps, float wd, int step, float ema_beta,
                                  float grad_scale, void* stream) {
  DSK_REQUIRE(p && g && m && v && numel && ntensors > 0 && ntensors <= 65535 && max_numel > 0 && step > 0,
              "dsk_adamw_ema_step: bad arguments");
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.0f - powf(beta2, (float)step));
  DSK_LAUNCH(adamw_ema_kernel, multi_grid(ntensors, max_numel), 256, 0, as_stream(stream), p, g, m, v, shadow, numel, lr,
             beta1, beta2, eps, wd, bc1, bc2s, 1.0f - ema_beta, grad_scale);
  return DSK_OK;
}
