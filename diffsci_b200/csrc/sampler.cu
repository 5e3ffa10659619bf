// sampler.cu -- K7: EDM preconditioning and fused integrator stages (HBM-bandwidth kernels).
//
// One elementwise pass per network evaluation does everything the reference spreads over
// ~30 tiny ATen kernels and 2-3 host syncs (karras/karrasmodule.py:690-733,
// karras/schedulers.py:247-274, karras/integrators.py:29-113): D = c_out*F + c_skip*x, the
// score, the probability-flow RHS, the Euler/Heun/EM/Karras update, the optional history
// write, the churn noise (Philox) and the c_in scaling of the NEXT network input.
// All step scalars come from a device-resident table row selected by a device counter, so a
// single captured CUDA graph is replayed for every step.
#include "common.cuh"
#include "philox.cuh"

namespace dsk {

struct Precond {
  float c_in, c_out, c_skip;
};

// EDMPreconditioner (karras/preconditioners.py:35-53), same operation order in fp32.
// kind 0: EDM ; kind 1: NullPreconditioner (c_in = c_out = 1, c_skip = 0; preconditioners.py:139-161)
__device__ __forceinline__ Precond edm_precond(float sigma, float sd, int kind = 0) {
  if (kind == 1) return Precond{1.0f, 1.0f, 0.0f};
  float sum = sigma * sigma + sd * sd;
  float rt = sqrtf(sum);
  Precond p;
  p.c_skip = (sd * sd) / sum;
  p.c_out = (sigma * sd) / rt;
  p.c_in = 1.0f / rt;
  return p;
}

// Scheduler.rhs for the EDM functions (schedulers.py:262-268): -(sigma*sigma')*score with
// score = (D - x)/sigma^2 (karrasmodule.py:733); returns score through *sc.
__device__ __forceinline__ float edm_rhs(float F, float x, float sigma, const Precond& p, float* sc) {
  float D = p.c_out * F + p.c_skip * x;
  float s = (D - x) / (sigma * sigma);
  *sc = s;
  return -(sigma * 1.0f) * s;
}

struct StageArgs {
  float* x;
  float* x_aux;
  float* r1;
  const void* F;
  void* xin;
  float* cnoise;
  const float* tab;
  const int* row;
  const float* noise;
  float* hist;
  uint64_t seed;
  int B, C;
  int64_t S;
  float sigma_data, sigma_max;
  int precond;   // 0 EDM, 1 Null
  // conditional path: the network input rows hold xin_ld >= C channels (PUNetGCond's channel-concatenated
  // conditioning lives in channels C..xin_ld-1 and is written once per run by the host); cfg != 0: the network runs a
  // 2B batch (rows [0,B) unconditional, [B,2B) conditional) and F = (1-g) F_u + g F_c (karrasmodule.py:705-713).
  int xin_ld, cfg;
  float guidance;
  // inpainting (Scheduler.inpaint, karras/schedulers.py:91-122): after every completed step -- and at the start of the run --
  // the known region is re-imposed, x <- x (1 - m) + y[level] m, with y the forward (data -> noise) history [blend_rows, B, C, S]
  // of the known data, level = blend_rows - 1 - (steps completed).  Null: no blending.  The mask has mask_n elements and is
  // broadcast over the leading dimensions.
  const float* blend_y;
  const float* blend_mask;
  int64_t mask_n;
  int blend_rows;
};

template <int V> struct Vec;
template <> struct Vec<1> {
  static __device__ __forceinline__ void ld(const float* p, int64_t i, float* o) { o[0] = p[i]; }
  static __device__ __forceinline__ void st(float* p, int64_t i, const float* v) { p[i] = v[0]; }
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, int64_t i, float* o) { o[0] = __bfloat162float(p[i]); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, int64_t i, const float* v) { p[i] = __float2bfloat16_rn(v[0]); }
  static __device__ __forceinline__ void ld(const __half* p, int64_t i, float* o) { o[0] = __half2float(p[i]); }
  static __device__ __forceinline__ void st(__half* p, int64_t i, const float* v) { p[i] = __float2half_rn(v[0]); }
};
template <> struct Vec<4> {
  static __device__ __forceinline__ void ld(const float* p, int64_t i, float* o) {
    float4 v = *reinterpret_cast<const float4*>(p + i);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  static __device__ __forceinline__ void st(float* p, int64_t i, const float* v) {
    *reinterpret_cast<float4*>(p + i) = make_float4(v[0], v[1], v[2], v[3]);
  }
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, int64_t i, float* o) {
    uint2 raw = *reinterpret_cast<const uint2*>(p + i);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x), b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
    o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, int64_t i, const float* v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 raw;
    raw.x = *reinterpret_cast<uint32_t*>(&a);
    raw.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p + i) = raw;
  }
  static __device__ __forceinline__ void ld(const __half* p, int64_t i, float* o) {
    uint2 raw = *reinterpret_cast<const uint2*>(p + i);
    __half2 a = *reinterpret_cast<__half2*>(&raw.x), b = *reinterpret_cast<__half2*>(&raw.y);
    o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
  }
  static __device__ __forceinline__ void st(__half* p, int64_t i, const float* v) {
    __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
    uint2 raw;
    raw.x = *reinterpret_cast<uint32_t*>(&a);
    raw.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p + i) = raw;
  }
};

// V == 4 requires C == 1 == xin_ld (channels-last == NCHW) and (B*S) % 4 == 0.
template <typename T, int V, int STAGE>
__global__ void __launch_bounds__(256) sampler_stage_kernel(StageArgs a) {
  const int rowi = a.row[0];
  // run-time seed lives in device memory (row[1], row[2]) so that a captured graph can be replayed
  // with a new seed; the by-value seed is a compile/capture-time base.
  const uint64_t seed = a.seed ^ (((uint64_t)(uint32_t)a.row[2] << 32) | (uint64_t)(uint32_t)a.row[1]);
  const float* r = a.tab + (int64_t)rowi * DSK_TAB_COLS;
  const float t = r[DSK_TAB_T], dt = r[DSK_TAB_DT], that = r[DSK_TAB_THAT];
  const float sd = a.sigma_data;
  const int64_t N = (int64_t)a.B * a.C * a.S;
  const int64_t CS = (int64_t)a.C * a.S;
  const T* Fp = reinterpret_cast<const T*>(a.F);
  T* xin = reinterpret_cast<T*>(a.xin);

  // sigma of the evaluation that produced F, and of the evaluation being prepared
  float sig_eval, sig_next;
  bool writes_hist = false, do_prep = true;
  int hist_slot = rowi + 1, noise_row = rowi;
  float ncoef = 0.0f;
  // INIT at row r starts a run at step r (r > 0: Scheduler.propagate_partial, schedulers.py:177-217): its history slot is r
  if (STAGE == DSK_STAGE_INIT) { sig_eval = 1.0f; sig_next = that; writes_hist = true; hist_slot = rowi; ncoef = r[DSK_TAB_CHURN]; }
  else if (STAGE == DSK_STAGE_EULER) { sig_eval = t; sig_next = r[DSK_TAB_TNEXT]; writes_hist = true; }
  else if (STAGE == DSK_STAGE_HEUN_MID) { sig_eval = t; sig_next = t + dt; }
  else if (STAGE == DSK_STAGE_HEUN_FIN) { sig_eval = t + dt; sig_next = r[DSK_TAB_TNEXT]; writes_hist = true; }
  else if (STAGE == DSK_STAGE_HEUN_LAST) { sig_eval = t; sig_next = 0.f; writes_hist = true; do_prep = false; }
  else if (STAGE == DSK_STAGE_EM) { sig_eval = t; sig_next = r[DSK_TAB_TNEXT]; writes_hist = true; }
  else if (STAGE == DSK_STAGE_KARRAS_MID) { sig_eval = that; sig_next = t + dt; }
  else if (STAGE == DSK_STAGE_KARRAS_FIN) { sig_eval = t + dt; sig_next = r[DSK_TAB_TNEXT]; writes_hist = true; ncoef = r[DSK_TAB_COLS + DSK_TAB_CHURN]; noise_row = rowi + 1; }
  else { sig_eval = that; sig_next = 0.f; writes_hist = true; do_prep = false; }  // KARRAS_LAST
  if (sig_next <= 0.0f) do_prep = false;

  const Precond pe = edm_precond(sig_eval, sd, a.precond);
  const Precond pn = edm_precond(do_prep ? sig_next : 1.0f, sd, a.precond);
  const float lang = r[DSK_TAB_LANG], nstr = r[DSK_TAB_NOISE], sqdt = r[DSK_TAB_SQDT];
  const float dth = (t + dt) - that;  // Karras: dt_noise = (t+dt) - t_noise (integrators.py:107)

  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (do_prep && tid < a.B) {
    const float cn = a.precond == 1 ? sig_next : 0.5f * logf(sig_next);
    a.cnoise[tid] = cn;
    if (a.cfg) a.cnoise[a.B + tid] = cn;
  }
  const bool strided = (V == 1) && (a.C > 1 || a.xin_ld != a.C);

  for (int64_t i = tid * V; i < N; i += (int64_t)gridDim.x * blockDim.x * V) {
    // channels-last index of element i (NCHW order); identical when C == 1
    int64_t cl = i, cli = i;   // cl: index into F rows [.., C]; cli: index into network-input rows [.., xin_ld]
    if (strided) {
      int64_t b = i / CS, rem = i - b * CS;
      int64_t c = rem / a.S, s = rem - c * a.S;
      cl = (b * a.S + s) * a.C + c;
      cli = (b * a.S + s) * a.xin_ld + c;
    }
    float xv[V], fv[V], av[V], rv[V], zv[V], xo[V], pv[V];
    Vec<V>::ld(a.x, i, xv);
    if (STAGE != DSK_STAGE_INIT) {
      Vec<V>::ld(Fp, cl, fv);
      if (a.cfg) {
        float fc[V];
        Vec<V>::ld(Fp, N + cl, fc);
#pragma unroll
        for (int k = 0; k < V; ++k) fv[k] = (1.0f - a.guidance) * fv[k] + a.guidance * fc[k];
      }
    }
    if (STAGE == DSK_STAGE_HEUN_FIN || STAGE == DSK_STAGE_KARRAS_FIN) {
      Vec<V>::ld(a.x_aux, i, av);
      Vec<V>::ld(a.r1, i, rv);
    }
    const bool need_noise = (STAGE == DSK_STAGE_EM) ||
                            ((STAGE == DSK_STAGE_INIT || STAGE == DSK_STAGE_KARRAS_FIN) && ncoef != 0.0f);
    if (need_noise) {
      if (a.noise != nullptr) {
        Vec<V>::ld(a.noise + (int64_t)noise_row * N, i, zv);
      } else {
        float q[4];
        philox_normal4(seed, (uint32_t)noise_row, (uint64_t)(i >> 2), q);
        if (V == 4) {
#pragma unroll
          for (int k = 0; k < V; ++k) zv[k] = q[k];
        } else {
          zv[0] = q[i & 3];
        }
      }
    }
    float auxo[V], r1o[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float sc;
      if (STAGE == DSK_STAGE_INIT) {
        // x = white_noise * maximum_scale (karrasmodule.py:881); history[0] is pre-churn
        xo[k] = xv[k] * a.sigma_max;
        pv[k] = (ncoef != 0.0f) ? xo[k] + ncoef * zv[k] : xo[k];
      } else if (STAGE == DSK_STAGE_EULER) {
        xo[k] = xv[k] + dt * edm_rhs(fv[k], xv[k], sig_eval, pe, &sc);
        pv[k] = xo[k];
      } else if (STAGE == DSK_STAGE_HEUN_MID) {
        r1o[k] = edm_rhs(fv[k], xv[k], sig_eval, pe, &sc);
        auxo[k] = xv[k] + dt * r1o[k];
        pv[k] = auxo[k];
      } else if (STAGE == DSK_STAGE_HEUN_FIN) {
        float r2 = edm_rhs(fv[k], av[k], sig_eval, pe, &sc);
        xo[k] = xv[k] + (0.5f * (rv[k] + r2)) * dt;
        pv[k] = xo[k];
      } else if (STAGE == DSK_STAGE_HEUN_LAST) {
        float q1 = edm_rhs(fv[k], xv[k], sig_eval, pe, &sc);
        xo[k] = xv[k] + (0.5f * (q1 + q1)) * dt;
        pv[k] = xo[k];
      } else if (STAGE == DSK_STAGE_EM) {
        float q1 = edm_rhs(fv[k], xv[k], sig_eval, pe, &sc);
        q1 += -(lang * sc);  // Langevin drift (schedulers.py:269-274)
        xo[k] = (xv[k] + q1 * dt) + ((nstr * zv[k]) * sqdt);
        pv[k] = xo[k];
      } else if (STAGE == DSK_STAGE_KARRAS_MID) {
        r1o[k] = edm_rhs(fv[k], xv[k], sig_eval, pe, &sc);
        auxo[k] = xv[k] + dth * r1o[k];
        pv[k] = auxo[k];
      } else if (STAGE == DSK_STAGE_KARRAS_FIN) {
        float r2 = edm_rhs(fv[k], av[k], sig_eval, pe, &sc);
        xo[k] = xv[k] + (0.5f * (rv[k] + r2)) * dth;
        pv[k] = (ncoef != 0.0f) ? xo[k] + ncoef * zv[k] : xo[k];
      } else {  // KARRAS_LAST
        float q1 = edm_rhs(fv[k], xv[k], sig_eval, pe, &sc);
        xo[k] = xv[k] + dth * q1;
        pv[k] = xo[k];
      }
    }
    if (a.blend_y != nullptr && STAGE != DSK_STAGE_HEUN_MID && STAGE != DSK_STAGE_KARRAS_MID) {
      // the state after this stage sits at level rowi (INIT) / rowi + 1 of the schedule: blend with the matching row of the
      // known data's forward history, then redo what depends on it (churned state / next network input)
      const int level = a.blend_rows - 1 - (STAGE == DSK_STAGE_INIT ? rowi : rowi + 1);
      float yv[V];
      Vec<V>::ld(a.blend_y + (int64_t)level * N, i, yv);
      if (STAGE == DSK_STAGE_INIT && writes_hist && a.hist != nullptr)     // history[0] is the state BEFORE the first blend
        Vec<V>::st(a.hist + (int64_t)hist_slot * N, i, xo);                //   (schedulers.py:105-111)
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float m = a.blend_mask[(i + k) % a.mask_n];
        const float churn = pv[k] - xo[k];                 // Karras churn perturbation added on top of the step result
        xo[k] = xo[k] * (1.0f - m) + yv[k] * m;
        pv[k] = xo[k] + churn;
      }
    }
    if (STAGE == DSK_STAGE_HEUN_MID || STAGE == DSK_STAGE_KARRAS_MID) {
      Vec<V>::st(a.x_aux, i, auxo);
      Vec<V>::st(a.r1, i, r1o);
    } else {
      if (writes_hist && a.hist != nullptr && !(STAGE == DSK_STAGE_INIT && a.blend_y != nullptr))
        Vec<V>::st(a.hist + (int64_t)hist_slot * N, i, xo);
      // the state carried to the next step includes the churn perturbation (x_hat)
      Vec<V>::st(a.x, i, (STAGE == DSK_STAGE_INIT || STAGE == DSK_STAGE_KARRAS_FIN) ? pv : xo);
    }
    if (do_prep) {
#pragma unroll
      for (int k = 0; k < V; ++k) pv[k] = pn.c_in * pv[k];
      Vec<V>::st(xin, cli, pv);
      if (a.cfg) Vec<V>::st(xin, (int64_t)a.B * a.S * a.xin_ld + cli, pv);
    }
  }
}

__global__ void advance_kernel(int* row) { *row += 1; }

// xin[b, s, c] = c_in[b] * x[b, c, s]  -- any preconditioner (karrasmodule.py:690-702)
// ld >= C: channel stride of the network-input rows; dup: also write the second half of a 2B (CFG) batch.
template <typename T>
__global__ void __launch_bounds__(256) precond_scale_kernel(const float* __restrict__ x, const float* __restrict__ c_in,
                                                             T* __restrict__ xin, int B, int C, int64_t S, int ld, int dup) {
  const int64_t N = (int64_t)B * C * S, CS = (int64_t)C * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / CS, rem = i - b * CS;
    int64_t c = rem / S, s = rem - c * S;
    const T v = from_f32<T>(c_in[b] * x[i]);
    xin[(b * S + s) * ld + c] = v;
    if (dup) xin[((b + B) * S + s) * ld + c] = v;
  }
}

// F_u <- (1 - g) F_u + g F_c over n elements (classifier-free guidance mix, karrasmodule.py:711-713)
template <typename T>
__global__ void __launch_bounds__(256) cfg_mix_kernel(T* __restrict__ Fu, const T* __restrict__ Fc, float g, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    Fu[i] = from_f32<T>((1.0f - g) * to_f32<T>(Fu[i]) + g * to_f32<T>(Fc[i]));
}

// D = c_out[b]*F + c_skip[b]*x ; score = (D - x)/sigma[b]^2  (karrasmodule.py:717-733)
template <typename T>
__global__ void __launch_bounds__(256) precond_denoise_kernel(const T* __restrict__ F, const float* __restrict__ x,
                                                               const float* __restrict__ c_out, const float* __restrict__ c_skip,
                                                               const float* __restrict__ sigma, float* __restrict__ D,
                                                               float* __restrict__ score, int B, int C, int64_t S) {
  const int64_t N = (int64_t)B * C * S, CS = (int64_t)C * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / CS, rem = i - b * CS;
    int64_t c = rem / S, s = rem - c * S;
    float xv = x[i];
    float d = c_out[b] * to_f32<T>(F[(b * S + s) * C + c]) + c_skip[b] * xv;
    if (D != nullptr) D[i] = d;
    if (score != nullptr) {
      float sg = sigma[b];
      score[i] = (d - xv) / (sg * sg);
    }
  }
}

template <typename T, int V>
static int launch_stage(int stage, const StageArgs& a, cudaStream_t st) {
  const int64_t N = (int64_t)a.B * a.C * a.S;
  const int grid = grid_for(N / V, 256, 16);
#define CASE(SG)                                                          \
  case SG:                                                                \
    DSK_LAUNCH((sampler_stage_kernel<T, V, SG>), grid, 256, 0, st, a);    \
    break;
  switch (stage) {
    CASE(DSK_STAGE_INIT)
    CASE(DSK_STAGE_EULER)
    CASE(DSK_STAGE_HEUN_MID)
    CASE(DSK_STAGE_HEUN_FIN)
    CASE(DSK_STAGE_HEUN_LAST)
    CASE(DSK_STAGE_EM)
    CASE(DSK_STAGE_KARRAS_MID)
    CASE(DSK_STAGE_KARRAS_FIN)
    CASE(DSK_STAGE_KARRAS_LAST)
    default:
      set_error("dsk_sampler_stage: unknown stage %d", stage);
      return DSK_ERR_ARG;
  }
#undef CASE
  return DSK_OK;
}

}  // namespace dsk

using namespace dsk;

extern "C" int dsk_sampler_stage(int stage, float* x, float* x_aux, float* r1, const void* F, void* xin,
                                 float* cnoise, const float* tab, const int* row, const float* noise, uint64_t seed,
                                 float* hist, int B, int C, int64_t S, float sigma_data, float sigma_max,
                                 int precond_kind, int act_dtype, void* stream) {
  return dsk_sampler_stage_cond(stage, x, x_aux, r1, F, xin, cnoise, tab, row, noise, seed, hist, B, C, S, sigma_data,
                                sigma_max, precond_kind, act_dtype, C, 0, 1.0f, stream);
}

extern "C" int dsk_sampler_stage_cond(int stage, float* x, float* x_aux, float* r1, const void* F, void* xin,
                                      float* cnoise, const float* tab, const int* row, const float* noise, uint64_t seed,
                                      float* hist, int B, int C, int64_t S, float sigma_data, float sigma_max,
                                      int precond_kind, int act_dtype, int xin_ld, int cfg, float guidance,
                                      void* stream) {
  return dsk_sampler_stage_blend(stage, x, x_aux, r1, F, xin, cnoise, tab, row, noise, seed, hist, B, C, S, sigma_data, sigma_max,
                                 precond_kind, act_dtype, xin_ld, cfg, guidance, nullptr, nullptr, 0, 0, stream);
}

extern "C" int dsk_sampler_stage_blend(int stage, float* x, float* x_aux, float* r1, const void* F, void* xin,
                                       float* cnoise, const float* tab, const int* row, const float* noise, uint64_t seed,
                                       float* hist, int B, int C, int64_t S, float sigma_data, float sigma_max,
                                       int precond_kind, int act_dtype, int xin_ld, int cfg, float guidance,
                                       const float* blend_y, const float* blend_mask, int64_t mask_n, int blend_rows,
                                       void* stream) {
  DSK_REQUIRE(xin_ld >= C, "dsk_sampler_stage_cond: xin_ld=%d < C=%d", xin_ld, C);
  DSK_REQUIRE(blend_y == nullptr || (blend_mask != nullptr && mask_n > 0 && blend_rows > 1 && ((int64_t)B * C * S) % mask_n == 0),
              "dsk_sampler_stage_blend: bad blend arguments");
  DSK_REQUIRE(x && tab && row && cnoise, "dsk_sampler_stage: null x/tab/row/cnoise");
  DSK_REQUIRE(B > 0 && C > 0 && S > 0, "dsk_sampler_stage: bad shape B=%d C=%d S=%lld", B, C, (long long)S);
  DSK_REQUIRE(stage == DSK_STAGE_INIT || F != nullptr, "dsk_sampler_stage: F is null");
  DSK_REQUIRE(precond_kind == 0 || precond_kind == 1, "dsk_sampler_stage: bad precond_kind %d", precond_kind);
  const bool preps = !(stage == DSK_STAGE_HEUN_LAST || stage == DSK_STAGE_KARRAS_LAST);
  DSK_REQUIRE(!preps || xin != nullptr, "dsk_sampler_stage: xin is null");
  const bool needs_aux = stage == DSK_STAGE_HEUN_MID || stage == DSK_STAGE_HEUN_FIN ||
                         stage == DSK_STAGE_KARRAS_MID || stage == DSK_STAGE_KARRAS_FIN;
  DSK_REQUIRE(!needs_aux || (x_aux && r1), "dsk_sampler_stage: x_aux/r1 are null");
  DSK_REQUIRE(act_dtype == DSK_F32 || is_h16(act_dtype), "dsk_sampler_stage: bad dtype %d", act_dtype);
  StageArgs a{x, x_aux, r1, F, xin, cnoise, tab, row, noise, hist, seed, B, C, S, sigma_data, sigma_max, precond_kind,
              xin_ld, cfg ? 1 : 0, guidance, blend_y, blend_mask, mask_n, blend_rows};
  const int64_t N = (int64_t)B * C * S;
  const bool vec = (C == 1) && (xin_ld == 1) && (N % 4 == 0);
  cudaStream_t st = as_stream(stream);
  if (act_dtype == DSK_F32) return vec ? launch_stage<float, 4>(stage, a, st) : launch_stage<float, 1>(stage, a, st);
  if (act_dtype == DSK_F16) return vec ? launch_stage<__half, 4>(stage, a, st) : launch_stage<__half, 1>(stage, a, st);
  return vec ? launch_stage<__nv_bfloat16, 4>(stage, a, st) : launch_stage<__nv_bfloat16, 1>(stage, a, st);
}

extern "C" int dsk_sampler_advance(int* row, void* stream) {
  DSK_REQUIRE(row, "dsk_sampler_advance: null row");
  DSK_LAUNCH(advance_kernel, 1, 1, 0, as_stream(stream), row);
  return DSK_OK;
}

extern "C" int dsk_precond_scale(const float* x, const float* c_in, void* xin, int B, int C, int64_t S, int act_dtype,
                                 void* stream) {
  return dsk_precond_scale_cond(x, c_in, xin, B, C, S, act_dtype, C, 0, stream);
}

extern "C" int dsk_precond_scale_cond(const float* x, const float* c_in, void* xin, int B, int C, int64_t S,
                                      int act_dtype, int xin_ld, int dup, void* stream) {
  DSK_REQUIRE(x && c_in && xin, "dsk_precond_scale: null pointer");
  DSK_REQUIRE(B > 0 && C > 0 && S > 0 && xin_ld >= C, "dsk_precond_scale: bad shape");
  const int grid = grid_for((int64_t)B * C * S, 256, 16);
  if (act_dtype == DSK_F32)
    DSK_LAUNCH(precond_scale_kernel<float>, grid, 256, 0, as_stream(stream), x, c_in, (float*)xin, B, C, S, xin_ld, dup);
  else if (act_dtype == DSK_BF16)
    DSK_LAUNCH(precond_scale_kernel<__nv_bfloat16>, grid, 256, 0, as_stream(stream), x, c_in, (__nv_bfloat16*)xin, B, C, S,
               xin_ld, dup);
  else if (act_dtype == DSK_F16)
    DSK_LAUNCH(precond_scale_kernel<__half>, grid, 256, 0, as_stream(stream), x, c_in, (__half*)xin, B, C, S, xin_ld, dup);
  else
    DSK_REQUIRE(false, "dsk_precond_scale: bad dtype %d", act_dtype);
  return DSK_OK;
}

extern "C" int dsk_cfg_mix(void* F_uncond, const void* F_cond, float guidance, int64_t n, int act_dtype, void* stream) {
  DSK_REQUIRE(F_uncond && F_cond && n > 0, "dsk_cfg_mix: bad arguments");
  const int grid = grid_for(n, 256, 16);
  if (act_dtype == DSK_F32)
    DSK_LAUNCH(cfg_mix_kernel<float>, grid, 256, 0, as_stream(stream), (float*)F_uncond, (const float*)F_cond, guidance, n);
  else if (act_dtype == DSK_BF16)
    DSK_LAUNCH(cfg_mix_kernel<__nv_bfloat16>, grid, 256, 0, as_stream(stream), (__nv_bfloat16*)F_uncond,
               (const __nv_bfloat16*)F_cond, guidance, n);
  else if (act_dtype == DSK_F16)
    DSK_LAUNCH(cfg_mix_kernel<__half>, grid, 256, 0, as_stream(stream), (__half*)F_uncond, (const __half*)F_cond, guidance, n);
  else
    DSK_REQUIRE(false, "dsk_cfg_mix: bad dtype %d", act_dtype);
  return DSK_OK;
}

extern "C" int dsk_precond_denoise(const void* F, const float* x, const float* c_out, const float* c_skip,
                                   const float* sigma, float* D, float* score, int B, int C, int64_t S, int act_dtype,
                                   void* stream) {
  DSK_REQUIRE(F && x && c_out && c_skip && (D || score), "dsk_precond_denoise: null pointer");
  DSK_REQUIRE(score == nullptr || sigma != nullptr, "dsk_precond_denoise: score needs sigma");
  DSK_REQUIRE(B > 0 && C > 0 && S > 0, "dsk_precond_denoise: bad shape");
  const int grid = grid_for((int64_t)B * C * S, 256, 16);
  if (act_dtype == DSK_F32)
    DSK_LAUNCH(precond_denoise_kernel<float>, grid, 256, 0, as_stream(stream), (const float*)F, x, c_out, c_skip, sigma, D, score, B, C, S);
  else if (act_dtype == DSK_BF16)
    DSK_LAUNCH(precond_denoise_kernel<__nv_bfloat16>, grid, 256, 0, as_stream(stream), (const __nv_bfloat16*)F, x, c_out, c_skip, sigma, D, score, B, C, S);
  else if (act_dtype == DSK_F16)
    DSK_LAUNCH(precond_denoise_kernel<__half>, grid, 256, 0, as_stream(stream), (const __half*)F, x, c_out, c_skip, sigma, D, score, B, C, S);
  else
    DSK_REQUIRE(false, "dsk_precond_denoise: bad dtype %d", act_dtype);
  return DSK_OK;
}
