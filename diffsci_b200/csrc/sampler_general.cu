// sampler_general.cu -- table-driven integrator stages for ANY scheduler / preconditioner pair (SURVEY 8f-3 on the graph
// engine): VP / VE / SR3 / custom objects, i.e. the non-constant-scaling branch of Scheduler.rhs (karras/schedulers.py:275-293)
// with any KarrasPreconditioner (karras/preconditioners.py:30-161).
//
// For fixed t, the drift is linear in the state and the network output:
//     rhs(x, t) = (s'/s) x + Bm * score(x/s, sigma),   score(z, sigma) = (c_out F + (c_skip - 1) z) / sigma^2,   F = net(c_in z, c_noise)
//              = P x + Q F,     P = s'/s + Bm (c_skip - 1) / (sigma^2 s),   Q = Bm c_out / sigma^2,   network input = (c_in / s) x
// (Bm = -s sigma' sigma [or the scheduler's pf_score_multiplier], minus langevin_factor / s on stochastic steps.)  The host
// evaluates (P, Q, c_in/s, c_noise) for both evaluation points of every step with the scheduler's and the preconditioner's own
// objects; the kernel is then one elementwise pass per network evaluation that knows nothing about the family, replayed as a
// captured graph like sampler.cu (engine.GeneralSamplerEngine; DSK_GENERAL_ENGINE=0 falls back to the Integrator.step seam).
#include "common.cuh"
#include "philox.cuh"

namespace dsk {

struct GenArgs {
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
  int B, C;
  int64_t S;
  float x_scale;      // INIT: x <- x_scale * x (the scheduler's maximum_scale)
  int xin_ld;
};

template <typename T, int STAGE>
__global__ void __launch_bounds__(256) general_stage_kernel(GenArgs a) {
  const int rowi = a.row[0];
  const uint64_t seed = ((uint64_t)(uint32_t)a.row[2] << 32) | (uint64_t)(uint32_t)a.row[1];
  const float* r = a.tab + (int64_t)rowi * DSK_GTAB_COLS;
  const float* rn = r + DSK_GTAB_COLS;                                   // next step's row (the table has one padding row)
  const float dt = r[DSK_GTAB_DT];
  const int64_t N = (int64_t)a.B * a.C * a.S, CS = (int64_t)a.C * a.S;
  const T* Fp = reinterpret_cast<const T*>(a.F);
  T* xin = reinterpret_cast<T*>(a.xin);

  float P = 0.f, Q = 0.f, xs = 0.f, cn = 0.f;      // rhs coefficients of the evaluation that produced F; input scale / c_noise being prepared
  bool prep = true, writes_hist = false;
  int hist_slot = rowi + 1;
  if (STAGE == DSK_GSTAGE_INIT) { xs = r[DSK_GTAB_XS1]; cn = r[DSK_GTAB_CN1]; writes_hist = true; hist_slot = rowi; }   // partial runs start at row > 0
  else if (STAGE == DSK_GSTAGE_STEP1) { P = r[DSK_GTAB_P1]; Q = r[DSK_GTAB_Q1]; xs = rn[DSK_GTAB_XS1]; cn = rn[DSK_GTAB_CN1]; writes_hist = true; }
  else if (STAGE == DSK_GSTAGE_HEUN_MID) { P = r[DSK_GTAB_P1]; Q = r[DSK_GTAB_Q1]; xs = r[DSK_GTAB_XS2]; cn = r[DSK_GTAB_CN2]; }
  else { P = r[DSK_GTAB_P2]; Q = r[DSK_GTAB_Q2]; xs = rn[DSK_GTAB_XS1]; cn = rn[DSK_GTAB_CN1]; writes_hist = true; }   // HEUN_FIN
  if (xs == 0.0f) prep = false;                     // padding row: the run is over
  const float nz = r[DSK_GTAB_NZ];                  // STEP1 only: noise_strength(t) * sqrt|dt| (0 on deterministic steps)

  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (prep && tid < a.B) a.cnoise[tid] = cn;
  const bool strided = a.C > 1 || a.xin_ld != a.C;

  for (int64_t i = tid; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t cl = i, cli = i;                        // channels-last index into F rows [.., C] / network-input rows [.., xin_ld]
    if (strided) {
      const int64_t b = i / CS, rem = i - b * CS;
      const int64_t c = rem / a.S, s = rem - c * a.S;
      cl = (b * a.S + s) * a.C + c;
      cli = (b * a.S + s) * a.xin_ld + c;
    }
    const float xv = a.x[i];
    float out, carry;                               // out: what the stage stores as state / history; carry: what the next evaluation sees
    if (STAGE == DSK_GSTAGE_INIT) {
      out = carry = xv * a.x_scale;
    } else if (STAGE == DSK_GSTAGE_STEP1) {         // Euler, Euler-Maruyama, and the single-evaluation last Heun step (t + dt == 0)
      const float rhs = P * xv + Q * to_f32<T>(Fp[cl]);
      out = xv + dt * rhs;
      if (nz != 0.0f) {
        float z;
        if (a.noise != nullptr) {
          z = a.noise[(int64_t)rowi * N + i];
        } else {
          float q[4];
          philox_normal4(seed, (uint32_t)rowi, (uint64_t)(i >> 2), q);
          z = q[i & 3];
        }
        out += nz * z;
      }
      carry = out;
    } else if (STAGE == DSK_GSTAGE_HEUN_MID) {
      const float rhs = P * xv + Q * to_f32<T>(Fp[cl]);
      a.r1[i] = rhs;
      out = carry = xv + dt * rhs;
      a.x_aux[i] = out;
    } else {                                        // HEUN_FIN: r2 at (x_aux, t + dt)
      const float r2 = P * a.x_aux[i] + Q * to_f32<T>(Fp[cl]);
      out = carry = xv + (0.5f * (a.r1[i] + r2)) * dt;
    }
    if (STAGE != DSK_GSTAGE_HEUN_MID) {
      a.x[i] = out;
      if (writes_hist && a.hist != nullptr) a.hist[(int64_t)hist_slot * N + i] = out;
    }
    if (prep) xin[cli] = from_f32<T>(xs * carry);
  }
}

template <typename T>
static int launch_general(int stage, const GenArgs& a, cudaStream_t st) {
  const int grid = grid_for((int64_t)a.B * a.C * a.S, 256, 16);
  switch (stage) {
    case DSK_GSTAGE_INIT: DSK_LAUNCH((general_stage_kernel<T, DSK_GSTAGE_INIT>), grid, 256, 0, st, a); break;
    case DSK_GSTAGE_STEP1: DSK_LAUNCH((general_stage_kernel<T, DSK_GSTAGE_STEP1>), grid, 256, 0, st, a); break;
    case DSK_GSTAGE_HEUN_MID: DSK_LAUNCH((general_stage_kernel<T, DSK_GSTAGE_HEUN_MID>), grid, 256, 0, st, a); break;
    case DSK_GSTAGE_HEUN_FIN: DSK_LAUNCH((general_stage_kernel<T, DSK_GSTAGE_HEUN_FIN>), grid, 256, 0, st, a); break;
    default: set_error("dsk_sampler_stage_general: unknown stage %d", stage); return DSK_ERR_ARG;
  }
  return DSK_OK;
}

}  // namespace dsk

using namespace dsk;

extern "C" int dsk_sampler_stage_general(int stage, float* x, float* x_aux, float* r1, const void* F, void* xin, float* cnoise,
                                         const float* tab, const int* row, const float* noise, float* hist, int B, int C,
                                         int64_t S, float x_scale, int act_dtype, int xin_ld, void* stream) {
  DSK_REQUIRE(x && xin && cnoise && tab && row, "dsk_sampler_stage_general: null pointer");
  DSK_REQUIRE(stage == DSK_GSTAGE_INIT || F != nullptr, "dsk_sampler_stage_general: F is required after INIT");
  DSK_REQUIRE((stage != DSK_GSTAGE_HEUN_MID && stage != DSK_GSTAGE_HEUN_FIN) || (x_aux && r1),
              "dsk_sampler_stage_general: the Heun stages need x_aux and r1");
  DSK_REQUIRE(B > 0 && C > 0 && S > 0 && xin_ld >= C, "dsk_sampler_stage_general: bad shape");
  GenArgs a{x, x_aux, r1, F, xin, cnoise, tab, row, noise, hist, B, C, S, x_scale, xin_ld};
  if (act_dtype == DSK_F32) return launch_general<float>(stage, a, as_stream(stream));
  if (act_dtype == DSK_BF16) return launch_general<__nv_bfloat16>(stage, a, as_stream(stream));
  if (act_dtype == DSK_F16) return launch_general<__half>(stage, a, as_stream(stream));
  DSK_REQUIRE(false, "dsk_sampler_stage_general: bad act_dtype %d", act_dtype);
  return DSK_ERR_ARG;
}
