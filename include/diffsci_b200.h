/*
 * diffsci_b200.h -- C ABI of libdiffsci_b200.so (hand-written sm_100a CUDA for the
 * Karras/EDM hot path of Lacadame/DiffSci).
 *
 * The reference is pure Python/PyTorch and defines no FFI of its own; every
 * entry point below names the reference call site (file:line under /root/reference) whose
 * arithmetic it replaces.  The reference-side binding is a ctypes stub.
 *
 * Conventions
 *  - every function returns 0 on success, a negative code on error; the message is in
 *    dsk_last_error() (thread-local).  There is NO CPU fallback: a non-sm_100 device, a
 *    null pointer or an unsupported shape is an error, never a silent slow path.
 *  - all pointers are DEVICE pointers unless named host_*; the caller owns every buffer
 *    (inputs, outputs, workspaces).  No hidden allocations, no internal threads.
 *  - every launch is asynchronous on the cudaStream_t passed as `stream` (void* here so
 *    that the header needs no CUDA include); functions are CUDA-graph capturable.
 *  - activations inside the networks are CHANNELS-LAST [B, D, H, W, C] (D = 1 for 2-D),
 *    dtype DSK_F32, DSK_BF16 or DSK_F16; user-facing tensors x / D(x) / noise are fp32 NC(D)HW as in
 *    the reference.
 */
#ifndef DIFFSCI_B200_H
#define DIFFSCI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSK_OK 0
#define DSK_ERR_ARG -1
#define DSK_ERR_CUDA -2
#define DSK_ERR_UNSUPPORTED -3

/* Storage / operand formats.  DSK_BF16 and DSK_F16 are the two 16-bit formats of the tensor-core path (same kernels, the
 * format is a run-time flag: tcgen05 kind::f16 takes either).  fp16 keeps 3 more mantissa bits than bf16 (a forward pass of
 * the score networks lands at ~2e-3 of the reference's fp32 result instead of ~1.5e-2, oracle/split_budget.py) and is the
 * inference format; bf16 (fp32's exponent range) is the training format, where gradients need the range. */
enum { DSK_F32 = 0, DSK_BF16 = 1, DSK_F16 = 2,
       /* split-fp16: a tensor of fp32-class values v stored as [.., 2C] fp16 -- channels [0, C) hold hi = fp16(v), channels
        * [C, 2C) hold lo = fp16(2^11 (v - hi)) (the scaling keeps lo a normal fp16 number); hi + 2^-11 lo carries 22 mantissa
        * bits.  The operand format of the tensor-core fp32-parity mode: a split convolution / GEMM runs 3 tcgen05 MMAs per
        * k-step -- hi*hi into one group of fp32 TMEM accumulators, hi*lo + lo*hi into another -- and combines them in the
        * epilogue (products below 2^-22 dropped).  tcgen05 accumulates with truncation (tools/probe_tc_accum.py), so the hi*hi
        * products of a convolution tile are spread over three accumulator sets added in fp32 afterwards.  Reproduces the
        * reference's fp32 arithmetic to 1e-5-class over a whole network at a third of the 16-bit tensor-core rate instead of
        * CUDA-core FFMA rate. */
       DSK_SPLIT_F16 = 3 };

/* ---- library ---------------------------------------------------------------------- */
int dsk_version(void);
const char* dsk_last_error(void);
/* number of kernels this library has launched (or recorded into a graph) in this process */
uint64_t dsk_launch_count(void);
/* fails unless `device` is compute capability 10.x (B200) */
int dsk_check_device(int device);

/* ---- K7: EDM preconditioning + sampler stages ---------------------------------------
 * Replaces KarrasModule.get_denoiser / get_score (karras/karrasmodule.py:673-733),
 * EDMPreconditioner (karras/preconditioners.py:30-53), Scheduler.rhs EDM branch
 * (karras/schedulers.py:247-274) and the integrator steps (karras/integrators.py:29-113).
 *
 * x, x_aux, r1, noise, hist_row: fp32 [B, C, S] (reference NC(D)HW layout, S = prod(spatial)).
 * F / xin: network output / input, channels-last [B, S, C], dtype `act_dtype`.
 * cnoise: fp32 [B] = c_noise(sigma), the network's second argument.
 */

/* xin[b,s,c] = c_in[b] * x[b,c,s] for ANY preconditioner: the [B] coefficient vectors come from the
 * host-side preconditioner object (preconditioners.py:30-161). */
int dsk_precond_scale(const float* x, const float* c_in, void* xin, int B, int C, int64_t S, int act_dtype,
                      void* stream);
/* D = c_out[b]*F + c_skip[b]*x (D may be NULL); score = (D - x)/sigma[b]^2 (score may be NULL). */
int dsk_precond_denoise(const void* F, const float* x, const float* c_out, const float* c_skip, const float* sigma,
                        float* D, float* score, int B, int C, int64_t S, int act_dtype, void* stream);

/* Step table: fp32 [nrows][DSK_TAB_COLS], one row per integrator step, built on the host
 * (diffsci_b200.models.karras.schedulers) exactly as EDMScheduler.create_steps does
 * (schedulers.py:377-385).  `row` is a DEVICE int32[4]: row[0] = current step (so that one captured
 * CUDA graph serves every step; dsk_sampler_advance increments it), row[1], row[2] = low / high word of
 * a run-time Philox seed that is XOR-ed into the by-value `seed` (replayable graphs, new noise). */
#define DSK_TAB_COLS 8
enum {
  DSK_TAB_T = 0,      /* t_i                                             */
  DSK_TAB_DT = 1,     /* dt_i = t_{i+1} - t_i                            */
  DSK_TAB_THAT = 2,   /* Karras: t_i + gamma_i*t_i ; otherwise t_i       */
  DSK_TAB_LANG = 3,   /* EM: langevin_factor(t_i) (0 outside interval)   */
  DSK_TAB_NOISE = 4,  /* EM: noise_injection(t_i) = sqrt(2*lang)         */
  DSK_TAB_SQDT = 5,   /* EM: sqrt(|dt_i|)                                */
  DSK_TAB_TNEXT = 6,  /* sigma of the first network evaluation of step i+1 (t_{i+1}, or its t_hat) */
  DSK_TAB_CHURN = 7   /* Karras: sqrt(that^2-t^2)*S_noise of THIS step (0 otherwise).  The table has
                         nsteps+1 rows (last row zero) because KARRAS_FIN reads the next row's churn. */
};
enum {
  DSK_STAGE_INIT = 0,       /* x *= sigma_max ; (Karras churn) ; prep first evaluation       */
  DSK_STAGE_EULER = 1,      /* x += dt*rhs ; prep next step                                  */
  DSK_STAGE_HEUN_MID = 2,   /* r1 = rhs ; x_aux = x + dt*r1 ; prep evaluation at t+dt        */
  DSK_STAGE_HEUN_FIN = 3,   /* r2 = rhs(x_aux) ; x += 0.5*(r1+r2)*dt ; prep next step        */
  DSK_STAGE_HEUN_LAST = 4,  /* t+dt == 0: x += 0.5*(r1+r1)*dt (integrators.py:49-53)         */
  DSK_STAGE_EM = 5,         /* x += rhs_stoch*dt + noise_strength*xi*sqrt|dt| ; prep next     */
  DSK_STAGE_KARRAS_MID = 6, /* r1 at (x_hat,t_hat); x_aux = x_hat + dt_hat*r1 ; prep at t+dt  */
  DSK_STAGE_KARRAS_FIN = 7, /* x = x_hat + 0.5*(r1+r2)*dt_hat ; churn + prep next step       */
  DSK_STAGE_KARRAS_LAST = 8 /* t+dt == 0: x = x_aux (Euler from t_hat, integrators.py:108-112) */
};
/* One fused elementwise pass of an integrator stage.  `noise` is an injected N(0,1) tensor
 * array (fp32 [nrows][B,C,S], indexed by the step that consumes it) or NULL, in which case a
 * counter-based Philox4x32-10 stream keyed by (seed, row, element) supplies it.  hist (may be NULL)
 * is fp32 [nrows+1][B,C,S]; slot row+1 receives the new x when the stage completes a step
 * (history[i+1], schedulers.py:86-87), slot 0 the scaled initial noise.
 * precond_kind: 0 = EDMPreconditioner(sigma_data), 1 = NullPreconditioner. */
int dsk_sampler_stage(int stage, float* x, float* x_aux, float* r1, const void* F, void* xin, float* cnoise,
                      const float* tab, const int* row, const float* noise, uint64_t seed, float* hist,
                      int B, int C, int64_t S, float sigma_data, float sigma_max, int precond_kind, int act_dtype,
                      void* stream);
int dsk_sampler_advance(int* row, void* stream);

/* Table-driven stages for ANY scheduler / preconditioner pair (SURVEY 8f-3 on the graph engine; the non-constant-scaling branch
 * of Scheduler.rhs, karras/schedulers.py:275-293, with any KarrasPreconditioner, karras/preconditioners.py:30-161).  For fixed t
 * the drift is linear in the state and the network output, rhs = P x + Q F with
 *   P = s'/s + Bm (c_skip - 1) / (sigma^2 s),  Q = Bm c_out / sigma^2,  network input = (c_in / s) x,
 * Bm = -(s sigma' sigma | pf_score_multiplier) [- langevin_factor / s on stochastic steps]; the host evaluates these with the
 * scheduler's and the preconditioner's own objects (diffsci_b200.models.karras.schedulers.Scheduler.general_step_table).
 * Table: fp32 [nsteps + 1][DSK_GTAB_COLS], last row zero padding. */
#define DSK_GTAB_COLS 12
enum dsk_gtab_col {
  DSK_GTAB_DT = 0,   /* dt_i                                                                   */
  DSK_GTAB_P1 = 1,   /* rhs coefficients of the evaluation at t_i                              */
  DSK_GTAB_Q1 = 2,
  DSK_GTAB_P2 = 3,   /* ... at t_i + dt_i (Heun's second evaluation)                           */
  DSK_GTAB_Q2 = 4,
  DSK_GTAB_XS1 = 5,  /* network-input scale c_in/s and c_noise of the evaluation at t_i        */
  DSK_GTAB_CN1 = 6,
  DSK_GTAB_XS2 = 7,  /* ... at t_i + dt_i                                                      */
  DSK_GTAB_CN2 = 8,
  DSK_GTAB_HAS2 = 9, /* 1: the step has a second evaluation (t_i + dt_i > 0, integrators.py:45) */
  DSK_GTAB_NZ = 10   /* Euler-Maruyama: noise_strength(t_i) * sqrt|dt_i| (integrators.py:66-69) */
};
enum dsk_gstage {
  DSK_GSTAGE_INIT = 0,      /* x *= x_scale ; history[0] ; prepare the evaluation at t_0                          */
  DSK_GSTAGE_STEP1 = 1,     /* x += dt (P1 x + Q1 F) (+ NZ xi) ; history ; prepare the next step                  */
  DSK_GSTAGE_HEUN_MID = 2,  /* r1 = P1 x + Q1 F ; x_aux = x + dt r1 ; prepare the evaluation at t + dt            */
  DSK_GSTAGE_HEUN_FIN = 3   /* r2 = P2 x_aux + Q2 F ; x += 0.5 (r1 + r2) dt ; history ; prepare the next step     */
};
int dsk_sampler_stage_general(int stage, float* x, float* x_aux, float* r1, const void* F, void* xin, float* cnoise,
                              const float* tab, const int* row, const float* noise, float* hist, int B, int C, int64_t S,
                              float x_scale, int act_dtype, int xin_ld, void* stream);

/* Conditional sampling (KarrasModule.get_denoiser with y / guidance, karrasmodule.py:703-716;
 * PUNetGCond's channel concatenation, nets/punetg.py:716-735).  Same stage, with
 *   xin_ld >= C : the network-input rows hold xin_ld channels, the state x fills channels [0, C); channels [C, xin_ld)
 *                 carry the channel-concatenated conditioning, written once per run by the caller;
 *   cfg != 0    : classifier-free guidance evaluated as ONE 2B-sample network call -- xin/cnoise rows [B, 2B) receive
 *                 copies of rows [0, B), and F is read as (1-guidance)*F[0:B] + guidance*F[B:2B] (rows [0,B) = the
 *                 unconditional evaluation, rows [B,2B) the conditional one). */
int dsk_sampler_stage_cond(int stage, float* x, float* x_aux, float* r1, const void* F, void* xin, float* cnoise,
                           const float* tab, const int* row, const float* noise, uint64_t seed, float* hist,
                           int B, int C, int64_t S, float sigma_data, float sigma_max, int precond_kind, int act_dtype,
                           int xin_ld, int cfg, float guidance, void* stream);
/* The same stage with the known-region blend of Scheduler.inpaint (karras/schedulers.py:91-122) fused in: after every completed
 * step (and after DSK_STAGE_INIT) x <- x (1 - mask) + blend_y[level] mask, level = blend_rows - 1 - (steps completed), before the
 * history write and before the next network input is prepared.  blend_y: fp32 [blend_rows][B, C, S], the forward (data -> noise)
 * history of the known data (blend_rows = nsteps + 1); blend_mask: fp32, mask_n elements broadcast over the leading dimensions.
 * blend_y == NULL: dsk_sampler_stage_cond.  DSK_STAGE_INIT at row r > 0 starts a run at step r of the table
 * (Scheduler.propagate_partial, schedulers.py:177-217) and writes history slot r. */
int dsk_sampler_stage_blend(int stage, float* x, float* x_aux, float* r1, const void* F, void* xin, float* cnoise,
                            const float* tab, const int* row, const float* noise, uint64_t seed, float* hist,
                            int B, int C, int64_t S, float sigma_data, float sigma_max, int precond_kind, int act_dtype,
                            int xin_ld, int cfg, float guidance, const float* blend_y, const float* blend_mask, int64_t mask_n,
                            int blend_rows, void* stream);
/* dsk_precond_scale into rows of xin_ld channels; dup != 0 also fills rows [B, 2B) (the CFG batch). */
int dsk_precond_scale_cond(const float* x, const float* c_in, void* xin, int B, int C, int64_t S, int act_dtype,
                           int xin_ld, int dup, void* stream);
/* F_uncond <- (1-guidance)*F_uncond + guidance*F_cond over n elements of `act_dtype` (karrasmodule.py:711-713). */
int dsk_cfg_mix(void* F_uncond, const void* F_cond, float guidance, int64_t n, int act_dtype, void* stream);

/* Integrator.step seam for foreign score functions (integrators.py:29-113; schedulers.py:266-274):
 * out = a0*x + a1*r1 + a2*r2 + a3*z, fp32, null inputs skipped; out may alias any input. */
int dsk_lincomb(float* out, int64_t n, const float* x, float a0, const float* r1, float a1, const float* r2, float a2,
                const float* z, float a3, void* stream);
/* N(0,1) draws from the same Philox4x32-10 stream the fused stages use: element i of stream
 * `stream_id` (= step row) under `seed`.  Replaces torch.randn_like (integrators.py:68,105). */
int dsk_philox_normal(float* out, int64_t n, uint64_t seed, uint32_t stream_id, void* stream);
/* Training-mode dropout, forward AND backward: y = x * keep / (1-p) (+ dres), keep(i) = [u_i >= p] from Philox stream
 * (seed, stream_id): the backward launch of a site passes the same (seed, stream_id) and regenerates the mask.
 * Replaces torch.nn.Dropout in ResnetBlockC / ADMBaseBlock (commonlayers.py:792,830; adm.py:312-313).  dtype: DSK_F32 | DSK_BF16;
 * y may alias x or dres. */
int dsk_dropout(const void* x, const void* dres, void* y, int64_t n, float p, uint64_t seed, uint32_t stream_id, int dtype,
                void* stream);

/* ---- K1: convolution / GEMM ----------------------------------------------------------
 * Replaces torch.nn.Conv2d/Conv3d(padding='same') in ResnetBlockC (nets/commonlayers.py:777-833),
 * DownSampler/UpSampler (commonlayers.py:53-58,123-128), convin/convout (nets/punetg.py:203-214),
 * ADM convs (nets/adm.py:268-284) and torch.nn.Linear / in/out projections of MultiheadAttention.
 *
 * in  : [B, Di, Hi, Wi, Cin]  (if up2: the conv sees nearest-upsampled x2 input, i.e. F.interpolate
 *        fused into the gather: commonlayers.py:129,145)
 * w   : packed weights. DSK_F32 path: fp32 [taps][Cin][Cout];  16-bit paths (DSK_BF16 | DSK_F16): [taps][Cout][Cin];
 *       DSK_SPLIT_F16: fp16 [taps][Cout][2 Cin] (hi | lo)   (16-bit + up2: the sub-pixel layout of dsk_pack_upconv_weight)
 * out : [B, D, H, W, Cout] ; y = conv(in) + bias[co] + chan_bias[b,co] + residual[b,..,co]
 * ksize in {1,3}; ndim in {2,3} (ndim 2 => D = 1, taps = k*k).
 */
typedef struct {
  int B, D, H, W;      /* OUTPUT spatial size (input is half of it when up2)            */
  int Cin, Cout;
  int ksize, ndim;
  int up2;             /* 1: nearest x2 upsample fused into the input gather             */
  int w_dtype;         /* DSK_F32: w is fp32 [taps][Cin][Cout] -> CUDA-core FFMA kernel;
                          DSK_BF16 | DSK_F16: 16-bit [taps][Cout][Cin] -> tcgen05 implicit-GEMM kernel (in_dtype = the same format);
                          DSK_SPLIT_F16: [taps][Cout][2 Cin] -> the same kernel with split operands: in_dtype = DSK_SPLIT_F16
                          ([B,D,H,W,2 Cin]), 3 MMAs per k-step (fp32-parity mode on the tensor cores).  in_dtype = DSK_SPLIT_F16
                          with w_dtype = DSK_F16: activations split only, 2 MMAs per k-step.                                   */
  int in_dtype;        /* dtype of `in`                                                   */
  int out_dtype;       /* dtype of `out` and `residual`: the 16-bit format of the operands, or DSK_F32 */
  int out_nchw_f32;    /* 1: write fp32 NC(D)HW (user layout) instead of channels-last    */
  int circular;        /* 1: circular ('periodic') padding on every spatial axis instead of zeros:
                          CircularConv2d / CircularConv3d (commonlayers.py:918-1032), PUNetGConfig(convolution_type=
                          "circular").  CUDA-core kernels wrap their gather; tcgen05 kernels read a halo-padded copy
                          (dsk_conv_fwd_circ / the workspace of dsk_conv_wgrad).                                       */
  int res_dtype;       /* DSK_RES_SAME (0): `residual` has out_dtype;  DSK_RES_F32: `residual` is fp32 whatever out_dtype is --
                          tcgen05 kernels only: a block of an fp32-storage mode whose output is read by ONE convolution writes
                          the 16-bit operand copy directly (fp32 residual stream in, fp16 out) instead of fp32 + a cast pass */
  int operand16;       /* 0, or -- with in_dtype = out_dtype = w_dtype = DSK_F32 on a few-input-channel convolution (convin, Cin <= 4:
                          the first layer of the fp32-storage modes) -- run it on the tensor cores instead of the CUDA-core fp32
                          kernel: DSK_BF16 | DSK_F16: operands rounded to this format; DSK_SPLIT_F16 (taps * Cin <= 32): fp16
                          operands split hi + lo inside one im2col row, x_hi w_hi + x_lo w_hi + x_hi w_lo: fp32-class result */
} dsk_conv_desc;
#define DSK_RES_SAME 0
#define DSK_RES_F32 1
int dsk_conv_fwd(const dsk_conv_desc* d, const void* in, const void* w, const float* bias,
                 const float* chan_bias, const void* residual, void* out, void* stream);
/* The same convolution with the statistics of the FOLLOWING per-channel norm fused into its epilogue (the reference
 * runs GroupNorm(C, C) / GroupRMSNorm(C, C) as separate passes over the conv output, commonlayers.py:824-831):
 * stats[b][slot][c] (float2: sum, sum of squares over the pixels one epilogue warp stored; dsk_conv_stats_slots() slots
 * per sample, fixed summation order) is consumed by dsk_norm_act_prestat.  Only the cta_group::2 tcgen05 kernel emits
 * them: query dsk_conv_stats_supported(d) first. */
int dsk_conv_stats_supported(const dsk_conv_desc* d);
int dsk_conv_stats_slots(void);
int dsk_conv_fwd_stats(const dsk_conv_desc* d, const void* in, const void* w, const float* bias, const float* chan_bias,
                       const void* residual, void* out, void* stats, void* stream);
/* Circular padding.  dsk_pad_circular: y[B, D+2, H+2, W+2, C] (ndim 2: [B, 1, H+2, W+2, C]) = x wrapped by one
 * pixel on every spatial axis (torch.nn.functional.pad(mode='circular'), commonlayers.py:960-967, 1015-1030).
 * dsk_conv_fwd_circ: dsk_conv_fwd / dsk_conv_fwd_stats (stats may be NULL) for d->circular = 1; `pad_ws` holds the padded
 * copy the tcgen05 kernels read through TMA (dsk_conv_pad_ws_bytes(d) bytes; 0 = this convolution runs on a CUDA-core
 * kernel that wraps its indices, pad_ws may be NULL). */
int dsk_pad_circular(const void* x, void* y, int B, int D, int H, int W, int C, int ndim, int dtype, void* stream);
int64_t dsk_conv_pad_ws_bytes(const dsk_conv_desc* d);
int dsk_conv_fwd_circ(const dsk_conv_desc* d, const void* in, const void* w, const float* bias, const float* chan_bias,
                      const void* residual, void* out, void* stats, void* pad_ws, void* stream);
/* d->circular = 2: `in` is ALREADY the halo-padded tensor [B, D+2, H+2, W+2, C] (written by dsk_norm_apply_padded): tcgen05
 * kernels only, no padding pass.  dsk_norm_apply_padded: the apply pass of dsk_norm_act / dsk_norm_act_prestat (called with
 * y = NULL: statistics + folded scale/shift table in `ws` only) writing that padded layout directly -- interior plus wrapped
 * halos -- so the GroupNorm + SiLU in front of a circular convolution (commonlayers.py:824-831, 918-1032) costs no extra pass. */
int dsk_norm_apply_padded(const void* x, void* y_padded, const void* ws, int B, int D, int H, int W, int C, int ndim, int silu,
                          int in_dtype, int out_dtype, void* stream);
/* nearest x2 upsample of a channels-last tensor (torch.nn.Upsample(scale_factor=2), commonlayers.py:129):
 * only needed in front of the tcgen05 conv; the FFMA conv fuses it into its gather (up2). bf16, C % 8 == 0. */
int dsk_upsample2x(const void* x, void* y, int B, int D, int H, int W, int C, int ndim, int dtype, void* stream);
/* Sub-pixel weights for the tcgen05 UpSampler conv (up2 = 1 with 16-bit weights): [2^ndim phases][2^ndim taps][Cout][Cin] in
 * `dtype` (DSK_BF16 | DSK_F16 | DSK_SPLIT_F16: rows of [hi | lo], 2 Cin long), the three taps of each axis pre-summed in fp32
 * onto the two input offsets an output parity sees (csrc/conv_tc.cu). */
int dsk_pack_upconv_weight(const float* w_ref, void* w_packed, int Cout, int Cin, int ndim, int dtype, void* stream);
/* repack a reference-layout weight [Cout, Cin, k(,k)(,k)] fp32 into the two packed layouts */
int dsk_pack_conv_weight(const float* w_ref, void* w_packed, int Cout, int Cin, int taps, int dtype, void* stream);

/* dsk_pack_conv_weight / dsk_pack_conv_weight_dgrad for MANY 16-bit weights in ONE launch (the re-pack after every optimizer
 * step of a training iteration, karrasmodule.py:1146-1155 + :497-507).  jobs: DEVICE array of
 *   struct { const float* w_ref; void* w_packed; int Cout, Cin, taps, dgrad, dtype, block0; }
 * (taps <= 27; dtype DSK_BF16 | DSK_F16 | DSK_SPLIT_F16), block0 = first block of the job = sum over the previous jobs of
 * (dgrad ? Cin : Cout) * ceil((dgrad ? Cout : Cin) / 64); total_blocks = that sum over all jobs. */
int dsk_pack_conv_weights_multi(const void* jobs, int njobs, int total_blocks, void* stream);

/* Batched C[b] = act(alpha * A[b] (MxK, row-major, lda) * B[b] + bias[n]) (act: 0 none, 1 SiLU, 2 ReLU);
 * transB = 1: B is [N][K] row-major (C = A*B^T) ; transB = 0: B is [K][N].  fp32 CUDA-core path. */
int dsk_gemm_f32(const float* A, const float* Bm, float* Cm, const float* bias, int M, int N, int K, int lda,
                 int ldb, int ldc, int64_t strideA, int64_t strideB, int64_t strideC, int batch, int transB,
                 float alpha, int act, void* stream);

/* General form: C[b] = act(alpha * op(A[b]) op(B[b]) + bias[n]) + beta * C[b];  transA = 1: A is stored [K][M] (lda >= M).
 * The A^T B and A B products of the attention / linear backward passes. */
int dsk_gemm_f32_ex(const float* A, const float* Bm, float* Cm, const float* bias, int M, int N, int K, int lda, int ldb,
                    int ldc, int64_t strideA, int64_t strideB, int64_t strideC, int batch, int transA, int transB,
                    float alpha, float beta, int act, void* stream);

/* Batched bf16 GEMM on tcgen05/TMEM (csrc/gemm_tc.cu): C[b] = alpha * A[b] B[b]^T + bias (+ residual).
 * A [M,K] (lda), B [N,K] (ldb): bf16, K contiguous; C bf16 or fp32 (out_f32) with leading dimension ldc;
 * bias fp32 per column (bias_rows = 0) or per row (bias_rows = 1); residual bf16 laid out like C.
 * A batch stride of 0 shares the operand across the batch.  K, ld*, strides: multiples of 8 elements.
 * transA / transB: the operand is stored [K, M] / [K, N] (lda >= M / ldb >= N) and is fed to the tensor core as an
 * MN-major operand -- the A^T B products of autograd's backward (weight gradients, dV = P^T dO, dK = dS^T Q).
 * The tensor-core form of the projections / QK^T / PV inside nn.MultiheadAttention (nets/attention.py:42-44). */
int dsk_gemm_bf16_tc(const void* A, const void* Bm, void* C, const float* bias, int bias_rows, const void* residual, int M,
                     int N, int K, int64_t lda, int64_t ldb, int64_t ldc, int64_t strideA, int64_t strideB,
                     int64_t strideC, int batch, float alpha, int out_f32, int transA, int transB, void* stream);
/* The same kernel for every tensor-core operand format.  dtype DSK_BF16 | DSK_F16: as above in that 16-bit format (residual
 * and 16-bit C in the same format; res_f32 = 1: the residual is fp32).  dtype DSK_SPLIT_F16: A and B are split-fp16 operands
 * (hi | lo along their CONTIGUOUS coordinate -- k for K-major, m / n for MN-major operands -- with the lo half a_lo / b_lo
 * elements after the hi half; lda / ldb are the full row lengths): C accumulates hi*hi + hi*lo + lo*hi, the fp32-parity form
 * of the projections / QK^T / PV of nn.MultiheadAttention (attention.py:42-44).  b_lo < 0: B is plain fp16 (2 MMAs per
 * k-step).  Split operands need K % 64 == 0. */
int dsk_gemm_tc(const void* A, const void* Bm, void* C, const float* bias, int bias_rows, const void* residual, int res_f32, int M,
                int N, int K, int64_t lda, int64_t ldb, int64_t ldc, int64_t strideA, int64_t strideB, int64_t strideC, int batch,
                float alpha, int out_f32, int transA, int transB, int dtype, int a_lo, int b_lo, void* stream);
/* softmax backward on rows: dS = P * (dP - rowsum(dP * P)); P bf16, dP fp32, dS bf16 (cols % 4 == 0, cols <= 8192) */
int dsk_softmax_bwd_rows_bf16(const void* P, const float* dP, void* dS, int64_t rows, int cols, void* stream);
/* P[b] = softmax(alpha Q[b] K[b]^T) written once as bf16 [batch, L, L] -- the middle of nn.MultiheadAttention
 * (nets/attention.py:42-44,68) without the fp32 score tensor: QK^T runs twice on the tensor cores, first with a
 * row-statistics epilogue, then with the exp / normalise epilogue.  Q, K: bf16 [L, C] per batch (K-major, ldq / ldk, batch
 * strides); L % 8 == 0; ws: dsk_attn_softmax_ws_bytes(batch, L) bytes. */
int64_t dsk_attn_softmax_ws_bytes(int batch, int L);
int dsk_attn_softmax_qk(const void* Q, const void* K, void* P, void* ws, int L, int C, int64_t ldq, int64_t ldk,
                        int64_t strideQ, int64_t strideK, int batch, float alpha, void* stream);
/* ... with Q, K, P in `dtype` (DSK_BF16 | DSK_F16) */
int dsk_attn_softmax_qk_h16(const void* Q, const void* K, void* P, void* ws, int L, int C, int64_t ldq, int64_t ldk,
                            int64_t strideQ, int64_t strideK, int batch, float alpha, int dtype, void* stream);
/* O[b] = softmax(alpha Q[b] K[b]^T) V[b] as ONE flash-style kernel (replaces the SDPA call inside
 * torch.nn.MultiheadAttention(C, 1 head), nets/attention.py:42-44, 93-102): scores, probabilities and the output accumulator
 * live in TMEM (S = Q K^T and O += P V on tcgen05, P as the TMEM A operand), online softmax with a lazily updated row
 * maximum -- nothing of size L x L reaches shared memory or HBM.  Q, K, V: `dtype` (DSK_BF16 | DSK_F16) [L, C] per batch with
 * leading dimensions ldq / ldk / ldv (e.g. the three column blocks of a packed projection [L, 3C]); C = 128 | 256; any L.
 * out_mode 0: O in `dtype` [L, ldo]; 1: fp32; 2 (DSK_F16 only): split fp16 rows [hi (C) | 2^11 lo (C)] for a split GEMM. */
int dsk_attn_flash(const void* Q, const void* K, const void* V, void* O, int L, int C, int64_t ldq, int64_t ldk, int64_t ldv,
                   int64_t ldo, int64_t strideQ, int64_t strideK, int64_t strideV, int64_t strideO, int batch, float alpha,
                   int dtype, int out_mode, void* stream);
/* softmax over the last dim: fp32 scores [rows, cols] -> bf16 probabilities (cols % 4 == 0, cols <= 8192) */
int dsk_softmax_rows_bf16(const float* S, void* P, int64_t rows, int cols, void* stream);
/* ... -> P in `dtype`: DSK_BF16 | DSK_F16 (rows of cols), DSK_SPLIT_F16 (rows of 2 cols: hi | lo, the A operand of a split PV) */
int dsk_softmax_rows_h16(const float* S, void* P, int64_t rows, int cols, int dtype, void* stream);

/* ---- K4: per-group norm + affine (+FiLM) + SiLU ---------------------------------------
 * Replaces torch.nn.GroupNorm(G,C) / GroupRMSNorm(G,C) (+ SiLU) in ResnetBlockC
 * (commonlayers.py:362-384, 824-831) and ADMBaseBlock (nets/adm.py:305-329).
 * mode 0 = LayerNorm-style (mean/var, biased), 1 = RMS.  eps = 1e-5.
 * x,y: [B, S, C]; G groups of C/G channels (G == C for PUNetG, G == 1 default for ADM).
 * film_scale/film_shift: optional fp32 [B, C] (y = norm*film_scale + film_shift, adm.py:305-308).
 * ws: workspace of dsk_norm_ws_bytes(B, S, C) bytes.
 */
int64_t dsk_norm_ws_bytes(int B, int64_t S, int C);
int dsk_norm_act(const void* x, void* y, const float* gamma, const float* beta, const float* film_scale,
                 const float* film_shift, void* ws, int B, int64_t S, int C, int G, int mode, int silu,
                 int in_dtype, int out_dtype, void* stream);
/* dsk_norm_act without the statistics pass: mean / rstd come from the `conv_stats` a preceding dsk_conv_fwd_stats left
 * (G == C only).  One finalize launch + the apply pass: 2 N s bytes instead of 3 N s. */
int dsk_norm_act_prestat(const void* x, void* y, const float* gamma, const float* beta, const float* film_scale,
                         const float* film_shift, const void* conv_stats, int nslots, void* ws, int B, int64_t S, int C,
                         int G, int mode, int silu, int in_dtype, int out_dtype, void* stream);

/* ---- K5: 2x pooling -------------------------------------------------------------------
 * Replaces torch.nn.MaxPool{2,3}d(2) (commonlayers.py:60-63,81) / AvgPool (adm.py:361-371).
 * x: [B, D, H, W, C] -> y: [B, D/2 (or 1), H/2, W/2, C] (floor, like torch). */
int dsk_pool2x(const void* x, void* y, int B, int D, int H, int W, int C, int ndim, int is_max, int dtype,
               void* stream);
/* ... fp32 input (C % 4 == 0, float4 per thread), output fp32 or rounded once to out_dtype = DSK_BF16 | DSK_F16: the fp32-storage
 * precision modes pool straight into the 16-bit operand copy of the DownSampler convolution (commonlayers.py:53-63) */
int dsk_pool2x_f32(const float* x, void* y, int B, int D, int H, int W, int C, int ndim, int is_max, int out_dtype, void* stream);
/* y = a + b (elementwise, same dtype); used for the additive U-Net skips (punetg.py:373,384) */
int dsk_add(const void* a, const void* b, void* y, int64_t n, int dtype, void* stream);
/* NCHW fp32 <-> channels-last conversions at the module boundary */
int dsk_nchw_to_cl(const float* x, void* y, int B, int C, int64_t S, int dtype, void* stream);
int dsk_cl_to_nchw(const void* x, float* y, int B, int C, int64_t S, int dtype, void* stream);
int dsk_cast(const void* x, void* y, int64_t n, int in_dtype, int out_dtype, void* stream);
/* fp32 [rows, C] -> split-fp16 [rows, 2C] (DSK_SPLIT_F16: hi | lo): the operand form of the split tensor-core kernels for
 * tensors no fused producer writes in that form (pooled / raw residual-stream inputs of DownSampler, UpSampler and convout,
 * commonlayers.py:53-58,123-128; punetg.py:209-214; attention tokens, attention.py:42-44).  C % 4 == 0. */
int dsk_split_f16(const float* x, void* y, int64_t rows, int C, void* stream);
/* channel concat of two channels-last tensors (ADM 'concat' skips, adm.py:296-297) */
int dsk_concat_channels(const void* a, const void* b, void* y, int64_t rows, int Ca, int Cb, int dtype,
                        void* stream);

/* out = x * (1 - mask) + y * mask on fp32 tensors: the known-region blend of Scheduler.inpaint / Scheduler.repaint
 * (karras/schedulers.py:111-116, 151, 162).  mask has mask_n elements and is broadcast over the leading dimensions. */
int dsk_mask_blend(float* out, const float* x, const float* y, const float* mask, int64_t n, int64_t mask_n, void* stream);

/* ---- K6: time embedding ----------------------------------------------------------------
 * Fourier features (commonlayers.py:175-190): out[b] = cat(sin(2*pi*t_b*W), cos(2*pi*t_b*W)), fp32. */
int dsk_fourier(const float* t, const float* W, float* out, int B, int half, void* stream);
/* Grouped small-batch linear: for g in [0,ngroups): Y_g[B,out_g] = act(X_g[B,in_g] W_g^T + b_g).
 * Device arrays of pointers / sizes (one launch for the time MLP layer of every ResNet block;
 * commonlayers.py:516-550, adm.py:331-343, nets/mlp.py:38-58).  act: 0 none, 1 SiLU, 2 ReLU.
 * Z (optional table): receives the pre-activation X W^T + b, which the backward pass needs. */
int dsk_grouped_linear(const float* const* X, const float* const* W, const float* const* bias, float* const* Y,
                       float* const* Z, const int* in_dim, const int* out_dim, int ngroups, int max_out, int B,
                       int act, void* stream);
/* Backward of the same layer for every group in two launches (what autograd runs below loss.backward() for the
 * nn.Linear + SiLU pairs of ResnetTimeBlock / ADM time embedding):
 *   dZ_g = dY_g * act'(Z_g)  (dZ_g may alias dY_g when act == 0),  dW_g = dZ_g^T X_g,  db_g = colsum(dZ_g),
 *   dX_g = dZ_g W_g -- or, with shared_dx, ONE dX (table entry 0) = sum_g dZ_g W_g for groups that read the same X.
 * dX == NULL (or a NULL entry) skips the input gradient; accumulate_dx adds into dX. */
int dsk_grouped_linear_bwd(const float* const* dY, const float* const* Z, const float* const* X,
                           const float* const* W, float* const* dZ, float* const* dW, float* const* db,
                           float* const* dX, const int* in_dim, const int* out_dim, int ngroups, int max_out,
                           int max_in, int B, int act, int shared_dx, int accumulate_dx, void* stream);

/* The same layers at large batch (B > 32): a grouped smem-tiled fp32 GEMM over a device table of
 *   struct { const float* A; const float* B; float* C; const float* bias; float* Z; int M, N, K, lda, ldb, ldc; }
 * C_g = act(op(A_g) op(B_g) + bias_g), Z_g (optional) = the pre-activation; transA: A stored [K][M]; transB: B stored [N][K].
 * Forward X W^T, weight gradient dZ^T X and input gradient dZ W are the three transpose combinations. */
int dsk_grouped_gemm_f32(const void* table, int ngroups, int max_m, int max_n, int transA, int transB, int act, void* stream);
/* dZ_g = dY_g * act'(Z_g) (skipped for act == 0) and db_g = colsum(dZ_g) for every group in one launch. */
int dsk_grouped_dz_bias(const float* const* dY, const float* const* Z, float* const* dZ, float* const* db, const int* out_dim,
                        int ngroups, int max_out, int B, int act, void* stream);

/* ---- K3: attention ----------------------------------------------------------------------
 * softmax over the last dim of S[batch*rows, cols] in place (fp32), the middle of
 * nn.MultiheadAttention (nets/attention.py:42-44,68): softmax(QK^T/sqrt(d)). */
int dsk_softmax_rows(float* S, int64_t rows, int cols, void* stream);

/* ---- K8/K9: training ---------------------------------------------------------------------
 * EDM loss forward + dL/dF (karrasmodule.py:590-650; noisesamplers.py:30-33):
 *   D = c_out*F + c_skip*(x+sigma*n);  L = mean(lambda(sigma) * l(D,x) * (1-mask)),
 *   l = Huber(delta=1) (loss_kind 0) | MSE (loss_kind 1);  dF = dL/dF (same layout/dtype as F, fp32).
 * loss_out: fp32 [1], must be zeroed by the caller.  mask may be NULL. */
int dsk_edm_loss_fwd_bwd(const float* F, const float* x, const float* noise, const float* sigma,
                         const float* mask, float* loss_out, float* dF, int B, int C, int64_t S,
                         float sigma_data, int loss_kind, void* stream);
/* The same loss for ANY preconditioner / noise sampler (VP, VE, SR3, uniform; preconditioners.py:56-136,
 * noisesamplers.py:44-110): the per-sample coefficients c_out[b], c_skip[b] and the loss weight lambda[b] are evaluated by
 * the host-side objects and passed in as fp32 [B] vectors. */
int dsk_precond_loss_fwd_bwd(const float* F, const float* x, const float* noise, const float* sigma, const float* c_out,
                             const float* c_skip, const float* weight, const float* mask, float* loss_out, float* dF,
                             int B, int C, int64_t S, int loss_kind, void* stream);
/* ... with torch.nn.HuberLoss's delta (loss_metric = {"huber": {"delta": d}}, karrasmodule.py:558-562):
 * l = r^2 / 2 for |r| <= delta, delta (|r| - delta / 2) otherwise; dl/dr = clamp(r, -delta, delta).  Ignored for MSE. */
int dsk_precond_loss_fwd_bwd_huber(const float* F, const float* x, const float* noise, const float* sigma, const float* c_out,
                                   const float* c_skip, const float* weight, const float* mask, float* loss_out, float* dF,
                                   int B, int C, int64_t S, int loss_kind, float delta, void* stream);
int dsk_precond_loss_rows_huber(const float* F, const float* x, const float* noise, const float* sigma, const float* c_out,
                                const float* c_skip, const float* weight, const float* mask, float* loss_b, float* dF,
                                int B, int C, int64_t S, int loss_kind, float delta, void* stream);
/* The same loss reduced PER SAMPLE: loss_b[b] (fp32 [B], zeroed by the caller) = sum over the sample of weight[b] * l * (1-mask)
 * / (B*C*S), so that sum_b loss_b = the scalar above.  For a per-sample factor applied on the host side that needs its own
 * gradient: the learned uncertainty weighting of has_dynamic_loss_weight (karrasmodule.py:594-602, DynamicLossWeight
 * :1256-1278): loss = sum_b exp(-u_b) loss_b + mean(u). */
int dsk_precond_loss_rows(const float* F, const float* x, const float* noise, const float* sigma, const float* c_out,
                          const float* c_skip, const float* weight, const float* mask, float* loss_b, float* dF,
                          int B, int C, int64_t S, int loss_kind, void* stream);
/* Ensemble training loss (EnsembleKarrasModule.loss_fn, karras/karrasmodule_new.py:963-1149, with the
 * ensemble-aware metrics of custom_losses.py:536-690 and :765-865).  Layouts: x fp32 [B][C*S]; noise, F, dF, out fp32
 * [B][E][C*S] (the B*E rows the network sees).
 *   dsk_ensemble_noise_add:  out[b][e] = x[b] + sigma[b] * noise[b][e]                  (karrasmodule_new.py:1014-1040)
 *   dsk_ensemble_loss_fwd_bwd: D[b][e] = c_out[b] F[b][e] + c_skip[b] (x[b] + sigma[b] noise[b][e]);
 *     loss_kind 0 Huber(delta=1) | 1 MSE:  loss += s1[b] * l(D[b][e] - x[b]) * (1 - mask)
 *     loss_kind 2 CRPS:                    loss += s1[b] * sum_e |D[b][e] - x[b]| - s2[b] * sum_{i<j} |D[b][i] - D[b][j]|
 *     dF = c_out[b] * dloss/dD.  s1, s2: fp32 [B], every constant of the reference's reductions folded in by the caller
 *     (mean lambda, 1/(B E N), valid-pixel counts).  mask: NULL or fp32 [B][mask_C][S], mask_C in {1, C}; not for CRPS
 *     (the reference's CRPS mask only rescales per sample, custom_losses.py:851-857).  E in 1..16.
 *     loss_out: fp32 [1], zeroed by the caller. */
int dsk_ensemble_noise_add(const float* x, const float* noise, const float* sigma, float* out, int B, int E, int64_t CS,
                           void* stream);
int dsk_ensemble_loss_fwd_bwd(const float* F, const float* x, const float* noise, const float* sigma, const float* c_out,
                              const float* c_skip, const float* s1, const float* s2, const float* mask, int mask_C,
                              float* loss_out, float* dF, int B, int E, int C, int64_t S, int loss_kind, void* stream);
/* Multi-tensor EMA: shadow_i <- lerp(shadow_i, p_i, 1-beta) (karras/ema.py:139-147) in one launch. */
int dsk_ema_update(float* const* shadow, const float* const* param, const int64_t* numel, int ntensors,
                   int64_t max_numel, float beta, void* stream);
/* Multi-tensor AdamW (reference default optimizer, karrasmodule.py:497-500) fused with the EMA lerp
 * (shadow may be NULL to skip EMA). */
int dsk_adamw_ema_step(float* const* p, const float* const* g, float* const* m, float* const* v,
                       float* const* shadow, const int64_t* numel, int ntensors, int64_t max_numel, float lr,
                       float beta1, float beta2, float eps, float wd, int step, float ema_beta, float grad_scale,
                       void* stream);

/* ---- K2: backward kernels ----------------------------------------------------------------------
 * What loss.backward() runs below KarrasModule.loss_fn in training_step (karras/karrasmodule.py:1146-1155;
 * the reference relies on ATen autograd of the layers cited per function).  Gradients of activations use the
 * activation dtype (channels-last), parameter gradients are fp32 in the REFERENCE parameter layout.
 * `dres` (optional, same layout/dtype as the output, may alias it) is added to the result: gradient accumulation
 * for tensors with several consumers (ResNet identity, U-Net skips) without a separate pass. */

/* Data gradient of conv_same = conv_same(dY) with the taps flipped and Cin/Cout exchanged: pack the weights with this
 * function and call dsk_conv_fwd with Cin' = Cout, Cout' = Cin (residual = dres).  Backward of torch.nn.Conv2d/3d
 * (commonlayers.py:777-833, 53-58, 123-128; punetg.py:203-214; adm.py:268-284). */
int dsk_pack_conv_weight_dgrad(const float* w_ref, void* w_packed, int Cout, int Cin, int taps, int dtype, void* stream);
/* Weight gradient: dw[co][ci][tap] (+)= sum_pixels dy[p][co] * x[p + tap][ci]  (fp32, reference layout [Cout,Cin,k..]).
 * `d` describes the FORWARD conv: B,D,H,W = output size, in_dtype = dtype of x, out_dtype = dtype of dy, up2 = x is the
 * low-resolution input of conv(nearest_up2(x)); w_dtype selects the path (DSK_BF16: tcgen05 where the shape allows,
 * otherwise / DSK_F32: CUDA-core split-K, fp32 accumulate, deterministic).  ws: dsk_conv_wgrad_ws_bytes(d) bytes. */
int64_t dsk_conv_wgrad_ws_bytes(const dsk_conv_desc* d);
int dsk_conv_wgrad(const dsk_conv_desc* d, const void* x, const void* dy, float* dw, void* ws, int accumulate, void* stream);
/* Channel sums of a channels-last tensor dy [B, S, C]: out[b][c] (per_sample = 1: gradient of the time-embedding vector
 * added per sample and channel, commonlayers.py:826-829) or out[c] (bias gradient).  ws: dsk_bwd_ws_bytes(B,S,C). */
int64_t dsk_bwd_ws_bytes(int B, int64_t S, int C);
int dsk_channel_sum(const void* dy, float* out, void* ws, int B, int64_t S, int C, int dtype, int per_sample, void* stream);
/* out[c] = sum_r in[r][c] for fp32 matrices (bias gradients of Linear layers) */
int dsk_colsum_f32(const float* in, float* out, int64_t rows, int cols, int ld, void* stream);
/* Backward of dsk_norm_act.  x: the forward input; dy: gradient of the forward output; fwd_ws: the workspace the forward
 * call filled for this x (folded scale/shift table + (mean, rstd)); dx = d/dx (+ dres).  dgamma/dbeta [C], dfilm_* [B, C]
 * (null iff the corresponding forward operand was null).  ws: dsk_bwd_ws_bytes(B,S,C).  Autograd of GroupNorm /
 * GroupRMSNorm (+FiLM) + SiLU (commonlayers.py:362-384, 824-831; adm.py:305-329). */
int dsk_norm_act_bwd(const void* x, const void* dy, const void* dres, void* dx, const float* gamma, const float* beta,
                     const float* film_scale, const void* fwd_ws, float* dgamma, float* dbeta, float* dfilm_scale,
                     float* dfilm_shift, void* ws, int B, int64_t S, int C, int G, int mode, int silu, int dtype,
                     void* stream);
/* Backward of dsk_pool2x (x: forward input, needed for max; D,H,W: INPUT size) and of nearest x2 upsampling
 * (D,H,W: low-resolution size; dx = sum of the 2^ndim children of dy). */
int dsk_pool2x_bwd(const void* x, const void* dy, const void* dres, void* dx, int B, int D, int H, int W, int C, int ndim,
                   int is_max, int dtype, void* stream);
int dsk_upsample2x_bwd(const void* dy, const void* dres, void* dx, int B, int D, int H, int W, int C, int ndim, int dtype,
                       void* stream);
/* Softmax backward on rows, in place on dP: dS = P * (dP - sum_j dP_j P_j)  (attention.py:42-44,68). */
int dsk_softmax_bwd_rows(const float* P, float* dP, int64_t rows, int cols, void* stream);
/* SiLU forward / backward on fp32 vectors (time MLPs, commonlayers.py:516-550; adm.py:1047-1053). */
int dsk_silu_fwd(const float* z, float* a, int64_t n, void* stream);
int dsk_silu_bwd(const float* z, const float* da, float* dz, int64_t n, void* stream);
/* ReLU of the default MLPUncond (nets/mlp.py:20-35) in the training graph: a = max(z, 0); dz = da * [z > 0]. */
int dsk_relu_fwd(const float* z, float* a, int64_t n, void* stream);
int dsk_relu_bwd(const float* z, const float* da, float* dz, int64_t n, void* stream);
/* y = a (+ b) with per-operand dtypes (b may be NULL: cast/copy).  Gradient accumulation across fp32 / bf16 buffers. */
int dsk_add_ex(const void* a, int a_dtype, const void* b, int b_dtype, void* y, int y_dtype, int64_t n, void* stream);
/* Backward of dsk_concat_channels: da = dy[:, :Ca] (+ ra), db = dy[:, Ca:] (+ rb); da or db may be NULL. */
int dsk_split_channels(const void* dy, const void* ra, const void* rb, void* da, void* db, int64_t rows, int Ca, int Cb,
                       int dtype, void* stream);

/* ---- Whole-network evaluation from a non-Python host (dsk_plan_* / dsk_denoiser_fwd) ------------------------------
 * The launch list of a network evaluation is assembled by the Python plan (models/nets/punetg.py, adm.py).  A TAPE is that list
 * recorded once for a (network, batch, shape, precision): every entry point of this header that was called, with its arguments
 * -- pointers as (buffer, offset), descriptors by value --, the sizes of all device buffers, and the contents of the constant
 * ones (packed weights, parameters, pointer tables with their relocations).  diffsci_b200.tape.export_denoiser(module, B, shape,
 * path) writes it; this API replays it from C with no Python in the process:
 *     D(x; sigma) = c_skip x + c_out F(c_in x, c_noise)     KarrasModule.get_denoiser (karrasmodule.py:673-719) with the
 *     EDMPreconditioner scalars (preconditioners.py:30-53) computed on the device by dsk_edm_coeffs.
 * No hidden allocations: the caller provides ONE device workspace of dsk_plan_info(plan, DSK_PLAN_WORKSPACE_BYTES) bytes;
 * dsk_plan_bind uploads the constants into it (asynchronously on `stream`); dsk_denoiser_fwd enqueues the launches.
 * x, out: fp32 [B, C, *S] (the reference's NC(D)HW), sigma: fp32 [B], all device memory.  One host thread per plan. */
typedef struct dsk_plan dsk_plan;
enum { DSK_PLAN_WORKSPACE_BYTES = 0, DSK_PLAN_BATCH = 1, DSK_PLAN_CHANNELS = 2, DSK_PLAN_SAMPLE_ELEMS = 3, DSK_PLAN_LAUNCHES = 4 };
int dsk_plan_create_from_tape(const void* tape, int64_t nbytes, dsk_plan** plan);
int dsk_plan_load(const char* path, dsk_plan** plan);
int64_t dsk_plan_info(const dsk_plan* plan, int what);
int dsk_plan_bind(dsk_plan* plan, void* workspace, void* stream);
int dsk_denoiser_fwd(dsk_plan* plan, const float* x, const float* sigma, float* out, void* stream);
int dsk_plan_destroy(dsk_plan* plan);
/* EDM preconditioner scalars of a batch of noise levels (preconditioners.py:30-53): c_in = 1 / sqrt(s^2 + sd^2),
 * c_out = s sd / sqrt(s^2 + sd^2), c_skip = sd^2 / (s^2 + sd^2), c_noise = log(s) / 2, in the reference's fp32 operation order. */
int dsk_edm_coeffs(const float* sigma, float sigma_data, float* c_in, float* c_out, float* c_skip, float* c_noise, int B,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFSCI_B200_H */
