// conv_small.cu -- the first and last convolutions of the U-Nets (convin: Cin <= 4 -> C; convout: C -> Cout <= 4;
// reference nets/punetg.py:203-214, nets/adm.py:189-196).  They carry < 0.2% of the FLOPs but touch a full-resolution
// activation tensor, so they are HBM-bandwidth kernels, not GEMMs: one coalesced read (convout) or write (convin) of the
// C-channel tensor, fp32 accumulation on the CUDA cores, neighbouring taps served by L1.
#include <type_traits>

#include "common.cuh"

namespace dsk {

struct SmallConvArgs {
  const void* in;
  const float* w;        // packed fp32 [taps][Cin][Cout]
  const float* bias;
  void* out;
  float* out_nchw;
  int B, D, H, W, Cin, Cout, ks, ndim;
};

template <typename T> __device__ __forceinline__ void ld8f(const T* p, float* o);
template <> __device__ __forceinline__ void ld8f<float>(const float* p, float* o) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
template <> __device__ __forceinline__ void ld8f<__nv_bfloat16>(const __nv_bfloat16* p, float* o) {
  uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int j = 0; j < 4; ++j) { o[2 * j] = __low2float(h[j]); o[2 * j + 1] = __high2float(h[j]); }
}
template <> __device__ __forceinline__ void ld8f<__half>(const __half* p, float* o) {
  uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int j = 0; j < 4; ++j) { o[2 * j] = __low2float(h[j]); o[2 * j + 1] = __high2float(h[j]); }
}
template <typename T> __device__ __forceinline__ void st8f(T* p, const float* v);
template <> __device__ __forceinline__ void st8f<__half>(__half* p, const float* v) {
  uint4 o;
  __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
  *reinterpret_cast<uint4*>(p) = o;
}
template <> __device__ __forceinline__ void st8f<float>(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void st8f<__nv_bfloat16>(__nv_bfloat16* p, const float* v) {
  uint4 o;
  __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) oh[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
  *reinterpret_cast<uint4*>(p) = o;
}

// ---- convout: C (multiple of 8) -> COUT <= 4.  Thread = (run of FO_PW pixels along w, 8-channel chunk): the kw
// taps of neighbouring outputs share their input vectors in registers (FO_PW + 2 loads feed 3 * FO_PW tap products),
// the C/8 chunk-threads of a pixel run are adjacent lanes and meet by shuffle.
constexpr int FO_PW = 4;
template <typename TI, typename TO, int COUT>
__global__ void __launch_bounds__(256) conv_few_out_kernel(SmallConvArgs a) {
  extern __shared__ float wsm[];                      // [taps][Cin][COUT]
  const int taps = a.ndim == 3 ? a.ks * a.ks * a.ks : a.ks * a.ks;
  for (int i = threadIdx.x; i < taps * a.Cin * COUT; i += blockDim.x) wsm[i] = a.w[i];
  __syncthreads();
  const int chunks = a.Cin >> 3;                      // power of two <= 32 (checked on the host)
  const int wruns = (a.W + FO_PW - 1) / FO_PW;
  const int64_t nruns = (int64_t)a.B * a.D * a.H * wruns;
  const int64_t S = (int64_t)a.D * a.H * a.W;
  const TI* in = reinterpret_cast<const TI*>(a.in);
  const int r = a.ks >> 1;
  const int64_t total = nruns * chunks;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t padded = (total + 31) & ~(int64_t)31;  // whole warps iterate together (shuffles below)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += stride) {
    const bool live = i < total;
    int64_t t = live ? i / chunks : 0;
    const int ch = (int)(i % chunks) * 8;
    const int w0 = (int)(t % wruns) * FO_PW; t /= wruns;
    const int h0 = (int)(t % a.H); t /= a.H;
    const int d0 = (int)(t % a.D);
    const int b = (int)(t / a.D);
    float acc[FO_PW][COUT];
#pragma unroll
    for (int px = 0; px < FO_PW; ++px)
#pragma unroll
      for (int co = 0; co < COUT; ++co) acc[px][co] = 0.0f;
    if (live) {
      for (int kd = 0; kd < (a.ndim == 3 ? a.ks : 1); ++kd)
        for (int kh = 0; kh < a.ks; ++kh) {
          const int zd = a.ndim == 3 ? d0 + kd - r : 0, zh = h0 + kh - r;
          if ((unsigned)zd >= (unsigned)a.D || (unsigned)zh >= (unsigned)a.H) continue;
          const TI* rowp = in + (((int64_t)b * a.D + zd) * a.H + zh) * a.W * a.Cin + ch;
          const int tap0 = (kd * a.ks + kh) * a.ks;
          // input columns w0 - r .. w0 + FO_PW - 1 + r
#pragma unroll
          for (int xi = 0; xi < FO_PW + 2; ++xi) {
            const int zw = w0 + xi - 1;                // ks == 3: r == 1; ks == 1 uses only xi = 1..FO_PW
            if (a.ks == 1 && (xi == 0 || xi == FO_PW + 1)) continue;
            if ((unsigned)zw >= (unsigned)a.W) continue;
            float x[8];
            ld8f<TI>(rowp + (int64_t)zw * a.Cin, x);
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const int px = xi - kw;                  // output pixel (relative) that sees this column through tap kw
              if (px < 0 || px >= FO_PW) continue;
              if (a.ks == 1 && kw != 1) continue;
              const float* wp = wsm + ((int64_t)(tap0 + (a.ks == 1 ? 0 : kw)) * a.Cin + ch) * COUT;
#pragma unroll
              for (int e = 0; e < 8; ++e)
#pragma unroll
                for (int co = 0; co < COUT; ++co) acc[px][co] = fmaf(x[e], wp[e * COUT + co], acc[px][co]);
            }
          }
        }
    }
    for (int o = chunks >> 1; o > 0; o >>= 1)
#pragma unroll
      for (int px = 0; px < FO_PW; ++px)
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[px][co] += __shfl_xor_sync(0xffffffffu, acc[px][co], o);
    if (live && ch == 0) {
#pragma unroll
      for (int px = 0; px < FO_PW; ++px) {
        if (w0 + px >= a.W) break;
        const int64_t pix = (((int64_t)b * a.D + d0) * a.H + h0) * a.W + w0 + px;
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
          if (co >= a.Cout) break;
          float v = acc[px][co] + (a.bias != nullptr ? a.bias[co] : 0.0f);
          if (a.out_nchw != nullptr) a.out_nchw[((int64_t)b * a.Cout + co) * S + (pix - (int64_t)b * S)] = v;
          else reinterpret_cast<TO*>(a.out)[pix * a.Cout + co] = from_f32<TO>(v);
        }
      }
    }
  }
}

// ---- convin: CIN <= 4 -> Cout (multiple of 32).  Thread = (pixel, 32 output channels): the taps of a pixel are a
// handful of scalars, amortised over 32 accumulators; weights are smem broadcasts; the store is 4 x 16 B.
constexpr int FI_CO = 32;
template <typename TI, typename TO, int CIN>
__global__ void __launch_bounds__(256) conv_few_in_kernel(SmallConvArgs a) {
  extern __shared__ float wsm[];                      // [taps][CIN][Cout]
  const int taps = a.ndim == 3 ? a.ks * a.ks * a.ks : a.ks * a.ks;
  for (int i = threadIdx.x; i < taps * CIN * a.Cout; i += blockDim.x) wsm[i] = a.w[i];
  __syncthreads();
  const int chunks = a.Cout / FI_CO;
  const int64_t npix = (int64_t)a.B * a.D * a.H * a.W;
  const TI* in = reinterpret_cast<const TI*>(a.in);
  TO* out = reinterpret_cast<TO*>(a.out);
  const int r = a.ks >> 1;
  const int64_t total = npix * chunks;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // pixel-major within a warp: consecutive lanes = consecutive pixels (coalesced tap loads), chunk outermost per block row
    const int64_t pix = i % npix;
    const int ch = (int)(i / npix) * FI_CO;
    int64_t t = pix;
    const int w0 = (int)(t % a.W); t /= a.W;
    const int h0 = (int)(t % a.H); t /= a.H;
    const int d0 = (int)(t % a.D);
    const int b = (int)(t / a.D);
    float acc[FI_CO];
#pragma unroll
    for (int e = 0; e < FI_CO; ++e) acc[e] = a.bias != nullptr ? a.bias[ch + e] : 0.0f;
    int tap = 0;
    for (int kd = 0; kd < (a.ndim == 3 ? a.ks : 1); ++kd)
      for (int kh = 0; kh < a.ks; ++kh)
        for (int kw = 0; kw < a.ks; ++kw, ++tap) {
          const int zd = a.ndim == 3 ? d0 + kd - r : 0, zh = h0 + kh - r, zw = w0 + kw - r;
          if ((unsigned)zd >= (unsigned)a.D || (unsigned)zh >= (unsigned)a.H || (unsigned)zw >= (unsigned)a.W) continue;
          const TI* ip = in + ((((int64_t)b * a.D + zd) * a.H + zh) * a.W + zw) * CIN;
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) {
            const float x = to_f32<TI>(ip[ci]);
            const float4* wp = reinterpret_cast<const float4*>(wsm + ((int64_t)tap * CIN + ci) * a.Cout + ch);
#pragma unroll
            for (int e4 = 0; e4 < FI_CO / 4; ++e4) {
              const float4 w4 = wp[e4];
              acc[4 * e4] = fmaf(x, w4.x, acc[4 * e4]);
              acc[4 * e4 + 1] = fmaf(x, w4.y, acc[4 * e4 + 1]);
              acc[4 * e4 + 2] = fmaf(x, w4.z, acc[4 * e4 + 2]);
              acc[4 * e4 + 3] = fmaf(x, w4.w, acc[4 * e4 + 3]);
            }
          }
        }
#pragma unroll
    for (int e8 = 0; e8 < FI_CO / 8; ++e8) st8f<TO>(out + pix * a.Cout + ch + e8 * 8, acc + e8 * 8);
  }
}

// k = 3, CIN * taps <= 54 (Cin = 1, or 2 in 2-D): thread = one pixel, ALL output channels.  The receptive field (27 or 9 taps
// per input channel) is gathered once into registers -- zero for taps outside the image, so the inner loops carry no bounds
// logic -- and reused for every 32-channel chunk of the output; 32-bit index arithmetic (the pixel-times-chunk form above
// spends as many instructions on its 64-bit divisions as on its FMAs).  Weights: shared-memory float4 broadcasts, which is what
// bounds it (8 LDS.128 per 32 FMAs: 387 us for C4's first layer, 537 MB of fp32 output).  Tried and dropped: the thread's weights
// in registers (4 channels x 2 pixels per thread, 174 registers): one resident block per SM, latency-bound, 697 us.  The
// 16-bit-operand modes take the tensor-core im2col kernel instead (convin_tc.cu, exact split rows).
template <typename TI, typename TO, int CIN, bool D3>
__global__ void __launch_bounds__(256) conv_few_in_px_kernel(SmallConvArgs a) {
  extern __shared__ float wsm[];                      // [taps][CIN][Cout]
  constexpr int TAPS = D3 ? 27 : 9, KD = D3 ? 3 : 1;
  for (int i = threadIdx.x; i < TAPS * CIN * a.Cout; i += blockDim.x) wsm[i] = a.w[i];
  __syncthreads();
  const TI* in = reinterpret_cast<const TI*>(a.in);
  TO* out = reinterpret_cast<TO*>(a.out);
  const uint32_t npix = (uint32_t)a.B * a.D * a.H * a.W;            // host: < 2^31
  for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
    uint32_t t = pix;
    const int w0 = (int)(t % (uint32_t)a.W); t /= (uint32_t)a.W;
    const int h0 = (int)(t % (uint32_t)a.H); t /= (uint32_t)a.H;
    const int d0 = D3 ? (int)(t % (uint32_t)a.D) : 0;
    float xv[TAPS * CIN];
#pragma unroll
    for (int kd = 0; kd < KD; ++kd)
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int zd = D3 ? d0 + kd - 1 : 0, zh = h0 + kh - 1, zw = w0 + kw - 1;
          const bool ok = (unsigned)zd < (unsigned)a.D && (unsigned)zh < (unsigned)a.H && (unsigned)zw < (unsigned)a.W;
          const int off = (((D3 ? kd - 1 : 0) * a.H + (kh - 1)) * a.W + (kw - 1)) * CIN;
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci)
            xv[((kd * 3 + kh) * 3 + kw) * CIN + ci] = ok ? to_f32<TI>(in[(int64_t)pix * CIN + off + ci]) : 0.0f;
        }
#pragma unroll 1
    for (int ch = 0; ch < a.Cout; ch += FI_CO) {
      float acc[FI_CO];
#pragma unroll
      for (int e4 = 0; e4 < FI_CO / 4; ++e4) {
        const float4 b4 = a.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(a.bias + ch) + e4) : make_float4(0.f, 0.f, 0.f, 0.f);
        acc[4 * e4] = b4.x; acc[4 * e4 + 1] = b4.y; acc[4 * e4 + 2] = b4.z; acc[4 * e4 + 3] = b4.w;
      }
#pragma unroll
      for (int k = 0; k < TAPS * CIN; ++k) {
        const float x = xv[k];
        const float4* wp = reinterpret_cast<const float4*>(wsm + k * a.Cout + ch);
#pragma unroll
        for (int e4 = 0; e4 < FI_CO / 4; ++e4) {
          const float4 w4 = wp[e4];
          acc[4 * e4] = fmaf(x, w4.x, acc[4 * e4]);
          acc[4 * e4 + 1] = fmaf(x, w4.y, acc[4 * e4 + 1]);
          acc[4 * e4 + 2] = fmaf(x, w4.z, acc[4 * e4 + 2]);
          acc[4 * e4 + 3] = fmaf(x, w4.w, acc[4 * e4 + 3]);
        }
      }
#pragma unroll
      for (int e8 = 0; e8 < FI_CO / 8; ++e8) st8f<TO>(out + (int64_t)pix * a.Cout + ch + e8 * 8, acc + e8 * 8);
    }
  }
}

template <typename TI, typename TO>
static int launch_few_out(const SmallConvArgs& a, cudaStream_t st) {
  const int taps = a.ndim == 3 ? a.ks * a.ks * a.ks : a.ks * a.ks;
  const size_t smem = (size_t)taps * a.Cin * 4 * sizeof(float);
  const int64_t total = (int64_t)a.B * a.D * a.H * ((a.W + FO_PW - 1) / FO_PW) * (a.Cin / 8);
  const int grid = grid_for(total, 256, 16);
  switch (a.Cout) {
    case 1: DSK_LAUNCH((conv_few_out_kernel<TI, TO, 1>), grid, 256, smem, st, a); break;
    case 2: DSK_LAUNCH((conv_few_out_kernel<TI, TO, 2>), grid, 256, smem, st, a); break;
    case 3: DSK_LAUNCH((conv_few_out_kernel<TI, TO, 3>), grid, 256, smem, st, a); break;
    default: DSK_LAUNCH((conv_few_out_kernel<TI, TO, 4>), grid, 256, smem, st, a); break;
  }
  return DSK_OK;
}

template <typename TI, typename TO>
static int launch_few_in(const SmallConvArgs& a, cudaStream_t st) {
  const int taps = a.ndim == 3 ? a.ks * a.ks * a.ks : a.ks * a.ks;
  const size_t smem = (size_t)taps * a.Cin * a.Cout * sizeof(float);
  const int64_t npix = (int64_t)a.B * a.D * a.H * a.W;
  if (a.ks == 3 && taps * a.Cin <= 54 && npix < (1ll << 31) && npix * a.Cin < (1ll << 31)) {   // pixel-major kernel (see above)
    const int pgrid = grid_for(npix, 256, 16);
    if (a.ndim == 3 && a.Cin == 1) DSK_LAUNCH((conv_few_in_px_kernel<TI, TO, 1, true>), pgrid, 256, smem, st, a);
    else if (a.ndim == 3) DSK_LAUNCH((conv_few_in_px_kernel<TI, TO, 2, true>), pgrid, 256, smem, st, a);
    else if (a.Cin == 1) DSK_LAUNCH((conv_few_in_px_kernel<TI, TO, 1, false>), pgrid, 256, smem, st, a);
    else if (a.Cin == 2) DSK_LAUNCH((conv_few_in_px_kernel<TI, TO, 2, false>), pgrid, 256, smem, st, a);
    else if (a.Cin == 3) DSK_LAUNCH((conv_few_in_px_kernel<TI, TO, 3, false>), pgrid, 256, smem, st, a);
    else DSK_LAUNCH((conv_few_in_px_kernel<TI, TO, 4, false>), pgrid, 256, smem, st, a);
    return DSK_OK;
  }
  const int64_t total = (int64_t)a.B * a.D * a.H * a.W * (a.Cout / FI_CO);
  const int grid = grid_for(total, 256, 16);
  switch (a.Cin) {
    case 1: DSK_LAUNCH((conv_few_in_kernel<TI, TO, 1>), grid, 256, smem, st, a); break;
    case 2: DSK_LAUNCH((conv_few_in_kernel<TI, TO, 2>), grid, 256, smem, st, a); break;
    case 3: DSK_LAUNCH((conv_few_in_kernel<TI, TO, 3>), grid, 256, smem, st, a); break;
    default: DSK_LAUNCH((conv_few_in_kernel<TI, TO, 4>), grid, 256, smem, st, a); break;
  }
  return DSK_OK;
}

int convin_tc_dispatch(const dsk_conv_desc* d, const void* in, const void* w, const float* bias, void* out, cudaStream_t st);

// returns 1 if handled, 0 if the shape is not a few-channel conv, negative on error
int conv_small_dispatch(const dsk_conv_desc* d, const void* in, const void* w, const float* bias, const float* chan_bias,
                        const void* residual, void* out, cudaStream_t st) {
  if (chan_bias != nullptr || residual != nullptr || d->up2) return 0;
  const int taps = d->ndim == 3 ? d->ksize * d->ksize * d->ksize : d->ksize * d->ksize;
  SmallConvArgs a{in, (const float*)w, bias, d->out_nchw_f32 ? nullptr : out, d->out_nchw_f32 ? (float*)out : nullptr,
                  d->B, d->D, d->H, d->W, d->Cin, d->Cout, d->ksize, d->ndim};
  const int chunks = d->Cin / 8;
  const bool few_out = d->Cout <= 4 && d->Cin % 8 == 0 && chunks <= 32 && (chunks & (chunks - 1)) == 0 &&
                       (size_t)taps * d->Cin * 4 * 4 <= 48 * 1024 && (d->ksize == 3 || d->ksize == 1);
  const bool few_in = d->Cin <= 4 && d->Cout % FI_CO == 0 && !d->out_nchw_f32 && (size_t)taps * d->Cin * d->Cout * 4 <= 48 * 1024;
  if (!few_out && !few_in) return 0;
  if (few_in && !few_out) {      // tensor-core im2col form (convin_tc.cu) where it applies (its gather wraps for circular padding)
    const int rc = convin_tc_dispatch(d, in, w, bias, out, st);
    if (rc != DSK_ERR_UNSUPPORTED) return rc == DSK_OK ? 1 : rc;
  }
  if (d->circular) return 0;     // the CUDA-core few-channel kernels zero-pad: hand over to the generic wrapping kernel
  const int ti = d->in_dtype, to = d->out_nchw_f32 ? DSK_F32 : d->out_dtype;
  int rc;
#define GO(FN)                                                                                        \
  if (ti == DSK_F32 && to == DSK_F32) rc = FN<float, float>(a, st);                                   \
  else if (ti == DSK_BF16 && to == DSK_BF16) rc = FN<__nv_bfloat16, __nv_bfloat16>(a, st);            \
  else if (ti == DSK_BF16 && to == DSK_F32) rc = FN<__nv_bfloat16, float>(a, st);                     \
  else if (ti == DSK_F16 && to == DSK_F16) rc = FN<__half, __half>(a, st);                            \
  else if (ti == DSK_F16 && to == DSK_F32) rc = FN<__half, float>(a, st);                             \
  else if (ti == DSK_F32 && to == DSK_F16) rc = FN<float, __half>(a, st);                             \
  else rc = FN<float, __nv_bfloat16>(a, st);
  if (few_out) { GO(launch_few_out) } else { GO(launch_few_in) }
#undef GO
  return rc == DSK_OK ? 1 : rc;
}

// ---- weight gradient of the few-channel convolutions --------------------------------------------------------------
//   dW[co][ci][tap] = sum_pix dY[pix][co] * X[pix + tap][ci]
// One side is WIDE (32 / 64 / 128 channels), the other NARROW (<= 4): FEW_IN: narrow = Cin (convin), wide = Cout;
// otherwise narrow = Cout (convout), wide = Cin.  Thread = (wide channel, group of taps); a block walks a contiguous
// pixel range keeping its [taps-in-group][narrow] partial sums in registers: the wide tensor is read exactly once
// (convin: dY once per pixel; convout: X once per tap, the 27 shifted re-reads hit L1), the narrow one is broadcast.
// Block partials go to ws[block][tap*Cin + ci][co] and are summed in a fixed order by wgrad_reduce_kernel.
constexpr int WF_NARROW = 4, WF_THREADS = 256;
template <typename TX, typename TG, bool FEW_IN, int TPG>
__global__ void __launch_bounds__(WF_THREADS) wgrad_few_kernel(const TX* __restrict__ x, const TG* __restrict__ dy, float* __restrict__ ws,
                                                                int B, int D, int H, int W, int Cin, int Cout, int ndim,
                                                                int64_t pix_per_block) {
  const int taps = ndim == 3 ? 27 : 9;
  const int Cw = FEW_IN ? Cout : Cin, Cn = FEW_IN ? Cin : Cout;
  const int wch = threadIdx.x % Cw, grp = threadIdx.x / Cw;
  const int t_lo = grp * TPG, t_hi = min(taps, t_lo + TPG);
  float acc[TPG][WF_NARROW];
#pragma unroll
  for (int t = 0; t < TPG; ++t)
#pragma unroll
    for (int n = 0; n < WF_NARROW; ++n) acc[t][n] = 0.0f;
  // the block owns a contiguous range of (b, d, h) rows; tap validity in d and h and the tap's pixel offset are per row,
  // only the w bound is checked per pixel
  const int64_t rows = (int64_t)B * D * H;
  const int64_t r0 = (int64_t)blockIdx.x * pix_per_block, r1 = min(rows, r0 + pix_per_block);
  for (int64_t row = r0; row < r1; ++row) {
    const int h0 = (int)(row % H), d0 = (int)((row / H) % D);
    const int64_t base = row * W;
    bool ok[TPG];
    int off[TPG], kwm[TPG];
#pragma unroll
    for (int t = 0; t < TPG; ++t) {
      const int tap = t_lo + t;
      const int kw = tap % 3, kh = (tap / 3) % 3, kd = tap / 9;
      const int zd = ndim == 3 ? d0 + kd - 1 : 0, zh = h0 + kh - 1;
      ok[t] = tap < t_hi && (unsigned)zd < (unsigned)D && (unsigned)zh < (unsigned)H;
      off[t] = ((zd - d0) * H + (zh - h0)) * W + kw - 1;
      kwm[t] = kw - 1;
    }
#pragma unroll 4
    for (int w0 = 0; w0 < W; ++w0) {
      const int64_t pix = base + w0;
      float gw = 0.0f, gn[WF_NARROW];
      if (FEW_IN) {
        gw = to_f32<TG>(dy[pix * Cout + wch]);
      } else {
#pragma unroll
        for (int n = 0; n < WF_NARROW; ++n) gn[n] = n < Cn ? to_f32<TG>(dy[pix * Cout + n]) : 0.0f;
      }
#pragma unroll
      for (int t = 0; t < TPG; ++t) {
        if (ok[t] && (unsigned)(w0 + kwm[t]) < (unsigned)W) {
          const int64_t q = pix + off[t];
          if (FEW_IN) {
#pragma unroll
            for (int n = 0; n < WF_NARROW; ++n)
              if (n < Cn) acc[t][n] = fmaf(gw, to_f32<TX>(x[q * Cin + n]), acc[t][n]);
          } else {
            const float xv = to_f32<TX>(x[q * Cin + wch]);
#pragma unroll
            for (int n = 0; n < WF_NARROW; ++n) acc[t][n] = fmaf(gn[n], xv, acc[t][n]);
          }
        }
      }
    }
  }
  float* o = ws + (int64_t)blockIdx.x * taps * Cin * Cout;
#pragma unroll
  for (int t = 0; t < TPG; ++t) {
    const int tap = t_lo + t;
    if (tap < t_hi) {
#pragma unroll
      for (int n = 0; n < WF_NARROW; ++n)
        if (n < Cn) {
          const int ci = FEW_IN ? n : wch, co = FEW_IN ? wch : n;
          o[((int64_t)tap * Cin + ci) * Cout + co] = acc[t][n];
        }
    }
  }
}

int wgrad_reduce_launch(const float* ws, float* dw, int Cout, int Cin, int taps, int nsplit, int accumulate, cudaStream_t st);

static bool wgrad_few_shape(const dsk_conv_desc* d, bool* few_in, int* tpg) {
  if (d->ksize != 3 || d->up2 || d->circular) return false;
  const int taps = d->ndim == 3 ? 27 : 9;
  const bool fi = d->Cin <= WF_NARROW, fo = d->Cout <= WF_NARROW;
  if (fi == fo) return false;                                   // both narrow (tiny) or neither
  const int Cw = fi ? d->Cout : d->Cin;
  if (Cw != 32 && Cw != 64 && Cw != 128) return false;
  const int groups = WF_THREADS / Cw;
  const int need = (taps + groups - 1) / groups;
  *few_in = fi;
  *tpg = need <= 7 ? 7 : 14;
  return need <= 14;
}

static int wgrad_few_blocks(const dsk_conv_desc* d) {
  const int64_t rows = (int64_t)d->B * d->D * d->H;             // the unit of work is one (b, d, h) row of W pixels
  int64_t blocks = rows;
  if (blocks > 16 * DSK_NUM_SMS) blocks = 16 * DSK_NUM_SMS;
  return (int)(blocks < 1 ? 1 : blocks);
}

// ---- few-channel weight gradient on the tensor cores (bf16) ----------------------------------------------------------
// The CUDA-core kernel above keeps one pixel in flight per block iteration and is latency-bound (2.3 ms for the convin of
// a 2 x 64^3 volume -- 100x its HBM roofline).  In bf16 mode the same sum is a GEMM with a tiny MN extent and a huge
// reduction:   FEW_IN : dW[(tap,ci)][co] = sum_pix  Ncol[pix][(tap,ci)] * dY[pix][co],   Ncol = im2col of the NARROW x
//              FEW_OUT: dW[ci][(tap,co)] = sum_q    X[q][ci] * Ncol[q][(tap,co)],        Ncol[q][(tap,co)] = dY[q - off(tap)][co]
// Only the narrow tensor is expanded (pixels x KP bf16, KP = taps*Cn rounded up to 8: 27 -> 32), the wide tensor is read once,
// in place, as an MN-major operand of the tcgen05 GEMM (dsk_gemm_bf16_tc, transA = transB = 1), split over the pixels into
// `nsplit` batch entries whose fp32 partials are summed in a fixed order.  Zero or circular padding is decided in the
// expansion kernel, so periodic convolutions take this path too.
struct FewTcPlan {
  bool ok, few_in;
  int taps, Cn, Cw, KP, nsplit;
  int64_t pix, kc;
};

static FewTcPlan few_tc_plan(const dsk_conv_desc* d) {
  FewTcPlan f{};
  if (d->ksize != 3 || d->up2 || d->in_dtype != DSK_BF16 || d->out_dtype != DSK_BF16) return f;
  const bool fi = d->Cin <= WF_NARROW, fo = d->Cout <= WF_NARROW;
  if (fi == fo) return f;
  f.few_in = fi;
  f.taps = d->ndim == 3 ? 27 : 9;
  f.Cn = fi ? d->Cin : d->Cout;
  f.Cw = fi ? d->Cout : d->Cin;
  if (f.Cw % 8 != 0 || f.Cw > 256) return f;
  f.KP = (f.taps * f.Cn + 7) & ~7;
  f.pix = (int64_t)d->B * d->D * d->H * d->W;
  // the GEMM reads the wide tensor in place: the split must tile the pixels exactly, in multiples of 8 (16-byte TMA rows)
  for (int ns = DSK_NUM_SMS; ns >= 16; --ns)
    if (f.pix % ns == 0 && (f.pix / ns) % 8 == 0) { f.nsplit = ns; break; }
  if (f.nsplit == 0) return f;
  f.kc = f.pix / f.nsplit;
  f.ok = true;
  return f;
}

static int64_t few_tc_part_bytes(const FewTcPlan& f) {
  const int64_t b = (int64_t)f.nsplit * f.KP * f.Cw * (int64_t)sizeof(float);
  return (b + 1023) & ~(int64_t)1023;
}

// Ncol[pix][j], j = tap*Cn + n (zero for j >= taps*Cn): the narrow tensor shifted by +off(tap) (FEW_IN: x under the tap)
// or by -off(tap) (FEW_OUT: the dY that this x pixel meets through the tap); zero outside the image, or wrapped.
__global__ void __launch_bounds__(256) few_expand_kernel(const __nv_bfloat16* __restrict__ nar, __nv_bfloat16* __restrict__ col, int B, int D,
                                                          int H, int W, int Cn, int taps, int KP, int ndim, int sign, int circ) {
  const int64_t pix = (int64_t)B * D * H * W;
  const int chunks = KP / 8;
  const int64_t total = pix * chunks;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chunks);
    int64_t t = i / chunks;
    const int w0 = (int)(t % W); t /= W;
    const int h0 = (int)(t % H); t /= H;
    const int d0 = (int)(t % D);
    const int b = (int)(t / D);
    uint4 o;
    __nv_bfloat16* ov = reinterpret_cast<__nv_bfloat16*>(&o);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = ch * 8 + e;
      __nv_bfloat16 v = __float2bfloat16_rn(0.0f);
      if (j < taps * Cn) {
        const int tap = j / Cn, n = j - tap * Cn;
        const int kw = tap % 3, kh = (tap / 3) % 3, kd = tap / 9;
        int zw = w0 + sign * (kw - 1), zh = h0 + sign * (kh - 1), zd = ndim == 3 ? d0 + sign * (kd - 1) : 0;
        bool ok = true;
        if (circ) {
          zw = zw < 0 ? zw + W : (zw >= W ? zw - W : zw);
          zh = zh < 0 ? zh + H : (zh >= H ? zh - H : zh);
          zd = zd < 0 ? zd + D : (zd >= D ? zd - D : zd);
        } else {
          ok = (unsigned)zw < (unsigned)W && (unsigned)zh < (unsigned)H && (unsigned)zd < (unsigned)D;
        }
        if (ok) v = nar[((((int64_t)b * D + zd) * H + zh) * W + zw) * Cn + n];
      }
      ov[e] = v;
    }
    reinterpret_cast<uint4*>(col)[i] = o;
  }
}

// dw[co][ci][tap] (+)= sum_z part[z][..]:  FEW_IN: part[z][tap*Cin + ci][co] (rows KP, cols Cout);
//                                          FEW_OUT: part[z][ci][tap*Cout + co] (rows Cin, cols KP)
__global__ void __launch_bounds__(256) few_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int Cout, int Cin, int taps,
                                                          int KP, int few_in, int nsplit, int accumulate) {
  const int total = Cout * Cin * taps;
  const int64_t zs = (int64_t)KP * (few_in ? Cout : Cin);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int co, ci, tap;
    int64_t src;
    if (few_in) {                 // i walks part rows: coalesced over co
      co = i % Cout; const int m = i / Cout; tap = m / Cin; ci = m - tap * Cin;
      src = (int64_t)m * Cout + co;
    } else {                      // i walks part row ci, column (tap, co)
      const int j = i % (taps * Cout); ci = i / (taps * Cout); tap = j / Cout; co = j - tap * Cout;
      src = (int64_t)ci * KP + j;
    }
    float s = 0.0f;
    for (int z = 0; z < nsplit; ++z) s += part[z * zs + src];
    float* o = dw + ((int64_t)co * Cin + ci) * taps + tap;
    *o = accumulate ? *o + s : s;
  }
}

}  // namespace dsk
extern "C" int dsk_gemm_bf16_tc(const void* A, const void* Bm, void* C, const float* bias, int bias_rows, const void* residual,
                                int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc, int64_t strideA, int64_t strideB,
                                int64_t strideC, int batch, float alpha, int out_f32, int transA, int transB, void* stream);
namespace dsk {

static int few_tc_run(const dsk_conv_desc* d, const FewTcPlan& f, const void* x, const void* dy, float* dw, void* ws, int accumulate,
                      cudaStream_t st) {
  float* part = (float*)ws;
  __nv_bfloat16* col = (__nv_bfloat16*)((uint8_t*)ws + few_tc_part_bytes(f));
  const void* nar = f.few_in ? x : dy;
  DSK_LAUNCH(few_expand_kernel, grid_for(f.pix * (f.KP / 8), 256, 16), 256, 0, st, (const __nv_bfloat16*)nar, col, d->B, d->D, d->H, d->W,
             f.Cn, f.taps, f.KP, d->ndim, f.few_in ? 1 : -1, d->circular);
  int rc;
  if (f.few_in)     // C[(tap,ci)][co] = sum_pix col[pix][(tap,ci)] dY[pix][co]
    rc = dsk_gemm_bf16_tc(col, dy, part, nullptr, 0, nullptr, f.KP, d->Cout, (int)f.kc, f.KP, d->Cout, d->Cout, f.kc * f.KP,
                          f.kc * d->Cout, (int64_t)f.KP * d->Cout, f.nsplit, 1.0f, 1, 1, 1, st);
  else              // C[ci][(tap,co)] = sum_q X[q][ci] col[q][(tap,co)]
    rc = dsk_gemm_bf16_tc(x, col, part, nullptr, 0, nullptr, d->Cin, f.KP, (int)f.kc, d->Cin, f.KP, f.KP, f.kc * d->Cin, f.kc * f.KP,
                          (int64_t)d->Cin * f.KP, f.nsplit, 1.0f, 1, 1, 1, st);
  if (rc != DSK_OK) return rc;
  DSK_LAUNCH(few_reduce_kernel, grid_for((int64_t)d->Cout * d->Cin * f.taps, 256, 8), 256, 0, st, part, dw, d->Cout, d->Cin, f.taps, f.KP,
             f.few_in ? 1 : 0, f.nsplit, accumulate);
  return DSK_OK;
}

int64_t wgrad_few_ws_bytes(const dsk_conv_desc* d) {
  const FewTcPlan f = few_tc_plan(d);
  if (f.ok) return few_tc_part_bytes(f) + f.pix * f.KP * 2;
  bool fi; int tpg;
  if (!wgrad_few_shape(d, &fi, &tpg)) return 0;
  return (int64_t)wgrad_few_blocks(d) * (d->ndim == 3 ? 27 : 9) * d->Cin * d->Cout * (int64_t)sizeof(float);
}

// returns 1 if handled, 0 if the shape is not a few-channel conv, negative on error
int wgrad_few_dispatch(const dsk_conv_desc* d, const void* x, const void* dy, float* dw, void* ws, int accumulate, cudaStream_t st) {
  static const int old_path = [] { const char* e = getenv("DSK_WGRAD_FEW_OLD"); return e ? atoi(e) : 0; }();   // A/B measurements
  const FewTcPlan f = few_tc_plan(d);
  if (f.ok && !old_path) {
    const int rc = few_tc_run(d, f, x, dy, dw, ws, accumulate, st);
    return rc == DSK_OK ? 1 : rc;
  }
  bool fi; int tpg;
  if (!wgrad_few_shape(d, &fi, &tpg)) return 0;
  const int blocks = wgrad_few_blocks(d);
  const int64_t total = (int64_t)d->B * d->D * d->H;           // rows
  const int64_t ppb = (total + blocks - 1) / blocks;
  const int used = (int)((total + ppb - 1) / ppb);
#define WF_GO(TX, TG)                                                                                                              \
  do {                                                                                                                             \
    if (fi && tpg == 7) DSK_LAUNCH((wgrad_few_kernel<TX, TG, true, 7>), used, WF_THREADS, 0, st, (const TX*)x, (const TG*)dy, (float*)ws, \
                                   d->B, d->D, d->H, d->W, d->Cin, d->Cout, d->ndim, ppb);                                          \
    else if (fi) DSK_LAUNCH((wgrad_few_kernel<TX, TG, true, 14>), used, WF_THREADS, 0, st, (const TX*)x, (const TG*)dy, (float*)ws,   \
                            d->B, d->D, d->H, d->W, d->Cin, d->Cout, d->ndim, ppb);                                                 \
    else if (tpg == 7) DSK_LAUNCH((wgrad_few_kernel<TX, TG, false, 7>), used, WF_THREADS, 0, st, (const TX*)x, (const TG*)dy,          \
                                  (float*)ws, d->B, d->D, d->H, d->W, d->Cin, d->Cout, d->ndim, ppb);                               \
    else DSK_LAUNCH((wgrad_few_kernel<TX, TG, false, 14>), used, WF_THREADS, 0, st, (const TX*)x, (const TG*)dy, (float*)ws, d->B,    \
                    d->D, d->H, d->W, d->Cin, d->Cout, d->ndim, ppb);                                                               \
  } while (0)
  if (d->in_dtype == DSK_F32 && d->out_dtype == DSK_F32) WF_GO(float, float);
  else if (d->in_dtype == DSK_BF16 && d->out_dtype == DSK_BF16) WF_GO(__nv_bfloat16, __nv_bfloat16);
  else if (d->in_dtype == DSK_F32 && d->out_dtype == DSK_BF16) WF_GO(float, __nv_bfloat16);
  else WF_GO(__nv_bfloat16, float);
#undef WF_GO
  const int rc = wgrad_reduce_launch((const float*)ws, dw, d->Cout, d->Cin, d->ndim == 3 ? 27 : 9, used, accumulate, st);
  return rc == DSK_OK ? 1 : rc;
}

}  // namespace dsk
