// convout_tc.cu -- the LAST convolution of the U-Nets (C -> Cout <= 3, k = 3; reference nets/punetg.py:209-214,
// nets/adm.py:189-196) on the tensor cores without paying for 64 output channels.
//
//   out[p, co] = b[co] + sum_{kd,kh,kw} sum_ci X[p + (kd,kh,kw) - 1, ci] W[co, ci, kd, kh, kw]
//
// An implicit GEMM with M = pixels and N = Cout wastes the machine: every tap re-reads the 128 x 64 activation tile from
// shared memory for an N of 1..3 (the N = 16 tile of conv_tc.cu ran at 1/10 of the HBM roofline).  Here the IN-PLANE TAPS
// are the N dimension instead:
//
//   Z_pp[r, (kh,kw,co)] = sum_kd  X_patch(pp + kd)[r, :] . W[kd][(kh,kw,co), :]          (tcgen05, accumulated over kd in TMEM)
//   out[pp, h, w, co]   = b[co] + sum_{kh,kw} Z_pp[(h + kh) * 10 + (w + kw), (kh,kw,co)]    (epilogue: shifted gather in smem)
//
// r runs over ALL 180 pixels of the halo'd 18 x 10 patch (two M = 128 row groups: rows 0..127 and 52..179, the second a
// row-shifted view of the same TMA patch), so an activation patch is read from shared memory 2 x (number of output planes
// it feeds) times instead of 27 x.  N = 9 * Cout padded to 16 / 32.  The kernel is then bound by the L2 -> SM patch
// stream (halo factor ~2.8), not by the tensor core's operand port.
#include <stdlib.h>

#include "tc_common.cuh"

namespace dsk {

constexpr int CO_BW = 8, CO_BH = 16, CO_PW = CO_BW + 2, CO_PH = CO_BH + 2;
constexpr int CO_ROWS = CO_PW * CO_PH;               // 180 patch pixels
constexpr int CO_PATCH_BYTES = CO_ROWS * 128;        // 23040
constexpr int CO_PATCH_STRIDE = 23552;
constexpr int CO_G1 = CO_ROWS - 128;                 // 52: first row of the second M = 128 row group
constexpr int CO_THREADS = 192;                      // warp 0: TMA, warp 1: MMA + TMEM owner, warps 2-5: epilogue
constexpr int CO_NA = 6;   // CO_P (template): output planes per tile -- 3-D depth halo (P + 2) / P of the L2 -> SM patch stream

struct CoParams {
  int B, D, H, W, Cin, Cout, KD;      // as the tensor map sees them (2-D: B = 1, D = batch of planes)
  int tiles_w, tiles_h, groups_d, total_tiles;
  const uint16_t* w;                  // packed [tap][16][w_ld] (PackedConv convout layout, zero rows beyond Cout)
  const float* bias;
  uint16_t* out;                      // channels-last [.., Cout] in the 16-bit format (fp32 if out_f32), or
  float* out_nchw;                    // fp32 NC(D)HW
  int planes_per_sample;
  int pad_hw, pad_d;                  // circular padding: the tensor map covers the halo-padded copy (see conv_tc.cu)
  int f16, out_f32;                   // 16-bit format (0 bf16, 1 fp16); channels-last output is fp32
  int vparts, a_lo_off, w_lo_off, w_ld;   // split operands: virtual K chunks per real chunk (conv_tc.cu), weight row length
};
__device__ __forceinline__ int co_vchunk_a(const CoParams& p, int vc) {
  const int c = vc / p.vparts, part = vc - c * p.vparts;
  return c * 64 + ((p.vparts > 1 && part == p.vparts - 1) ? p.a_lo_off : 0);
}
__device__ __forceinline__ int co_vchunk_w(const CoParams& p, int vc) {
  const int c = vc / p.vparts, part = vc - c * p.vparts;
  return c * 64 + ((p.vparts == 3 && part == 1) ? p.w_lo_off : 0);
}

template <int COUT, int CO_P = 2>
__global__ void __launch_bounds__(CO_THREADS, 1)
convout_tc_kernel(const __grid_constant__ CUtensorMap tmapA, const CoParams p) {
  constexpr int NT = 9 * COUT, N_PAD = NT <= 16 ? 16 : 32;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int KD = p.KD, nchunks = (p.Cin / 64) * p.vparts;           // virtual K chunks
  uint8_t* sA = smem;                                               // CO_NA patches
  uint8_t* sW = smem + (size_t)CO_NA * CO_PATCH_STRIDE;             // [kd][chunk][N_PAD rows x 128 B], SWIZZLE_128B
  const int wbytes = KD * nchunks * N_PAD * 128;
  float* sZ = reinterpret_cast<float*>(sW + ((wbytes + 1023) & ~1023));   // [P][180][NT] gather buffer
  __shared__ uint64_t full_a[CO_NA], empty_a[CO_NA], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  constexpr uint32_t ACC_COLS = CO_P * 2 * N_PAD;                   // [plane][row group] x N_PAD columns
  // split operands: a second accumulator set takes the hi*lo + lo*hi products (lo operands are stored times 2^11); the epilogue
  // combines z = hh + 2^-11 lo.  Chains are short here (taps are the N dimension: KD * 4 MMAs per accumulator and chunk).
  const uint32_t nsets = p.vparts > 1 ? 2u : 1u;
  const uint32_t BUF_COLS = nsets * ACC_COLS;
  const uint32_t TMEM_COLS = 2 * BUF_COLS;                          // double buffered: 128 .. 512 columns
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NJ = CO_P + KD - 1, dpad = KD >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < CO_NA; ++i) { mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // weights -> shared memory once, in the K-major SWIZZLE_128B layout the B descriptor expects:
  // row n = (kh*3+kw)*Cout + co of tile (kd, chunk); 16-byte piece j of the row goes to n*128 + ((j ^ (n & 7)) << 4)
  {
    const int pieces = KD * nchunks * N_PAD * 8;
    for (int i = threadIdx.x; i < pieces; i += CO_THREADS) {
      const int j = i & 7, n = (i >> 3) % N_PAD, tc = (i >> 3) / N_PAD;     // tc = kd * nchunks + chunk
      const int chunk = tc % nchunks, kd = tc / nchunks;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (n < NT) {
        const int khw = n / COUT, co = n - khw * COUT;
        const int tap = kd * 9 + khw;
        v = *reinterpret_cast<const uint4*>(p.w + ((int64_t)tap * 16 + co) * p.w_ld + co_vchunk_w(p, chunk) + j * 8);
      }
      *reinterpret_cast<uint4*>(sW + (size_t)tc * N_PAD * 128 + n * 128 + ((j ^ (n & 7)) << 4)) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
  }
  if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapA)) : "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  auto coord = [&](int t, int& w0, int& h0, int& d0, int& b) {
    w0 = (t % p.tiles_w) * CO_BW; t /= p.tiles_w;
    h0 = (t % p.tiles_h) * CO_BH; t /= p.tiles_h;
    d0 = (t % p.groups_d) * CO_P; t /= p.groups_d;
    b = t;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one_sync()) {
      uint32_t seq = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        int w0, h0, d0, b;
        coord(t, w0, h0, d0, b);
        for (int c = 0; c < nchunks; ++c)
          for (int j = 0; j < NJ; ++j, ++seq) {
            const uint32_t slot = seq % CO_NA, ph = (seq / CO_NA) & 1;
            mbar_wait(&empty_a[slot], ph ^ 1);
            mbar_expect_tx(&full_a[slot], CO_PATCH_BYTES);
            asm volatile(
                "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                    smem_u32(sA + (size_t)slot * CO_PATCH_STRIDE)),
                "l"(reinterpret_cast<uint64_t>(&tmapA)), "r"(co_vchunk_a(p, c)), "r"(w0 - 1 + p.pad_hw), "r"(h0 - 1 + p.pad_hw), "r"(d0 + j - dpad + p.pad_d), "r"(b),
                "r"(smem_u32(&full_a[slot]))
                : "memory");
          }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc_h16(N_PAD, 128, p.f16);
    constexpr uint32_t HI = umma_desc_hi(1024);                     // 8 consecutive patch rows = 1024 B (both operands)
    uint32_t seq = 0, it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      mbar_wait(&acc_empty[as], aph ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem_base + as * BUF_COLS;
      uint32_t used[2] = {0u, 0u};                                  // per set: output planes already written in this tile
      int part = -1;
      for (int c = 0; c < nchunks; ++c) {
        if (++part == p.vparts) part = 0;                           // part of virtual chunk c (no division in the issue path)
        for (int j = 0; j < NJ; ++j, ++seq) {
          const uint32_t slot = seq % CO_NA;
          mbar_wait(&full_a[slot], (seq / CO_NA) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a0 = umma_desc_lo(smem_u32(sA + (size_t)slot * CO_PATCH_STRIDE));
          const uint32_t a1 = a0 + ((CO_G1 * 128) >> 4);
          if (elect_one_sync()) {
            // input plane j feeds output plane pp = j - kd with depth tap kd
            for (int kd = 0; kd < KD; ++kd) {
              const int pp = j - kd;
              if (pp < 0 || pp >= CO_P) continue;
              const uint32_t b_lo = umma_desc_lo(smem_u32(sW + (size_t)(kd * nchunks + c) * N_PAD * 128));
              const uint32_t set = part ? 1u : 0u;
              const uint32_t first = (used[set] >> pp) & 1u;         // 0: first MMA into this (set, plane) accumulator
              used[set] |= 1u << pp;
              const uint32_t accs = acc + set * ACC_COLS;
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                umma_bf16(accs + (pp * 2 + 0) * N_PAD, umma_desc64(a0 + k4 * 2, HI), umma_desc64(b_lo + k4 * 2, HI), idesc,
                          k4 == 0 ? first : 1u);
                umma_bf16(accs + (pp * 2 + 1) * N_PAD, umma_desc64(a1 + k4 * 2, HI), umma_desc64(b_lo + k4 * 2, HI), idesc,
                          k4 == 0 ? first : 1u);
              }
            }
            umma_commit(&empty_a[slot]);
          }
          __syncwarp();
        }
      }
      if (elect_one_sync()) umma_commit(&acc_full[as]);
      __syncwarp();
    }
  } else {
    // ===================== epilogue: TMEM -> Z in shared memory -> shifted gather -> global =====================
    const int q = warp & 3;                          // TMEM lane quadrant of this warp (warps 2,3,4,5 -> 2,3,0,1)
    const int r0 = q * 32 + lane;                    // row of row group 0; row group 1 holds patch row r0 + 52
    const int et = (warp - 2) * 32 + lane;           // 0..127: the output pixel this thread gathers
    const int oline = et >> 3, owp = et & 7;
    constexpr int ZS = NT;                           // row stride of Z (9 / 18 / 27 words)
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      int w0, h0, d0, b;
      coord(t, w0, h0, d0, b);
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      mbar_wait(&acc_full[as], aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int pp = 0; pp < CO_P; ++pp)
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const uint32_t taddr = tmem_base + as * BUF_COLS + (pp * 2 + g) * N_PAD + ((uint32_t)(q * 32) << 16);
          const int prow = g == 0 ? r0 : r0 + CO_G1;
          const bool mine = g == 0 || prow >= 128;   // rows 52..127 are computed by both groups: group 0 owns them
          float* zrow = sZ + ((size_t)pp * CO_ROWS + prow) * ZS;
          if constexpr (N_PAD == 16) {
            uint32_t v[16];
            DSK_TMEM_LD_X16(v, taddr);
            if (nsets > 1) {
              uint32_t u[16];
              DSK_TMEM_LD_X16(u, taddr + ACC_COLS);
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(fmaf(__uint_as_float(u[e]), 1.0f / 2048.0f, __uint_as_float(v[e])));
            }
            if (mine)
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (e < NT) zrow[e] = __uint_as_float(v[e]);
          } else {
            uint32_t v[32];
            DSK_TMEM_LD_X32(v, taddr);
            if (nsets > 1) {
              uint32_t u[32];
              DSK_TMEM_LD_X32(u, taddr + ACC_COLS);
#pragma unroll
              for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(fmaf(__uint_as_float(u[e]), 1.0f / 2048.0f, __uint_as_float(v[e])));
            }
            if (mine)
#pragma unroll
              for (int e = 0; e < 32; ++e)
                if (e < NT) zrow[e] = __uint_as_float(v[e]);
          }
        }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);    // accumulators are free again: the MMAs of the tile after next may start
      asm volatile("bar.sync 1, 128;" ::: "memory"); // Z complete (epilogue warps only)
      const int h = h0 + oline, w = w0 + owp;
      if (h < p.H && w < p.W) {
#pragma unroll
        for (int pp = 0; pp < CO_P; ++pp) {
          const int d = d0 + pp;
          if (d >= p.D) continue;
          float acc[COUT];
#pragma unroll
          for (int co = 0; co < COUT; ++co) acc[co] = 0.0f;
#pragma unroll
          for (int khw = 0; khw < 9; ++khw) {
            const float* z = sZ + ((size_t)pp * CO_ROWS + (oline + khw / 3) * CO_PW + owp + khw % 3) * ZS + khw * COUT;
#pragma unroll
            for (int co = 0; co < COUT; ++co) acc[co] += z[co];
          }
          const int64_t pix = (((int64_t)b * p.D + d) * p.H + h) * p.W + w;
#pragma unroll
          for (int co = 0; co < COUT; ++co) {
            const float x = acc[co] + (p.bias != nullptr ? __ldg(p.bias + co) : 0.0f);
            if (p.out_nchw != nullptr) {
              const int64_t S = (int64_t)p.planes_per_sample * p.H * p.W;
              const int64_t sample = ((int64_t)b * p.D + d) / p.planes_per_sample;
              p.out_nchw[(sample * COUT + co) * S + (pix - sample * S)] = x;
            } else {
              if (p.out_f32) reinterpret_cast<float*>(p.out)[pix * COUT + co] = x;
              else p.out[pix * COUT + co] = pack_h1(x, p.f16);
            }
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory"); // gather done before the next tile overwrites Z
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

template <int COUT, int CO_P = 2>
static int launch_convout(const CUtensorMap& ta, const CoParams& p, cudaStream_t st) {
  constexpr int N_PAD = 9 * COUT <= 16 ? 16 : 32;
  const int nchunks = (p.Cin / 64) * p.vparts;
  const size_t wbytes = ((size_t)p.KD * nchunks * N_PAD * 128 + 1023) & ~(size_t)1023;
  const size_t smem = (size_t)CO_NA * CO_PATCH_STRIDE + wbytes + (size_t)CO_P * CO_ROWS * 9 * p.Cout * sizeof(float) + 1024;
  if (smem > 227 * 1024) return DSK_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(convout_tc_kernel<COUT, CO_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("convout_tc: cudaFuncSetAttribute(%zu B smem): %s", smem, cudaGetErrorString(e)); return DSK_ERR_CUDA; }
  const int grid = p.total_tiles < DSK_NUM_SMS ? p.total_tiles : DSK_NUM_SMS;
  DSK_LAUNCH((convout_tc_kernel<COUT, CO_P>), grid, CO_THREADS, smem, st, ta, p);
  return DSK_OK;
}

// returns DSK_ERR_UNSUPPORTED for shapes it does not take (the caller falls back to the N = 16 tile of conv_tc.cu)
// padded != 0: `in` is the halo-padded copy [B, D+2 (3-D), H+2, W+2, Cin] of a circular convolution
int convout_tc_dispatch(const dsk_conv_desc* d, const void* in, const void* w, const float* bias, void* out, cudaStream_t st, int padded) {
  const bool a_split = d->in_dtype == DSK_SPLIT_F16, w_split = d->w_dtype == DSK_SPLIT_F16;
  if (d->ksize != 3 || d->up2 || !(is_h16(d->in_dtype) || a_split) || d->Cin % 64 != 0 || d->Cout > 3) return DSK_ERR_UNSUPPORTED;
  if (!d->out_nchw_f32 && !(is_h16(d->out_dtype) || d->out_dtype == DSK_F32)) return DSK_ERR_UNSUPPORTED;
  const int a_ld = d->Cin * (a_split ? 2 : 1);
  EncodeTiledFn encode = get_encode();
  if (encode == nullptr) return DSK_ERR_UNSUPPORTED;
  const int KD = d->ndim == 3 ? 3 : 1;
  const int planes = d->ndim == 3 ? d->D : d->B, batch = d->ndim == 3 ? d->B : 1;
  CUtensorMap ta;
  const int pad_hw = padded ? 1 : 0, pad_d = (padded && d->ndim == 3) ? 1 : 0;
  const int tW = d->W + 2 * pad_hw, tH = d->H + 2 * pad_hw, tP = planes + 2 * pad_d;
  cuuint64_t dims[5] = {(cuuint64_t)a_ld, (cuuint64_t)tW, (cuuint64_t)tH, (cuuint64_t)tP, (cuuint64_t)batch};
  cuuint64_t strides[4] = {(cuuint64_t)a_ld * 2, (cuuint64_t)tW * a_ld * 2, (cuuint64_t)tH * tW * a_ld * 2,
                           (cuuint64_t)tP * tH * tW * a_ld * 2};
  cuuint32_t box[5] = {64, CO_PW, CO_PH, 1, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(&ta, tmap_h16(d->in_dtype == DSK_BF16 ? DSK_BF16 : DSK_F16), 5, const_cast<void*>(in), dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("convout_tc: activation tensor map failed (CUresult %d)", (int)r); return DSK_ERR_CUDA; }
  CoParams p;
  p.B = batch; p.D = planes; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout; p.KD = KD;
  // planes per tile: 3-D single-channel outputs with plain operands take 8 or 4 planes per tile (depth halo 1.25 / 1.5 instead of
  // 2.0: the kernel is bound by the L2 -> SM patch stream) as long as the grid keeps >= 6 rounds of tiles; DSK_CONVOUT_P forces.
  static const int force_p = [] { const char* e = getenv("DSK_CONVOUT_P"); return e ? atoi(e) : 0; }();
  p.tiles_w = (d->W + CO_BW - 1) / CO_BW; p.tiles_h = (d->H + CO_BH - 1) / CO_BH;
  int P = 2;
  if (KD == 3 && d->Cout == 1 && !a_split) {
    for (int cand = 8; cand >= 4; cand >>= 1)
      if (planes % cand == 0 && (int64_t)p.tiles_w * p.tiles_h * (planes / cand) * batch >= 6 * DSK_NUM_SMS) { P = cand; break; }
    if (force_p == 2 || ((force_p == 4 || force_p == 8) && planes % force_p == 0)) P = force_p;
  }
  p.groups_d = (planes + P - 1) / P;
  p.total_tiles = p.tiles_w * p.tiles_h * p.groups_d * batch;
  p.w = (const uint16_t*)w; p.bias = bias;
  p.out = d->out_nchw_f32 ? nullptr : (uint16_t*)out;
  p.f16 = d->in_dtype == DSK_BF16 ? 0 : 1;
  p.out_f32 = (!d->out_nchw_f32 && d->out_dtype == DSK_F32) ? 1 : 0;
  p.vparts = a_split ? (w_split ? 3 : 2) : 1;
  p.a_lo_off = d->Cin; p.w_lo_off = d->Cin; p.w_ld = d->Cin * (w_split ? 2 : 1);
  p.out_nchw = d->out_nchw_f32 ? (float*)out : nullptr;
  p.planes_per_sample = d->ndim == 3 ? d->D : 1;
  p.pad_hw = pad_hw; p.pad_d = pad_d;
  if (d->Cout == 1 && P == 8) return launch_convout<1, 8>(ta, p, st);
  if (d->Cout == 1 && P == 4) return launch_convout<1, 4>(ta, p, st);
  if (d->Cout == 1) return launch_convout<1>(ta, p, st);
  if (d->Cout == 2) return launch_convout<2>(ta, p, st);
  return launch_convout<3>(ta, p, st);
}

}  // namespace dsk
