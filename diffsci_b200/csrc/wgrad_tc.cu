// wgrad_tc.cu -- K2 (bf16 throughput mode): convolution WEIGHT GRADIENT on the 5th-gen tensor cores
// (tcgen05.mma, accumulators in TMEM, operands by TMA).  sm_100a only.
//
// Reference: autograd of torch.nn.Conv2d/Conv3d(k=3, padding='same') reached by loss.backward() in
// KarrasModule.training_step (karras/karrasmodule.py:1146-1155; layers nets/commonlayers.py:777-833, 53-58, 123-128).
//
//   dW[tap][ci][co] = sum_pixels X[pix + tap][ci] * dY[pix][co]
//
// Formulation.  The reduction runs over output pixels, so both operands are used exactly as they lie in HBM
// (channels-last: a pixel is one 128-byte row per 64-channel chunk) as MN-MAJOR UMMA operands -- no transposition:
//   A (M x K) = X patch,  M = 2 taps x 64 input channels, K = 16 pixels per instruction
//   B (N x K) = dY tile,  N = 64 output channels
// A CTA tile is 16 (h) x 8 (w) output pixels of one plane.  One TMA box brings the halo'd (18 x 10) X patch, another
// the dY tile.  As in the forward kernel (conv_tc.cu) the 9 in-plane taps are shifted VIEWS of the one patch; here two
// taps are fused into one M = 128 instruction by pointing the descriptor's leading-dimension offset (the stride between
// the two 64-element MN atoms) at the second tap's view: LBO = 128 B (kw -> kw+1) or 1024 B ((kh,2) -> (kh+1,0)).
// 9 taps = 4 pairs + 1 padded pair -> 5 accumulators of 64 fp32 columns in TMEM, which stay resident for the CTA's whole
// pixel range: there is no epilogue inside the main loop.
// Work split: blockIdx.x = (depth tap kd, 64-chunk of Cin, 64-chunk of Cout), blockIdx.y = slice of the pixel tiles.
// Each CTA writes its partial [9 taps x 64 ci] x [64 co] block to ws[slice][tap*Cin + ci][co]; the slices are summed in a
// fixed order by wgrad_reduce_kernel (gemm_ffma.cu), which also emits the reference layout [Cout][Cin][taps].
//
// Pipeline: warp 0 = TMA producer (5-stage ring of {patch, dY tile}), warp 2 = MMA issuer / TMEM owner,
// warps 4-7 = final TMEM -> global drain.
#include "tc_common.cuh"

namespace dsk {

constexpr int WG_BW = 8, WG_BH = 16;
constexpr int WG_PW = WG_BW + 2, WG_PH = WG_BH + 2;
constexpr int WG_PATCH_BYTES = WG_PW * WG_PH * 128;   // 23040
constexpr int WG_PATCH_STRIDE = 23552;                // 23 KB, keeps the dY tile 1024-B aligned
constexpr int WG_DY_BYTES = WG_BW * WG_BH * 128;      // 16384
constexpr int WG_STAGE = WG_PATCH_STRIDE + WG_DY_BYTES;
constexpr int WG_STAGES = 5;
constexpr int WG_THREADS = 256;
constexpr int WG_NACC = 5;                            // tap pairs (0,1) (2,3) (4,5) (6,7) (8,-)
constexpr uint32_t WG_TMEM_COLS = 512;                // 5 x 64 rounded up to a power of two

struct WgParams {
  int B, D, H, W, Cin, Cout, KD;   // B, D: batch / planes as the tensor maps see them (2-D: B = 1, D = batch)
  int tiles_w, tiles_h, total_tiles, tiles_per_split;
  int n_ci, n_co;
  float* ws;                       // [nsplit][taps*Cin][Cout]
  int pad_hw, pad_d;               // circular padding: x tensor map covers the halo-padded copy (see conv_tc.cu)
};

// MN-major SWIZZLE_128B descriptors: lo = start >> 4 | (LBO >> 4) << 16 ; hi = SBO >> 4 | version | swizzle mode
__device__ __forceinline__ uint32_t wg_desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return (saddr >> 4) | ((lbo_bytes >> 4) << 16); }

// bf16 x bf16 -> fp32, A and B MN-major (bits 15, 16), M = 128
__device__ __forceinline__ uint32_t wg_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapDY, const WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[WG_STAGES], empty[WG_STAGES], acc_full;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int unit = blockIdx.x;
  const int kd = unit % p.KD; unit /= p.KD;
  const int ci_chunk = unit % p.n_ci;
  const int co_chunk = unit / p.n_ci;
  const int split = blockIdx.y;
  const int t0 = split * p.tiles_per_split;
  const int t1 = min(p.total_tiles, t0 + p.tiles_per_split);
  const int dpad = p.KD >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(WG_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapX)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapDY)) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one_sync()) {
      uint32_t seq = 0;
      for (int t = t0; t < t1; ++t, ++seq) {
        int r = t;
        const int w0 = (r % p.tiles_w) * WG_BW; r /= p.tiles_w;
        const int h0 = (r % p.tiles_h) * WG_BH; r /= p.tiles_h;
        const int d = r % p.D;
        const int b = r / p.D;
        const uint32_t slot = seq % WG_STAGES, ph = (seq / WG_STAGES) & 1;
        mbar_wait(&empty[slot], ph ^ 1);
        mbar_expect_tx(&full[slot], WG_PATCH_BYTES + WG_DY_BYTES);
        uint8_t* st = smem + (size_t)slot * WG_STAGE;
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                smem_u32(st)),
            "l"(reinterpret_cast<uint64_t>(&tmapX)), "r"(ci_chunk * 64), "r"(w0 - 1 + p.pad_hw), "r"(h0 - 1 + p.pad_hw), "r"(d + kd - dpad + p.pad_d), "r"(b),
            "r"(smem_u32(&full[slot]))
            : "memory");
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                smem_u32(st + WG_PATCH_STRIDE)),
            "l"(reinterpret_cast<uint64_t>(&tmapDY)), "r"(co_chunk * 64), "r"(w0), "r"(h0), "r"(d), "r"(b), "r"(smem_u32(&full[slot]))
            : "memory");
      }
    }
  } else if (warp == 2) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = wg_idesc(64);
    constexpr uint32_t A_HI = umma_desc_hi(WG_PW * 128), B_HI = umma_desc_hi(1024);
    uint32_t seq = 0;
    for (int t = t0; t < t1; ++t, ++seq) {
      const uint32_t slot = seq % WG_STAGES;
      mbar_wait(&full[slot], (seq / WG_STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sx = smem_u32(smem + (size_t)slot * WG_STAGE), sy = sx + WG_PATCH_STRIDE;
      if (elect_one_sync()) {
#pragma unroll
        for (int ks = 0; ks < WG_BH / 2; ++ks) {                 // 16 pixels = 2 tile rows per instruction
          const uint64_t db = umma_desc64(wg_desc_lo(sy + ks * 2048, 16), B_HI);
          const uint32_t acc = (seq | (uint32_t)ks) == 0 ? 0u : 1u;
#pragma unroll
          for (int a = 0; a < WG_NACC; ++a) {
            // first tap of the pair: t = 2a -> (kh, kw) = (t / 3, t % 3); second tap t + 1 (a == 4: the tap itself again)
            const int ta = 2 * a, kh = ta / 3, kw = ta % 3;
            // (the padded pair's second half reads one pixel further along the row: in-bounds garbage, rows 64..127 of
            //  accumulator 4 are never stored)
            const uint32_t lbo = kw == 2 && a != 4 ? (uint32_t)(WG_PW - 2) * 128u : 128u;
            const uint32_t start = sx + (uint32_t)(((2 * ks + kh) * WG_PW + kw) * 128);
            umma_bf16(tmem_base + a * 64, umma_desc64(wg_desc_lo(start, lbo), A_HI), db, idesc, acc);
          }
        }
        umma_commit(&empty[slot]);
      }
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit(&acc_full);
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== drain: TMEM -> ws[split][tap*Cin + ci][co] =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;             // accumulator row: (second tap of the pair ? 64 : 0) + ci
    const int half = row >> 6, ci = row & 63;
    const int taps = p.KD * 9;
    const int64_t MN = (int64_t)taps * p.Cin * p.Cout;
    float* wsz = p.ws + (int64_t)split * MN;
    if (t1 > t0) {
      mbar_wait(&acc_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
#pragma unroll
    for (int a = 0; a < WG_NACC; ++a) {
      const int tap_hw = 2 * a + half;
      const bool live = tap_hw < 9;
      float* dst = wsz + ((int64_t)(kd * 9 + (live ? tap_hw : 0)) * p.Cin + ci_chunk * 64 + ci) * p.Cout + co_chunk * 64;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        if (t1 > t0) {
          const uint32_t taddr = tmem_base + a * 64 + c0 + ((uint32_t)(q * 32) << 16);
          DSK_TMEM_LD_X32(v, taddr);
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = 0u;
        }
        if (live) {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<float4*>(dst + c0 + e) =
                make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(WG_TMEM_COLS));
}

// dw[co][ci][tap] (+)= sum_z ws[z][tap*Cin + ci][co]   (gemm_ffma.cu)
int wgrad_reduce_launch(const float* ws, float* dw, int Cout, int Cin, int taps, int nsplit, int accumulate, cudaStream_t st);

struct WgPlan {
  bool ok;
  int KD, planes, batch, tiles_w, tiles_h, total_tiles, units, nsplit, tiles_per_split;
};

static WgPlan wg_plan(const dsk_conv_desc* d) {
  WgPlan w{};
  w.ok = d->ksize == 3 && !d->up2 && d->in_dtype == DSK_BF16 && d->out_dtype == DSK_BF16 && d->Cin % 64 == 0 && d->Cout % 64 == 0 &&
         ((d->ndim == 2 && d->D == 1) || d->ndim == 3);
  if (!w.ok) return w;
  w.KD = d->ndim == 3 ? 3 : 1;
  w.planes = d->ndim == 3 ? d->D : d->B;
  w.batch = d->ndim == 3 ? d->B : 1;
  w.tiles_w = (d->W + WG_BW - 1) / WG_BW;
  w.tiles_h = (d->H + WG_BH - 1) / WG_BH;
  const int64_t total = (int64_t)w.tiles_w * w.tiles_h * w.planes * w.batch;
  if (total > 0x7fffffff) { w.ok = false; return w; }
  w.total_tiles = (int)total;
  w.units = w.KD * (d->Cin / 64) * (d->Cout / 64);
  int ns = DSK_NUM_SMS / w.units;
  if (ns < 1) ns = 1;
  if (ns > w.total_tiles) ns = w.total_tiles;
  w.tiles_per_split = (w.total_tiles + ns - 1) / ns;
  w.nsplit = (w.total_tiles + w.tiles_per_split - 1) / w.tiles_per_split;
  return w;
}

static int64_t wg_split_bytes(const dsk_conv_desc* d, const WgPlan& w) {
  const int64_t b = (int64_t)w.nsplit * (w.KD * 9) * d->Cin * d->Cout * (int64_t)sizeof(float);
  return (b + 1023) & ~(int64_t)1023;
}
// circular padding: the halo-padded copy of x lives behind the split-K slices in the same workspace
static int64_t wg_pad_bytes(const dsk_conv_desc* d) {
  if (!d->circular) return 0;
  const int pd = d->ndim == 3 ? 1 : 0;
  return (int64_t)d->B * (d->D + 2 * pd) * (d->H + 2) * (d->W + 2) * d->Cin * 2;
}

}  // namespace dsk

using namespace dsk;

extern "C" int64_t dsk_conv_wgrad_tc_ws_bytes(const dsk_conv_desc* d) {
  if (!d || d->w_dtype != DSK_BF16) return 0;
  const WgPlan w = wg_plan(d);
  if (!w.ok) return 0;
  return wg_split_bytes(d, w) + wg_pad_bytes(d);
}

extern "C" int dsk_conv_wgrad_tc(const dsk_conv_desc* d, const void* x, const void* dy, float* dw, void* ws, int accumulate,
                                 void* stream) {
  const WgPlan w = wg_plan(d);
  if (!w.ok) return DSK_ERR_UNSUPPORTED;
  EncodeTiledFn encode = get_encode();
  DSK_REQUIRE(encode != nullptr, "dsk_conv_wgrad(tc): cuTensorMapEncodeTiled is unavailable");
  CUtensorMap tx, ty;
  const int pad_hw = d->circular ? 1 : 0, pad_d = (d->circular && d->ndim == 3) ? 1 : 0;
  if (d->circular) {
    void* xp = (uint8_t*)ws + wg_split_bytes(d, w);
    const int rc = pad_circular_launch(x, xp, d->B, d->D, d->H, d->W, d->Cin, d->ndim, DSK_BF16, as_stream(stream));
    if (rc != DSK_OK) return rc;
    x = xp;
  }
  {
    const int tW = d->W + 2 * pad_hw, tH = d->H + 2 * pad_hw, tP = w.planes + 2 * pad_d;
    cuuint64_t dims[5] = {(cuuint64_t)d->Cin, (cuuint64_t)tW, (cuuint64_t)tH, (cuuint64_t)tP, (cuuint64_t)w.batch};
    cuuint64_t strides[4] = {(cuuint64_t)d->Cin * 2, (cuuint64_t)tW * d->Cin * 2, (cuuint64_t)tH * tW * d->Cin * 2,
                             (cuuint64_t)tP * tH * tW * d->Cin * 2};
    cuuint32_t box[5] = {64, WG_PW, WG_PH, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DSK_REQUIRE(r == CUDA_SUCCESS, "dsk_conv_wgrad(tc): x tensor map failed (CUresult %d)", (int)r);
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)d->Cout, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)w.planes, (cuuint64_t)w.batch};
    cuuint64_t strides[4] = {(cuuint64_t)d->Cout * 2, (cuuint64_t)d->W * d->Cout * 2, (cuuint64_t)d->H * d->W * d->Cout * 2,
                             (cuuint64_t)w.planes * d->H * d->W * d->Cout * 2};
    cuuint32_t box[5] = {64, WG_BW, WG_BH, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&ty, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DSK_REQUIRE(r == CUDA_SUCCESS, "dsk_conv_wgrad(tc): dy tensor map failed (CUresult %d)", (int)r);
  }
  WgParams p;
  p.B = w.batch; p.D = w.planes; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout; p.KD = w.KD;
  p.tiles_w = w.tiles_w; p.tiles_h = w.tiles_h; p.total_tiles = w.total_tiles; p.tiles_per_split = w.tiles_per_split;
  p.n_ci = d->Cin / 64; p.n_co = d->Cout / 64;
  p.ws = (float*)ws;
  p.pad_hw = pad_hw; p.pad_d = pad_d;
  const size_t smem = (size_t)WG_STAGES * WG_STAGE + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("wgrad_tc: cudaFuncSetAttribute(%zu B smem): %s", smem, cudaGetErrorString(e)); return DSK_ERR_CUDA; }
    configured = true;
  }
  cudaStream_t st = as_stream(stream);
  dim3 grid(w.units, w.nsplit);
  DSK_LAUNCH(wgrad_tc_kernel, grid, WG_THREADS, smem, st, tx, ty, p);
  return wgrad_reduce_launch((const float*)ws, dw, d->Cout, d->Cin, w.KD * 9, w.nsplit, accumulate, st);
}
