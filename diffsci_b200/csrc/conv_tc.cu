// conv_tc.cu -- K1 (bf16 throughput mode): implicit-GEMM 3x3(x3) convolution on the 5th-gen tensor
// cores (tcgen05.mma, accumulators in TMEM) fed by TMA.  sm_100a only.
//
// Reference call sites: torch.nn.Conv2d/Conv3d(k=3, padding='same') in ResnetBlockC
// (nets/commonlayers.py:777-833), DownSampler/UpSampler (commonlayers.py:53-58,123-128).
//
// Formulation.  Activations are channels-last bf16 [B, D, H, W, C]; a pixel is one 128-byte row per
// 64-channel chunk.  A CTA tile is 16 (h) x 8 (w) output pixels (= the 128 rows of one UMMA, M = 128)
// on P consecutive d-planes, all N_TILE output channels; accumulators: P x N_TILE fp32 TMEM columns.
// For every input plane the tile touches, ONE TMA box load brings the (16+2) x (8+2) halo'd patch of a
// 64-channel chunk into shared memory (zero padding = TMA out-of-bounds fill, signed coordinates).
// The 9 in-plane taps are then just 9 shifted *views* of that patch: the UMMA shared-memory descriptor
// (K-major, SWIZZLE_128B) takes start = patch + (kh*10 + kw)*128 B and SBO = 10*128 B -- the hardware
// swizzle is a function of the absolute smem address (verified by tests/cuda/umma_probe.cu), so views
// that are not 1024-B aligned read exactly "pixel rows r .. r+7" of each 8-row group.  No im2col matrix
// is ever materialised, in HBM or in shared memory: each activation byte crosses L2->SM about 1.4x
// (halo) x (P+2)/P (depth halo) instead of 27x.
// Weights are bf16 [tap][Cout][Cin] (K-major B operand), streamed per tap through a TMA ring and shared
// by the P planes of the tile.
//
// Pipeline (persistent CTAs, 1 per SM, static tile schedule):
//   warp 0   TMA producer, activations (patch ring, per-patch full/empty mbarriers)
//   warp 1   TMA producer, weights     (tap ring)
//   warp 2   MMA issuer (one elected thread), owns the TMEM allocation
//   warps 4-7 epilogue: tcgen05.ld -> +bias +time-embedding vector +residual -> bf16 -> global
// TMEM accumulators are double buffered so the epilogue of tile i overlaps the main loop of tile i+1.
#include <stdlib.h>

#include "tc_common.cuh"

namespace dsk {

constexpr int TC_BW = 8, TC_BH = 16;                 // output tile (w, h) = 128 rows
constexpr int TC_PW = TC_BW + 2, TC_PH = TC_BH + 2;  // halo'd patch
constexpr int TC_PATCH_BYTES = TC_PW * TC_PH * 128;  // 23040
constexpr int TC_PATCH_STRIDE = 23552;               // rounded up to 1024
constexpr int TC_THREADS = 256;

struct TcParams {
  int B, D, H, W, Cin, Cout;
  int KD;                 // 3 for 3-D convs, 1 for 2-D (then D is the batch-of-planes axis)
  int tiles_w, tiles_h, groups_d, n_tiles, total_tiles;
  const float* bias;      // [Cout] or null
  const float* chan_bias; // [B][Cout] or null
  const uint16_t* residual;  // like out, or null (16-bit format `f16`, or fp32 when out_f32)
  uint16_t* out;          // [B, D, H, W, Cout]
  int f16;                // 16-bit format of the operands and of 16-bit outputs: 0 = bfloat16, 1 = IEEE half
  int out_f32;            // out / residual are fp32 (split parity mode)
  // Split operands (DSK_SPLIT_F16: rows of [hi | lo]): every real 64-channel K chunk becomes `vparts` virtual chunks that
  // accumulate into the same TMEM columns -- vparts = 3: (A hi, W hi), (A hi, W lo), (A lo, W hi); vparts = 2: the weights are
  // plain fp16, (A hi, W), (A lo, W); vparts = 1: plain 16-bit operands.  a_lo_off / w_lo_off: channel offset of the lo half.
  int vparts, a_lo_off, w_lo_off;
  // Accumulator sets (split mode only, nsets = 4; N_TILE = 64): tcgen05 adds into its fp32 accumulators with truncation
  // (round toward zero; tools/probe_tc_accum.py: -0.7 * 2^-23 relative per chained MMA for same-sign data), so a long chain of
  // MMAs into ONE accumulator loses what the split operands gained.  The hi*hi products of a tile are therefore dealt
  // round-robin over three accumulator sets, the hi*lo + lo*hi products -- whose lo operands are stored scaled by 2^11
  // (DSK_SPLIT_F16), out of fp16's subnormal range -- go to a fourth, and the epilogue adds the four in fp32 (round to
  // nearest):  acc = (s0 + s1 + s2) + 2^-11 * s3.  The four sets occupy all 512 TMEM columns: no double buffering.
  int nsets;
  int dbg;                // DSK_CONV_DBG (measurements only): 1 = epilogue frees the accumulators without reading / storing them
  int planes_per_sample;  // D for 3-D; 1 for 2-D  (chan_bias row = plane / planes_per_sample)
  int cout_real;          // N_TILE = 16 path (convout): the first cout_real (<= 16) channels are real, the rest zero padding
  float* out_nchw;        // N_TILE = 16 path: fp32 NC(D)HW output (user layout) instead of channels-last bf16
  int nphase;             // sub-pixel UpSampler conv: 8 (3-D) / 4 (2-D) output parities per input-resolution tile, else 1
  float2* stats;          // cta_group::2 kernel: per-(sample, slot, channel) partial (sum, sum of squares) of the output, or null
  int samples;            // number of samples the stats are kept for (3-D: B; 2-D: the planes ARE the samples)
  int pad_hw, pad_d;      // circular padding: the activation tensor map covers a halo-padded copy ([.., D+2*pad_d, H+2, W+2, C]);
                          // these offsets move the patch coordinates into it (0: zero 'same' padding through TMA OOB fill)
  int pair_d;             // cta_group::2 kernel: the two CTAs of a pair take neighbouring plane groups instead of neighbouring
                          // w-tiles (odd number of w-tiles, e.g. the 7x7 planes of MNIST's bottom level); no fused statistics
  int res_f32;            // residual is fp32 (== out_f32 unless the output is a 16-bit operand copy of an fp32-storage mode)
  // tile index -> coordinates without integer division (cta_group::2 kernel): divisors nphase, pairs along the fastest tile
  // axis (w-pairs, or tiles_w when pairing along d), tiles_h, plane groups (or pairs of them), B
  uint32_t dv_m[5], dv_s[5];
  int bh;                 // cta_group::2 kernel: output rows per CTA tile = 16 * T (T = 2: two 16-row sub-tiles share every weight load)
};
// n / d for n < 2^31, d < 2^31 as a multiply + shift (Granlund-Montgomery, N = 31): m = ceil(2^(31+l) / d), l = ceil(log2 d)
struct FastDivHost {
  uint32_t m, s;
  explicit FastDivHost(uint32_t d) {
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;
    s = 31 + l;
    m = (uint32_t)(((1ull << s) + d - 1) / d);      // d == 1: 2^31; d > 2^(l-1) keeps m < 2^32
  }
};
__device__ __forceinline__ uint32_t fast_div(uint32_t n, uint32_t m, uint32_t s) { return (uint32_t)(((uint64_t)n * m) >> s); }
constexpr int TC_STAT_SLOTS = DSK_NUM_SMS;          // one slot per CTA
constexpr float kLoScale = 1.0f / 2048.0f;          // 2^-11: the lo half of a split operand is stored times 2^11
// Virtual K chunks in issue order: (real 64-channel chunk c, part) with part fastest.  A counter pair instead of vc / vparts:
// an integer division per tap in the MMA issue path costs more than the tap's MMAs take to execute.
struct VChunk {
  int c = 0, part = 0;
  __device__ __forceinline__ void next(const TcParams& p) { if (++part == p.vparts) { part = 0; ++c; } }
  // channel coordinate in the activation / weight tensor map
  __device__ __forceinline__ int a(const TcParams& p) const { return c * 64 + ((p.vparts > 1 && part == p.vparts - 1) ? p.a_lo_off : 0); }
  __device__ __forceinline__ int w(const TcParams& p) const { return c * 64 + ((p.vparts == 3 && part == 1) ? p.w_lo_off : 0); }
};

struct TileCoord {
  int w0, h0, d0, b, n0;
  int pa, pb, pc;   // output parity (depth, height, width) of a sub-pixel phase; 0 otherwise
  int phase;
};
__device__ __forceinline__ TileCoord tile_coord(const TcParams& p, int t, int n_tile_size, int P) {
  TileCoord c;
  c.phase = t % p.nphase; t /= p.nphase;            // phases of one tile run back to back (same patches: L2 hits)
  c.pc = c.phase & 1; c.pb = (c.phase >> 1) & 1; c.pa = (c.phase >> 2) & 1;
  c.w0 = (t % p.tiles_w) * TC_BW; t /= p.tiles_w;
  c.h0 = (t % p.tiles_h) * TC_BH; t /= p.tiles_h;
  c.d0 = (t % p.groups_d) * P;    t /= p.groups_d;
  c.b = t % p.B;                  t /= p.B;
  c.n0 = t * n_tile_size;
  return c;
}

// N_TILE: output channels per CTA tile (64 | 128); P: planes per tile; NA: patch ring depth; NB: weight ring depth
// UPS: conv(nearest_upsample_x2(x)) evaluated at INPUT resolution: for each output parity (a,b,c) the 3 taps of an axis
// collapse onto 2 input offsets {a-1, a} (weights pre-summed by pack_upconv_weight_kernel), i.e. 8 phase-specific
// 2x2x2 convolutions -- 64 tap-GEMMs per 8 outputs instead of 27 per output (3.4x fewer MACs), no 8x tensor in HBM.
template <int N_TILE, int P, int NA, int NB, bool UPS>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapW, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                   // NA patches
  uint8_t* sB = smem + (size_t)NA * TC_PATCH_STRIDE;    // NB weight tiles of N_TILE x 128 B
  constexpr int B_BYTES = N_TILE * 128;
  __shared__ uint64_t full_a[NA], empty_a[NA], full_b[NB], empty_b[NB], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  constexpr uint32_t TMEM_COLS = 2 * P * N_TILE;        // double-buffered accumulators
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS >= 32, "TMEM columns");
  // accumulator sets per tile: 1 (plain), 2 (hh | lo; activations-split mode) or 4 (three hh sets + lo; fp32-parity mode,
  // N_TILE = 64 only).  Double-buffered across tiles when two buffers fit the 512 TMEM columns (1 set; 2 sets at N_TILE = 64),
  // else single-buffered (the epilogue of a tile then runs between its MMAs and the next tile's).
  const int nsets = p.nsets;
  constexpr uint32_t SET_COLS = P * N_TILE;             // columns of one accumulator set
  const uint32_t buf_cols = nsets * SET_COLS, nbuf = 2 * buf_cols <= 512 ? 2u : 1u, tmem_cols = nbuf * buf_cols;
  const uint32_t lo_set = nsets - 1, nhh = nsets > 1 ? nsets - 1 : 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KD = p.KD, NJ = P + KD - 1;                 // patches per (tile, chunk)
  const int nchunks = (p.Cin / 64) * p.vparts;          // virtual K chunks (split operands: 2 or 3 per real chunk)
  const int dpad = KD >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapW)) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================== TMA producer: activation patches =====================
    if (elect_one_sync()) {
      uint32_t seq = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord tc = tile_coord(p, t, N_TILE, P);
        VChunk vc;
        for (int c = 0; c < nchunks; ++c, vc.next(p))
          for (int j = 0; j < NJ; ++j, ++seq) {
            const uint32_t slot = seq % NA, ph = (seq / NA) & 1;
            mbar_wait(&empty_a[slot], ph ^ 1);
            mbar_expect_tx(&full_a[slot], TC_PATCH_BYTES);
            asm volatile(
                "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                    smem_u32(sA + (size_t)slot * TC_PATCH_STRIDE)),
                "l"(reinterpret_cast<uint64_t>(&tmapA)), "r"(vc.a(p)), "r"(tc.w0 - 1 + p.pad_hw), "r"(tc.h0 - 1 + p.pad_hw), "r"(tc.d0 + j - dpad + p.pad_d), "r"(tc.b),
                "r"(smem_u32(&full_a[slot]))
                : "memory");
          }
      }
    }
  } else if (warp == 1) {
    // ===================== TMA producer: weight taps =====================
    if (elect_one_sync()) {
      uint32_t seq = 0;
      const int ntaps = UPS ? (KD == 3 ? 8 : 4) : KD * 9;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord tc = tile_coord(p, t, N_TILE, P);
        const int tap_base = UPS ? tc.phase * ntaps : 0;
        VChunk vc;
        for (int c = 0; c < nchunks; ++c, vc.next(p))
          for (int tap = tap_base; tap < tap_base + ntaps; ++tap, ++seq) {
            const uint32_t slot = seq % NB, ph = (seq / NB) & 1;
            mbar_wait(&empty_b[slot], ph ^ 1);
            mbar_expect_tx(&full_b[slot], B_BYTES);
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                    smem_u32(sB + (size_t)slot * B_BYTES)),
                "l"(reinterpret_cast<uint64_t>(&tmapW)), "r"(vc.w(p)), "r"(tap * p.Cout + tc.n0), "r"(smem_u32(&full_b[slot]))
                : "memory");
          }
      }
    }
  } else if (warp == 2) {
    // ===================== MMA issuer =====================
    // The whole warp walks the control flow (barrier waits are warp-uniform); one elected lane issues.
    const uint32_t idesc = umma_idesc_h16(N_TILE, 128, p.f16);
    constexpr uint32_t A_HI = umma_desc_hi(TC_PW * 128), B_HI = umma_desc_hi(1024);
    uint32_t seq_a = 0, seq_b = 0, it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const uint32_t as = nbuf == 1 ? 0u : (it & 1), aph = nbuf == 1 ? (it & 1) : ((it >> 1) & 1);
      mbar_wait(&acc_empty[as], aph ^ 1);               // epilogue has drained this accumulator set
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_acc = tmem_base + as * buf_cols;
      uint32_t used = 0, hh = 0;                         // accumulator sets written in this tile; round-robin counter
      int part = 0;                                      // part of the current virtual chunk (0: hi * hi)
      const TileCoord tc = tile_coord(p, t, N_TILE, P);
      for (int c = 0; c < nchunks; ++c) {
        const uint32_t seq_c = seq_a;                    // first patch of this (tile, chunk)
        if constexpr (UPS) {
          // all patches of the chunk up front; taps kd in {a, a+1}, kh in {b, b+1}, kw in {c, c+1}
          for (int j = 0; j < NJ; ++j) mbar_wait(&full_a[(seq_c + j) % NA], ((seq_c + j) / NA) & 1);
          const int ntd = KD == 3 ? 2 : 1;
          for (int td = 0; td < ntd; ++td) {
            const int kd = KD == 3 ? tc.pa + td : 0;
            uint32_t a_lo[P];
#pragma unroll
            for (int pp = 0; pp < P; ++pp)
              a_lo[pp] = umma_desc_lo(smem_u32(sA + (size_t)((seq_c + pp + kd) % NA) * TC_PATCH_STRIDE));
            for (int thw = 0; thw < 4; ++thw, ++seq_b) {
              const int kh = tc.pb + (thw >> 1), kw = tc.pc + (thw & 1);
              const uint32_t bs = seq_b % NB;
              mbar_wait(&full_b[bs], (seq_b / NB) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t b_lo = umma_desc_lo(smem_u32(sB + (size_t)bs * B_BYTES));
              const uint32_t tap_off = (kh * TC_PW + kw) * 8;
              uint32_t set = 0;                                 // accumulator set of this tap's MMAs
              if (nsets > 1) { if (part) set = lo_set; else { set = hh; hh = hh + 1 == nhh ? 0 : hh + 1; } }
              const uint32_t first = (used >> set) & 1u;
              used |= 1u << set;
              const uint32_t acc_set = tmem_acc + set * SET_COLS;
              if (elect_one_sync()) {
#pragma unroll
                for (int pp = 0; pp < P; ++pp) {
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4)
                    umma_bf16(acc_set + pp * N_TILE, umma_desc64(a_lo[pp] + tap_off + k4 * 2, A_HI),
                              umma_desc64(b_lo + k4 * 2, B_HI), idesc, k4 == 0 ? first : 1u);
                }
                umma_commit(&empty_b[bs]);
              }
              __syncwarp();
            }
          }
          if (elect_one_sync()) {
            for (int j = 0; j < NJ; ++j) umma_commit(&empty_a[(seq_c + j) % NA]);
          }
          __syncwarp();
        } else {
        for (int kd = 0; kd < KD; ++kd) {
          // patches first needed at this kd: j = 0..P-1 at kd == 0, else j = P-1+kd
          const int jlo = kd == 0 ? 0 : P - 1 + kd, jhi = P - 1 + kd;
          for (int j = jlo; j <= jhi; ++j) {
            const uint32_t s = seq_c + j;
            mbar_wait(&full_a[s % NA], (s / NA) & 1);
          }
          uint32_t a_lo[P];
#pragma unroll
          for (int pp = 0; pp < P; ++pp)
            a_lo[pp] = umma_desc_lo(smem_u32(sA + (size_t)((seq_c + pp + kd) % NA) * TC_PATCH_STRIDE));
#pragma unroll
          for (int khw = 0; khw < 9; ++khw, ++seq_b) {
            const uint32_t bs = seq_b % NB;
            mbar_wait(&full_b[bs], (seq_b / NB) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t b_lo = umma_desc_lo(smem_u32(sB + (size_t)bs * B_BYTES));
            const uint32_t tap_off = ((khw / 3) * TC_PW + (khw % 3)) * 8;      // (kh*10 + kw) * 128 B >> 4
            uint32_t set = 0;                                 // accumulator set of this tap's MMAs
            if (nsets > 1) { if (part) set = lo_set; else { set = hh; hh = hh + 1 == nhh ? 0 : hh + 1; } }
            const uint32_t first = (used >> set) & 1u;
            used |= 1u << set;
            const uint32_t acc_set = tmem_acc + set * SET_COLS;
            if (elect_one_sync()) {
#pragma unroll
              for (int pp = 0; pp < P; ++pp) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                  umma_bf16(acc_set + pp * N_TILE, umma_desc64(a_lo[pp] + tap_off + k4 * 2, A_HI),
                            umma_desc64(b_lo + k4 * 2, B_HI), idesc, k4 == 0 ? first : 1u);
              }
              umma_commit(&empty_b[bs]);                 // weight slot free once these MMAs retire
            }
            __syncwarp();
          }
          // patches whose last use was this kd
          const int rlo = kd, rhi = (kd == KD - 1) ? NJ - 1 : kd;
          if (elect_one_sync()) {
            for (int j = rlo; j <= rhi; ++j) umma_commit(&empty_a[(seq_c + j) % NA]);
          }
          __syncwarp();
        }
        }
        seq_a += NJ;
        if (++part == p.vparts) part = 0;
      }
      if (elect_one_sync()) umma_commit(&acc_full[as]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;                               // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;                        // accumulator row = pixel of the 16x8 tile
    const int line = row >> 3, wp = row & 7;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const TileCoord tc = tile_coord(p, t, N_TILE, P);
      const uint32_t as = nbuf == 1 ? 0u : (it & 1), aph = nbuf == 1 ? (it & 1) : ((it >> 1) & 1);
      mbar_wait(&acc_full[as], aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int h = tc.h0 + line, w = tc.w0 + wp;
      const bool in_hw = h < p.H && w < p.W;
#pragma unroll
      for (int pp = 0; pp < P; ++pp) {
        const int d = tc.d0 + pp;
        const bool valid = in_hw && d < p.D;
        int64_t pix;
        if constexpr (UPS) {   // output lives on the 2x grid: (2d+a, 2h+b, 2w+c); 2-D convs do not upsample the plane axis
          const int od = p.KD == 3 ? 2 * d + tc.pa : d, OD = p.KD == 3 ? 2 * p.D : p.D;
          pix = (((int64_t)tc.b * OD + od) * (2 * p.H) + (2 * h + tc.pb)) * (2 * p.W) + (2 * w + tc.pc);
        } else {
          pix = (((int64_t)tc.b * p.D + d) * p.H + h) * p.W + w;
        }
        const int brow = (tc.b * p.D + d) / p.planes_per_sample;
        const uint32_t taddr = tmem_base + as * buf_cols + pp * N_TILE + ((uint32_t)(q * 32) << 16);
        if constexpr (N_TILE == 16) {
          // few-output-channel conv (convout, reference punetg.py:209-214): only cout_real columns are stored
          uint32_t v[16];
          DSK_TMEM_LD_X16(v, taddr);
          if (nsets > 1) {                                  // split mode: (s0 + s1 + s2) + 2^-11 * s3
            uint32_t u[16];
#pragma unroll 1
            for (int sidx = 1; sidx < nsets; ++sidx) {
              DSK_TMEM_LD_X16(u, taddr + sidx * SET_COLS);
              const float sc = sidx == (int)lo_set ? kLoScale : 1.0f;
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(fmaf(__uint_as_float(u[e]), sc, __uint_as_float(v[e])));
            }
          }
          if (valid) {
            const int64_t S = (int64_t)p.planes_per_sample * p.H * p.W;
#pragma unroll
            for (int co = 0; co < 16; ++co) {
              if (co < p.cout_real) {
                float x = __uint_as_float(v[co]);
                if (p.bias != nullptr) x += __ldg(p.bias + co);
                if (p.out_nchw != nullptr) p.out_nchw[((int64_t)brow * p.cout_real + co) * S + (pix - (int64_t)brow * S)] = x;
                else if (p.out_f32) reinterpret_cast<float*>(p.out)[pix * p.cout_real + co] = x;
                else p.out[pix * p.cout_real + co] = pack_h1(x, p.f16);
              }
            }
          }
        } else
#pragma unroll
        for (int c0 = 0; c0 < N_TILE; c0 += 32) {
          uint32_t v[32];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
              : "r"(taddr + c0));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (nsets > 1) {                                  // split mode: (s0 + s1 + s2) + 2^-11 * s3
            uint32_t u[32];
#pragma unroll 1
            for (int sidx = 1; sidx < nsets; ++sidx) {
              DSK_TMEM_LD_X32(u, taddr + sidx * SET_COLS + c0);
              const float sc = sidx == (int)lo_set ? kLoScale : 1.0f;
#pragma unroll
              for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(fmaf(__uint_as_float(u[e]), sc, __uint_as_float(v[e])));
            }
          }
          if (valid) {
            const int n = tc.n0 + c0;
            const int64_t off = pix * p.Cout + n;
#pragma unroll
            for (int g = 0; g < 4; ++g) {                  // 4 x 8 channels
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                // acc + (bias + chan_bias): the order of the pair kernel's staged bias sum -- which of the two kernels runs depends
                // on the batch (parity of the plane-group count), and sharded sampling must reproduce the single-GPU run bit for bit
                float bsum = p.bias != nullptr ? __ldg(p.bias + n + g * 8 + e) : 0.0f;
                if (p.chan_bias != nullptr) bsum += __ldg(p.chan_bias + (int64_t)brow * p.Cout + n + g * 8 + e);
                f[e] = __uint_as_float(v[g * 8 + e]) + bsum;
              }
              if (p.residual != nullptr) {                 // warp-uniform; the residual's format is independent of the output's
                if (p.res_f32) {
                  const float4* rf = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + off + g * 8);
                  const float4 r0 = rf[0], r1 = rf[1];
                  f[0] += r0.x; f[1] += r0.y; f[2] += r0.z; f[3] += r0.w; f[4] += r1.x; f[5] += r1.y; f[6] += r1.z; f[7] += r1.w;
                } else {
                  const uint4 rr = *reinterpret_cast<const uint4*>(p.residual + off + g * 8);
                  const uint32_t* rw = reinterpret_cast<const uint32_t*>(&rr);
#pragma unroll
                  for (int e = 0; e < 4; ++e) { const float2 t = unpack_h2(rw[e], p.f16); f[2 * e] += t.x; f[2 * e + 1] += t.y; }
                }
              }
              if (p.out_f32) {                             // warp-uniform
                float* of = reinterpret_cast<float*>(p.out) + off + g * 8;
                reinterpret_cast<float4*>(of)[0] = make_float4(f[0], f[1], f[2], f[3]);
                reinterpret_cast<float4*>(of)[1] = make_float4(f[4], f[5], f[6], f[7]);
              } else {
                uint4 o;
                uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e) ow[e] = pack_h2(f[2 * e], f[2 * e + 1], p.f16);
                *reinterpret_cast<uint4*>(p.out + off + g * 8) = o;
              }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);          // 4 epilogue warps -> count 4
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
}

// ====================================================================================================================
// cta_group::2 variant: a CTA PAIR (cluster of 2 = one TPC) works on two neighbouring pixel tiles (16 x 8 each, side by
// side in w) of the same output-channel tile.  One tcgen05.mma.cta_group::2 of M = 256 covers both tiles: every CTA feeds
// its own 128 activation rows from its own shared memory, and only HALF of the weight tile (N_TILE/2 rows) -- the tensor
// core reads the two halves from the two SMs.  Per SM this halves the weight bytes read per MMA from shared memory (the
// operand port is what bounds the N = 64 layers: 6 KB per 32 tensor-clocks at cta_group::1, 5 KB here) and halves the
// L2 -> SM weight stream.  Protocol:
//   * both CTAs run the two TMA producer warps; the loads are cta_group::2 loads whose mbarrier is the LEADER's (rank 0)
//     full barrier, which expects the bytes of both CTAs;
//   * only the leader's warp 2 issues MMAs; its tcgen05.commit multicasts the arrival to the empty / acc_full barriers of
//     both CTAs;
//   * each CTA's epilogue warps drain their own TMEM (rows of their own tile) and arrive -- remotely for rank 1 -- on the
//     leader's acc_empty barrier (count 8).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  // relaxed: the arrival only hands the TMEM accumulator back (ordered by tcgen05.fence::before_thread_sync); a release here
  // would make the lane wait for all of its global stores to drain (ERRBAR: 7% of the samples in the ncu profile)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;    // shared::cluster address of the even (leader) CTA of a pair

__device__ __forceinline__ TileCoord tile_coord2(const TcParams& p, int u_, int rank, int n_tile_size, int P) {
  TileCoord c;
  uint32_t u = (uint32_t)u_, q;
  // u -> (phase, w, h, d, b, n-tile), fastest first; quotients by multiply + shift (dv_m / dv_s, host-computed): the five
  // integer divisions of the plain form were ~10 % of the epilogue warps' samples
  q = fast_div(u, p.dv_m[0], p.dv_s[0]); c.phase = (int)(u - q * (uint32_t)p.nphase); u = q;
  c.pc = c.phase & 1; c.pb = (c.phase >> 1) & 1; c.pa = (c.phase >> 2) & 1;
  if (p.pair_d) {
    q = fast_div(u, p.dv_m[1], p.dv_s[1]); c.w0 = (int)(u - q * (uint32_t)p.tiles_w) * TC_BW; u = q;
    q = fast_div(u, p.dv_m[2], p.dv_s[2]); c.h0 = (int)(u - q * (uint32_t)p.tiles_h) * p.bh; u = q;
    const uint32_t pairs_d = (uint32_t)p.groups_d >> 1;
    q = fast_div(u, p.dv_m[3], p.dv_s[3]); c.d0 = (int)((u - q * pairs_d) * 2 + rank) * P; u = q;
  } else {
    const uint32_t pairs_w = (uint32_t)p.tiles_w >> 1;
    q = fast_div(u, p.dv_m[1], p.dv_s[1]); c.w0 = (int)((u - q * pairs_w) * 2 + rank) * TC_BW; u = q;
    q = fast_div(u, p.dv_m[2], p.dv_s[2]); c.h0 = (int)(u - q * (uint32_t)p.tiles_h) * p.bh; u = q;
    q = fast_div(u, p.dv_m[3], p.dv_s[3]); c.d0 = (int)(u - q * (uint32_t)p.groups_d) * P; u = q;
  }
  q = fast_div(u, p.dv_m[4], p.dv_s[4]); c.b = (int)(u - q * (uint32_t)p.B); u = q;
  c.n0 = (int)u * n_tile_size;
  return c;
}

constexpr int TC2_THREADS = 384;   // warp 0/1: TMA producers, warp 2: MMA issuer, warp 3: idle, warps 4-11: epilogue
// T = 2 ("merged depth taps", 3-D, N_TILE = 64, plain 16-bit operands): an input plane j of the tile feeds up to two output planes
// (plane 0 through depth tap kd, plane 1 through kd - 1), and the two accumulators sit side by side in TMEM -- so ONE MMA of
// N = 128 whose B operand is [W(kd, khw) ; W(kd - 1, khw)] does both.  The activation rows (4 KB per MMA and SM) are then read
// once for two tap-GEMMs: 6 KB of shared-memory operands per 64 tensor clocks instead of 2 x 5 KB, which takes the 64-channel
// layers off the operand-port ceiling (0.80 of the tensor peak with N = 64 pairs).  The two CTAs of a pair each supply one of
// the two stacked taps (the hardware takes B rows 0..63 from rank 0 and 64..127 from rank 1), so a CTA loads whole taps (8 KB)
// where the N = 64 form loads half taps: to keep the L2 -> SM weight stream where it was, a CTA tile is T = 2 sub-tiles of
// 16 x 8 pixels stacked in h (one 34 x 10 patch per input plane) that share every weight slot; accumulators [t][plane][64],
// 256 columns, double-buffered = all 512.  Input planes are walked in the order 1, 2, 0, 3 (relative to d0 - 1) so that the
// first MMA into a buffer is a merged one that initialises both planes.
__device__ __forceinline__ int t2_plane(int jj) { return jj == 0 ? 1 : (jj == 1 ? 2 : (jj == 2 ? 0 : 3)); }
template <int T> struct Tc2Geom {
  static constexpr int PATCH_ROWS = TC_PW * (TC_BH * T + 2);
  static constexpr int PATCH_BYTES = PATCH_ROWS * 128;                       // T = 2: 43520
  static constexpr int PATCH_STRIDE = (PATCH_BYTES + 1023) / 1024 * 1024;    // T = 1: 23552, T = 2: 44032
};
template <int N_TILE, int P, int NA, int NB, bool UPS, int T = 1>
__global__ void __launch_bounds__(TC2_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapW, const TcParams p) {
  static_assert(T == 1 || (T == 2 && N_TILE == 64 && P == 2 && !UPS), "merged depth taps: N_TILE = 64, P = 2, no sub-pixel phases");
  constexpr int PATCH_BYTES = Tc2Geom<T>::PATCH_BYTES, PATCH_STRIDE = Tc2Geom<T>::PATCH_STRIDE;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)NA * PATCH_STRIDE;
  constexpr int B_HALF = T == 2 ? N_TILE * 128 : (N_TILE / 2) * 128;   // weight ring slot: this CTA's half of a tap tile (T = 2: a whole tap)
  __shared__ uint64_t full_a[NA], empty_a[NA], full_b[NB], empty_b[NB], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint4 stage_s[8][32 * 4];                  // per epilogue warp: 32 rows x 64 B (coalescing stage of the stores)
  __shared__ __align__(16) float bias_s[8][N_TILE];     // per epilogue warp: bias + chan_bias row of its current (sample, n-tile)
  constexpr uint32_t TMEM_COLS = 2 * T * P * N_TILE;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS >= 32, "TMEM columns");
  // accumulator sets per tile: 1 (plain), 2 (hh | lo; activations-split mode) or 4 (three hh sets + lo; fp32-parity mode,
  // N_TILE = 64 only).  Double-buffered across tiles when two buffers fit the 512 TMEM columns (1 set; 2 sets at N_TILE = 64),
  // else single-buffered (the epilogue of a tile then runs between its MMAs and the next tile's).
  const int nsets = p.nsets;
  constexpr uint32_t SET_COLS = T * P * N_TILE;         // columns of one accumulator set ([t][plane][N_TILE])
  const uint32_t buf_cols = nsets * SET_COLS, nbuf = 2 * buf_cols <= 512 ? 2u : 1u, tmem_cols = nbuf * buf_cols;
  const uint32_t lo_set = nsets - 1, nhh = nsets > 1 ? nsets - 1 : 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int total_pairs = p.total_tiles >> 1;
  const int KD = p.KD, NJ = P + KD - 1;
  const int nchunks = (p.Cin / 64) * p.vparts;          // virtual K chunks (split operands: 2 or 3 per real chunk)
  const int dpad = KD >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapW)) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                                   // barriers of both CTAs initialised before any remote arrive / TMA
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================== TMA producer: activation patches (both CTAs, own tile) =====================
    if (elect_one_sync()) {
      uint32_t seq = 0;
      for (int u = cluster_id; u < total_pairs; u += nclusters) {
        const TileCoord tc = tile_coord2(p, u, rank, N_TILE, P);
        VChunk vc;
        for (int c = 0; c < nchunks; ++c, vc.next(p))
          for (int jj = 0; jj < NJ; ++jj, ++seq) {
            const int j = T == 2 ? t2_plane(jj) : jj;       // T = 2: planes in the order 1, 2, 0, 3
            const uint32_t slot = seq % NA, ph = (seq / NA) & 1;
            mbar_wait(&empty_a[slot], ph ^ 1);
            if (leader) mbar_expect_tx(&full_a[slot], 2 * PATCH_BYTES);
            asm volatile(
                "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                    smem_u32(sA + (size_t)slot * PATCH_STRIDE)),
                "l"(reinterpret_cast<uint64_t>(&tmapA)), "r"(vc.a(p)), "r"(tc.w0 - 1 + p.pad_hw), "r"(tc.h0 - 1 + p.pad_hw), "r"(tc.d0 + j - dpad + p.pad_d), "r"(tc.b),
                "r"(smem_u32(&full_a[slot]) & kPeerBitMask)
                : "memory");
          }
      }
    }
  } else if (warp == 1) {
    // ===================== TMA producer: this CTA's half of every weight tap =====================
    if (elect_one_sync()) {
      uint32_t seq = 0;
      const int ntaps = UPS ? (KD == 3 ? 8 : 4) : KD * 9;
      for (int u = cluster_id; u < total_pairs; u += nclusters) {
        const TileCoord tc = tile_coord2(p, u, rank, N_TILE, P);
        const int tap_base = UPS ? tc.phase * ntaps : 0;
        VChunk vc;
        if constexpr (T == 2) {
          // slot (jj, khw): merged planes (jj < 2) take ONE WHOLE tap per CTA -- rank 0 the tap of output plane 0 (kd = j),
          // rank 1 the tap of output plane 1 (kd = j - 1) -- as two 32-row boxes; the outer planes take this CTA's half of
          // the single tap (kd = 0 for plane 0 of the tile, kd = 2 for plane 3)
          for (int c = 0; c < nchunks; ++c, vc.next(p))
            for (int jj = 0; jj < 4; ++jj) {
              const int j = t2_plane(jj);
              const bool wide = jj < 2;
              const int kd = wide ? j - (int)rank : (j == 0 ? 0 : 2);
              for (int khw = 0; khw < 9; ++khw, ++seq) {
                const uint32_t slot = seq % NB, ph = (seq / NB) & 1;
                mbar_wait(&empty_b[slot], ph ^ 1);
                if (leader) mbar_expect_tx(&full_b[slot], wide ? 2 * B_HALF : B_HALF);
                const int row = (kd * 9 + khw) * p.Cout + tc.n0 + (wide ? 0 : (int)rank * 32);
                for (int hb = 0; hb < (wide ? 2 : 1); ++hb)
                  asm volatile(
                      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                          smem_u32(sB + (size_t)slot * B_HALF + hb * 4096)),
                      "l"(reinterpret_cast<uint64_t>(&tmapW)), "r"(vc.w(p)), "r"(row + hb * 32),
                      "r"(smem_u32(&full_b[slot]) & kPeerBitMask)
                      : "memory");
              }
            }
          continue;
        }
        for (int c = 0; c < nchunks; ++c, vc.next(p))
          for (int tap = tap_base; tap < tap_base + ntaps; ++tap, ++seq) {
            const uint32_t slot = seq % NB, ph = (seq / NB) & 1;
            mbar_wait(&empty_b[slot], ph ^ 1);
            if (leader) mbar_expect_tx(&full_b[slot], 2 * B_HALF);
            asm volatile(
                "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                    smem_u32(sB + (size_t)slot * B_HALF)),
                "l"(reinterpret_cast<uint64_t>(&tmapW)), "r"(vc.w(p)), "r"(tap * p.Cout + tc.n0 + (int)rank * (N_TILE / 2)),
                "r"(smem_u32(&full_b[slot]) & kPeerBitMask)
                : "memory");
          }
      }
    }
  } else if (warp == 2) {
    if (leader) {
    // ===================== MMA issuer (leader CTA only) =====================
    const uint32_t idesc = umma_idesc_h16(N_TILE, 256, p.f16);
    constexpr uint32_t A_HI = umma_desc_hi(TC_PW * 128), B_HI = umma_desc_hi(1024);
    uint32_t seq_a = 0, seq_b = 0, it = 0;
    for (int u = cluster_id; u < total_pairs; u += nclusters, ++it) {
      const uint32_t as = nbuf == 1 ? 0u : (it & 1), aph = nbuf == 1 ? (it & 1) : ((it >> 1) & 1);
      mbar_wait(&acc_empty[as], aph ^ 1);               // the epilogues of BOTH CTAs have drained this accumulator set
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_acc = tmem_base + as * buf_cols;
      uint32_t used = 0, hh = 0;                         // accumulator sets written in this tile; round-robin counter
      int part = 0;                                      // part of the current virtual chunk (0: hi * hi)
      const TileCoord tc = tile_coord2(p, u, 0, N_TILE, P);
      for (int c = 0; c < nchunks; ++c) {
        const uint32_t seq_c = seq_a;
        if constexpr (T == 2) {
          const uint32_t idesc_w = umma_idesc_h16(2 * N_TILE, 256, p.f16);
          for (int jj = 0; jj < 4; ++jj) {
            const uint32_t s = seq_c + jj;
            mbar_wait(&full_a[s % NA], (s / NA) & 1);
            const uint32_t a_lo = umma_desc_lo(smem_u32(sA + (size_t)(s % NA) * PATCH_STRIDE));
            const bool wide = jj < 2;
            const uint32_t dcol = jj == 3 ? N_TILE : 0;      // plane 3 feeds output plane 1 only
            const uint32_t id = wide ? idesc_w : idesc;
#pragma unroll
            for (int khw = 0; khw < 9; ++khw, ++seq_b) {
              const uint32_t bs = seq_b % NB;
              mbar_wait(&full_b[bs], (seq_b / NB) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t b_lo = umma_desc_lo(smem_u32(sB + (size_t)bs * B_HALF));
              const uint32_t tap_off = ((khw / 3) * TC_PW + (khw % 3)) * 8;
              const uint32_t first = (c == 0 && jj == 0 && khw == 0) ? 0u : 1u;
              if (elect_one_sync()) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4)
                    umma_bf16_2cta(tmem_acc + t * (P * N_TILE) + dcol, umma_desc64(a_lo + tap_off + t * (TC_BH * TC_PW * 8) + k4 * 2, A_HI),
                                   umma_desc64(b_lo + k4 * 2, B_HI), id, k4 == 0 ? first : 1u);
                }
                umma_commit_2cta(&empty_b[bs]);
              }
              __syncwarp();
            }
            if (elect_one_sync()) umma_commit_2cta(&empty_a[s % NA]);
            __syncwarp();
          }
        } else if constexpr (UPS) {
          for (int j = 0; j < NJ; ++j) mbar_wait(&full_a[(seq_c + j) % NA], ((seq_c + j) / NA) & 1);
          const int ntd = KD == 3 ? 2 : 1;
          for (int td = 0; td < ntd; ++td) {
            const int kd = KD == 3 ? tc.pa + td : 0;
            uint32_t a_lo[P];
#pragma unroll
            for (int pp = 0; pp < P; ++pp)
              a_lo[pp] = umma_desc_lo(smem_u32(sA + (size_t)((seq_c + pp + kd) % NA) * TC_PATCH_STRIDE));
            for (int thw = 0; thw < 4; ++thw, ++seq_b) {
              const int kh = tc.pb + (thw >> 1), kw = tc.pc + (thw & 1);
              const uint32_t bs = seq_b % NB;
              mbar_wait(&full_b[bs], (seq_b / NB) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t b_lo = umma_desc_lo(smem_u32(sB + (size_t)bs * B_HALF));
              const uint32_t tap_off = (kh * TC_PW + kw) * 8;
              uint32_t set = 0;                                 // accumulator set of this tap's MMAs
              if (nsets > 1) { if (part) set = lo_set; else { set = hh; hh = hh + 1 == nhh ? 0 : hh + 1; } }
              const uint32_t first = (used >> set) & 1u;
              used |= 1u << set;
              const uint32_t acc_set = tmem_acc + set * SET_COLS;
              if (elect_one_sync()) {
#pragma unroll
                for (int pp = 0; pp < P; ++pp) {
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4)
                    umma_bf16_2cta(acc_set + pp * N_TILE, umma_desc64(a_lo[pp] + tap_off + k4 * 2, A_HI),
                                   umma_desc64(b_lo + k4 * 2, B_HI), idesc, k4 == 0 ? first : 1u);
                }
                umma_commit_2cta(&empty_b[bs]);
              }
              __syncwarp();
            }
          }
          if (elect_one_sync()) {
            for (int j = 0; j < NJ; ++j) umma_commit_2cta(&empty_a[(seq_c + j) % NA]);
          }
          __syncwarp();
        } else {
        for (int kd = 0; kd < KD; ++kd) {
          const int jlo = kd == 0 ? 0 : P - 1 + kd, jhi = P - 1 + kd;
          for (int j = jlo; j <= jhi; ++j) {
            const uint32_t s = seq_c + j;
            mbar_wait(&full_a[s % NA], (s / NA) & 1);
          }
          uint32_t a_lo[P];
#pragma unroll
          for (int pp = 0; pp < P; ++pp)
            a_lo[pp] = umma_desc_lo(smem_u32(sA + (size_t)((seq_c + pp + kd) % NA) * TC_PATCH_STRIDE));
#pragma unroll
          for (int khw = 0; khw < 9; ++khw, ++seq_b) {
            const uint32_t bs = seq_b % NB;
            mbar_wait(&full_b[bs], (seq_b / NB) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t b_lo = umma_desc_lo(smem_u32(sB + (size_t)bs * B_HALF));
            const uint32_t tap_off = ((khw / 3) * TC_PW + (khw % 3)) * 8;
            uint32_t set = 0;                                 // accumulator set of this tap's MMAs
            if (nsets > 1) { if (part) set = lo_set; else { set = hh; hh = hh + 1 == nhh ? 0 : hh + 1; } }
            const uint32_t first = (used >> set) & 1u;
            used |= 1u << set;
            const uint32_t acc_set = tmem_acc + set * SET_COLS;
            if (elect_one_sync()) {
#pragma unroll
              for (int pp = 0; pp < P; ++pp) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                  umma_bf16_2cta(acc_set + pp * N_TILE, umma_desc64(a_lo[pp] + tap_off + k4 * 2, A_HI),
                                 umma_desc64(b_lo + k4 * 2, B_HI), idesc, k4 == 0 ? first : 1u);
              }
              umma_commit_2cta(&empty_b[bs]);
            }
            __syncwarp();
          }
          const int rlo = kd, rhi = (kd == KD - 1) ? NJ - 1 : kd;
          if (elect_one_sync()) {
            for (int j = rlo; j <= rhi; ++j) umma_commit_2cta(&empty_a[(seq_c + j) % NA]);
          }
          __syncwarp();
        }
        }
        seq_a += NJ;
        if (++part == p.vparts) part = 0;
      }
      if (elect_one_sync()) umma_commit_2cta(&acc_full[as]);
      __syncwarp();
    }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own tile / own TMEM) =====================
    // 8 warps: warp w serves TMEM lane quadrant q = w & 3 (hardware rule) and plane pp = (w - 4) >> 2 of the tile, so the two
    // planes drain concurrently.  Everything the epilogue reads from global memory for a tile (residual rows) is requested
    // BEFORE the wait on the accumulator, so the load latency hides behind the tile's MMAs.
    static_assert(P == 2, "one epilogue warp set per plane");
    const int q = warp & 3, pp = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int line = row >> 3, wp = row & 7;
    // Fused norm statistics (optional): per output channel, sum and sum of squares of the fp32 results over the pixels
    // this warp stores, accumulated in registers while the CTA stays inside one (n-tile, sample) range of the tile order
    // and flushed (combined over the CTA's epilogue warps) to stats[sample][slot = cta][channel] when it leaves it -- no
    // atomics, fixed summation order.
    // Lane l of the warp ends up owning channel c0 + l of every 32-channel group (butterfly transpose-reduce).
    constexpr int NG = N_TILE / 32;
    constexpr int DEP = NG < 2 ? NG : 2;                  // residual prefetch depth in 32-channel groups
    float st_s[NG], st_q[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) { st_s[g] = 0.0f; st_q[g] = 0.0f; }
    const bool do_stats = p.stats != nullptr;
    // a "range" = all tiles of one (n-tile, sample).  3-D: sample = batch entry (both planes of a tile belong to it);
    // 2-D: the plane axis IS the batch, a tile holds P samples and this warp (plane pp) owns sample d0 + pp.
    const bool is3d = p.KD == 3;
    const int srange = is3d ? p.B : p.groups_d;              // sample ranges per n-tile
    const int per_range = p.pair_d ? 1 : p.nphase * (p.tiles_w >> 1) * p.tiles_h * (is3d ? p.groups_d : 1);   // pair_d: no statistics
    const int nranges = total_pairs / per_range;             // n_tiles * srange
    int cur_range = -1;
    uint4* stg = stage_s[warp - 4];
    // bias[n0 + c] + chan_bias[sample][n0 + c] of the warp's current (sample row, n-tile): the tile order keeps both fixed for
    // many tiles, so the 2 x N_TILE scalar loads per row (13 % of the epilogue's samples when issued per tile behind the
    // accumulator wait) happen once per change and the tiles read the sum as shared-memory broadcasts
    float* bsm = bias_s[warp - 4];
    int bias_key = -1;
    const bool any_bias = p.bias != nullptr || p.chan_bias != nullptr;
    // flush: the 8 epilogue warps of the CTA reach a range change together (same tile sequence); their partials meet in
    // shared memory (the store stage is idle between tiles) and ONE slot per CTA goes to global memory, summed in a fixed
    // order.  3-D: all 8 warps belong to one sample; 2-D: the warps of plane pp belong to sample d0 + pp.
    auto flush = [&](int range, bool zero) {
      const int nt = range / srange, sr = range - nt * srange;  // range -> (n-tile, sample range)
      float2* mine = reinterpret_cast<float2*>(stage_s[warp - 4]);      // each warp parks its partials in its OWN stage
      // channel of the 32-channel group this lane's running sums belong to: l (accumulator-mapping butterfly, 16-bit outputs) or
      // 16 b4 + 4 (l & 3) + 2 b3 + b2 (store-mapping reduce-scatter, fp32 outputs)
      const int own = p.out_f32 ? (((lane >> 4) & 1) * 16 + (lane & 3) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)) : lane;
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        mine[g * 32 + own] = zero ? make_float2(0.0f, 0.0f) : make_float2(st_s[g], st_q[g]);
        st_s[g] = 0.0f; st_q[g] = 0.0f;
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const bool lead = is3d ? warp == 4 : (warp == 4 || warp == 8);
      const int sg = is3d ? sr : sr * P + pp;
      if (lead && sg < p.samples) {
        const int w0 = warp - 4, nw = is3d ? 8 : 4;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          float2 t = make_float2(0.0f, 0.0f);
          for (int w = w0; w < w0 + nw; ++w) {
            const float2 v = reinterpret_cast<const float2*>(stage_s[w])[g * 32 + lane];
            t.x += v.x; t.y += v.y;
          }
          p.stats[((int64_t)sg * TC_STAT_SLOTS + blockIdx.x) * p.Cout + nt * N_TILE + g * 32 + lane] = t;
        }
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
    };
    uint32_t it = 0;
    for (int u = cluster_id; u < total_pairs; u += nclusters, ++it) {
      const TileCoord tc0 = tile_coord2(p, u, rank, N_TILE, P);
      if (do_stats) {
        const int r = u / per_range;
        if (r != cur_range) {
          if (cur_range >= 0) flush(cur_range, false);
          for (int z = cur_range + 1; z < r; ++z) flush(z, true);
          cur_range = r;
        }
      }
#pragma unroll 1
      for (int t = 0; t < T; ++t) {                           // T = 2: the two 16-row sub-tiles of the CTA tile, one accumulator buffer
      TileCoord tc = tc0;
      tc.h0 += t * TC_BH;
      const int h = tc.h0 + line, w = tc.w0 + wp;
      const int d = tc.d0 + pp;
      const bool valid = h < p.H && w < p.W && d < p.D;
      int64_t pix;
      if constexpr (UPS) {
        const int od = p.KD == 3 ? 2 * d + tc.pa : d, OD = p.KD == 3 ? 2 * p.D : p.D;
        pix = (((int64_t)tc.b * OD + od) * (2 * p.H) + (2 * h + tc.pb)) * (2 * p.W) + (2 * w + tc.pc);
      } else {
        pix = (((int64_t)tc.b * p.D + d) * p.H + h) * p.W + w;
      }
      const int brow = (tc.b * p.D + d) / p.planes_per_sample;
      // store-side mapping (see the staged store below): this lane writes column chunk lane&3 of rows (line q*4+k, w lane>>2)
      const int w_st = tc.w0 + (lane >> 2);
      const bool st_ok = w_st < p.W && d < p.D;
      int64_t st_pix;
      if constexpr (UPS) {
        const int od = p.KD == 3 ? 2 * d + tc.pa : d, OD = p.KD == 3 ? 2 * p.D : p.D;
        st_pix = (((int64_t)tc.b * OD + od) * (2 * p.H) + (2 * (tc.h0 + q * 4) + tc.pb)) * (2 * p.W) + (2 * w_st + tc.pc);
      } else {
        st_pix = (((int64_t)tc.b * p.D + d) * p.H + tc.h0 + q * 4) * p.W + w_st;
      }
      const bool of32 = p.out_f32 != 0;                     // fp32 output (fp32-storage modes); warp-uniform
      const bool rf32 = p.res_f32 != 0;                     // fp32 residual (independent of the output format); warp-uniform
      const int64_t roff = pix * p.Cout + tc.n0;
      const uint16_t* rrow = (p.residual != nullptr && valid && !rf32) ? p.residual + roff : nullptr;
      const float* rrow32 = (p.residual != nullptr && valid && rf32) ? reinterpret_cast<const float*>(p.residual) + roff : nullptr;
      if (any_bias) {
        // d >= p.D: the second plane of a tile past an odd plane count (e.g. a 2-D batch of 1) -- its chan_bias row does not exist
        const int key = (d < p.D ? brow : p.B * p.D) * p.n_tiles + tc.n0 / N_TILE;
        if (key != bias_key) {                              // warp-uniform
          bias_key = key;
          __syncwarp();
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            float bv = p.bias != nullptr ? __ldg(p.bias + tc.n0 + g * 32 + lane) : 0.0f;
            if (p.chan_bias != nullptr && d < p.D) bv += __ldg(p.chan_bias + (int64_t)brow * p.Cout + tc.n0 + g * 32 + lane);
            bsm[g * 32 + lane] = bv;
          }
          __syncwarp();
        }
      }
      uint4 rr[DEP][4];
      if (rrow != nullptr) {
#pragma unroll
        for (int gg = 0; gg < DEP; ++gg)
#pragma unroll
          for (int g = 0; g < 4; ++g) rr[gg][g] = *reinterpret_cast<const uint4*>(rrow + gg * 32 + g * 8);
      }
      // fp32 OUTPUT: residual add and statistics happen in the STORE mapping (after the stage has transposed the tile: lane =
      // (row 8k + lane/4, 16-byte chunk lane&3)), so the residual is read exactly like the output is written -- 4 lanes per
      // 64 contiguous bytes, 8 wavefronts per load instruction.  Reading it in the accumulator mapping (a thread per pixel row,
      // 32 lines per instruction) cost 256 L1 wavefronts per warp and 32-channel group in a kernel that is bound by the
      // shared-memory / L1 data path: 4x the stage traffic it avoided.
      const bool res_st = of32 && p.residual != nullptr;
      const float* res32 = reinterpret_cast<const float*>(p.residual);
      auto st_row = [&](int k) -> int64_t {                 // pixel index of stage row 8k + lane/4
        if constexpr (UPS) return st_pix + (int64_t)k * 2 * (2 * p.W);
        else return st_pix + (int64_t)k * p.W;
      };
      auto load_res_st = [&](int c0) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            rr[hh][k] = make_uint4(0, 0, 0, 0);
            if (st_ok && tc.h0 + q * 4 + k < p.H)
              rr[hh][k] = __ldcs(reinterpret_cast<const uint4*>(res32 + st_row(k) * p.Cout + tc.n0 + c0 + hh * 16 + (lane & 3) * 4));
          }
      };
      if constexpr (DEP == 2) {
        if (res_st) {
          load_res_st(0);
        } else if (rrow32 != nullptr) {
          // fp32 residual into a 16-bit output (operand copy): accumulator mapping, ONE 32-channel group (8 float4) in the same
          // 32 registers, requested before the accumulator wait and refilled with the next group while the current one is stored
#pragma unroll
          for (int k = 0; k < 8; ++k) rr[k >> 2][k & 3] = *reinterpret_cast<const uint4*>(rrow32 + k * 4);
        }
      }
      if (p.residual != nullptr && u + nclusters < total_pairs) {
        // the residual rows of the NEXT tile start their trip from HBM now (one tile time ahead, no registers held)
        const TileCoord tn = tile_coord2(p, u + nclusters, rank, N_TILE, P);
        const int hn = tn.h0 + t * TC_BH + line, wn = tn.w0 + wp, dn = tn.d0 + pp;
        if (hn < p.H && wn < p.W && dn < p.D) {
          int64_t pn;
          if constexpr (UPS) {
            const int od = p.KD == 3 ? 2 * dn + tn.pa : dn, OD = p.KD == 3 ? 2 * p.D : p.D;
            pn = (((int64_t)tn.b * OD + od) * (2 * p.H) + (2 * hn + tn.pb)) * (2 * p.W) + (2 * wn + tn.pc);
          } else {
            pn = (((int64_t)tn.b * p.D + dn) * p.H + hn) * p.W + wn;
          }
          const int es = rf32 ? 4 : 2;
          const char* rn = reinterpret_cast<const char*>(p.residual) + (pn * p.Cout + tn.n0) * es;
          for (int c = 0; c < N_TILE * es / 128; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(rn + c * 128));
        }
      }
      const uint32_t as = nbuf == 1 ? 0u : (it & 1), aph = nbuf == 1 ? (it & 1) : ((it >> 1) & 1);
      if (t == 0) mbar_wait(&acc_full[as], aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + as * buf_cols + (t * P + pp) * N_TILE + ((uint32_t)(q * 32) << 16);
      if (p.dbg & 1) {                                       // measurement: MMA / TMA side alone
        if (t == T - 1) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&acc_empty[as], 0);
        }
        continue;
      }
#pragma unroll
      for (int gi = 0; gi < NG; ++gi) {
        const int c0 = gi * 32;
        uint32_t v[32];
        DSK_TMEM_LD_X32(v, taddr + c0);
        if (nsets > 1) {                                    // split mode: (s0 + s1 + s2) + 2^-11 * s3, fp32 round-to-nearest adds
          uint32_t u[32];
#pragma unroll 1
          for (int sidx = 1; sidx < nsets; ++sidx) {
            DSK_TMEM_LD_X32(u, taddr + sidx * SET_COLS + c0);
            const float sc = sidx == (int)lo_set ? kLoScale : 1.0f;
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(fmaf(__uint_as_float(u[e]), sc, __uint_as_float(v[e])));
          }
        }
        float f[32];
        if (any_bias) {
          const float4* b4 = reinterpret_cast<const float4*>(bsm + c0);
#pragma unroll
          for (int e4 = 0; e4 < 8; ++e4) {
            const float4 bb = b4[e4];
            f[4 * e4] = __uint_as_float(v[4 * e4]) + bb.x; f[4 * e4 + 1] = __uint_as_float(v[4 * e4 + 1]) + bb.y;
            f[4 * e4 + 2] = __uint_as_float(v[4 * e4 + 2]) + bb.z; f[4 * e4 + 3] = __uint_as_float(v[4 * e4 + 3]) + bb.w;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
        }
        if (rrow != nullptr) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(&rr[gi % DEP][g]);
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float2 t = unpack_h2(rw[e], p.f16); f[g * 8 + 2 * e] += t.x; f[g * 8 + 2 * e + 1] += t.y; }
          }
          if (gi + DEP < NG) {                               // refill the slot just consumed
#pragma unroll
            for (int g = 0; g < 4; ++g) rr[gi % DEP][g] = *reinterpret_cast<const uint4*>(rrow + (gi + DEP) * 32 + g * 8);
          }
        }
        if constexpr (DEP == 2) {
          if (rrow32 != nullptr && !res_st) {
#pragma unroll
            for (int e4 = 0; e4 < 8; ++e4) {
              const uint4 r = rr[e4 >> 2][e4 & 3];
              f[4 * e4] += __uint_as_float(r.x); f[4 * e4 + 1] += __uint_as_float(r.y);
              f[4 * e4 + 2] += __uint_as_float(r.z); f[4 * e4 + 3] += __uint_as_float(r.w);
            }
            if (gi + 1 < NG) {
#pragma unroll
              for (int k = 0; k < 8; ++k) rr[k >> 2][k & 3] = *reinterpret_cast<const uint4*>(rrow32 + (gi + 1) * 32 + k * 4);
            }
          }
        }
        if (of32) {
          // fp32 rows are 128 B per 32-channel group: the same 2 KB stage takes them as two 16-channel halves.  (Measured: storing
          // the thread's full 128-byte line straight from registers instead -- no shared-memory traffic next to the operand
          // fetches -- is SLOWER: 64->64 @ 64^3, B = 8: 403 -> 424 us plain, 539 -> 585 us with residual + statistics.)
          float* outf = reinterpret_cast<float*>(p.out);
          float ps[8], pq[8];                                // this lane's (half, channel-of-chunk) partial sums over its 4 rows
#pragma unroll
          for (int e = 0; e < 8; ++e) { ps[e] = 0.0f; pq[e] = 0.0f; }
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              stg[lane * 4 + (g ^ ((lane >> 1) & 3))] =
                  make_uint4(__float_as_uint(f[hh * 16 + g * 4]), __float_as_uint(f[hh * 16 + g * 4 + 1]),
                             __float_as_uint(f[hh * 16 + g * 4 + 2]), __float_as_uint(f[hh * 16 + g * 4 + 3]));
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int r = 8 * k + (lane >> 2);
              const uint4 sv = stg[r * 4 + ((lane & 3) ^ ((r >> 1) & 3))];
              float4 val = make_float4(__uint_as_float(sv.x), __uint_as_float(sv.y), __uint_as_float(sv.z), __uint_as_float(sv.w));
              if (res_st) {
                val.x += __uint_as_float(rr[hh][k].x); val.y += __uint_as_float(rr[hh][k].y);
                val.z += __uint_as_float(rr[hh][k].z); val.w += __uint_as_float(rr[hh][k].w);
              }
              if (st_ok && tc.h0 + q * 4 + k < p.H) {
                __stcs(reinterpret_cast<float4*>(outf + st_row(k) * p.Cout + tc.n0 + c0 + hh * 16 + (lane & 3) * 4), val);
                if (do_stats) {
                  ps[hh * 4] += val.x; ps[hh * 4 + 1] += val.y; ps[hh * 4 + 2] += val.z; ps[hh * 4 + 3] += val.w;
                  pq[hh * 4] = fmaf(val.x, val.x, pq[hh * 4]); pq[hh * 4 + 1] = fmaf(val.y, val.y, pq[hh * 4 + 1]);
                  pq[hh * 4 + 2] = fmaf(val.z, val.z, pq[hh * 4 + 2]); pq[hh * 4 + 3] = fmaf(val.w, val.w, pq[hh * 4 + 3]);
                }
              }
            }
            __syncwarp();
          }
          if (res_st && gi + 1 < NG) load_res_st((gi + 1) * 32);   // next group's residual while this one's statistics reduce
          if (do_stats) {
            // reduce-scatter over the 8 lanes that share a chunk (lane bits 4, 3, 2): 8 values -> 1 per lane in 3 halving steps
            // (7 + 7 shuffles; the accumulator-mapping butterfly of the 16-bit path needs 31 + 31).  Lane keeps index
            // 4 b4 + 2 b3 + b2 = (half b4, channel 2 b3 + b2 of its chunk): channel 16 b4 + 4 (lane & 3) + 2 b3 + b2 of the group.
#pragma unroll
            for (int o = 4; o >= 1; o >>= 1) {
              const bool up = (lane & (o << 2)) != 0;
#pragma unroll
              for (int i = 0; i < o; ++i) {
                const float ks = up ? ps[i + o] : ps[i], gs = up ? ps[i] : ps[i + o];
                const float kq = up ? pq[i + o] : pq[i], gq = up ? pq[i] : pq[i + o];
                ps[i] = ks + __shfl_xor_sync(0xffffffffu, gs, o << 2);
                pq[i] = kq + __shfl_xor_sync(0xffffffffu, gq, o << 2);
              }
            }
            st_s[gi] += ps[0];
            st_q[gi] += pq[0];
          }
        } else {
          // (slot swizzle g ^ ((row >> 1) & 3): rows are 64 B apart, so rows r and r + 2 start in the same bank and the 8 lanes of a
          // 128-bit store phase need 4 distinct slot permutations; the first version used g ^ (row & 3) and paid a 2-way bank
          // conflict on every stage write -- 28 % of the epilogue's shared-memory wavefronts in a kernel bound by that port)
          // staged, transposed store: a thread holds one pixel ROW, so a direct st.global.v4 touches 32 lines per
          // instruction; through this warp's 2 KB stage (XOR-swizzled 16-byte slots) 4 lanes write 64 contiguous bytes
          // of a row and one instruction covers 8 rows.  Row r = 8k + lane/4 of the warp is pixel (line q*4 + k, w lane/4).
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o;
            uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) ow[e] = pack_h2(f[g * 8 + 2 * e], f[g * 8 + 2 * e + 1], p.f16);
            stg[lane * 4 + (g ^ ((lane >> 1) & 3))] = o;
          }
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int r = 8 * k + (lane >> 2);
            const uint4 val = stg[r * 4 + ((lane & 3) ^ ((r >> 1) & 3))];
            if (st_ok && tc.h0 + q * 4 + k < p.H) {
              int64_t px;
              if constexpr (UPS) px = st_pix + (int64_t)k * 2 * (2 * p.W);
              else px = st_pix + (int64_t)k * p.W;
              *reinterpret_cast<uint4*>(p.out + px * p.Cout + tc.n0 + c0 + (lane & 3) * 8) = val;
            }
          }
          __syncwarp();
        }
        if (do_stats && !of32) {
          // butterfly transpose-reduce over the 32 lanes (= 32 pixels): lane l keeps channel c0 + l
          float sq[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) { f[e] = valid ? f[e] : 0.0f; sq[e] = f[e] * f[e]; }
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int i = 0; i < o; ++i) {
              const float ks = up ? f[i + o] : f[i], gs = up ? f[i] : f[i + o];
              const float kq = up ? sq[i + o] : sq[i], gq = up ? sq[i] : sq[i + o];
              f[i] = ks + __shfl_xor_sync(0xffffffffu, gs, o);
              sq[i] = kq + __shfl_xor_sync(0xffffffffu, gq, o);
            }
          }
          st_s[gi] += f[0];
          st_q[gi] += sq[0];
        }
      }
      if (t == T - 1) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&acc_empty[as], 0);   // 2 CTAs x 8 epilogue warps -> count 16 on the leader
      }
      }
    }
    if (do_stats) {
      if (cur_range >= 0) flush(cur_range, false);
      for (int z = cur_range + 1; z < nranges; ++z) flush(z, true);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                                    // the peer's shared memory / barriers stay alive until both are done
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
}

template <int N_TILE, int P, int NA, int NB, bool UPS, int T = 1>
static int launch_tc2(const CUtensorMap& ta, const CUtensorMap& tw, const TcParams& p, cudaStream_t st) {
  const size_t smem = (size_t)NA * Tc2Geom<T>::PATCH_STRIDE + (size_t)NB * (T == 2 ? N_TILE : N_TILE / 2) * 128 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<N_TILE, P, NA, NB, UPS, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("conv_tc2: cudaFuncSetAttribute(%zu B smem): %s", smem, cudaGetErrorString(e)); return DSK_ERR_CUDA; }
    configured = true;
  }
  const int pairs = p.total_tiles / 2;
  const int clusters = pairs < DSK_NUM_SMS / 2 ? pairs : DSK_NUM_SMS / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(TC2_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc2_kernel<N_TILE, P, NA, NB, UPS, T>, ta, tw, p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) { set_error("conv_tc2_kernel launch failed: %s", cudaGetErrorString(e)); return DSK_ERR_CUDA; }
  return DSK_OK;
}

// nearest x2 upsample of a channels-last bf16 tensor (input of the UpSampler conv on the tcgen05 path)
__global__ void __launch_bounds__(256) upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int D, int H, int W,
                                                          int C8, int ndim) {
  const int Do = ndim == 3 ? D * 2 : D, Ho = H * 2, Wo = W * 2;
  const int64_t total = (int64_t)B * Do * Ho * Wo * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C8);
    int64_t r = i / C8;
    int w = (int)(r % Wo); r /= Wo;
    int h = (int)(r % Ho); r /= Ho;
    int d = (int)(r % Do);
    int b = (int)(r / Do);
    const int ds = ndim == 3 ? d >> 1 : d;
    y[i] = x[((((int64_t)b * D + ds) * H + (h >> 1)) * W + (w >> 1)) * C8 + c];
  }
}

template <int N_TILE, int P, int NA, int NB, bool UPS>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tw, const TcParams& p, cudaStream_t st) {
  const size_t smem = (size_t)NA * TC_PATCH_STRIDE + (size_t)NB * N_TILE * 128 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<N_TILE, P, NA, NB, UPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("conv_tc: cudaFuncSetAttribute(%zu B smem): %s", smem, cudaGetErrorString(e)); return DSK_ERR_CUDA; }
    configured = true;
  }
  const int grid = p.total_tiles < DSK_NUM_SMS ? p.total_tiles : DSK_NUM_SMS;
  DSK_LAUNCH((conv_tc_kernel<N_TILE, P, NA, NB, UPS>), grid, TC_THREADS, smem, st, ta, tw, p);
  return DSK_OK;
}

// Sub-pixel weights of conv3(nearest_up2(x)):  out[2i+a] = sum_k w[k] x[i + floor((a+k-1)/2)].  Per axis the three taps
// collapse onto two input offsets: a = 0: offset -1 <- {k0}, 0 <- {k1,k2};  a = 1: 0 <- {k0,k1}, +1 <- {k2}.
// Layout: bf16 [phase = (a*2+b)*2+c][tap = (td*2+th)*2+tw][Cout][Cin] (2-D: phase = b*2+c, tap = th*2+tw), summed in fp32.
__global__ void __launch_bounds__(256) pack_upconv_weight_kernel(const float* __restrict__ w, uint16_t* __restrict__ o, int Cout,
                                                                  int Cin, int ndim, int fmt) {
  const int nax = ndim, nph = 1 << nax, ntap = 1 << nax, k3 = ndim == 3 ? 27 : 9;
  const int64_t total = (int64_t)nph * ntap * Cout * Cin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    int64_t r = i / Cin;
    const int co = (int)(r % Cout); r /= Cout;
    const int tap = (int)(r % ntap);
    const int ph = (int)(r / ntap);
    // per-axis parity and 2-tap index, innermost axis = w
    int par[3], tt[3];
    for (int ax = 0; ax < nax; ++ax) { par[nax - 1 - ax] = (ph >> ax) & 1; tt[nax - 1 - ax] = (tap >> ax) & 1; }
    float acc = 0.0f;
    const float* wp = w + ((int64_t)co * Cin + ci) * k3;
    for (int k = 0; k < k3; ++k) {
      int kk[3];
      int q = k;
      for (int ax = nax - 1; ax >= 0; --ax) { kk[ax] = q % 3; q /= 3; }
      bool match = true;
      for (int ax = 0; ax < nax; ++ax) {
        // low-res offset of tap kk for parity par: floor((par + kk - 1) / 2) in {-1, 0, 1}; 2-tap index = offset - (par - 1)
        const int off = (par[ax] + kk[ax] - 1 + 2) / 2 - 1;
        if (off - (par[ax] - 1) != tt[ax]) match = false;
      }
      if (match) acc += wp[k];
    }
    put16(o, i / Cin, Cin, ci, acc, fmt);
  }
}

// The same packing with one thread per (co, ci) pair: the 3^d reference taps are read once (contiguous) and scattered into the
// 2^d x 2^d (phase, tap) sums held in registers, in ascending-k order -- bit-identical to the kernel above, which spends 27
// predicated iterations per OUTPUT element (121 us per 256x128x27 weight, twice per training iteration).
template <int NAX>
__global__ void __launch_bounds__(128) pack_upconv_weight_pair_kernel(const float* __restrict__ w, uint16_t* __restrict__ o, int Cout,
                                                                       int Cin, int fmt) {
  constexpr int NPH = 1 << NAX, K3 = NAX == 3 ? 27 : 9;
  const int64_t pairs = (int64_t)Cout * Cin;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pairs) return;
  float acc[NPH][NPH];
#pragma unroll
  for (int a = 0; a < NPH; ++a)
#pragma unroll
    for (int b = 0; b < NPH; ++b) acc[a][b] = 0.0f;
  const float* wp = w + i * K3;
#pragma unroll
  for (int k = 0; k < K3; ++k) {
    const float v = wp[k];
    int kk[3] = {0, 0, 0};
    {
      int q = k;
#pragma unroll
      for (int ax = NAX - 1; ax >= 0; --ax) { kk[ax] = q % 3; q /= 3; }
    }
#pragma unroll
    for (int ph = 0; ph < NPH; ++ph) {
      int tap = 0;
#pragma unroll
      for (int ax = 0; ax < NAX; ++ax) {                       // axis ax: parity bit / tap bit (NAX-1-ax), as in the kernel above
        const int par = (ph >> (NAX - 1 - ax)) & 1;
        const int off = (par + kk[ax] - 1 + 2) / 2 - 1;
        tap |= (off - (par - 1)) << (NAX - 1 - ax);
      }
      acc[ph][tap] += v;
    }
  }
#pragma unroll
  for (int ph = 0; ph < NPH; ++ph)
#pragma unroll
    for (int tap = 0; tap < NPH; ++tap)
      put16(o, ((int64_t)ph * NPH + tap) * Cout + i / Cin, Cin, (int)(i % Cin), acc[ph][tap], fmt);
}

}  // namespace dsk

using namespace dsk;

extern "C" int dsk_pack_upconv_weight(const float* w_ref, void* w_packed, int Cout, int Cin, int ndim, int dtype, void* stream) {
  DSK_REQUIRE(w_ref && w_packed && Cout > 0 && Cin > 0 && (ndim == 2 || ndim == 3), "dsk_pack_upconv_weight: bad arguments");
  DSK_REQUIRE(is_h16(dtype) || dtype == DSK_SPLIT_F16, "dsk_pack_upconv_weight: bad dtype %d", dtype);
  static const int old_path = [] { const char* e = getenv("DSK_PACK_UPCONV_OLD"); return e ? atoi(e) : 0; }();   // A/B + bit-exactness tests
  const int64_t pairs = (int64_t)Cout * Cin;
  if (!old_path) {
    const int grid = (int)((pairs + 127) / 128);
    if (ndim == 3) DSK_LAUNCH(pack_upconv_weight_pair_kernel<3>, grid, 128, 0, as_stream(stream), w_ref, (uint16_t*)w_packed, Cout, Cin, dtype);
    else DSK_LAUNCH(pack_upconv_weight_pair_kernel<2>, grid, 128, 0, as_stream(stream), w_ref, (uint16_t*)w_packed, Cout, Cin, dtype);
    return DSK_OK;
  }
  const int64_t total = (int64_t)(1 << ndim) * (1 << ndim) * pairs;
  DSK_LAUNCH(pack_upconv_weight_kernel, grid_for(total, 256, 8), 256, 0, as_stream(stream), w_ref, (uint16_t*)w_packed, Cout, Cin,
             ndim, dtype);
  return DSK_OK;
}

extern "C" int dsk_upsample2x(const void* x, void* y, int B, int D, int H, int W, int C, int ndim, int dtype, void* stream) {
  DSK_REQUIRE(x && y && B > 0 && D > 0 && H > 0 && W > 0 && C > 0, "dsk_upsample2x: bad arguments");
  DSK_REQUIRE(is_h16(dtype) && C % 8 == 0, "dsk_upsample2x: 16-bit tensors with C %% 8 == 0 only (got dtype %d, C %d)", dtype, C);
  const int64_t total = (int64_t)B * (ndim == 3 ? 2 * D : D) * 2 * H * 2 * W * (C / 8);
  DSK_LAUNCH(upsample2x_kernel, grid_for(total, 256, 16), 256, 0, as_stream(stream), (const uint4*)x, (uint4*)y, B, D, H, W, C / 8, ndim);
  return DSK_OK;
}

namespace dsk {
int convout_tc_dispatch(const dsk_conv_desc* d, const void* in, const void* w, const float* bias, void* out, cudaStream_t st, int padded);
}

// operand / output format combinations the tcgen05 convolutions take:
//   in bf16, w bf16 | in fp16, w fp16 | in split-fp16, w split-fp16 (3 MMAs per k-step) | in split-fp16, w fp16 (2 MMAs);
//   out: the 16-bit format of the operands, or fp32 (split parity mode: residual fp32 too)
static bool tc_formats_ok(const dsk_conv_desc* d) {
  const bool ops = (d->in_dtype == DSK_BF16 && d->w_dtype == DSK_BF16) || (d->in_dtype == DSK_F16 && d->w_dtype == DSK_F16) ||
                   (d->in_dtype == DSK_SPLIT_F16 && (d->w_dtype == DSK_SPLIT_F16 || d->w_dtype == DSK_F16));
  if (!ops) return false;
  if (d->out_nchw_f32) return true;
  const int fmt16 = d->in_dtype == DSK_BF16 ? DSK_BF16 : DSK_F16;
  if (d->res_dtype != DSK_RES_SAME && d->res_dtype != DSK_RES_F32) return false;
  return d->out_dtype == fmt16 || d->out_dtype == DSK_F32;
}
static bool tc_pair_eligible(const dsk_conv_desc* d) {
  static const int force_cg = [] { const char* e = getenv("DSK_CONV_CG"); return e ? atoi(e) : 0; }();
  if (force_cg == 1 || d->ksize != 3 || !tc_formats_ok(d) || d->out_nchw_f32 || d->Cin % 64 != 0 || d->Cout % 64 != 0)
    return false;
  const int iW = d->up2 ? d->W / 2 : d->W;
  return (((iW + TC_BW - 1) / TC_BW) % 2) == 0;
}

// odd number of w-tiles: pair along the plane axis instead (both CTAs still share every weight tap); needs an even number of
// plane groups per batch entry and no fused statistics (the two CTAs of a pair may work on different samples)
static bool tc_pair_d_eligible(const dsk_conv_desc* d, bool stats) {
  static const int force_cg = [] { const char* e = getenv("DSK_CONV_CG"); return e ? atoi(e) : 0; }();
  if (stats || force_cg == 1 || d->ksize != 3 || !tc_formats_ok(d) || d->out_nchw_f32 || d->Cin % 64 != 0 || d->Cout % 64 != 0 ||
      tc_pair_eligible(d))
    return false;
  const int iD = (d->up2 && d->ndim == 3) ? d->D / 2 : d->D;
  const int planes = d->ndim == 3 ? iD : d->B;
  return (((planes + 1) / 2) % 2) == 0;      // groups_d with P = 2
}

// 1 if dsk_conv_fwd_stats should emit fused norm statistics for this convolution: the cta_group::2 kernel, 3-D only.
// (The epilogue can do it for 2-D too -- tests exercise it with DSK_CONV_STATS_2D=1 -- but a 2-D tile has 3x fewer MMAs to
// hide the reduction behind and one slot set per sample of a large batch: measured on C5 / C2 it does not pay.)
// 2-D, fp32 OUTPUT (the fp32-storage modes): since the statistics of that epilogue became a 14-shuffle reduce-scatter in the store
// mapping they pay in 2-D as well -- where a sample spans enough tile pairs that the per-range flush is rare (>= 32: planes of
// 128^2 and up; C5 evaluation at B = 32 10.27 -> 10.0 ms; the 28^2 planes of MNIST flush every 4 tiles and lose).  16-bit outputs
// keep the butterfly and stay off in 2-D (C5 bf16 8.14 -> 8.89 ms when forced).  DSK_CONV_STATS_2D = 1 forces all on, -1 all off.
extern "C" int dsk_conv_stats_supported(const dsk_conv_desc* d) {
  static const int allow2d = [] { const char* e = getenv("DSK_CONV_STATS_2D"); return e ? atoi(e) : 0; }();
  if (d == nullptr || !tc_pair_eligible(d)) return 0;
  if (d->ndim == 3 || allow2d > 0) return 1;
  if (allow2d < 0 || d->out_dtype != DSK_F32 || d->out_nchw_f32) return 0;
  const int iH = d->up2 ? d->H / 2 : d->H, iW = d->up2 ? d->W / 2 : d->W;
  const int pairs = (((iW + TC_BW - 1) / TC_BW) / 2) * ((iH + TC_BH - 1) / TC_BH) * (d->up2 ? 4 : 1);
  return pairs >= 32 ? 1 : 0;
}
extern "C" int dsk_conv_stats_slots(void) { return TC_STAT_SLOTS; }

static int conv_fwd_tc_impl(const dsk_conv_desc* d, const void* in, const void* w, const float* bias, const float* chan_bias,
                            const void* residual, void* out, float2* stats, void* pad_ws, void* stream);

extern "C" int dsk_conv_fwd_tc(const dsk_conv_desc* d, const void* in, const void* w, const float* bias,
                               const float* chan_bias, const void* residual, void* out, void* stream) {
  DSK_REQUIRE(d == nullptr || !d->circular, "dsk_conv_fwd: circular padding on the tcgen05 path needs dsk_conv_fwd_circ (padded copy)");
  return conv_fwd_tc_impl(d, in, w, bias, chan_bias, residual, out, nullptr, nullptr, stream);
}

extern "C" int dsk_conv_fwd_ffma(const dsk_conv_desc*, const void*, const void*, const float*, const float*, const void*, void*, void*);

// bytes of the halo-padded input copy a circular convolution on the tcgen05 path reads (0: CUDA-core kernel, wraps in place)
extern "C" int64_t dsk_conv_pad_ws_bytes(const dsk_conv_desc* d) {
  if (d == nullptr || d->circular != 1 || d->w_dtype == DSK_F32) return 0;     // circular == 2: the input is already padded
  const int iD = (d->up2 && d->ndim == 3) ? d->D / 2 : d->D, iH = d->up2 ? d->H / 2 : d->H, iW = d->up2 ? d->W / 2 : d->W;
  const int pd = d->ndim == 3 ? 1 : 0;
  return (int64_t)d->B * (iD + 2 * pd) * (iH + 2) * (iW + 2) * d->Cin * (d->in_dtype == DSK_SPLIT_F16 ? 4 : 2);
}

extern "C" int dsk_conv_fwd_circ(const dsk_conv_desc* d, const void* in, const void* w, const float* bias, const float* chan_bias,
                                 const void* residual, void* out, void* stats, void* pad_ws, void* stream) {
  DSK_REQUIRE(d != nullptr && d->circular, "dsk_conv_fwd_circ: descriptor is not circular");
  if (d->w_dtype == DSK_F32) {
    DSK_REQUIRE(d->circular == 1, "dsk_conv_fwd_circ: a pre-padded input (circular = 2) needs the tcgen05 path");
    DSK_REQUIRE(stats == nullptr, "dsk_conv_fwd_circ: fused statistics need the tcgen05 path");
    return dsk_conv_fwd_ffma(d, in, w, bias, chan_bias, residual, out, stream);
  }
  DSK_REQUIRE(pad_ws != nullptr || d->circular == 2, "dsk_conv_fwd_circ: the tcgen05 path needs pad_ws (dsk_conv_pad_ws_bytes)");
  DSK_REQUIRE(stats == nullptr || dsk_conv_stats_supported(d), "dsk_conv_fwd_circ: this convolution cannot emit fused statistics");
  return conv_fwd_tc_impl(d, in, w, bias, chan_bias, residual, out, (float2*)stats, pad_ws, stream);
}

extern "C" int dsk_conv_fwd_stats(const dsk_conv_desc* d, const void* in, const void* w, const float* bias, const float* chan_bias,
                                  const void* residual, void* out, void* stats, void* stream) {
  DSK_REQUIRE(stats != nullptr && dsk_conv_stats_supported(d), "dsk_conv_fwd_stats: this convolution cannot emit fused statistics");
  DSK_REQUIRE(!d->circular, "dsk_conv_fwd_stats: circular padding needs dsk_conv_fwd_circ");
  return conv_fwd_tc_impl(d, in, w, bias, chan_bias, residual, out, (float2*)stats, nullptr, stream);
}

static int conv_fwd_tc_impl(const dsk_conv_desc* d, const void* in, const void* w, const float* bias, const float* chan_bias,
                            const void* residual, void* out, float2* stats, void* pad_ws, void* stream) {
  DSK_REQUIRE(d && in && w && out, "dsk_conv_fwd(tc): null pointer");
  const bool few_out = d->Cout <= 16 && !d->up2 && chan_bias == nullptr && residual == nullptr;
  if (d->ksize != 3 || (d->out_nchw_f32 && !few_out) || !tc_formats_ok(d) || d->Cin % 64 != 0 || (d->Cout % 64 != 0 && !few_out)) {
    set_error("dsk_conv_fwd: the tcgen05 path takes k=3, 16-bit / split-fp16 operands of one format, Cin %% 64 == 0, Cout %% 64 == 0 "
              "(got k=%d Cin=%d Cout=%d up2=%d in=%d w=%d out=%d nchw=%d)", d->ksize, d->Cin, d->Cout, d->up2, d->in_dtype, d->w_dtype,
              d->out_dtype, d->out_nchw_f32);
    return DSK_ERR_UNSUPPORTED;
  }
  const bool a_split = d->in_dtype == DSK_SPLIT_F16, w_split = d->w_dtype == DSK_SPLIT_F16;
  const int a_ld = d->Cin * (a_split ? 2 : 1), w_ld = d->Cin * (w_split ? 2 : 1);     // channel-row lengths of the two tensor maps
  const int fmt16 = d->in_dtype == DSK_BF16 ? DSK_BF16 : DSK_F16;
  DSK_REQUIRE((d->ndim == 2 && d->D == 1) || d->ndim == 3, "dsk_conv_fwd(tc): bad ndim/D");
  // circular padding: TMA boxes cannot wrap, so the kernels read a halo-padded copy (one extra pass over the input) and the
  // patch coordinates are shifted into it; nothing else in the kernels changes (no OOB fill is ever hit inside the image).
  const int iD = (d->up2 && d->ndim == 3) ? d->D / 2 : d->D, iH = d->up2 ? d->H / 2 : d->H, iW = d->up2 ? d->W / 2 : d->W;
  const int pad_hw = d->circular ? 1 : 0, pad_d = (d->circular && d->ndim == 3) ? 1 : 0;
  if (d->circular == 1) {
    DSK_REQUIRE(pad_ws != nullptr, "dsk_conv_fwd(tc): circular padding needs pad_ws");
    const int rc = pad_circular_launch(in, pad_ws, d->B, d->ndim == 3 ? iD : 1, iH, iW, a_ld, d->ndim, DSK_BF16, as_stream(stream));
    if (rc != DSK_OK) return rc;
    in = pad_ws;
  }
  if (few_out) {     // in-plane taps as the N dimension (convout_tc.cu); DSK_CONVOUT_OLD=1 keeps the N = 16 tile for A/B runs
    static const int old_path = [] { const char* e = getenv("DSK_CONVOUT_OLD"); return e ? atoi(e) : 0; }();
    if (!old_path) {
      const int rc = convout_tc_dispatch(d, in, w, bias, out, as_stream(stream), d->circular);
      if (rc != DSK_ERR_UNSUPPORTED) return rc;
    }
  }
  EncodeTiledFn encode = get_encode();
  DSK_REQUIRE(encode != nullptr, "dsk_conv_fwd(tc): cuTensorMapEncodeTiled is unavailable");
  // 2-D: the batch is the plane axis (no depth taps); 3-D: planes = D with zero padding per sample
  const int KD = d->ndim == 3 ? 3 : 1;
  // up2: D/H/W in the descriptor are the OUTPUT size; the kernel tiles the INPUT (half-size) grid
  const int planes = d->ndim == 3 ? iD : d->B;
  const int batch = d->ndim == 3 ? d->B : 1;
  const int tW = iW + 2 * pad_hw, tH = iH + 2 * pad_hw, tP = planes + 2 * pad_d;
  // merged depth taps (conv_tc2_kernel<.., T = 2>): 3-D, 64-channel output tiles, plain 16-bit operands, CTA pairs side by side in w,
  // rows in whole 32-row tiles.  DSK_CONV_T2=0 keeps the N = 64 pair kernel (A/B measurements).
  static const int allow_t2 = [] { const char* e = getenv("DSK_CONV_T2"); return e ? atoi(e) : 1; }();
  const bool t2 = allow_t2 && !few_out && !d->up2 && KD == 3 && !a_split && !w_split && d->Cout % 128 != 0 && tc_pair_eligible(d) &&
                  iH % (2 * TC_BH) == 0;
  CUtensorMap ta, tw;
  {
    cuuint64_t dims[5] = {(cuuint64_t)a_ld, (cuuint64_t)tW, (cuuint64_t)tH, (cuuint64_t)tP, (cuuint64_t)batch};
    cuuint64_t strides[4] = {(cuuint64_t)a_ld * 2, (cuuint64_t)tW * a_ld * 2, (cuuint64_t)tH * tW * a_ld * 2,
                             (cuuint64_t)tP * tH * tW * a_ld * 2};
    cuuint32_t box[5] = {64, TC_PW, (cuuint32_t)(t2 ? 2 * TC_BH + 2 : TC_PH), 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&ta, tmap_h16(fmt16), 5, const_cast<void*>(in), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DSK_REQUIRE(r == CUDA_SUCCESS, "dsk_conv_fwd(tc): activation tensor map failed (CUresult %d)", (int)r);
  }
  const int nphase = d->up2 ? (KD == 3 ? 8 : 4) : 1;
  const int ntaps = d->up2 ? nphase * (KD == 3 ? 8 : 4) : KD * 9;     // weight rows: [phase][tap][Cout] or [tap][Cout]
  // fully split operands: four accumulator sets of P * 64 columns fill the TMEM -> 64-channel tiles only
  const int n_tile = few_out ? 16 : ((d->Cout % 128 == 0 && !(a_split && w_split)) ? 128 : 64);
  const int w_rows = few_out ? 16 : d->Cout;                          // few_out weights are zero-padded to 16 output channels
  {
    cuuint64_t dims[2] = {(cuuint64_t)w_ld, (cuuint64_t)ntaps * w_rows};
    cuuint64_t strides[1] = {(cuuint64_t)w_ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)n_tile};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&tw, tmap_h16(fmt16), 2, const_cast<void*>(w), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DSK_REQUIRE(r == CUDA_SUCCESS, "dsk_conv_fwd(tc): weight tensor map failed (CUresult %d)", (int)r);
  }
  TcParams p;
  p.B = batch; p.D = planes; p.H = iH; p.W = iW; p.Cin = d->Cin; p.Cout = w_rows; p.KD = KD;
  p.cout_real = d->Cout;
  p.out_nchw = d->out_nchw_f32 ? (float*)out : nullptr;
  constexpr int P = 2;
  p.nphase = nphase;
  p.tiles_w = (iW + TC_BW - 1) / TC_BW;
  p.bh = t2 ? 2 * TC_BH : TC_BH;
  p.tiles_h = (iH + p.bh - 1) / p.bh;
  p.groups_d = (planes + P - 1) / P;
  p.n_tiles = w_rows / n_tile;
  p.total_tiles = p.tiles_w * p.tiles_h * p.groups_d * batch * p.n_tiles * nphase;
  p.bias = bias; p.chan_bias = chan_bias;
  p.residual = (const uint16_t*)residual; p.out = (uint16_t*)out;
  p.f16 = fmt16 == DSK_F16 ? 1 : 0;
  p.out_f32 = (!d->out_nchw_f32 && d->out_dtype == DSK_F32) ? 1 : 0;
  p.res_f32 = d->res_dtype == DSK_RES_F32 ? 1 : (d->res_dtype == DSK_RES_SAME ? p.out_f32 : 0);
  p.vparts = a_split ? (w_split ? 3 : 2) : 1;
  p.a_lo_off = d->Cin; p.w_lo_off = d->Cin;
  p.nsets = a_split ? (w_split ? 4 : 2) : 1;
  static const int dbg = [] { const char* e = getenv("DSK_CONV_DBG"); return e ? atoi(e) : 0; }();
  p.dbg = dbg;
  p.planes_per_sample = d->ndim == 3 ? iD : 1;
  p.stats = stats; p.samples = d->B;
  p.pad_hw = pad_hw; p.pad_d = pad_d;
  p.pair_d = (!few_out && tc_pair_d_eligible(d, stats != nullptr)) ? 1 : 0;
  {
    const uint32_t dv[5] = {(uint32_t)nphase, (uint32_t)(p.pair_d ? p.tiles_w : (p.tiles_w >> 1 ? p.tiles_w >> 1 : 1)), (uint32_t)p.tiles_h,
                            (uint32_t)(p.pair_d ? (p.groups_d >> 1 ? p.groups_d >> 1 : 1) : p.groups_d), (uint32_t)batch};
    for (int i = 0; i < 5; ++i) { const FastDivHost f(dv[i]); p.dv_m[i] = f.m; p.dv_s[i] = f.s; }
  }
  cudaStream_t st = as_stream(stream);
  // cta_group::2 (CTA pairs): needs an even number of w-tiles (the pair sits side by side in w).  DSK_CONV_CG=1 forces
  // the single-CTA kernel (A/B measurements).
  if (!few_out && (tc_pair_eligible(d) || p.pair_d)) {
    if (stats != nullptr && p.total_tiles / 2 < DSK_NUM_SMS / 2) {
      // fewer CTAs than slots: the slots of the CTAs that do not exist read as zero
      cudaError_t e = cudaMemsetAsync(stats, 0, (size_t)d->B * TC_STAT_SLOTS * d->Cout * sizeof(float2), st);
      DSK_REQUIRE(e == cudaSuccess, "dsk_conv_fwd_stats: memset failed: %s", cudaGetErrorString(e));
    }
    CUtensorMap tw2;
    cuuint64_t dims[2] = {(cuuint64_t)w_ld, (cuuint64_t)ntaps * w_rows};
    cuuint64_t strides[1] = {(cuuint64_t)w_ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(n_tile / 2)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&tw2, tmap_h16(fmt16), 2, const_cast<void*>(w), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DSK_REQUIRE(r == CUDA_SUCCESS, "dsk_conv_fwd(tc2): weight tensor map failed (CUresult %d)", (int)r);
    // smem: 7 patches (161 KB) + 8 half-taps (N=64: 32 KB, N=128: 64 KB... 6 patches then)
    if (d->up2) {
      if (n_tile == 64) return launch_tc2<64, P, 7, 8, true>(ta, tw2, p, st);
      return launch_tc2<128, P, 6, 8, true>(ta, tw2, p, st);
    }
    if (t2) return launch_tc2<64, P, 3, 8, false, 2>(ta, tw2, p, st);   // 3 patches of 43 KB + 8 whole taps of 8 KB = 196 KB
    if (n_tile == 64) return launch_tc2<64, P, 7, 8, false>(ta, tw2, p, st);
    return launch_tc2<128, P, 6, 8, false>(ta, tw2, p, st);
  }
  // smem: N_TILE=64: 6 patches (138 KB) + 8 taps x 8 KB (64 KB) = 202 KB; N_TILE=128: 6 patches + 5 x 16 KB = 218 KB
  if (few_out) return launch_tc<16, P, 6, 8, false>(ta, tw, p, st);
  if (d->up2) {
    if (n_tile == 64) return launch_tc<64, P, 6, 8, true>(ta, tw, p, st);
    return launch_tc<128, P, 6, 5, true>(ta, tw, p, st);
  }
  if (n_tile == 64) return launch_tc<64, P, 6, 8, false>(ta, tw, p, st);
  return launch_tc<128, P, 6, 5, false>(ta, tw, p, st);
}
