// convin_tc.cu -- the FIRST convolution of the U-Nets (Cin <= 4 -> C, k = 3; reference nets/punetg.py:203-208,
// nets/adm.py:183-188) on the tensor cores.
//
// With Cin = 1..3 the whole receptive field of an output pixel is K = taps * Cin <= 64 numbers: the im2col row of a pixel
// is ONE 128-byte shared-memory row.  Four builder warps gather those rows (the input tensor is tiny and L1/L2 resident)
// straight into the K-major SWIZZLE_128B layout, one tcgen05 MMA group (M = 128 pixels, N = Cout, K = 64) turns a tile
// into accumulators, and two sets of four epilogue warps (one per accumulator buffer, alternating tiles) stream the
// C-channel result out through a shared-memory stage so that every store instruction writes whole 128-byte lines -- the kernel
// is bound by the write of the output tensor (the CUDA-core version was bound by shared-memory weight reads at 1/10 of that).
// Input fp32 or 16-bit (rounded to the operand format while gathering), output 16-bit or fp32 (the fp32-storage modes).
#include <stdlib.h>

#include "tc_common.cuh"

namespace dsk {

constexpr int CI_BW = 8, CI_BH = 16;               // 128 output pixels per tile
constexpr int CI_NG = 2;                           // builder groups of 4 warps: CI_NG tiles are gathered concurrently
constexpr int CI_BUILD = CI_NG * 4;                // warps 0..CI_BUILD-1: im2col builders
constexpr int CI_EPI = 8;                          // epilogue warps: set (w >> 2) drains accumulator buffer (tile seq & 1)
constexpr int CI_THREADS = (CI_BUILD + 1 + CI_EPI) * 32;  // + warp CI_BUILD: MMA / TMEM owner
constexpr int CI_STG_BYTES = 32 * 128;             // per epilogue warp: 32 rows x 128 B (one 32-channel fp32 group)
constexpr int CI_PW = 12;                          // patch row pitch in values (10 used)
constexpr int CI_PATCH_WORDS = 3 * 18 * CI_PW;     // 648 32-bit containers (fp32 bits, or a zero-extended 16-bit value)
constexpr int CI_PATCH_LD = (CI_PATCH_WORDS + 127) / 128;   // loads per thread of a 128-thread builder group
constexpr int CI_STAGES = 6;
constexpr int CI_A_BYTES = 128 * 128;

struct CiParams {
  const uint16_t* x;        // channels-last [B, D, H, W, Cin], 16-bit format f16 ? half : bfloat16 (fp32 when in_f32)
  int in_f32, out_f32;      // fp32 input (rounded to the 16-bit operand format in the gather) / fp32 output
  int split;                // fp32 input only: EXACT fp32-class product on fp16 operands -- the im2col row becomes
                            // [x_hi (K) | x_lo (K) | x_hi (K)] and the weight row [w_hi | w_hi | w_lo] (3K <= 128: two 128-byte
                            // K chunks), i.e. x_hi w_hi + x_lo w_hi + x_hi w_lo in one accumulator chain of <= 8 MMAs
  int nstages;              // A ring depth (6; 4 in split mode: stages are twice as large)
  int staged;               // CIN == 1: the builder group loads the tile's halo'd input patch (3 planes x 18 x 12 values) cooperatively
                            // -- ~5 coalesced loads per thread, requested ONE TILE AHEAD -- into shared memory and every thread reads
                            // its 27 taps from there.  (27 predicated global loads per thread and tile, issued when the tile starts,
                            // left the two builder groups waiting on memory latency: 305 us for a 100 us write.)
  const float* w;           // packed fp32 [taps][Cin][Cout]
  const float* bias;
  uint16_t* out;            // channels-last [B, D, H, W, Cout]
  int f16;
  int B, D, H, W, Cin, Cout, ndim;
  int tiles_w, tiles_h, total_tiles;
  int circ;                 // circular padding: the gather wraps instead of zero-filling (CircularConv, commonlayers.py:918-1032)
};

template <int CIN, bool D3>
__global__ void __launch_bounds__(CI_THREADS, 1) convin_tc_kernel(const CiParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int nch = p.split ? 2 : 1;                           // 128-byte K chunks per im2col row
  const int NST = p.nstages;
  const size_t a_stage = (size_t)nch * CI_A_BYTES;
  uint8_t* sA = smem;                                        // NST x nch x [128 rows x 128 B]
  uint8_t* sB = smem + (size_t)NST * a_stage;                // nch x [Cout rows x 128 B]
  uint8_t* sStg = sB + (size_t)nch * p.Cout * 128;           // CI_EPI x [32 rows x 128 B] store stages
  float* sBias = reinterpret_cast<float*>(sStg + (size_t)CI_EPI * CI_STG_BYTES);   // [Cout]
  __shared__ uint32_t patch_s[CI_NG][2][CI_PATCH_WORDS];     // per builder group: double-buffered input patch (staged mode)
  __shared__ uint64_t full_a[CI_STAGES], empty_a[CI_STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.Cout;
  const uint32_t tmem_cols = N <= 64 ? 128u : (N <= 128 ? 256u : 512u);   // 2 x N, power of two
  constexpr int TAPS = D3 ? 27 : 9;
  constexpr int K = TAPS * CIN;                              // <= 64 (host-checked)

  if (threadIdx.x == 0) {
    for (int i = 0; i < CI_STAGES; ++i) { mbar_init(&full_a[i], 4); mbar_init(&empty_a[i], 1); }   // the first p.nstages are used
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == CI_BUILD) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // weights -> bf16 K-major SWIZZLE_128B rows: B[co][k = tap*CIN + ci] = w[k][co], zero beyond K
  for (int i = threadIdx.x; i < N * 8 * nch; i += CI_THREADS) {
    const int j = i & 7, co = (i >> 3) % N, c = (i >> 3) / N;
    uint32_t h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float ab[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int kk = c * 64 + j * 8 + 2 * e + u;           // position in the (virtual) K axis
        float val = 0.0f;
        if (!p.split) {
          if (kk < K) val = p.w[(int64_t)kk * N + co];
        } else if (kk < 3 * K) {
          const int k0 = kk < K ? kk : (kk < 2 * K ? kk - K : kk - 2 * K);
          const float wv = p.w[(int64_t)k0 * N + co];
          val = kk < 2 * K ? wv : wv - __half2float(__float2half_rn(wv));   // [w_hi | w_hi | w_lo] (pack rounds w_hi below)
        }
        ab[u] = val;
      }
      h[e] = pack_h2(ab[0], ab[1], p.f16);
    }
    *reinterpret_cast<uint4*>(sB + (size_t)c * N * 128 + co * 128 + ((j ^ (co & 7)) << 4)) = *reinterpret_cast<const uint4*>(h);
  }
  for (int i = threadIdx.x; i < N; i += CI_THREADS) sBias[i] = p.bias != nullptr ? p.bias[i] : 0.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  auto coord = [&](int t, int& w0, int& h0, int& d, int& b) {
    w0 = (t % p.tiles_w) * CI_BW; t /= p.tiles_w;
    h0 = (t % p.tiles_h) * CI_BH; t /= p.tiles_h;
    d = t % p.D;
    b = t / p.D;
  };

  if (warp < CI_BUILD) {
    // ===================== im2col builders: group g gathers tiles seq = g, g + CI_NG, ...; thread = one pixel row ===========
    const int grp = warp >> 2;
    const int r = threadIdx.x & 127, line = r >> 3, wp = r & 7;
    constexpr int kdn = D3 ? 3 : 1;
    uint32_t seq = grp;
    // ---- staged mode (CIN == 1) ----
    uint32_t pre[CI_PATCH_LD];                                 // this thread's share of the NEXT tile's patch, in flight
    auto patch_request = [&](int t_) {                         // issue the loads of tile t_'s patch (zero / wrapped outside the image)
      int w0_, h0_, d_, b_;
      coord(t_, w0_, h0_, d_, b_);
#pragma unroll
      for (int i = 0; i < CI_PATCH_LD; ++i) {
        const int idx = r + i * 128;
        pre[i] = 0;
        if (idx < kdn * 18 * CI_PW) {
          const int pcol = idx % CI_PW, prow = idx / CI_PW, kd = prow / 18, hh = prow - kd * 18;
          int z = D3 ? d_ + kd - 1 : 0, y = h0_ + hh - 1, x = w0_ + pcol - 1;
          bool ok = pcol < 10;
          if (p.circ) {
            z = z < 0 ? z + p.D : (z >= p.D ? z - p.D : z);
            y = y < 0 ? y + p.H : (y >= p.H ? y - p.H : y);
            x = x < 0 ? x + p.W : (x >= p.W ? x - p.W : x);
            ok = ok && (unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W;      // rows / columns of ragged tiles far outside
          } else {
            ok = ok && (unsigned)z < (unsigned)p.D && (unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W;
          }
          if (ok) {
            const int64_t off = (((int64_t)b_ * p.D + z) * p.H + y) * p.W + x;
            pre[i] = p.in_f32 ? __float_as_uint(reinterpret_cast<const float*>(p.x)[off]) : (uint32_t)p.x[off];
          }
        }
      }
    };
    auto patch_commit = [&](uint32_t buf) {                    // registers -> shared memory, then the group meets
#pragma unroll
      for (int i = 0; i < CI_PATCH_LD; ++i) {
        const int idx = r + i * 128;
        if (idx < kdn * 18 * CI_PW) patch_s[grp][buf][idx] = pre[i];
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
    };
    uint32_t pn = 0;
    if (p.staged) {
      const int t0 = blockIdx.x + grp * gridDim.x;
      if (t0 < p.total_tiles) { patch_request(t0); patch_commit(0); }
    }
    for (int t = blockIdx.x + grp * gridDim.x; t < p.total_tiles; t += CI_NG * gridDim.x, seq += CI_NG, ++pn) {
      int w0, h0, d, b;
      coord(t, w0, h0, d, b);
      const int h = h0 + line, w = w0 + wp;
      const uint32_t slot = seq % NST, ph = (seq / NST) & 1;
      if (p.staged) {
        if constexpr (CIN == 1) {
          const int tn = t + CI_NG * gridDim.x;
          const bool more = tn < p.total_tiles;                // group-uniform
          if (more) patch_request(tn);                         // next tile's loads fly while this tile is packed
          const uint32_t* pat = patch_s[grp][pn & 1];
          uint32_t tv[K];
#pragma unroll
          for (int kd = 0; kd < kdn; ++kd)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                const uint32_t raw = pat[(kd * 18 + line + kh) * CI_PW + wp + kw];
                const int k = (kd * 3 + kh) * 3 + kw;
                if (p.in_f32) {
                  const float xf = __uint_as_float(raw);
                  const uint32_t hi = pack_h1(xf, p.f16);
                  tv[k] = hi | ((uint32_t)pack_h1(xf - unpack_h1((uint16_t)hi, p.f16), p.f16) << 16);
                } else {
                  tv[k] = raw;
                }
              }
          mbar_wait(&empty_a[slot], ph ^ 1);
          uint8_t* row = sA + (size_t)slot * a_stage + r * 128;
          auto elem = [&](int i, bool split) -> uint32_t {
            if (!split) return i < K ? (tv[i < K ? i : 0] & 0xffffu) : 0u;
            if (i < K) return tv[i < K ? i : 0] & 0xffffu;
            if (i < 2 * K) return tv[(i - K) < K ? (i - K) : 0] >> 16;
            if (i < 3 * K) return tv[(i - 2 * K) < K ? (i - 2 * K) : 0] & 0xffffu;
            return 0u;
          };
          if (!p.split) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint4 piece;
              piece.x = elem(j * 8, false) | (elem(j * 8 + 1, false) << 16);
              piece.y = elem(j * 8 + 2, false) | (elem(j * 8 + 3, false) << 16);
              piece.z = elem(j * 8 + 4, false) | (elem(j * 8 + 5, false) << 16);
              piece.w = elem(j * 8 + 6, false) | (elem(j * 8 + 7, false) << 16);
              *reinterpret_cast<uint4*>(row + ((j ^ (r & 7)) << 4)) = piece;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              uint4 piece;
              piece.x = elem(j * 8, true) | (elem(j * 8 + 1, true) << 16);
              piece.y = elem(j * 8 + 2, true) | (elem(j * 8 + 3, true) << 16);
              piece.z = elem(j * 8 + 4, true) | (elem(j * 8 + 5, true) << 16);
              piece.w = elem(j * 8 + 6, true) | (elem(j * 8 + 7, true) << 16);
              *reinterpret_cast<uint4*>(row + (j >> 3) * CI_A_BYTES + (((j & 7) ^ (r & 7)) << 4)) = piece;
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_a[slot]);
          // the other buffer was last read one tile ago, and every thread of the group has passed the barrier since
          if (more) patch_commit((pn + 1) & 1);
          continue;
        }
      }
      // gather the receptive field first (loads in flight while waiting for the slot); validity factorises per axis
      bool dv[3], hv[3], wv[3];
      const int sH = p.W * CIN, sD = p.H * sH;
      int od[3], oh[3], ow[3];                                  // element offsets of the three taps of each axis from the centre
      const bool pv = h < p.H && w < p.W;                       // ragged tiles: pixels outside the image gather nothing
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        int zd = d + k - 1, zh = h + k - 1, zw = w + k - 1;
        if (p.circ) {
          zd = zd < 0 ? zd + p.D : (zd >= p.D ? zd - p.D : zd);
          zh = zh < 0 ? zh + p.H : (zh >= p.H ? zh - p.H : zh);
          zw = zw < 0 ? zw + p.W : (zw >= p.W ? zw - p.W : zw);
          dv[k] = pv && (p.ndim == 3 || k == 0);
          hv[k] = pv;
          wv[k] = pv;
        } else {
          dv[k] = p.ndim == 3 ? (unsigned)zd < (unsigned)p.D : k == 0;
          hv[k] = (unsigned)zh < (unsigned)p.H;
          wv[k] = (unsigned)zw < (unsigned)p.W;
        }
        od[k] = p.ndim == 3 ? (zd - d) * sD : 0;
        oh[k] = (zh - h) * sH;
        ow[k] = (zw - w) * CIN;
      }
      const int64_t ctr_off = ((((int64_t)b * p.D + d) * p.H + h) * p.W + w) * CIN;
      const uint16_t* ctr = p.x + ctr_off;
      const float* ctr32 = reinterpret_cast<const float*>(p.x) + ctr_off;
      // one 32-bit register per tap holding its 16-bit operand(s): hi in the low half; split mode: lo = x - fp16(x) in the high
      // half.  Every index below is a compile-time constant after unrolling, and the 16-byte pieces are assembled from these
      // registers with shifts / PRMT -- a uint16 array reinterpreted as uint4 goes through local memory (the first version did).
      uint32_t tv[K];
#pragma unroll
      for (int k = 0; k < K; ++k) tv[k] = 0;                   // zero bits = 0.0 in both 16-bit formats
#pragma unroll
      for (int kd = 0; kd < 3; ++kd) {
        if (kd >= kdn) break;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const bool ok = dv[kd] && hv[kh] && wv[kw];
            const int so = od[kd] + oh[kh] + ow[kw];
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci)
              if (ok) {
                const int k = ((kd * 3 + kh) * 3 + kw) * CIN + ci;
                if (p.in_f32) {
                  const float xf = ctr32[so + ci];
                  const uint32_t hi = pack_h1(xf, p.f16);
                  tv[k] = hi | ((uint32_t)pack_h1(xf - unpack_h1((uint16_t)hi, p.f16), p.f16) << 16);
                } else {
                  tv[k] = ctr[so + ci];
                }
              }
          }
      }
      mbar_wait(&empty_a[slot], ph ^ 1);
      uint8_t* row = sA + (size_t)slot * a_stage + r * 128;
      // element at virtual K position i of the row: plain: x_hi[i]; split: [x_hi (K) | x_lo (K) | x_hi (K)]; zero beyond
      auto elem = [&](int i, bool split) -> uint32_t {
        if (!split) return i < K ? (tv[i < K ? i : 0] & 0xffffu) : 0u;
        if (i < K) return tv[i < K ? i : 0] & 0xffffu;
        if (i < 2 * K) return tv[(i - K) < K ? (i - K) : 0] >> 16;
        if (i < 3 * K) return tv[(i - 2 * K) < K ? (i - 2 * K) : 0] & 0xffffu;
        return 0u;
      };
      if (!p.split) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 piece;
          piece.x = elem(j * 8, false) | (elem(j * 8 + 1, false) << 16);
          piece.y = elem(j * 8 + 2, false) | (elem(j * 8 + 3, false) << 16);
          piece.z = elem(j * 8 + 4, false) | (elem(j * 8 + 5, false) << 16);
          piece.w = elem(j * 8 + 6, false) | (elem(j * 8 + 7, false) << 16);
          *reinterpret_cast<uint4*>(row + ((j ^ (r & 7)) << 4)) = piece;
        }
      } else if (3 * K <= 128) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          uint4 piece;
          piece.x = elem(j * 8, true) | (elem(j * 8 + 1, true) << 16);
          piece.y = elem(j * 8 + 2, true) | (elem(j * 8 + 3, true) << 16);
          piece.z = elem(j * 8 + 4, true) | (elem(j * 8 + 5, true) << 16);
          piece.w = elem(j * 8 + 6, true) | (elem(j * 8 + 7, true) << 16);
          *reinterpret_cast<uint4*>(row + (j >> 3) * CI_A_BYTES + (((j & 7) ^ (r & 7)) << 4)) = piece;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_a[slot]);
    }
  } else if (warp == CI_BUILD) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc_h16(N, 128, p.f16);
    constexpr uint32_t HI = umma_desc_hi(1024);
    const uint32_t b_lo = umma_desc_lo(smem_u32(sB));
    uint32_t seq = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++seq) {
      const uint32_t as = seq & 1, aph = (seq >> 1) & 1;
      const uint32_t slot = seq % NST;
      mbar_wait(&acc_empty[as], aph ^ 1);
      mbar_wait(&full_a[slot], (seq / NST) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_lo = umma_desc_lo(smem_u32(sA + (size_t)slot * a_stage));
      const int Kv = p.split ? 3 * K : K;                     // virtual K
      if (elect_one_sync()) {
#pragma unroll
        for (int k8 = 0; k8 < 8; ++k8) {
          const uint32_t c = k8 >> 2, k4 = k8 & 3;
          if (k8 * 16 < Kv)
            umma_bf16(tmem_base + as * N, umma_desc64(a_lo + c * (CI_A_BYTES >> 4) + k4 * 2, HI),
                      umma_desc64(b_lo + c * (uint32_t)((N * 128) >> 4) + k4 * 2, HI), idesc, k8 != 0);
        }
        umma_commit(&empty_a[slot]);
        umma_commit(&acc_full[as]);
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue =====================
    // warp e = warp - CI_BUILD - 1: TMEM lane quadrant q = warp & 3 (hardware rule), set es = e >> 2 takes the tiles whose
    // sequence number has parity es (= accumulator buffer es).  A thread holds one pixel ROW of the tile; the 32 x 32 block of a
    // channel group goes through the warp's stage (XOR-swizzled 16-byte slots) and leaves as whole lines: fp32: 8 lanes per
    // row, 4 rows per instruction; 16-bit: 4 lanes per row, 8 rows per instruction.
    const int e = warp - CI_BUILD - 1, q = warp & 3, es = e >> 2;
    uint4* stg = reinterpret_cast<uint4*>(sStg + (size_t)e * CI_STG_BYTES);
    for (int t = blockIdx.x + es * gridDim.x, seq = es; t < p.total_tiles; t += 2 * gridDim.x, seq += 2) {
      int w0, h0, d, b;
      coord(t, w0, h0, d, b);
      const uint32_t as = es, aph = (seq >> 1) & 1;
      mbar_wait(&acc_full[as], aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int64_t plane = ((int64_t)b * p.D + d) * p.H;
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        DSK_TMEM_LD_X32(v, tmem_base + as * N + c0 + ((uint32_t)(q * 32) << 16));
        float f[32];
        const float4* b4 = reinterpret_cast<const float4*>(sBias + c0);
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          const float4 bb = b4[e4];
          f[4 * e4] = __uint_as_float(v[4 * e4]) + bb.x; f[4 * e4 + 1] = __uint_as_float(v[4 * e4 + 1]) + bb.y;
          f[4 * e4 + 2] = __uint_as_float(v[4 * e4 + 2]) + bb.z; f[4 * e4 + 3] = __uint_as_float(v[4 * e4 + 3]) + bb.w;
        }
        if (p.out_f32) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            stg[lane * 8 + (g ^ (lane & 7))] = make_uint4(__float_as_uint(f[4 * g]), __float_as_uint(f[4 * g + 1]),
                                                          __float_as_uint(f[4 * g + 2]), __float_as_uint(f[4 * g + 3]));
          __syncwarp();
          // row r = 4k + lane / 8 of the warp = pixel (line q*4 + k/2, column 4 (k & 1) + lane / 8) of the tile: ONE 64-bit base per
          // tile and 32-bit row offsets with compile-time structure (eight hoisted 64-bit addresses spilled to local memory)
          const int j = lane & 7, l3 = lane >> 3;
          float* ob = reinterpret_cast<float*>(p.out) + ((plane + h0 + q * 4) * p.W + w0 + l3) * N + c0 + j * 4;
          int wn = p.W * N, n4 = 4 * N;
          asm volatile("" : "+r"(wn), "+r"(n4));                // opaque per pass: keeps the row offsets out of the loop-invariant set
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int r = 4 * k + l3;
            const uint4 val = stg[r * 8 + (j ^ (r & 7))];
            if (h0 + q * 4 + (k >> 1) < p.H && w0 + 4 * (k & 1) + l3 < p.W)
              *reinterpret_cast<uint4*>(ob + (k >> 1) * wn + (k & 1) * n4) = val;
          }
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o;
            uint32_t* oh = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int k = 0; k < 4; ++k) oh[k] = pack_h2(f[g * 8 + 2 * k], f[g * 8 + 2 * k + 1], p.f16);
            stg[lane * 4 + (g ^ ((lane >> 1) & 3))] = o;
          }
          __syncwarp();
          const int j = lane & 3, l2 = lane >> 2;               // row r = 8k + lane / 4 = pixel (line q*4 + k, column lane / 4)
          uint16_t* ob = p.out + ((plane + h0 + q * 4) * p.W + w0 + l2) * N + c0 + j * 8;
          int wn = p.W * N;
          asm volatile("" : "+r"(wn));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int r = 8 * k + l2;
            const uint4 val = stg[r * 4 + (j ^ ((r >> 1) & 3))];
            if (h0 + q * 4 + k < p.H && w0 + l2 < p.W) *reinterpret_cast<uint4*>(ob + k * wn) = val;
          }
        }
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == CI_BUILD) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
}

template <int CIN, bool D3>
static int launch_convin(const CiParams& p, cudaStream_t st) {
  const int nch = p.split ? 2 : 1;
  const size_t smem = (size_t)p.nstages * nch * CI_A_BYTES + (size_t)nch * p.Cout * 128 + (size_t)CI_EPI * CI_STG_BYTES +
                      (size_t)p.Cout * 4 + 1024;
  cudaError_t e = cudaFuncSetAttribute(convin_tc_kernel<CIN, D3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("convin_tc: cudaFuncSetAttribute(%zu B smem): %s", smem, cudaGetErrorString(e)); return DSK_ERR_CUDA; }
  const int grid = p.total_tiles < DSK_NUM_SMS ? p.total_tiles : DSK_NUM_SMS;
  DSK_LAUNCH((convin_tc_kernel<CIN, D3>), grid, CI_THREADS, smem, st, p);
  return DSK_OK;
}

// returns DSK_ERR_UNSUPPORTED for shapes it does not take (the caller keeps the CUDA-core few-channel kernel)
int convin_tc_dispatch(const dsk_conv_desc* d, const void* in, const void* w, const float* bias, void* out, cudaStream_t st) {
  static const int old_path = [] { const char* e = getenv("DSK_CONVIN_OLD"); return e ? atoi(e) : 0; }();
  const int taps = d->ndim == 3 ? 27 : 9;
  // formats: 16-bit in -> the same 16-bit out (the 16-bit modes), or -- descriptor field `operand16` = DSK_BF16 | DSK_F16 naming
  // the tensor-core operand format -- fp32 in -> fp32 out (the fp32-storage modes with 16-bit operands)
  const bool h16 = is_h16(d->in_dtype) && d->out_dtype == d->in_dtype;
  const bool split = d->operand16 == DSK_SPLIT_F16 && 3 * taps * d->Cin <= 128;      // exact: [x_hi | x_lo | x_hi] . [w_hi | w_hi | w_lo]
  const bool f32 = d->in_dtype == DSK_F32 && d->out_dtype == DSK_F32 && (is_h16(d->operand16) || split);
  if (old_path || d->ksize != 3 || d->up2 || d->out_nchw_f32 || !(h16 || f32) || d->w_dtype != DSK_F32 ||
      d->Cin > 4 || taps * d->Cin > 64 || (d->Cout != 64 && d->Cout != 128 && d->Cout != 256))
    return DSK_ERR_UNSUPPORTED;
  CiParams p;
  p.x = (const uint16_t*)in; p.w = (const float*)w; p.bias = bias; p.out = (uint16_t*)out;
  p.f16 = split ? 1 : ((f32 ? d->operand16 : d->in_dtype) == DSK_F16 ? 1 : 0);
  p.in_f32 = f32 ? 1 : 0; p.out_f32 = f32 ? 1 : 0;
  p.split = (f32 && split) ? 1 : 0;
  p.nstages = p.split ? 4 : CI_STAGES;
  static const int gather = [] { const char* e = getenv("DSK_CONVIN_GATHER"); return e ? atoi(e) : 0; }();
  p.staged = (d->Cin == 1 && !gather) ? 1 : 0;
  p.B = d->B; p.D = d->D; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout; p.ndim = d->ndim;
  p.circ = d->circular;
  p.tiles_w = (d->W + CI_BW - 1) / CI_BW; p.tiles_h = (d->H + CI_BH - 1) / CI_BH;
  p.total_tiles = p.tiles_w * p.tiles_h * d->D * d->B;
  if (d->ndim == 3) return d->Cin == 1 ? launch_convin<1, true>(p, st) : launch_convin<2, true>(p, st);   // 27 * Cin <= 64
  switch (d->Cin) {
    case 1: return launch_convin<1, false>(p, st);
    case 2: return launch_convin<2, false>(p, st);
    case 3: return launch_convin<3, false>(p, st);
    default: return launch_convin<4, false>(p, st);
  }
}

}  // namespace dsk
