// The LAST ResBlock1 pair of every MRF branch of a stage, the branch sum, the 1/nk average and the next leaky-relu
// in ONE launch (narrow stages, C = 32):
//     X = lrelu( ( sum_j  c2_j( lrelu( c1_j(P_j) + b1_j ) ) + b2_j + x(P_j) ) / nk )     models.py:278-284, modules.py:211-221
// P_j = a-form input of branch j's last pair, c1_j = Conv1d(C, C, k_j, dilation d_j), c2_j = Conv1d(C, C, k_j, 1).
//
// Why: before, each branch's last c1 was its own launch (HBM-bound: one tensor read, one written) and the fused MRF
// launch then re-read SIX tensors (three h, three residuals): 9 tensor reads + 4 writes per stage for three pairs.
// Here the three h tiles never leave the SM and the residuals come from the activation tiles that are resident for
// c1 anyway: 3 reads + 1 write.
//
// Tile: every branch's h tile covers the SAME 256 time rows [t0 - hmax, t0 - hmax + 256), hmax = max_j (k_j - 1)/2, so
// that the three c2 convs accumulate into one TMEM tile of 256 - 2*hmax valid output rows; branch j's c2 reads its h
// tile at row offset hmax - hk_j + tap.  Pipeline per CTA (tile i, branch j):
//   TMA A(i,j) -> c1(i,j) [acc1, double buffered] -> epi1: h(i,j) -> smem [double buffered] -> c2(i,j) [+= acc2(i)]
//   after j = nbr-1: epi2(i): acc2 + sum b2 + sum_j x(P_j) (from the resident tiles) -> /nk -> lrelu -> global
// Two MMA-issuing warps (c1 stream, c2 stream) and two epilogue groups of 8 warps, as in conv_pair.cu; all weights
// (2 * sum k_j taps x 2 KB) stay resident.
#include <algorithm>

#include "common.cuh"
#include "conv_mrfp.h"
#include "ptx.cuh"

namespace vd {

constexpr int kMpEpiWarps = 16;
constexpr int kMpThreads = 64 + 32 * kMpEpiWarps + 32;  // producer, c1 issuer, 16 epilogue warps, c2 issuer
constexpr int kMpC2Warp = 2 + kMpEpiWarps;
constexpr int kMpHRows = 272;                           // 256 + 2*hmax rounded up (hmax <= 8)

struct MrfpMaps {
  CUtensorMap a[kMpMaxBr];
};

template <int CH, bool F16>
__global__ void __launch_bounds__(kMpThreads, 1)
conv_mrfp_kernel(const __grid_constant__ MrfpMaps tm, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ MrfpParams p) {
  constexpr int KC = CH, ROWB = KC * 2;
  constexpr int B_STAGE = CH * ROWB;      // one tap's weights [CH][KC]
  constexpr int ACC_COLS = 2 * CH;        // two 128-row accumulators per conv
  constexpr int TMEM_COLS = 4 * ACC_COLS; // acc1[2] + acc2[2]
  constexpr int CHUNKS = CH / 16, NW = kMpEpiWarps / 8;   // warps per quadrant per group
  static_assert(CH == 32, "mrfp kernel: C = 32 (one 16-column chunk per epilogue warp)");
  static_assert(TMEM_COLS <= 512 && CHUNKS == NW, "mrfp kernel: a warp owns one 16-column chunk of both accumulators");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NA = p.na_stages;
  uint8_t* smemA = smem;
  uint8_t* smemW = smemA + NA * p.a_stage_bytes;
  uint8_t* smemH = smemW + p.ntaps * B_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemH + 2 * kMpHRows * ROWB);
  uint64_t* a_full = bars;                  // [kMpMaxNA]
  uint64_t* a_empty = a_full + kMpMaxNA;    // [kMpMaxNA]  1 (c1 retired) + 8 (epi2 warps read the residual)
  uint64_t* acc1_full = a_empty + kMpMaxNA;
  uint64_t* acc1_empty = acc1_full + 2;
  uint64_t* acc2_full = acc1_empty + 2;
  uint64_t* acc2_empty = acc2_full + 2;
  uint64_t* h_full = acc2_empty + 2;
  uint64_t* h_empty = h_full + 2;
  uint64_t* w_full = h_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  float* sbias = reinterpret_cast<float*>(bars + 32);    // 256 B of barriers, then (nbr + 1) * CH floats

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nbr = p.nbr;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    for (int j = 0; j < nbr; ++j) tma_prefetch_desc(&tm.a[j]);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < kMpMaxNA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1 + kMpEpiWarps / 2); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc1_full[i], 1); mbar_init(&acc1_empty[i], kMpEpiWarps / 2);
      mbar_init(&acc2_full[i], 1); mbar_init(&acc2_empty[i], kMpEpiWarps / 2);
      mbar_init(&h_full[i], kMpEpiWarps / 2);
      mbar_init(&h_empty[i], 1);
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < (kMpMaxBr + 1) * CH; i += kMpThreads) {
    const int j = i / CH, c = i % CH;
    sbias[i] = j < kMpMaxBr ? (j < nbr ? p.bias1[j][c] : 0.f) : p.bias2sum[c];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int hmax = p.hmax;
  const int my_tiles = p.total_tiles > (int)blockIdx.x ? (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(w_full, p.ntaps * B_STAGE);
      for (int tap = 0; tap < p.ntaps; ++tap) tma_load_3d(&tmW, w_full, smemW + tap * B_STAGE, 0, 0, tap);
      pdl_wait();   // the weights (static) load while the previous launch drains; activations only from here on
      uint32_t sa = 0, pa = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const uint32_t tile = blockIdx.x + i * gridDim.x;
        uint32_t b, mt;
        p.div_m.divmod(tile, b, mt);
        const int t0 = mt * p.bmo;
        for (int j = 0; j < nbr; ++j) {
          mbar_wait(&a_empty[sa], pa ^ 1);
          mbar_expect_tx(&a_full[sa], p.nboxes[j] * 64 * ROWB);
          for (int bx = 0; bx < p.nboxes[j]; ++bx)
            tma_load_3d(&tm.a[j], &a_full[sa], smemA + sa * p.a_stage_bytes + bx * 64 * ROWB, 0,
                        t0 + p.a_lo[j] + bx * 64, (int)b);
          if (++sa == (uint32_t)NA) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == kMpC2Warp) {
    // ------------------------------------------------------------ MMA issuers (warp-uniform; elected lane issues)
    constexpr uint32_t idesc = umma_idesc_f16(CH, F16);
    constexpr uint32_t desc_hi = umma_desc_hi(ROWB);
    const uint32_t leader = elect_one();
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(smemA)), w_lo0 = umma_desc_lo(smem_u32(smemW));
    const uint32_t h_lo0 = umma_desc_lo(smem_u32(smemH));
    const uint32_t a_stage16 = (uint32_t)p.a_stage_bytes >> 4;
    constexpr uint32_t h_buf16 = (uint32_t)(kMpHRows * ROWB) >> 4;
    mbar_wait(w_full, 0);
    tc_fence_after();
    if (warp == 1) {
      // c1 stream: acc1[n & 1] = c1_j(A(i, j)), n = running (tile, branch) index
      uint32_t sa = 0, pa = 0, n = 0;
      for (int i = 0; i < my_tiles; ++i) {
        for (int j = 0; j < nbr; ++j, ++n) {
          const uint32_t as = n & 1;
          mbar_wait(&acc1_empty[as], ((n >> 1) & 1) ^ 1);
          mbar_wait(&a_full[sa], pa);
          tc_fence_after();
          const uint32_t d_base = tmem_base + as * ACC_COLS;
          const uint32_t tap_step16 = (uint32_t)(p.dil[j] * ROWB) >> 4;
          uint32_t at = a_lo0 + sa * a_stage16, wt = w_lo0 + p.w1_tap[j] * (B_STAGE >> 4);
          const int k = p.k[j];
          for (int tap = 0; tap < k; ++tap, at += tap_step16, wt += B_STAGE >> 4) {
#pragma unroll
            for (int acc = 0; acc < 2; ++acc)
#pragma unroll
              for (int kk = 0; kk < KC / 16; ++kk)
                umma_f16_lohi(d_base + acc * CH, at + ((acc * 128 * ROWB + kk * 32) >> 4), desc_hi, wt + ((kk * 32) >> 4),
                              desc_hi, idesc, (tap > 0 || kk > 0) ? 1u : 0u, leader);
          }
          if (leader) {
            umma_commit(&acc1_full[as]);
            umma_commit(&a_empty[sa]);
          }
          if (++sa == (uint32_t)NA) { sa = 0; pa ^= 1; }
        }
      }
    } else {
      // c2 stream: acc2[i & 1] (+)= c2_j(h(i, j)); the branch sum is formed in the accumulator
      uint32_t n = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const uint32_t as = i & 1;
        const uint32_t d_base = tmem_base + 2 * ACC_COLS + as * ACC_COLS;
        for (int j = 0; j < nbr; ++j, ++n) {
          const uint32_t hb = n & 1;
          mbar_wait(&h_full[hb], (n >> 1) & 1);
          if (j == 0) mbar_wait(&acc2_empty[as], ((i >> 1) & 1) ^ 1);
          tc_fence_after();
          uint32_t ht = h_lo0 + hb * h_buf16 + ((uint32_t)((hmax - p.hk[j]) * ROWB) >> 4);
          uint32_t wt = w_lo0 + p.w2_tap[j] * (B_STAGE >> 4);
          const int k = p.k[j];
          for (int tap = 0; tap < k; ++tap, ht += ROWB >> 4, wt += B_STAGE >> 4) {
#pragma unroll
            for (int acc = 0; acc < 2; ++acc)
#pragma unroll
              for (int kk = 0; kk < KC / 16; ++kk)
                umma_f16_lohi(d_base + acc * CH, ht + ((acc * 128 * ROWB + kk * 32) >> 4), desc_hi, wt + ((kk * 32) >> 4),
                              desc_hi, idesc, (j > 0 || tap > 0 || kk > 0) ? 1u : 0u, leader);
          }
          if (leader) {
            umma_commit(&h_empty[hb]);
            if (j == nbr - 1) umma_commit(&acc2_full[as]);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int q = warp & 3;
    const int grp = (warp - 2) >> 3;          // 0: epi1 (h producer), 1: epi2 (output)
    const int hsel = ((warp - 2) & 7) >> 2;   // which of the group's two warps on this TMEM lane quadrant
    const int c0 = hsel * 16;                 // this warp's 16-column chunk (of both 128-row accumulators)
    const int L = p.L, C = CH;
    const float slope = p.slope, res_gain = p.res_gain;
    pdl_wait();   // output stores may overwrite a buffer the previous launch still reads

    if (grp == 0) {
      // h(i, j) = lrelu(c1_j + b1_j), zero outside the utterance, written as c2's swizzled K-major A operand
      float4 breg[kMpMaxBr][4];   // biases of this warp's 16 columns, per branch, in registers
#pragma unroll
      for (int j = 0; j < kMpMaxBr; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) breg[j][e] = *reinterpret_cast<const float4*>(sbias + j * CH + c0 + 4 * e);
      uint32_t n = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const uint32_t tile = blockIdx.x + i * gridDim.x;
        uint32_t b, mt;
        p.div_m.divmod(tile, b, mt);
        const int t0 = mt * p.bmo;
#pragma unroll
        for (int j = 0; j < kMpMaxBr; ++j) {
          if (j >= nbr) break;
          const uint32_t as = n & 1, hb = n & 1, ph = (n >> 1) & 1;
          uint8_t* const hbuf = smemH + hb * kMpHRows * ROWB;
          mbar_wait(&acc1_full[as], ph);
          tc_fence_after();
          bool h_free = false;
#pragma unroll
          for (int acc = 0; acc < 2; ++acc) {
            const int r = acc * 128 + q * 32 + lane;
            const int th = t0 - hmax + r;
            uint32_t a[16];
            __syncwarp();
            tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + as * ACC_COLS + acc * CH + c0, a);
            tmem_ld_wait();
            const bool inside = th >= 0 && th < L;
            uint4 o[2];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              uint32_t* o2 = reinterpret_cast<uint32_t*>(&o[h2]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int jj = h2 * 8 + e * 2;
                const float4 bq = breg[j][jj >> 2];
                float v0 = __uint_as_float(a[jj]) + ((jj & 3) == 0 ? bq.x : bq.z);
                float v1 = __uint_as_float(a[jj + 1]) + ((jj & 3) == 0 ? bq.y : bq.w);
                v0 = inside ? fmaxf(v0, v0 * slope) : 0.f;
                v1 = inside ? fmaxf(v1, v1 * slope) : 0.f;
                o2[e] = pack_act2<F16>(v0, v1);
              }
            }
            if (!h_free) {  // the c2 that last read this h buffer must have retired before it is overwritten
              mbar_wait(&h_empty[hb], ph ^ 1);
              h_free = true;
            }
            const uint32_t sw = (r >> 1) & 3;   // 64-byte rows: 16-byte chunk index XOR address bits [7, 9)
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2)
              *reinterpret_cast<uint4*>(hbuf + r * ROWB + ((((c0 >> 3) + h2) ^ sw) << 4)) = o[h2];
          }
          fence_proxy_async();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&h_full[hb]);
            mbar_arrive(&acc1_empty[as]);
          }
          ++n;
        }
      }
    } else {
      // X = lrelu((acc2 + sum b2 + sum_j x(P_j)) / nk): residuals from the resident activation tiles
      float4 b2[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) b2[e] = *reinterpret_cast<const float4*>(sbias + kMpMaxBr * CH + c0 + 4 * e);
      const float scale = p.scale, out_slope = p.out_slope;
      __nv_bfloat16* const out = p.out;
      uint32_t sa = 0;   // stage of (i, 0)
      for (int i = 0; i < my_tiles; ++i) {
        const uint32_t tile = blockIdx.x + i * gridDim.x;
        uint32_t b, mt;
        p.div_m.divmod(tile, b, mt);
        const int t0 = mt * p.bmo;
        const uint32_t as = i & 1;
        mbar_wait(&acc2_full[as], (i >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int acc = 0; acc < 2; ++acc) {
          const int i0 = acc * 128 + q * 32;            // first output row of this warp's 32
          uint32_t a[16];
          __syncwarp();
          tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + 2 * ACC_COLS + as * ACC_COLS + acc * CH + c0, a);
          // residual rows of the three branches from their resident tiles (issued while the TMEM load is in flight)
          uint4 rx[kMpMaxBr][2];
          uint32_t st = sa;
#pragma unroll
          for (int j = 0; j < kMpMaxBr; ++j) {
            if (j >= nbr) break;
            const uint8_t* atile = smemA + st * p.a_stage_bytes;
            const int ra = i0 + lane + hmax + p.hk[j] * p.dil[j];   // row of x(t0 + i0 + lane) in branch j's tile
            const uint32_t sw = (ra >> 1) & 3;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2)
              rx[j][h2] = *reinterpret_cast<const uint4*>(atile + ra * ROWB + ((((c0 >> 3) + h2) ^ sw) << 4));
            if (++st == (uint32_t)NA) st = 0;
          }
          tmem_ld_wait();
          // same summation order as the unfused schedule's MRF epilogue (epilogue.cuh EPI 3): acc + bias, + residuals in
          // branch order, * 1/nk -- the two schedules stay bit-identical
          float v[16];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[4 * e] = __uint_as_float(a[4 * e]) + b2[e].x;
            v[4 * e + 1] = __uint_as_float(a[4 * e + 1]) + b2[e].y;
            v[4 * e + 2] = __uint_as_float(a[4 * e + 2]) + b2[e].z;
            v[4 * e + 3] = __uint_as_float(a[4 * e + 3]) + b2[e].w;
          }
#pragma unroll
          for (int j = 0; j < kMpMaxBr; ++j) {
            if (j >= nbr) break;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const uint32_t* r2 = reinterpret_cast<const uint32_t*>(&rx[j][h2]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 xr = unpack_act2<F16>(r2[e]);
                v[h2 * 8 + e * 2] += xr.x >= 0.f ? xr.x : xr.x * res_gain;        // a-form -> residual stream
                v[h2 * 8 + e * 2 + 1] += xr.y >= 0.f ? xr.y : xr.y * res_gain;
              }
            }
          }
          uint4 ov[2];
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            uint32_t* o2 = reinterpret_cast<uint32_t*>(&ov[h2]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int jj = h2 * 8 + e * 2;
              const float v0 = v[jj] * scale, v1 = v[jj + 1] * scale;
              o2[e] = pack_act2<F16>(fmaxf(v0, v0 * out_slope), fmaxf(v1, v1 * out_slope));
            }
          }
          const int rows_valid = min(32, max(0, min(p.bmo - i0, L - (t0 + i0))));
          if (lane < rows_valid) st_global_v8(out + ((long)b * L + t0 + i0 + lane) * C + c0, ov[0], ov[1]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&acc2_empty[as]);
          uint32_t st = sa;
          for (int j = 0; j < nbr; ++j) {
            mbar_arrive(&a_empty[st]);
            if (++st == (uint32_t)NA) st = 0;
          }
        }
        sa += nbr;
        while (sa >= (uint32_t)NA) sa -= NA;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------- host side
int encode_tmap_3d(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                   bool swizzle);

static constexpr int kMpSmemBudget = 227 * 1024 - 1024 /*align*/ - 256 /*barriers*/ - 1024 /*bias*/;

struct MpGeom {
  int hmax, ntaps, a_stage, na;
  int nboxes[kMpMaxBr];
  bool ok;
};

static MpGeom mp_geom(int channels, int nbr, const int* k, const int* dil) {
  MpGeom g{};
  g.ok = false;
  if (channels != 32 || nbr < 1 || nbr > kMpMaxBr) return g;
  const int rowb = channels * 2;
  int sumk = 0;
  for (int j = 0; j < nbr; ++j) {
    if (k[j] % 2 == 0 || k[j] < 1 || k[j] > 17 || dil[j] < 1) return g;
    const int hk = (k[j] - 1) / 2;
    g.hmax = std::max(g.hmax, hk);
    g.nboxes[j] = (256 + 2 * hk * dil[j] + 63) / 64;
    g.a_stage = std::max(g.a_stage, g.nboxes[j] * 64 * rowb);
    sumk += k[j];
  }
  if (256 + 2 * g.hmax > kMpHRows) return g;
  g.ntaps = 2 * sumk;
  const int fixed = g.ntaps * channels * rowb + 2 * kMpHRows * rowb;
  g.na = std::min(kMpMaxNA, (kMpSmemBudget - fixed) / g.a_stage);
  if (g.na < nbr + 1) return g;   // the tiles of one output tile stay resident until its epilogue has read the residuals
  g.ok = true;
  return g;
}

bool mrfp_supported(int channels, int nbr, const int* k, const int* dil) { return mp_geom(channels, nbr, k, dil).ok; }

int plan_conv_mrfp(MrfpPlan* pl, int B, int L, int channels, int nbr, const int* k, const int* dil,
                   const __nv_bfloat16* const* xs, const __nv_bfloat16* w, int num_sms) {
  const MpGeom g = mp_geom(channels, nbr, k, dil);
  VD_CHECK(g.ok, "conv_mrfp: unsupported shape");
  MrfpParams& p = pl->p;
  p = MrfpParams{};
  p.B = B; p.L = L; p.nbr = nbr;
  p.hmax = g.hmax;
  p.bmo = 256 - 2 * g.hmax;
  int tap = 0;
  for (int j = 0; j < nbr; ++j) {
    p.k[j] = k[j]; p.dil[j] = dil[j]; p.hk[j] = (k[j] - 1) / 2;
    p.nboxes[j] = g.nboxes[j];
    p.a_lo[j] = -(g.hmax + p.hk[j] * dil[j]);
    p.w1_tap[j] = tap;
    tap += k[j];
  }
  for (int j = 0; j < nbr; ++j) { p.w2_tap[j] = tap; tap += k[j]; }
  p.ntaps = tap;
  p.a_stage_bytes = g.a_stage;
  p.na_stages = g.na;
  p.m_tiles = (L + p.bmo - 1) / p.bmo;
  p.total_tiles = B * p.m_tiles;
  p.div_m.init(p.m_tiles);
  p.scale = 1.f / nbr;
  pl->channels = channels;
  pl->grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  const int rowb = channels * 2;
  pl->smem = 1024 + (size_t)p.na_stages * p.a_stage_bytes + (size_t)p.ntaps * channels * rowb + 2 * kMpHRows * rowb + 256 + 1024;
  for (int j = 0; j < kMpMaxBr; ++j) {
    if (j < nbr) {
      if (encode_tmap_3d(&pl->tmA[j], xs[j], channels, L, B, channels, 64, true)) return 1;
    } else {
      pl->tmA[j] = pl->tmA[0];
    }
  }
  if (encode_tmap_3d(&pl->tmW, w, channels, channels, p.ntaps, channels, channels, true)) return 1;
  return 0;
}

template <int CH, bool F16>
static int launch_mrfp_typed(const MrfpPlan& pl, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VD_CUDA(cudaFuncSetAttribute(conv_mrfp_kernel<CH, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const size_t smem = pl.smem > (size_t)120 * 1024 ? pl.smem : (size_t)120 * 1024;   // one CTA per SM, see conv_tc.cu
  MrfpMaps maps;
  for (int j = 0; j < kMpMaxBr; ++j) maps.a[j] = pl.tmA[j];
  conv_mrfp_kernel<CH, F16><<<pl.grid, kMpThreads, smem, stream>>>(maps, pl.tmW, pl.p);
  VD_CUDA(cudaGetLastError());
  return 0;
}

int launch_conv_mrfp(MrfpPlan& pl, const float* const* bias1, const float* bias2sum, float slope, float out_slope,
                     __nv_bfloat16* out, cudaStream_t stream, int f16) {
  pl.p.f16 = f16;
  for (int j = 0; j < pl.p.nbr; ++j) pl.p.bias1[j] = bias1[j];
  pl.p.bias2sum = bias2sum;
  pl.p.slope = slope;
  pl.p.res_gain = 1.f / slope;
  pl.p.out_slope = out_slope;
  pl.p.out = out;
  if (pl.channels == 32) return f16 ? launch_mrfp_typed<32, true>(pl, stream) : launch_mrfp_typed<32, false>(pl, stream);
  set_error("conv_mrfp: no kernel instance");
  return 1;
}

}  // namespace vd
