// The tail of a 128-channel MRF stage in ONE launch (models.py:278-284, modules.py:211-221 last iteration):
//     X = lrelu( ( sum_j  c2_j( lrelu( c1_j(P_j) + b1_j ) ) + b2_j + x(P_j) ) / nbr , out_slope )
// i.e. the LAST ResBlock1 pair of every branch (c1_j: k_j taps, dilation d_j; c2_j: k_j taps, dilation 1), the branch sum,
// the 1/nk average and the next leaky-relu.  Before: one c1 launch per branch (P_j -> H_j through HBM) and a fused-MRF
// launch that re-reads three H_j and three P_j: 3.05 GB of DRAM traffic per 16 x 10 s step for this stage; here the three h
// tiles never leave the SM and the three residual tiles are the inputs just read: 0.86 GB.  The launch is no faster in
// isolation (the same MMAs), but under the power cap every run of the step sits at, DRAM traffic is time (DESIGN.md 4.5:
// 2.6 GB less = -1.5 % per step).  conv_mrfp.cu is the same fusion for the C = 32 stage (resident weights, time-as-M);
// here the weights stream and the tiles are conv_pairf.cu's: M = 128 output channels, N = up to 256 time rows.
//
// Per CTA tile (WO output rows) and branch j (sub-step n = 3 i + j), all on the tensor pipe in this order:
//   TMA x_j -> c1_j: D1[:, 0..N1_j) = sum_t W1_j[t] . X_j[row + t d_j]             (D1: TMEM columns [0, 256))
//   h epilogue: h_j = lrelu(D1 + b1_j), zero outside the utterance -> shared memory (c2's swizzled K-major B operand)
//   c2_j: D2[:, 0..WO) (+)= sum_t W2_j[t] . H_j[row + t]                            (D2: TMEM columns [256, 512))
//   after the last branch: output epilogue (D2 + sum b2 + sum_j x(P_j)) / nbr -> lrelu -> bf16, residual rows from L2.
// One x buffer, one h buffer, one D1, one D2: c1_{j+1} can only be issued once the h epilogue has drained D1, so every h
// epilogue is exposed (~2.5 k of ~21 k cycles per sub-step); the output epilogue overlaps the next tile's first c1.
#include <algorithm>

#include "common.cuh"
#include "conv_mrf128.h"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace vd {

constexpr int kM8EpiWarps = 16;
constexpr int kM8Threads = 96 + 32 * kM8EpiWarps;   // warp 0: weight producer, 1: MMA issuer, 2: x producer, 3..18: epilogue

template <bool F16>
__global__ void __launch_bounds__(kM8Threads, 1)
conv_mrf128_kernel(const __grid_constant__ Mrf128Maps tm, const __grid_constant__ CUtensorMap tmW,
                   const __grid_constant__ Mrf128Params p) {
  constexpr int KC = 64, ROWB = 128, NCH = 2;
  constexpr int B_STAGE = 128 * ROWB;      // one K-chunk of one tap: [128 out channels][64]
  constexpr int HS_SUB = 256 * ROWB;       // one K-chunk of the h tile
  constexpr int D2_COL = 256;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* XS = smem;                                  // [NCH] chunks of xr_max rows
  uint8_t* HS = XS + p.xbuf_bytes;                     // [NCH] chunks of 256 rows
  uint8_t* WS = HS + NCH * HS_SUB;                     // weight ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(WS + p.nw * B_STAGE);
  uint64_t* x_full = bars;
  uint64_t* x_empty = bars + 1;     // c1 of the sub-step has retired: the x buffer may take the next branch's tile
  uint64_t* d1_full = bars + 2;
  uint64_t* h_ready = bars + 3;     // h written (and D1 drained)
  uint64_t* h_empty = bars + 4;     // c2 of the sub-step has retired: the h buffer may be overwritten
  uint64_t* d2_full = bars + 5;
  uint64_t* d2_empty = bars + 6;
  uint64_t* w_full = bars + 8;
  uint64_t* w_empty = w_full + kM8MaxW;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + kM8MaxW);
  float* sbias = reinterpret_cast<float*>(bars + 32);                 // 256 B of barriers, then (nbr + 1) x 128 floats
  uint8_t* scratch_base = reinterpret_cast<uint8_t*>(sbias) + 2048;   // 16 warps x 1 KB

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NW = p.nw, nbr = p.nbr;

  if (warp == 0 && lane == 0) {
    for (int j = 0; j < nbr; ++j) tma_prefetch_desc(&tm.x[j]);
    tma_prefetch_desc(&tmW);
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    mbar_init(d1_full, 1);
    mbar_init(h_ready, kM8EpiWarps);
    mbar_init(h_empty, 1);
    mbar_init(d2_full, 1);
    mbar_init(d2_empty, kM8EpiWarps);
    for (int i = 0; i < kM8MaxW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < (kM8MaxBr + 1) * 128; i += kM8Threads) {
    const int j = i >> 7, c = i & 127;
    sbias[i] = j < kM8MaxBr ? (j < nbr ? p.bias1[j][c] : 0.f) : p.bias2sum[c];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int my_tiles =
      p.total_tiles > (int)blockIdx.x ? (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    // ------------------------------------------------------------ weight producer: per sub-step c1's chunks, then c2's
    if (lane == 0) {
      uint32_t sw = 0, pw = 0;
      for (int i = 0; i < my_tiles; ++i)
        for (int j = 0; j < nbr; ++j)
          for (int conv = 0; conv < 2; ++conv) {
            const int tbase = conv == 0 ? p.wbase1[j] : p.wbase2[j];
            for (int ch = 0; ch < NCH; ++ch)
              for (int tap = 0; tap < p.nt[j]; ++tap) {
                mbar_wait(&w_empty[sw], pw ^ 1);
                mbar_expect_tx(&w_full[sw], B_STAGE);
                tma_load_3d(&tmW, &w_full[sw], WS + sw * B_STAGE, ch * KC, 0, tbase + tap);
                if (++sw == (uint32_t)NW) { sw = 0; pw ^= 1; }
              }
          }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ x producer: branch j's tile (two K-chunks)
    if (lane == 0) {
      uint32_t n = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const uint32_t tile = blockIdx.x + i * gridDim.x;
        uint32_t b, mt;
        p.div_m.divmod(tile, b, mt);
        const int m0 = (int)mt * p.WO;
        for (int j = 0; j < nbr; ++j, ++n) {
          mbar_wait(x_empty, (n & 1) ^ 1);
          mbar_expect_tx(x_full, NCH * p.XR[j] * ROWB);
          const int row0 = m0 + p.xrow0[j];   // = m0 - hk_j - hk_j d_j
          const int nbx = p.xboxes[j], bxr = p.xbox_rows[j];
          for (int ch = 0; ch < NCH; ++ch)
            for (int bx = 0; bx < nbx; ++bx)
              tma_load_5d(&tm.x[j], x_full, XS + ch * p.xsub_bytes + bx * bxr * ROWB, 0, ch, 0, row0 + bx * bxr, (int)b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform; elected lane issues)
    constexpr uint32_t desc_hi = umma_desc_hi(ROWB);
    const uint32_t idesc2 = umma_idesc_f16(p.WO, F16);
    const uint32_t leader = elect_one();
    const uint32_t x_lo0 = umma_desc_lo(smem_u32(XS)), h_lo0 = umma_desc_lo(smem_u32(HS));
    const uint32_t w_lo0 = umma_desc_lo(smem_u32(WS));
    const uint32_t xsub16 = (uint32_t)p.xsub_bytes >> 4;
    uint32_t sw = 0, pw = 0, n = 0;
    for (int i = 0; i < my_tiles; ++i) {
      for (int j = 0; j < nbr; ++j, ++n) {
        const uint32_t par = n & 1;
        const int nt = p.nt[j];
        // ---- c1_j: D1 is free (the wait for h_ready of the previous sub-step below implies it)
        mbar_wait(x_full, par);
        tc_fence_after();
        const uint32_t idesc1 = umma_idesc_f16(p.N1[j], F16);
        uint32_t started = 0;
        for (int ch = 0; ch < NCH; ++ch)
          for (int tap = 0; tap < nt; ++tap) {
            mbar_wait(&w_full[sw], pw);
            tc_fence_after();
            const uint32_t w_lo = w_lo0 + sw * (B_STAGE >> 4);
            const uint32_t x_lo = x_lo0 + ch * xsub16 + ((uint32_t)(tap * p.dstep[j] * ROWB) >> 4);
#pragma unroll
            for (int kk = 0; kk < KC / 16; ++kk)
              umma_f16_lohi(tmem_base, w_lo + kk * 2, desc_hi, x_lo + kk * 2, desc_hi, idesc1, kk == 0 ? started : 1u,
                            leader);
            started = 1;
            if (leader) umma_commit(&w_empty[sw]);
            if (++sw == (uint32_t)NW) { sw = 0; pw ^= 1; }
          }
        if (leader) {
          umma_commit(d1_full);
          umma_commit(x_empty);
        }
        // ---- c2_j: D2 (+)= W2_j . h_j
        mbar_wait(h_ready, par);
        if (j == 0) mbar_wait(d2_empty, (i & 1) ^ 1);   // the previous tile's output epilogue has drained D2
        tc_fence_after();
        started = j > 0 ? 1u : 0u;
        for (int ch = 0; ch < NCH; ++ch)
          for (int tap = 0; tap < nt; ++tap) {
            mbar_wait(&w_full[sw], pw);
            tc_fence_after();
            const uint32_t w_lo = w_lo0 + sw * (B_STAGE >> 4);
            const uint32_t h_lo = h_lo0 + ((uint32_t)(ch * HS_SUB + tap * ROWB) >> 4);
#pragma unroll
            for (int kk = 0; kk < KC / 16; ++kk)
              umma_f16_lohi(tmem_base + D2_COL, w_lo + kk * 2, desc_hi, h_lo + kk * 2, desc_hi, idesc2,
                            kk == 0 ? started : 1u, leader);
            started = 1;
            if (leader) umma_commit(&w_empty[sw]);
            if (++sw == (uint32_t)NW) { sw = 0; pw ^= 1; }
          }
        if (leader) {
          umma_commit(h_empty);
          if (j == nbr - 1) umma_commit(d2_full);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int q = warp & 3;                 // TMEM lane quadrant (hardware rule: warp % 4)
    const int sub = (warp - 3) >> 2;        // which of the quadrant's four warps
    uint8_t* scratch = scratch_base + (warp - 3) * 1024;
    uint8_t* const dummy = scratch + lane * 16;  // stmatrix target for rows that fall outside the h tile
    // fragment layout (tmem_ld_frag): this thread's channels are q*32 + 8m + lane/4, m = 0..3
    float b2f[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) b2f[m] = sbias[kM8MaxBr * 128 + q * 32 + 8 * m + (lane >> 2)];
    const float slope = p.slope, res_gain = p.res_gain, out_slope = p.out_slope, scale = p.scale;
    const int L = p.L, WO = p.WO;
    const int n_oitems = WO / 16;
    ConvEpilogue ep{};
    for (int j = 0; j < kM8MaxBr; ++j) ep.res[j] = p.res[j < nbr ? j : 0];
    ep.nres = nbr;
    ep.out = p.out;
    ep.epi_smem = 1;
    const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16);
    // this warp's 32 channels inside a 128-byte K-chunk row of h: chunk q / 2, 16-byte pieces (q % 2) * 4 + lane / 8
    const int hchunk = q >> 1;
    const uint32_t c16 = (uint32_t)((q & 1) * 4 + (lane >> 3));

    uint32_t n = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const uint32_t tile = blockIdx.x + i * gridDim.x;
      uint32_t b, mt;
      p.div_m.divmod(tile, b, mt);
      const int m0 = (int)mt * WO;
      // ---- h_j = lrelu(c1_j + b1_j), zero outside the utterance, as c2's B operand
      for (int j = 0; j < nbr; ++j, ++n) {
        float b1f[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) b1f[m] = sbias[j * 128 + q * 32 + 8 * m + (lane >> 2)];
        const int hbase = m0 - p.hk[j];   // sample of h row 0
        const int HR = p.HR[j];
        const int n_hitems = p.N1[j] / 16;
        mbar_wait(d1_full, n & 1);
        mbar_wait(h_empty, (n & 1) ^ 1);   // c2 of the previous sub-step no longer reads the h buffer
        tc_fence_after();
        for (int it = sub; it < n_hitems; it += 4) {
          const int rel0 = it * 16;
          uint32_t a[16];
          __syncwarp();
          tmem_ld_frag(t_base + it * 16, a);
          uint32_t haddr[2];
#pragma unroll
          for (int cg = 0; cg < 2; ++cg) {
            const int rel = rel0 + cg * 8 + (lane & 7);
            haddr[cg] = rel < HR ? smem_u32(HS + hchunk * HS_SUB + rel * ROWB + ((c16 ^ (uint32_t)(rel & 7)) << 4))
                                 : smem_u32(dummy);
          }
          bool inside[2][2];
#pragma unroll
          for (int cg = 0; cg < 2; ++cg)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int arow = hbase + rel0 + cg * 8 + 2 * (lane & 3) + e;
              inside[cg][e] = arow >= 0 && arow < L;
            }
          tmem_ld_wait();
#pragma unroll
          for (int cg = 0; cg < 2; ++cg) {
            uint32_t pk[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              float v0 = __uint_as_float(a[frag_idx(m, cg, 0)]) + b1f[m];
              float v1 = __uint_as_float(a[frag_idx(m, cg, 1)]) + b1f[m];
              v0 = inside[cg][0] ? fmaxf(v0, v0 * slope) : 0.f;
              v1 = inside[cg][1] ? fmaxf(v1, v1 * slope) : 0.f;
              pk[m] = pack_act2<F16>(v0, v1);
            }
            stmatrix_x4_trans(haddr[cg], pk[0], pk[1], pk[2], pk[3]);
          }
        }
        fence_proxy_async();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(h_ready);
      }

      // ---- X = lrelu((D2 + sum b2 + sum_j x(P_j)) / nbr): conv_tc.cu's channels-as-M epilogue with three residuals
      auto ocoords = [&](int it, EpiItem& e) {
        const int t = m0 + it * 16;
        e.b = (int)b;
        e.n = q * 32;
        e.rows_valid = min(16, max(0, L - t));
        e.row0 = (long)b * L + t;
        e.base = e.row0 * 128 + q * 32;
        e.tcol = D2_COL + it * 16;
        e.t = t;
      };
      EpiLoads ld;
      EpiItem cur{};
      if (sub < n_oitems) {
        ocoords(sub, cur);
        epiT_issue_loads<0>(ep, cur, 128, lane, ld);
      }
      mbar_wait(d2_full, i & 1);
      tc_fence_after();
      for (int it = sub; it < n_oitems; it += 4) {
        uint32_t acc[kIW];
        float v[kIW];
        __syncwarp();
        tmem_ld_frag(t_base + cur.tcol, acc);
        tmem_ld_wait();
        epiT_accumulate<0, F16>(ep, b2f, scratch, cur, 128, lane, res_gain, acc, ld, v);
        const bool last = it + 4 >= n_oitems;
        if (last) {   // accumulator fully read by this warp: hand D2 back before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d2_empty);
        }
        const EpiItem done = cur;
        if (!last) {
          ocoords(it + 4, cur);
          epiT_issue_loads<0>(ep, cur, 128, lane, ld);
        }
#pragma unroll
        for (int e2 = 0; e2 < kIW; ++e2) v[e2] *= scale;
        epiT_store<1, F16>(ep, scratch, done, 128, lane, out_slope, 1.f, v);
      }
      if (sub >= n_oitems) {   // a warp without output items still owes its arrival
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d2_empty);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------- host side
int encode_tmap_3d(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                   bool swizzle);
int encode_tmap_act(CUtensorMap* m, const void* base, uint32_t kc, uint64_t d1, uint64_t s1, uint64_t d2, uint64_t s2,
                    uint64_t rows, uint64_t srow, uint64_t B, uint64_t sb, uint32_t box_rows, uint32_t box_d2);

static constexpr int kM8SmemBudget = 227 * 1024 - 1024 /*align*/ - 256 /*barriers*/ - 2048 /*bias*/ - 16384 /*scratch*/;

struct M8Geom {
  int WO, nw, xr_max;
  int HR[kM8MaxBr], N1[kM8MaxBr], XR[kM8MaxBr], nbx[kM8MaxBr], bxr[kM8MaxBr];
  bool ok;
};

static M8Geom m8_geom(int channels, int nbr, const int* k, const int* dil) {
  M8Geom g{};
  g.ok = false;
  if (channels != 128 || nbr < 1 || nbr > kM8MaxBr) return g;
  for (int j = 0; j < nbr; ++j)
    if (k[j] % 2 == 0 || k[j] < 1 || k[j] > 15 || dil[j] < 1 || dil[j] > 8) return g;
  for (int wo = 240; wo >= 64; wo -= 16) {
    bool fits = true;
    int xr_max = 0;
    for (int j = 0; j < nbr && fits; ++j) {
      const int hr = wo + k[j] - 1;
      const int n1 = (hr + 15) / 16 * 16;
      const int need = n1 + (k[j] - 1) * dil[j];       // rows c1's taps reach
      const int nbx = (need + 255) / 256;              // a TMA box has at most 256 rows; boxes start on 8-row swizzle atoms
      const int bxr = ((need + nbx - 1) / nbx + 7) / 8 * 8;
      const int xr = nbx * bxr;
      if (hr > 256 || n1 > 256) fits = false;
      g.HR[j] = hr; g.N1[j] = n1; g.XR[j] = xr; g.nbx[j] = nbx; g.bxr[j] = bxr;
      xr_max = std::max(xr_max, xr);
    }
    if (!fits) continue;
    const int xbuf = 2 * xr_max * 128;
    const int nw = std::min(kM8MaxW, (kM8SmemBudget - xbuf - 2 * 256 * 128) / (128 * 128));
    if (nw < 4) continue;
    g.WO = wo; g.nw = nw; g.xr_max = xr_max;
    g.ok = true;
    return g;
  }
  return g;
}

bool mrf128_supported(int channels, int nbr, const int* k, const int* dil) { return m8_geom(channels, nbr, k, dil).ok; }

int plan_conv_mrf128(Mrf128Plan* pl, int B, int L, int nbr, const int* k, const int* dil, const __nv_bfloat16* const* xs,
                     const __nv_bfloat16* w, int num_sms) {
  const M8Geom g = m8_geom(128, nbr, k, dil);
  VD_CHECK(g.ok, "conv_mrf128: unsupported shape");
  Mrf128Params& p = pl->p;
  p = Mrf128Params{};
  p.B = B; p.L = L; p.nbr = nbr; p.WO = g.WO; p.nw = g.nw;
  p.xsub_bytes = g.xr_max * 128;
  p.xbuf_bytes = 2 * p.xsub_bytes;
  int tap = 0;
  for (int j = 0; j < nbr; ++j) { p.wbase1[j] = tap; tap += k[j]; }   // packed order: c1 of every branch, then c2 of every branch
  for (int j = 0; j < nbr; ++j) { p.wbase2[j] = tap; tap += k[j]; }
  for (int j = 0; j < kM8MaxBr; ++j) {
    const int jj = j < nbr ? j : 0;
    const int hk = (k[jj] - 1) / 2;
    p.nt[j] = k[jj]; p.dstep[j] = dil[jj]; p.hk[j] = hk;
    p.xrow0[j] = -hk - hk * dil[jj];
    p.HR[j] = g.HR[jj]; p.N1[j] = g.N1[jj]; p.XR[j] = g.XR[jj];
    p.xboxes[j] = g.nbx[jj]; p.xbox_rows[j] = g.bxr[jj];
    p.res[j] = xs[jj];
  }
  p.m_tiles = (L + p.WO - 1) / p.WO;
  p.total_tiles = B * p.m_tiles;
  p.div_m.init(p.m_tiles);
  p.scale = 1.f / nbr;
  pl->grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  pl->smem = 1024 + (size_t)p.xbuf_bytes + (size_t)2 * 256 * 128 + (size_t)g.nw * 128 * 128 + 256 + 2048 + 16384;
  for (int j = 0; j < kM8MaxBr; ++j) {
    if (j < nbr) {   // plain [B][L][128] view: two 64-channel K-chunks per row, boxes of XR_j rows
      if (encode_tmap_act(&pl->tm.x[j], xs[j], 64, 2, 64, 1, 128, L, 128, B, (uint64_t)L * 128, g.bxr[j], 1)) return 1;
    } else {
      pl->tm.x[j] = pl->tm.x[0];
    }
  }
  if (encode_tmap_3d(&pl->tmW, w, 128, 128, tap, 64, 128, true)) return 1;
  return 0;
}

template <bool F16>
static int launch_m8_typed(const Mrf128Plan& pl, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VD_CUDA(cudaFuncSetAttribute(conv_mrf128_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  conv_mrf128_kernel<F16><<<pl.grid, kM8Threads, pl.smem, stream>>>(pl.tm, pl.tmW, pl.p);
  VD_CUDA(cudaGetLastError());
  return 0;
}

int launch_conv_mrf128(Mrf128Plan& pl, const float* const* bias1, const float* bias2sum, float slope, float out_slope,
                       __nv_bfloat16* out, cudaStream_t stream, int f16) {
  for (int j = 0; j < pl.p.nbr; ++j) pl.p.bias1[j] = bias1[j];
  for (int j = pl.p.nbr; j < kM8MaxBr; ++j) pl.p.bias1[j] = bias1[0];
  pl.p.bias2sum = bias2sum;
  pl.p.slope = slope;
  pl.p.res_gain = 1.f / slope;
  pl.p.out_slope = out_slope;
  pl.p.out = out;
  return f16 ? launch_m8_typed<true>(pl, stream) : launch_m8_typed<false>(pl, stream);
}

}  // namespace vd
