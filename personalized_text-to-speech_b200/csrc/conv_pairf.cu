// Fused ResBlock1 pair on 128-(virtual-)channel channels-as-M tiles: time-folded for the narrow stages (C = 32 / 64), and
// on the plain view for C = 128 (r = 1: the k = 3 pairs of stage 1):
//     y = lrelu( c2( lrelu( c1(a) + b1 ) ) + b2 + x(a) )          (modules.py:211-221, one loop iteration)
// Same fusion as conv_pair.cu (h never leaves the SM), but both convs run on the folded view of decoder.cu
// fold_geom: r = 128/C time samples per row, 128 virtual channels, block-Toeplitz weights, channels-as-M tiles
// (M = 128 virtual output channels, N = up to 256 folded rows per instruction).  The N = 32/64 time-as-M tiles of
// conv_pair.cu read 4 KB of activations from shared memory per 16/32-cycle MMA and run at 25-45 % of the tensor
// rate; here one MMA reads 4 KB of weights + 8 KB of activations per 128 cycles.
//
// Per CTA tile: WO folded output rows (c2's N), HR = WO + nt - 1 rows of h (c2's halo).
//   c1 (dilation d): for each sub-sequence rho of its dilated view, D1[:, rho*N1 ..] = sum_taps W1' . X(rho, phase)
//        X sub-tiles arrive by TMA (one per (rho, phase), 5-d view), weights stream through a ring
//   h  : TMEM -> registers (+b1, leaky-relu, zero outside the utterance) -> shared memory in the NATURAL folded
//        layout (row = sample / r, K-chunk = sample % r), i.e. c2's swizzled K-major B operand; a dilated c1
//        produces samples d*(r*n+phi)+rho, so this store is where the sub-sequences are interleaved back
//   c2 (dilation 1): D2 = sum_taps W2' . H
//   out: D2 + b2 + x (residual re-read from global memory: the tile was loaded a moment ago, L2-hot) -> lrelu -> bf16
//
// Two tiles in flight.  A tile owns one SLOT = 256 TMEM columns + one shared-memory buffer for its whole life:
//   buffer:  x(i) --c1--> dead --h epilogue writes h(i) IN PLACE--> c2 --> dead --> x(i+2)
//   columns: D1(i) --h epilogue--> dead --c2 writes D2(i) over it--> output epilogue --> D1(i+2)
// so the tensor pipe runs c1(i+1) / c2(i+1) of the other slot while this slot's accumulator is in an epilogue.  (The
// first version held one tile: D1 and D2 side by side, x and h side by side, and every h epilogue was exposed:
// cycles/tile = c1 + h epilogue + c2, profiles/r01_trace_pairf.txt.)
//   MMA order       c1(0) c1(1) | c2(j) c2(j+1) c1(j+2) c1(j+3) | ...       j = 0, 2, 4, ...
//   epilogue order  h(0)  h(1)  | o(j)  o(j+1)  h(j+2)  h(j+3)  | ...       (all 16 warps work on one accumulator at a time)
#include <algorithm>

#include "common.cuh"
#include "conv_pairf.h"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace vd {

#ifndef VITSDEC_TRACE
#define VITSDEC_TRACE 0
#endif
constexpr bool kPfTrace = VITSDEC_TRACE != 0;
constexpr int kPfEpiWarps = 16;
constexpr int kPfThreads = 96 + 32 * kPfEpiWarps;  // warp 0: weight producer, 1: MMA issuer, 2: x producer, 3..18: epilogue
constexpr int kPfMaxOItems = 4;                    // output items (16 columns) per epilogue warp and tile: WO <= 256

// K-chunks are always 64 virtual channels = 128-byte rows (SWIZZLE_128B): with 64-byte rows (one 32-channel phase
// per chunk) the 32-byte K=16 slices of 8 consecutive rows fall on the same banks twice and every MMA ran at half
// rate (320 vs 160 cycles at N=256, profiles/r01_trace_pairf.txt).  For C = 32 a chunk is two time phases.
template <int CH, bool F16>
__global__ void __launch_bounds__(kPfThreads, 1)
conv_pairf_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ PairFParams p) {
  constexpr int KC = 64, ROWB = 128;
  constexpr int R = 128 / CH;              // time samples (phases) per folded row (1: plain 128-channel view)
  constexpr int RSH = CH == 32 ? 2 : (CH == 64 ? 1 : 0);    // log2(R)
  constexpr int PPC = CH <= KC ? KC / CH : 1;               // phases per K-chunk (a dilated view needs CH <= 64)
  constexpr int NCH = 2;                   // K-chunks per folded row
  constexpr int B_STAGE = 128 * ROWB;      // one K-chunk of one folded tap: [128 virtual out channels][64]
  constexpr int HS_SUB = 256 * ROWB;       // one K-chunk of the h tile
  constexpr int SLOT_COLS = 256;           // TMEM columns of one tile slot

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int xs_sub = p.XR * ROWB;
  const int buf_bytes = p.buf_bytes;                // one slot: x as [d][NCH] sub-tiles of XR rows, then h as [NCH] chunks of 256 rows
  uint8_t* BUF = smem;
  uint8_t* WS = BUF + 2 * buf_bytes;                // weight ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(WS + p.nw * B_STAGE);
  uint64_t* x_full = bars;            // [2] per slot
  uint64_t* x_empty = bars + 2;       // [2] c2 of the slot's tile has retired: the buffer may take the next x tile
  uint64_t* d1_full = bars + 4;       // [2]
  uint64_t* h_ready = bars + 6;       // [2]
  uint64_t* d2_full = bars + 8;       // [2]
  uint64_t* d2_empty = bars + 10;     // [2] the output epilogue has read D2: the slot's columns may take the next D1
  uint64_t* w_full = bars + 12;
  uint64_t* w_empty = w_full + kPfMaxW;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + kPfMaxW);
  float* sbias = reinterpret_cast<float*>(bars + 32);                 // 256 B of barriers, then 2 x 128 floats
  uint8_t* scratch_base = reinterpret_cast<uint8_t*>(sbias) + 1024;   // 16 warps x 1 KB

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NW = p.nw;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], 1);
      mbar_init(&d1_full[s], 1);
      mbar_init(&h_ready[s], kPfEpiWarps);
      mbar_init(&d2_full[s], 1);
      mbar_init(&d2_empty[s], kPfEpiWarps);
    }
    for (int i = 0; i < kPfMaxW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 256; i += kPfThreads) sbias[i] = i < 128 ? p.bias1[i % CH] : p.bias2[i % CH];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int my_tiles =
      p.total_tiles > (int)blockIdx.x ? (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int dR = p.d * R;
  // first row of sub-sequence rho that holds a sample >= R*hbase:  floor((R*hbase - rho) / (d*R))
  auto nlo_of = [&](int hbase, int rho) -> int {
    return (int)p.div_dr.quot((uint32_t)(R * hbase - rho + 1024 * dR)) - 1024;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ weight producer: the chunks of every conv, in MMA order
    if (lane == 0) {
      uint32_t sw = 0, pw = 0;   // ring stage / phase
      auto stream = [&](int conv) {
        for (int ch = 0; ch < NCH; ++ch) {
          for (int tap = 0; tap < p.nt; ++tap) {
            if (!((p.kmask[tap] >> ch) & 1u)) continue;
            mbar_wait(&w_empty[sw], pw ^ 1);
            mbar_expect_tx(&w_full[sw], B_STAGE);
            tma_load_3d(&tmW, &w_full[sw], WS + sw * B_STAGE, ch * KC, 0, conv * p.nt + tap);
            if (++sw == (uint32_t)NW) { sw = 0; pw ^= 1; }
          }
        }
      };
      if (my_tiles > 0) stream(0);
      if (my_tiles > 1) stream(0);
      for (int j = 0; j < my_tiles; j += 2) {
        stream(1);
        if (j + 1 < my_tiles) stream(1);
        if (j + 2 < my_tiles) stream(0);
        if (j + 3 < my_tiles) stream(0);
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ x producer: d*R sub-tiles per tile
    if (lane == 0) {
      for (int i = 0; i < my_tiles; ++i) {
        const uint32_t tile = blockIdx.x + i * gridDim.x;
        uint32_t b, mt;
        p.div_m.divmod(tile, b, mt);
        const int hbase = (int)mt * p.WO + p.smin;
        const int slot = i & 1;
        uint8_t* XS = BUF + slot * buf_bytes;
        mbar_wait(&x_empty[slot], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&x_full[slot], p.d * NCH * xs_sub);
        for (int rho = 0; rho < p.d; ++rho) {
          const int row0 = nlo_of(hbase, rho) + p.xmin;
          for (int ch = 0; ch < NCH; ++ch) {
            uint8_t* dst = XS + (rho * NCH + ch) * xs_sub;
            // dilated view [C][rho][phase][row][b]: a chunk is PPC phases (box {C, 1, PPC, XR, 1} -> 128-byte rows)
            if (p.d > 1) tma_load_5d(&tmX, &x_full[slot], dst, 0, rho, ch * PPC, row0, (int)b);
            else tma_load_5d(&tmX, &x_full[slot], dst, 0, ch, 0, row0, (int)b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform; elected lane issues)
    const uint32_t idesc1 = umma_idesc_f16(p.N1, F16), idesc2 = umma_idesc_f16(p.WO, F16);
    constexpr uint32_t desc_hi = umma_desc_hi(ROWB);
    const uint32_t leader = elect_one();
    const uint32_t buf_lo0 = umma_desc_lo(smem_u32(BUF));
    const uint32_t w_lo0 = umma_desc_lo(smem_u32(WS));
    uint32_t sw = 0, pw = 0;   // ring stage / phase
    auto c1 = [&](int i) {
      const int slot = i & 1;
      const uint32_t par = (i >> 1) & 1;
      uint8_t* XS = BUF + slot * buf_bytes;
      const uint32_t x_lo0 = buf_lo0 + ((uint32_t)(slot * buf_bytes) >> 4);
      const uint32_t d_base = tmem_base + slot * SLOT_COLS;
      const bool tr = kPfTrace && p.trace && blockIdx.x == 0 && i < 256 && lane == 0;
      if (tr) p.trace[i * 12 + 8] = clock64();
      mbar_wait(&d2_empty[slot], par ^ 1);   // the slot's columns: output epilogue of tile i - 2 done
      mbar_wait(&x_full[slot], par);
      tc_fence_after();
      if (tr) p.trace[i * 12 + 0] = clock64();
      if (p.d > 1) {
        // rows of the dilated view past the utterance end alias the next utterance: zero them (see conv_tc.cu)
        const uint32_t tile = blockIdx.x + i * gridDim.x;
        uint32_t b, mt;
        p.div_m.divmod(tile, b, mt);
        const int hbase = (int)mt * p.WO + p.smin;
        bool wrote = false;
        for (int rho = 0; rho < p.d; ++rho) {
          const int idx = p.rows_rho - 1 - (nlo_of(hbase, rho) + p.xmin);
          if (idx < 0 || idx >= p.XR) continue;
          for (int psi = 0; psi < R; ++psi) {
            const int rem = p.L - rho - p.d * psi;
            const int nlim = rem > 0 ? (int)p.div_dr.quot(rem + dR - 1) : 0;
            if (nlim < p.rows_rho) {
              // phase psi = CH channels of the row: 16-byte chunks (psi % PPC) * CH/8 + lane, 128B-swizzled
              if (lane < CH / 8) {
                const uint32_t c16 = (uint32_t)((psi % PPC) * (CH / 8) + lane);
                *reinterpret_cast<uint4*>(XS + (rho * NCH + psi / PPC) * xs_sub + idx * ROWB + ((c16 ^ (idx & 7)) << 4)) =
                    make_uint4(0, 0, 0, 0);
              }
              wrote = true;
            }
          }
        }
        if (wrote) {
          fence_proxy_async();
          __syncwarp();
        }
      }
      uint32_t started = 0;
      for (int ch = 0; ch < NCH; ++ch) {
        for (int tap = 0; tap < p.nt; ++tap) {
          if (!((p.kmask[tap] >> ch) & 1u)) continue;
          mbar_wait(&w_full[sw], pw);
          tc_fence_after();
          const uint32_t w_lo = w_lo0 + sw * (B_STAGE >> 4);
          for (int rho = 0; rho < p.d; ++rho) {
            const uint32_t x_lo = x_lo0 + ((uint32_t)((rho * NCH + ch) * xs_sub + tap * p.dstep * ROWB) >> 4);
#pragma unroll
            for (int kk = 0; kk < KC / 16; ++kk)
              umma_f16_lohi(d_base + rho * p.N1, w_lo + kk * 2, desc_hi, x_lo + kk * 2, desc_hi, idesc1,
                            kk == 0 ? started : 1u, leader);
          }
          started = 1;
          if (leader) umma_commit(&w_empty[sw]);
          if (++sw == (uint32_t)NW) { sw = 0; pw ^= 1; }
        }
      }
      if (leader) umma_commit(&d1_full[slot]);
      if (tr) p.trace[i * 12 + 1] = clock64();
    };
    auto c2 = [&](int i) {
      const int slot = i & 1;
      const uint32_t par = (i >> 1) & 1;
      const uint32_t h_lo0 = buf_lo0 + ((uint32_t)(slot * buf_bytes) >> 4);
      const uint32_t d_base = tmem_base + slot * SLOT_COLS;
      const bool tr = kPfTrace && p.trace && blockIdx.x == 0 && i < 256 && lane == 0;
      if (tr) p.trace[i * 12 + 9] = clock64();
      mbar_wait(&h_ready[slot], par);   // (the h epilogue has also finished reading D1: D2 may overwrite it)
      tc_fence_after();
      if (tr) p.trace[i * 12 + 2] = clock64();
      uint32_t started = 0;
      for (int ch = 0; ch < NCH; ++ch) {
        for (int tap = 0; tap < p.nt; ++tap) {
          if (!((p.kmask[tap] >> ch) & 1u)) continue;
          mbar_wait(&w_full[sw], pw);
          tc_fence_after();
          const uint32_t w_lo = w_lo0 + sw * (B_STAGE >> 4);
          const uint32_t h_lo = h_lo0 + ((uint32_t)(ch * HS_SUB + tap * ROWB) >> 4);
#pragma unroll
          for (int kk = 0; kk < KC / 16; ++kk)
            umma_f16_lohi(d_base, w_lo + kk * 2, desc_hi, h_lo + kk * 2, desc_hi, idesc2, kk == 0 ? started : 1u, leader);
          started = 1;
          if (leader) umma_commit(&w_empty[sw]);
          if (++sw == (uint32_t)NW) { sw = 0; pw ^= 1; }
        }
      }
      if (leader) {
        umma_commit(&d2_full[slot]);
        umma_commit(&x_empty[slot]);
      }
      if (tr) p.trace[i * 12 + 3] = clock64();
    };
    if (my_tiles > 0) c1(0);
    if (my_tiles > 1) c1(1);
    for (int j = 0; j < my_tiles; j += 2) {
      c2(j);
      if (j + 1 < my_tiles) c2(j + 1);
      if (j + 2 < my_tiles) c1(j + 2);
      if (j + 3 < my_tiles) c1(j + 3);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int q = warp & 3;                 // TMEM lane quadrant (hardware rule: warp % 4)
    const int sub = (warp - 3) >> 2;        // which of the quadrant's four warps
    uint8_t* scratch = scratch_base + (warp - 3) * 1024;
    const int phi = (q * 32) / CH;          // time phase of this warp's virtual channels (phase * C + channel)
    // fragment layout (tmem_ld_frag): this thread's channels are q*32 + 8m + lane/4, m = 0..3
    float b1f[4], b2f[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      b1f[m] = sbias[q * 32 + 8 * m + (lane >> 2)];
      b2f[m] = sbias[128 + q * 32 + 8 * m + (lane >> 2)];
    }
    const int co0 = (q * 32) % CH;          // first real channel of this warp's 32 virtual channels
    uint8_t* const dummy = scratch + lane * 16;  // stmatrix target for rows that fall outside the h tile
    const float slope = p.slope, res_gain = p.res_gain;
    const int d = p.d, N1 = p.N1, HR = p.HR, Lf = p.Lf, WO = p.WO;
    const int ipr = N1 / 16;                // h items per sub-sequence
    const int n_hitems = d * ipr, n_oitems = WO / 16;
    ConvEpilogue ep{};
    ep.res[0] = p.x;
    ep.out = p.out;
    ep.epi_smem = 1;

    // ---- h(i) = lrelu(c1 + b1), zero outside the utterance, stored over x(i) as c2's B operand (natural folded layout)
    auto hepi = [&](int i) {
      const int slot = i & 1;
      const uint32_t par = (i >> 1) & 1;
      uint8_t* HS = BUF + slot * buf_bytes;
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + slot * SLOT_COLS;
      const uint32_t tile = blockIdx.x + i * gridDim.x;
      uint32_t b, mt;
      p.div_m.divmod(tile, b, mt);
      const int hbase = (int)mt * WO + p.smin;
      mbar_wait(&d1_full[slot], par);   // every c1 MMA of the tile has retired: x(i) is dead, D1(i) complete
      tc_fence_after();
      const bool tr = kPfTrace && p.trace && blockIdx.x == 0 && warp == 3 && lane == 0 && i < 256;
      if (tr) p.trace[i * 12 + 4] = clock64();
      for (int it = sub; it < n_hitems; it += 4) {
        int rho = 0, jn = it;
        while (jn >= ipr) { jn -= ipr; ++rho; }
        const int u = d * phi + rho;                       // sample offset inside a row of the dilated view
        const int phase = u & (R - 1), rowoff = u >> RSH;  // its place in the natural view
        const int rel0 = d * (nlo_of(hbase, rho) + jn * 16) + rowoff - hbase;  // h-tile row of column 0; +d per column
        uint32_t a[16];
        __syncwarp();
        tmem_ld_frag(t_base + it * 16, a);
        // stmatrix row addresses: lane i stores row (i & 7) of the 8-channel chunk (i >> 3), for columns cg*8 + (i & 7)
        uint32_t haddr[2];
#pragma unroll
        for (int cg = 0; cg < 2; ++cg) {
          const int rel = rel0 + d * (cg * 8 + (lane & 7));
          const int vcol0 = phase * CH + co0;   // first virtual channel of this warp's 32 in the natural folded row
          const uint32_t c16 = (uint32_t)((vcol0 & (KC - 1)) >> 3) + (lane >> 3);
          haddr[cg] = (rel >= 0 && rel < HR)
                          ? smem_u32(HS + (vcol0 >> 6) * HS_SUB + rel * ROWB + ((c16 ^ (rel & 7)) << 4))
                          : smem_u32(dummy);
        }
        // zero outside the utterance (c2's zero padding): per column held by this thread
        bool inside[2][2];
#pragma unroll
        for (int cg = 0; cg < 2; ++cg)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int arow = rel0 + d * (cg * 8 + 2 * (lane & 3) + e) + hbase;
            inside[cg][e] = arow >= 0 && arow < Lf;
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
      if (lane == 0) mbar_arrive(&h_ready[slot]);
      if (tr) p.trace[i * 12 + 5] = clock64();
    };

    // ---- y(i) = lrelu(c2 + b2 + x): same item pipeline as conv_tc.cu's channels-as-M epilogue; the residual rows of ALL
    // of this warp's items are requested before the accumulator is waited for (they come from L2)
    auto oepi = [&](int i) {
      const int slot = i & 1;
      const uint32_t par = (i >> 1) & 1;
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + slot * SLOT_COLS;
      const uint32_t tile = blockIdx.x + i * gridDim.x;
      uint32_t b, mt;
      p.div_m.divmod(tile, b, mt);
      const int n0 = (int)mt * WO;
      auto ocoords = [&](int it, EpiItem& e) {
        const int t = n0 + it * 16;
        e.b = (int)b;
        e.n = q * 32;
        e.rows_valid = min(16, max(0, Lf - t));
        e.row0 = (long)b * Lf + t;
        e.base = e.row0 * 128 + q * 32;
        e.tcol = it * 16;
      };
      // residual rows of the first two items are requested before the accumulator is waited for, the others two items
      // ahead of their use (they come from L2: the tile was loaded a moment ago)
      EpiLoads ld[2];
#pragma unroll
      for (int kq = 0; kq < 2; ++kq) {
        const int it = sub + 4 * kq;
        if (it < n_oitems) {
          EpiItem e{};
          ocoords(it, e);
          epiT_issue_loads<2>(ep, e, 128, lane, ld[kq]);
        }
      }
      mbar_wait(&d2_full[slot], par);
      tc_fence_after();
      const bool tr = kPfTrace && p.trace && blockIdx.x == 0 && warp == 3 && lane == 0 && i < 256;
      if (tr) p.trace[i * 12 + 6] = clock64();
#pragma unroll
      for (int kq = 0; kq < kPfMaxOItems; ++kq) {
        const int it = sub + 4 * kq;
        if (it < n_oitems) {
          EpiItem cur{};
          ocoords(it, cur);
          uint32_t acc[kIW];
          float v[kIW];
          __syncwarp();
          tmem_ld_frag(t_base + cur.tcol, acc);
          tmem_ld_wait();
          epiT_accumulate<2, F16>(ep, b2f, scratch, cur, 128, lane, res_gain, acc, ld[kq & 1], v);
          if (it + 4 >= n_oitems) {  // accumulator fully read by this warp: hand the slot's columns back before the stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d2_empty[slot]);
          }
          if (kq + 2 < kPfMaxOItems && it + 8 < n_oitems) {
            EpiItem e{};
            ocoords(it + 8, e);
            epiT_issue_loads<2>(ep, e, 128, lane, ld[kq & 1]);
          }
          epiT_store<2, F16>(ep, scratch, cur, 128, lane, slope, 1.f, v);
        }
      }
      if (tr) p.trace[i * 12 + 7] = clock64();
      if (sub >= n_oitems) {  // a warp without output items still owes its arrival
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d2_empty[slot]);
      }
    };

    if (my_tiles > 0) hepi(0);
    if (my_tiles > 1) hepi(1);
    for (int j = 0; j < my_tiles; j += 2) {
      oepi(j);
      if (j + 1 < my_tiles) oepi(j + 1);
      if (j + 2 < my_tiles) hepi(j + 2);
      if (j + 3 < my_tiles) hepi(j + 3);
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

static constexpr int kPfSmemBudget = 227 * 1024 - 1024 /*align*/ - 256 /*barriers*/ - 1024 /*bias*/ - 16384 /*scratch*/;

static int pf_floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
static int hk_of(int k) { return (k - 1) / 2; }

struct PfGeom {
  int r, nt, smin, WO, HR, N1, XR, nw, buf;
  bool ok;
};

static PfGeom pf_geom(int channels, int k, int dil) {
  PfGeom g{};
  g.ok = false;
  if ((channels != 32 && channels != 64 && channels != 128) || k % 2 == 0 || k > 15 || dil < 1 || dil > 8) return g;
  // C = 32 with a dilated c1 would need two time phases side by side in one 128-byte shared-memory row (TMA box
  // {32, 1, 2, rows}); that load did not produce the expected layout on hardware (tools/probes/tma_box_probe.cu) and the
  // 64-byte-row alternative runs the MMAs at half rate, so those pairs stay on conv_pair.cu.
  if (channels == 32 && dil > 1) return g;
  const int r = 128 / channels, hk = (k - 1) / 2, rowb = 128;
  g.r = r;
  g.smin = pf_floordiv(-hk, r);
  g.nt = pf_floordiv(r - 1 + hk, r) - g.smin + 1;
  if (g.nt > kPfMaxTaps) return g;
  for (int wo = 240; wo >= 64; wo -= 16) {
    const int hr = wo + g.nt - 1;
    if (hr > 256) continue;
    // rows of one sub-sequence that hold a sample of the h tile (hr*r consecutive samples); on the plain view (r = 1) a
    // dilated c1 is one tile whose taps are `dil` rows apart, not `dil` sub-sequences
    const bool plain = r == 1;
    const int nsub = plain ? 1 : dil;
    const int need = (dil == 1 || plain) ? hr : (hr * r - 1) / (dil * r) + 2;
    const int n1 = (need + 15) / 16 * 16;
    if (nsub * n1 > 256) continue;
    const int xr = (n1 + (g.nt - 1) * (plain ? dil : 1) + 7) / 8 * 8;
    if (xr > 256) continue;
    // one slot buffer holds the x sub-tiles, then (in place) the h tile; two slots
    const int buf = (std::max(nsub * 2 * xr * rowb, 2 * 256 * rowb) + 1023) / 1024 * 1024;
    const int nw = std::min(kPfMaxW, (kPfSmemBudget - 2 * buf) / (128 * rowb));
    if (nw < 3) continue;
    g.WO = wo; g.HR = hr; g.N1 = n1; g.XR = xr; g.nw = nw; g.buf = buf;
    g.ok = true;
    return g;
  }
  return g;
}

int pairf_taps(int channels, int k) { return pf_geom(channels, k, 1).nt; }
bool pairf_supported(int channels, int k, int dil) { return pf_geom(channels, k, dil).ok; }
// Where this kernel is the default inside the 16 x 10 s decode (ncu launch lists under profiles/, A/B runs of bench.py):
//   C = 128 (plain view): every non-final pair.  The k = 3 pairs were HBM-bound launches on their own (212 us fused, 231
//           as two launches); for k = 7 / 11 the two launches are no slower in isolation, but each fused pair saves
//           ~640 MB of DRAM traffic, and under the power cap every run of the step sits at that is what counts:
//           -1.5 % per step with all four of them fused (9.16 -> 9.02 ms A/B, option pairf 1 vs 3-style rules);
//   C = 64 (2-sample folded view): the dilation-1 pairs with k >= 7 -- the N = 64 time-as-M tiles of conv_pair.cu run the
//           tensor pipe at 48 cycles per 32-cycle MMA and the k = 11 pairs did not fit there at all (two launches, 5 tensor
//           passes); dilated pairs on a folded view and k = 3 stay where they were (sub-sequence MMAs of N = 80, see
//           above: the dilated k = 11 pair measured 368 us here against 301 us as two folded launches);
//   C = 32: conv_mrfp.cu (the 4-sample fold doubles the MACs of a k = 3 conv).
// Option pairf: 0 never, 1 this rule, 2 wherever the kernel exists (tests), 3 C = 128 only and there k <= 5 only (the
// first round-2 rule, A/B).
bool pairf_preferred(int channels, int k, int dil) {
  return channels == 128 || (channels == 64 && dil == 1 && k >= 7);
}

int plan_conv_pairf(PairFPlan* pl, int B, int L, int channels, int k, int dil, const __nv_bfloat16* x,
                    const __nv_bfloat16* w_fold, int num_sms) {
  const PfGeom g = pf_geom(channels, k, dil);
  VD_CHECK(g.ok, "conv_pairf: unsupported shape");
  VD_CHECK(L % g.r == 0, "conv_pairf: the utterance length must be a multiple of the fold factor");
  PairFParams& p = pl->p;
  const bool plain = g.r == 1;
  p.B = B; p.L = L; p.Lf = L / g.r; p.d = plain ? 1 : dil;
  p.dstep = plain ? dil : 1;
  p.xmin = plain ? -hk_of(k) * dil : g.smin;
  p.WO = g.WO; p.HR = g.HR; p.N1 = g.N1; p.XR = g.XR; p.nt = g.nt; p.smin = g.smin; p.nw = g.nw; p.buf_bytes = g.buf;
  const int hk = (k - 1) / 2;
  for (int t = 0; t < g.nt; ++t) {
    uint32_t mask = 0;
    for (int psi = 0; psi < g.r; ++psi)
      for (int phi = 0; phi < g.r; ++phi)
        if (std::abs(g.r * (g.smin + t) + psi - phi) <= hk) mask |= 1u << (psi * channels / 64);  // 64-channel chunks
    p.kmask[t] = plain ? 3u : mask;   // plain 128-channel rows: both K-chunks of every tap
  }
  p.rows_rho = (L + p.d * g.r - 1) / (p.d * g.r);
  p.m_tiles = (p.Lf + p.WO - 1) / p.WO;
  p.total_tiles = B * p.m_tiles;
  p.div_m.init(p.m_tiles);
  p.div_dr.init(p.d * g.r);
  p.x = x;
  p.trace = nullptr;
  pl->channels = channels;
  pl->grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  const int rowb = 128;
  pl->smem = 1024 + (size_t)2 * g.buf + (size_t)g.nw * 128 * rowb + 256 + 1024 + 16384;
  if (p.d > 1) {
    // [C][rho][phase][row][utterance]: sample t = dil*(r*row + phase) + rho
    // a K-chunk = 64/C phases: box {C, 1, 64/C, XR, 1} lands as 128-byte rows [phase][channel]
    if (encode_tmap_act(&pl->tmX, x, channels, dil, channels, g.r, (uint64_t)dil * channels, p.rows_rho,
                        (uint64_t)dil * g.r * channels, B, (uint64_t)L * channels, g.XR, 64 / channels))
      return 1;
  } else {
    if (encode_tmap_act(&pl->tmX, x, 64, 2, 64, 1, 128, p.Lf, 128, B, (uint64_t)L * channels, g.XR, 1)) return 1;
  }
  if (encode_tmap_3d(&pl->tmW, w_fold, 128, 128, 2 * g.nt, 64, 128, true)) return 1;
  return 0;
}

template <int CH, bool F16>
static int launch_pairf_typed(const PairFPlan& pl, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VD_CUDA(cudaFuncSetAttribute(conv_pairf_kernel<CH, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  conv_pairf_kernel<CH, F16><<<pl.grid, kPfThreads, pl.smem, stream>>>(pl.tmX, pl.tmW, pl.p);
  VD_CUDA(cudaGetLastError());
  return 0;
}

template <int CH>
static int launch_pairf_inst(const PairFPlan& pl, cudaStream_t stream) {
  return pl.p.f16 ? launch_pairf_typed<CH, true>(pl, stream) : launch_pairf_typed<CH, false>(pl, stream);
}

int launch_conv_pairf(PairFPlan& pl, const float* bias1, const float* bias2, float slope, __nv_bfloat16* out,
                      cudaStream_t stream, int f16) {
  pl.p.f16 = f16;
  pl.p.bias1 = bias1;
  pl.p.bias2 = bias2;
  pl.p.slope = slope;
  pl.p.res_gain = 1.f / slope;
  pl.p.out = out;
  if (pl.channels == 32) return launch_pairf_inst<32>(pl, stream);
  if (pl.channels == 64) return launch_pairf_inst<64>(pl, stream);
  if (pl.channels == 128) return launch_pairf_inst<128>(pl, stream);
  set_error("conv_pairf: no kernel instance");
  return 1;
}

}  // namespace vd
