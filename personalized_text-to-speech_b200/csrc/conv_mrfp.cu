// Fused ResBlock1 pairs of the C = 32 stage on the sm_100a tensor cores, on the 2-sample time-folded view:
//     X = lrelu( ( sum_j  c2_j( lrelu( c1_j(P_j) + b1_j ) ) + b2_j + x(P_j) ) / nbr , out_slope )
//   nbr = 1: one ResBlock1 iteration (modules.py:211-221)       c1 = Conv1d(C, C, k, dilation d), c2 = Conv1d(C, C, k, 1)
//   nbr = 3: the LAST iteration of every MRF branch of the stage, the branch sum, the 1/nk average and the next
//            leaky-relu (models.py:278-284) -- three h tiles that never leave the SM, one output tensor.
//
// Why the folded view.  With 32 channels a time-as-M MMA (M = 128 samples, N = 32, K = 16) reads 4 KB of activations
// + 1 KB of weights from shared memory for 16 cycles of math: measured 40 cycles (tools/probes/mma_rate_probe.cu:
// cycles = 32 + N/4 below N = 128, the operand fetch at 128 B/cycle), 40 % of the tensor rate at best, and the fused
// kernels were bound by the hand-offs between MMA issue and the epilogue groups per 256-sample tile on top of that.
// [B][L][32] is bit for bit [B][L/2][64]: row m holds samples 2m, 2m+1 ("phases").  On that view
//   * a dilation-1 conv is a conv over folded rows with N = 64: phase-0 outputs take tap t0, phase-1 outputs tap
//     t0 - 1 of the SAME input chunk (row shift s, input phase psi; t0 = 2s + psi + hk), so with the taps stored in
//     reverse order two neighbouring tap blocks ARE the stacked 64-row B operand: k + 1 chunks, no extra weights, 48
//     cycles per MMA for twice the samples of the 40-cycle N = 32 form;
//   * a dilated conv (odd dilation flips the phase) runs as 2k N = 32 block jobs (tap, output phase) -> (s, psi);
//   * a tile of 2 x 128 folded rows covers 512 samples: twice the work per hand-off of the plain kernels.
// MMA jobs are a host-built table (MrfpParams::jobs): the issuing warps just walk it.
//
// Per CTA (tile i, branch j; n = running (i, j) index):
//   TMA A(n) -> c1(n) [acc1, 2 TMEM buffers] -> epi1: h(n) = lrelu(acc1 + b1_j), 0 outside the utterance -> smem
//   [swizzled K-major A operand of c2, 1-2 buffers] -> c2(n) [+= acc2(i), 2 TMEM buffers]
//   after the last branch: epi2(i) = (acc2 + sum b2 + sum_j x(P_j)) / nbr -> lrelu -> global.
// The residual rows x(P_j) come from the resident activation tiles when shared memory has a stage to spare beyond the
// tile's own (single pairs: the tile stays until its output epilogue has read it); with three branches there is room for
// two stages only, so a stage is released as soon as its c1 has retired and the residual rows are re-read from global
// memory (they hit L2 there: 679 MB of DRAM reads for 678 MB of inputs; for the short k = 3 pairs the same re-read
// missed L2 for 3/4 of the rows and exposed the DRAM latency once per item).  Two MMA-issuing warps (c1 stream / c2 stream), two epilogue groups of
// 8 warps (h producer / output), as in conv_pair.cu.
#include <algorithm>

#include "common.cuh"
#include "conv_mrfp.h"
#include "ptx.cuh"

namespace vd {

#ifndef VITSDEC_TRACE
#define VITSDEC_TRACE 0
#endif
constexpr bool kMpTrace = VITSDEC_TRACE != 0;   // tools/trace_mrfp.py: stamps 0/1 c1 issue, 2/3 c2 issue, 4/5 epi1, 6/7 epi2,
                                                // 8 epi2 wait begin, 9 epi1 wait begin, 10 TMA issued, 11 c1 wait begin
constexpr int kMpEpiWarps = 16;
constexpr int kMpThreads = 64 + 32 * kMpEpiWarps + 32;  // producer, c1 issuer, 16 epilogue warps, c2 issuer
constexpr int kMpC2Warp = 2 + kMpEpiWarps;
constexpr int kMpRowB = 128;                            // folded row: 2 samples x 32 channels x 2 bytes
constexpr int kMpWBlock = 32 * 64;                      // one tap: [32 out][32 in] bf16, 64-byte rows (SWIZZLE_64B)

struct MrfpMaps {
  CUtensorMap a[kMpMaxBr];
};

template <int NACC, bool F16>
__global__ void __launch_bounds__(kMpThreads, 1)
conv_mrfp_kernel(const __grid_constant__ MrfpMaps tm, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ MrfpParams p) {
  constexpr int ROWB = kMpRowB;
  constexpr int MROWS = 128 * NACC;       // folded rows per tile (h rows; output rows incl. the invalid margin)
  constexpr int ACC_COLS = NACC * 64;     // NACC accumulators of 128 rows x (2 phases x 32 channels)
  constexpr int TMEM_COLS = 4 * ACC_COLS; // acc1[2] + acc2[2]
  static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns must be a power of two");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NA = p.na_stages, NH = p.nh;
  const int hrows_bytes = (MROWS + 2 * p.hm + 7) / 8 * 8 * ROWB;   // h buffer: MROWS written rows + the rows c2's taps reach
  uint8_t* smemA = smem;
  uint8_t* smemH = smemA + NA * p.a_stage_bytes;
  uint8_t* smemW = smemH + NH * hrows_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemW + p.wbytes);
  uint64_t* a_full = bars;                  // [kMpMaxNA]
  uint64_t* a_empty = a_full + kMpMaxNA;    // [kMpMaxNA]
  uint64_t* acc1_full = a_empty + kMpMaxNA;
  uint64_t* acc1_empty = acc1_full + 2;
  uint64_t* acc2_full = acc1_empty + 2;
  uint64_t* acc2_empty = acc2_full + 2;
  uint64_t* h_full = acc2_empty + 2;
  uint64_t* h_empty = h_full + 2;
  uint64_t* w_full = h_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  float* sbias = reinterpret_cast<float*>(bars + 32);    // 256 B of barriers, then (kMpMaxBr + 1) * 64 floats (per column)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nbr = p.nbr;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    for (int j = 0; j < nbr; ++j) tma_prefetch_desc(&tm.a[j]);
    tma_prefetch_desc(&tmW);
    // an activation stage is free when c1 has retired -- and, when the residual is read from it, when the 8 output
    // epilogue warps are done with it too
    for (int i = 0; i < kMpMaxNA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], p.res_smem ? 1 + kMpEpiWarps / 2 : 1); }
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
  for (int i = threadIdx.x; i < (kMpMaxBr + 1) * 64; i += kMpThreads) {   // bias per accumulator column
    const int j = i >> 6, c = p.fold == 2 ? (i & 31) : (i & 63);           // folded view: column = phase * 32 + channel
    sbias[i] = j < kMpMaxBr ? (j < nbr ? p.bias1[j][c] : 0.f) : p.bias2sum[c];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int hm = p.hm;
  const int my_tiles = p.total_tiles > (int)blockIdx.x ? (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      // weights: every tap block once; within a conv the blocks are stored in REVERSE tap order, so that the blocks of
      // taps t0, t0 - 1 form the stacked [64][32] operand of a dilation-1 chunk
      mbar_expect_tx(w_full, p.wbytes);
      if (p.fold == 2) {
        for (int c = 0; c < 2 * nbr; ++c)
          for (int t = 0; t < p.conv_k[c]; ++t)
            tma_load_3d(&tmW, w_full, smemW + (p.conv_base[c] + p.conv_k[c] - 1 - t) * kMpWBlock, 0, 0, p.conv_base[c] + t);
      } else {   // C = 64: a tap is two [64 out][32 in] K-halves of 4 KB, natural order
        for (int t = 0; t < p.ntaps; ++t)
          for (int kh = 0; kh < 2; ++kh) tma_load_3d(&tmW, w_full, smemW + (2 * t + kh) * 2 * kMpWBlock, kh * 32, 0, t);
      }
      pdl_wait();   // the weights (static) load while the previous launch drains; activations only from here on
      uint32_t sa = 0, pa = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const uint32_t tile = blockIdx.x + i * gridDim.x;
        uint32_t b, mt;
        p.div_m.divmod(tile, b, mt);
        const int m0 = mt * p.bmo;
        for (int j = 0; j < nbr; ++j) {
          mbar_wait(&a_empty[sa], pa ^ 1);
          if (kMpTrace && p.trace && blockIdx.x == 0 && i * nbr + j < 256) p.trace[(i * nbr + j) * 12 + 10] = clock64();
          mbar_expect_tx(&a_full[sa], p.a_boxes[j] * 32 * ROWB);
          for (int bx = 0; bx < p.a_boxes[j]; ++bx)
            tma_load_3d(&tm.a[j], &a_full[sa], smemA + sa * p.a_stage_bytes + bx * 32 * ROWB, 0,
                        m0 + p.a_lo[j] + bx * 32, (int)b);
          if (++sa == (uint32_t)NA) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == kMpC2Warp) {
    // ------------------------------------------------------------ MMA issuers (warp-uniform; elected lane issues)
    constexpr uint32_t idesc32 = umma_idesc_f16(32, F16), idesc64 = umma_idesc_f16(64, F16);
    constexpr uint32_t a_hi = umma_desc_hi(ROWB);    // activations / h: 128-byte rows, SWIZZLE_128B
    constexpr uint32_t w_hi = umma_desc_hi(64);      // weight blocks: 64-byte rows, SWIZZLE_64B
    const uint32_t leader = elect_one();
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(smemA)), w_lo0 = umma_desc_lo(smem_u32(smemW));
    const uint32_t h_lo0 = umma_desc_lo(smem_u32(smemH));
    const uint32_t a_stage16 = (uint32_t)p.a_stage_bytes >> 4;
    const uint32_t h_buf16 = (uint32_t)hrows_bytes >> 4;
    mbar_wait(w_full, 0);
    tc_fence_after();
    // all MMAs of one conv on one operand tile: NACC accumulators x K = 32 per job
    auto run_jobs = [&](int jb, int je, uint32_t src_lo, uint32_t d_base, bool always_acc) {
      for (int q = jb; q < je; ++q) {
        const MpJob job = p.jobs[q];
        const uint32_t idesc = (job.flags & 1) ? idesc64 : idesc32;
        const uint32_t init = (job.flags & 2) && !always_acc ? 0u : 1u;
        const uint32_t a0 = src_lo + job.a_off16, w0 = w_lo0 + job.w_off16, d0 = d_base + job.d_off;
#pragma unroll
        for (int acc = 0; acc < NACC; ++acc)
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
            umma_f16_lohi(d0 + acc * 64, a0 + ((acc * 128 * ROWB + kk * 32) >> 4), a_hi, w0 + ((kk * 32) >> 4), w_hi, idesc,
                          kk == 0 ? init : 1u, leader);
      }
    };
    if (warp == 1) {
      // c1 stream: acc1[n & 1] = c1_j(A(n))
      uint32_t sa = 0, pa = 0, n = 0;
      for (int i = 0; i < my_tiles; ++i) {
        for (int j = 0; j < nbr; ++j, ++n) {
          const uint32_t as = n & 1;
          const bool tr = kMpTrace && p.trace && blockIdx.x == 0 && n < 256 && lane == 0;
          if (tr) p.trace[n * 12 + 11] = clock64();
          mbar_wait(&acc1_empty[as], ((n >> 1) & 1) ^ 1);
          mbar_wait(&a_full[sa], pa);
          tc_fence_after();
          if (tr) p.trace[n * 12 + 0] = clock64();
          run_jobs(p.job_beg[2 * j], p.job_beg[2 * j + 1], a_lo0 + sa * a_stage16, tmem_base + as * ACC_COLS, false);
          if (leader) {
            umma_commit(&acc1_full[as]);
            umma_commit(&a_empty[sa]);
          }
          if (tr) p.trace[n * 12 + 1] = clock64();
          if (++sa == (uint32_t)NA) { sa = 0; pa ^= 1; }
        }
      }
    } else {
      // c2 stream: acc2[i & 1] (+)= c2_j(h(n)); the branch sum is formed in the accumulator
      uint32_t hb = 0, ph = 0, n = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const uint32_t as = i & 1;
        for (int j = 0; j < nbr; ++j, ++n) {
          mbar_wait(&h_full[hb], ph);
          if (j == 0) mbar_wait(&acc2_empty[as], ((i >> 1) & 1) ^ 1);
          tc_fence_after();
          const bool tr = kMpTrace && p.trace && blockIdx.x == 0 && n < 256 && lane == 0;
          if (tr) p.trace[n * 12 + 2] = clock64();
          run_jobs(p.job_beg[2 * j + 1], p.job_beg[2 * j + 2], h_lo0 + hb * h_buf16, tmem_base + 2 * ACC_COLS + as * ACC_COLS,
                   j > 0);
          if (leader) {
            umma_commit(&h_empty[hb]);
            if (j == nbr - 1) umma_commit(&acc2_full[as]);
          }
          if (tr) p.trace[n * 12 + 3] = clock64();
          if (++hb == (uint32_t)NH) { hb = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int q = warp & 3;
    const int grp = (warp - 2) >> 3;          // 0: epi1 (h producer), 1: epi2 (output)
    const int hsel = ((warp - 2) & 7) >> 2;   // which of the group's two warps on this TMEM lane quadrant
    const int ch0 = hsel * 16;                // this warp's 16 channels, of both time phases (columns ch0 and 32 + ch0)
    const int Lf = p.Lf;
    const float slope = p.slope, res_gain = p.res_gain;
    pdl_wait();   // output stores may overwrite a buffer the previous launch still reads

    if (grp == 0) {
      // h(n) = lrelu(c1_j + b1_j), zero outside the utterance, written as c2's swizzled K-major A operand
      uint32_t n = 0, hb = 0, ph = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const uint32_t tile = blockIdx.x + i * gridDim.x;
        uint32_t b, mt;
        p.div_m.divmod(tile, b, mt);
        const int m0 = mt * p.bmo;
        const bool edge_tile = m0 - hm < 0 || m0 - hm + MROWS > Lf;   // warp-uniform
#pragma unroll
        for (int j = 0; j < kMpMaxBr; ++j) {
          if (j >= nbr) break;
          const uint32_t as = n & 1;
          uint8_t* const hbuf = smemH + hb * hrows_bytes;
          // biases of this warp's two 16-column chunks: shared-memory reads issued before the wait, which hides them
          float4 breg[2][4];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh)
#pragma unroll
            for (int e = 0; e < 4; ++e) breg[hh][e] = *reinterpret_cast<const float4*>(sbias + j * 64 + ch0 + hh * 32 + 4 * e);
          const bool tr = kMpTrace && p.trace && blockIdx.x == 0 && warp == 2 && lane == 0 && n < 256;
          if (tr) p.trace[n * 12 + 9] = clock64();
          mbar_wait(&acc1_full[as], (n >> 1) & 1);
          tc_fence_after();
          if (tr) p.trace[n * 12 + 4] = clock64();
          bool h_free = false;
#pragma unroll
          for (int it = 0; it < 2 * NACC; ++it) {
            const int acc = it >> 1, c0 = ch0 + (it & 1) * 32;   // column chunk: (phase it & 1, channels ch0 ..)
            const int r = acc * 128 + q * 32 + lane;
            const int fr = m0 - hm + r;
            uint32_t a[16];
            __syncwarp();
            tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + as * ACC_COLS + acc * 64 + c0, a);
            tmem_ld_wait();
            // leaky-relu as max(v, slope * v) (slope in (0, 1]); rows outside the utterance are c2's zero padding -- only
            // the first / last tiles of an utterance have any, the others skip the per-element select
            const bool inside = fr >= 0 && fr < Lf;
            const float2 sl2 = make_float2(slope, slope);
            uint4 o[2];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              uint32_t* o2 = reinterpret_cast<uint32_t*>(&o[h2]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int jj = h2 * 8 + e * 2;
                const float4 bq = breg[it & 1][jj >> 2];
                const float2 v = fadd2(make_float2(__uint_as_float(a[jj]), __uint_as_float(a[jj + 1])),
                                       (jj & 3) == 0 ? make_float2(bq.x, bq.y) : make_float2(bq.z, bq.w));
                const float2 t = fmul2(v, sl2);
                o2[e] = pack_act2<F16>(fmaxf(v.x, t.x), fmaxf(v.y, t.y));
              }
            }
            if (edge_tile && !inside) o[0] = o[1] = make_uint4(0, 0, 0, 0);
            if (!h_free) {  // the c2 that last read this h buffer must have retired before it is overwritten
              mbar_wait(&h_empty[hb], ph ^ 1);
              h_free = true;
            }
            const uint32_t sw = r & 7;   // 128-byte rows: 16-byte chunk index XOR address bits [7, 10)
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
          if (tr) p.trace[n * 12 + 5] = clock64();
          ++n;
          if (++hb == (uint32_t)NH) { hb = 0; ph ^= 1; }
        }
      }
    } else {
      // X = lrelu((acc2 + sum b2 + sum_j x(P_j)) / nbr): residual rows re-read from global memory (L2)
      float4 b2r[2][4];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int e = 0; e < 4; ++e) b2r[hh][e] = *reinterpret_cast<const float4*>(sbias + kMpMaxBr * 64 + ch0 + hh * 32 + 4 * e);
      const float scale = p.scale, out_slope = p.out_slope;
      __nv_bfloat16* const out = p.out;
      const __nv_bfloat16* const res0 = p.res[0];
      const __nv_bfloat16* const res1 = p.res[1];
      const __nv_bfloat16* const res2 = p.res[2];
      const bool res_smem = p.res_smem != 0;
      uint32_t sa = 0;   // activation stage of (tile i, branch 0)
      for (int i = 0; i < my_tiles; ++i) {
        const uint32_t tile = blockIdx.x + i * gridDim.x;
        uint32_t b, mt;
        p.div_m.divmod(tile, b, mt);
        const int m0 = mt * p.bmo;
        const uint32_t as = i & 1;
        bool waited = false;
        const int ntr = i * nbr + nbr - 1;
        const bool tr = kMpTrace && p.trace && blockIdx.x == 0 && warp == 2 + kMpEpiWarps / 2 && lane == 0 && ntr < 256;
        if (tr) p.trace[ntr * 12 + 8] = clock64();
#pragma unroll
        for (int it = 0; it < 2 * NACC; ++it) {
          const int acc = it >> 1, c0 = ch0 + (it & 1) * 32;
          const int row = acc * 128 + q * 32 + lane;      // output folded row of this thread
          const bool valid = row < p.bmo && m0 + row < Lf;
          const long off = ((long)b * Lf + m0 + row) * 64 + c0;
          // residual rows of the branches: this thread's 16 columns are one aligned 32-byte sector
          uint4 rx[kMpMaxBr][2];
          if (!res_smem) {
#pragma unroll
            for (int j = 0; j < kMpMaxBr; ++j) {
              rx[j][0] = rx[j][1] = make_uint4(0, 0, 0, 0);
              if (j < nbr && valid) ld_stream_v8((j == 0 ? res0 : (j == 1 ? res1 : res2)) + off, rx[j][0], rx[j][1]);
            }
          }
          if (!waited) {
            mbar_wait(&acc2_full[as], (i >> 1) & 1);   // (implies that every c1 of the tile has retired: the tiles are complete)
            tc_fence_after();
            waited = true;
            if (tr) p.trace[ntr * 12 + 6] = clock64();
          }
          uint32_t a[16];
          __syncwarp();
          tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + 2 * ACC_COLS + as * ACC_COLS + acc * 64 + c0, a);
          if (res_smem) {
            uint32_t st = sa;
#pragma unroll
            for (int j = 0; j < kMpMaxBr; ++j) {
              if (j >= nbr) break;
              const uint8_t* atile = smemA + st * p.a_stage_bytes;
              const int ra = row - p.a_lo[j];             // row of this output row's own samples in branch j's tile
              const uint32_t sw = ra & 7;
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2)
                rx[j][h2] = *reinterpret_cast<const uint4*>(atile + ra * ROWB + ((((c0 >> 3) + h2) ^ sw) << 4));
              if (++st == (uint32_t)NA) st = 0;
            }
          }
          tmem_ld_wait();
          // same summation order as the unfused schedule's epilogues (epilogue.cuh): acc + bias, + residuals in branch
          // order, * 1/nbr; packed fp32 adds / multiplies (two elements per issue slot)
          const float4 (&b2)[4] = b2r[it & 1];
          float2 v[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[2 * e] = fadd2(make_float2(__uint_as_float(a[4 * e]), __uint_as_float(a[4 * e + 1])), make_float2(b2[e].x, b2[e].y));
            v[2 * e + 1] = fadd2(make_float2(__uint_as_float(a[4 * e + 2]), __uint_as_float(a[4 * e + 3])),
                                 make_float2(b2[e].z, b2[e].w));
          }
          const float2 g2 = make_float2(res_gain, res_gain);
#pragma unroll
          for (int j = 0; j < kMpMaxBr; ++j) {
            if (j >= nbr) break;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const uint32_t* r2 = reinterpret_cast<const uint32_t*>(&rx[j][h2]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 xr = unpack_act2<F16>(r2[e]);
                const float2 xg = fmul2(xr, g2);
                // a-form -> residual stream: a >= 0 ? a : a * res_gain  ==  min(a, a * res_gain) for res_gain >= 1
                v[h2 * 4 + e] = fadd2(v[h2 * 4 + e], make_float2(fminf(xr.x, xg.x), fminf(xr.y, xg.y)));
              }
            }
          }
          const float2 os2 = make_float2(out_slope, out_slope), sc2 = make_float2(scale, scale);
          uint4 ov[2];
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            uint32_t* o2 = reinterpret_cast<uint32_t*>(&ov[h2]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float2 w = v[h2 * 4 + e];
              if (nbr > 1) w = fmul2(w, sc2);     // (x 1 would be exact anyway; a single pair skips the instruction)
              const float2 t = fmul2(w, os2);
              o2[e] = pack_act2<F16>(fmaxf(w.x, t.x), fmaxf(w.y, t.y));
            }
          }
          if (valid) st_global_v8(out + off, ov[0], ov[1]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&acc2_empty[as]);
          if (res_smem) {
            uint32_t st = sa;
            for (int j = 0; j < nbr; ++j) {
              mbar_arrive(&a_empty[st]);
              if (++st == (uint32_t)NA) st = 0;
            }
          }
        }
        if (tr) p.trace[ntr * 12 + 7] = clock64();
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

static int mp_floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

struct MpGeom {
  int hm, ntaps, a_stage, na, nh, njobs;
  int lo[kMpMaxBr], rows[kMpMaxBr];   // halo rows before the h tile / rows of the activation tile (multiple of 32)
  bool ok;
};

static MpGeom mp_geom(int channels, int nbr, const int* k, const int* dil, int nacc) {
  MpGeom g{};
  g.ok = false;
  if ((channels != 32 && channels != 64) || nbr < 1 || nbr > kMpMaxBr) return g;
  const int fold = 64 / channels;   // samples per row
  const int mrows = 128 * nacc;
  int sumk = 0;
  for (int j = 0; j < nbr; ++j) {
    if (k[j] % 2 == 0 || k[j] < 1 || k[j] > 15 || dil[j] < 1 || dil[j] > 16) return g;
    const int hk = (k[j] - 1) / 2;
    // sample offsets -d*hk .. d*hk (+ fold - 1 for the last phase) -> rows floor(-d*hk / fold) .. floor((fold - 1 + d*hk) / fold)
    g.hm = std::max(g.hm, -mp_floordiv(-hk, fold));
    g.lo[j] = -mp_floordiv(-dil[j] * hk, fold);
    const int hi = mp_floordiv(fold - 1 + dil[j] * hk, fold);
    g.rows[j] = (mrows + g.lo[j] + hi + 31) / 32 * 32;
    g.a_stage = std::max(g.a_stage, g.rows[j] * kMpRowB);
    sumk += k[j];
    g.njobs += fold == 2 ? (dil[j] == 1 ? k[j] + 1 : 2 * k[j]) + k[j] + 1 : 4 * k[j];
  }
  if (g.njobs > kMpMaxJobs) return g;
  g.ntaps = 2 * sumk;
  const int hbytes = (mrows + 2 * g.hm + 7) / 8 * 8 * kMpRowB;
  const int wbytes = g.ntaps * channels * channels * 2;
  // two h buffers when two activation stages still fit, else one
  g.nh = (wbytes + 2 * hbytes + 2 * g.a_stage <= kMpSmemBudget) ? 2 : 1;
  g.na = std::min(kMpMaxNA, (kMpSmemBudget - wbytes - g.nh * hbytes) / g.a_stage);
  if (g.na < 2) return g;
  g.ok = true;
  return g;
}

bool mrfp_supported(int channels, int nbr, const int* k, const int* dil) { return mp_geom(channels, nbr, k, dil, 2).ok; }

// jobs of one conv (taps k, dilation d) reading an operand buffer whose row 0 is `row_base` rows before the first
// output row's own row; the conv's weights start at tap `wbase` of the set in shared memory
static void mp_conv_jobs(MrfpParams& p, int& nj, int k, int d, int row_base, int wbase) {
  const int hk = (k - 1) / 2;
  auto add = [&](int s, int psi, int w_byte_off, int d_off, bool n64, bool init) {
    MpJob& j = p.jobs[nj++];
    j.a_off16 = (uint16_t)(((s + row_base) * kMpRowB + psi * 64) >> 4);
    j.w_off16 = (uint16_t)(w_byte_off >> 4);
    j.d_off = (uint8_t)d_off;
    j.flags = (uint8_t)((n64 ? 1 : 0) | (init ? 2 : 0));
    j.pad = 0;
  };
  if (p.fold == 1) {
    // C = 64, plain rows: tap t is two K = 32 jobs (input channels 0-31 / 32-63) with [64 out][32 in] weight blocks
    for (int t = 0; t < k; ++t)
      for (int kh = 0; kh < 2; ++kh) add((t - hk) * d, kh, ((wbase + t) * 2 + kh) * 2 * kMpWBlock, 0, true, t == 0 && kh == 0);
    return;
  }
  auto blk = [&](int t) { return (wbase + k - 1 - t) * kMpWBlock; };   // C = 32: reverse tap order in shared memory
  if (d == 1) {
    // chunk (s, psi) <-> t0 = 2s + psi + hk: phase-0 outputs use tap t0, phase-1 outputs tap t0 - 1
    auto sp = [&](int t0, int& s, int& psi) { s = mp_floordiv(t0 - hk, 2); psi = (t0 - hk) - 2 * s; };
    int s, psi;
    sp(0, s, psi);  add(s, psi, blk(0), 0, false, true);          // t0 = 0: only phase 0 (tap 0)
    sp(k, s, psi);  add(s, psi, blk(k - 1), 32, false, true);     // t0 = k: only phase 1 (tap k - 1)
    for (int t0 = 1; t0 < k; ++t0) {                              // interior chunks: blocks of taps t0, t0 - 1 are adjacent
      sp(t0, s, psi);
      add(s, psi, blk(t0), 0, true, false);
    }
  } else {
    bool started[2] = {false, false};
    for (int t = 0; t < k; ++t)
      for (int phi = 0; phi < 2; ++phi) {
        const int off = phi + d * (t - hk);
        const int s = mp_floordiv(off, 2), psi = off - 2 * s;
        add(s, psi, blk(t), phi * 32, false, !started[phi]);
        started[phi] = true;
      }
  }
}

int plan_conv_mrfp(MrfpPlan* pl, int B, int L, int channels, int nbr, const int* k, const int* dil,
                   const __nv_bfloat16* const* xs, const __nv_bfloat16* w, int num_sms) {
  VD_CHECK(channels == 32 || channels == 64, "conv_mrfp: 32 or 64 channels");
  const int fold = 64 / channels;
  VD_CHECK(L % fold == 0, "conv_mrfp: the utterance length must be even (2-sample folded view)");
  const int Lf = L / fold;
  // tiles of 256 rows; small problems (fewer tiles than SMs) take 128-row tiles to occupy more SMs
  int nacc = 2;
  {
    const MpGeom g2 = mp_geom(channels, nbr, k, dil, 2);
    VD_CHECK(g2.ok, "conv_mrfp: unsupported shape");
    const int bmo2 = 256 - 2 * g2.hm;
    if ((long)B * ((Lf + bmo2 - 1) / bmo2) < num_sms && mp_geom(channels, nbr, k, dil, 1).ok) nacc = 1;
  }
  const MpGeom g = mp_geom(channels, nbr, k, dil, nacc);
  MrfpParams& p = pl->p;
  p = MrfpParams{};
  p.B = B; p.Lf = Lf; p.nbr = nbr;
  p.fold = fold;
  p.hm = g.hm;
  p.bmo = 128 * nacc - 2 * g.hm;
  int tap = 0;
  for (int j = 0; j < nbr; ++j) {           // global packed order: all c1 convs, then all c2 convs
    p.conv_base[2 * j] = tap; p.conv_k[2 * j] = k[j];
    tap += k[j];
  }
  for (int j = 0; j < nbr; ++j) {
    p.conv_base[2 * j + 1] = tap; p.conv_k[2 * j + 1] = k[j];
    tap += k[j];
  }
  p.ntaps = tap;
  p.wbytes = tap * channels * channels * 2;
  int nj = 0;
  for (int j = 0; j < nbr; ++j) {
    p.a_lo[j] = -(g.hm + g.lo[j]);
    p.a_boxes[j] = g.rows[j] / 32;
    p.job_beg[2 * j] = nj;
    mp_conv_jobs(p, nj, k[j], dil[j], g.lo[j], p.conv_base[2 * j]);       // c1: tile row = h row + s + lo
    p.job_beg[2 * j + 1] = nj;
    mp_conv_jobs(p, nj, k[j], 1, g.hm, p.conv_base[2 * j + 1]);          // c2: h row = output row + s + hm
  }
  p.job_beg[2 * nbr] = nj;
  VD_CHECK(nj <= kMpMaxJobs, "conv_mrfp: job table overflow");
  p.a_stage_bytes = g.a_stage;
  p.na_stages = g.na;
  p.nh = g.nh;
  p.res_smem = g.na > nbr ? 1 : 0;   // room to keep a tile's stages until its output epilogue and still prefetch
  p.m_tiles = (p.Lf + p.bmo - 1) / p.bmo;
  p.total_tiles = B * p.m_tiles;
  p.div_m.init(p.m_tiles);
  p.scale = 1.f / nbr;
  for (int j = 0; j < kMpMaxBr; ++j) p.res[j] = xs[j < nbr ? j : 0];
  pl->channels = channels;
  pl->nacc = nacc;
  pl->pdl = false;
  pl->grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  const int hbytes = (128 * nacc + 2 * g.hm + 7) / 8 * 8 * kMpRowB;
  pl->smem = 1024 + (size_t)p.na_stages * p.a_stage_bytes + (size_t)p.nh * hbytes + (size_t)p.wbytes + 256 + 1024;
  for (int j = 0; j < kMpMaxBr; ++j) {
    if (j < nbr) {   // rows of 64 (virtual) channels: [B][L / fold][64], 32-row boxes, 128-byte rows
      if (encode_tmap_3d(&pl->tmA[j], xs[j], 64, Lf, B, 64, 32, true)) return 1;
    } else {
      pl->tmA[j] = pl->tmA[0];
    }
  }
  // weight blocks with 64-byte rows: [32 out][32 in] per tap (C = 32), two [64 out][32 in] K-halves per tap (C = 64)
  if (encode_tmap_3d(&pl->tmW, w, channels, channels, p.ntaps, 32, channels, true)) return 1;
  return 0;
}

template <int NACC, bool F16>
static int launch_mrfp_typed(const MrfpPlan& pl, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VD_CUDA(cudaFuncSetAttribute(conv_mrfp_kernel<NACC, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const size_t smem = pl.smem > (size_t)120 * 1024 ? pl.smem : (size_t)120 * 1024;   // one CTA per SM, see conv_tc.cu
  MrfpMaps maps;
  for (int j = 0; j < kMpMaxBr; ++j) maps.a[j] = pl.tmA[j];
  if (pl.pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pl.grid);
    cfg.blockDim = dim3(kMpThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    VD_CUDA(cudaLaunchKernelEx(&cfg, conv_mrfp_kernel<NACC, F16>, maps, pl.tmW, pl.p));
    return 0;
  }
  conv_mrfp_kernel<NACC, F16><<<pl.grid, kMpThreads, smem, stream>>>(maps, pl.tmW, pl.p);
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
  VD_CHECK(pl.channels == 32 || pl.channels == 64, "conv_mrfp: no kernel instance");
  if (pl.nacc == 2) return f16 ? launch_mrfp_typed<2, true>(pl, stream) : launch_mrfp_typed<2, false>(pl, stream);
  return f16 ? launch_mrfp_typed<1, true>(pl, stream) : launch_mrfp_typed<1, false>(pl, stream);
}

}  // namespace vd
