// Implicit-GEMM convolution on the sm_100a tensor cores (tcgen05.mma, accumulators in TMEM, operands
// staged by TMA).  One kernel serves conv_pre, the polyphase ConvTranspose1d stages and every ResBlock
// conv of the decoder (reference: models.py:271-287, modules.py:210-223).
//
// Mapping.  Activations are channels-last bf16 [B][L][C]; a CTA tile is BM = 128*NACC time rows x BN
// output columns.  D[time, n] += A[time, ci] * W[n, ci]^T per (tap, 64-channel chunk):
//   * A operand  = the activation tile, K-major (channels contiguous), 128B-swizzled rows.  ONE TMA
//     load of rows [t0+halo_lo, t0+BM+halo_hi) serves every tap: a tap is a row offset of the UMMA
//     shared-memory descriptor into the same tile (no im2col, no per-tap reload).  TMA out-of-bounds
//     zero fill implements the zero padding at both utterance ends.
//   * B operand  = packed weights W[tap][n][ci], streamed through an NB-stage ring.
//   * D          = fp32 in TMEM, double buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue
// (tcgen05.ld -> bias / residual / MRF / leaky-relu -> bf16 channels-last stores).
#include <algorithm>

#include "common.cuh"
#include "conv_pairf.h"
#include "conv_tc.h"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace vd {

constexpr int kMaxNA = 8;      // activation-chunk stages (runtime count <= this)
constexpr int kMaxNB = 8;      // weight-tile stages when weights are streamed
#ifndef VITSDEC_TRACE
#define VITSDEC_TRACE 0   // 1: build with the per-tile clock64 trace hooks (tools/trace_probe.py); costs ~10 % in the epilogue
#endif
constexpr bool kTrace = VITSDEC_TRACE != 0;
constexpr int kEpiWarps = 16;  // four warps per TMEM lane quadrant: the epilogue is instruction-latency bound, TLP hides it
constexpr int kTcThreads = 64 + 32 * kEpiWarps;
constexpr int kSmemBudget = 227 * 1024 - 1024 /*alignment slack*/ - 512 /*barriers*/ - 8192 /*bias*/ - 32768 /*scratch*/;

template <int BN, int KC>
struct TcCfg {
  static constexpr int NACC = BN >= 256 ? 1 : 2;
  static constexpr int ROWB = KC * 2;
  static constexpr int B_STAGE = BN * ROWB;
  static constexpr int ACC_COLS = NACC * BN;
  static constexpr int RBOXC = BN < 64 ? BN : 64;   // residual prefetch box: RBOXC columns x 64 rows
  static constexpr int NBUF = 512 / ACC_COLS > 8 ? 8 : 512 / ACC_COLS;  // accumulator buffers in TMEM (2..8)
  static constexpr int TMEM_COLS = NBUF * ACC_COLS;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns must be a power of two <= 512");
};

template <int BN, int KC, int EPI, bool SWAP, bool F16>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ TmapPack tm, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ ConvTcParams p) {
  using C = TcCfg<BN, KC>;
  constexpr int NACC = C::NACC, ROWB = C::ROWB, B_STAGE = C::B_STAGE, ACC_COLS = C::ACC_COLS;
  constexpr int BM = 128 * NACC;
  // channels-as-M tiles cover swap_rows (256, 128 or 64) time rows: small decodes use narrower tiles to occupy more SMs
  const int BMr = SWAP ? p.swap_rows : BM;

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: the 128B swizzle pattern is a function of the shared-memory address bits
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smemA = smem;
  const int NA = p.na_stages, NB = p.nb_stages;
  uint8_t* smemB = smem + NA * p.a_stage_bytes;   // streamed ring, or the whole [tap][kc] weight set (stationary)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemB + p.b_region_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kMaxNA;
  uint64_t* b_full = a_empty + kMaxNA;
  uint64_t* b_empty = b_full + kMaxNB;
  uint64_t* acc_full = b_empty + kMaxNB;
  uint64_t* acc_empty = acc_full + 8;
  uint64_t* w_full = acc_empty + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  float* sbias = reinterpret_cast<float*>(bars + 64);  // 512 bytes of barrier space, then the bias vector

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // Programmatic dependent launch (small decodes, ConvTcPlan::pdl): the next launch of the stream may begin its prologue
  // (barrier init, TMEM allocation, descriptor prefetch, bias staging -- nothing the previous kernel writes) on SMs this
  // grid leaves idle; every access to activations comes after pdl_wait() below.
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    for (int sg = 0; sg < p.g.nseg; ++sg) tma_prefetch_desc(&tm.a[sg]);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < kMaxNA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kMaxNB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 8; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kEpiWarps); }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < p.g.n_total; i += kTcThreads) sbias[i] = p.ep.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nkc = p.g.c_in / KC;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      // ring positions are (stage, phase) counters: a runtime `it % NA` is ~150 cycles of dependent integer code per
      // K-chunk / tap in a single-thread stream
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      if (p.stationary) {
        // all taps x K-chunks of this layer's weights stay resident: one bulk load per CTA, no per-tap handshake
        // non-zero K-chunks only, packed in (tap, chunk) order: slot = w_slot[tap] + rank of the chunk in the mask
        int nchunks = 0;
        for (int tap = 0; tap < p.g.ntaps; ++tap) nchunks += __popc(p.g.tap_kmask[tap] & ((1u << nkc) - 1u));
        mbar_expect_tx(w_full, nchunks * B_STAGE);
        int slot = 0;
        for (int tap = 0; tap < p.g.ntaps; ++tap)
          for (int kc = 0; kc < nkc; ++kc)
            if ((p.g.tap_kmask[tap] >> kc) & 1u)
              tma_load_3d(&tmW, w_full, smemB + (slot++) * B_STAGE, kc * KC, 0, tap);
      }
      pdl_wait();   // resident weights (static data) are already on their way; activations only from here on
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        uint32_t mb, nt, bq, mt, bu, rho;
        p.div_n.divmod(tile, mb, nt);
        p.div_m.divmod(mb, bq, mt);
        p.div_rho.divmod(bq, bu, rho);
        const int b = bu;
        const int t0 = mt * BMr;
        const int n0 = nt * BN;
        if (p.res_prefetch) {
          // the residual tiles this tile's epilogue will read: start them towards L2 now (the producer runs NA
          // activation stages ahead of the MMAs, so this is early enough to hide the DRAM latency)
          for (int i = 0; i < p.ep.nres; ++i)
            for (int r = 0; r < BMr; r += 64)
              for (int c = 0; c < BN; c += C::RBOXC) tma_prefetch_3d(&tm.r[i], n0 + c, t0 + r, b);
        }
        int tap0 = 0;
        for (int sg = 0; sg < p.g.nseg; ++sg) {
          const int tap1 = p.g.seg_tap_end[sg];
          const int nbx = p.seg_nboxes[sg];
          for (int kc = 0; kc < nkc; ++kc) {
            mbar_wait(&a_empty[sa], pa ^ 1);
            if (kTrace && p.trace && blockIdx.x == 0 && sg == 0 && kc == 0 && tile / gridDim.x < 256)
              p.trace[(tile / gridDim.x) * 12 + 0] = clock64();
            mbar_expect_tx(&a_full[sa], nbx * 64 * ROWB);
            // activation map dims: [KC][K-chunk][1][row][utterance], or for a dilated folded view
            // [C][sub-sequence rho][phase = K-chunk][row][utterance] (conv_params.h)
            const int c1 = p.rho_d > 1 ? (int)rho : kc, c2 = p.rho_d > 1 ? kc : 0;
            for (int bx = 0; bx < nbx; ++bx)
              tma_load_5d(&tm.a[sg], &a_full[sa], smemA + sa * p.a_stage_bytes + bx * 64 * ROWB, 0, c1, c2,
                          t0 + p.seg_halo_lo[sg] + bx * 64, b);
            if (++sa == (uint32_t)NA) { sa = 0; pa ^= 1; }
            if (p.stationary) continue;
            for (int tap = tap0; tap < tap1; ++tap) {
              if (p.g.tap_nlo[tap] >= n0 + BN || p.g.tap_nhi[tap] <= n0 || !((p.g.tap_kmask[tap] >> kc) & 1u)) continue;
              mbar_wait(&b_empty[sb], pb ^ 1);
              mbar_expect_tx(&b_full[sb], B_STAGE);
              tma_load_3d(&tmW, &b_full[sb], smemB + sb * B_STAGE, kc * KC, n0, tap);
              if (++sb == (uint32_t)NB) { sb = 0; pb ^= 1; }
            }
          }
          tap0 = tap1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // The whole warp runs this (warp-uniform) control flow; only the elected lane issues tcgen05.mma / commit.
    // Descriptors are (lo, hi) pairs: hi is a constant, lo advances by precomputed 16-byte-unit deltas, so the
    // issue loop is a handful of uniform-register adds per MMA.
    // SWAP: D[channel, time] = W[channel, ci] * X[time, ci]^T -- the 128 output channels are the MMA's M, the 256 time
    // rows its N, so one instruction reads 4 KB of weights + 8 KB of activations per 128 cycles (96 B/cycle of
    // shared-memory operand traffic) instead of 2 x (4 KB + 4 KB) per 2 x 64 cycles (128 B/cycle, the smem limit).
    const uint32_t idesc = umma_idesc_f16(SWAP ? p.swap_rows : BN, F16);
    constexpr uint32_t desc_hi = umma_desc_hi(ROWB);
    const uint32_t leader = elect_one();
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(smemA)), b_lo0 = umma_desc_lo(smem_u32(smemB));
    const uint32_t a_stage16 = (uint32_t)p.a_stage_bytes >> 4;
    uint32_t itt = 0;
    uint32_t sa = 0, pa = 0, sb_next = 0, pb = 0;   // ring (stage, phase) counters, see the producer
    pdl_wait();   // (this warp writes one row of a staged activation tile in the dilated folded view)
    if (p.stationary) {
      mbar_wait(w_full, 0);
      tc_fence_after();
    }
    int tapmask_n0 = -1;     // the taps that reach an N-tile depend on the tile's column range only: recompute on change
    uint32_t tapmask = 0;    // (2 constant-bank loads per tap; one N-tile per layer is the common case)
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++itt) {
      const uint32_t mb_ = p.div_n.quot(tile);
      const int n0 = (tile - mb_ * p.n_tiles) * BN;
      uint32_t bq_, mt_, bu_, rho_;
      p.div_m.divmod(mb_, bq_, mt_);
      p.div_rho.divmod(bq_, bu_, rho_);
      if (n0 != tapmask_n0) {
        tapmask = 0;
        for (int tap = 0; tap < p.g.ntaps; ++tap)
          if (p.g.tap_nlo[tap] < n0 + BN && p.g.tap_nhi[tap] > n0) tapmask |= 1u << tap;
        tapmask_n0 = n0;
      }
      const uint32_t as = itt % C::NBUF, pacc = (itt / C::NBUF) & 1;
      mbar_wait(&acc_empty[as], pacc ^ 1);
      tc_fence_after();
      const bool tr = kTrace && p.trace && blockIdx.x == 0 && itt < 256 && lane == 0;
      if (tr) p.trace[itt * 12 + 1] = clock64();
      const uint32_t d_base = tmem_base + as * ACC_COLS;
      uint32_t accum = 0;  // 0 for the first MMA of each accumulator of this tile
      int tap0 = 0;
      for (int sg = 0; sg < p.g.nseg; ++sg) {
      const int tap1 = p.g.seg_tap_end[sg];
      for (int kc = 0; kc < nkc; ++kc) {
        mbar_wait(&a_full[sa], pa);
        tc_fence_after();
        if (tr && sg == 0 && kc == 0) p.trace[itt * 12 + 2] = clock64();
        if (p.rho_d > 1) {
          // Dilated folded view: the last row of a sub-sequence may hold samples t >= L_real, whose addresses alias the
          // next utterance (the tensor map bounds rows, not samples).  The convolution needs zeros there: overwrite
          // that one row of the staged tile before the MMAs read it.
          const int rem = p.L_real - (int)rho_ - p.rho_d * kc;  // K-chunk == phase of the row's samples
          const int nlim = rem > 0 ? (int)p.div_dr.quot(rem + p.rho_d * p.r_fold - 1) : 0;
          const int idx = p.g.L - 1 - ((int)mt_ * BMr + p.seg_halo_lo[sg]);
          if (nlim < p.g.L && idx >= 0 && idx < p.seg_nboxes[sg] * 64) {
            if (lane < ROWB / 16)
              *reinterpret_cast<uint4*>(smemA + sa * p.a_stage_bytes + idx * ROWB + lane * 16) = make_uint4(0, 0, 0, 0);
            fence_proxy_async();
            __syncwarp();
          }
        }
        const uint32_t a_lo_stage = a_lo0 + sa * a_stage16;
        for (int tap = tap0; tap < tap1; ++tap) {
          if (!((tapmask >> tap) & 1u) || !((p.g.tap_kmask[tap] >> kc) & 1u)) continue;
          uint32_t sb = 0;
          uint32_t b_lo;
          if (p.stationary) {
            const uint32_t slot = p.w_slot[tap] + __popc(p.g.tap_kmask[tap] & ((1u << kc) - 1u));
            b_lo = b_lo0 + slot * (B_STAGE >> 4);
          } else {
            sb = sb_next;
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            b_lo = b_lo0 + sb * (B_STAGE >> 4);
          }
          const uint32_t a_lo = a_lo_stage + p.tap_delta16[tap];
          if constexpr (SWAP) {
#pragma unroll
            for (int k = 0; k < KC / 16; ++k)
              umma_f16_lohi(d_base, b_lo + ((k * 32) >> 4), desc_hi, a_lo + ((k * 32) >> 4), desc_hi, idesc,
                            k == 0 ? accum : 1u, leader);
          } else {
#pragma unroll
            for (int acc = 0; acc < NACC; ++acc) {
#pragma unroll
              for (int k = 0; k < KC / 16; ++k) {
                umma_f16_lohi(d_base + acc * BN, a_lo + ((acc * 128 * ROWB + k * 32) >> 4), desc_hi,
                              b_lo + ((k * 32) >> 4), desc_hi, idesc, k == 0 ? accum : 1u, leader);
              }
            }
          }
          accum = 1;
          if (!p.stationary) {
            if (leader) umma_commit(&b_empty[sb]);  // weights stage free once these MMAs retire
            if (++sb_next == (uint32_t)NB) { sb_next = 0; pb ^= 1; }
          }
        }
        if (leader) umma_commit(&a_empty[sa]);
        if (++sa == (uint32_t)NA) { sa = 0; pa ^= 1; }
      }
      tap0 = tap1;
      }
      if (leader) umma_commit(&acc_full[as]);
      if (tr) p.trace[itt * 12 + 3] = clock64();
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue warps
    // TMEM lane quadrant = warp % 4 (hardware rule); the four warps of a quadrant take a tile's items round-robin.
    static_assert(!SWAP || (BN == 128 && NACC == 2), "SWAP tiles are 128 channels x 256 time rows");
    const FastDiv div_rho = p.div_rho, div_dr = p.div_dr;
    const int rho_d = p.rho_d, c_shift = p.c_shift, r_fold = p.r_fold, L_real = p.L_real, rowstride = p.rowstride;
    const long bstride = p.bstride;
    constexpr int CHUNKS = BN / kIW, NW = kEpiWarps / 4;
    static_assert(NACC * CHUNKS >= NW, "every epilogue warp needs at least one item per tile");
    const int NITEMS = SWAP ? p.swap_rows / kIW : NACC * CHUNKS;   // swap_rows >= 64: at least one item per warp
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    // two 1 KB transposition buffers per warp: with TMA output stores an item's buffer is still being read by the TMA
    // engine while the next item is staged in the other one
    uint8_t* scratch = reinterpret_cast<uint8_t*>(sbias) + 8192 + (warp - 2) * 2048;
    constexpr bool kTmaEpi = SWAP && EPI >= 1 && EPI <= 3;
    const bool tma_out = kTmaEpi && p.tma_epi != 0;   // launch constant: plain [B][L][n_total] output view
    uint32_t nitem = 0;
    // launch constants live in registers for the whole loop (each read of the __grid_constant__ block is an LDC)
    const FastDiv div_n = p.div_n, div_m = p.div_m;
    const int L = p.g.L, n_total = p.g.n_total, total_tiles = p.total_tiles, tile_step = gridDim.x;
    const float out_slope = p.ep.out_slope, mrf_scale = p.ep.mrf_scale, res_gain = p.ep.res_gain;
    ConvEpilogue ep = p.ep;
    pdl_wait();   // residual reads and output stores depend on the previous launch
    auto coords = [&](int tile, int it, EpiItem& e) {
      uint32_t mb, nt, bq, mt;
      div_n.divmod(tile, mb, nt);
      div_m.divmod(mb, bq, mt);
      e.b = bq;
      if constexpr (SWAP) {
        const int t = mt * BMr + it * kIW;      // item = 16 time rows (TMEM columns) x this warp's 32 channels
        e.n = nt * BN + q * 32;
        // utterance x sub-sequence -> (utterance, rho); the warp's 32 columns are channels [c, c+32) of phase phi
        uint32_t bu, rho;
        div_rho.divmod(bq, bu, rho);
        e.b = bu;
        const int phi = e.n >> c_shift, c = e.n & ((1 << c_shift) - 1);
        int nlim = L;
        if (rho_d > 1) {
          const int rem = L_real - (int)rho - rho_d * phi;
          nlim = rem > 0 ? (int)div_dr.quot(rem + rho_d * r_fold - 1) : 0;
        }
        e.rows_valid = min(kIW, max(0, nlim - t));
        e.row0 = (long)e.b * L + t;
        e.base = (long)e.b * bstride + (((int)rho + rho_d * phi) << c_shift) + c + (long)t * rowstride;
        e.tcol = it * kIW;
        e.t = t;
      } else {
        const int acc = it / CHUNKS, c0 = (it % CHUNKS) * kIW;
        const int t = mt * BM + acc * 128 + q * 32;
        e.n = nt * BN + c0;
        e.rows_valid = min(32, max(0, L - t));
        e.row0 = (long)e.b * L + t;
        e.tcol = acc * BN + c0;
      }
    };
    auto issue = [&](const EpiItem& e, EpiLoads& l) {
      if constexpr (SWAP) epiT_issue_loads<EPI>(ep, e, rowstride, lane, l);
      else epi_issue_loads<EPI>(ep, e, n_total, lane, l);
    };
    int tile = blockIdx.x, it = hsel;
    uint32_t itt = 0;
    EpiLoads ld;
    EpiItem cur{};
    if (tile < total_tiles) {
      coords(tile, it, cur);
      issue(cur, ld);
    }
    while (tile < total_tiles) {
      const bool first = it < NW, last = it + NW >= NITEMS;
      const uint32_t as = itt % C::NBUF, pacc = (itt / C::NBUF) & 1;
      if (first) {
        mbar_wait(&acc_full[as], pacc);
        tc_fence_after();
      }
      const bool tr = kTrace && p.trace && blockIdx.x == 0 && warp == 2 && lane == 0 && first && itt < 256;
      if (tr) p.trace[itt * 12 + 4] = clock64();
      uint32_t acc[kIW];
      float v[kIW];
      __syncwarp();
      if constexpr (SWAP) tmem_ld_frag(tmem_base + ((uint32_t)(q * 32) << 16) + as * ACC_COLS + cur.tcol, acc);
      else tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + as * ACC_COLS + cur.tcol, acc);
      float4 bv[kIW / 4];  // bias for this item's columns: read from smem while the TMEM load is in flight
      float bias4[4] = {0.f, 0.f, 0.f, 0.f};  // channels-as-M: this thread's channels are 8m + lane/4 (fragment layout)
      if constexpr (SWAP) {
        // (keeping these four values in registers across the items of an N-tile measured 0.8 % SLOWER, tools/ab.sh)
#pragma unroll
        for (int m = 0; m < 4; ++m) bias4[m] = sbias[cur.n + 8 * m + (lane >> 2)];
      } else {
#pragma unroll
        for (int j = 0; j < kIW / 4; ++j) bv[j] = *reinterpret_cast<const float4*>(sbias + cur.n + 4 * j);
      }
      tmem_ld_wait();
      if (tr) p.trace[itt * 12 + 5] = clock64();
      if constexpr (EPI == 4) {
        // conv_post: this lane's folded column is (phase, channel) = ((32q + lane) / C, (32q + lane) % C); only
        // channel 0 is a real output.  Columns of the item are consecutive folded rows: samples r apart.
        const int pc = ep.post_c;
        if ((q * 32) % pc == 0) {  // warp-uniform: this warp's first channel is channel 0 of phase 32q / C
          // fragment layout: lanes 0..3 hold channel 0 (bf16(w) sums), lanes 4..7 channel 1 (the weights' bf16
          // remainders), four columns each: add the two rows, then tanh
          const int r = n_total / pc;
          float* o = ep.out_f32 + cur.row0 * r + (q * 32) / pc;
#pragma unroll
          for (int cg = 0; cg < 2; ++cg)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float hi = __uint_as_float(acc[frag_idx(0, cg, e)]);
              const float x = hi + __shfl_down_sync(0xffffffffu, hi, 4);
              const int col = cg * 8 + 2 * (lane & 3) + e;
              if (lane < 4 && col < cur.rows_valid) o[(long)col * r] = tanhf(x);
            }
        }
      } else if constexpr (SWAP) {
        uint8_t* buf = scratch;
        if constexpr (kTmaEpi) {
          if (tma_out) {   // the store that last read this buffer (two items ago) must have drained it
            buf = scratch + (nitem & 1) * 1024;
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
          }
        }
        epiT_accumulate<EPI, F16>(ep, bias4, buf, cur, n_total, lane, res_gain, acc, ld, v);
      }
      else epi_accumulate<EPI, F16>(ep, bv, scratch, cur, n_total, lane, res_gain, acc, ld, v);
      if (tr) p.trace[itt * 12 + 6] = clock64();
      if (last) {  // accumulator fully read: hand the TMEM buffer back before the stores
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);
        ++itt;
      }
      if (tr) p.trace[(itt - (last ? 1 : 0)) * 12 + 8] = clock64();
      const EpiItem done = cur;
      // next item: its global reads go out now and land while this item is stored and the next accumulator is awaited
      int ntile = tile, nit = it + NW;
      if (nit >= NITEMS) { nit = hsel; ntile += tile_step; }
      if (ntile < total_tiles) {
        coords(ntile, nit, cur);
        issue(cur, ld);
      }
      if (tr) p.trace[(itt - (last ? 1 : 0)) * 12 + 9] = clock64();
      if constexpr (EPI == 4) { (void)done; }
      else if constexpr (SWAP) {
        if constexpr (kTmaEpi) {
          if (tma_out) epiT_store_tma<EPI, F16>(&tm.o, scratch + (nitem & 1) * 1024, done, lane, out_slope, mrf_scale, v);
          else epiT_store<EPI, F16>(ep, scratch, done, rowstride, lane, out_slope, mrf_scale, v);
          ++nitem;
        } else {
          epiT_store<EPI, F16>(ep, scratch, done, rowstride, lane, out_slope, mrf_scale, v);
        }
      }
      else epi_store<EPI, F16>(ep, scratch, done, n_total, lane, out_slope, mrf_scale, v);
      if (tr) p.trace[(itt - (last ? 1 : 0)) * 12 + 7] = clock64();
      tile = ntile; it = nit;
    }
    if constexpr (kTmaEpi) {
      if (tma_out && lane == 0) bulk_wait<0>();   // every output store of this warp has completed before the CTA retires
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// bf16 tensor [d2][d1][d0] (d0 contiguous), box {b0, b1, 1}, swizzle span = b0*2 bytes
static int encode_3d(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0,
                     uint32_t b1, bool swizzle = true) {
  EncodeTiledFn fn = get_encode_fn();
  VD_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = !swizzle ? CU_TENSOR_MAP_SWIZZLE_NONE
                          : b0 * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                          : (b0 * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) dims=(%llu,%llu,%llu) box=(%u,%u)", (int)r,
             (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, b0, b1);
    set_error(buf);
    return 1;
  }
  return 0;
}

// Activation tensor as the kernel's 5-d view (bf16): dims {kc, d1, d2, rows, B} with element strides
// {1, s1, s2, srow, sb}; box {kc, 1, 1, 64, 1}, swizzle span = kc*2 bytes.
static int encode_act(CUtensorMap* m, const void* base, uint32_t kc, uint64_t d1, uint64_t s1, uint64_t d2, uint64_t s2,
                      uint64_t rows, uint64_t srow, uint64_t B, uint64_t sb, uint32_t box_rows = 64,
                      uint32_t box_d2 = 1) {
  EncodeTiledFn fn = get_encode_fn();
  VD_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[5] = {kc, d1, d2, rows, B};
  cuuint64_t strides[4] = {s1 * 2, s2 * 2, srow * 2, sb * 2};
  cuuint32_t box[5] = {kc, 1, box_d2, box_rows, 1};  // box_d2 > 1: several phases side by side in one shared-memory row
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const uint32_t row_bytes = kc * 2 * box_d2;
  CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                           : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[320];
    snprintf(buf, sizeof buf,
             "cuTensorMapEncodeTiled(5d) failed (%d) dims=(%u,%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu,%llu)", (int)r,
             kc, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)rows, (unsigned long long)B,
             (unsigned long long)s1, (unsigned long long)s2, (unsigned long long)srow, (unsigned long long)sb);
    set_error(buf);
    return 1;
  }
  return 0;
}

int encode_tmap_act(CUtensorMap* m, const void* base, uint32_t kc, uint64_t d1, uint64_t s1, uint64_t d2, uint64_t s2,
                    uint64_t rows, uint64_t srow, uint64_t B, uint64_t sb, uint32_t box_rows, uint32_t box_d2) {
  return encode_act(m, base, kc, d1, s1, d2, s2, rows, srow, B, sb, box_rows, box_d2);
}

int encode_tmap_3d(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                   bool swizzle) {
  return encode_3d(m, base, d0, d1, d2, b0, b1, swizzle);
}

template <int BN, int KC, int EPI, bool SWAP, bool F16>
static int launch_typed(const ConvTcPlan& pl, cudaStream_t stream) {
  static bool attr_set = false;  // benign race: idempotent
  if (!attr_set) {
    VD_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, KC, EPI, SWAP, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024));
    attr_set = true;
  }
  // More than half an SM's shared memory for EVERY launch of this kernel: CTAs of two overlapping launches (programmatic
  // dependent launch, concurrent MRF branches) can then never share an SM, so a dependent's CTA can never take the TMEM
  // columns that a CTA of the launch it waits for is about to allocate (deadlock).  Today the register file already
  // forbids such co-residency; this makes it independent of the compiler's register count.
  const size_t smem = pl.smem > (size_t)120 * 1024 ? pl.smem : (size_t)120 * 1024;
  if (pl.pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pl.grid);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    VD_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, KC, EPI, SWAP, F16>, pl.tm, pl.tmW, pl.p));
    return 0;
  }
  conv_tc_kernel<BN, KC, EPI, SWAP, F16><<<pl.grid, kTcThreads, smem, stream>>>(pl.tm, pl.tmW, pl.p);
  VD_CUDA(cudaGetLastError());
  return 0;
}

// ConvEpilogue::f16 selects the fp16-storage instance of the same kernel (operand format bits of the instruction
// descriptor and the epilogue's 16-bit conversions are compile-time)
template <int BN, int KC, int EPI, bool SWAP = false>
static int launch_one(const ConvTcPlan& pl, cudaStream_t stream) {
  return pl.p.ep.f16 ? launch_typed<BN, KC, EPI, SWAP, true>(pl, stream)
                     : launch_typed<BN, KC, EPI, SWAP, false>(pl, stream);
}

// channels-as-M variant (128 channels x 256 time rows per tile); never with the fp32 MRF fallback
static int launch_swapped(const ConvTcPlan& pl, cudaStream_t stream) {
  const ConvEpilogue& e = pl.p.ep;
  const bool simple = e.bias_b == nullptr;
  if (pl.kc == 32) {  // dilated folded view of a 32-channel tensor: K-chunk = one time phase
    if (simple && e.mrf_mode == 0 && e.nres == 0) return launch_one<128, 32, 1, true>(pl, stream);
    if (simple && e.mrf_mode == 0 && e.nres == 1) return launch_one<128, 32, 2, true>(pl, stream);
    VD_CHECK(e.mrf_mode == 0 || e.mrf_mode == 3, "conv_tc: epilogue not available for 32-channel K-chunks");
    return launch_one<128, 32, 0, true>(pl, stream);
  }
  if (e.mrf_mode == 4) return launch_one<128, 64, 4, true>(pl, stream);
  if (simple && e.mrf_mode == 0 && e.nres == 0) return launch_one<128, 64, 1, true>(pl, stream);
  if (simple && e.mrf_mode == 0 && e.nres == 1) return launch_one<128, 64, 2, true>(pl, stream);
  if (simple && e.mrf_mode == 3 && e.nres == 3) return launch_one<128, 64, 3, true>(pl, stream);
  return launch_one<128, 64, 0, true>(pl, stream);
}

// Specialised epilogues exist for the shapes of the shipped configuration; everything else takes the generic one.
template <int BN, int KC, bool SPECIALISED>
static int launch_inst(const ConvTcPlan& pl, cudaStream_t stream) {
  const ConvEpilogue& e = pl.p.ep;
  if constexpr (SPECIALISED) {
    const bool simple = e.bias_b == nullptr && e.rowmask == nullptr && !e.gate && !e.split_col &&
                        (e.mrf_mode == 0 || (e.mrf_mode == 3 && e.mrf == nullptr));
    if (simple && e.mrf_mode == 0 && e.nres == 0) return launch_one<BN, KC, 1>(pl, stream);
    if (simple && e.mrf_mode == 0 && e.nres == 1) return launch_one<BN, KC, 2>(pl, stream);
    if (simple && e.mrf_mode == 3 && e.nres == 3) return launch_one<BN, KC, 3>(pl, stream);
  }
  return launch_one<BN, KC, 0>(pl, stream);
}

int plan_conv_tc(ConvTcPlan* pl, const ConvGeom& g, const __nv_bfloat16* const* xs, const __nv_bfloat16* w,
                 int num_sms, int desc_mode, bool allow_swap) {
  const int raw_desc_mode = desc_mode;
  const bool force_streaming = (desc_mode & 2) != 0;  // test knob: exercise the streamed-weights path everywhere
  // The producer's TMA L2 prefetch of the residual tiles is OFF since round 2 (desc_mode bit 2 turns it on): with the
  // ResBlock pairs fused, what is left of the residual reads are tiles the same SM loaded as operands a moment ago, and
  // the step measured 1.0 % faster without the prefetches (8.63 vs 8.72 ms, three alternating pairs of runs on one box).
  pl->no_res_prefetch = (desc_mode & 4) == 0;
  const bool force_no_swap = (desc_mode & 8) != 0;   // test knob: keep wide layers on the time-as-M form
  // experiment knob: channels-as-M epilogue with register transposes (movmatrix) and 4-byte global accesses instead
  // of the shared-memory transposition.  It removes ~15 % of the SM's shared-memory traffic but its partial-sector
  // loads/stores cost far more: 11.97 vs 10.52 ms per 16 x 10 s step.
  pl->epi_smem = (desc_mode & 128) == 0;
  pl->pdl = false;
  pl->no_tma_epi = (desc_mode & 2048) != 0;   // experiment knob: LSU output stores everywhere (the round-1 epilogue)
  pl->out_bound = nullptr;
  const int na_stream = (desc_mode & 32) ? 3 : ((desc_mode & 64) ? 4 : 2);  // experiment knob: activation stages when streaming
  desc_mode &= 1;
  VD_CHECK(g.c_in % 32 == 0, "conv_tc: c_in must be a multiple of 32");
  VD_CHECK(g.n_total % 32 == 0, "conv_tc: output columns must be a multiple of 32");
  VD_CHECK(g.ntaps >= 1 && g.ntaps <= kMaxTaps, "conv_tc: 1..32 taps supported");
  VD_CHECK(g.nseg >= 1 && g.nseg <= kMaxSeg && g.seg_tap_end[g.nseg - 1] == g.ntaps, "conv_tc: bad segment table");
  VD_CHECK(g.n_total <= 2048, "conv_tc: at most 2048 output columns per row");
  const int rho_d = std::max(1, g.rho_d);
  int kc = (g.c_in % 64 == 0) ? 64 : 32;
  if (rho_d > 1) {
    VD_CHECK((g.c_real == 32 || g.c_real == 64) && g.c_in % g.c_real == 0 && g.n_total == 128 && allow_swap,
             "conv_tc: a dilated folded view needs 32 or 64 real channels and 128 folded columns");
    kc = g.c_real;  // one time phase per K-chunk: phases of a dilated view are not contiguous in memory
  }
  VD_CHECK(g.c_in / kc <= 31, "conv_tc: too many K-chunks per tap");
  int bn = 32;
  for (int c : {256, 128, 64}) {
    if (g.n_total % c == 0) { bn = c; break; }
  }
  // Small problems (the flow's frame-rate convs): with 256 x 128 tiles a launch may have barely more tiles than SMs (162
  // on 148 = two rounds, the second almost empty); halving the tile width evens the rounds out.
  if (!allow_swap && bn == 128) {
    const long tiles128 = (long)g.B * ((g.L + 255) / 256) * (g.n_total / 128);
    const long r128 = (tiles128 + num_sms - 1) / num_sms, r64 = (2 * tiles128 + num_sms - 1) / num_sms;
    if (tiles128 <= 4L * num_sms && r64 < 2 * r128 && !(raw_desc_mode & 256)) bn = 64;
  }
  // Wide layers whose weights do not stay resident run channels-as-M (128 channels x 256 time rows): at N=128 the
  // time-as-M form is bound by shared-memory operand bandwidth, at N=256 by L2 weight streaming.
  const bool want_swap = allow_swap && (kc == 64 || rho_d > 1) && g.n_total % 128 == 0 && !force_no_swap;
  if (want_swap) bn = 128;
  pl->swap = want_swap;
  VD_CHECK(rho_d == 1 || want_swap, "conv_tc: a dilated folded view runs channels-as-M only");
  const int nacc = bn >= 256 ? 1 : 2;
  int bm = 128 * nacc;
  if (want_swap) {
    // Small decodes: a 2 s utterance has 6-44 tiles of 256 rows per launch in stages 0-1 for 148 SMs (x3 with the
    // MRF branches running concurrently).  Narrower tiles (N = 128 / 64 per MMA) are less efficient per MAC but put
    // the launch on more SMs; large batches keep 256 rows.
    while (bm > 64 && 3L * g.B * rho_d * ((g.L + bm - 1) / bm) * (g.n_total / bn) <= num_sms) bm /= 2;
    if (raw_desc_mode & 512) bm = 256;   // experiment knob: always 256-row tiles
  }
  // Paired tiles (conv_tc2.cu, tcgen05.mma.cta_group::2): layers with exactly 256 output channels whose taps all span
  // every column and every K-chunk (the ResBlock convs and the fused MRF launch of a 256-channel stage), full-size tiles.
  bool cta2 = want_swap && rho_d == 1 && kc == 64 && g.n_total == 256 && bm == 256 && !(raw_desc_mode & 4096);
  for (int t = 0; t < g.ntaps && cta2; ++t)
    cta2 = g.tap_nlo[t] <= 0 && g.tap_nhi[t] >= g.n_total && (g.tap_kmask[t] & ((1u << (g.c_in / kc)) - 1u)) == ((1u << (g.c_in / kc)) - 1u);
  const int clusters2 = cta2 ? max_clusters_tc2((size_t)227 * 1024) : 0;
  if (clusters2 <= 0) cta2 = false;
  pl->cta2 = cta2;
  const int rows_cta = cta2 ? 128 : bm;   // time rows of the tile staged by ONE CTA
  ConvTcParams& p = pl->p;
  p.g = g;
  p.swap_rows = bm;
  p.a_stage_bytes = 0;
  int tap0 = 0;
  for (int sg = 0; sg < g.nseg; ++sg) {
    const int tap1 = g.seg_tap_end[sg];
    VD_CHECK(tap1 > tap0, "conv_tc: empty segment");
    int lo = g.tap_off[tap0], hi = g.tap_off[tap0];
    for (int i = tap0 + 1; i < tap1; ++i) { lo = std::min(lo, g.tap_off[i]); hi = std::max(hi, g.tap_off[i]); }
    p.seg_halo_lo[sg] = lo;
    p.seg_nboxes[sg] = (rows_cta + (hi - lo) + 63) / 64;
    p.a_stage_bytes = std::max(p.a_stage_bytes, p.seg_nboxes[sg] * 64 * kc * 2);
    for (int i = tap0; i < tap1; ++i) p.tap_delta16[i] = (uint32_t)((g.tap_off[i] - lo) * kc * 2) >> 4;
    tap0 = tap1;
  }
  p.m_tiles = (g.L + bm - 1) / bm;
  p.n_tiles = g.n_total / bn;
  p.total_tiles = cta2 ? g.B * p.m_tiles : g.B * rho_d * p.m_tiles * p.n_tiles;   // cta2: one tile = both channel halves
  p.div_n.init(p.n_tiles);
  p.div_m.init(p.m_tiles);
  p.div_rho.init(rho_d);
  p.rho_d = rho_d;
  if (rho_d > 1) {
    p.r_fold = g.c_in / g.c_real;
    p.c_shift = g.c_real == 32 ? 5 : 6;
    p.L_real = g.L_real;
    p.bstride = (long)g.L_real * g.c_real;
    p.rowstride = rho_d * p.r_fold * g.c_real;
    VD_CHECK(g.L == (g.L_real + rho_d * p.r_fold - 1) / (rho_d * p.r_fold), "conv_tc: folded row count mismatch");
  } else {
    p.r_fold = 1;
    p.c_shift = 5;
    p.L_real = g.L;
    p.bstride = (long)g.L * g.n_total;
    p.rowstride = g.n_total;
  }
  p.div_dr.init(rho_d * p.r_fold);
  p.desc_mode = desc_mode;
  p.res_prefetch = 0;
  p.tma_epi = 0;
  p.cta2_relay = (raw_desc_mode & 8192) ? 1 : 0;
  p.trace = nullptr;
  pl->bn = bn;
  pl->kc = kc;
  pl->grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  if (cta2) pl->grid = 2 * std::min(p.total_tiles, std::min(clusters2, num_sms / 2));
  const int b_stage = bn * kc * 2;
  int w_chunks = 0;
  for (int t = 0; t < g.ntaps; ++t) {
    p.w_slot[t] = (uint16_t)w_chunks;
    w_chunks += __builtin_popcount(g.tap_kmask[t] & ((1u << (g.c_in / kc)) - 1u));
  }
  const int w_all = w_chunks * b_stage;  // non-zero K-chunks only
  // weights stay resident in shared memory when the whole layer fits next to >= 2 activation stages
  p.stationary = (p.n_tiles == 1 && w_all + 2 * p.a_stage_bytes <= kSmemBudget) ? 1 : 0;
  if (force_streaming || cta2) p.stationary = 0;
  if (p.stationary) {
    p.b_region_bytes = w_all;
    p.nb_stages = 0;
    p.na_stages = std::min(kMaxNA, (kSmemBudget - w_all) / p.a_stage_bytes);
    p.na_stages = std::min(p.na_stages, std::max(2, (96 * 1024) / p.a_stage_bytes));  // deeper did not help (r1f)
  } else {
    p.na_stages = 2;
    VD_CHECK(2 * p.a_stage_bytes + 2 * b_stage <= kSmemBudget, "conv_tc: dilation halo too large for shared memory");
    if (na_stream * p.a_stage_bytes + 3 * b_stage <= kSmemBudget) p.na_stages = na_stream;
    if (cta2 && 3 * p.a_stage_bytes + 6 * b_stage <= kSmemBudget) p.na_stages = 3;   // half-size stages: a third one is cheap
    p.nb_stages = std::min(kMaxNB, (kSmemBudget - p.na_stages * p.a_stage_bytes) / b_stage);
    p.b_region_bytes = p.nb_stages * b_stage;
  }
  pl->smem = 1024 + (size_t)p.na_stages * p.a_stage_bytes + p.b_region_bytes + 512 + 8192 + 32768;
  for (int sg = 0; sg < kMaxSeg; ++sg) {
    if (sg >= g.nseg) {   // unused segments: valid placeholders without another driver call (an encode costs ~5 us)
      pl->tm.a[sg] = pl->tm.a[0];
      pl->tm.r[sg] = pl->tm.r[0];
      pl->res_bound[sg] = nullptr;
      continue;
    }
    const void* x = xs[sg];
    if (rho_d > 1) {
      // [C][rho][phase][row][utterance]: sample t = rho_d*(r*row + phase) + rho
      if (encode_act(&pl->tm.a[sg], x, kc, rho_d, g.c_real, p.r_fold, (uint64_t)rho_d * g.c_real, g.L,
                     (uint64_t)p.rowstride, g.B, (uint64_t)p.bstride))
        return 1;
    } else {
      if (encode_act(&pl->tm.a[sg], x, kc, g.c_in / kc, kc, 1, g.c_in, g.L, g.c_in, g.B, (uint64_t)g.L * g.c_in))
        return 1;
    }
    if (sg == 0 && encode_3d(&pl->tm.r[0], x, g.c_in, g.L, g.B, kc, 64)) return 1;
    pl->tm.r[sg] = pl->tm.r[0];  // placeholder until a residual is bound
    pl->res_bound[sg] = nullptr;
  }
  if (encode_3d(&pl->tmW, w, g.c_in, g.n_total, g.ntaps, kc, bn)) return 1;
  return 0;
}

// true when the launch takes one of the channels-as-M epilogues that can write its output through the TMA engine:
// a plain [B][L][n_total] 16-bit output view, specialised epilogue (launch_swapped: EPI 1-3)
static bool tma_epilogue_ok(const ConvTcPlan& pl, const ConvEpilogue& ep) {
  const bool simple = ep.bias_b == nullptr && ep.mrf == nullptr && !ep.split_col && !ep.gate && ep.rowmask == nullptr;
  const bool epi123 = (ep.mrf_mode == 0 && ep.nres <= 1) || (ep.mrf_mode == 3 && ep.nres == 3);
  return pl.swap && pl.kc == 64 && pl.p.rho_d == 1 && simple && epi123 && ep.out != nullptr && !pl.no_tma_epi;
}

int bind_residual_tc(ConvTcPlan& pl, const ConvEpilogue& ep) {
  VD_CHECK(ep.nres >= 0 && ep.nres <= kMaxSeg - 1, "conv_tc: at most 3 residual tensors");
  if (tma_epilogue_ok(pl, ep) && pl.out_bound != ep.out) {
    // output items of the channels-as-M epilogue: {32 channels, 16 rows} boxes, SWIZZLE_64B (= the scratch layout)
    if (encode_3d(&pl.tm.o, ep.out, pl.p.g.n_total, pl.p.g.L, pl.p.g.B, 32, 16)) return 1;
    pl.out_bound = ep.out;
  }
  if (pl.p.rho_d > 1 || ep.split_col) return 0;  // strided / split residual rows: no L2 prefetch maps
  for (int i = 0; i < ep.nres; ++i) {
    if (pl.res_bound[i] != ep.res[i]) {
      if (encode_3d(&pl.tm.r[i], ep.res[i], pl.p.g.n_total, pl.p.g.L, pl.p.g.B, pl.bn < 64 ? pl.bn : 64, 64, false))
        return 1;
      pl.res_bound[i] = ep.res[i];
    }
  }
  return 0;
}

int launch_conv_tc(ConvTcPlan& pl, const ConvEpilogue& ep, cudaStream_t stream) {
  pl.p.ep = ep;
  pl.p.ep.epi_smem = pl.epi_smem ? 1 : 0;
  if (bind_residual_tc(pl, ep)) return 1;  // no-op when the plan was built with these residuals
  pl.p.res_prefetch = (ep.nres > 0 && !pl.no_res_prefetch && pl.p.rho_d == 1 && !ep.split_col) ? 1 : 0;
  pl.p.tma_epi = tma_epilogue_ok(pl, ep) ? 1 : 0;
  if (!pl.p.tma_epi) { pl.tm.o = pl.tm.a[0]; pl.out_bound = nullptr; }   // valid placeholder
  if (pl.cta2) {
    const bool epi123 = ep.bias_b == nullptr && ep.mrf == nullptr &&
                        ((ep.mrf_mode == 0 && ep.nres <= 1) || (ep.mrf_mode == 3 && ep.nres == 3));
    VD_CHECK(epi123 && ep.rowmask == nullptr && !ep.gate && !ep.split_col,
             "conv_tc: this epilogue needs a plan built with desc_mode bit 12 (no paired tiles)");
    return launch_conv_tc2(pl, stream);
  }
  if (pl.swap) {
    VD_CHECK(ep.rowmask == nullptr && !ep.gate && !ep.split_col,
             "conv_tc: row mask / gate / split epilogue need a time-as-M plan (allow_swap = false)");
    VD_CHECK(ep.mrf == nullptr, "conv_tc: the channels-as-M variant has no fp32 MRF accumulator path");
    return launch_swapped(pl, stream);
  }
  VD_CHECK(ep.mrf_mode != 4, "conv_tc: the conv_post epilogue needs a channels-as-M (folded) launch");
  switch (pl.bn * 100 + pl.kc) {
    case 25664: return launch_inst<256, 64, true>(pl, stream);
    case 12864: return launch_inst<128, 64, true>(pl, stream);
    case 6464:  return launch_inst<64, 64, true>(pl, stream);
    case 3264:  return launch_inst<32, 64, false>(pl, stream);
    case 25632: return launch_inst<256, 32, false>(pl, stream);
    case 12832: return launch_inst<128, 32, false>(pl, stream);
    case 6432:  return launch_inst<64, 32, false>(pl, stream);
    case 3232:  return launch_inst<32, 32, true>(pl, stream);
  }
  set_error("conv_tc: no kernel instance");
  return 1;
}

}  // namespace vd
