// Fused ResBlock1 pair on the sm_100a tensor cores:
//     y = lrelu( c2( lrelu( c1(a) + b1 ) ) + b2 + x(a) )          (modules.py:211-221, one loop iteration)
// with a = lrelu(x) the stored a-form input, c1 = Conv1d(C, C, k, dilation d), c2 = Conv1d(C, C, k, dilation 1).
//
// Why: for the narrow stages (C = 32, 64) one conv per launch is bound by HBM and by the epilogue, not by the tensor
// pipe: every conv reads and writes a 226 MB tensor (16 x 10 s) for a few hundred MACs per element.  Here the
// intermediate h never leaves the SM: c1's accumulator goes TMEM -> registers (bias, leaky-relu, zero outside the
// utterance) -> shared memory in the 128B/64B-swizzled K-major layout, where c2's tcgen05.mma reads it as its A
// operand; the residual x is recovered from the activation tile that is already resident for c1.  HBM traffic per
// pair drops from 5 tensor passes to 2, launches from 2 to 1.
//
// Tile: 256 rows of h per CTA tile (two 128-row accumulators), of which 256 - (k-1) output rows are valid (c2 needs a
// (k-1)/2 halo of h on each side), so tiles advance by 256 - (k-1) rows.  Both convs' weights stay resident.
// Pipeline per CTA (tile i): TMA A(i) -> c1(i) -> epi1(i) [h -> smem] -> c2(i) -> epi2(i) [-> global], with c1(i+1)
// issued before c2(i) so the tensor pipe works on the next tile while the epilogue warps build h.
#include <algorithm>

#include "common.cuh"
#include "conv_pair.h"
#include "ptx.cuh"

namespace vd {

#ifndef VITSDEC_TRACE
#define VITSDEC_TRACE 0
#endif
constexpr bool kPairTrace = VITSDEC_TRACE != 0;
constexpr int kPairEpiWarps = 16;
constexpr int kPairThreads = 64 + 32 * kPairEpiWarps + 32;  // producer, c1 issuer, 16 epilogue warps, c2 issuer
constexpr int kPairC2Warp = 2 + kPairEpiWarps;
constexpr int kPairMaxNA = 4;
constexpr int kPairHRows = 272;  // 256 + (k-1) rounded up, k <= 15

__device__ __forceinline__ uint4 ld_shared_u4(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }

template <int ROWB>
__device__ __forceinline__ uint32_t swz_row(int row) {  // chunk XOR term of the TMA/UMMA swizzle for a 1024B-aligned tile
  return ROWB == 128 ? (row & 7) : ((row >> 1) & 3);
}

template <int CH, int KC, int NACC, bool F16>
__global__ void __launch_bounds__(kPairThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ PairParams p) {
  constexpr int ROWB = KC * 2;
  constexpr int B_STAGE = CH * ROWB;      // one tap's weights [CH][KC]
  constexpr int ACC_COLS = NACC * CH;     // NACC 128-row accumulators per conv (2, or 1 when smem is tight)
  constexpr int TMEM_COLS = 4 * ACC_COLS; // acc1[2] + acc2[2]
  constexpr int CHUNKS = CH / 16, NITEMS = NACC * CHUNKS, NW = kPairEpiWarps / 8;  // warps per quadrant per group
  constexpr int HROWS = NACC == 2 ? kPairHRows : 144;
  static_assert(KC == CH, "pair kernel: one K chunk per tap");
  static_assert(TMEM_COLS <= 512 && NITEMS >= NW, "pair kernel: C must be 32 or 64");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NA = p.na_stages;
  uint8_t* smemA = smem;
  uint8_t* smemW = smemA + NA * p.a_stage_bytes;
  uint8_t* smemH = smemW + 2 * p.k * B_STAGE;
  const int NH = p.nh;                     // 1 or 2 buffers for the intermediate h
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemH + NH * HROWS * ROWB);
  uint64_t* a_full = bars;                 // [kPairMaxNA]
  uint64_t* a_empty = a_full + kPairMaxNA; // [kPairMaxNA]  1 (c1 retired) + 16 (epilogue warps read the residual)
  uint64_t* acc1_full = a_empty + kPairMaxNA;
  uint64_t* acc1_empty = acc1_full + 2;
  uint64_t* acc2_full = acc1_empty + 2;
  uint64_t* acc2_empty = acc2_full + 2;
  uint64_t* h_full = acc2_empty + 2;      // [2]
  uint64_t* h_empty = h_full + 2;         // [2]
  uint64_t* w_full = h_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  float* sbias = reinterpret_cast<float*>(bars + 32);                 // 256 B of barriers, then 2*CH floats
  uint8_t* scratch_base = reinterpret_cast<uint8_t*>(sbias) + 1024;   // 16 warps x 1 KB

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_launch_dependents();   // see conv_tc_kernel: the next launch may start its prologue; activations are touched after pdl_wait()
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    // epilogue warps work in two groups of 8: group 0 turns c1's accumulator into h, group 1 finishes c2's
    for (int i = 0; i < kPairMaxNA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1 + kPairEpiWarps / 2); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc1_full[i], 1); mbar_init(&acc1_empty[i], kPairEpiWarps / 2);
      mbar_init(&acc2_full[i], 1); mbar_init(&acc2_empty[i], kPairEpiWarps / 2);
      mbar_init(&h_full[i], kPairEpiWarps / 2);
      mbar_init(&h_empty[i], 1);
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 2 * CH; i += kPairThreads) sbias[i] = i < CH ? p.bias1[i] : p.bias2[i - CH];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int hk = p.hk;
  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int my_tiles = p.total_tiles > (int)blockIdx.x ? (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(w_full, 2 * p.k * B_STAGE);
      for (int tap = 0; tap < 2 * p.k; ++tap) tma_load_3d(&tmW, w_full, smemW + tap * B_STAGE, 0, 0, tap);
      pdl_wait();   // the weights (static) load while the previous launch drains; activations only from here on
      uint32_t sa = 0, pa = 0;   // stage / phase counters: a runtime i % NA is ~150 cycles of dependent integer code
      for (int i = 0; i < my_tiles; ++i) {
        const uint32_t tile = blockIdx.x + i * gridDim.x;
        uint32_t b, mt;
        p.div_m.divmod(tile, b, mt);
        const int t0 = mt * p.bmo;
        mbar_wait(&a_empty[sa], pa ^ 1);
        if (kPairTrace && p.trace && blockIdx.x == 0 && i < 256) p.trace[i * 12 + 10] = clock64();
        mbar_expect_tx(&a_full[sa], p.nboxes * 64 * ROWB);
        for (int bx = 0; bx < p.nboxes; ++bx)
          tma_load_3d(&tmA, &a_full[sa], smemA + sa * p.a_stage_bytes + bx * 64 * ROWB, 0,
                      t0 - hk - hk * p.dil + bx * 64, (int)b);
        if (++sa == (uint32_t)NA) { sa = 0; pa ^= 1; }
      }
    }
  } else if (warp == 1 || warp == kPairC2Warp) {
    // ------------------------------------------------------------ MMA issuers (warp-uniform; elected lane issues)
    // Two issuing warps: warp 1 runs c1 of every tile, the last warp c2.  With ONE issuer the fixed cost of its
    // in-order stream per tile (4 mbarrier waits, 4 commits, index arithmetic: ~2000 cycles, gaps of ~600 cycles
    // between every c1 and c2 in profiles/r01_trace_pair.txt) overlapped with nothing and bounded every pair; the
    // tensor pipe executes the two streams in issue order and the hand-offs (acc1 / h / acc2 / activation stages) are
    // the same mbarriers as before.  tcgen05.commit tracks the MMAs of the committing thread only.
    constexpr uint32_t idesc = umma_idesc_f16(CH, F16);
    constexpr uint32_t desc_hi = umma_desc_hi(ROWB);
    const uint32_t leader = elect_one();
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(smemA)), w_lo0 = umma_desc_lo(smem_u32(smemW));
    const uint32_t h_lo0 = umma_desc_lo(smem_u32(smemH));
    const uint32_t a_stage16 = (uint32_t)p.a_stage_bytes >> 4;
    const uint32_t h_buf16 = (uint32_t)(HROWS * ROWB) >> 4;
    const int k = p.k;
    mbar_wait(w_full, 0);
    tc_fence_after();
    if (warp == 1) {
      const uint32_t tap_step16 = (uint32_t)(p.dil * ROWB) >> 4;
      uint32_t sa = 0, pa = 0;   // stage / phase counters instead of i % NA, i / NA
      for (int i = 0; i < my_tiles; ++i) {
        const uint32_t as = i & 1;
        if (kPairTrace && p.trace && blockIdx.x == 0 && i < 256 && lane == 0) p.trace[i * 12 + 11] = clock64();
        mbar_wait(&acc1_empty[as], ((i >> 1) & 1) ^ 1);
        mbar_wait(&a_full[sa], pa);
        tc_fence_after();
        const bool tr = kPairTrace && p.trace && blockIdx.x == 0 && i < 256 && lane == 0;
        if (tr) p.trace[i * 12 + 0] = clock64();
        const uint32_t d_base = tmem_base + as * ACC_COLS;
        uint32_t at = a_lo0 + sa * a_stage16, wt = w_lo0;
        for (int tap = 0; tap < k; ++tap, at += tap_step16, wt += B_STAGE >> 4) {
#pragma unroll
          for (int acc = 0; acc < NACC; ++acc)
#pragma unroll
            for (int kk = 0; kk < KC / 16; ++kk)
              umma_f16_lohi(d_base + acc * CH, at + ((acc * 128 * ROWB + kk * 32) >> 4), desc_hi, wt + ((kk * 32) >> 4),
                            desc_hi, idesc, (tap > 0 || kk > 0) ? 1u : 0u, leader);
        }
        if (leader) {
          umma_commit(&acc1_full[as]);
          umma_commit(&a_empty[sa]);
        }
        if (tr) p.trace[i * 12 + 1] = clock64();
        if (++sa == (uint32_t)NA) { sa = 0; pa ^= 1; }
      }
    } else {
      uint32_t hb = 0, ph = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const uint32_t as = i & 1;
        mbar_wait(&h_full[hb], ph);
        mbar_wait(&acc2_empty[as], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const bool tr = kPairTrace && p.trace && blockIdx.x == 0 && i < 256 && lane == 0;
        if (tr) p.trace[i * 12 + 2] = clock64();
        const uint32_t d_base = tmem_base + 2 * ACC_COLS + as * ACC_COLS;
        uint32_t ht = h_lo0 + hb * h_buf16, wt = w_lo0 + k * (B_STAGE >> 4);
        for (int tap = 0; tap < k; ++tap, ht += ROWB >> 4, wt += B_STAGE >> 4) {
#pragma unroll
          for (int acc = 0; acc < NACC; ++acc)
#pragma unroll
            for (int kk = 0; kk < KC / 16; ++kk)
              umma_f16_lohi(d_base + acc * CH, ht + ((acc * 128 * ROWB + kk * 32) >> 4), desc_hi, wt + ((kk * 32) >> 4),
                            desc_hi, idesc, (tap > 0 || kk > 0) ? 1u : 0u, leader);
        }
        if (leader) {
          umma_commit(&acc2_full[as]);
          umma_commit(&h_empty[hb]);
        }
        if (tr) p.trace[i * 12 + 3] = clock64();
        if (++hb == (uint32_t)NH) { hb = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int q = warp & 3;
    const int grp = (warp - 2) >> 3;          // 0: epi1 (h producer), 1: epi2 (output)
    const int hsel = ((warp - 2) & 7) >> 2;   // which of the group's two warps on this TMEM lane quadrant
    uint8_t* scratch = scratch_base + (warp - 2) * 1024;
    const int L = p.L, C = CH;
    const float slope = p.slope, res_gain = p.res_gain;
    __nv_bfloat16* const out = p.out;
    pdl_wait();   // output stores may overwrite a buffer the previous launch still reads

    // h = lrelu(c1 + b1), zero outside the utterance, written as c2's swizzled K-major A operand
    uint32_t hb = 0, ph = 0;   // epi1: h buffer / phase counters;  epi2: activation stage counter (no runtime i % N)
    uint32_t sa = 0;
    // A warp's items always cover the same NBS 16-column groups (item it = hsel + NW*m -> columns 16*(it % CHUNKS)):
    // their biases live in registers instead of being re-read from shared memory for every item, where the loads queue
    // behind the tensor core's operand reads
    constexpr int NBS = CHUNKS / NW;
    float4 breg[NBS][4];
#pragma unroll
    for (int s2 = 0; s2 < NBS; ++s2)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        breg[s2][j] = *reinterpret_cast<const float4*>(sbias + grp * CH + (hsel + NW * s2) * 16 + 4 * j);
    auto epi1 = [&](int i) {
      const uint32_t tile = blockIdx.x + i * gridDim.x;
      uint32_t b, mt;
      p.div_m.divmod(tile, b, mt);
      const int t0 = mt * p.bmo;
      const uint32_t as = i & 1;
      uint8_t* const hbuf = smemH + hb * HROWS * ROWB;
      if (kPairTrace && p.trace && blockIdx.x == 0 && warp == 2 && lane == 0 && i < 256) p.trace[i * 12 + 9] = clock64();
      mbar_wait(&acc1_full[as], (i >> 1) & 1);
      tc_fence_after();
      const bool tr = kPairTrace && p.trace && blockIdx.x == 0 && warp == 2 && lane == 0 && i < 256;
      if (tr) p.trace[i * 12 + 4] = clock64();
      bool h_free = false;
#pragma unroll
      for (int m = 0; m < NITEMS / NW; ++m) {
        const int it = hsel + NW * m;
        const int acc = it / CHUNKS, c0 = (it % CHUNKS) * 16;
        const float4 (&bv)[4] = breg[m % NBS];
        const int r = acc * 128 + q * 32 + lane;
        const int th = t0 - hk + r;
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
            const int j = h2 * 8 + e * 2;
            const float4 bq = bv[j >> 2];
            float v0 = __uint_as_float(a[j]) + ((j & 3) == 0 ? bq.x : bq.z);
            float v1 = __uint_as_float(a[j + 1]) + ((j & 3) == 0 ? bq.y : bq.w);
            v0 = inside ? fmaxf(v0, v0 * slope) : 0.f;
            v1 = inside ? fmaxf(v1, v1 * slope) : 0.f;
            o2[e] = pack_act2<F16>(v0, v1);
          }
        }
        if (!h_free) {  // the c2 that last read this h buffer must have retired before it is overwritten
          mbar_wait(&h_empty[hb], ph ^ 1);
          h_free = true;
        }
        const uint32_t sw = swz_row<ROWB>(r);
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
      if (++hb == (uint32_t)NH) { hb = 0; ph ^= 1; }
      if (tr) p.trace[i * 12 + 5] = clock64();
    };

    // y = lrelu(c2 + b2 + x), x recovered from the resident activation tile; coalesced channels-last store
    auto epi2 = [&](int i) {
      const uint32_t tile = blockIdx.x + i * gridDim.x;
      uint32_t b, mt;
      p.div_m.divmod(tile, b, mt);
      const int t0 = mt * p.bmo;
      const uint32_t as = i & 1;
      const uint8_t* atile = smemA + sa * p.a_stage_bytes;
      if (kPairTrace && p.trace && blockIdx.x == 0 && warp == 2 + kPairEpiWarps / 2 && lane == 0 && i < 256)
        p.trace[i * 12 + 8] = clock64();
      mbar_wait(&acc2_full[as], (i >> 1) & 1);
      tc_fence_after();
      const bool tr = kPairTrace && p.trace && blockIdx.x == 0 && warp == 2 + kPairEpiWarps / 2 && lane == 0 && i < 256;
      if (tr) p.trace[i * 12 + 6] = clock64();
#pragma unroll
      for (int m = 0; m < NITEMS / NW; ++m) {
        const int it = hsel + NW * m;
        const int acc = it / CHUNKS, c0 = (it % CHUNKS) * 16;
        const float4 (&bv)[4] = breg[m % NBS];
        const int i0 = acc * 128 + q * 32;            // first output row of this warp's 32
        const int ra = i0 + lane + hk + hk * p.dil;   // row of x(t0 + i0 + lane) in the activation tile
        uint32_t a[16];
        __syncwarp();
        tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + 2 * ACC_COLS + as * ACC_COLS + acc * CH + c0, a);
        const uint32_t sw = swz_row<ROWB>(ra);
        uint4 rx[2];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) rx[h2] = ld_shared_u4(atile + ra * ROWB + ((((c0 >> 3) + h2) ^ sw) << 4));
        tmem_ld_wait();
        uint4 ov[2];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const uint32_t* r2 = reinterpret_cast<const uint32_t*>(&rx[h2]);
          uint32_t* o2 = reinterpret_cast<uint32_t*>(&ov[h2]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = h2 * 8 + e * 2;
            const float4 bq = bv[j >> 2];
            const float2 xr = unpack_act2<F16>(r2[e]);
            float v0 = __uint_as_float(a[j]) + ((j & 3) == 0 ? bq.x : bq.z) + (xr.x >= 0.f ? xr.x : xr.x * res_gain);
            float v1 = __uint_as_float(a[j + 1]) + ((j & 3) == 0 ? bq.y : bq.w) + (xr.y >= 0.f ? xr.y : xr.y * res_gain);
            o2[e] = pack_act2<F16>(fmaxf(v0, v0 * slope), fmaxf(v1, v1 * slope));
          }
        }
        // this thread's row segment is one aligned 32-byte sector: a single 256-bit store, no shared-memory transposition
        // (the scratch round trip queued behind the tensor core's operand reads, profiles/r01_trace_pair.txt)
        const int rows_valid = min(32, max(0, min(p.bmo - i0, L - (t0 + i0))));
        if (lane < rows_valid) st_global_v8(out + ((long)b * L + t0 + i0 + lane) * C + c0, ov[0], ov[1]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&acc2_empty[as]);
        mbar_arrive(&a_empty[sa]);
      }
      if (++sa == (uint32_t)NA) sa = 0;
      if (tr) p.trace[i * 12 + 7] = clock64();
    };

    if (grp == 0) {
      for (int i = 0; i < my_tiles; ++i) epi1(i);
    } else {
      for (int i = 0; i < my_tiles; ++i) epi2(i);
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

static constexpr int kPairSmemBudget = 227 * 1024 - 1024 /*align*/ - 256 /*barriers*/ - 1024 /*bias*/ - 16384 /*scratch*/;

// rows of h per tile: 256 when three activation stages + both weight sets + h fit, else 128, else 0 (unsupported)
static int pair_tile_rows(int channels, int k, int dil) {
  if (channels != 32 && channels != 64) return 0;
  if (k % 2 == 0 || k > 15) return 0;
  const int rowb = channels * 2;
  for (int rows : {256, 128}) {
    const int nboxes = (rows + (k - 1) * dil + 63) / 64;
    const int a_stage = nboxes * 64 * rowb;
    const int hrows = rows == 256 ? kPairHRows : 144;
    const int need = 3 * a_stage + 2 * k * channels * rowb + hrows * rowb;  // one h buffer is the minimum
    if (rows == 128 && channels != 64) continue;  // the 128-row form needs >= 4 epilogue items per tile
    if (need <= kPairSmemBudget) return rows;
  }
  return 0;
}

bool pair_supported(int channels, int k, int dil) { return pair_tile_rows(channels, k, dil) != 0; }

int plan_conv_pair(PairPlan* pl, int B, int L, int channels, int k, int dil, const __nv_bfloat16* x,
                   const __nv_bfloat16* w_pair, int num_sms) {
  VD_CHECK(pair_supported(channels, k, dil), "conv_pair: unsupported shape");
  PairParams& p = pl->p;
  p.B = B; p.L = L; p.k = k; p.dil = dil; p.hk = (k - 1) / 2;
  const int rows = pair_tile_rows(channels, k, dil);
  pl->nacc = rows / 128;
  p.bmo = rows - (k - 1);
  const int rowb = channels * 2;
  p.nboxes = (rows + (k - 1) * dil + 63) / 64;
  p.a_stage_bytes = p.nboxes * 64 * rowb;
  const int hbytes = (rows == 256 ? kPairHRows : 144) * rowb;
  int fixed = 2 * k * channels * rowb + hbytes;
  p.nh = 1;
  if (fixed + hbytes + 3 * p.a_stage_bytes <= kPairSmemBudget) {  // double-buffer h when there is room
    p.nh = 2;
    fixed += hbytes;
  }
  p.na_stages = std::min(kPairMaxNA, (kPairSmemBudget - fixed) / p.a_stage_bytes);
  p.m_tiles = (L + p.bmo - 1) / p.bmo;
  p.total_tiles = B * p.m_tiles;
  p.div_m.init(p.m_tiles);
  p.trace = nullptr;
  pl->channels = channels;
  pl->pdl = false;
  pl->grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  pl->smem = 1024 + (size_t)p.na_stages * p.a_stage_bytes + fixed + 256 + 1024 + 16384;
  if (encode_tmap_3d(&pl->tmA, x, channels, L, B, channels, 64, true)) return 1;
  if (encode_tmap_3d(&pl->tmW, w_pair, channels, channels, 2 * k, channels, channels, true)) return 1;
  return 0;
}

template <int CH, int NACC, bool F16>
static int launch_pair_typed(const PairPlan& pl, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VD_CUDA(cudaFuncSetAttribute(conv_pair_kernel<CH, CH, NACC, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024));
    attr_set = true;
  }
  const size_t smem = pl.smem > (size_t)120 * 1024 ? pl.smem : (size_t)120 * 1024;   // one CTA per SM, see conv_tc.cu
  if (pl.pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pl.grid);
    cfg.blockDim = dim3(kPairThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    VD_CUDA(cudaLaunchKernelEx(&cfg, conv_pair_kernel<CH, CH, NACC, F16>, pl.tmA, pl.tmW, pl.p));
    return 0;
  }
  conv_pair_kernel<CH, CH, NACC, F16><<<pl.grid, kPairThreads, smem, stream>>>(pl.tmA, pl.tmW, pl.p);
  VD_CUDA(cudaGetLastError());
  return 0;
}

template <int CH, int NACC>
static int launch_pair_inst(const PairPlan& pl, cudaStream_t stream) {
  return pl.p.f16 ? launch_pair_typed<CH, NACC, true>(pl, stream) : launch_pair_typed<CH, NACC, false>(pl, stream);
}

int launch_conv_pair(PairPlan& pl, const float* bias1, const float* bias2, float slope, __nv_bfloat16* out,
                     cudaStream_t stream, int f16) {
  pl.p.f16 = f16;
  pl.p.bias1 = bias1;
  pl.p.bias2 = bias2;
  pl.p.slope = slope;
  pl.p.res_gain = 1.f / slope;
  pl.p.out = out;
  if (pl.channels == 32 && pl.nacc == 2) return launch_pair_inst<32, 2>(pl, stream);
  if (pl.channels == 64 && pl.nacc == 2) return launch_pair_inst<64, 2>(pl, stream);
  if (pl.channels == 64 && pl.nacc == 1) return launch_pair_inst<64, 1>(pl, stream);
  set_error("conv_pair: no kernel instance");
  return 1;
}

}  // namespace vd
