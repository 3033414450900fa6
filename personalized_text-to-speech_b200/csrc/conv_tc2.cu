// The 256-output-channel layers (stage 0 of the shipped configuration) as PAIRED tensor-core tiles: tcgen05.mma.cta_group::2,
// M = 256 output channels across the two CTAs of a cluster, N = 256 time rows, fp32 accumulators in both CTAs' TMEM.
//
// Why.  conv_tc.cu's channels-as-M tile (M = 128, N = 256) fetches 4 KB of weights + 8 KB of activations from shared
// memory per 128-cycle MMA: 96 of the 128 wavefronts a cycle budget allows, and the weight ring's TMA writes (32), the
// activation tile's (7) and the epilogue's staging (15) bring it to ~150 -- the measured 163-168 cycles per MMA
// (profiles/r02_ncu_stage1_k11.txt).  A paired MMA shares the activation operand between the two SMs: CTA r
// holds the weights of ITS 128 output channels (A operand, as before) and HALF of the tile's time rows (B operand: rows
// [128 r, 128 r + 128) plus the taps' halo); the hardware exchanges the halves.  Per CTA and MMA: 32 + 32 operand + 32 ring
// + 4 tile + 15 epilogue = 115 wavefronts for 128 cycles, and half the activation bytes through L2 / TMA.
//
// Protocol (rank 0 = leader: it alone issues MMAs; every barrier sits at the same shared-memory offset in both CTAs):
//   a_full / b_full   each CTA's TMA producer fills its own stages with cp.async.bulk.tensor.cta_group::2 loads whose
//                     transaction bytes are counted on the LEADER's barrier (shared::cluster address of rank 0); the
//                     leader's producer expects the bytes of both halves, so one wait tells the leader that both halves
//                     of the operands have landed.  (First version, ConvTcParams::cta2_relay: local barriers + a relay
//                     thread in the peer that forwards each completion with a remote arrive -- two more hops per stage.)
//   a_empty / b_empty / acc_full   tcgen05.commit with cluster multicast: one commit arrives in both CTAs.
//   acc_empty         leader only, 2 x 16 arrivals: its own epilogue warps and (remotely) the peer's.
// Tiles: (utterance, 256-row block); cluster c takes tiles c, c + #clusters, ...  Both CTAs always have the same tiles, so
// the rings of the two CTAs advance in lockstep by construction.  Epilogues are conv_tc.cu's channels-as-M epilogues
// (epilogue.cuh) on each CTA's 128 channels.  Verified instruction forms: tools/probes/cta2_probe.cu.
#include <algorithm>

#include "common.cuh"
#include "conv_tc.h"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace vd {

constexpr int kT2EpiWarps = 16;
constexpr int kT2Threads = 64 + 32 * kT2EpiWarps;
constexpr int kT2MaxNA = 8, kT2MaxNB = 8;

// ---------------------------------------------------------------- cluster / cta_group::2 forms (see the probe)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait with cluster-scope acquire: the phase may have been completed by the peer CTA's arrive
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0, ok = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 20)) {
      printf("vitsdec: cluster mbarrier timeout block %d thread %d bar@%u parity %u\n", (int)blockIdx.x, (int)threadIdx.x,
             smem_u32(bar), parity);
      __trap();
    }
  }
}
// TMA loads of a CTA pair: the data lands in THIS CTA's shared memory, the transaction bytes are counted on a barrier
// given by its shared::cluster address -- the leader's (mapa rank 0) -- so one barrier collects both CTAs' halves
__device__ __forceinline__ void tma2_load_3d(const CUtensorMap* m, uint32_t bar_cluster, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma2_load_5d(const CUtensorMap* m, uint32_t bar_cluster, void* dst, int c0, int c1, int c2,
                                             int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.ne.b32 q, %7, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
// all MMAs issued so far by this thread arrive (once) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)0x3)
      : "memory");
}
// kind::f16 instruction descriptor with M = 256 (the pair's rows)
__host__ __device__ constexpr uint32_t umma_idesc_f16_m256(int n, bool fp16_operands) {
  return (1u << 4) | ((fp16_operands ? 0u : 1u) << 7) | ((fp16_operands ? 0u : 1u) << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(256 >> 4) << 24);
}

template <int EPI, bool F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kT2Threads, 1)
conv_tc2_kernel(const __grid_constant__ TmapPack tm, const __grid_constant__ CUtensorMap tmW,
                const __grid_constant__ ConvTcParams p) {
  constexpr int KC = 64, ROWB = 128, BNH = 128;   // K-chunk, its row bytes, output channels per CTA
  constexpr int B_STAGE = BNH * ROWB;             // one (tap, K-chunk) of this CTA's channels: 16 KB
  constexpr int ACC_COLS = 256, NBUF = 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smemA = smem;
  const int NA = p.na_stages, NB = p.nb_stages;
  uint8_t* smemB = smem + NA * p.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemB + p.b_region_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kT2MaxNA;
  uint64_t* b_full = a_empty + kT2MaxNA;
  uint64_t* b_empty = b_full + kT2MaxNB;
  uint64_t* acc_full = b_empty + kT2MaxNB;
  uint64_t* acc_empty = acc_full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 8 + 1);
  float* sbias = reinterpret_cast<float*>(bars + 64);   // 512 bytes of barrier space, then this CTA's 128 biases

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int n0 = (int)rank * BNH;                 // this CTA's output channels [n0, n0 + 128)

  if (warp == 0 && lane == 0) {
    for (int sg = 0; sg < p.g.nseg; ++sg) tma_prefetch_desc(&tm.a[sg]);
    tma_prefetch_desc(&tmW);
    // the leader's full barriers complete on its own producer's expect_tx AND the peer relay's remote arrive
    const uint32_t full_count = (rank == 0 && p.cta2_relay) ? 2 : 1;
    for (int i = 0; i < kT2MaxNA; ++i) { mbar_init(&a_full[i], full_count); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kT2MaxNB; ++i) { mbar_init(&b_full[i], full_count); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 8; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 2 * kT2EpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc2(tmem_slot, NBUF * ACC_COLS);
    tmem_relinquish2();
  }
  for (int i = threadIdx.x; i < BNH; i += kT2Threads) sbias[i] = p.ep.bias[n0 + i];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers and TMEM are in place before anything arrives remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nkc = p.g.c_in / KC;
  const int total_tiles = p.total_tiles;   // (utterance, 256-row block) pairs

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: this CTA's operand halves
    if (lane == 0) {
      const bool relay = p.cta2_relay != 0;
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
        uint32_t b, mt;
        p.div_m.divmod(tile, b, mt);
        const int t0 = (int)mt * 256;
        if (p.res_prefetch) {
          for (int i = 0; i < p.ep.nres; ++i)
            for (int r = 0; r < 256; r += 64)
              for (int c = 0; c < BNH; c += 64) tma_prefetch_3d(&tm.r[i], n0 + c, t0 + r, (int)b);
        }
        int tap0 = 0;
        for (int sg = 0; sg < p.g.nseg; ++sg) {
          const int tap1 = p.g.seg_tap_end[sg];
          const int nbx = p.seg_nboxes[sg];
          for (int kc = 0; kc < nkc; ++kc) {
            mbar_wait(&a_empty[sa], pa ^ 1);
            // rows [t0 + 128 rank + halo_lo, ...): this CTA's half of the tile's rows
            const int row0 = t0 + (int)rank * 128 + p.seg_halo_lo[sg];
            uint8_t* adst = smemA + sa * p.a_stage_bytes;
            if (relay) {
              mbar_expect_tx(&a_full[sa], nbx * 64 * ROWB);
              for (int bx = 0; bx < nbx; ++bx)
                tma_load_5d(&tm.a[sg], &a_full[sa], adst + bx * 64 * ROWB, 0, kc, 0, row0 + bx * 64, (int)b);
            } else {
              // both CTAs' bytes are counted on the LEADER's barrier (its producer expects them all)
              if (rank == 0) mbar_expect_tx(&a_full[sa], 2 * nbx * 64 * ROWB);
              const uint32_t bar = mapa_u32(&a_full[sa], 0);
              for (int bx = 0; bx < nbx; ++bx)
                tma2_load_5d(&tm.a[sg], bar, adst + bx * 64 * ROWB, 0, kc, 0, row0 + bx * 64, (int)b);
            }
            if (++sa == (uint32_t)NA) { sa = 0; pa ^= 1; }
            for (int tap = tap0; tap < tap1; ++tap) {
              mbar_wait(&b_empty[sb], pb ^ 1);
              if (relay) {
                mbar_expect_tx(&b_full[sb], B_STAGE);
                tma_load_3d(&tmW, &b_full[sb], smemB + sb * B_STAGE, kc * KC, n0, tap);
              } else {
                if (rank == 0) mbar_expect_tx(&b_full[sb], 2 * B_STAGE);
                tma2_load_3d(&tmW, mapa_u32(&b_full[sb], 0), smemB + sb * B_STAGE, kc * KC, n0, tap);
              }
              if (++sb == (uint32_t)NB) { sb = 0; pb ^= 1; }
            }
          }
          tap0 = tap1;
        }
      }
    }
  } else if (warp == 1 && rank == 1) {
    // ------------------------------------------------------------ peer relay: "my half has landed" -> the leader's barriers
    if (lane == 0 && p.cta2_relay) {   // (only with the relay protocol; by default the TMA loads signal the leader directly)
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
        int tap0 = 0;
        for (int sg = 0; sg < p.g.nseg; ++sg) {
          const int tap1 = p.g.seg_tap_end[sg];
          for (int kc = 0; kc < nkc; ++kc) {
            mbar_wait(&a_full[sa], pa);
            mbar_arrive_remote(mapa_u32(&a_full[sa], 0));
            if (++sa == (uint32_t)NA) { sa = 0; pa ^= 1; }
            for (int tap = tap0; tap < tap1; ++tap) {
              mbar_wait(&b_full[sb], pb);
              mbar_arrive_remote(mapa_u32(&b_full[sb], 0));
              if (++sb == (uint32_t)NB) { sb = 0; pb ^= 1; }
            }
          }
          tap0 = tap1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA; warp-uniform, elected lane issues)
    const uint32_t idesc = umma_idesc_f16_m256(256, F16);
    constexpr uint32_t desc_hi = umma_desc_hi(ROWB);
    const uint32_t leader = elect_one();
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(smemA)), b_lo0 = umma_desc_lo(smem_u32(smemB));
    const uint32_t a_stage16 = (uint32_t)p.a_stage_bytes >> 4;
    uint32_t itt = 0, sa = 0, pa = 0, sb = 0, pb = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++itt) {
      const uint32_t as = itt & 1, pacc = (itt >> 1) & 1;
      mbar_wait_cluster(&acc_empty[as], pacc ^ 1);
      tc_fence_after();
      const uint32_t d_base = tmem_base + as * ACC_COLS;
      uint32_t accum = 0;
      int tap0 = 0;
      for (int sg = 0; sg < p.g.nseg; ++sg) {
        const int tap1 = p.g.seg_tap_end[sg];
        for (int kc = 0; kc < nkc; ++kc) {
          mbar_wait_cluster(&a_full[sa], pa);
          tc_fence_after();
          const uint32_t a_lo_stage = a_lo0 + sa * a_stage16;
          for (int tap = tap0; tap < tap1; ++tap) {
            mbar_wait_cluster(&b_full[sb], pb);
            tc_fence_after();
            const uint32_t w_lo = b_lo0 + sb * (B_STAGE >> 4);
            const uint32_t x_lo = a_lo_stage + p.tap_delta16[tap];
#pragma unroll
            for (int k = 0; k < KC / 16; ++k)   // D[channel, time] += W[channel, ci] * X[time, ci]^T
              umma2_f16_lohi(d_base, w_lo + ((k * 32) >> 4), desc_hi, x_lo + ((k * 32) >> 4), desc_hi, idesc,
                             k == 0 ? accum : 1u, leader);
            accum = 1;
            if (leader) umma2_commit_mc(&b_empty[sb]);
            if (++sb == (uint32_t)NB) { sb = 0; pb ^= 1; }
          }
          if (leader) umma2_commit_mc(&a_empty[sa]);
          if (++sa == (uint32_t)NA) { sa = 0; pa ^= 1; }
        }
        tap0 = tap1;
      }
      if (leader) umma2_commit_mc(&acc_full[as]);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue warps (both CTAs): 128 channels x 256 rows each
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    constexpr int NW = kT2EpiWarps / 4, NITEMS = 256 / kIW;
    uint8_t* scratch = reinterpret_cast<uint8_t*>(sbias) + 8192 + (warp - 2) * 2048;
    const FastDiv div_m = p.div_m;
    const int L = p.g.L, n_total = p.g.n_total, rowstride = p.rowstride;
    const long bstride = p.bstride;
    const float out_slope = p.ep.out_slope, mrf_scale = p.ep.mrf_scale, res_gain = p.ep.res_gain;
    ConvEpilogue ep = p.ep;
    const bool tma_out = p.tma_epi != 0;
    const uint32_t acc_empty_leader0 = mapa_u32(&acc_empty[0], 0), acc_empty_leader1 = mapa_u32(&acc_empty[1], 0);
    uint32_t nitem = 0;
    auto coords = [&](int tile, int it, EpiItem& e) {
      uint32_t b, mt;
      div_m.divmod(tile, b, mt);
      const int t = (int)mt * 256 + it * kIW;
      e.b = (int)b;
      e.n = n0 + q * 32;
      e.rows_valid = min(kIW, max(0, L - t));
      e.row0 = (long)e.b * L + t;
      e.base = (long)e.b * bstride + e.n + (long)t * rowstride;
      e.tcol = it * kIW;
      e.t = t;
    };
    int tile = cluster_id, it = hsel;
    uint32_t itt = 0;
    EpiLoads ld;
    EpiItem cur{};
    if (tile < total_tiles) {
      coords(tile, it, cur);
      epiT_issue_loads<EPI>(ep, cur, rowstride, lane, ld);
    }
    while (tile < total_tiles) {
      const bool first = it < NW, last = it + NW >= NITEMS;
      const uint32_t as = itt & 1, pacc = (itt >> 1) & 1;
      if (first) {
        mbar_wait(&acc_full[as], pacc);
        tc_fence_after();
      }
      uint32_t acc[kIW];
      float v[kIW];
      __syncwarp();
      tmem_ld_frag(tmem_base + ((uint32_t)(q * 32) << 16) + as * ACC_COLS + cur.tcol, acc);
      float bias4[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) bias4[m] = sbias[q * 32 + 8 * m + (lane >> 2)];
      tmem_ld_wait();
      uint8_t* buf = scratch;
      if (tma_out) {   // the store that last read this buffer (two items ago) must have drained it
        buf = scratch + (nitem & 1) * 1024;
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
      }
      epiT_accumulate<EPI, F16>(ep, bias4, buf, cur, n_total, lane, res_gain, acc, ld, v);
      if (last) {   // accumulator fully read by this warp: hand the buffer back to the leader's MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive(&acc_empty[as]);
          else mbar_arrive_remote(as ? acc_empty_leader1 : acc_empty_leader0);
        }
        ++itt;
      }
      const EpiItem done = cur;
      int ntile = tile, nit = it + NW;
      if (nit >= NITEMS) { nit = hsel; ntile += n_clusters; }
      if (ntile < total_tiles) {
        coords(ntile, nit, cur);
        epiT_issue_loads<EPI>(ep, cur, rowstride, lane, ld);
      }
      if (tma_out) epiT_store_tma<EPI, F16>(&tm.o, buf, done, lane, out_slope, mrf_scale, v);
      else epiT_store<EPI, F16>(ep, buf, done, rowstride, lane, out_slope, mrf_scale, v);
      ++nitem;
      tile = ntile; it = nit;
    }
    if (tma_out && lane == 0) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // neither CTA may retire while the peer can still arrive on its barriers or read its operands
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, NBUF * ACC_COLS);
  }
}

// ------------------------------------------------------------------------------------------- host side
template <int EPI, bool F16>
static int launch_tc2_typed(const ConvTcPlan& pl, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VD_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<EPI, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const size_t smem = pl.smem > (size_t)120 * 1024 ? pl.smem : (size_t)120 * 1024;   // one CTA per SM, see conv_tc.cu
  conv_tc2_kernel<EPI, F16><<<pl.grid, kT2Threads, smem, stream>>>(pl.tm, pl.tmW, pl.p);
  VD_CUDA(cudaGetLastError());
  return 0;
}

template <int EPI>
static int launch_tc2_epi(const ConvTcPlan& pl, cudaStream_t stream) {
  return pl.p.ep.f16 ? launch_tc2_typed<EPI, true>(pl, stream) : launch_tc2_typed<EPI, false>(pl, stream);
}

int launch_conv_tc2(const ConvTcPlan& pl, cudaStream_t stream) {
  const ConvEpilogue& e = pl.p.ep;
  VD_CHECK(e.bias_b == nullptr && e.mrf == nullptr && !e.split_col && !e.gate && e.rowmask == nullptr,
           "conv_tc2: specialised epilogues only");
  if (e.mrf_mode == 0 && e.nres == 0) return launch_tc2_epi<1>(pl, stream);
  if (e.mrf_mode == 0 && e.nres == 1) return launch_tc2_epi<2>(pl, stream);
  if (e.mrf_mode == 3 && e.nres == 3) return launch_tc2_epi<3>(pl, stream);
  set_error("conv_tc2: no kernel instance for this epilogue");
  return 1;
}

// largest number of 2-CTA clusters of this kernel that can be resident at once (GPCs with an odd number of free SMs
// strand one): the persistent grid is sized to it so that every cluster of a launch runs concurrently
int max_clusters_tc2(size_t smem) {
  static int cached = -1;
  if (cached >= 0) return cached;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148);
  cfg.blockDim = dim3(kT2Threads);
  cfg.dynamicSmemBytes = smem > (size_t)120 * 1024 ? smem : (size_t)120 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaFuncSetAttribute(conv_tc2_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
      cudaOccupancyMaxActiveClusters(&n, conv_tc2_kernel<1, false>, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = 0;
  }
  cached = n;
  return n;
}

}  // namespace vd
