// Epilogue building blocks shared by the tcgen05 convolution kernels (conv_tc.cu, conv_pairf.cu): work items of
// 32 TMEM lanes x 16 columns, coalesced global I/O through a per-warp shared-memory transposition.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace vd {

constexpr int kIW = 16;        // epilogue work item: 32 rows (TMEM lanes of one warp) x kIW output columns

// Epilogue work item = 32 rows x 16 output columns per warp (one row per thread).  Global reads (residuals) are
// issued one item AHEAD of their use so their DRAM/L2 latency overlaps the TMEM load, the math and the stores of the
// current item.  Sixteen epilogue warps (four per scheduler) because the per-item instruction stream is a long
// dependent chain: with two warps per scheduler it ran at ~7 cycles/instruction and bounded every k=3 layer
// (profiles/r01_trace_probe.txt).
//
// Per-warp 1 KB transpose scratch: a 32-row x 32-byte item, 16-byte chunks XOR-swizzled so that both access patterns
// below are bank-conflict free.  Threads OWN rows for the math (TMEM lane == row), but global memory wants each warp
// instruction to cover contiguous row segments (16 rows x 32 B per LDG/STG.128 instead of 32 rows x 16 B).
__device__ __forceinline__ uint32_t scr_off(int row, int chunk) { return row * 32 + ((chunk ^ ((row >> 2) & 1)) << 4); }

constexpr int kMaxRes = kMaxSeg - 1;
struct EpiLoads {
  uint4 res[kMaxRes][2];  // residual tensors, COALESCED mapping: element j = row 16*j + lane/2, chunk lane%2
};

__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// One epilogue work item: rows [row0, row0+32) x columns [n, n+kIW) of utterance b; rows_valid of them exist.
struct EpiItem {
  int b, n, rows_valid;
  int t;           // channels-as-M: first time row of the item inside its utterance (TMA coordinates)
  long row0;       // b*L + t of lane 0's row
  long base;       // channels-as-M: element offset of (first row of the item, this warp's first channel)
  uint32_t tcol;   // TMEM column of the item inside its accumulator buffer
};

// EPI specialisation: straight-line epilogue code for the common cases (the generic path re-tests half a dozen
// launch constants per item, and a lone warp pays ~20-30 cycles per resolved branch):
//   0 generic (runtime flags: per-utterance bias, any residual count, fp32 MRF fallback)
//   1 plain (bias + leaky-relu)      2 one residual      3 three residuals + 1/nk scale (fused MRF)
//   4 conv_post on the folded view (tanh, fp32 waveform; channels-as-M only)
template <int EPI>
__device__ __forceinline__ bool epi_has_res(const ConvEpilogue& ep, int i) {
  if constexpr (EPI == 0) return i < ep.nres;
  if constexpr (EPI == 1 || EPI == 4) return false;
  if constexpr (EPI == 2) return i < 1;
  return i < 3;
}

template <int EPI>
__device__ __forceinline__ void epi_issue_loads(const ConvEpilogue& ep, const EpiItem& it, int n_total, int lane,
                                                EpiLoads& ld) {
  if (EPI == 0 && ep.split_col) {   // split launch: only the low columns carry a residual, with their own row stride
    if (it.n >= ep.split_col) return;
    n_total = ep.split_col;
  }
#pragma unroll
  for (int i = 0; i < kMaxRes; ++i) {
    if (epi_has_res<EPI>(ep, i)) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int row = 16 * j + (lane >> 1);
        if (row < it.rows_valid)
          ld.res[i][j] = ld_stream_u4(reinterpret_cast<const uint4*>(ep.res[i] + (it.row0 + row) * n_total + it.n) +
                                      (lane & 1));
      }
    }
  }
}

// v = acc + bias (+ per-utterance bias) (+ residuals) (+ MRF accumulator)
template <int EPI, bool F16 = false>
__device__ __forceinline__ void epi_accumulate(const ConvEpilogue& ep, const float4 (&bias)[4], uint8_t* scratch,
                                               const EpiItem& it, int n_total, int lane, float res_gain,
                                               const uint32_t (&acc)[kIW], const EpiLoads& ld, float (&v)[kIW]) {
#pragma unroll
  for (int j = 0; j < kIW; j += 4) {
    const float4 bv = bias[j >> 2];
    v[j + 0] = __uint_as_float(acc[j + 0]) + bv.x;
    v[j + 1] = __uint_as_float(acc[j + 1]) + bv.y;
    v[j + 2] = __uint_as_float(acc[j + 2]) + bv.z;
    v[j + 3] = __uint_as_float(acc[j + 3]) + bv.w;
  }
  if (EPI == 0 && ep.bias_b) {
    const float* bb = ep.bias_b + (long)it.b * n_total + it.n;
#pragma unroll
    for (int j = 0; j < kIW; j += 4) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(bb + j));
      v[j + 0] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
    }
  }
  const bool split_hi = EPI == 0 && ep.split_col && it.n >= ep.split_col;
#pragma unroll
  for (int i = 0; i < kMaxRes; ++i) {
    if (epi_has_res<EPI>(ep, i) && !split_hi) {
      // coalesced registers -> scratch -> row-owner registers
#pragma unroll
      for (int j = 0; j < 2; ++j)
        *reinterpret_cast<uint4*>(scratch + scr_off(16 * j + (lane >> 1), lane & 1)) = ld.res[i][j];
      __syncwarp();
      uint4 mine[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) mine[c] = *reinterpret_cast<const uint4*>(scratch + scr_off(lane, c));
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint32_t* r2 = reinterpret_cast<const uint32_t*>(&mine[q]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = unpack_act2<F16>(r2[e]);
          v[q * 8 + e * 2 + 0] += a.x >= 0.f ? a.x : a.x * res_gain;
          v[q * 8 + e * 2 + 1] += a.y >= 0.f ? a.y : a.y * res_gain;
        }
      }
    }
  }
  if (EPI == 0 && ep.split_col) {
    if (split_hi && ep.mrf_mode == 2 && lane < it.rows_valid) {
      const float4* mp = reinterpret_cast<const float4*>(ep.mrf + (it.row0 + lane) * (n_total - ep.split_col) + it.n -
                                                         ep.split_col);
#pragma unroll
      for (int j = 0; j < kIW / 4; ++j) {
        const float4 m = mp[j];
        v[4 * j] += m.x; v[4 * j + 1] += m.y; v[4 * j + 2] += m.z; v[4 * j + 3] += m.w;
      }
    }
    return;
  }
  if (EPI == 0 && (ep.mrf_mode == 2 || (ep.mrf_mode == 3 && ep.mrf)) && lane < it.rows_valid) {  // fp32 fallback path
    const float4* mp = reinterpret_cast<const float4*>(ep.mrf + (it.row0 + lane) * n_total + it.n);
#pragma unroll
    for (int j = 0; j < kIW / 4; ++j) {
      const float4 m = mp[j];
      v[4 * j] += m.x; v[4 * j + 1] += m.y; v[4 * j + 2] += m.z; v[4 * j + 3] += m.w;
    }
  }
}

template <int EPI, bool F16 = false>
__device__ __forceinline__ void epi_store(const ConvEpilogue& ep, uint8_t* scratch, const EpiItem& it, int n_total,
                                          int lane, float out_slope, float mrf_scale, float (&v)[kIW]) {
  if (EPI == 0 && ep.rowmask) {  // `* x_mask`: one mask value per (utterance, time) row
    const float mk = lane < it.rows_valid ? __ldg(ep.rowmask + it.row0 + lane) : 0.f;
#pragma unroll
    for (int j = 0; j < kIW; ++j) v[j] *= mk;
  }
  if (EPI == 0 && ep.gate) {
    // gated activation of the flow's WN: the layer's output channels were interleaved at pack time so that this
    // thread's 16 columns are 8 (tanh input, sigmoid input) pairs -> 8 bf16 outputs = one 16-byte store per row
    uint4 ov;
    uint32_t* o2 = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float g0 = tanhf(v[4 * e]) * (1.f / (1.f + __expf(-v[4 * e + 1])));
      const float g1 = tanhf(v[4 * e + 2]) * (1.f / (1.f + __expf(-v[4 * e + 3])));
      o2[e] = pack_act2<F16>(g0, g1);
    }
    if (lane < it.rows_valid)
      *reinterpret_cast<uint4*>(ep.out + (it.row0 + lane) * (n_total / 2) + it.n / 2) = ov;
    return;
  }
  if (EPI == 0 && ep.split_col) {
    if (it.n >= ep.split_col) {   // skip half: fp32 accumulator
      if (lane < it.rows_valid) {
        float4* mp = reinterpret_cast<float4*>(ep.mrf + (it.row0 + lane) * (n_total - ep.split_col) + it.n - ep.split_col);
#pragma unroll
        for (int j = 0; j < kIW / 4; ++j) mp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      return;
    }
    n_total = ep.split_col;       // residual half: the ordinary bf16 store below, with its own row stride
  } else if (EPI == 0 && (ep.mrf_mode == 1 || ep.mrf_mode == 2)) {
    if (lane < it.rows_valid) {
      float4* mp = reinterpret_cast<float4*>(ep.mrf + (it.row0 + lane) * n_total + it.n);
#pragma unroll
      for (int j = 0; j < kIW / 4; ++j) mp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    return;
  }
  if (EPI == 3 || (EPI == 0 && ep.mrf_mode == 3)) {
#pragma unroll
    for (int j = 0; j < kIW; ++j) v[j] *= mrf_scale;
  }
  // row-owner registers -> scratch -> coalesced 32-byte row segments
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint4 ov;
    uint32_t* o2 = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      o2[e] = pack_act2<F16>(fmaxf(v[q * 8 + e * 2], v[q * 8 + e * 2] * out_slope),
                             fmaxf(v[q * 8 + e * 2 + 1], v[q * 8 + e * 2 + 1] * out_slope));  // slope in (0,1]
    *reinterpret_cast<uint4*>(scratch + scr_off(lane, q)) = ov;
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int row = 16 * j + (lane >> 1);
    const uint4 ov = *reinterpret_cast<const uint4*>(scratch + scr_off(row, lane & 1));
    if (row < it.rows_valid)
      *(reinterpret_cast<uint4*>(ep.out + (it.row0 + row) * n_total + it.n) + (lane & 1)) = ov;
  }
  __syncwarp();
}

// ---- SWAP epilogue (channels on TMEM lanes, time on TMEM columns) --------------------------------------------------
// Item = 32 channels (one per thread) x 16 time rows.  Global memory is [time][channel]: the item is 16 rows of 64
// bytes.  Scratch holds it row-major (64-byte rows, chunk-swizzled); the thread<->time transposition happens in the
// 2-byte shared-memory accesses, global accesses stay 16 bytes per lane over 8 rows x 64 B per instruction.
__device__ __forceinline__ uint32_t scrT_off(int row, int chunk) { return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4); }

template <int EPI>
__device__ __forceinline__ void epiT_issue_loads(const ConvEpilogue& ep, const EpiItem& it, int rowstride, int lane,
                                                 EpiLoads& ld) {
#pragma unroll
  for (int i = 0; i < kMaxRes; ++i) {
    if (epi_has_res<EPI>(ep, i)) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int row = 8 * j + (lane >> 2);
        if (ep.epi_smem) {
          if (row < it.rows_valid)
            ld.res[i][j] = ld_stream_u4(reinterpret_cast<const uint4*>(ep.res[i] + it.base + (long)row * rowstride) +
                                        (lane & 3));
        } else {
          // 8x8 blocks (time rows 8j + lane/4, channels 8m + 2(lane%4), +1) in the fragment layout of the [time][channel]
          // matrix: a register transpose (movmatrix) turns each into the accumulator's [channel][time] fragment
          uint32_t r[4] = {0u, 0u, 0u, 0u};
          if (row < it.rows_valid) {
            const __nv_bfloat16* p = ep.res[i] + it.base + (long)row * rowstride + 2 * (lane & 3);
#pragma unroll
            for (int m = 0; m < 4; ++m) r[m] = ld_stream_u32(p + 8 * m);
          }
          ld.res[i][j] = make_uint4(r[0], r[1], r[2], r[3]);
        }
      }
    }
  }
}

// Channels-as-M items are held in the mma-fragment layout (tmem_ld_16x256b_x2): thread t owns channels
// 8m + t/4 (m = 0..3) and columns (time rows) cg*8 + 2(t%4) + e.  Register index of (m, cg, e):
__device__ __forceinline__ constexpr int frag_idx(int m, int cg, int e) { return (m >> 1) * 8 + cg * 4 + (m & 1) * 2 + e; }

__device__ __forceinline__ void tmem_ld_frag(uint32_t taddr, uint32_t (&a)[kIW]) {
  tmem_ld_16x256b_x2(taddr, &a[0]);
  tmem_ld_16x256b_x2(taddr + (16u << 16), &a[8]);
}

template <int EPI, bool F16 = false>
__device__ __forceinline__ void epiT_accumulate(const ConvEpilogue& ep, const float (&bias)[4], uint8_t* scratch,
                                                const EpiItem& it, int n_total, int lane, float res_gain,
                                                const uint32_t (&acc)[kIW], const EpiLoads& ld, float (&v)[kIW]) {
  float bv[4] = {bias[0], bias[1], bias[2], bias[3]};
  if (EPI == 0 && ep.bias_b) {
#pragma unroll
    for (int m = 0; m < 4; ++m) bv[m] += __ldg(ep.bias_b + (long)it.b * n_total + it.n + 8 * m + (lane >> 2));
  }
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int cg = 0; cg < 2; ++cg)
#pragma unroll
      for (int e = 0; e < 2; ++e) v[frag_idx(m, cg, e)] = __uint_as_float(acc[frag_idx(m, cg, e)]) + bv[m];
  const uint32_t sbase = smem_u32(scratch);
#pragma unroll
  for (int i = 0; i < kMaxRes; ++i) {
    if (epi_has_res<EPI>(ep, i)) {
      if (ep.epi_smem) {
        // coalesced registers -> scratch rows [time][32 channels] -> fragment registers (ldmatrix .trans)
#pragma unroll
        for (int j = 0; j < 2; ++j)
          *reinterpret_cast<uint4*>(scratch + scrT_off(8 * j + (lane >> 2), lane & 3)) = ld.res[i][j];
        __syncwarp();
      }
#pragma unroll
      for (int cg = 0; cg < 2; ++cg) {
        uint32_t r[4];
        if (ep.epi_smem) {
          ldmatrix_x4_trans(sbase + scrT_off(cg * 8 + (lane & 7), lane >> 3), r);
        } else {
          r[0] = movmatrix_trans(ld.res[i][cg].x);
          r[1] = movmatrix_trans(ld.res[i][cg].y);
          r[2] = movmatrix_trans(ld.res[i][cg].z);
          r[3] = movmatrix_trans(ld.res[i][cg].w);
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const float2 a = unpack_act2<F16>(r[m]);
          // a-form -> residual stream: a >= 0 ? a : a * res_gain  ==  min(a, a * res_gain) for res_gain >= 1
          v[frag_idx(m, cg, 0)] += fminf(a.x, a.x * res_gain);
          v[frag_idx(m, cg, 1)] += fminf(a.y, a.y * res_gain);
        }
      }
      if (ep.epi_smem) __syncwarp();
    }
  }
}

template <int EPI, bool F16 = false>
__device__ __forceinline__ void epiT_store(const ConvEpilogue& ep, uint8_t* scratch, const EpiItem& it, int rowstride,
                                           int lane, float out_slope, float mrf_scale, float (&v)[kIW]) {
  if (EPI == 3 || (EPI == 0 && ep.mrf_mode == 3)) {
#pragma unroll
    for (int j = 0; j < kIW; ++j) v[j] *= mrf_scale;
  }
  const uint32_t sbase = smem_u32(scratch);
#pragma unroll
  for (int cg = 0; cg < 2; ++cg) {
    uint32_t pk[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const float x0 = v[frag_idx(m, cg, 0)], x1 = v[frag_idx(m, cg, 1)];
      pk[m] = pack_act2<F16>(fmaxf(x0, x0 * out_slope), fmaxf(x1, x1 * out_slope));  // slope in (0,1]
    }
    if (ep.epi_smem) {
      stmatrix_x4_trans(sbase + scrT_off(cg * 8 + (lane & 7), lane >> 3), pk[0], pk[1], pk[2], pk[3]);
    } else {
      // register transpose: afterwards this thread holds (time cg*8 + lane/4, channels 8m + 2(lane%4), +1)
      const int row = cg * 8 + (lane >> 2);
      __nv_bfloat16* p = ep.out + it.base + (long)row * rowstride + 2 * (lane & 3);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const uint32_t t = movmatrix_trans(pk[m]);
        if (row < it.rows_valid) *reinterpret_cast<uint32_t*>(p + 8 * m) = t;
      }
    }
  }
  if (!ep.epi_smem) return;
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int row = 8 * j + (lane >> 2);
    const uint4 ov = *reinterpret_cast<const uint4*>(scratch + scrT_off(row, lane & 3));
    if (row < it.rows_valid)
      *(reinterpret_cast<uint4*>(ep.out + it.base + (long)row * rowstride) + (lane & 3)) = ov;
  }
  __syncwarp();
}

// The same store through the TMA engine: the item is staged as above (the scratch layout IS the SWIZZLE_64B layout of a
// {32 channels, 16 rows} box: 16-byte chunk index XOR address bits [7, 9)) and one bulk tensor store writes it, clipping
// rows past the end of the utterance.  Against epiT_store this saves the LDS.128 read-back and the STG.128s, i.e. 0.7 K of
// the ~13 K L1 data-pipe wavefronts a 128 x 256 tile of a k = 11 layer costs (the pipe the tensor core's operand fetch
// and the TMEM loads share).  `buf` must not be touched again before bulk_wait_read<> says so.
template <int EPI, bool F16 = false>
__device__ __forceinline__ void epiT_store_tma(const CUtensorMap* tmO, uint8_t* buf, const EpiItem& it, int lane,
                                               float out_slope, float mrf_scale, float (&v)[kIW]) {
  if (EPI == 3) {
#pragma unroll
    for (int j = 0; j < kIW; ++j) v[j] *= mrf_scale;
  }
  const uint32_t sbase = smem_u32(buf);
#pragma unroll
  for (int cg = 0; cg < 2; ++cg) {
    uint32_t pk[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const float x0 = v[frag_idx(m, cg, 0)], x1 = v[frag_idx(m, cg, 1)];
      pk[m] = pack_act2<F16>(fmaxf(x0, x0 * out_slope), fmaxf(x1, x1 * out_slope));  // slope in (0,1]
    }
    stmatrix_x4_trans(sbase + scrT_off(cg * 8 + (lane & 7), lane >> 3), pk[0], pk[1], pk[2], pk[3]);
  }
  fence_proxy_async();   // generic-proxy stores -> visible to the TMA engine's async-proxy reads
  __syncwarp();
  if (lane == 0) {
    if (it.rows_valid > 0) tma_store_3d(tmO, buf, it.n, it.t, it.b);
    bulk_commit();
  }
}

}  // namespace vd
