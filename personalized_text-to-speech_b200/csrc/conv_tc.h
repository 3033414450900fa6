// Host-visible plan for one tcgen05 convolution launch (see conv_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "conv_params.h"

namespace vd {

// Division by a launch-time constant without the ~150-cycle integer-divide sequence (Granlund-Montgomery round-up
// method; exact for dividends below 2^31, which tile indices always are).
struct FastDiv {
  uint32_t mul, shr, div;
  void init(uint32_t d) {
    div = d;
    shr = 0;
    while ((1u << shr) < d) ++shr;
    mul = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << shr) - d)) / d) + 1;
  }
#ifdef __CUDACC__
  __device__ __forceinline__ uint32_t quot(uint32_t n) const { return (__umulhi(mul, n) + n) >> shr; }
  __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
    q = quot(n);
    r = n - q * div;
  }
#endif
};

struct ConvTcParams {
  ConvGeom g;
  ConvEpilogue ep;
  int seg_halo_lo[kMaxSeg];   // min tap offset of the segment
  int seg_nboxes[kMaxSeg];    // 64-row TMA boxes that cover BM + halo rows of the segment
  int a_stage_bytes;          // max over segments of nboxes * 64 * KC * 2
  int na_stages;              // activation stages in flight
  int nb_stages;              // streamed weight stages (0 when stationary)
  int b_region_bytes;         // bytes of the weight region (ring or resident set)
  int stationary;             // 1: the layer's whole weight set stays resident in shared memory
  int m_tiles;                // per utterance
  int n_tiles;
  int total_tiles;
  FastDiv div_n, div_m;       // tile -> (n-tile, m-block) -> (utterance x sub-sequence, m-tile)
  FastDiv div_rho;            // (utterance x sub-sequence) -> (utterance, rho)
  FastDiv div_dr;             // division by rho_d * r (rows of a dilated folded view)
  int swap_rows;              // time rows per channels-as-M tile (256, or 128 / 64 for small decodes)
  int rho_d;                  // sub-sequences per utterance (1 = ordinary view), see ConvGeom
  int c_shift;                // log2(channels per time sample) of a dilated folded view (5 otherwise)
  int r_fold;                 // time samples per folded row
  int L_real;                 // time samples per utterance
  long bstride;               // elements between utterances in the output / residual tensors
  int rowstride;              // elements between consecutive rows of one lane group (channels-as-M epilogue)
  uint16_t w_slot[kMaxTaps];  // resident weights: shared-memory slot of the tap's first non-zero K-chunk
  uint32_t tap_delta16[kMaxTaps];  // (tap_off - seg_halo_lo) * row_bytes >> 4: descriptor start-address delta per tap
  int res_prefetch;           // 1: the producer prefetches the residual tiles (tmR) into L2
  int cta2_relay;             // conv_tc2.cu: 1 = the first full-barrier protocol (peer relay thread), experiment knob
  int tma_epi;                // 1: the channels-as-M epilogue writes its output items with TMA tensor stores (tm.o)
  unsigned long long* trace;  // debug: per-tile clock64 stamps of CTA 0 ([tile][8]) or null
  int desc_mode;              // debug knob for the A descriptor base-offset field (0 = none)
};

struct TmapPack {
  CUtensorMap a[kMaxSeg];     // activation tensors, one per segment
  CUtensorMap r[kMaxSeg];     // residual tensors (L2 prefetch only)
  CUtensorMap o;              // output tensor, {32 channels, 16 rows} boxes (channels-as-M epilogue, tma_epi)
};

struct ConvTcPlan {
  TmapPack tm;
  CUtensorMap tmW;
  const void* res_bound[kMaxSeg];
  const void* out_bound;     // output tensor tm.o was encoded for
  bool no_tma_epi;           // experiment knob (desc_mode bit 11)
  ConvTcParams p;
  int bn, kc;
  int grid;
  bool no_res_prefetch;
  bool epi_smem;    // channels-as-M epilogue transposes through shared memory instead of registers (experiment knob)
  bool swap;        // channels-as-M variant (128 channels x 256 time rows per tile)
  bool cta2;        // paired tiles: 256 channels x 256 time rows per 2-CTA cluster, tcgen05.mma.cta_group::2 (conv_tc2.cu)
  bool pdl;         // launch with the programmatic-serialization attribute (set by the decoder for launches that leave
                    // SMs idle: the kernel's prologue then overlaps the previous launch of the stream)
  size_t smem;
};

// Encodes the TMA descriptors for the activation tensors xs[0..g.nseg) (each [B][L][c_in]) and the packed weights
// `w` [ntaps][n_total][c_in].
int plan_conv_tc(ConvTcPlan* pl, const ConvGeom& g, const __nv_bfloat16* const* xs, const __nv_bfloat16* w,
                 int num_sms, int desc_mode, bool allow_swap = true);
int bind_residual_tc(ConvTcPlan& pl, const ConvEpilogue& ep);
int launch_conv_tc(ConvTcPlan& pl, const ConvEpilogue& ep, cudaStream_t stream);
int launch_conv_tc2(const ConvTcPlan& pl, cudaStream_t stream);   // conv_tc2.cu; pl.p.ep already bound
int max_clusters_tc2(size_t smem);                                // resident 2-CTA clusters of conv_tc2_kernel (0: none)
int launch_conv_simt(const ConvGeom& g, const ConvEpilogue& ep, const __nv_bfloat16* const* xs,
                     const __nv_bfloat16* w, cudaStream_t stream);

}  // namespace vd
