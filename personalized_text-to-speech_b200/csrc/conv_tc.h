// Host-visible plan for one tcgen05 convolution launch (see conv_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "conv_params.h"

namespace vd {

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
  uint32_t tap_delta16[kMaxTaps];  // (tap_off - seg_halo_lo) * row_bytes >> 4: descriptor start-address delta per tap
  int res_prefetch;           // 1: the producer prefetches the residual tiles (tmR) into L2
  int desc_mode;              // debug knob for the A descriptor base-offset field (0 = none)
};

struct TmapPack {
  CUtensorMap a[kMaxSeg];     // activation tensors, one per segment
  CUtensorMap r[kMaxSeg];     // residual tensors (L2 prefetch only)
};

struct ConvTcPlan {
  TmapPack tm;
  CUtensorMap tmW;
  const void* res_bound[kMaxSeg];
  ConvTcParams p;
  int bn, kc;
  int grid;
  size_t smem;
};

// Encodes the TMA descriptors for the activation tensors xs[0..g.nseg) (each [B][L][c_in]) and the packed weights
// `w` [ntaps][n_total][c_in].
int plan_conv_tc(ConvTcPlan* pl, const ConvGeom& g, const __nv_bfloat16* const* xs, const __nv_bfloat16* w,
                 int num_sms, int desc_mode);
int bind_residual_tc(ConvTcPlan& pl, const ConvEpilogue& ep);
int launch_conv_tc(ConvTcPlan& pl, const ConvEpilogue& ep, cudaStream_t stream);
int launch_conv_simt(const ConvGeom& g, const ConvEpilogue& ep, const __nv_bfloat16* const* xs,
                     const __nv_bfloat16* w, cudaStream_t stream);

}  // namespace vd
