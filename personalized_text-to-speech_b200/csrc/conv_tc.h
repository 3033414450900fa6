// Host-visible plan for one tcgen05 convolution launch (see conv_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "conv_params.h"

namespace vd {

struct ConvTcParams {
  ConvGeom g;
  ConvEpilogue ep;
  int halo_lo;        // min tap offset
  int nboxes;         // 64-row TMA boxes per A stage
  int a_stage_bytes;  // nboxes * 64 * KC * 2
  int na_stages;      // activation stages in flight
  int nb_stages;      // streamed weight stages (0 when stationary)
  int b_region_bytes; // bytes of the weight region (ring or resident set)
  int stationary;     // 1: the layer's whole weight set stays resident in shared memory
  int m_tiles;        // per utterance
  int n_tiles;
  int total_tiles;
  uint32_t tap_delta16[kMaxTaps];  // (tap_off - halo_lo) * row_bytes >> 4: descriptor start-address delta per tap
  int res_prefetch;   // 1: the producer prefetches the residual tile (tmR) into L2
  int desc_mode;      // debug knob for the A descriptor base-offset field (0 = none)
};

struct ConvTcPlan {
  CUtensorMap tmA, tmW, tmR;
  const void* res_bound;
  ConvTcParams p;
  int bn, kc;
  int grid;
  size_t smem;
};

// Encodes the TMA descriptors for activations `x` [B][L][c_in] and packed weights `w` [ntaps][n_total][c_in].
int plan_conv_tc(ConvTcPlan* pl, const ConvGeom& g, const __nv_bfloat16* x, const __nv_bfloat16* w, int num_sms,
                 int desc_mode);
int bind_residual_tc(ConvTcPlan& pl, const __nv_bfloat16* res);
int launch_conv_tc(ConvTcPlan& pl, const ConvEpilogue& ep, cudaStream_t stream);
int launch_conv_simt(const ConvGeom& g, const ConvEpilogue& ep, const __nv_bfloat16* x, const __nv_bfloat16* w,
                     cudaStream_t stream);

}  // namespace vd
