// Host-visible plan for the fused ResBlock1 pair kernel (conv_pair.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "conv_tc.h"

namespace vd {

struct PairParams {
  int B, L;
  int k, dil, hk;        // taps per conv, dilation of the first conv, (k-1)/2
  int bmo;               // valid output rows per tile = 256 - (k-1)
  int nboxes;            // 64-row TMA boxes per activation stage (256 + (k-1)*dil rows)
  int a_stage_bytes;
  int na_stages;
  int nh;                // buffers for the intermediate h (2 when shared memory allows)
  int m_tiles, total_tiles;
  FastDiv div_m;
  const float* bias1;
  const float* bias2;
  float slope, res_gain;
  __nv_bfloat16* out;
  unsigned long long* trace;
  int f16;               // 1: fp16 storage instead of bf16 (ConvEpilogue::f16)
};

struct PairPlan {
  CUtensorMap tmA, tmW;
  PairParams p;
  int channels;
  int nacc;         // 128-row accumulators per conv per tile (2, or 1 when shared memory is tight)
  int grid;
  size_t smem;
  bool pdl;         // see ConvTcPlan::pdl
};

// true when both convs' weights, three activation stages and the h tile fit in shared memory
bool pair_supported(int channels, int k, int dil);
// x: a-form input [B][L][C]; w_pair: packed [2k][C][C] (taps of c1, then taps of c2)
int plan_conv_pair(PairPlan* pl, int B, int L, int channels, int k, int dil, const __nv_bfloat16* x,
                   const __nv_bfloat16* w_pair, int num_sms);
int launch_conv_pair(PairPlan& pl, const float* bias1, const float* bias2, float slope, __nv_bfloat16* out,
                     cudaStream_t stream, int f16 = 0);

}  // namespace vd
