// Host-visible plan for the fused "last pair of every MRF branch" kernel (conv_mrfp.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "conv_tc.h"

namespace vd {

constexpr int kMpMaxBr = 3;    // MRF branches (resblock kernels) per stage
constexpr int kMpMaxNA = 6;    // activation stages in flight

struct MrfpParams {
  int B, L;
  int nbr;                      // branches
  int k[kMpMaxBr], dil[kMpMaxBr], hk[kMpMaxBr];   // taps, dilation of c1, (k-1)/2 per branch (c2 has dilation 1)
  int hmax;                     // max hk: every branch's h tile covers times [t0 - hmax, t0 - hmax + 256)
  int bmo;                      // valid output rows per tile = 256 - 2*hmax
  int nboxes[kMpMaxBr];         // 64-row TMA boxes of branch j's activation tile (256 + 2*hk*dil rows)
  int a_lo[kMpMaxBr];           // first row of that tile relative to t0: -(hmax + hk*dil)
  int w1_tap[kMpMaxBr];         // first tap of c1_j / c2_j inside the packed weight set
  int w2_tap[kMpMaxBr];
  int ntaps;                    // total taps in the packed set (2 * sum k)
  int a_stage_bytes, na_stages;
  int m_tiles, total_tiles;
  FastDiv div_m;
  const float* bias1[kMpMaxBr]; // c1 biases
  const float* bias2sum;        // sum of the branches' c2 biases
  float slope;                  // leaky-relu inside the ResBlock (h, and the a-form the inputs are stored in)
  float res_gain;               // 1 / slope
  float out_slope;              // leaky-relu applied to the stage output (0.1, or 0.01 before conv_post)
  float scale;                  // 1 / nbr
  __nv_bfloat16* out;
  int f16;
};

struct MrfpPlan {
  CUtensorMap tmA[kMpMaxBr];
  CUtensorMap tmW;
  MrfpParams p;
  int channels;
  int grid;
  size_t smem;
};

// true when the three (or fewer) branches' last pairs fit the kernel: C = 32, both weight sets of every branch resident
bool mrfp_supported(int channels, int nbr, const int* k, const int* dil);
// xs[j]: a-form input of branch j's last pair [B][L][C]; w: packed [ntaps][C][C] in the order c1_0, c1_1, .., c2_0, ..
int plan_conv_mrfp(MrfpPlan* pl, int B, int L, int channels, int nbr, const int* k, const int* dil,
                   const __nv_bfloat16* const* xs, const __nv_bfloat16* w, int num_sms);
int launch_conv_mrfp(MrfpPlan& pl, const float* const* bias1, const float* bias2sum, float slope, float out_slope,
                     __nv_bfloat16* out, cudaStream_t stream, int f16 = 0);

}  // namespace vd
