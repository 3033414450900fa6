// Host-visible plan for the fused "last ResBlock pair of every MRF branch + branch average" launch of a 128-channel
// stage (conv_mrf128.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "conv_tc.h"

namespace vd {

constexpr int kM8MaxBr = 3;   // branches per launch
constexpr int kM8MaxW = 8;    // weight ring stages

struct Mrf128Params {
  int B, L;                 // utterances, samples per utterance
  int nbr;                  // branches
  int WO;                   // output rows per tile (c2's N)
  int nw;                   // weight ring stages
  int xsub_bytes;           // one K-chunk of the staged x tile (largest branch)
  int xbuf_bytes;           // both K-chunks
  int nt[kM8MaxBr];         // taps of branch j's convs (k_j)
  int dstep[kM8MaxBr];      // rows between c1_j's taps (its dilation)
  int hk[kM8MaxBr];         // (k_j - 1) / 2: h row 0 is sample m0 - hk
  int xrow0[kM8MaxBr];      // first row of the x tile relative to m0: -hk - hk * dilation
  int HR[kM8MaxBr];         // rows of h per tile = WO + k_j - 1
  int N1[kM8MaxBr];         // c1_j's N (HR rounded up to 16)
  int XR[kM8MaxBr];         // rows of the staged x tile = xboxes * xbox_rows
  int xboxes[kM8MaxBr], xbox_rows[kM8MaxBr];   // TMA boxes of the x tile (a box has at most 256 rows)
  int wbase1[kM8MaxBr], wbase2[kM8MaxBr];      // first tap of c1_j / c2_j in the packed weights
  int m_tiles, total_tiles;
  FastDiv div_m;
  const float* bias1[kM8MaxBr];
  const float* bias2sum;    // sum of the branches' c2 biases
  const __nv_bfloat16* res[kM8MaxBr];   // the branch inputs again (residual rows, re-read from L2)
  float slope, res_gain, out_slope, scale;
  __nv_bfloat16* out;
};

struct Mrf128Maps {
  CUtensorMap x[kM8MaxBr];
};

struct Mrf128Plan {
  Mrf128Maps tm;
  CUtensorMap tmW;
  Mrf128Params p;
  int grid;
  size_t smem;
};

bool mrf128_supported(int channels, int nbr, const int* k, const int* dil);
// xs[j]: a-form input of branch j's pair [B][L][128]; w: packed taps [ntaps][128][128] in the order c1_0, c1_1, .., c2_0, ..
int plan_conv_mrf128(Mrf128Plan* pl, int B, int L, int nbr, const int* k, const int* dil, const __nv_bfloat16* const* xs,
                     const __nv_bfloat16* w, int num_sms);
int launch_conv_mrf128(Mrf128Plan& pl, const float* const* bias1, const float* bias2sum, float slope, float out_slope,
                       __nv_bfloat16* out, cudaStream_t stream, int f16 = 0);

}  // namespace vd
