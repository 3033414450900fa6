// Host-visible plan for the fused ResBlock1 pair / "last pair of every MRF branch" kernel of the narrow stages
// (conv_mrfp.cu): 64 (virtual) channels per row -- C = 32 on the 2-sample time-folded view, C = 64 on the plain one.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "conv_tc.h"

namespace vd {

constexpr int kMpMaxBr = 3;     // branches per launch (1 = a plain ResBlock pair, 3 = the MRF of the shipped config)
constexpr int kMpMaxNA = 4;     // activation stages in flight
constexpr int kMpMaxJobs = 160; // tensor-core jobs of all convs of a launch

// One tensor-core job = NACC x 2 MMAs (K = 32): D[:, d_off .. d_off + N) += A[rows + a_off, K slice] * W_block^T
struct MpJob {
  uint16_t a_off16;   // (row offset * 128 + input phase * 64) >> 4, relative to the tile (A) or the h buffer
  uint16_t w_off16;   // byte offset of the weight block in shared memory >> 4
  uint8_t d_off;      // first accumulator column: 0, or 32 for an output-phase-1 block
  uint8_t flags;      // bit 0: N = 64 (both output phases, two stacked tap blocks), else N = 32; bit 1: first write of its columns
  uint16_t pad;
};

struct MrfpParams {
  int B, Lf;                    // utterances; folded rows per utterance (L / 2)
  int nbr;                      // branches
  int hm;                       // h halo in folded rows: every branch's h tile covers rows [m0 - hm, m0 - hm + 128*NACC)
  int bmo;                      // valid output folded rows per tile = 128*NACC - 2*hm
  int a_lo[kMpMaxBr];           // first folded row of branch j's activation tile relative to m0
  int a_boxes[kMpMaxBr];        // 32-row TMA boxes of that tile
  int job_beg[2 * kMpMaxBr + 1];// jobs of conv c = [job_beg[c], job_beg[c+1]); conv order c1_0, c2_0, c1_1, c2_1, ...
  int conv_base[2 * kMpMaxBr];  // first tap of conv c in the packed weights (global order: all c1, then all c2)
  int conv_k[2 * kMpMaxBr];
  int ntaps;                    // taps of all convs of the launch
  int fold;                     // samples per row: 2 (C = 32, two phases side by side) or 1 (C = 64, plain rows)
  int wbytes;                   // bytes of all weight blocks in shared memory
  int a_stage_bytes, na_stages, nh;
  int res_smem;                 // 1: the residual rows come from the resident activation tiles (which then stay until the
                                // output epilogue has read them: needs na_stages > nbr); 0: re-read from global memory / L2
  int m_tiles, total_tiles;
  FastDiv div_m;
  const float* bias1[kMpMaxBr]; // c1 biases
  const float* bias2sum;        // sum of the branches' c2 biases
  const __nv_bfloat16* res[kMpMaxBr];  // the branch inputs again, for res_smem = 0
  float slope;                  // leaky-relu inside the ResBlock (h, and the a-form the inputs are stored in)
  float res_gain;               // 1 / slope
  float out_slope;              // leaky-relu of the output (the ResBlock slope, or the next stage's / conv_post's)
  float scale;                  // 1 / nbr
  __nv_bfloat16* out;
  int f16;
  unsigned long long* trace;    // debug (VITSDEC_TRACE=1 builds): per-(tile, branch) clock64 stamps of CTA 0, [n][12]
  MpJob jobs[kMpMaxJobs];
};

struct MrfpPlan {
  CUtensorMap tmA[kMpMaxBr];
  CUtensorMap tmW;
  MrfpParams p;
  int channels;
  int nacc;         // 128-row accumulators per conv per tile (2; 1 for small problems)
  int grid;
  size_t smem;
  bool pdl;
};

// true when the launch fits: C = 32 (even length) or C = 64, both weight sets of every branch resident next to the tiles
bool mrfp_supported(int channels, int nbr, const int* k, const int* dil);
// xs[j]: a-form input of branch j's pair [B][L][C], L even; w: packed taps [ntaps][C][C] in the order c1_0, c1_1, .., c2_0, ..
int plan_conv_mrfp(MrfpPlan* pl, int B, int L, int channels, int nbr, const int* k, const int* dil,
                   const __nv_bfloat16* const* xs, const __nv_bfloat16* w, int num_sms);
int launch_conv_mrfp(MrfpPlan& pl, const float* const* bias1, const float* bias2sum, float slope, float out_slope,
                     __nv_bfloat16* out, cudaStream_t stream, int f16 = 0);

}  // namespace vd
