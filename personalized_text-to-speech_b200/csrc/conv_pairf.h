// Host-visible plan for the time-folded fused ResBlock1 pair kernel (conv_pairf.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "conv_tc.h"

namespace vd {

constexpr int kPfMaxTaps = 12;   // folded taps per conv (k <= 15, r >= 2)
constexpr int kPfMaxW = 8;       // weight ring stages

struct PairFParams {
  int B, L, Lf;          // utterances, samples per utterance, folded rows per utterance (L / r)
  int d;                 // sub-sequences of c1's view: its dilation on a folded view (C = 32 / 64), 1 on the plain one
  int dstep;             // rows between c1's taps: 1, or c1's dilation on the plain 128-channel view (r = 1)
  int xmin;              // first row of the staged x tile relative to the h tile's first row
  int WO, HR;            // output folded rows per tile (c2's N); rows of h per tile = WO + nt - 1
  int N1, XR;            // c1's N per sub-sequence; rows per staged x sub-tile
  int nt, smin;          // folded taps per conv and the first tap's row offset (both convs share k)
  uint32_t kmask[kPfMaxTaps];  // bit c: 64-channel K-chunk c of the folded tap is non-zero
  int rows_rho;          // rows per sub-sequence of c1's view = ceil(L / (d*r))
  int nw;                // weight ring stages
  int buf_bytes;         // one tile slot's shared-memory buffer: the x sub-tiles, then (in place) the h tile
  int m_tiles, total_tiles;
  FastDiv div_m, div_dr;
  const float* bias1;
  const float* bias2;
  float slope, res_gain;
  const __nv_bfloat16* x;
  __nv_bfloat16* out;
  unsigned long long* trace;  // debug (VITSDEC_TRACE=1 builds): per-tile clock64 stamps of CTA 0, [tile][12]
  int f16;               // 1: fp16 storage instead of bf16 (ConvEpilogue::f16)
};

struct PairFPlan {
  CUtensorMap tmX, tmW;
  PairFParams p;
  int channels;
  int grid;
  size_t smem;
};

// folded taps of a k-tap conv at fold factor r = 128 / channels
int pairf_taps(int channels, int k);
bool pairf_supported(int channels, int k, int dil);
bool pairf_preferred(int channels, int k, int dil);
// x: a-form input [B][L][C] (L a multiple of r = 128/C); w_fold: [2*nt][128][128] block-Toeplitz taps of c1, then c2
int plan_conv_pairf(PairFPlan* pl, int B, int L, int channels, int k, int dil, const __nv_bfloat16* x,
                    const __nv_bfloat16* w_fold, int num_sms);
int launch_conv_pairf(PairFPlan& pl, const float* bias1, const float* bias2, float slope, __nv_bfloat16* out,
                      cudaStream_t stream, int f16 = 0);

int encode_tmap_act(CUtensorMap* m, const void* base, uint32_t kc, uint64_t d1, uint64_t s1, uint64_t d2, uint64_t s2,
                    uint64_t rows, uint64_t srow, uint64_t B, uint64_t sb, uint32_t box_rows, uint32_t box_d2 = 1);

}  // namespace vd
