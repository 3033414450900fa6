// Shared description of one fused convolution launch (channels-last activations).
//
// Activations are [B][L][C] bf16 ("rows" = time samples, channels contiguous) so that a filter tap is a
// row shift of the same tile.  Every convolution on the decoder path -- conv_pre (models.py:271), the
// polyphase form of the four ConvTranspose1d `ups` (models.py:277) and the 72 ResBlock convs
// (modules.py:210-223) -- is   y[b,t,n] = sum_tap sum_ci  W[tap][n][ci] * x[b, t + off[tap], ci]
// followed by a fused epilogue.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace vd {

constexpr int kMaxTaps = 32;
constexpr int kMaxSeg = 4;  // input tensors whose convolutions accumulate into one output (MRF fusion)

struct ConvGeom {
  int B;        // utterances
  int L;        // rows (time samples) per utterance, input == output rows
  int c_in;     // K per tap
  int n_total;  // output columns per row (C_out, or stride*C_out for a polyphase transposed conv)
  int ntaps;
  int tap_off[kMaxTaps];  // input-row offset of each tap
  int tap_nlo[kMaxTaps];  // tap contributes only to columns [nlo, nhi) (polyphase zero blocks)
  int tap_nhi[kMaxTaps];
  // bit kc set: K-chunk kc of the tap (64 channels, or 32 when c_in is not a multiple of 64) holds non-zero weights.
  // All ones for ordinary layers; time-folded layers (decoder.cu fold_geom) have structurally zero chunks.
  uint32_t tap_kmask[kMaxTaps];
  // Segments: taps [seg_tap_end[s-1], seg_tap_end[s]) read input tensor s.  One segment for an ordinary conv; the
  // last convs of the MRF branches (models.py:279-284) run as ONE launch with one segment per branch, so the branch
  // sum is formed in the accumulator instead of in HBM.
  int nseg;
  int seg_tap_end[kMaxSeg];
  // Dilated time-folded view (decoder.cu fold_geom, rho > 1): the tensor is C = c_real channels x L_real samples and
  // row n of sub-sequence rho holds samples t = rho_d*(r*n + phi) + rho, phi = 0..r-1 (r = c_in / c_real), so that a
  // dilation-rho_d convolution is a dilation-1 convolution on each sub-sequence.  L is then rows per sub-sequence,
  // ceil(L_real / (rho_d*r)).  0/1 = ordinary contiguous view.
  int rho_d;
  int c_real;
  int L_real;
};

// Epilogue:  v = acc + bias[n] (+ bias_b[b][n]) (+ sum_i unlrelu(res[i][b,t,n]))
//   mrf_mode 0: out = lrelu(v, out_slope)
//   mrf_mode 1: mrf  = v                       (first MRF branch)          no bf16 output
//   mrf_mode 2: mrf += v                       (middle MRF branches)       no bf16 output
//   mrf_mode 3: out = lrelu((mrf + v) * mrf_scale, out_slope)  (last branch; mrf may be null -- it is when the
//               branches were accumulated in TMEM by a multi-segment launch)
//   mrf_mode 4: conv_post on the time-folded view (channels-as-M tiles only): out_f32[b][r*n + phi] = tanh(acc) for
//               the folded output column phi*post_c + 0; no bias, no bf16 output            models.py:286-287
// Activations are stored post-leaky-relu ("a-form"): the next conv's tensor-core operand.  The
// residual stream x is recovered exactly (up to the bf16 rounding of a) as x = a >= 0 ? a : a * res_gain
// with res_gain = 1/slope, which is what lets one bf16 tensor serve as both operand and residual.
struct ConvEpilogue {
  const float* bias;           // [n_total]
  const float* bias_b;         // [B][n_total] or null
  const __nv_bfloat16* res[kMaxSeg];  // nres residual tensors [B][L][n_total] in a-form
  int nres;
  float res_gain;              // 1/slope of the a-form stored in `res`
  float* mrf;                  // [B][L][n_total] fp32 or null
  int mrf_mode;
  float mrf_scale;
  float out_slope;             // 1.0f = identity
  __nv_bfloat16* out;          // [B][L][n_total]
  const float* rowmask;        // optional [B][L]: every output row is multiplied by its mask value (x_mask of the
                               // flow, modules.py:171,176; time-as-M tiles with the generic epilogue only)
  int split_col;               // > 0 (generic time-as-M only): ONE launch, two epilogues.  Columns [0, split_col) take the
                               // residual res[0] and the bf16 store to `out` (both with split_col columns per row);
                               // columns [split_col, n_total) go to the fp32 accumulator `mrf` (n_total - split_col
                               // columns per row, mrf_mode 1 = store / 2 = add).  WN's res_skip_layers, modules.py:169-173
  int gate;                    // 1: columns are (a_j, b_j) pairs; out[b,t,j] = tanh(a_j) * sigmoid(b_j), n_total/2 output
                               // columns (WN's fused_add_tanh_sigmoid_multiply, commons.py:103-110; generic time-as-M only)
  int epi_smem;                // channels-as-M epilogue: 1 = transpose through shared memory (ldmatrix/stmatrix; default),
                               // 0 = register transposes (movmatrix) and 4-byte global accesses (slower, kept as a knob)
  float* out_f32;              // mrf_mode 4: waveform [B][L * n_total / post_c]
  int post_c;                  // mrf_mode 4: real channels per time sample (n_total = r * post_c)
  int f16;                     // 1: operands, residuals and the stored output are IEEE fp16 instead of bf16 (option
                               // "fp16"): same tensor-core rate (kind::f16), 3 more mantissa bits per stored activation;
                               // stores saturate at +-65504.  Pointers stay typed __nv_bfloat16 (opaque 16-bit storage).
};

}  // namespace vd
