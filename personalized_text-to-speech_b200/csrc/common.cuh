#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "conv_params.h"

namespace vd {

// ---- host error plumbing: every C-ABI entry returns 0 / non-zero and records a thread-local message
void set_error(const std::string& msg);
#define VD_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      vd::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                       \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)
#define VD_CHECK(cond, msg)        \
  do {                             \
    if (!(cond)) {                 \
      vd::set_error(msg);          \
      return 1;                    \
    }                              \
  } while (0)

__device__ __forceinline__ float lrelu(float v, float slope) { return v >= 0.f ? v : v * slope; }

// scalar epilogue (SIMT path and tails); see ConvEpilogue for the semantics
__device__ __forceinline__ void epilogue_scalar(const ConvEpilogue& ep, int b, long row, int n, int n_total,
                                                float acc) {
  const long idx = row * n_total + n;  // row = b*L + t
  float v = acc + ep.bias[n];
  if (ep.bias_b) v += ep.bias_b[(long)b * n_total + n];
  for (int i = 0; i < ep.nres; ++i) {
    float a = __bfloat162float(ep.res[i][idx]);
    v += a >= 0.f ? a : a * ep.res_gain;
  }
  if (ep.mrf_mode == 1) {
    ep.mrf[idx] = v;
  } else if (ep.mrf_mode == 2) {
    ep.mrf[idx] += v;
  } else {
    if (ep.mrf_mode == 3) v = ((ep.mrf ? ep.mrf[idx] : 0.f) + v) * ep.mrf_scale;
    ep.out[idx] = __float2bfloat16_rn(lrelu(v, ep.out_slope));
  }
}

}  // namespace vd
