#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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

// ---- 16-bit activation / weight storage: bf16 (default) or fp16 (ConvEpilogue::f16, option "fp16").
// F16 is a template parameter of the tcgen05 kernels (no per-element branch on the hot path); the *_rt forms take a
// run-time flag for pack kernels and the CUDA-core cross-check.
constexpr float kF16Max = 65504.f;
template <bool F16>
__device__ __forceinline__ uint32_t pack_act2(float lo, float hi) {
  if constexpr (F16) {
    const __half2 h = __floats2half2_rn(fminf(fmaxf(lo, -kF16Max), kF16Max), fminf(fmaxf(hi, -kF16Max), kF16Max));
    return *reinterpret_cast<const uint32_t*>(&h);
  } else {
    const __nv_bfloat162 o = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&o);
  }
}
template <bool F16>
__device__ __forceinline__ float2 unpack_act2(uint32_t u) {
  if constexpr (F16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
  else return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ __nv_bfloat16 pack_act_rt(float v, int f16) {
  if (f16) {
    const __half h = __float2half_rn(fminf(fmaxf(v, -kF16Max), kF16Max));
    return *reinterpret_cast<const __nv_bfloat16*>(&h);
  }
  return __float2bfloat16_rn(v);
}
__device__ __forceinline__ float unpack_act_rt(__nv_bfloat16 a, int f16) {
  return f16 ? __half2float(*reinterpret_cast<const __half*>(&a)) : __bfloat162float(a);
}

// scalar epilogue (SIMT path and tails); see ConvEpilogue for the semantics
__device__ __forceinline__ void epilogue_scalar(const ConvEpilogue& ep, int b, long row, int n, int n_total,
                                                float acc) {
  const long idx = row * n_total + n;  // row = b*L + t
  float v = acc + ep.bias[n];
  if (ep.bias_b) v += ep.bias_b[(long)b * n_total + n];
  for (int i = 0; i < ep.nres; ++i) {
    float a = unpack_act_rt(ep.res[i][idx], ep.f16);
    v += a >= 0.f ? a : a * ep.res_gain;
  }
  if (ep.mrf_mode == 1) {
    ep.mrf[idx] = v;
  } else if (ep.mrf_mode == 2) {
    ep.mrf[idx] += v;
  } else {
    if (ep.mrf_mode == 3) v = ((ep.mrf ? ep.mrf[idx] : 0.f) + v) * ep.mrf_scale;
    ep.out[idx] = pack_act_rt(lrelu(v, ep.out_slope), ep.f16);
  }
}

}  // namespace vd
