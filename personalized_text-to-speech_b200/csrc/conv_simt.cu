// CUDA-core direct convolution over the same packed operands / epilogue as the tcgen05 kernel.
// Role: GPU-side cross-check for the tensor-core path in tests (VITSDEC impl "simt"), selected
// explicitly -- it is never a silent fallback.
#include "common.cuh"

namespace vd {

constexpr int kSimtRows = 4;

struct SimtInputs {
  const __nv_bfloat16* x[kMaxSeg];
};

__global__ void __launch_bounds__(128) conv_simt_kernel(ConvGeom g, ConvEpilogue ep, SimtInputs in,
                                                        const __nv_bfloat16* __restrict__ w) {
  const int n = blockIdx.y * 32 + (threadIdx.x & 31);
  const int tgroup = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  const int t0 = tgroup * kSimtRows;
  if (n >= g.n_total || t0 >= g.L) return;
  float acc[kSimtRows];
#pragma unroll
  for (int r = 0; r < kSimtRows; ++r) acc[r] = 0.f;
  int seg = 0;
  for (int tap = 0; tap < g.ntaps; ++tap) {
    while (tap >= g.seg_tap_end[seg]) ++seg;
    const __nv_bfloat16* __restrict__ x = in.x[seg];
    if (n < g.tap_nlo[tap] || n >= g.tap_nhi[tap]) continue;
    const uint4* wr = reinterpret_cast<const uint4*>(w + ((long)tap * g.n_total + n) * g.c_in);
    for (int c8 = 0; c8 < g.c_in / 8; ++c8) {
      const uint4 wv = __ldg(wr + c8);
      const uint32_t* w2 = reinterpret_cast<const uint32_t*>(&wv);
#pragma unroll
      for (int r = 0; r < kSimtRows; ++r) {
        const int ti = t0 + r + g.tap_off[tap];
        if (ti < 0 || ti >= g.L) continue;
        const uint4 xv = __ldg(reinterpret_cast<const uint4*>(x + ((long)b * g.L + ti) * g.c_in) + c8);
        const uint32_t* x2 = reinterpret_cast<const uint32_t*>(&xv);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 wf = ep.f16 ? unpack_act2<true>(w2[q]) : unpack_act2<false>(w2[q]);
          const float2 xf = ep.f16 ? unpack_act2<true>(x2[q]) : unpack_act2<false>(x2[q]);
          acc[r] = fmaf(wf.x, xf.x, acc[r]);
          acc[r] = fmaf(wf.y, xf.y, acc[r]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kSimtRows; ++r) {
    const int t = t0 + r;
    if (t < g.L) epilogue_scalar(ep, b, (long)b * g.L + t, n, g.n_total, acc[r]);
  }
}

int launch_conv_simt(const ConvGeom& g, const ConvEpilogue& ep, const __nv_bfloat16* const* xs, const __nv_bfloat16* w,
                     cudaStream_t stream) {
  SimtInputs in{};
  for (int s = 0; s < g.nseg; ++s) in.x[s] = xs[s];
  VD_CHECK(g.c_in % 8 == 0, "conv_simt: c_in must be a multiple of 8");
  dim3 grid((g.L + 4 * kSimtRows - 1) / (4 * kSimtRows), (g.n_total + 31) / 32, g.B);
  conv_simt_kernel<<<grid, 128, 0, stream>>>(g, ep, in, w);
  VD_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vd
