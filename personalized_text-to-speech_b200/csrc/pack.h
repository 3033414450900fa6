// Launchers for the load-time / boundary kernels in pack.cu.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>

namespace vd {
int launch_wn_scale(const float* v, const float* g, float* scale, int rows, int inner, cudaStream_t st);
int launch_pack_conv(const float* w, const float* scale, __nv_bfloat16* wp, int c_out, int c_in, int k,
                     cudaStream_t st, int interleave = 0, int f16 = 0, int c_in_src = 0, int c_out_src = 0);
int launch_interleave_bias(const float* b, float* out, int c_out, cudaStream_t st);
int launch_pack_conv_fold(const float* w, const float* scale, __nv_bfloat16* wp, int C, int c_out, int k, int r,
                          cudaStream_t st, int lo_part = 0, int f16 = 0, int c_in_src = 0);
int launch_pack_convT(const float* w, const float* scale, __nv_bfloat16* wp, int c_in, int c_out, int k, int s, int p,
                      int ntaps, int off0, cudaStream_t st, int f16 = 0, int c_in_src = 0, int c_out_src = 0);
int launch_replicate_bias(const float* b, float* out, int c_out, int reps, cudaStream_t st, int c_out_src = 0);
int launch_pcm16(const float* wav, int16_t* out, long n, cudaStream_t st);
int launch_sum_bias(const float* b0, const float* b1, const float* b2, const float* b3, float* out, int n,
                    cudaStream_t st);
int launch_pack_z(const float* z, long sb, long sc, __nv_bfloat16* out, int B, int C, int T, cudaStream_t st,
                  int f16 = 0, int c_pad = 0);
int launch_cond(const float* wc, const float* bc, const float* g, float* cb, int B, int c_out, int gin,
                cudaStream_t st);
int launch_conv_post(const __nv_bfloat16* x, const float* w, float* out, int B, int L, int C, cudaStream_t st,
                     int f16 = 0);
int launch_unpack_debug(const __nv_bfloat16* a, float gain, float* out, int B, int L, int C, cudaStream_t st,
                        int f16 = 0);
// c_in_src / c_out_src (0 = same as the packed size): dimensions of the SOURCE weight / bias when the packed operand is
// zero-padded to the 32-channel granularity of the tensor-core tiles.
// The f16 argument of the packers selects the 16-bit storage format: 0 = bf16, 1 = IEEE fp16 (option "fp16").
}  // namespace vd
