// Load-time and boundary kernels: weight-norm fold + operand packing, latent transposition,
// speaker-conditioning GEMV, conv_post + tanh, debug read-back.
#include "common.cuh"
#include "pack.h"

namespace vd {

// scale[r] = g[r] / ||v[r, :]||_2  (torch.nn.utils.weight_norm, dim=0; reference models.py:254, modules.py:191-206)
__global__ void wn_scale_kernel(const float* __restrict__ v, const float* __restrict__ g, float* __restrict__ scale,
                                int inner) {
  const int r = blockIdx.x;
  float s = 0.f;
  if (g != nullptr) {
    for (int i = threadIdx.x; i < inner; i += blockDim.x) {
      const float x = v[(long)r * inner + i];
      s = fmaf(x, x, s);
    }
  }
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (blockDim.x + 31) / 32; ++i) t += red[i];
    // an all-zero row (pruned / dead channel of a remove_weight_norm checkpoint re-expressed as v = w, g = ||w|| = 0):
    // 0 * v = 0 like the reference's folded weight, not 0/0 = NaN
    scale[r] = g != nullptr ? (t > 0.f ? g[r] / sqrtf(t) : 0.f) : 1.f;
  }
}

// Conv1d weight [co][ci][k] -> packed [tap][co][ci] bf16 (B operand rows = co, K = ci contiguous).
// interleave != 0: packed row 2j holds weight row j, packed row 2j+1 weight row c_out/2 + j (gate pairs, ConvEpilogue::gate)
__global__ void pack_conv_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                 __nv_bfloat16* __restrict__ wp, int c_out, int c_in, int k, int interleave, int f16,
                                 int c_in_src, int c_out_src) {
  // c_in_src < c_in / c_out_src < c_out: the source weight is narrower than the packed operand, which is zero-padded
  // (conv_pre of a latent whose width is not a multiple of 32; stages narrower than 32 channels)
  const long total = (long)k * c_out * c_in;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ci = i % c_in;
    int co = (i / c_in) % c_out;
    if (interleave) co = (co & 1) ? c_out / 2 + (co >> 1) : (co >> 1);
    const int j = i / ((long)c_in * c_out);
    wp[i] = pack_act_rt(ci < c_in_src && co < c_out_src ? w[((long)co * c_in_src + ci) * k + j] * scale[co] : 0.f, f16);
  }
}

// Conv1d weight [co][ci][k] (C x C, dilation 1) -> time-folded block-Toeplitz operand [s][phi*C+co][psi*C+ci]
// (decoder.cu fold_geom; output channels c_out..C-1 are zero padding): folded row n holds time samples r*n..r*n+r-1; tap j contributes where
// j - (k-1)/2 = r*(s_min+s) + psi - phi.
__global__ void pack_conv_fold_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                      __nv_bfloat16* __restrict__ wp, int C, int c_out, int k, int r, int s_min,
                                      int ntaps, int lo_part, int f16, int c_in_src) {
  const int rc = r * C;
  const long total = (long)ntaps * rc * rc;
  const int hk = (k - 1) / 2;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int cc = i % rc;
    const int nn = (i / rc) % rc;
    const int s = s_min + (int)(i / ((long)rc * rc));
    const int psi = cc / C, ci = cc % C, phi = nn / C, co = nn % C;
    const int j = r * s + psi - phi + hk;
    float val = 0.f;
    // lo_part (two-term weights of a single-output layer): output row 0 holds bf16(w), row 1 the bf16 remainder
    // w - bf16(w) of the SAME source row; the epilogue adds the two accumulator rows
    const int src = lo_part ? 0 : co;
    if (j >= 0 && j < k && co < (lo_part ? 2 : c_out) && ci < c_in_src)
      val = w[((long)src * c_in_src + ci) * k + j] * (scale ? scale[src] : 1.f);
    if (lo_part && co == 1) val -= unpack_act_rt(pack_act_rt(val, f16), f16);
    wp[i] = pack_act_rt(val, f16);
  }
}

// ConvTranspose1d weight [ci][co][k] (stride s, padding p) -> polyphase packed [tap][r*c_out+co][ci]:
// output sample s*i + r takes input rows i + off; the contributing kernel index is j = r + p - s*off.
__global__ void pack_convT_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                  __nv_bfloat16* __restrict__ wp, int c_in, int c_out, int k, int s, int p,
                                  int ntaps, int off0, int f16, int c_in_src, int c_out_src) {
  const long total = (long)ntaps * s * c_out * c_in;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ci = i % c_in;
    const long rest = i / c_in;
    const int n = rest % (s * c_out);
    const int tap = rest / (s * c_out);
    const int r = n / c_out, co = n % c_out;
    const int j = r + p - s * (off0 + tap);
    float val = 0.f;
    if (j >= 0 && j < k && ci < c_in_src && co < c_out_src) val = w[((long)ci * c_out_src + co) * k + j] * scale[ci];
    wp[i] = pack_act_rt(val, f16);
  }
}

__global__ void replicate_bias_kernel(const float* __restrict__ b, float* __restrict__ out, int c_out, int reps,
                                      int c_out_src) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c_out * reps) out[i] = (b && i % c_out < c_out_src) ? b[i % c_out] : 0.f;   // padded channels: zero bias
}
__global__ void interleave_bias_kernel(const float* __restrict__ b, float* __restrict__ out, int c_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c_out) out[i] = b[(i & 1) ? c_out / 2 + (i >> 1) : (i >> 1)];
}

// combined bias of a fused MRF launch: sum of the member layers' biases
__global__ void sum_bias_kernel(const float* __restrict__ b0, const float* __restrict__ b1,
                                const float* __restrict__ b2, const float* __restrict__ b3, float* __restrict__ out,
                                int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = b0[i];
  if (b1) s += b1[i];
  if (b2) s += b2[i];
  if (b3) s += b3[i];
  out[i] = s;
}

// z fp32 [B][C][T] (strided) -> bf16 [B][T][Cp], channels [C, Cp) zero
__global__ void pack_z_kernel(const float* __restrict__ z, long sb, long sc, __nv_bfloat16* __restrict__ out, int C,
                              int T, int f16, int Cp) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < T) ? z[b * sb + c * sc + t] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T && c < Cp) out[((long)b * T + t) * Cp + c] = pack_act_rt(tile[threadIdx.x][i], f16);
  }
}

// cb[b][co] = cond.bias[co] + sum_ci cond.weight[co][ci] * g[b][ci]   (models.py:272-273; one warp per output)
__global__ void cond_kernel(const float* __restrict__ wc, const float* __restrict__ bc, const float* __restrict__ g,
                            float* __restrict__ cb, int c_out, int gin) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  if (warp >= c_out) return;
  float s = 0.f;
  for (int i = lane; i < gin; i += 32) s = fmaf(wc[(long)warp * gin + i], g[(long)b * gin + i], s);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) cb[(long)b * c_out + warp] = s + bc[warp];
}

// out[b][t] = tanh( sum_j sum_c w[c][j] * x[b][t + j - 3][c] ), x already leaky-relu'ed (models.py:285-287).
// M = 1 output channel: CUDA cores.  A block stages 512 + 6 rows in shared memory; each thread produces 4 outputs
// 128 rows apart so that (a) every weight vector read from shared memory is used 4 times (the first version issued
// one broadcast LDS per MAC and was LSU-bound at 168 us) and (b) a warp's row reads stay bank-conflict free.
constexpr int kPostThreads = 128;
template <int kPostPer>
__global__ void __launch_bounds__(kPostThreads) conv_post_kernel(const __nv_bfloat16* __restrict__ x,
                                                                const float* __restrict__ w, float* __restrict__ out,
                                                                int L, int C, int f16) {
  constexpr int kPostTile = kPostThreads * kPostPer;
  extern __shared__ uint8_t sm[];
  float* ws = reinterpret_cast<float*>(sm);                                  // [7][C]
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(sm + 7 * C * 4);     // [kPostTile + 6][C + 8] (padded rows)
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kPostTile;
  const int pitch = C + 8;
  for (int i = threadIdx.x; i < 7 * C; i += kPostThreads) ws[i] = w[(i % C) * 7 + i / C];  // conv_post.weight [1][C][7]
  const int vec_per_row = C / 8;
  for (int i = threadIdx.x; i < (kPostTile + 6) * vec_per_row; i += kPostThreads) {
    const int r = i / vec_per_row, v = i % vec_per_row;
    const int t = t0 + r - 3;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (t >= 0 && t < L) val = __ldg(reinterpret_cast<const uint4*>(x + ((long)b * L + t) * C) + v);
    *reinterpret_cast<uint4*>(xs + r * pitch + v * 8) = val;
  }
  __syncthreads();
  float acc[kPostPer];
#pragma unroll
  for (int o = 0; o < kPostPer; ++o) acc[o] = 0.f;
  for (int j = 0; j < 7; ++j) {
    for (int c = 0; c < C; c += 8) {
      const float4 w0 = *reinterpret_cast<const float4*>(ws + j * C + c);
      const float4 w1 = *reinterpret_cast<const float4*>(ws + j * C + c + 4);
#pragma unroll
      for (int o = 0; o < kPostPer; ++o) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xs + (threadIdx.x + o * kPostThreads + j) * pitch + c);
        float2 f0, f1, f2, f3;
        if (f16) {
          f0 = unpack_act2<true>(xv.x); f1 = unpack_act2<true>(xv.y); f2 = unpack_act2<true>(xv.z); f3 = unpack_act2<true>(xv.w);
        } else {
          f0 = unpack_act2<false>(xv.x); f1 = unpack_act2<false>(xv.y); f2 = unpack_act2<false>(xv.z); f3 = unpack_act2<false>(xv.w);
        }
        float a = acc[o];
        a = fmaf(f0.x, w0.x, a); a = fmaf(f0.y, w0.y, a); a = fmaf(f1.x, w0.z, a); a = fmaf(f1.y, w0.w, a);
        a = fmaf(f2.x, w1.x, a); a = fmaf(f2.y, w1.y, a); a = fmaf(f3.x, w1.z, a); a = fmaf(f3.y, w1.w, a);
        acc[o] = a;
      }
    }
  }
#pragma unroll
  for (int o = 0; o < kPostPer; ++o) {
    const int t = t0 + threadIdx.x + o * kPostThreads;
    if (t < L) out[(long)b * L + t] = tanhf(acc[o]);
  }
}

// a-form bf16 [B][L][C] -> residual-stream fp32 NCL [B][C][L]
__global__ void unpack_debug_kernel(const __nv_bfloat16* __restrict__ a, float gain, float* __restrict__ out, int L,
                                    int C, int f16) {
  const long total = (long)gridDim.y * L * C;
  (void)total;
  const int b = blockIdx.y;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < (long)L * C; i += (long)gridDim.x * blockDim.x) {
    const int c = i % C;
    const long t = i / C;
    const float v = unpack_act_rt(a[(long)b * L * C + i], f16);
    out[((long)b * C + c) * L + t] = v >= 0.f ? v : v * gain;
  }
}

// ------------------------------------------------------------------------------------------ launchers
// Output side (cmd_inference.py:114-117): waveform fp32 [-1, 1] -> 16-bit PCM on the device, so that the D2H copy moves
// half the bytes and lands directly behind the WAV header in a pinned host buffer.  round-to-nearest-even of
// clip(x, -1, 1) * 32767, the usual float -> int16 export; 8 samples per thread, 16-byte stores.
__global__ void pcm16_kernel(const float* __restrict__ wav, int16_t* __restrict__ out, long n) {
  const long i8 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i8 + 8 <= n && (reinterpret_cast<uintptr_t>(out + i8) & 15) == 0 && (reinterpret_cast<uintptr_t>(wav + i8) & 15) == 0) {
    const float4 a = *reinterpret_cast<const float4*>(wav + i8), b = *reinterpret_cast<const float4*>(wav + i8 + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int lo = __float2int_rn(fminf(fmaxf(v[2 * j], -1.f), 1.f) * 32767.f);
      const int hi = __float2int_rn(fminf(fmaxf(v[2 * j + 1], -1.f), 1.f) * 32767.f);
      pk[j] = ((uint32_t)hi << 16) | ((uint32_t)lo & 0xffffu);
    }
    *reinterpret_cast<uint4*>(out + i8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  } else {
    for (long i = i8; i < n && i < i8 + 8; ++i)
      out[i] = (int16_t)__float2int_rn(fminf(fmaxf(wav[i], -1.f), 1.f) * 32767.f);
  }
}

int launch_wn_scale(const float* v, const float* g, float* scale, int rows, int inner, cudaStream_t st) {
  wn_scale_kernel<<<rows, 256, 0, st>>>(v, g, scale, inner);
  VD_CUDA(cudaGetLastError());
  return 0;
}
int launch_interleave_bias(const float* b, float* out, int c_out, cudaStream_t st) {
  interleave_bias_kernel<<<(c_out + 255) / 256, 256, 0, st>>>(b, out, c_out);
  VD_CUDA(cudaGetLastError());
  return 0;
}
int launch_pack_conv(const float* w, const float* scale, __nv_bfloat16* wp, int c_out, int c_in, int k,
                     cudaStream_t st, int interleave, int f16, int c_in_src, int c_out_src) {
  const long total = (long)k * c_out * c_in;
  pack_conv_kernel<<<(int)std::min<long>((total + 255) / 256, 4096), 256, 0, st>>>(w, scale, wp, c_out, c_in, k,
                                                                                   interleave, f16,
                                                                                   c_in_src > 0 ? c_in_src : c_in,
                                                                                   c_out_src > 0 ? c_out_src : c_out);
  VD_CUDA(cudaGetLastError());
  return 0;
}
int launch_pack_conv_fold(const float* w, const float* scale, __nv_bfloat16* wp, int C, int c_out, int k, int r,
                          cudaStream_t st, int lo_part, int f16, int c_in_src) {
  const int hk = (k - 1) / 2;
  const int s_min = -((hk + r - 1) / r), s_max = (r - 1 + hk) / r;  // floor(-hk/r), floor((r-1+hk)/r)
  const int ntaps = s_max - s_min + 1;
  const long total = (long)ntaps * r * C * r * C;
  pack_conv_fold_kernel<<<(int)std::min<long>((total + 255) / 256, 4096), 256, 0, st>>>(w, scale, wp, C, c_out, k, r,
                                                                                        s_min, ntaps, lo_part, f16,
                                                                                        c_in_src > 0 ? c_in_src : C);
  VD_CUDA(cudaGetLastError());
  return 0;
}
int launch_pack_convT(const float* w, const float* scale, __nv_bfloat16* wp, int c_in, int c_out, int k, int s, int p,
                      int ntaps, int off0, cudaStream_t st, int f16, int c_in_src, int c_out_src) {
  const long total = (long)ntaps * s * c_out * c_in;
  pack_convT_kernel<<<(int)std::min<long>((total + 255) / 256, 4096), 256, 0, st>>>(w, scale, wp, c_in, c_out, k, s, p,
                                                                                   ntaps, off0, f16,
                                                                                   c_in_src > 0 ? c_in_src : c_in,
                                                                                   c_out_src > 0 ? c_out_src : c_out);
  VD_CUDA(cudaGetLastError());
  return 0;
}
int launch_replicate_bias(const float* b, float* out, int c_out, int reps, cudaStream_t st, int c_out_src) {
  replicate_bias_kernel<<<(c_out * reps + 255) / 256, 256, 0, st>>>(b, out, c_out, reps,
                                                                    c_out_src > 0 ? c_out_src : c_out);
  VD_CUDA(cudaGetLastError());
  return 0;
}
int launch_pcm16(const float* wav, int16_t* out, long n, cudaStream_t st) {
  if (n <= 0) return 0;
  const long threads = (n + 7) / 8;
  pcm16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(wav, out, n);
  VD_CUDA(cudaGetLastError());
  return 0;
}
int launch_sum_bias(const float* b0, const float* b1, const float* b2, const float* b3, float* out, int n,
                    cudaStream_t st) {
  sum_bias_kernel<<<(n + 255) / 256, 256, 0, st>>>(b0, b1, b2, b3, out, n);
  VD_CUDA(cudaGetLastError());
  return 0;
}
int launch_pack_z(const float* z, long sb, long sc, __nv_bfloat16* out, int B, int C, int T, cudaStream_t st,
                  int f16, int c_pad) {
  const int Cp = c_pad > C ? c_pad : C;
  dim3 grid((T + 31) / 32, (Cp + 31) / 32, B), block(32, 8);
  pack_z_kernel<<<grid, block, 0, st>>>(z, sb, sc, out, C, T, f16, Cp);
  VD_CUDA(cudaGetLastError());
  return 0;
}
int launch_cond(const float* wc, const float* bc, const float* g, float* cb, int B, int c_out, int gin,
                cudaStream_t st) {
  dim3 grid((c_out * 32 + 255) / 256, B);
  cond_kernel<<<grid, 256, 0, st>>>(wc, bc, g, cb, c_out, gin);
  VD_CUDA(cudaGetLastError());
  return 0;
}
int launch_conv_post(const __nv_bfloat16* x, const float* w, float* out, int B, int L, int C, cudaStream_t st,
                     int f16) {
  VD_CHECK(C % 8 == 0, "conv_post: channels must be a multiple of 8");
  auto smem_for = [&](int per) { return 7 * C * 4 + (size_t)(kPostThreads * per + 6) * (C + 8) * 2; };
  if (smem_for(4) <= 48 * 1024) {
    dim3 grid((L + kPostThreads * 4 - 1) / (kPostThreads * 4), B);
    conv_post_kernel<4><<<grid, kPostThreads, smem_for(4), st>>>(x, w, out, L, C, f16);
  } else {
    VD_CHECK(smem_for(1) <= 48 * 1024, "conv_post: too many channels");
    dim3 grid((L + kPostThreads - 1) / kPostThreads, B);
    conv_post_kernel<1><<<grid, kPostThreads, smem_for(1), st>>>(x, w, out, L, C, f16);
  }
  VD_CUDA(cudaGetLastError());
  return 0;
}
int launch_unpack_debug(const __nv_bfloat16* a, float gain, float* out, int B, int L, int C, cudaStream_t st,
                        int f16) {
  dim3 grid((unsigned)std::min<long>(((long)L * C + 255) / 256, 8192), B);
  unpack_debug_kernel<<<grid, 256, 0, st>>>(a, gain, out, L, C, f16);
  VD_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vd
