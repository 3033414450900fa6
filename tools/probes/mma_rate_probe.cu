// Probe: cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) as a function of N and of the shared-memory
// row pitch of the A / B operands (64-byte rows = SWIZZLE_64B, 128-byte rows = SWIZZLE_128B), issued back to back by one
// thread the way the conv kernels issue them (A start address shifted by a few rows per "tap", K advanced by 32 bytes).
// Answers: what does an N = 32 / 64 time-as-M MMA cost when the activation rows are 128 bytes (two time phases of a
// 32-channel tensor side by side, "r = 2 fold") instead of 64 bytes?
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/mma_rate_probe tools/probes/mma_rate_probe.cu
#include <cstdint>
#include <cstdio>

#include "../../personalized_text-to-speech_b200/csrc/ptx.cuh"

using namespace vd;

struct Cfg { int n, a_rowb, b_rowb, koff; };

__global__ void __launch_bounds__(128, 1) probe(const Cfg* cfgs, int ncfg, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                  // 320 rows x 128 B
  uint8_t* sB = smem + 320 * 128;      // 256 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 256 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (320 + 256) * 128 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (threadIdx.x < 32) { tmem_alloc(slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x < 32) {
    const uint32_t leader = elect_one();
    uint32_t phase = 0;
    for (int c = 0; c < ncfg; ++c) {
      const Cfg cf = cfgs[c];
      const uint32_t idesc = umma_idesc_f16(cf.n, false);
      const uint32_t a_hi = umma_desc_hi(cf.a_rowb), b_hi = umma_desc_hi(cf.b_rowb);
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(sA)), b_lo0 = umma_desc_lo(smem_u32(sB));
      __syncwarp();
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t tap = i % 7;                                   // row shift of the A tile, like a filter tap
        const uint32_t kk = (i & 1) * 2 + ((cf.koff * (i & 2)) >> 1) * 4;   // K advance: 32 B (+ 64 B when koff)
        umma_f16_lohi(tmem, a_lo0 + ((tap * cf.a_rowb) >> 4) + kk, a_hi, b_lo0 + kk, b_hi, idesc, i > 0, leader);
      }
      if (leader) umma_commit(bar);
      __syncwarp();
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      const long long t1 = clock64();
      if (threadIdx.x == 0) out[blockIdx.x * ncfg + c] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int main() {
  const Cfg h[] = {{32, 64, 64, 0},   {32, 128, 64, 0},  {32, 128, 128, 0}, {32, 128, 128, 1}, {64, 128, 128, 0},
                   {64, 128, 128, 1}, {64, 64, 64, 0},   {128, 128, 128, 0}, {256, 128, 128, 0}, {16, 128, 128, 0},
                   {96, 128, 128, 0}, {48, 128, 128, 0}};
  const int ncfg = sizeof(h) / sizeof(h[0]), iters = 2048, grid = 148;
  Cfg* d; long long* o;
  cudaMalloc(&d, sizeof(h)); cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaMalloc(&o, grid * ncfg * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int rep = 0; rep < 2; ++rep) probe<<<grid, 128, 100 * 1024>>>(d, ncfg, iters, o);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  static long long r[148 * 16];
  cudaMemcpy(r, o, grid * ncfg * sizeof(long long), cudaMemcpyDeviceToHost);
  printf("# cycles per M=128,K=16 tcgen05.mma (one issuing thread, %d back-to-back MMAs, 148 CTAs at once; min / median over CTAs)\n", iters);
  printf("#   N  A_row_bytes  B_row_bytes  K+64B   cycles/MMA(min)  (max)   ideal(N/2)\n");
  for (int c = 0; c < ncfg; ++c) {
    long long mn = 1LL << 60, mx = 0;
    for (int b = 0; b < grid; ++b) { mn = r[b * ncfg + c] < mn ? r[b * ncfg + c] : mn; mx = r[b * ncfg + c] > mx ? r[b * ncfg + c] : mx; }
    printf("  %4d  %6d  %10d  %6d   %10.1f  %10.1f   %6.1f\n", h[c].n, h[c].a_rowb, h[c].b_rowb, h[c].koff, (double)mn / iters,
           (double)mx / iters, h[c].n / 2.0);
  }
  return 0;
}
