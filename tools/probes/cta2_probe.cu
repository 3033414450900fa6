// Probe (NOT YET RUN ON HARDWARE -- compiled for sm_100a only; first thing to run in the next round, under `timeout 20`):
// one tcgen05.mma.cta_group::2 tile, M = 256 (two CTAs x 128 rows of A), N = 256, K = 64, bf16 -> fp32, with the operand
// layout conv_tc.cu would use for a 256-output-channel layer (stage 0):
//   A = weights  [256 channels][64 ci]      CTA r holds channels [128 r, 128 r + 128)
//   B = activations [256 time rows][64 ci]  CTA r holds rows     [128 r, 128 r + 128)   (both K-major, SWIZZLE_128B)
//   D[channel][time] in TMEM: CTA r's 128 lanes x 256 columns.
// What it must answer: (1) the instruction / commit / alloc forms below are accepted and complete, (2) each CTA's TMEM
// holds rows 128 r .. of D = A * B^T, i.e. the pair reads the peer's half of B at the SAME shared-memory offset.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probes/cta2_probe tools/probes/cta2_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// K-major SWIZZLE_128B shared-memory descriptor (rows of 128 bytes, 8-row atoms 1024 bytes apart), as conv_tc.cu builds it
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;  // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                       // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: bf16 A/B, fp32 D, K-major both, N at [17,23) in units of 8, M at [24,29) in units of 16
__device__ __forceinline__ uint32_t umma_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

constexpr int kRows = 128, kK = 64, kN = 256, kCols = 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
cta2_probe(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D,
           int* __restrict__ status) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                       // [128][64] bf16, 16 KB
  uint8_t* sB = smem + kRows * kK * 2;      // [128][64] bf16, 16 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + kRows * kK * 2);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const uint32_t rank = cluster_ctarank();
  const int t = threadIdx.x, warp = t >> 5;

  // operands: CTA r takes rows [128 r, 128 r + 128) of A and of B; 16-byte chunks XOR-swizzled by (row & 7)
  for (int i = t; i < kRows * (kK / 8); i += blockDim.x) {
    const int row = i / (kK / 8), chunk = i % (kK / 8);
    const uint4 a = *reinterpret_cast<const uint4*>(A + ((size_t)(rank * kRows + row) * kK + chunk * 8));
    const uint4 b = *reinterpret_cast<const uint4*>(B + ((size_t)(rank * kRows + row) * kK + chunk * 8));
    *reinterpret_cast<uint4*>(sA + row * 128 + ((chunk ^ (row & 7)) << 4)) = a;
    *reinterpret_cast<uint4*>(sB + row * 128 + ((chunk ^ (row & 7)) << 4)) = b;
  }
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy operand stores -> async proxy (tensor core)
  if (warp == 0) {   // the same logical warp of BOTH CTAs, same destination offset (cute::TMEM::Allocator2Sm's contract)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(kCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();   // both CTAs' operands, barriers and TMEM are in place
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;

  if (rank == 0 && warp == 1) {   // the leader CTA issues for the pair; one elected lane
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    if (leader) {
      const uint32_t idesc = umma_idesc(256, kN);
      for (int k = 0; k < kK / 16; ++k) {
        const uint64_t da = umma_desc(smem_u32(sA) + k * 32), db = umma_desc(smem_u32(sB) + k * 32);
        const uint32_t acc = k > 0 ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
            : "memory");
      }
      // completion -> the mbarrier at the same offset in BOTH CTAs
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                       smem_u32(bar)),
                   "h"((uint16_t)0x3)
                   : "memory");
    }
    __syncwarp();
  }

  // every thread waits for the accumulator (bounded: a protocol mistake must not hang the box)
  uint32_t ok = 0;
  for (uint32_t spin = 0; spin < (1u << 22) && !ok; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(0u)
                 : "memory");
  }
  if (!ok) {
    if (t == 0) status[rank] = -1;   // timed out
  } else {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp w reads TMEM lanes [32 w, 32 w + 32) = channels 128 rank + 32 w + lane; 256 columns = time rows
    for (int c0 = 0; c0 < kN; c0 += 16) {
      uint32_t v[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float* out = D + (size_t)(rank * kRows + t) * kN + c0;
      for (int j = 0; j < 16; ++j) out[j] = __uint_as_float(v[j]);
    }
    if (t == 0) status[rank] = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();   // neither CTA may free while the peer can still be reading its shared memory / TMEM
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kCols));
}

int main() {
  const int M = 256, N = kN, K = kK;
  std::vector<__nv_bfloat16> hA((size_t)M * K), hB((size_t)N * K);
  std::vector<float> fA(hA.size()), fB(hB.size());
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16((rand() % 17 - 8) / 8.f); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16((rand() % 13 - 6) / 4.f); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  int* dS;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, (size_t)M * N * 4); cudaMalloc(&dS, 8);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, (size_t)M * N * 4); cudaMemset(dS, 0, 8);
  const size_t smem = 2 * kRows * kK * 2 + 1024 + 64;
  cudaFuncSetAttribute(cta2_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cta2_probe<<<2, 128, smem>>>(dA, dB, dD, dS);
  cudaError_t e = cudaDeviceSynchronize();
  int st[2] = {0, 0};
  std::vector<float> hD((size_t)M * N);
  cudaMemcpy(st, dS, 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  printf("launch: %s, status cta0 %d cta1 %d (1 = accumulator arrived, -1 = timed out)\n", cudaGetErrorString(e), st[0], st[1]);
  double worst = 0;
  int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)fA[(size_t)m * K + k] * fB[(size_t)n * K + k];
      const double d = fabs(ref - hD[(size_t)m * N + n]);
      if (d > worst) worst = d;
      if (d > 1e-3 && bad++ < 8) printf("  D[%d][%d] = %f, expected %f\n", m, n, hD[(size_t)m * N + n], ref);
    }
  printf("max |D - A*B^T| = %g over %d x %d, mismatches %d\n", worst, M, N, bad);
  return bad != 0;
}
