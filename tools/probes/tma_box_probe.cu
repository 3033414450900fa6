// Probe: shared-memory layout of a TMA box {32 ch, 1, 2 phases, 8 rows, 1} (bf16) with SWIZZLE_128B.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tm, uint16_t* out) {
  __shared__ __align__(1024) uint16_t tile[8 * 64];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1024));
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(tile)), "l"(&tm), "r"(smem_u32(&bar)), "r"(0), "r"(1), "r"(2), "r"(0), "r"(0) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.b32 %0,1,0,p; }" : "=r"(ok) : "r"(smem_u32(&bar)));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = tile[i];
}
int main() {
  // tensor x[b][t][c], C=32, d=3, r=4: value = t*32 + c (as integer in u16, t < 2000)
  const int C = 32, d = 3, r = 4, L = 1200;
  uint16_t* h = new uint16_t[L * C];
  for (int t = 0; t < L; ++t) for (int c = 0; c < C; ++c) h[t * C + c] = (uint16_t)(t * 32 + c);
  uint16_t* dx; cudaMalloc(&dx, L * C * 2 + 8192); cudaMemcpy(dx, h, L * C * 2, cudaMemcpyHostToDevice);
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  CUtensorMap tm;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)d, (cuuint64_t)r, (cuuint64_t)(L / (d * r)), 1};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)d * C * 2, (cuuint64_t)d * r * C * 2, (cuuint64_t)L * C * 2};
  cuuint32_t box[5] = {32, 1, 2, 8, 1}, es[5] = {1, 1, 1, 1, 1};
  CUresult rc = ((EncodeTiledFn)f)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc %d\n", (int)rc);
  uint16_t* dout; cudaMalloc(&dout, 1024);
  probe<<<1, 128>>>(tm, dout);
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  uint16_t o[512]; cudaMemcpy(o, dout, 1024, cudaMemcpyDeviceToHost);
  // coordinates (0, rho=1, psi=2, row=0): expected element (row n, psi, c) = t*32+c with t = d*(r*n+psi)+rho
  for (int line = 0; line < 8; ++line) {
    printf("smem 128B line %d:", line);
    for (int ch = 0; ch < 8; ++ch) { int v = o[line * 64 + ch * 8]; printf("  [t=%d c=%d]", v / 32, v % 32); }
    printf("\n");
  }
  return 0;
}
