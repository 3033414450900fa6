// Probe: register layout of tcgen05.ld.16x256b and of stmatrix/ldmatrix .trans (used by the fragment-layout epilogue).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(uint32_t* out) {
  __shared__ uint32_t slot;
  __shared__ __align__(16) uint16_t sm[32 * 8 * 4];
  const int t = threadIdx.x;
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(32));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  __syncthreads();
  const uint32_t tb = slot;
  uint32_t v[16];
  for (int j = 0; j < 16; ++j) v[j] = t * 100 + j;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(tb),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
               "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]));
  asm volatile("tcgen05.wait::st.sync.aligned;");
  __syncwarp();
  uint32_t r[8];
  for (int half = 0; half < 2; ++half) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tb + ((uint32_t)(half * 16) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int i = 0; i < 8; ++i) out[(half * 32 + t) * 8 + i] = r[i];
  }
  // stmatrix .trans: thread t holds M_m[t/4][2(t%4)], M_m[t/4][2(t%4)+1] for matrices m = 0..3; id = m*1000 + a*10 + b
  uint32_t f[4];
  for (int m = 0; m < 4; ++m) {
    const int a = t / 4, b = 2 * (t % 4);
    f[m] = (uint32_t)(m * 1000 + a * 10 + b) | ((uint32_t)(m * 1000 + a * 10 + b + 1) << 16);
  }
  // lane i gives the address of row (i%8) of matrix (i/8): rows are 16 bytes; matrix m occupies sm[m*64 .. m*64+63]
  const uint32_t addr = smem_u32(sm + (t / 8) * 64 + (t % 8) * 8);
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(f[0]), "r"(f[1]), "r"(f[2]), "r"(f[3]));
  __syncwarp();
  for (int i = t; i < 256; i += 32) out[512 + i] = sm[i];
  __syncwarp();
  // ldmatrix .trans of the same memory
  uint32_t g[4];
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(g[0]), "=r"(g[1]), "=r"(g[2]), "=r"(g[3]) : "r"(addr));
  for (int m = 0; m < 4; ++m) out[768 + t * 4 + m] = g[m];
  __syncthreads();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(32));
}
int main() {
  uint32_t* d; cudaMalloc(&d, 4096 * 4); cudaMemset(d, 0, 4096 * 4);
  probe<<<1, 32>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  static uint32_t h[4096]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  printf("tcgen05.ld.16x256b.x2 (value = lane*100 + col):\n");
  for (int half = 0; half < 2; ++half) for (int t = 0; t < 32; t += (t < 8 ? 1 : 8)) {
    printf(" half %d thread %2d:", half, t);
    for (int i = 0; i < 8; ++i) printf(" %5u", h[(half * 32 + t) * 8 + i]);
    printf("\n");
  }
  printf("stmatrix.trans smem (matrix m rows of 8 b16; id = m*1000 + a*10 + b, thread t holds a=t/4,b=2(t%%4),+1):\n");
  for (int m = 0; m < 4; ++m) for (int r = 0; r < 8; r += (m == 0 ? 1 : 4)) {
    printf(" m%d row %d:", m, r);
    for (int c = 0; c < 8; ++c) printf(" %5u", h[512 + m * 64 + r * 8 + c]);
    printf("\n");
  }
  printf("ldmatrix.trans back (thread: g0 lo/hi ... ):\n");
  for (int t = 0; t < 8; ++t) { printf(" thread %d:", t); for (int m = 0; m < 4; ++m) printf(" (%u,%u)", h[768 + t * 4 + m] & 0xffff, h[768 + t * 4 + m] >> 16); printf("\n"); }
  return 0;
}
