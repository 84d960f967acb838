// Stand-alone probe for the register <-> (lane, column) mapping of tcgen05.ld shapes (debug tool, not part of the library).
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/tmem_ld_probe.cu -o tools/tmem_ld_probe.bin
// Every thread writes value lane * 1000 + column along its own TMEM lane (32x32b stores), then warp 0 reads columns 0..15 of
// lanes 0..15 and of lanes 16..31 with 16x256b.x2 and prints what each thread's registers hold.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(int *out) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base_s;
  const uint32_t lane_addr = base + ((uint32_t)(warp * 32) << 16);
  uint32_t v[16];
  for (int c = 0; c < 16; ++c) v[c] = (uint32_t)(tid * 1000 + c);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(lane_addr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
               "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int half = 0; half < 2; ++half) {
    uint32_t r[8];
    const uint32_t a = lane_addr + ((uint32_t)(half * 16) << 16);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(a) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int k = 0; k < 8; ++k) out[((warp * 2 + half) * 32 + lane) * 8 + k] = (int)r[k];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(32) : "memory");
}

int main() {
  int *d, h[4 * 2 * 32 * 8];
  cudaMalloc(&d, sizeof(h));
  cudaMemset(d, 0xFF, sizeof(h));
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int w = 0; w < 4; w += 3)
    for (int half = 0; half < 2; ++half)
      for (int t = 0; t < 32; ++t) {
        printf("warp %d half %d thread %2d:", w, half, t);
        for (int k = 0; k < 8; ++k) { const int x = h[((w * 2 + half) * 32 + t) * 8 + k]; printf(" r%d=(lane %3d,col %2d)", k, x / 1000, x % 1000); }
        printf("\n");
      }
  return 0;
}
