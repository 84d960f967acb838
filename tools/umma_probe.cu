// Stand-alone probe for tcgen05.mma.kind::tf32 operand layouts (debug tool, not part of the library).
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/umma_probe.cu -o tools/umma_probe.bin
// One CTA, M=128, N=NN, K=32 (four K=8 MMAs).  Shared memory is filled by ordinary stores in a
// chosen layout; the result D is compared against a host product.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int M = 128, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

struct Cfg {
  int n;            // N (multiple of 32, <= 256)
  int b_mode;       // 0: K-major SW128, 1: MN-major SW128 (32-column blocks 4 KB apart)
  uint32_t b_lbo, b_sbo, b_step;  // descriptor fields and per-K=8 start-address step for B
  uint32_t idesc;
  int with_mask;    // pass the 4-word disable mask form
  uint32_t b_layout;
};

__global__ void __launch_bounds__(128, 1) probe(const float *A, const float *B, float *D, Cfg cfg) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t *g = raw + (base - smem_u32(raw));
  float *sa = reinterpret_cast<float *>(g);                 // 16 KB
  float *sb = reinterpret_cast<float *>(g + 16384);         // up to 32 KB
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < M * K; i += 128) {
    const int m = i / K, k = i % K;
    const int off = (m / 8) * 1024 + (m % 8) * 128 + (((k / 4) ^ (m % 8)) * 16) + (k % 4) * 4;
    sa[off / 4] = A[m * K + k];
  }
  for (int i = tid; i < K * cfg.n; i += 128) {
    const int k = i / cfg.n, n = i % cfg.n;
    int off;
    if (cfg.b_mode == 0) off = (n / 8) * 1024 + (n % 8) * 128 + (((k / 4) ^ (n % 8)) * 16) + (k % 4) * 4;
    else if (cfg.b_mode == 1) off = (n / 32) * 4096 + k * 128 + ((((n % 32) / 4) ^ (k % 8)) * 16) + (n % 4) * 4;
    else off = (n / 32) * 4096 + k * 128 + ((((n % 32) / 8) ^ (k % 4)) * 32) + (n % 8) * 4;  // 32B chunks, 4-row period
    sb[off / 4] = B[k * cfg.n + n];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> async proxy (MMA)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    for (int kk = 0; kk < 4; ++kk) {
      const uint64_t da = make_desc(base + kk * 32, 16, 1024, 2);
      const uint64_t db = make_desc(base + 16384 + kk * cfg.b_step, cfg.b_lbo, cfg.b_sbo, cfg.b_layout);
      const uint32_t acc = kk ? 1u : 0u;
      if (cfg.with_mask) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db),
                     "r"(cfg.idesc), "r"(acc), "r"(0)
                     : "memory");
      } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(cfg.idesc),
                     "r"(acc)
                     : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everyone waits for the MMAs
  {
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
      if (clock64() - t0 > 2000000000LL) __trap();
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = warp * 32 + (tid & 31);
  for (int c = 0; c < cfg.n; c += 32) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int t = 0; t < 32; ++t) D[row * cfg.n + c + t] = __uint_as_float(r[t]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

static uint32_t idesc(int n, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

int main() {
  std::vector<float> A(M * K), B, D, want;
  srand(1);
  for (auto &v : A) v = (float)(rand() % 7 - 3);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, M * K * 4); cudaMalloc(&dB, K * 256 * 4); cudaMalloc(&dD, M * 256 * 4);
  cudaMemcpy(dA, A.data(), M * K * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024);
  struct V { const char *name; Cfg c; };
  std::vector<V> vs;
  for (int n : {32, 64, 256}) {
    const int mask = 0;
    vs.push_back({"B K-major  sbo=1024 step=32", {n, 0, 16, 1024, 32, idesc(n, 0), mask, 2}});
    vs.push_back({"B MN SW128 lbo=4096 sbo=1024 step=1024", {n, 1, 4096, 1024, 1024, idesc(n, 1), mask, 2}});
    vs.push_back({"B MN BASE32B lbo=4096 sbo=512 step=1024", {n, 2, 4096, 512, 1024, idesc(n, 1), mask, 1}});
    vs.push_back({"B MN BASE32B lbo=512 sbo=4096 step=1024", {n, 2, 512, 4096, 1024, idesc(n, 1), mask, 1}});
    vs.push_back({"B MN BASE32B lbo=4096 sbo=1024 step=1024", {n, 2, 4096, 1024, 1024, idesc(n, 1), mask, 1}});
  }
  for (auto &v : vs) {
    const int n = v.c.n;
    B.assign(K * n, 0.f); D.assign(M * n, -777.f); want.assign(M * n, 0.f);
    for (auto &x : B) x = (float)(rand() % 7 - 3);
    for (int m = 0; m < M; ++m)
      for (int j = 0; j < n; ++j) {
        float s = 0;
        for (int k = 0; k < K; ++k) s += A[m * K + k] * B[k * n + j];
        want[m * n + j] = s;
      }
    cudaMemcpy(dB, B.data(), K * n * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xFF, M * 256 * 4);
    probe<<<1, 128, 16384 + 32768 + 1024>>>(dA, dB, dD, v.c);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s n=%3d mask=%d: CUDA error %s\n", v.name, n, v.c.with_mask, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, M * n * 4, cudaMemcpyDeviceToHost);
    int bad = 0, zeros = 0;
    for (int i = 0; i < M * n; ++i) { bad += D[i] != want[i]; zeros += D[i] == 0.f; }
    printf("%-44s n=%3d mask=%d: bad=%d/%d zeros=%d  D[0..3]=%g %g %g %g want %g %g %g %g\n", v.name, n, v.c.with_mask, bad, M * n, zeros,
           D[0], D[1], D[2], D[3], want[0], want[1], want[2], want[3]);
  }
  return 0;
}
