// Shared by fps.cu (register/cluster kernel) and fps_bucket.cu (spatially bucketed kernel).
#pragma once
#include "common.cuh"

namespace ws3d {

// Tie-break order of the reference's shared-memory tree (sampling_gpu.cu:86-91,139-205): among equal
// running distances the winner is the smallest bit-reversed (k mod BS), then the smallest k.
constexpr uint32_t kNoKey = 0xFFFFFFFFu;

// key(k): top L bits = bit-reversed (k mod 2^L), low 32-L bits = k >> L.
__host__ __device__ inline uint32_t brev32(uint32_t v) {
#ifdef __CUDA_ARCH__
  return __brev(v);
#else
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
  v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
  return (v >> 16) | (v << 16);
#endif
}
__host__ __device__ inline uint32_t fps_key(uint32_t k, int L) {
  if (L == 0) return k;
  const uint32_t lowmask = 0xFFFFFFFFu >> L;
  return (brev32(k) & ~lowmask) | (k >> L);
}
__host__ __device__ inline uint32_t fps_unkey(uint32_t key, int L) {
  if (L == 0) return key;
  const uint32_t lowmask = 0xFFFFFFFFu >> L;
  return ((key & lowmask) << L) | brev32(key & ~lowmask);
}


struct FpsParams {
  int n, m, L;       // points, samples, log2(reference block size)
  int log2T;         // blockDim.x == 1 << log2T
  const float *xyz;  // (B,N,3)
  float *temp;       // (B,N) or null
  int *idx;          // (B,M)
  float *new_xyz;    // (B,M,3) or null
};

// ---- cluster / mbarrier / DSMEM primitives (raw PTX) --------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d,
                                            uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
               "r"(a), "r"(b), "r"(c), "r"(d), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void st_async_v2(uint32_t raddr, uint32_t a, uint32_t b, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(raddr), "r"(a),
               "r"(b), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t raddr, uint32_t a, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(a),
               "r"(rbar)
               : "memory");
}

// fps_bucket.cu
bool fps_bucket_applicable(int b, int n, int m);
int fps_bucket_launch(const FpsParams &prm, int b, cudaStream_t stream);
// fps_smem.cu: same shapes, running distances in shared memory, several clouds per SM
int fps_smem_launch(const FpsParams &prm, int b, cudaStream_t stream);
int fps_smem_clouds_per_cta(int b, int n);

}  // namespace ws3d
