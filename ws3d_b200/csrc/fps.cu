// Furthest point sampling for B200.
//
// Replaces pointnet2_lib/pointnet2/src/sampling_gpu.cu:93-253 (one 1024-thread CTA per
// cloud, xyz and the running min-distance re-read from global memory every iteration,
// ten __syncthreads per iteration).
//
// Design: FPS is a latency chain of m-1 dependent iterations, not a bandwidth problem.
//   * a cloud is split over a thread-block CLUSTER of C CTAs (C*B ~ number of SMs);
//   * every thread keeps its P points (x, y, z, running min-distance, tie key) in
//     REGISTERS for the whole kernel: after the first load nothing touches HBM/L2 again;
//   * per iteration: P distance updates per thread, a warp arg-max with two redux.sync,
//     one __syncthreads, warp 0 picks the CTA winner and posts {value, key, xyz} into the
//     shared memory of every CTA of the cluster with st.async (DSMEM) signalling an
//     mbarrier there; every warp then picks the cluster winner from C records.
//   * the sampled coordinates are emitted from the same loop (fused gather).
//
// Bit-exactness: the reference's winner is the maximum running distance; ties go to the
// smallest bit-reversed (k mod BS) -- its shared-memory tree keeps the lower slot of each
// pair -- and then to the smallest k (strict '>' inside a thread), BS = opt_n_threads(n).
// That order is encoded as a 32-bit key (smaller wins) so any thread/CTA decomposition
// reproduces it.  Distances use the reference's rounding order (sqdist_ref).
#include <limits.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "fps_common.cuh"
#include "tma.cuh"

namespace ws3d {
namespace {

constexpr int kMaxCluster = 16;

struct __align__(16) Rec {  // one candidate: 20 payload bytes in a 32-byte slot
  int v;                    // running-distance bits (>= 0 for real points)
  uint32_t key;
  float x, y;
  float z;
  float pad[3];
};


template <int P>
__device__ __forceinline__ void sort_by_key(uint32_t (&key)[P], float (&x)[P], float (&y)[P], float (&z)[P],
                                            float (&t)[P]) {
  // odd-even transposition network on registers (fully unrolled, init-time only)
#pragma unroll
  for (int round = 0; round < P; ++round) {
#pragma unroll
    for (int i = (round & 1); i + 1 < P; i += 2) {
      if (key[i + 1] < key[i]) {
        uint32_t tk = key[i]; key[i] = key[i + 1]; key[i + 1] = tk;
        float f;
        f = x[i]; x[i] = x[i + 1]; x[i + 1] = f;
        f = y[i]; y[i] = y[i + 1]; y[i + 1] = f;
        f = z[i]; z[i] = z[i + 1]; z[i + 1] = f;
        f = t[i]; t[i] = t[i + 1]; t[i + 1] = f;
      }
    }
  }
}

// One cluster of C CTAs (C == 1 when !CLUSTER) per cloud; grid = (C, B).
template <int P, bool CLUSTER>
__global__ void __launch_bounds__(1024, 1) fps_cluster_kernel(FpsParams prm) {
  extern __shared__ float4 s_pts[];                 // P*T own points, slot = p*T + tid (load order)
  __shared__ int2 s_wrec[2][32];                    // per-warp candidate {v, key}
  __shared__ Rec s_crec[2][kMaxCluster];            // per-CTA candidates of the cluster
  __shared__ __align__(8) unsigned long long s_bar[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = blockDim.x, nwarps = T >> 5;
  const uint32_t rank = CLUSTER ? cluster_ctarank() : 0u;
  const uint32_t C = CLUSTER ? cluster_nctarank() : 1u;
  const int CT = (int)C * T;
  const int log2CT = 31 - __clz(CT);
  const int g = (int)rank * T + tid;
  const int n = prm.n, m = prm.m, L = prm.L;
  const size_t cloud = blockIdx.y;
  const float *xyz = prm.xyz + cloud * (size_t)n * 3;
  float *temp = prm.temp ? prm.temp + cloud * (size_t)n : nullptr;
  int *idx = prm.idx + cloud * (size_t)m;
  float *new_xyz = prm.new_xyz ? prm.new_xyz + cloud * (size_t)m * 3 : nullptr;

  const uint32_t bar_base = smem_u32(&s_bar[0]);
  if (CLUSTER) {
    if (tid == 0) {
      mbar_init(bar_base, 1);
      mbar_init(bar_base + 8, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }

  float x[P], y[P], z[P], t[P];
  uint32_t key[P];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int k = p * CT + g;
    if (k < n) {
      x[p] = __ldg(xyz + (size_t)k * 3 + 0);
      y[p] = __ldg(xyz + (size_t)k * 3 + 1);
      z[p] = __ldg(xyz + (size_t)k * 3 + 2);
      t[p] = temp ? temp[k] : 1e10f;
      key[p] = fps_key((uint32_t)k, L);
    } else {
      x[p] = y[p] = z[p] = 0.f;
      t[p] = -1.f;  // never beats a real point (real running distances are >= 0)
      key[p] = kNoKey;
    }
    s_pts[p * T + tid] = make_float4(x[p], y[p], z[p], 0.f);
  }
  // strict '>' below must meet a thread's points in ascending key order
  sort_by_key<P>(key, x, y, z, t);

  float cx = __ldg(xyz + 0), cy = __ldg(xyz + 1), cz = __ldg(xyz + 2);  // idx[0] = 0
  if (g == 0 && m > 0) {
    idx[0] = 0;
    if (new_xyz) { new_xyz[0] = cx; new_xyz[1] = cy; new_xyz[2] = cz; }
  }
  if (CLUSTER) cluster_sync_all(); else __syncthreads();

  // remote addresses of this CTA's record slot and of the barrier in CTA `lane`, per parity (loop invariant)
  uint32_t rdst[2] = {0u, 0u}, rbars[2] = {0u, 0u};
  if (CLUSTER && lane < (int)C) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      rdst[q] = mapa(smem_u32(&s_crec[q][rank]), (uint32_t)lane);
      rbars[q] = mapa(bar_base + 8u * (uint32_t)q, (uint32_t)lane);
    }
  }

  for (int it = 0; it + 1 < m; ++it) {
    const int par = it & 1;
    const uint32_t bar = bar_base + 8u * (uint32_t)par;
    float best = -1.f;
    uint32_t bkey = kNoKey;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const float d = sqdist_ref(x[p] - cx, y[p] - cy, z[p] - cz);
      t[p] = fminf(d, t[p]);
      if (t[p] > best) { best = t[p]; bkey = key[p]; }
    }
    const int vb = __float_as_int(best);
    const int wv = __reduce_max_sync(0xFFFFFFFFu, vb);
    const uint32_t wk = __reduce_min_sync(0xFFFFFFFFu, vb == wv ? bkey : kNoKey);
    if (lane == 0) s_wrec[par][warp] = make_int2(wv, (int)wk);
    __syncthreads();

    // CTA winner: every warp reduces the <= 32 warp candidates redundantly (no second barrier);
    // in cluster mode only warp 0 needs it, to post the CTA's candidate to its peers.
    uint32_t win_key = kNoKey;
    if (!CLUSTER || warp == 0) {
      int2 r = lane < nwarps ? s_wrec[par][lane] : make_int2(INT_MIN, (int)kNoKey);
      const int bv = __reduce_max_sync(0xFFFFFFFFu, r.x);
      const uint32_t bk = __reduce_min_sync(0xFFFFFFFFu, r.x == bv ? (uint32_t)r.y : kNoKey);
      float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bk != kNoKey) {
        const int k = (int)fps_unkey(bk, L);
        // C and T are powers of two: the slot of point k is a shift and a mask (an integer division here costs
        // ~20 dependent instructions on the iteration's critical path)
        const int p = k >> log2CT, gg = k & (CT - 1);
        pt = s_pts[p * T + (gg - (int)rank * T)];
      }
      if (CLUSTER) {
        if (lane == 0) mbar_expect_tx(bar, C * 20u);
        if (lane < (int)C) {
          const uint32_t dst = par ? rdst[1] : rdst[0];
          const uint32_t rbar = par ? rbars[1] : rbars[0];
          st_async_v4(dst, (uint32_t)bv, bk, __float_as_uint(pt.x), __float_as_uint(pt.y), rbar);
          st_async_b32(dst + 16, __float_as_uint(pt.z), rbar);
        }
      } else {
        win_key = bk; cx = pt.x; cy = pt.y; cz = pt.z;
      }
    }

    if (CLUSTER) {
      mbar_wait(bar, (uint32_t)(it >> 1) & 1u);
      int v = INT_MIN;
      uint32_t kk = kNoKey;
      float rx = 0.f, ry = 0.f, rz = 0.f;
      if (lane < (int)C) {
        const Rec &rc = s_crec[par][lane];
        v = rc.v; kk = rc.key; rx = rc.x; ry = rc.y; rz = rc.z;
      }
      const int bv = __reduce_max_sync(0xFFFFFFFFu, v);
      win_key = __reduce_min_sync(0xFFFFFFFFu, v == bv ? kk : kNoKey);
      const int src = __ffs(__ballot_sync(0xFFFFFFFFu, v == bv && kk == win_key)) - 1;
      cx = __shfl_sync(0xFFFFFFFFu, rx, src);
      cy = __shfl_sync(0xFFFFFFFFu, ry, src);
      cz = __shfl_sync(0xFFFFFFFFu, rz, src);
    }
    if (g == 0) {
      idx[it + 1] = (int)fps_unkey(win_key, L);
      if (new_xyz) {
        new_xyz[(size_t)(it + 1) * 3 + 0] = cx;
        new_xyz[(size_t)(it + 1) * 3 + 1] = cy;
        new_xyz[(size_t)(it + 1) * 3 + 2] = cz;
      }
    }
  }

  if (temp) {
#pragma unroll
    for (int p = 0; p < P; ++p)
      if (key[p] != kNoKey) temp[fps_unkey(key[p], L)] = t[p];
  }
  if (CLUSTER) cluster_sync_all();  // nobody leaves while a peer could still post to it
}

// One-level exchange: every WARP posts its candidate {value, key} straight to every CTA of the cluster and every
// warp then reduces the C * nwarps candidates itself.  Compared with fps_cluster_kernel this removes the CTA
// barrier and warp 0's intermediate arg-max from the serial chain of an iteration (warp arg-max -> DSMEM post ->
// final arg-max instead of warp arg-max -> barrier -> CTA arg-max -> DSMEM post -> final arg-max).  The winner's
// coordinates are not shipped: every CTA keeps the whole cloud in shared memory (3 n floats <= 192 KB).
constexpr int kFlatMaxRecs = 128;   // C * warps per CTA: four candidates per lane in the final reduction

template <int P, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) fps_flat_kernel(FpsParams prm) {
  extern __shared__ __align__(16) float s_cloud[];     // x[n], y[n], z[n]
  __shared__ __align__(8) uint2 s_rec[2][kFlatMaxRecs];
  __shared__ __align__(8) unsigned long long s_bar[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = blockDim.x, nwarps = T >> 5;
  const uint32_t rank = cluster_ctarank();
  const uint32_t C = cluster_nctarank();
  const int CT = (int)C * T;
  const int g = (int)rank * T + tid;
  const int R = (int)C * nwarps;
  const int n = prm.n, m = prm.m, L = prm.L;
  const size_t cloud = blockIdx.y;
  const float *xyz = prm.xyz + cloud * (size_t)n * 3;
  float *temp = prm.temp ? prm.temp + cloud * (size_t)n : nullptr;
  int *idx = prm.idx + cloud * (size_t)m;
  float *new_xyz = prm.new_xyz ? prm.new_xyz + cloud * (size_t)m * 3 : nullptr;
  float *s_x = s_cloud, *s_y = s_cloud + n, *s_z = s_cloud + 2 * n;

  const uint32_t bar_base = smem_u32(&s_bar[0]);
  if (tid == 0) {
    mbar_init(bar_base, 1);
    mbar_init(bar_base + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int e = tid; e < n * 3; e += T) {   // coalesced read of the packed cloud, de-interleaved into x / y / z
    const float v = __ldg(xyz + e);
    const int k = e / 3, a = e - k * 3;
    s_cloud[a * n + k] = v;
  }

  float x[P], y[P], z[P], t[P];
  uint32_t key[P];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int k = p * CT + g;
    if (k < n) {
      x[p] = __ldg(xyz + (size_t)k * 3 + 0);
      y[p] = __ldg(xyz + (size_t)k * 3 + 1);
      z[p] = __ldg(xyz + (size_t)k * 3 + 2);
      t[p] = temp ? temp[k] : 1e10f;
      key[p] = fps_key((uint32_t)k, L);
    } else {
      x[p] = y[p] = z[p] = 0.f;
      t[p] = -1.f;  // never beats a real point (real running distances are >= 0)
      key[p] = kNoKey;
    }
  }
  sort_by_key<P>(key, x, y, z, t);  // strict '>' below must meet a thread's points in ascending key order

  float cx = __ldg(xyz + 0), cy = __ldg(xyz + 1), cz = __ldg(xyz + 2);  // idx[0] = 0
  if (g == 0 && m > 0) {
    idx[0] = 0;
    if (new_xyz) { new_xyz[0] = cx; new_xyz[1] = cy; new_xyz[2] = cz; }
  }
  // remote address of this warp's record slot and of the barrier in CTA `lane`, per parity
  uint32_t rdst[2] = {0u, 0u}, rbars[2] = {0u, 0u};
  if (lane < (int)C) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      rdst[q] = mapa(smem_u32(&s_rec[q][rank * nwarps + warp]), (uint32_t)lane);
      rbars[q] = mapa(bar_base + 8u * (uint32_t)q, (uint32_t)lane);
    }
  }
  cluster_sync_all();  // barriers and cloud copies are ready everywhere

  for (int it = 0; it + 1 < m; ++it) {
    const int par = it & 1;
    const uint32_t bar = bar_base + 8u * (uint32_t)par;
    if (tid == 0) mbar_expect_tx(bar, (uint32_t)R * 8u);
    // distance update, two points per packed instruction; running arg-max in ascending key order (strict '>').
    // (Measured: replacing the running scan by a max tree + key tree is slower -- more instructions to issue.)
    float best = -1.f;
    uint32_t bkey = kNoKey;
    if (P >= 2) {
#pragma unroll
      for (int p = 0; p + 1 < P; p += 2) {
        float d0, d1;
        sqdist_ref_x2(x[p], x[p + 1], y[p], y[p + 1], z[p], z[p + 1], cx, cy, cz, d0, d1);
        t[p] = fminf(d0, t[p]);
        t[p + 1] = fminf(d1, t[p + 1]);
        if (t[p] > best) { best = t[p]; bkey = key[p]; }
        if (t[p + 1] > best) { best = t[p + 1]; bkey = key[p + 1]; }
      }
    } else {
      t[0] = fminf(sqdist_ref(x[0] - cx, y[0] - cy, z[0] - cz), t[0]);
      if (t[0] > best) { best = t[0]; bkey = key[0]; }
    }
    const int vb = __float_as_int(best);
    const int wv = __reduce_max_sync(0xFFFFFFFFu, vb);
    const uint32_t wk = __reduce_min_sync(0xFFFFFFFFu, vb == wv ? bkey : kNoKey);
    if (lane < (int)C) st_async_v2(par ? rdst[1] : rdst[0], (uint32_t)wv, wk, par ? rbars[1] : rbars[0]);

    mbar_wait(bar, (uint32_t)(it >> 1) & 1u);
    int bv = INT_MIN;
    uint32_t bk = kNoKey;
#pragma unroll
    for (int s = 0; s < kFlatMaxRecs / 32; ++s) {
      const int ri = s * 32 + lane;
      if (ri < R) {
        const uint2 rc = s_rec[par][ri];
        const int v = (int)rc.x;
        if (v > bv || (v == bv && rc.y < bk)) { bv = v; bk = rc.y; }
      }
    }
    const int gv = __reduce_max_sync(0xFFFFFFFFu, bv);
    const uint32_t win_key = __reduce_min_sync(0xFFFFFFFFu, bv == gv ? bk : kNoKey);
    const int wi = (int)fps_unkey(win_key, L);
    cx = s_x[wi]; cy = s_y[wi]; cz = s_z[wi];
    if (g == 0) {
      idx[it + 1] = wi;
      if (new_xyz) {
        new_xyz[(size_t)(it + 1) * 3 + 0] = cx;
        new_xyz[(size_t)(it + 1) * 3 + 1] = cy;
        new_xyz[(size_t)(it + 1) * 3 + 2] = cz;
      }
    }
  }

  if (temp) {
#pragma unroll
    for (int p = 0; p < P; ++p)
      if (key[p] != kNoKey) temp[fps_unkey(key[p], L)] = t[p];
  }
  cluster_sync_all();  // nobody leaves while a peer could still post to it
}

// Any-size fallback: one CTA per cloud, running distances in global memory (L2-resident).
__global__ void __launch_bounds__(1024, 1) fps_generic_kernel(FpsParams prm) {
  __shared__ int2 s_wrec[32];
  __shared__ int s_win[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = blockDim.x, nwarps = T >> 5;
  const int n = prm.n, m = prm.m, L = prm.L;
  const size_t cloud = blockIdx.x;
  const float *xyz = prm.xyz + cloud * (size_t)n * 3;
  float *temp = prm.temp + cloud * (size_t)n;  // required here
  int *idx = prm.idx + cloud * (size_t)m;
  float *new_xyz = prm.new_xyz ? prm.new_xyz + cloud * (size_t)m * 3 : nullptr;
  int old = 0;
  if (tid == 0 && m > 0) idx[0] = 0;
  for (int j = 1; j < m; ++j) {
    const float cx = xyz[(size_t)old * 3], cy = xyz[(size_t)old * 3 + 1], cz = xyz[(size_t)old * 3 + 2];
    if (tid == 0 && new_xyz) {
      new_xyz[(size_t)(j - 1) * 3] = cx; new_xyz[(size_t)(j - 1) * 3 + 1] = cy; new_xyz[(size_t)(j - 1) * 3 + 2] = cz;
    }
    float best = -1.f;
    uint32_t bkey = kNoKey;
    for (int k = tid; k < n; k += T) {
      const float d = sqdist_ref(xyz[(size_t)k * 3] - cx, xyz[(size_t)k * 3 + 1] - cy, xyz[(size_t)k * 3 + 2] - cz);
      const float d2 = fminf(d, temp[k]);
      temp[k] = d2;
      const uint32_t kk = fps_key((uint32_t)k, L);
      if (d2 > best || (d2 == best && kk < bkey)) { best = d2; bkey = kk; }
    }
    const int vb = __float_as_int(best);
    const int wv = __reduce_max_sync(0xFFFFFFFFu, vb);
    const uint32_t wk = __reduce_min_sync(0xFFFFFFFFu, vb == wv ? bkey : kNoKey);
    if (lane == 0) s_wrec[warp] = make_int2(wv, (int)wk);
    __syncthreads();
    if (warp == 0) {
      int2 r = lane < nwarps ? s_wrec[lane] : make_int2(INT_MIN, (int)kNoKey);
      const int bv = __reduce_max_sync(0xFFFFFFFFu, r.x);
      const uint32_t bk = __reduce_min_sync(0xFFFFFFFFu, r.x == bv ? (uint32_t)r.y : kNoKey);
      if (lane == 0) s_win[j & 1] = (int)fps_unkey(bk, L);
    }
    __syncthreads();
    old = s_win[j & 1];
    if (tid == 0) idx[j] = old;
  }
  if (tid == 0 && new_xyz && m > 0) {
    new_xyz[(size_t)(m - 1) * 3] = xyz[(size_t)old * 3];
    new_xyz[(size_t)(m - 1) * 3 + 1] = xyz[(size_t)old * 3 + 1];
    new_xyz[(size_t)(m - 1) * 3 + 2] = xyz[(size_t)old * 3 + 2];
  }
}

__global__ void fill_kernel(float *p, size_t n, float v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// pointnet2_lib/pointnet2/src/cuda_utils.h:10-14 -- same expression, same libm.
int ref_block_size(int work_size) {
  const int pow_2 = (int)(log((double)work_size) / log(2.0));
  int t = 1 << pow_2;
  if (t > 1024) t = 1024;
  if (t < 1) t = 1;
  return t;
}
int ilog2(int v) { int l = 0; while ((1 << (l + 1)) <= v) ++l; return l; }
int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }
int env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

template <int P, bool CLUSTER>
int launch_cluster(const FpsParams &prm, int b, int C, int T, cudaStream_t stream) {
  auto kern = fps_cluster_kernel<P, CLUSTER>;
  const size_t smem = (size_t)P * T * sizeof(float4);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess && C > 8) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (e != cudaSuccess) { set_error("fps: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)C, (unsigned)b, 1);
  cfg.blockDim = dim3((unsigned)T, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CLUSTER ? 1 : 0;
  e = cudaLaunchKernelEx(&cfg, kern, prm);
  if (e != cudaSuccess) { set_error("fps: launch (P=%d C=%d T=%d): %s", P, C, T, cudaGetErrorString(e)); return (int)e; }
  return check_launch("furthest_point_sampling");
}

template <int P, int MAXT = 512>
int launch_flat(const FpsParams &prm, int b, int C, int T, cudaStream_t stream) {
  auto kern = fps_flat_kernel<P, MAXT>;
  const size_t smem = (size_t)prm.n * 3 * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess && C > 8) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (e != cudaSuccess) { set_error("fps: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)C, (unsigned)b, 1);
  cfg.blockDim = dim3((unsigned)T, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, prm);
  if (e != cudaSuccess) { set_error("fps: launch (flat P=%d C=%d T=%d): %s", P, C, T, cudaGetErrorString(e)); return (int)e; }
  return check_launch("furthest_point_sampling");
}

int fps_dispatch(int b, int n, int m, const float *xyz, float *temp, int *idx, float *new_xyz, cudaStream_t stream) {
  if (b < 0 || n < 0 || m < 0 || (b > 0 && n > 0 && m > 0 && (!xyz || !idx))) return fail_arg("furthest_point_sampling");
  if (b == 0 || m == 0) return 0;
  if (n == 0) return fail_arg("furthest_point_sampling (n == 0 with m > 0)");
  if (b > 65535) {  // gridDim.y limit: split the batch
    for (int b0 = 0; b0 < b; b0 += 32768) {
      const int bb = (b - b0 < 32768) ? b - b0 : 32768;
      int rc = fps_dispatch(bb, n, m, xyz + (size_t)b0 * n * 3, temp ? temp + (size_t)b0 * n : nullptr,
                            idx + (size_t)b0 * m, new_xyz ? new_xyz + (size_t)b0 * m * 3 : nullptr, stream);
      if (rc) return rc;
    }
    return 0;
  }
  FpsParams prm;
  prm.n = n; prm.m = m; prm.L = ilog2(ref_block_size(n));
  prm.xyz = xyz; prm.temp = temp; prm.idx = idx; prm.new_xyz = new_xyz;
  prm.log2T = 0;
  // large clouds: spatially bucketed kernel, one CTA per cloud (fps_bucket.cu)
  if (fps_bucket_applicable(b, n, m)) return fps_bucket_launch(prm, b, stream);

  // --- decomposition: C CTAs per cloud, T threads per CTA, P points per thread
  // (tuning overrides are read from the environment ONCE: getenv walks the whole environment block)
  static const int env_allow16 = env_int("WS3D_FPS_ALLOW16", 0), env_c = env_int("WS3D_FPS_C", 0),
                   env_ppt = env_int("WS3D_FPS_PPT", 0), env_t = env_int("WS3D_FPS_T", 0);
  int cmax = num_sms() / (b < 1 ? 1 : b);
  cmax = cmax >= 16 ? 16 : cmax >= 8 ? 8 : cmax >= 4 ? 4 : cmax >= 2 ? 2 : 1;
  if (cmax > 8) cmax = env_allow16 ? 16 : 8;
  // measured on B200 (profiles/r1_op_bench_v2.json): one CTA (T=512, P=8) beats any cluster up to 4096 points;
  // above that a cluster of 8 x 512 threads x 4 points is best when the SMs are there
  int C = n <= 4096 ? 1 : pow2_ceil(ceil_div(n, 2048));
  if (C > cmax) C = cmax;
  while (ceil_div(n, C) > 8 * 1024 && C < 8) C <<= 1;  // registers: P <= 8 at T = 1024
  if (env_c > 0) C = env_c;
  int npc = ceil_div(n, C);
  int T = pow2_ceil(ceil_div(npc, env_ppt > 0 ? env_ppt : (npc > 2048 ? 8 : 4)));
  if (T < 32) T = 32;
  if (T > 1024) T = 1024;
  if (env_t > 0) T = env_t;
  int P = pow2_ceil(ceil_div(npc, T));
  prm.log2T = ilog2(T);

  const bool ok = (C >= 1 && C <= kMaxCluster && (C & (C - 1)) == 0 && T >= 32 && T <= 1024 && (T & (T - 1)) == 0 &&
                   P <= 8 && (long long)P * T * C >= n);
  {
    // Clouds of 2048+ points that fit every CTA's shared memory: one-level exchange (fps_flat_kernel) on a
    // cluster of 4 CTAs x 16 (n > 8192) or 8 points per thread (n = 4096: 0.40 ms against 0.50 ms on one CTA).  Measured at b = 16, n = 16384 (profiles/r1_fps_bench_v4.json):
    // 2.09 ms against 2.76 ms for the two-level kernel on 8 CTAs; more CTAs or more warps per cluster lose to
    // the DSMEM traffic of the all-to-all post (C = 8, T = 512: 5.2 ms), fewer to the per-thread update.
    static const int flat = env_int("WS3D_FPS_FLAT", 1);
    static const int flat_min = env_int("WS3D_FPS_FLAT_MIN", 2048);
    if (flat && n >= flat_min && (size_t)n * 12 <= 200 * 1024 && cmax >= 2) {
      int Cf = env_c > 0 ? env_c : (cmax >= 4 ? 4 : 2);
      int Tf = pow2_ceil(ceil_div(ceil_div(n, Cf), env_ppt > 0 ? env_ppt : (n > 8192 ? 16 : 8)));
      if (Tf < 32) Tf = 32;
      if (Tf > 512) Tf = 512;
      const int Pf = pow2_ceil(ceil_div(ceil_div(n, Cf), Tf));
      if (Cf > 1 && Cf <= kMaxCluster && (Cf & (Cf - 1)) == 0 && Pf <= 16 && Cf * (Tf / 32) <= kFlatMaxRecs &&
          (long long)Pf * Tf * Cf >= n) {
        prm.log2T = ilog2(Tf);
        switch (Pf) {
          case 1: return launch_flat<1>(prm, b, Cf, Tf, stream);
          case 2: return launch_flat<2>(prm, b, Cf, Tf, stream);
          case 4: return launch_flat<4>(prm, b, Cf, Tf, stream);
          case 8: return launch_flat<8>(prm, b, Cf, Tf, stream);
          case 16: return launch_flat<16>(prm, b, Cf, Tf, stream);
        }
      }
    }
  }
  if (!ok) {
    // generic path needs the scratch array
    float *tp = temp;
    if (!tp) {
      tp = (float *)scratch((size_t)b * n * sizeof(float), 0);
      if (!tp) return (int)cudaErrorMemoryAllocation;
      const size_t tot = (size_t)b * n;
      fill_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(tp, tot, 1e10f);
      int rc = check_launch("fps fill");
      if (rc) return rc;
    }
    prm.temp = tp;
    fps_generic_kernel<<<b, 1024, 0, stream>>>(prm);
    return check_launch("furthest_point_sampling (generic)");
  }
#define WS3D_FPS_CASE(PP)                                                          \
  case PP:                                                                         \
    return C > 1 ? launch_cluster<PP, true>(prm, b, C, T, stream)                  \
                 : launch_cluster<PP, false>(prm, b, C, T, stream);
  switch (P) {
    WS3D_FPS_CASE(1)
    WS3D_FPS_CASE(2)
    WS3D_FPS_CASE(4)
    WS3D_FPS_CASE(8)
  }
#undef WS3D_FPS_CASE
  return fail_arg("furthest_point_sampling (no kernel for P)");
}

}  // namespace
}  // namespace ws3d

WS3D_API int ws3d_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx,
                                          ws3d_stream_t stream) {
  return ws3d::fps_dispatch(b, n, m, xyz, temp, idx, nullptr, ws3d::to_stream(stream));
}

WS3D_API int ws3d_furthest_point_sampling_gather(int b, int n, int m, const float *xyz, float *temp, int *idx,
                                                 float *new_xyz, ws3d_stream_t stream) {
  return ws3d::fps_dispatch(b, n, m, xyz, temp, idx, new_xyz, ws3d::to_stream(stream));
}
