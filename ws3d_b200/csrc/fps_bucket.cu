// Furthest point sampling with spatial buckets: one CTA per cloud, exact.
//
// The register/cluster kernel of fps.cu updates every point's running distance in every one of the
// m-1 iterations (16384 x 4095 = 67 M distance evaluations per cloud for SA1).  Almost all of them are
// no-ops: a point's running distance t can only drop if the new sample is closer than sqrt(t), and
// every t is <= T_max, the running distance of the sample just selected (it was the maximum).  So only
// points within sqrt(T_max) of the new sample can change -- a handful once a few hundred samples exist.
//
// This kernel sorts the cloud along a Morton curve once (cub::BlockRadixSort in shared memory), cuts
// the sorted order into buckets of 32 consecutive points (one point per lane) with a bounding box each,
// and per iteration
//   1. every lane tests one bucket of its warp: box-to-sample distance^2 (shrunk by 1e-4 to stay
//      conservative under rounding) >= T_max  ->  nothing in the bucket can change;
//   2. the warp updates only its active buckets (coordinates from shared memory, running distances in
//      registers) with exactly the reference arithmetic, so the skipped updates are provably no-ops;
//   3. lanes whose own maximum was lowered rescan their registers; warps that changed redo their
//      arg-max (two redux.sync), the rest re-post their cached candidate; one __syncthreads; every
//      warp picks the winner from the <= 16 warp candidates.
// Result: bit-identical sample order (same values, same tie-break key as fps.cu), ~3x fewer cycles per
// iteration, and one SM per cloud instead of a cluster of 8.
//
// Layout: sorted rank r = j*T + tid lives in lane (tid & 31) of warp (tid >> 5), register t[j];
// bucket (warp, j) = ranks j*T + 32*warp .. +31; lane j of the warp holds that bucket's box.
#include <cub/block/block_radix_sort.cuh>

#include <limits.h>
#include <stdlib.h>

#include "common.cuh"
#include "fps_common.cuh"

namespace ws3d {
namespace {

__device__ unsigned long long g_stats[8];  // debug counters (WS3D_FPS_STATS=1)

constexpr float kCullShrink = 0.9999f;  // >> the 4 ulp the box distance and the point distance can differ by

struct __align__(16) WarpRec {  // a warp's candidate
  int v;                        // running-distance bits
  uint32_t key;
  float x, y, z;
  float pad[3];
};

// order-preserving float <-> int map (finite values), for integer redux.sync min/max
__device__ __forceinline__ int f2ord(float f) {
  const int i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7FFFFFFF);
}
__device__ __forceinline__ float ord2f(int o) { return __int_as_float(o ^ ((o >> 31) & 0x7FFFFFFF)); }

__device__ __forceinline__ uint32_t spread3(uint32_t v) {  // 6 bits -> every third bit
  v &= 0x3Fu;
  v = (v | (v << 8)) & 0x300Fu;
  v = (v | (v << 4)) & 0x30C3u;
  v = (v | (v << 2)) & 0x9249u;
  return v;
}

template <int T, int P>
__global__ void __launch_bounds__(T, 1) fps_bucket_kernel(FpsParams prm) {
  static_assert(P <= 32 && T % 32 == 0 && T <= 1024, "one bucket per lane");
  using Sort = cub::BlockRadixSort<uint32_t, T, P>;
  constexpr int kCap = T * P;
  constexpr int kWarps = T / 32;
  static_assert(sizeof(typename Sort::TempStorage) <= (size_t)kCap * 14, "sort scratch must fit in the point arrays");

  extern __shared__ __align__(16) unsigned char s_raw[];
  float *s_x = reinterpret_cast<float *>(s_raw);
  float *s_y = s_x + kCap;
  float *s_z = s_y + kCap;
  unsigned short *s_k = reinterpret_cast<unsigned short *>(s_z + kCap);  // original index of each sorted rank
  __shared__ WarpRec s_rec[2][kWarps];
  __shared__ int s_box[6][kWarps];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = prm.n, m = prm.m, L = prm.L;
  const size_t cloud = blockIdx.x;
  const float *xyz = prm.xyz + cloud * (size_t)n * 3;
  float *temp = prm.temp ? prm.temp + cloud * (size_t)n : nullptr;
  int *idx = prm.idx + cloud * (size_t)m;
  float *new_xyz = prm.new_xyz ? prm.new_xyz + cloud * (size_t)m * 3 : nullptr;

  // ---- setup 1: bounding box of the finite points -> Morton quantisation
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  for (int k = tid; k < n; k += T) {
    const float x = __ldg(xyz + (size_t)k * 3), y = __ldg(xyz + (size_t)k * 3 + 1), z = __ldg(xyz + (size_t)k * 3 + 2);
    if (isfinite(x) && isfinite(y) && isfinite(z)) {
      lo[0] = min(lo[0], f2ord(x)); hi[0] = max(hi[0], f2ord(x));
      lo[1] = min(lo[1], f2ord(y)); hi[1] = max(hi[1], f2ord(y));
      lo[2] = min(lo[2], f2ord(z)); hi[2] = max(hi[2], f2ord(z));
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = __reduce_min_sync(0xFFFFFFFFu, lo[a]);
    hi[a] = __reduce_max_sync(0xFFFFFFFFu, hi[a]);
    if (lane == 0) { s_box[a][warp] = lo[a]; s_box[3 + a][warp] = hi[a]; }
  }
  __syncthreads();
  float org[3], inv_cell;
  {
    float ext = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      int l = lane < kWarps ? s_box[a][lane] : INT_MAX, h = lane < kWarps ? s_box[3 + a][lane] : INT_MIN;
      l = __reduce_min_sync(0xFFFFFFFFu, l);
      h = __reduce_max_sync(0xFFFFFFFFu, h);
      org[a] = l <= h ? ord2f(l) : 0.f;
      ext = fmaxf(ext, l <= h ? ord2f(h) - ord2f(l) : 0.f);
    }
    inv_cell = (ext > 0.f && isfinite(ext)) ? 64.f / ext : 0.f;
    if (!isfinite(inv_cell)) inv_cell = 0.f;
  }

  // ---- setup 2: sort (morton << 14 | original index); padded slots sort last
  uint32_t keys[P];
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int k = j * T + tid;
    uint32_t key = 0xFFFFFFFFu;
    if (k < n) {
      const float x = __ldg(xyz + (size_t)k * 3), y = __ldg(xyz + (size_t)k * 3 + 1), z = __ldg(xyz + (size_t)k * 3 + 2);
      uint32_t mort = 0x3FFFFu;
      if (isfinite(x) && isfinite(y) && isfinite(z)) {
        const uint32_t qx = (uint32_t)fminf(fmaxf((x - org[0]) * inv_cell, 0.f), 63.f);
        const uint32_t qy = (uint32_t)fminf(fmaxf((y - org[1]) * inv_cell, 0.f), 63.f);
        const uint32_t qz = (uint32_t)fminf(fmaxf((z - org[2]) * inv_cell, 0.f), 63.f);
        mort = spread3(qx) | (spread3(qz) << 1) | (spread3(qy) << 2);
      }
      // the all-ones code is the padding's: a real point never shares it (the sort looks at the code only, so padding must sort
      // strictly last; and (code << 14 | k) of a real point must never equal the padding key 0xFFFFFFFF)
      if (mort == 0x3FFFFu) mort = 0x3FFFEu;
      key = (mort << 14) | (uint32_t)k;
    }
    keys[j] = key;
  }
  __syncthreads();
  Sort(*reinterpret_cast<typename Sort::TempStorage *>(s_raw)).SortBlockedToStriped(keys, 14, 32);
  __syncthreads();  // the sort scratch becomes the point arrays

  // ---- setup 3: load the points in sorted order; per-bucket boxes
  float t[P];
  float bx0 = 0.f, by0 = 0.f, bz0 = 0.f, bx1 = 0.f, by1 = 0.f, bz1 = 0.f;  // box of bucket (warp, lane)
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int r = j * T + tid;
    float x = 0.f, y = 0.f, z = 0.f;
    int k = 0;
    bool real = keys[j] != 0xFFFFFFFFu;
    if (real) {
      k = (int)(keys[j] & 0x3FFFu);
      x = __ldg(xyz + (size_t)k * 3); y = __ldg(xyz + (size_t)k * 3 + 1); z = __ldg(xyz + (size_t)k * 3 + 2);
      t[j] = temp ? temp[k] : 1e10f;
    } else {
      t[j] = -1.f;  // never beats a real point (real running distances are >= 0)
    }
    s_x[r] = x; s_y[r] = y; s_z[r] = z; s_k[r] = real ? (unsigned short)k : (unsigned short)0xFFFFu;
    const bool fin = real && isfinite(x) && isfinite(y) && isfinite(z);  // others can never change: not in the box
    const int l0 = __reduce_min_sync(0xFFFFFFFFu, fin ? f2ord(x) : INT_MAX), h0 = __reduce_max_sync(0xFFFFFFFFu, fin ? f2ord(x) : INT_MIN);
    const int l1 = __reduce_min_sync(0xFFFFFFFFu, fin ? f2ord(y) : INT_MAX), h1 = __reduce_max_sync(0xFFFFFFFFu, fin ? f2ord(y) : INT_MIN);
    const int l2 = __reduce_min_sync(0xFFFFFFFFu, fin ? f2ord(z) : INT_MAX), h2 = __reduce_max_sync(0xFFFFFFFFu, fin ? f2ord(z) : INT_MIN);
    if (lane == j) {
      const float kInf = __int_as_float(0x7f800000);
      const bool any = l0 <= h0;
      bx0 = any ? ord2f(l0) : kInf; bx1 = any ? ord2f(h0) : -kInf;
      by0 = any ? ord2f(l1) : kInf; by1 = any ? ord2f(h1) : -kInf;
      bz0 = any ? ord2f(l2) : kInf; bz1 = any ? ord2f(h2) : -kInf;
    }
  }

  float cx = __ldg(xyz + 0), cy = __ldg(xyz + 1), cz = __ldg(xyz + 2);  // idx[0] = 0
  if (tid == 0 && m > 0) {
    idx[0] = 0;
    if (new_xyz) { new_xyz[0] = cx; new_xyz[1] = cy; new_xyz[2] = cz; }
  }
  __syncthreads();

  float t_max = __int_as_float(0x7f800000);  // upper bound of every running distance
  float bm = -2.f;          // this lane's maximum, its tie key and register slot
  uint32_t bkey = kNoKey;
  int bpos = 0;
  bool dirty = true;
  WarpRec mine;             // the warp's cached candidate (valid in every lane)
  mine.v = INT_MIN; mine.key = kNoKey; mine.x = mine.y = mine.z = 0.f;
  bool warp_stale = true;

  for (int it = 0; it + 1 < m; ++it) {
    const int par = it & 1;
    // ---- 1. which of this warp's buckets can change?  (NaN sample -> comparison false -> active)
    bool act = false;
    if (lane < P) {
      const float ax = fmaxf(fmaxf(bx0 - cx, cx - bx1), 0.f);
      const float ay = fmaxf(fmaxf(by0 - cy, cy - by1), 0.f);
      const float az = fmaxf(fmaxf(bz0 - cz, cz - bz1), 0.f);
      const float lb = ax * ax + ay * ay + az * az;
      act = !(lb * kCullShrink >= t_max);
    }
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, act);
    if (prm.log2T == 777 && lane == 0) {
      atomicAdd(&g_stats[0], (unsigned long long)__popc(mask));
      atomicAdd(&g_stats[1], mask ? 1ull : 0ull);
      atomicAdd(&g_stats[2], 1ull);
      if (it >= m / 2) { atomicAdd(&g_stats[4], (unsigned long long)__popc(mask)); atomicAdd(&g_stats[5], 1ull); }
    }
    // ---- 2. exact update of the active buckets.  The running distances live in registers and the bucket index
    //         is only known at run time, so the register is read through a select tree and written back with
    //         predicated moves -- branch-free, because the warp that has work is the iteration's critical path.
    if (mask) {
      bool changed = false;
      uint32_t mm = mask;
      while (mm) {
        const int j = __ffs(mm) - 1;
        mm &= mm - 1;
        const int r = j * T + tid;
        const float d = sqdist_ref(s_x[r] - cx, s_y[r] - cy, s_z[r] - cz);
        float sel[P];
#pragma unroll
        for (int q = 0; q < P; ++q) sel[q] = t[q];
#pragma unroll
        for (int w = 1; w < P; w <<= 1) {
          const bool hi_half = (j & w) != 0;
#pragma unroll
          for (int q = 0; q + w < P; q += 2 * w) sel[q] = hi_half ? sel[q + w] : sel[q];
        }
        const float old = sel[0];
        const float nt = fminf(d, old);
        if (nt != old) {
          changed = true;
          dirty = dirty || (j == bpos);
        }
#pragma unroll
        for (int q = 0; q < P; ++q) t[q] = (q == j) ? nt : t[q];
      }
      warp_stale = warp_stale || __any_sync(0xFFFFFFFFu, changed);
    }
    // ---- 3. lanes whose maximum was lowered rescan their registers: max tree, then the set of registers that
    //         hold the maximum as a bit mask (OR tree), then the tie key of the (usually single) holder
    if (prm.log2T == 777) { const uint32_t dm = __ballot_sync(0xFFFFFFFFu, dirty); if (lane == 0) atomicAdd(&g_stats[3], (unsigned long long)__popc(dm)); }
    if (dirty) {
      float v[P];
#pragma unroll
      for (int j = 0; j < P; ++j) v[j] = t[j];
#pragma unroll
      for (int w = 1; w < P; w <<= 1) {
#pragma unroll
        for (int j = 0; j + w < P; j += 2 * w) v[j] = fmaxf(v[j], v[j + w]);
      }
      bm = v[0];
      uint32_t e[P];
#pragma unroll
      for (int j = 0; j < P; ++j) e[j] = (t[j] == bm) ? (1u << j) : 0u;
#pragma unroll
      for (int w = 1; w < P; w <<= 1) {
#pragma unroll
        for (int j = 0; j + w < P; j += 2 * w) e[j] |= e[j + w];
      }
      uint32_t eq = e[0];
      bkey = kNoKey;
      if (bm >= 0.f) {
        while (eq) {
          const int j = __ffs(eq) - 1;
          eq &= eq - 1;
          const uint32_t kk = fps_key((uint32_t)s_k[j * T + tid], L);
          if (kk < bkey) { bkey = kk; bpos = j; }
        }
      }
      dirty = false;
    }
    // ---- 4. warp candidate (recomputed only if something in the warp changed), one barrier, CTA winner
    if (warp_stale) {
      const int vb = __float_as_int(bm);
      const int wv = __reduce_max_sync(0xFFFFFFFFu, vb);
      const uint32_t wk = __reduce_min_sync(0xFFFFFFFFu, vb == wv ? bkey : kNoKey);
      const int src = __ffs(__ballot_sync(0xFFFFFFFFu, vb == wv && bkey == wk)) - 1;
      float px = 0.f, py = 0.f, pz = 0.f;
      if (lane == src) {
        const int r = bpos * T + tid;
        px = s_x[r]; py = s_y[r]; pz = s_z[r];
      }
      mine.v = wv; mine.key = wk;
      mine.x = __shfl_sync(0xFFFFFFFFu, px, src);
      mine.y = __shfl_sync(0xFFFFFFFFu, py, src);
      mine.z = __shfl_sync(0xFFFFFFFFu, pz, src);
      warp_stale = false;
    }
    if (lane == 0) s_rec[par][warp] = mine;
    __syncthreads();
    {
      int v = INT_MIN;
      uint32_t kk = kNoKey;
      float rx = 0.f, ry = 0.f, rz = 0.f;
      if (lane < kWarps) {
        const WarpRec &rc = s_rec[par][lane];
        v = rc.v; kk = rc.key; rx = rc.x; ry = rc.y; rz = rc.z;
      }
      const int bv = __reduce_max_sync(0xFFFFFFFFu, v);
      const uint32_t win_key = __reduce_min_sync(0xFFFFFFFFu, v == bv ? kk : kNoKey);
      const int src = __ffs(__ballot_sync(0xFFFFFFFFu, v == bv && kk == win_key)) - 1;
      cx = __shfl_sync(0xFFFFFFFFu, rx, src);
      cy = __shfl_sync(0xFFFFFFFFu, ry, src);
      cz = __shfl_sync(0xFFFFFFFFu, rz, src);
      t_max = __int_as_float(bv);
      if (tid == 0) {
        idx[it + 1] = (int)fps_unkey(win_key, L);
        if (new_xyz) {
          new_xyz[(size_t)(it + 1) * 3 + 0] = cx;
          new_xyz[(size_t)(it + 1) * 3 + 1] = cy;
          new_xyz[(size_t)(it + 1) * 3 + 2] = cz;
        }
      }
    }
  }

  if (temp) {
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const unsigned short k = s_k[j * T + tid];
      if (k != 0xFFFFu) temp[k] = t[j];
    }
  }
}

int env_int2(const char *name, int dflt) {
  const char *s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

template <int T, int P>
int launch_bucket(const FpsParams &prm, int b, cudaStream_t stream) {
  auto kern = fps_bucket_kernel<T, P>;
  const size_t smem = (size_t)T * P * 14;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("fps (bucket): smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  static const int stats = env_int2("WS3D_FPS_STATS", 0);
  if (stats) {
    FpsParams q = prm;
    q.log2T = 777;
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_stats, z, sizeof(z));
    kern<<<b, T, smem, stream>>>(q);
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(z, g_stats, sizeof(z));
    fprintf(stderr, "[fps stats] T=%d P=%d b=%d n=%d m=%d: warp-iterations %llu, active buckets/warp-it %.3f, warps active %.3f, "
                    "dirty lanes/warp-it %.3f, second half: active buckets/warp-it %.3f\n", T, P, b, prm.n, prm.m, z[2],
            (double)z[0] / z[2], (double)z[1] / z[2], (double)z[3] / z[2], z[5] ? (double)z[4] / z[5] : 0.0);
    return check_launch("furthest_point_sampling (bucket)");
  }
  kern<<<b, T, smem, stream>>>(prm);
  return check_launch("furthest_point_sampling (bucket)");
}

}  // namespace

// Large clouds only: below 2048 points the whole cloud is a few buckets and the register kernels win;
// 16384 points is what fits one SM's shared memory (coordinates + 16-bit indices = 224 KB).
// It needs ONE SM per cloud, so it is the choice when the batch leaves fewer than 2 SMs per cloud for the
// cluster kernels (b > 74; measured on B200 at b = 64 x 16384 points: 3.55 ms against 4.39 ms for the two-level
// cluster kernel, 3.06 ms for the flat one on 2 SMs per cloud); WS3D_FPS_BUCKET=1/0 forces it.
bool fps_bucket_applicable(int b, int n, int m) {
  static const int env_mode = env_int2("WS3D_FPS_BUCKET", -1);
  // ws3d_set_fps_mode(): 1 = throughput (one SM per cloud whenever this kernel applies), 2 = latency (never)
  const int rt = fps_mode();
  const int mode = rt == 1 ? 1 : rt == 2 ? 0 : env_mode;
  if (mode == 0 || n < 2048 || n > 16384 || m < 64 || b < 1 || b > 65535) return false;
  if (mode == 1) return true;
  // automatic: against the cluster kernels of fps.cu measured at 4 / 2 / 1 SMs per cloud (profiles/r2_fps_bench.json, 16384 -> 4096:
  // 2.11 / 2.23 / 3.09 ms for b = 16 / 32 / 64 against 2.15 ms for the paired-sample kernel on ONE SM per cloud; 4096 -> 1024:
  // 0.40 ms at b = 16 and 0.50 at b = 96 against 0.54) -- large clouds switch as soon as a cloud gets fewer than 4 SMs, small
  // ones only when there are more clouds than SMs
  if (n > 8192) return b * 4 > num_sms();
  if (n > 4096) return b * 2 > num_sms();
  return b > num_sms();
}

int fps_bucket_launch(const FpsParams &prm, int b, cudaStream_t stream) {
  const int n = prm.n;
  // default: the shared-memory-distance kernel (fps_smem.cu: several clouds per SM); WS3D_FPS_SMEM=0 keeps the kernel of this file
  static const int use_smem = env_int2("WS3D_FPS_SMEM", 1);
  if (use_smem) return fps_smem_launch(prm, b, stream);
  // buckets per warp: 16 with twice the warps up to 8192 points (measured at b = 16, 4096 -> 1024: 0.70 ms against 0.84:
  // half the select tree and half the rescan), 32 at 16384 points, where 1024 threads would cap at 64 registers and spill (4.3 ms against 3.5)
  if (n <= 4096) return launch_bucket<256, 16>(prm, b, stream);
  if (n <= 8192) return launch_bucket<512, 16>(prm, b, stream);
  return launch_bucket<512, 32>(prm, b, stream);
}

}  // namespace ws3d
