// Furthest point sampling for THROUGHPUT: several clouds per SM, exact.
//
// The bucketed kernel of fps_bucket.cu keeps a cloud's coordinates in shared memory (196 KB at 16384 points) and its running
// distances in registers (a quarter of the register file), so a cloud owns a whole SM for the 3.5 ms of its latency chain while
// issuing on a third of the cycles -- a third of the machine's SM-time in the streamed pipeline.  This kernel
// turns the storage round: the running distances (4 B per point) live in SHARED memory, the Morton-sorted coordinates
// (16 B per point, with the original index in the fourth word) in a scratch array that stays in L2 and is read only for the
// buckets a new sample can still lower (6.7 of 512 per iteration at 16384 -> 4096), and every lane owns whole BUCKETS (box,
// maximum, tie key: 8 registers; the coordinates of the maximum in shared memory) instead of one point of every bucket.
// Consequences:
//   * no register-indexed state: an active bucket is updated by the 32 lanes of its warp (one point each) and its maximum is
//     recomputed on the spot with two redux.sync -- no select tree, no dirty-lane rescan;
//   * 72 KB of shared memory and 59 registers per thread for a 16384-point cloud: TWO clouds share an SM (eight at 4096 points)
//     as independent sub-blocks of one CTA and cover each other's stalls, so a cloud costs half an SM for the length of its
//     chain instead of a whole one (a third cloud would need the state under 40 registers: measured, spills).
// Same culling rule (tightened to the bucket's own maximum), same update arithmetic and same tie key as fps_bucket.cu / fps.cu:
// bit-identical sample order.  DESIGN.md section 3.1c has the measurements.
//
// Two launches: fps_sort_kernel (one CTA per cloud: Morton sort in shared memory, writes the sorted float4 array) and
// fps_smem_kernel (the m - 1 iterations).
#include <cub/block/block_radix_sort.cuh>

#include <limits.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "fps_common.cuh"

namespace ws3d {
namespace {

constexpr float kCullShrink = 0.9999f;  // >> the 4 ulp the box distance and the point distance can differ by
constexpr int kIterThreads = 512;       // 16 warps; lane l of warp w owns bucket l * 16 + w: a new sample's neighbourhood is a few runs of
                                        // CONSECUTIVE Morton buckets, which this interleaving hands to different warps (one L2 round each)
constexpr int kIterWarps = kIterThreads / 32;

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

// ---- launch 1: Morton order of a cloud -> sorted (x, y, z, original index) in global memory; slots beyond n hold index -1
template <int T, int P>
__global__ void __launch_bounds__(T, 1) fps_sort_kernel(int n, const float *__restrict__ xyz_g, float4 *__restrict__ sorted_g) {
  using Sort = cub::BlockRadixSort<uint32_t, T, P>;
  constexpr int kWarps = T / 32;
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ int s_box[6][kWarps];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t cloud = blockIdx.x;
  const float *xyz = xyz_g + cloud * (size_t)n * 3;
  float4 *sorted = sorted_g + cloud * (size_t)(T * P);

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
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int r = j * T + tid;            // sorted rank
    float4 v = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    if (keys[j] != 0xFFFFFFFFu) {
      const int k = (int)(keys[j] & 0x3FFFu);
      v.x = __ldg(xyz + (size_t)k * 3); v.y = __ldg(xyz + (size_t)k * 3 + 1); v.z = __ldg(xyz + (size_t)k * 3 + 2);
      v.w = __int_as_float(k);
    }
    sorted[r] = v;
  }
}

struct __align__(16) WarpRec {   // a warp's candidate (value bits, tie key, coordinates), 32 bytes
  int v;
  uint32_t key;
  float x, y, z;
  int pad[3];
};

struct SmemFpsParams {
  int b, n, m, L, cap;         // clouds, points, samples, log2(reference block size), sorted slots per cloud
  int clouds_per_cta, cloud_smem;   // sub-blocks of a CTA (one cloud each) and the bytes of shared memory each one owns
  const float4 *sorted;        // (B, cap)
  float *temp;                 // (B, n) or null
  int *idx;                    // (B, m)
  float *new_xyz;              // (B, m, 3) or null
  const float *xyz;            // (B, n, 3): the first sample is point 0
};

__device__ __forceinline__ void sub_barrier(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// ---- launch 2: the iterations.  A CTA is cut into independent sub-blocks of kWarps warps, one cloud each (own named barrier, own
// slice of the shared memory), so that SEVERAL CLOUDS SHARE ONE SM and cover each other's latency chain -- the hardware would
// otherwise spread one-cloud CTAs over as many SMs as there are clouds.  Lane l of warp w owns bucket l * kWarps + w:
// a new sample's neighbourhood is a few runs of CONSECUTIVE Morton buckets, which this interleaving hands to different
// warps (one L2 round each).
// Shared memory of a sub-block: coordinates of every bucket's current maximum (16 B x nb: written by the lane that holds that
// point, read by its own warp only), running distance of every sorted slot (4 B x 32 nb), the warps' standing candidates and the
// two parities of the exchange records.
__device__ unsigned long long g_smem_stats[16];   // debug counters (WS3D_FPS_STATS=1): see fps_smem_launch

template <int kWarps, int kBpl, int kMaxThreads, bool kStats = false>
__global__ void __launch_bounds__(kMaxThreads, 1) fps_smem_kernel(SmemFpsParams prm) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  constexpr int kSub = kWarps * 32;
  const int sub = threadIdx.x / kSub, tid = threadIdx.x % kSub, lane = tid & 31, warp = tid >> 5;
  const int cloud_i = blockIdx.x * prm.clouds_per_cta + sub;
  if (sub >= prm.clouds_per_cta || cloud_i >= prm.b) return;        // whole sub-blocks leave: their barrier is their own
  const int n = prm.n, m = prm.m, L = prm.L;
  const int nb = (n + 31) >> 5;                       // buckets of 32 consecutive sorted slots
  unsigned char *base = s_dyn + (size_t)sub * prm.cloud_smem;
  float4 *s_m4 = reinterpret_cast<float4 *>(base);                              // [nb]
  float *s_t = reinterpret_cast<float *>(base + (size_t)nb * 16);               // [nb * 32], -1 for padding
  WarpRec *s_mine = reinterpret_cast<WarpRec *>(base + (size_t)nb * 144);       // [kWarps]   (private to each warp)
  WarpRec *s_rec = s_mine + kWarps;                                             // [2][kWarps]
  const size_t cloud = cloud_i;
  const float4 *pts = prm.sorted + cloud * (size_t)prm.cap;
  int *idx = prm.idx + cloud * (size_t)m;
  float *new_xyz = prm.new_xyz ? prm.new_xyz + cloud * (size_t)m * 3 : nullptr;
  auto bucket_of = [&](int j) { return j * kWarps + warp; };   // this warp's buckets, j < 32 kBpl: lane j & 31, slot j >> 5

  // per-lane state of the buckets bucket_of(lane + 32 i), i < kBpl (every index below is a compile-time constant after
  // unrolling: the arrays live in registers)
  float bx0[kBpl], by0[kBpl], bz0[kBpl], bx1[kBpl], by1[kBpl], bz1[kBpl];
  int bmax[kBpl];              // bits of the bucket's largest running distance (>= 0: int order == float order)
  uint32_t bkey[kBpl];         // tie key of that point
#pragma unroll
  for (int i = 0; i < kBpl; ++i) {
    bx0[i] = by0[i] = bz0[i] = __int_as_float(0x7f800000);   // empty box: never active
    bx1[i] = by1[i] = bz1[i] = __int_as_float(0xff800000);
    bmax[i] = INT_MIN; bkey[i] = kNoKey;
  }

  // (Re)compute the maximum of bucket b = bucket_of(j) from the warp's 32 values.
  auto bucket_max = [&](int j, int b, float t, const float4 &p) {
    const int orig = __float_as_int(p.w);
    const int vb = __float_as_int(t);
    const int wv = __reduce_max_sync(0xFFFFFFFFu, vb);
    const uint32_t kk = (vb == wv && orig >= 0) ? fps_key((uint32_t)orig, L) : kNoKey;
    const uint32_t wk = __reduce_min_sync(0xFFFFFFFFu, kk);
    if (kk == wk && kk != kNoKey) s_m4[b] = p;          // keys are unique: exactly one lane (none if the bucket is all padding)
#pragma unroll
    for (int i = 0; i < kBpl; ++i)
      if (lane == (j & 31) && i == (j >> 5)) { bmax[i] = wv; bkey[i] = wk; }
  };

  // ---- setup: running distances, boxes, bucket maxima
  {
    const float *temp = prm.temp ? prm.temp + cloud * (size_t)n : nullptr;
#pragma unroll
    for (int slot = 0; slot < kBpl; ++slot)
    for (int jl = 0; jl < 32; ++jl) {
      const int j = slot * 32 + jl;
      const int b = bucket_of(j);
      if (b >= nb) break;                               // warp-uniform (buckets grow with j: the later slots are empty too)
      const float4 p = pts[b * 32 + lane];
      const int orig = __float_as_int(p.w);
      const float t = orig >= 0 ? (temp ? temp[orig] : 1e10f) : -1.f;
      s_t[b * 32 + lane] = t;
      const bool fin = orig >= 0 && isfinite(p.x) && isfinite(p.y) && isfinite(p.z);  // others can never change: not in the box
      const int l0 = __reduce_min_sync(0xFFFFFFFFu, fin ? f2ord(p.x) : INT_MAX), h0 = __reduce_max_sync(0xFFFFFFFFu, fin ? f2ord(p.x) : INT_MIN);
      const int l1 = __reduce_min_sync(0xFFFFFFFFu, fin ? f2ord(p.y) : INT_MAX), h1 = __reduce_max_sync(0xFFFFFFFFu, fin ? f2ord(p.y) : INT_MIN);
      const int l2 = __reduce_min_sync(0xFFFFFFFFu, fin ? f2ord(p.z) : INT_MAX), h2 = __reduce_max_sync(0xFFFFFFFFu, fin ? f2ord(p.z) : INT_MIN);
#pragma unroll
      for (int i = 0; i < kBpl; ++i)
        if (lane == jl && i == slot && l0 <= h0) {
          bx0[i] = ord2f(l0); bx1[i] = ord2f(h0);
          by0[i] = ord2f(l1); by1[i] = ord2f(h1);
          bz0[i] = ord2f(l2); bz1[i] = ord2f(h2);
        }
      if (lane == 0) s_m4[b] = make_float4(0.f, 0.f, 0.f, 0.f);
      __syncwarp();
      bucket_max(j, b, t, p);
    }
  }
  float cx = __ldg(prm.xyz + cloud * (size_t)n * 3), cy = __ldg(prm.xyz + cloud * (size_t)n * 3 + 1),
        cz = __ldg(prm.xyz + cloud * (size_t)n * 3 + 2);       // idx[0] = 0
  if (tid == 0 && m > 0) {
    idx[0] = 0;
    if (new_xyz) { new_xyz[0] = cx; new_xyz[1] = cy; new_xyz[2] = cz; }
  }
  float t_max = __int_as_float(0x7f800000);           // upper bound of every running distance
  bool warp_stale = true;
  unsigned long long st_active = 0, st_crit = 0, st_hist[6] = {0, 0, 0, 0, 0, 0};   // kStats only

  for (int it = 0; it + 1 < m; ++it) {
    // ---- 1. which of this warp's buckets can still be lowered?  (NaN sample -> comparison false -> active)
    uint32_t mm[kBpl];
#pragma unroll
    for (int i = 0; i < kBpl; ++i) {
      const float ax = fmaxf(fmaxf(bx0[i] - cx, cx - bx1[i]), 0.f);
      const float ay = fmaxf(fmaxf(by0[i] - cy, cy - by1[i]), 0.f);
      const float az = fmaxf(fmaxf(bz0[i] - cz, cz - bz1[i]), 0.f);
      const float lb = ax * ax + ay * ay + az * az;
      // a point p of the bucket changes only if d(sample, p) < t_p <= the bucket's OWN largest running distance (which
      // this lane holds), a tighter radius than the global maximum t_max the bucket kernel of fps_bucket.cu tests against
      const bool act = bucket_of(lane + 32 * i) < nb && !(lb * kCullShrink >= fminf(t_max, __int_as_float(bmax[i])));
      mm[i] = __ballot_sync(0xFFFFFFFFu, act);
    }
    // ---- 2. exact update of the active buckets: one point per lane, coordinates from the L2-resident sorted copy.
    //         Two buckets are fetched per round so that their L2 latencies overlap.
    auto pop = [&]() -> int {                         // next active bucket slot of this warp, or -1 (warp-uniform)
      int j = -1;
#pragma unroll
      for (int i = 0; i < kBpl; ++i)
        if (j < 0 && mm[i]) { j = __ffs(mm[i]) - 1 + 32 * i; mm[i] &= mm[i] - 1; }
      return j;
    };
    int st_rounds = 0;
    if (kStats) {
#pragma unroll
      for (int i = 0; i < kBpl; ++i) st_active += (unsigned)__popc(mm[i]);
    }
    for (int j0 = pop(); j0 >= 0; j0 = pop()) {
      if (kStats) ++st_rounds;
      int j1 = pop();
      const bool two = j1 >= 0;
      if (!two) j1 = j0;
      const int q0 = bucket_of(j0), q1 = bucket_of(j1);
      const int r0 = q0 * 32 + lane, r1 = q1 * 32 + lane;
      const float4 p0 = pts[r0];
      const float4 p1 = pts[r1];                      // the same line again when there is no second bucket
      {
        const float d = sqdist_ref(p0.x - cx, p0.y - cy, p0.z - cz);
        const float old = s_t[r0];
        const float nt = fminf(d, old);
        const bool ch = nt != old;
        if (ch) s_t[r0] = nt;
        if (__any_sync(0xFFFFFFFFu, ch)) { bucket_max(j0, q0, nt, p0); warp_stale = true; }
      }
      if (two) {
        const float d = sqdist_ref(p1.x - cx, p1.y - cy, p1.z - cz);
        const float old = s_t[r1];
        const float nt = fminf(d, old);
        const bool ch = nt != old;
        if (ch) s_t[r1] = nt;
        if (__any_sync(0xFFFFFFFFu, ch)) { bucket_max(j1, q1, nt, p1); warp_stale = true; }
      }
    }
    // ---- 3. warp candidate (only if one of its buckets changed), one barrier, sub-block winner
    WarpRec *rec = s_rec + (it & 1) * kWarps;
    if (warp_stale) {
      int lv = bmax[0], li = 0;                       // this lane's best bucket
      uint32_t lk = bkey[0];
#pragma unroll
      for (int i = 1; i < kBpl; ++i)
        if (bmax[i] > lv || (bmax[i] == lv && bkey[i] < lk)) { lv = bmax[i]; lk = bkey[i]; li = i; }
      const int wv = __reduce_max_sync(0xFFFFFFFFu, lv);
      const uint32_t wk = __reduce_min_sync(0xFFFFFFFFu, lv == wv ? lk : kNoKey);
      __syncwarp();                                   // s_m4 of this warp's buckets was written by other lanes
      if (lv == wv && lk == wk) {                     // one lane when the key is real; any of them otherwise (all padding: never wins)
        const float4 c = s_m4[bucket_of(lane + 32 * li)];
        WarpRec w;
        w.v = wv; w.key = wk; w.x = c.x; w.y = c.y; w.z = c.z;
        w.pad[0] = w.pad[1] = w.pad[2] = 0;
        s_mine[warp] = w;
        rec[warp] = w;
      }
      warp_stale = false;
    } else if (lane == 0) {
      rec[warp] = s_mine[warp];
    }
    if (kStats) {
      __syncwarp();
      if (lane == 0) rec[warp].pad[0] = st_rounds;
    }
    sub_barrier(1 + sub, kSub);
    if (kStats) {   // rounds of the slowest warp = what this iteration's chain paid for the updates
      const int crit = __reduce_max_sync(0xFFFFFFFFu, lane < kWarps ? rec[lane].pad[0] : 0);
      st_crit += (unsigned)crit;
#pragma unroll
      for (int h = 0; h < 6; ++h) st_hist[h] += (min(crit, 5) == h);
    }
    {
      int v = INT_MIN;
      uint32_t kk = kNoKey;
      if (lane < kWarps) { v = rec[lane].v; kk = rec[lane].key; }
      const int bv = __reduce_max_sync(0xFFFFFFFFu, v);
      const uint32_t win_key = __reduce_min_sync(0xFFFFFFFFu, v == bv ? kk : kNoKey);
      const int src = __ffs(__ballot_sync(0xFFFFFFFFu, v == bv && kk == win_key)) - 1;
      cx = rec[src].x; cy = rec[src].y; cz = rec[src].z;          // broadcast reads
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

  if (kStats) {
    if (lane == 0) atomicAdd(&g_smem_stats[0], st_active);
    if (tid == 0) {
      atomicAdd(&g_smem_stats[1], st_crit);
      atomicAdd(&g_smem_stats[2], (unsigned long long)(m > 0 ? m - 1 : 0));
      for (int h = 0; h < 6; ++h) atomicAdd(&g_smem_stats[4 + h], st_hist[h]);
    }
  }
  if (prm.temp) {
    float *temp = prm.temp + cloud * (size_t)n;
    sub_barrier(1 + sub, kSub);
    for (int r = tid; r < nb * 32; r += kSub) {
      const int orig = __float_as_int(pts[r].w);
      if (orig >= 0) temp[orig] = s_t[r];
    }
  }
}

// ---- launch 2, default form: TWO samples per traversal of the latency chain whenever the second one is provably the next.
// After the barrier every warp knows the best bucket maximum A (the new sample) and the best maximum B among all OTHER buckets.
// B is exactly the sample that would follow A -- before A's update has been applied -- if
//   (1) A's update leaves B alone:            !(d(A, B) < t_B)                        (same arithmetic as the update itself),
//   (2) nothing in A's own bucket can stay above B:   min(t2_A, diag2_A) < t_B,  t2_A = the second largest running distance of
//       A's bucket (kept with the bucket maximum), diag2_A = the squared diagonal of its box (+inf if the bucket holds a
//       non-finite point): every other point C of that bucket ends at min(t_C, d(A, C)) <= both,
//   (3) A and B are finite (a non-finite sample changes nothing and would be selected again) and B is not the last sample.
// Every other point ends at or below its bucket's old maximum <= t_B, and ties at t_B already lost to B's tie key, so the
// (distance, key) arg-max after A's update is B.  Both updates are then applied in ONE traversal (t <- min(d_B, min(d_A, t)),
// the reference's order) and the pair costs one barrier.  On KITTI-like scenes the conditions hold for 98 % of the samples
// (16384 -> 4096): 2074 traversals instead of 4095.  Shared memory per cloud grows by the second-largest value and the box
// diagonal of every bucket (8 B per bucket) and by the second candidate of every warp.
template <int kWarps>
__global__ void __launch_bounds__(1024, 1) fps_pair_kernel(SmemFpsParams prm) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  constexpr int kSub = kWarps * 32, kRecs = 2 * kWarps;
  const int sub = threadIdx.x / kSub, tid = threadIdx.x % kSub, lane = tid & 31, warp = tid >> 5;
  const int cloud_i = blockIdx.x * prm.clouds_per_cta + sub;
  if (sub >= prm.clouds_per_cta || cloud_i >= prm.b) return;        // whole sub-blocks leave: their barrier is their own
  const int n = prm.n, m = prm.m, L = prm.L;
  const int nb = (n + 31) >> 5;                       // buckets of 32 consecutive sorted slots
  unsigned char *base = s_dyn + (size_t)sub * prm.cloud_smem;
  float4 *s_m4 = reinterpret_cast<float4 *>(base);                              // [nb] coordinates of the bucket maximum
  float *s_t = reinterpret_cast<float *>(base + (size_t)nb * 16);               // [nb * 32], -1 for padding
  int *s_b2 = reinterpret_cast<int *>(base + (size_t)nb * 144);                 // [nb] bits of the second largest value
  float *s_diag = reinterpret_cast<float *>(base + (size_t)nb * 148);           // [nb] upper bound of d(p, q) inside the bucket
  WarpRec *s_mine = reinterpret_cast<WarpRec *>(base + (((size_t)nb * 152 + 15) & ~(size_t)15));   // [2 kWarps] (private to each warp)
  WarpRec *s_rec = s_mine + kRecs;                                              // [2][2 kWarps]
  const size_t cloud = cloud_i;
  const float4 *pts = prm.sorted + cloud * (size_t)prm.cap;
  int *idx = prm.idx + cloud * (size_t)m;
  float *new_xyz = prm.new_xyz ? prm.new_xyz + cloud * (size_t)m * 3 : nullptr;
  const int my_bucket = lane * kWarps + warp;         // lane j of this warp owns bucket j * kWarps + warp

  float bx0 = __int_as_float(0x7f800000), by0 = bx0, bz0 = bx0, bx1 = -bx0, by1 = -bx0, bz1 = -bx0;   // empty box: never active
  int bmax = INT_MIN;          // bits of the bucket's largest running distance (>= 0: int order == float order)
  uint32_t bkey = kNoKey;      // tie key of that point

  // (Re)compute maximum, tie key and second largest value of bucket b (owned by lane j) from the warp's 32 values.
  auto bucket_max = [&](int j, int b, float t, const float4 &p) {
    const int orig = __float_as_int(p.w);
    const int vb = __float_as_int(t);
    const int wv = __reduce_max_sync(0xFFFFFFFFu, vb);
    const uint32_t kk = (vb == wv && orig >= 0) ? fps_key((uint32_t)orig, L) : kNoKey;
    const uint32_t wk = __reduce_min_sync(0xFFFFFFFFu, kk);
    const bool top = kk == wk && kk != kNoKey;          // keys are unique: exactly one lane (none if the bucket is all padding)
    const int v2 = __reduce_max_sync(0xFFFFFFFFu, top ? INT_MIN : vb);
    if (top) { s_m4[b] = p; s_b2[b] = v2; }
    if (lane == j) { bmax = wv; bkey = wk; }
  };

  // ---- setup: running distances, boxes, bucket maxima
  {
    const float *temp = prm.temp ? prm.temp + cloud * (size_t)n : nullptr;
    for (int j = 0; j < 32; ++j) {
      const int b = j * kWarps + warp;
      if (b >= nb) break;                               // warp-uniform
      const float4 p = pts[b * 32 + lane];
      const int orig = __float_as_int(p.w);
      const float t = orig >= 0 ? (temp ? temp[orig] : 1e10f) : -1.f;
      s_t[b * 32 + lane] = t;
      const bool fin = orig >= 0 && isfinite(p.x) && isfinite(p.y) && isfinite(p.z);  // others can never change: not in the box
      const bool odd = __any_sync(0xFFFFFFFFu, orig >= 0 && !fin);                     // a real point outside the box
      const int l0 = __reduce_min_sync(0xFFFFFFFFu, fin ? f2ord(p.x) : INT_MAX), h0 = __reduce_max_sync(0xFFFFFFFFu, fin ? f2ord(p.x) : INT_MIN);
      const int l1 = __reduce_min_sync(0xFFFFFFFFu, fin ? f2ord(p.y) : INT_MAX), h1 = __reduce_max_sync(0xFFFFFFFFu, fin ? f2ord(p.y) : INT_MIN);
      const int l2 = __reduce_min_sync(0xFFFFFFFFu, fin ? f2ord(p.z) : INT_MAX), h2 = __reduce_max_sync(0xFFFFFFFFu, fin ? f2ord(p.z) : INT_MIN);
      if (lane == j) {
        float diag = __int_as_float(0x7f800000);
        if (l0 <= h0) {
          bx0 = ord2f(l0); bx1 = ord2f(h0);
          by0 = ord2f(l1); by1 = ord2f(h1);
          bz0 = ord2f(l2); bz1 = ord2f(h2);
          // the update computes d(A, C) = sqdist_ref of coordinate differences that are bounded by the box extents; rounding is
          // monotone, the factor covers the rest
          if (!odd) diag = sqdist_ref(bx1 - bx0, by1 - by0, bz1 - bz0) * 1.0001f;
        }
        s_diag[b] = diag;
        s_m4[b] = make_float4(0.f, 0.f, 0.f, 0.f);
        s_b2[b] = INT_MIN;
      }
      __syncwarp();
      bucket_max(j, b, t, p);
    }
    if (lane < 2) {
      WarpRec w;
      w.v = INT_MIN; w.key = kNoKey; w.x = w.y = w.z = 0.f; w.pad[0] = INT_MIN; w.pad[1] = 0x7f800000; w.pad[2] = 0;
      s_mine[2 * warp + lane] = w;
      s_rec[2 * warp + lane] = w;
      s_rec[kRecs + 2 * warp + lane] = w;
    }
    __syncwarp();
  }
  float c1x = __ldg(prm.xyz + cloud * (size_t)n * 3), c1y = __ldg(prm.xyz + cloud * (size_t)n * 3 + 1),
        c1z = __ldg(prm.xyz + cloud * (size_t)n * 3 + 2);       // idx[0] = 0
  float c2x = 0.f, c2y = 0.f, c2z = 0.f;
  bool two_c = false;                                           // a second sample is pending with the first
  if (tid == 0 && m > 0) {
    idx[0] = 0;
    if (new_xyz) { new_xyz[0] = c1x; new_xyz[1] = c1y; new_xyz[2] = c1z; }
  }
  bool warp_stale = true;
  unsigned round = 0;

  for (int k = 1; k < m; ++round) {     // k = samples selected so far; the pending one or two have not been applied yet
    // ---- 1. which of this warp's buckets can still be lowered?  A point p changes only if d(sample, p) < t_p <= the bucket's own
    //         largest running distance.  (NaN sample -> comparison false -> active)
    const float bound = __int_as_float(bmax);
    bool act;
    {
      const float ax = fmaxf(fmaxf(bx0 - c1x, c1x - bx1), 0.f);
      const float ay = fmaxf(fmaxf(by0 - c1y, c1y - by1), 0.f);
      const float az = fmaxf(fmaxf(bz0 - c1z, c1z - bz1), 0.f);
      act = !((ax * ax + ay * ay + az * az) * kCullShrink >= bound);
    }
    if (two_c) {
      const float ax = fmaxf(fmaxf(bx0 - c2x, c2x - bx1), 0.f);
      const float ay = fmaxf(fmaxf(by0 - c2y, c2y - by1), 0.f);
      const float az = fmaxf(fmaxf(bz0 - c2z, c2z - bz1), 0.f);
      act = act || !((ax * ax + ay * ay + az * az) * kCullShrink >= bound);
    }
    uint32_t mm = __ballot_sync(0xFFFFFFFFu, act && my_bucket < nb);
    // ---- 2. exact update of the active buckets: one point per lane, coordinates from the L2-resident sorted copy, two buckets
    //         per round so that their L2 latencies overlap.  An update that culling would have skipped is a no-op anyway.
    while (mm) {
      const int j0 = __ffs(mm) - 1;
      mm &= mm - 1;
      const bool two = mm != 0;
      const int j1 = two ? __ffs(mm) - 1 : j0;
      mm &= mm - 1;                                   // no-op when mm is already 0
      const int q0 = j0 * kWarps + warp, q1 = j1 * kWarps + warp;
      const int r0 = q0 * 32 + lane, r1 = q1 * 32 + lane;
      const float4 p0 = pts[r0];
      const float4 p1 = pts[r1];                      // the same line again when there is no second bucket
      {
        const float old = s_t[r0];
        float nt = fminf(sqdist_ref(p0.x - c1x, p0.y - c1y, p0.z - c1z), old);
        if (two_c) nt = fminf(sqdist_ref(p0.x - c2x, p0.y - c2y, p0.z - c2z), nt);
        const bool ch = nt != old;
        if (ch) s_t[r0] = nt;
        if (__any_sync(0xFFFFFFFFu, ch)) { bucket_max(j0, q0, nt, p0); warp_stale = true; }
      }
      if (two) {
        const float old = s_t[r1];
        float nt = fminf(sqdist_ref(p1.x - c1x, p1.y - c1y, p1.z - c1z), old);
        if (two_c) nt = fminf(sqdist_ref(p1.x - c2x, p1.y - c2y, p1.z - c2z), nt);
        const bool ch = nt != old;
        if (ch) s_t[r1] = nt;
        if (__any_sync(0xFFFFFFFFu, ch)) { bucket_max(j1, q1, nt, p1); warp_stale = true; }
      }
    }
    // ---- 3. the warp's best and second best bucket (only if one of its buckets changed), one barrier
    WarpRec *rec = s_rec + (round & 1u) * kRecs;
    if (warp_stale) {
      const int wv = __reduce_max_sync(0xFFFFFFFFu, bmax);
      const uint32_t wk = __reduce_min_sync(0xFFFFFFFFu, bmax == wv ? bkey : kNoKey);
      const bool top = bmax == wv && bkey == wk;      // one lane when the key is real; every lane when the warp has no real bucket
      const int lv2 = top ? INT_MIN : bmax;
      const int wv2 = __reduce_max_sync(0xFFFFFFFFu, lv2);
      const uint32_t wk2 = __reduce_min_sync(0xFFFFFFFFu, (!top && lv2 == wv2) ? bkey : kNoKey);
      const bool second = !top && bmax == wv2 && bkey == wk2;
      __syncwarp();                                   // s_m4 / s_b2 of this warp's buckets were written by other lanes
      if (top || second) {
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        int b2 = INT_MIN;
        float dg = __int_as_float(0x7f800000);
        if (my_bucket < nb) { c = s_m4[my_bucket]; b2 = s_b2[my_bucket]; dg = s_diag[my_bucket]; }
        WarpRec w;
        w.v = top ? wv : wv2; w.key = top ? wk : wk2; w.x = c.x; w.y = c.y; w.z = c.z;
        w.pad[0] = b2; w.pad[1] = __float_as_int(dg); w.pad[2] = 0;
        const int e = 2 * warp + (top ? 0 : 1);
        s_mine[e] = w;
        rec[e] = w;
      }
      warp_stale = false;
    } else if (lane < 2) {
      rec[2 * warp + lane] = s_mine[2 * warp + lane];
    }
    sub_barrier(1 + sub, kSub);
    // ---- 4. the two best buckets of the cloud: A = the new sample, B = the candidate for the one after it
    {
      int v = INT_MIN;
      uint32_t kk = kNoKey;
      if (lane < kRecs) { v = rec[lane].v; kk = rec[lane].key; }
      const int bv = __reduce_max_sync(0xFFFFFFFFu, v);
      const uint32_t key_a = __reduce_min_sync(0xFFFFFFFFu, v == bv ? kk : kNoKey);
      const int src_a = __ffs(__ballot_sync(0xFFFFFFFFu, v == bv && kk == key_a)) - 1;
      const int vv = lane == src_a ? INT_MIN : v;
      const int bv2 = __reduce_max_sync(0xFFFFFFFFu, vv);
      const uint32_t key_b = __reduce_min_sync(0xFFFFFFFFu, (lane != src_a && vv == bv2) ? kk : kNoKey);
      const int src_b = __ffs(__ballot_sync(0xFFFFFFFFu, lane != src_a && vv == bv2 && kk == key_b)) - 1;
      const WarpRec &ra = rec[src_a];
      const float ax = ra.x, ay = ra.y, az = ra.z;
      bool pair = false;
      float bx = 0.f, by = 0.f, bz = 0.f;
      if (k + 2 < m && bv2 > 0 && src_b >= 0 && key_b != kNoKey) {
        const WarpRec &rb = rec[src_b];
        bx = rb.x; by = rb.y; bz = rb.z;
        const float t_b = __int_as_float(bv2);
        const float d_ab = sqdist_ref(bx - ax, by - ay, bz - az);        // what B's lane would compute in A's update
        const float reach = fminf(__int_as_float(ra.pad[0]), __int_as_float(ra.pad[1]));
        pair = !(d_ab < t_b) && reach < t_b && isfinite(ax) && isfinite(ay) && isfinite(az) && isfinite(bx) && isfinite(by) &&
               isfinite(bz);
      }
      if (tid == 0) {
        idx[k] = (int)fps_unkey(key_a, L);
        if (new_xyz) { new_xyz[(size_t)k * 3 + 0] = ax; new_xyz[(size_t)k * 3 + 1] = ay; new_xyz[(size_t)k * 3 + 2] = az; }
        if (pair) {
          idx[k + 1] = (int)fps_unkey(key_b, L);
          if (new_xyz) { new_xyz[(size_t)k * 3 + 3] = bx; new_xyz[(size_t)k * 3 + 4] = by; new_xyz[(size_t)k * 3 + 5] = bz; }
        }
      }
      c1x = ax; c1y = ay; c1z = az;
      c2x = bx; c2y = by; c2z = bz;
      two_c = pair;
      k += pair ? 2 : 1;
    }
  }

  if (prm.temp) {
    float *temp = prm.temp + cloud * (size_t)n;
    sub_barrier(1 + sub, kSub);
    for (int r = tid; r < nb * 32; r += kSub) {
      const int orig = __float_as_int(pts[r].w);
      if (orig >= 0) temp[orig] = s_t[r];
    }
  }
}

int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

template <int T, int P>
int launch_sort(int b, int n, const float *xyz, float4 *sorted, cudaStream_t stream) {
  using Sort = cub::BlockRadixSort<uint32_t, T, P>;
  auto kern = fps_sort_kernel<T, P>;
  const size_t smem = sizeof(typename Sort::TempStorage);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("fps (sort): smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  kern<<<b, T, smem, stream>>>(n, xyz, sorted);
  return check_launch("furthest_point_sampling (sort)");
}


// Sub-block shape by bucket count; clouds per CTA by what 1024 threads and the shared memory of one SM hold.  In throughput
// mode (ws3d_set_fps_mode(1)) CTAs are packed full: the sampler is meant to leave the SMs to other work.  Otherwise (a batch
// too large for the cluster kernels) clouds are packed only as far as the batch exceeds the SM count.
// WS3D_FPS_SMEM_CLOUDS caps it (1 = one cloud per CTA, which the hardware spreads over as many SMs).
struct SmemShape { int warps, bpl, cloud_smem, clouds_per_cta; };
SmemShape smem_shape(int b, int n) {
  static const int max_clouds = std::max(1, env_int("WS3D_FPS_SMEM_CLOUDS", 8));
  // 16384 points (512 buckets): 0 = 16 warps x 1 bucket per lane, two clouds per CTA (1024 threads); 1 = 8 warps x 2, three
  // clouds (768 threads); 2 = 4 warps x 4, three clouds (384 threads).  Shared memory admits three clouds per SM.
  static const int wide_shape = env_int("WS3D_FPS_SMEM_SHAPE", 0);
  const int nb = (n + 31) / 32;
  SmemShape sh;
  sh.bpl = 1;
  sh.warps = nb <= 128 ? 4 : nb <= 256 ? 8 : 16;
  if (sh.warps == 16 && wide_shape == 1) { sh.warps = 8; sh.bpl = 2; }
  if (sh.warps == 16 && wide_shape == 2) { sh.warps = 4; sh.bpl = 4; }
  sh.cloud_smem = nb * 144 + 3 * sh.warps * (int)sizeof(WarpRec);
  const int by_threads = sh.bpl == 1 ? 32 / sh.warps : 3;
  int cpc = std::min(std::min(by_threads, (226 * 1024) / sh.cloud_smem), max_clouds);
  if (fps_mode() != 1) cpc = std::min(cpc, (b + num_sms() - 1) / num_sms());
  sh.clouds_per_cta = std::max(1, std::min(cpc, b));
  return sh;
}

}  // namespace

int fps_smem_clouds_per_cta(int b, int n) { return smem_shape(b, n).clouds_per_cta; }

// Same shapes as the bucketed kernel: 2048 <= n <= 16384, m >= 64.
int fps_smem_launch(const FpsParams &prm, int b, cudaStream_t stream) {
  const int n = prm.n;
  const int cap = n <= 4096 ? 4096 : n <= 8192 ? 8192 : 16384;
  float4 *sorted = (float4 *)scratch((size_t)b * cap * sizeof(float4), 8);
  if (!sorted) return (int)cudaErrorMemoryAllocation;
  int rc = n <= 4096 ? launch_sort<256, 16>(b, n, prm.xyz, sorted, stream)
         : n <= 8192 ? launch_sort<512, 16>(b, n, prm.xyz, sorted, stream)
                     : launch_sort<512, 32>(b, n, prm.xyz, sorted, stream);
  if (rc) return rc;
  SmemFpsParams q;
  q.b = b; q.n = n; q.m = prm.m; q.L = prm.L; q.cap = cap;
  const int nb = (n + 31) / 32;
  q.sorted = sorted; q.temp = prm.temp; q.idx = prm.idx; q.new_xyz = prm.new_xyz; q.xyz = prm.xyz;
  const SmemShape sh = smem_shape(b, n);
  const int warps = sh.warps, cpc = sh.clouds_per_cta;
  q.cloud_smem = sh.cloud_smem;
  q.clouds_per_cta = cpc;
  const size_t smem = (size_t)cpc * q.cloud_smem;
  const int grid = (b + cpc - 1) / cpc;
  cudaError_t e = cudaSuccess;
#define WS3D_FPS_SMEM_LAUNCH(W, BPL, MAXT)                                                                           \
  do {                                                                                                                \
    auto kern = fps_smem_kernel<W, BPL, MAXT>;                                                                        \
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);                          \
    if (e == cudaSuccess) kern<<<grid, cpc * W * 32, smem, stream>>>(q);                                              \
  } while (0)
  static const int stats = env_int("WS3D_FPS_STATS", 0);
  static const int use_pair = env_int("WS3D_FPS_PAIR", 1);   // 0: one sample per traversal (fps_smem_kernel), for A/B runs
  const int pair_smem = ((nb * 152 + 15) & ~15) + 6 * warps * (int)sizeof(WarpRec);
  if (use_pair && !stats && sh.bpl == 1 && (size_t)cpc * pair_smem <= 226u * 1024u) {
    q.cloud_smem = pair_smem;
    const size_t smem_p = (size_t)cpc * pair_smem;
#define WS3D_FPS_PAIR_LAUNCH(W)                                                                                      \
  do {                                                                                                                \
    auto kern = fps_pair_kernel<W>;                                                                                   \
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);                          \
    if (e == cudaSuccess) kern<<<grid, cpc * W * 32, smem_p, stream>>>(q);                                            \
  } while (0)
    if (warps == 4) WS3D_FPS_PAIR_LAUNCH(4);
    else if (warps == 8) WS3D_FPS_PAIR_LAUNCH(8);
    else WS3D_FPS_PAIR_LAUNCH(16);
#undef WS3D_FPS_PAIR_LAUNCH
  } else if (stats && warps == 16 && sh.bpl == 1) {   // debug: active buckets and update rounds per iteration (synchronises)
    unsigned long long z[16] = {};
    cudaMemcpyToSymbol(g_smem_stats, z, sizeof(z));
    auto kern = fps_smem_kernel<16, 1, 1024, true>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e == cudaSuccess) kern<<<grid, cpc * 16 * 32, smem, stream>>>(q);
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(z, g_smem_stats, sizeof(z));
    const double it = (double)std::max(1ull, z[2]);
    fprintf(stderr, "[fps smem stats] b=%d n=%d m=%d clouds/CTA=%d: active buckets per iteration %.2f, update rounds of the slowest warp "
                    "per iteration %.3f, share of iterations with 0/1/2/3/4/5+ rounds: %.3f %.3f %.3f %.3f %.3f %.3f\n", b, n, prm.m, cpc,
            z[0] / it, z[1] / it, z[4] / it, z[5] / it, z[6] / it, z[7] / it, z[8] / it, z[9] / it);
  } else if (sh.bpl == 4) WS3D_FPS_SMEM_LAUNCH(4, 4, 384);
  else if (sh.bpl == 2) WS3D_FPS_SMEM_LAUNCH(8, 2, 768);
  else if (warps == 4) WS3D_FPS_SMEM_LAUNCH(4, 1, 1024);
  else if (warps == 8) WS3D_FPS_SMEM_LAUNCH(8, 1, 1024);
  else WS3D_FPS_SMEM_LAUNCH(16, 1, 1024);
#undef WS3D_FPS_SMEM_LAUNCH
  if (e != cudaSuccess) { set_error("fps (shared-memory distances): smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  return check_launch("furthest_point_sampling (shared-memory distances)");
}

}  // namespace ws3d
