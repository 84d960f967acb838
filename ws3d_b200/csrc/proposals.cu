// Proposal clustering between the two WS3D stages.
//
// The live WS3D inference script replaces box NMS / roipool3d by two BEV operations written as dense PyTorch
// (tools/eval_auto.py): a greedy "radius NMS" over the predicted centres (:263-279, a Python loop over a P x P
// distance matrix; served here by ws3d_radius_nms in iou3d.cu, which shares the NMS mask + greedy-scan kernels)
// and a 4 m cylinder crop of the scene round every surviving centre (:289-291, :327-343, a P x N distance matrix
// and one boolean-mask gather per centre).  This file is the crop: one warp per centre scans the points 32 at a
// time and compacts the members in index order with ballot / popc, exactly like roipool3d's box crop but with the
// cylinder predicate and without a sample cap on the count.
//
// Exactness: the membership test is distance_2(centre, point) < radius with distance_2 as lib/utils/distance.py:3
// evaluates it in float32 torch kernels -- subtract, square, two-term sum, sqrt, one IEEE rounding each.
#include "common.cuh"

namespace ws3d {
namespace {

__device__ __forceinline__ float bev_dist(float ax, float az, float bx, float bz) {
  const float dx = __fsub_rn(ax, bx), dz = __fsub_rn(az, bz);
  return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dz, dz)));
}

constexpr int kWarps = 4;        // centres per CTA
constexpr int kChunk = 4096;     // points staged in shared memory at a time (x, z: 32 KB)

// The CTA's warps share every chunk of points (one coalesced pass over pts per 4 centres instead of one per centre,
// and the scan itself reads shared memory); members are appended in index order because chunks and the 32-point
// steps inside them are visited in order and ballot / popc orders the lanes.
__global__ void __launch_bounds__(kWarps * 32) cylinder_query_kernel(int n, int m, int cap, float radius,
                                                                     const float *__restrict__ pts,
                                                                     const float *__restrict__ centers,
                                                                     int *__restrict__ idx, int *__restrict__ cnt,
                                                                     unsigned char *__restrict__ any) {
  __shared__ float2 s_xz[kChunk];
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * kWarps + (threadIdx.x >> 5);
  const bool live = c < m;
  const float cx = live ? __ldg(centers + 2 * (size_t)c) : 0.f, cz = live ? __ldg(centers + 2 * (size_t)c + 1) : 0.f;
  int *row = idx + (size_t)(live ? c : 0) * cap;
  int found = 0;
  for (int base = 0; base < n; base += kChunk) {
    const int len = min(kChunk, n - base);
    __syncthreads();
    for (int k = threadIdx.x; k < len; k += kWarps * 32)
      s_xz[k] = make_float2(__ldg(pts + 3 * (size_t)(base + k)), __ldg(pts + 3 * (size_t)(base + k) + 2));
    __syncthreads();
    if (!live) continue;
#pragma unroll 4
    for (int k0 = 0; k0 < len; k0 += 32) {
      const int k = k0 + lane;
      bool hit = false;
      if (k < len) {
        const float2 p = s_xz[k];
        hit = bev_dist(cx, cz, p.x, p.y) < radius;
      }
      const unsigned ballot = __ballot_sync(0xFFFFFFFFu, hit);
      if (hit) {
        const int slot = found + __popc(ballot & ((1u << lane) - 1u));
        if (slot < cap) row[slot] = base + k;
        if (any) any[base + k] = 1;   // benign race: every writer stores 1
      }
      found += __popc(ballot);
    }
  }
  if (live && lane == 0) cnt[c] = found;
}

}  // namespace
}  // namespace ws3d

using namespace ws3d;

WS3D_API int ws3d_cylinder_query(int n, int m, int cap, float radius, const float *pts, const float *centers, int *idx,
                                 int *cnt, unsigned char *any, ws3d_stream_t stream) {
  const char *what = "cylinder_query";
  if (n < 0 || m < 0 || cap < 0) return fail_arg(what);
  if (m == 0) return 0;
  if (!centers || !cnt || (cap > 0 && !idx) || (n > 0 && !pts)) return fail_arg(what);
  cylinder_query_kernel<<<(unsigned)ceil_div(m, kWarps), kWarps * 32, 0, to_stream(stream)>>>(n, m, cap, radius, pts, centers,
                                                                                            idx, cnt, any);
  return check_launch(what);
}
