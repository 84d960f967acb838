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

// ---------------------------------------------------------------------------------------------------------------
// Stage-1 training labels on the GPU.
// KittiRCNNDataset.generate_gaussian_training_labels (lib/datasets/kitti_rcnn_dataset.py:529-573) runs per sample in
// numpy inside the data loader (num_workers = 0): a (points x boxes) distance matrix, a running minimum, a scipy
// Gaussian and an argmin-driven regression target.  Here: one thread per point, the scene's boxes in shared memory.
// float32 arithmetic as numpy evaluates it (np.power(v, 2) is v * v; the Python-float constants are cast to float32
// first), the Gaussian in float64 like scipy.
namespace ws3d {
namespace {

constexpr int kLabelThreads = 256;
constexpr int kLabelMaxBoxes = 512;

__global__ void __launch_bounds__(kLabelThreads) gaussian_labels_kernel(int n, int g_max, const float *__restrict__ pts,
                                                                        const float *__restrict__ boxes,
                                                                        const int *__restrict__ num_gt, float gauss_height,
                                                                        float gauss_status, double gauss_cov, float fg_radius,
                                                                        float *__restrict__ cls_label,
                                                                        float *__restrict__ reg_label) {
  __shared__ float s_bx[kLabelMaxBoxes], s_bz[kLabelMaxBoxes];
  const size_t scene = blockIdx.y;
  const int g = num_gt ? min(max(__ldg(num_gt + scene), 0), g_max) : g_max;
  for (int k = threadIdx.x; k < g; k += kLabelThreads) {
    s_bx[k] = __ldg(boxes + (scene * g_max + k) * 7 + 0);
    s_bz[k] = __ldg(boxes + (scene * g_max + k) * 7 + 2);
  }
  __syncthreads();
  const int i = blockIdx.x * kLabelThreads + threadIdx.x;
  if (i >= n) return;
  const float *p = pts + (scene * n + i) * 3;
  const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
  const float hy = __fmul_rn(y, gauss_height);
  const float hy2 = __fmul_rn(hy, hy);
  float centre_dist = 100.f;                 // point_center_dist (:531)
  float best = 0.f;                          // min / argmin of dist_points2box (:561-562; first minimum wins)
  int target = -1;
  for (int k = 0; k < g; ++k) {
    const float dx = __fsub_rn(x, s_bx[k]), dz = __fsub_rn(z, s_bz[k]);
    const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), hy2), __fmul_rn(dz, dz)));   // :545-548
    const float c = fminf(fmaxf(__fsub_rn(d, gauss_status), 0.f), 100.f);                           // np.clip(d - status, 0, 100)
    centre_dist = (c != c || centre_dist != centre_dist) ? __fadd_rn(c, centre_dist) : fminf(centre_dist, c);   // np.minimum propagates NaN
    if (target < 0 || d < best) { best = d; target = k; }   // np.argmin: NaN handling is not reproduced (finite inputs)
  }
  float cls = 0.f, rx = 0.f, rz = 0.f;
  if (g > 0) {
    const double dd = (double)centre_dist;
    cls = (float)exp(-0.5 * dd * dd / gauss_cov);            // pdf(d; 0, cov) / (1 / sqrt(2 pi cov))   (:563-564)
    if (best < fg_radius) {                                  // foreground_big_mask (:568)
      rx = __fsub_rn(s_bx[target], x);
      rz = __fsub_rn(s_bz[target], z);
    }
  }
  cls_label[scene * n + i] = cls;
  float *r = reg_label + (scene * n + i) * 3;
  r[0] = rx; r[1] = 0.f; r[2] = rz;
}

}  // namespace
}  // namespace ws3d

WS3D_API int ws3d_gaussian_rpn_labels(int b, int n, int g_max, const float *pts, const float *gt_boxes3d, const int *num_gt,
                                      float gauss_height, float gauss_status, float gauss_cov, float fg_radius,
                                      float *cls_label, float *reg_label, ws3d_stream_t stream) {
  const char *what = "gaussian_rpn_labels";
  if (b < 0 || n < 0 || g_max < 0 || g_max > kLabelMaxBoxes || b > 65535) return fail_arg(what);
  if (b == 0 || n == 0) return 0;
  if (!pts || !cls_label || !reg_label || (g_max > 0 && !gt_boxes3d) || !(gauss_cov > 0.f)) return fail_arg(what);
  dim3 grid((unsigned)ceil_div(n, kLabelThreads), (unsigned)b);
  gaussian_labels_kernel<<<grid, kLabelThreads, 0, to_stream(stream)>>>(n, g_max, pts, gt_boxes3d, num_gt, gauss_height,
                                                                         gauss_status, (double)gauss_cov, fg_radius, cls_label,
                                                                         reg_label);
  return check_launch(what);
}
