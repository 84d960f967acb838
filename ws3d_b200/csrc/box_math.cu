// Loss-side box math and the loader's point subsampling.
//
//  * boxes3d_to_corners3d_torch (lib/utils/kitti_utils.py:104-131): ~15 eager torch kernels per call (fills, eight
//    concatenations, cos / sin, a batched 3x3 matmul, broadcast add, permute); called three times per Stage-2 training
//    step for the corner loss (lib/net/train_functions.py:266-269).  One thread per box here.
//  * the corner distance of that loss (:270-271): min(|P - G|, |P - G_flipped|) per corner, forward and the gradient
//    with respect to the predicted box (the ground truth carries none), fused with the three corner computations.
//  * the 16384-point subsampling of the data loader (lib/datasets/kitti_rcnn_dataset.py:424-452): depth split,
//    random choice, shuffle -- as one launch driven by the permutations numpy's RNG would draw, so that a loader that
//    keeps the raw cloud on the GPU reproduces the reference's sample bit for bit.
#include <math.h>

#include "common.cuh"

namespace ws3d {
namespace {

constexpr float kPi = 3.14159265358979323846f;   // (float)np.pi: `ry + np.pi` on a float32 tensor adds the rounded constant

struct Corners { float x[8], y[8], z[8]; };

// kitti_utils.py:104-131.  The batched matmul R (3x3) x corners (3x8) accumulates k = 0, 1, 2 in order with FMAs from
// a zero accumulator; the middle term multiplies an exact 0 (or 1), so x' = fma(sin, zc, rn(cos * xc)),
// y' = yc, z' = fma(cos, zc, rn(-sin * xc)), then one rounded add of the centre.
__device__ __forceinline__ void box_corners(const float *__restrict__ b, bool flip, Corners &c) {
  const float cx = b[0], cy = b[1], cz = b[2], h = b[3], w = b[4], l = b[5];
  const float ry = flip ? __fadd_rn(b[6], kPi) : b[6];
  const float cosa = cosf(ry), sina = sinf(ry);
  const float hl = l / 2.f, hw = w / 2.f;
  const float xs[8] = {hl, hl, -hl, -hl, hl, hl, -hl, -hl};
  const float zs[8] = {hw, -hw, -hw, hw, hw, -hw, -hw, hw};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float yc = k < 4 ? 0.f : -h;
    c.x[k] = __fadd_rn(__fmaf_rn(sina, zs[k], __fmul_rn(cosa, xs[k])), cx);
    c.y[k] = __fadd_rn(yc, cy);
    c.z[k] = __fadd_rn(__fmaf_rn(cosa, zs[k], __fmul_rn(-sina, xs[k])), cz);
  }
}

__global__ void __launch_bounds__(128) corners3d_kernel(int n, const float *__restrict__ boxes, int flip, float *__restrict__ out) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= n) return;
  float b[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) b[k] = __ldg(boxes + (size_t)i * 7 + k);
  Corners c;
  box_corners(b, flip != 0, c);
  float *o = out + (size_t)i * 24;
#pragma unroll
  for (int k = 0; k < 8; ++k) { o[3 * k] = c.x[k]; o[3 * k + 1] = c.y[k]; o[3 * k + 2] = c.z[k]; }
}

// torch.norm(d, dim=-1) of a 3-vector: sqrt(sum of squares), each step rounded
__device__ __forceinline__ float norm3(float dx, float dy, float dz) {
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}

// dist (n, 8) = min(|P - G|, |P - Gf|); with grad_dist: grad_pred (n, 7) = d(sum_k grad_dist[k] * dist[k]) / d(pred box).
// Subgradients as torch's autograd takes them: norm at 0 -> 0; a tie of the min -> half to either branch.
template <bool kGrad>
__global__ void __launch_bounds__(128) corner_distance_kernel(int n, const float *__restrict__ pred, const float *__restrict__ gt,
                                                               const float *__restrict__ grad_dist, float *__restrict__ dist,
                                                               float *__restrict__ grad_pred) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= n) return;
  float p[7], g[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) { p[k] = __ldg(pred + (size_t)i * 7 + k); g[k] = __ldg(gt + (size_t)i * 7 + k); }
  Corners P, G, Gf;
  box_corners(p, false, P);
  box_corners(g, false, G);
  box_corners(g, true, Gf);
  const float cosa = cosf(p[6]), sina = sinf(p[6]);
  float gx = 0.f, gy = 0.f, gz = 0.f, gh = 0.f, gw = 0.f, gl = 0.f, gr = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float ax = __fsub_rn(P.x[k], G.x[k]), ay = __fsub_rn(P.y[k], G.y[k]), az = __fsub_rn(P.z[k], G.z[k]);
    const float bx = __fsub_rn(P.x[k], Gf.x[k]), by = __fsub_rn(P.y[k], Gf.y[k]), bz = __fsub_rn(P.z[k], Gf.z[k]);
    const float d1 = norm3(ax, ay, az), d2 = norm3(bx, by, bz);
    if (!kGrad) {
      dist[(size_t)i * 8 + k] = fminf(d1, d2);   // NaN-free inputs: fminf == torch.min
    } else {
      const float go = __ldg(grad_dist + (size_t)i * 8 + k);
      const float w1 = d1 < d2 ? 1.f : (d1 == d2 ? 0.5f : 0.f), w2 = 1.f - w1;
      const float s1 = d1 > 0.f ? go * w1 / d1 : 0.f, s2 = d2 > 0.f ? go * w2 / d2 : 0.f;
      const float dxk = s1 * ax + s2 * bx, dyk = s1 * ay + s2 * by, dzk = s1 * az + s2 * bz;   // dL / dP_k
      const float sx = (k & 2) ? -0.5f : 0.5f;                         // xc = sx * l
      const float sz = (k == 0 || k == 3 || k == 4 || k == 7) ? 0.5f : -0.5f;   // zc = sz * w
      const float xc = sx * p[5], zc = sz * p[4];
      gx += dxk; gy += dyk; gz += dzk;
      if (k >= 4) gh -= dyk;
      gl += sx * (cosa * dxk - sina * dzk);
      gw += sz * (sina * dxk + cosa * dzk);
      gr += dxk * (-sina * xc + cosa * zc) + dzk * (-cosa * xc - sina * zc);
    }
  }
  if (kGrad) {
    float *o = grad_pred + (size_t)i * 7;
    o[0] = gx; o[1] = gy; o[2] = gz; o[3] = gh; o[4] = gw; o[5] = gl; o[6] = gr;
  }
}

// ---- subsample_points -----------------------------------------------------------------------------
// One CTA.  n > npoints: `choice` = the perm_near-selected (npoints - n_far) of the near points (depth < near_depth, in
// index order, as np.where lists them) followed by every far point; n <= npoints: choice[j] = perm[j] mod n (the tiled
// arange of kitti_rcnn_dataset.py:436-440).  Then out[i] = pts[choice[order[i]]] (np.random.shuffle as a permutation).
constexpr int kSubThreads = 1024;

__global__ void __launch_bounds__(kSubThreads) subsample_kernel(int n, int c, int npoints, int n_near, float near_depth,
                                                                float sub_last, const float *__restrict__ pts,
                                                                const float *__restrict__ depth, const int *__restrict__ perm,
                                                                const int *__restrict__ order, int *__restrict__ lists,
                                                                float *__restrict__ out, int *__restrict__ choice_out,
                                                                int *__restrict__ status) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int *near_idx = lists, *far_idx = lists + n;
  const bool down = n > npoints;
  if (down) {
    // stable compaction of the near / far indices (block scan, kSubThreads points per round)
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < n; base += kSubThreads) {
      const int k = base + tid;
      const bool near = k < n && (__ldg(depth + k) < near_depth);
      const unsigned ballot = __ballot_sync(0xFFFFFFFFu, near);
      if (lane == 0) s_warp[warp] = __popc(ballot);
      __syncthreads();
      int before = 0;
      for (int w = 0; w < warp; ++w) before += s_warp[w];
      const int nb = s_base;                       // near points before this round
      const int pos = nb + before + __popc(ballot & ((1u << lane) - 1u));
      if (k < n) {
        if (near) near_idx[pos] = k;
        else far_idx[k - pos] = k;                 // far points before k = k - (near points before k)
      }
      __syncthreads();
      if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < kSubThreads / 32; ++w) tot += s_warp[w];
        s_base = nb + tot;
      }
      __syncthreads();
    }
    if (tid == 0) *status = (s_base == n_near) ? 0 : 1;   // the host drew perm for a different near count
    __syncthreads();
    if (s_base != n_near) return;
  } else if (tid == 0) {
    *status = 0;
  }
  const int k_near = down ? npoints - (n - n_near) : 0;
  for (int i = tid; i < npoints; i += kSubThreads) {
    const int j = __ldg(order + i);
    int src;
    if (down) src = j < k_near ? near_idx[__ldg(perm + j)] : far_idx[j - k_near];
    else src = __ldg(perm + j) % n;
    if (choice_out) choice_out[i] = src;
    for (int q = 0; q < c; ++q) {
      float v = __ldg(pts + (size_t)src * c + q);
      if (q == c - 1 && c > 3) v = __fsub_rn(v, sub_last);
      out[(size_t)i * c + q] = v;
    }
  }
}

}  // namespace
}  // namespace ws3d

using namespace ws3d;

WS3D_API int ws3d_boxes3d_to_corners3d(int n, const float *boxes3d, int flip, float *corners, ws3d_stream_t stream) {
  if (n < 0) return fail_arg("boxes3d_to_corners3d");
  if (n == 0) return 0;
  if (!boxes3d || !corners) return fail_arg("boxes3d_to_corners3d (null pointer)");
  corners3d_kernel<<<ceil_div(n, 128), 128, 0, to_stream(stream)>>>(n, boxes3d, flip, corners);
  return check_launch("boxes3d_to_corners3d");
}

WS3D_API int ws3d_corner_distance(int n, const float *pred_boxes3d, const float *gt_boxes3d, float *dist, ws3d_stream_t stream) {
  if (n < 0) return fail_arg("corner_distance");
  if (n == 0) return 0;
  if (!pred_boxes3d || !gt_boxes3d || !dist) return fail_arg("corner_distance (null pointer)");
  corner_distance_kernel<false><<<ceil_div(n, 128), 128, 0, to_stream(stream)>>>(n, pred_boxes3d, gt_boxes3d, nullptr, dist, nullptr);
  return check_launch("corner_distance");
}

WS3D_API int ws3d_corner_distance_grad(int n, const float *pred_boxes3d, const float *gt_boxes3d, const float *grad_dist,
                                       float *grad_pred, ws3d_stream_t stream) {
  if (n < 0) return fail_arg("corner_distance_grad");
  if (n == 0) return 0;
  if (!pred_boxes3d || !gt_boxes3d || !grad_dist || !grad_pred) return fail_arg("corner_distance_grad (null pointer)");
  corner_distance_kernel<true><<<ceil_div(n, 128), 128, 0, to_stream(stream)>>>(n, pred_boxes3d, gt_boxes3d, grad_dist, nullptr, grad_pred);
  return check_launch("corner_distance_grad");
}

WS3D_API int ws3d_subsample_points(int n, int c, int npoints, int n_near, float near_depth, float sub_last, const float *pts,
                                   const float *depth, const int *perm, const int *order, float *out, int *choice,
                                   int *status, ws3d_stream_t stream) {
  const char *what = "subsample_points";
  if (n <= 0 || c < 3 || npoints <= 0) return fail_arg(what);
  if (!pts || !perm || !order || !out || !status) return fail_arg("subsample_points (null pointer)");
  if (n > npoints) {
    const int k_near = npoints - (n - n_near);
    if (!depth || n_near < 0 || n_near > n || k_near < 0 || k_near > n_near)
      return fail_arg("subsample_points (more far points than npoints, or fewer near points than needed: numpy's choice raises too)");
  }
  int *lists = (int *)scratch((size_t)2 * n * sizeof(int), 6);
  if (!lists) return (int)cudaErrorMemoryAllocation;
  subsample_kernel<<<1, kSubThreads, 0, to_stream(stream)>>>(n, c, npoints, n_near, near_depth, sub_last, pts, depth, perm, order, lists,
                                                             out, choice, status);
  return check_launch(what);
}
