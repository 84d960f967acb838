/*
 * ws3d_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A plain-C CPU restatement of the reference's GPU algorithms for the WS3D
 * set-abstraction / roipool3d / iou3d hot path.  It exists so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg can check and time
 * the path without a GPU.  The product (ws3d_b200/) must never import, link
 * or call anything in this directory.
 *
 * Parity status: the reference ships NO golden vectors or tests for this path
 * (SURVEY.md section 4).  This restatement is pinned instead by
 *   (1) fixtures under tests/golden/ that were produced by running the
 *       reference's own kernels (oracle/_ref, built by oracle/build_ref.sh
 *       from the unmodified sources) on a B200 -- see tests/golden/README.md;
 *   (2) the reference's own CPU entry points (roipool3d_cpu,
 *       pts_in_boxes3d_cpu) executed from oracle/_ref in the CPU test-suite.
 *
 * Floating point: the reference kernels are compiled by nvcc with FMA
 * contraction on.  Which product of "a*b +- c*d" is fused was read out of the
 * SASS of the reference objects (cuobjdump -sass oracle/_ref/obj/ objects, sm_100a,
 * nvcc 12.9 -O2) and is spelled here with explicit fmaf(); compile this file
 * with -ffp-contract=off so the host compiler adds none of its own.
 * Transcendentals (cosf/sinf/atan2f) come from the host libm here and from
 * libdevice on the GPU: they can differ in the last ulp, so for iou3d the
 * GPU oracle (oracle/_ref) arbitrates bit-exactness and this file is held to
 * 1e-5; all index-producing pointnet2 ops are bit-exact by construction.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

ORACLE_API void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* squared distance exactly as the three pointnet2 kernels compute it:
 * t = rn(dy*dy); t = fma(dx,dx,t); d = fma(dz,dz,t)    */
static inline float sqdist(float dx, float dy, float dz) {
  float t = dy * dy;
  t = fmaf(dx, dx, t);
  return fmaf(dz, dz, t);
}

/* ------------------------------------------------------------------------ */
/* pointnet2_lib/pointnet2/src/cuda_utils.h:10-14  (opt_n_threads)           */
ORACLE_API int oracle_opt_n_threads(int work_size) {
  const int pow_2 = (int)(log((double)work_size) / log(2.0));
  int t = 1 << pow_2;
  if (t > 1024) t = 1024;
  if (t < 1) t = 1;
  return t;
}

/* pointnet2_lib/pointnet2/src/sampling_gpu.cu:93-209
 * (furthest_point_sampling_kernel<block_size>), one "block" per cloud.
 * dataset (B,N,3), temp (B,N) running min-dist (caller fills 1e10), idxs (B,M).
 * The block structure is kept because it defines the tie-break: thread tid
 * owns k = tid, tid+BS, ... (strict > keeps its lowest k), and the shared-memory
 * tree (__update, :86-91) keeps the LOWER slot of each pair on ties.          */
ORACLE_API void oracle_furthest_point_sampling(int b, int n, int m, const float *dataset,
                                               float *temp, int *idxs) {
  if (m <= 0) return;
  const int bs = oracle_opt_n_threads(n);
#pragma omp parallel for schedule(dynamic, 1)
  for (int bi = 0; bi < b; ++bi) {
    const float *xyz = dataset + (size_t)bi * n * 3;
    float *tmp = temp + (size_t)bi * n;
    int *out = idxs + (size_t)bi * m;
    float *dists = (float *)malloc(sizeof(float) * bs);
    int *dists_i = (int *)malloc(sizeof(int) * bs);
    int old = 0;
    out[0] = old;
    for (int j = 1; j < m; ++j) {
      const float x1 = xyz[old * 3 + 0], y1 = xyz[old * 3 + 1], z1 = xyz[old * 3 + 2];
      for (int t = 0; t < bs; ++t) { dists[t] = -1.0f; dists_i[t] = 0; }
      for (int k = 0; k < n; ++k) {          /* ascending k == each thread's own order */
        const int t = k % bs;
        const float d = sqdist(xyz[k * 3 + 0] - x1, xyz[k * 3 + 1] - y1, xyz[k * 3 + 2] - z1);
        const float d2 = fminf(d, tmp[k]);
        tmp[k] = d2;
        if (d2 > dists[t]) { dists[t] = d2; dists_i[t] = k; }
      }
      for (int s = bs / 2; s >= 1; s >>= 1) {
        for (int t = 0; t < s; ++t) {
          const float v1 = dists[t], v2 = dists[t + s];
          const int i1 = dists_i[t], i2 = dists_i[t + s];
          dists[t] = fmaxf(v1, v2);
          dists_i[t] = v2 > v1 ? i2 : i1;
        }
      }
      old = dists_i[0];
      out[j] = old;
    }
    free(dists);
    free(dists_i);
  }
}

/* sampling_gpu.cu:8-24  out[b,c,j] = points[b,c,idx[b,j]] */
ORACLE_API void oracle_gather_points(int b, int c, int n, int m, const float *points,
                                     const int *idx, float *out) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci) {
      const float *src = points + ((size_t)bi * c + ci) * n;
      float *dst = out + ((size_t)bi * c + ci) * m;
      const int *id = idx + (size_t)bi * m;
      for (int j = 0; j < m; ++j) dst[j] = src[id[j]];
    }
}

/* sampling_gpu.cu:46-63  grad_points[b,c,idx[b,j]] += grad_out[b,c,j]
 * (atomicAdd on the GPU: summation order is unspecified there; index order here) */
ORACLE_API void oracle_gather_points_grad(int b, int c, int n, int m, const float *grad_out,
                                          const int *idx, float *grad_points) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci) {
      const float *g = grad_out + ((size_t)bi * c + ci) * m;
      float *dst = grad_points + ((size_t)bi * c + ci) * n;
      const int *id = idx + (size_t)bi * m;
      for (int j = 0; j < m; ++j) dst[id[j]] += g[j];
    }
}

/* ball_query_gpu.cu:9-45  first nsample hits in ascending point index,
 * first hit replicated into every slot, rows without a hit are left alone
 * (the caller zero-fills, pointnet2_utils.py:218).  radius2 = rn(r*r) in f32;
 * dx = new_x - x.                                                            */
ORACLE_API void oracle_ball_query(int b, int n, int m, float radius, int nsample,
                                  const float *new_xyz, const float *xyz, int *idx) {
  const float radius2 = radius * radius;
#pragma omp parallel for collapse(2) schedule(static)
  for (int bi = 0; bi < b; ++bi)
    for (int j = 0; j < m; ++j) {
      const float *q = new_xyz + ((size_t)bi * m + j) * 3;
      const float *p = xyz + (size_t)bi * n * 3;
      int *o = idx + ((size_t)bi * m + j) * nsample;
      const float qx = q[0], qy = q[1], qz = q[2];
      int cnt = 0;
      for (int k = 0; k < n; ++k) {
        const float d2 = sqdist(qx - p[k * 3 + 0], qy - p[k * 3 + 1], qz - p[k * 3 + 2]);
        if (d2 < radius2) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) o[l] = k;
          o[cnt] = k;
          ++cnt;
          if (cnt >= nsample) break;
        }
      }
    }
}

/* group_points_gpu.cu:47-66  out[b,c,j,s] = points[b,c,idx[b,j,s]] */
ORACLE_API void oracle_group_points(int b, int c, int n, int npoints, int nsample,
                                    const float *points, const int *idx, float *out) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci) {
      const float *src = points + ((size_t)bi * c + ci) * n;
      float *dst = out + ((size_t)bi * c + ci) * npoints * nsample;
      const int *id = idx + (size_t)bi * npoints * nsample;
      for (int e = 0; e < npoints * nsample; ++e) dst[e] = src[id[e]];
    }
}

/* group_points_gpu.cu:8-25 */
ORACLE_API void oracle_group_points_grad(int b, int c, int n, int npoints, int nsample,
                                         const float *grad_out, const int *idx,
                                         float *grad_points) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci) {
      const float *g = grad_out + ((size_t)bi * c + ci) * npoints * nsample;
      float *dst = grad_points + ((size_t)bi * c + ci) * n;
      const int *id = idx + (size_t)bi * npoints * nsample;
      for (int e = 0; e < npoints * nsample; ++e) dst[id[e]] += g[e];
    }
}

/* interpolate_gpu.cu:9-52  three nearest known points, strict '<' cascade
 * (ascending distance, lowest index first on ties); outputs SQUARED distances.
 * best* are doubles initialised to 1e40 in the reference (:30): a slot that is
 * never filled is stored as (float)1e40 = +inf with index 0.                  */
ORACLE_API void oracle_three_nn(int b, int n, int m, const float *unknown, const float *known,
                                float *dist2, int *idx) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int bi = 0; bi < b; ++bi)
    for (int i = 0; i < n; ++i) {
      const float *u = unknown + ((size_t)bi * n + i) * 3;
      const float *kn = known + (size_t)bi * m * 3;
      const float ux = u[0], uy = u[1], uz = u[2];
      double best1 = 1e40, best2 = 1e40, best3 = 1e40;
      int besti1 = 0, besti2 = 0, besti3 = 0;
      for (int k = 0; k < m; ++k) {
        const float d = sqdist(ux - kn[k * 3 + 0], uy - kn[k * 3 + 1], uz - kn[k * 3 + 2]);
        if (d < best1) {
          best3 = best2; besti3 = besti2;
          best2 = best1; besti2 = besti1;
          best1 = d; besti1 = k;
        } else if (d < best2) {
          best3 = best2; besti3 = besti2;
          best2 = d; besti2 = k;
        } else if (d < best3) {
          best3 = d; besti3 = k;
        }
      }
      float *od = dist2 + ((size_t)bi * n + i) * 3;
      int *oi = idx + ((size_t)bi * n + i) * 3;
      od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
      oi[0] = besti1; oi[1] = besti2; oi[2] = besti3;
    }
}

/* interpolate_gpu.cu:77-97  out = w0*p0 + w1*p1 + w2*p2, contracted by nvcc to
 * fma(w2,p2, fma(w0,p0, rn(w1*p1)))  (SURVEY.md appendix B).                  */
ORACLE_API void oracle_three_interpolate(int b, int c, int m, int n, const float *points,
                                         const int *idx, const float *weight, float *out) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci) {
      const float *src = points + ((size_t)bi * c + ci) * m;
      float *dst = out + ((size_t)bi * c + ci) * n;
      const int *id = idx + (size_t)bi * n * 3;
      const float *w = weight + (size_t)bi * n * 3;
      for (int i = 0; i < n; ++i) {
        float t = w[i * 3 + 1] * src[id[i * 3 + 1]];
        t = fmaf(w[i * 3 + 0], src[id[i * 3 + 0]], t);
        dst[i] = fmaf(w[i * 3 + 2], src[id[i * 3 + 2]], t);
      }
    }
}

/* interpolate_gpu.cu:120-142 */
ORACLE_API void oracle_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                                              const int *idx, const float *weight,
                                              float *grad_points) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci) {
      const float *g = grad_out + ((size_t)bi * c + ci) * n;
      float *dst = grad_points + ((size_t)bi * c + ci) * m;
      const int *id = idx + (size_t)bi * n * 3;
      const float *w = weight + (size_t)bi * n * 3;
      for (int i = 0; i < n; ++i)
        for (int t = 0; t < 3; ++t) dst[id[i * 3 + t]] += g[i] * w[i * 3 + t];
    }
}

/* ------------------------------------------------------------------------ */
/* lib/utils/iou3d/src/iou3d_kernel.cu                                       */

#define IOU_EPS 1e-8f

typedef struct { float x, y; } pt2;

/* iou3d_kernel.cu:98-102 rotate_around_center; SASS shape:
 * new_x = fma(dx,c, rn(dy*s)) + cx ; new_y = fma(dy,c, -rn(dx*s)) + cy       */
static inline pt2 rot_about(float px, float py, float cx, float cy, float c, float s) {
  const float dx = px - cx, dy = py - cy;
  pt2 r;
  r.x = fmaf(dx, c, dy * s) + cx;
  r.y = fmaf(dy, c, -(dx * s)) + cy;
  return r;
}

/* iou3d_kernel.cu:50-65 check_in_box2d (MARGIN 1e-5, strict comparisons) */
static inline int in_box2d(const float *box, pt2 p) {
  const float MARGIN = 1e-5f;
  const float cx = (box[0] + box[2]) / 2, cy = (box[1] + box[3]) / 2;
  const float c = cosf(-box[4]), s = sinf(-box[4]);
  const pt2 r = rot_about(p.x, p.y, cx, cy, c, s);
  return (r.x > box[0] - MARGIN && r.x < box[2] + MARGIN && r.y > box[1] - MARGIN &&
          r.y < box[3] + MARGIN);
}

/* iou3d_kernel.cu:67-96 intersection(p1,p0,q1,q0,ans) incl. check_rect_cross :42-48 */
static inline int seg_intersection(pt2 p1, pt2 p0, pt2 q1, pt2 q0, pt2 *ans) {
  if (!(fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) &&
        fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) && fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y)))
    return 0;
  /* cross(a,b,o) = (a.x-o.x)*(b.y-o.y) - (b.x-o.x)*(a.y-o.y); left product fused */
  const float s1 = fmaf(q0.x - p0.x, p1.y - p0.y, -((p1.x - p0.x) * (q0.y - p0.y)));
  /* s2 and s5 share their two products in the reference build: neither is fused */
  const float pa = (p1.x - p0.x) * (q1.y - p0.y);
  const float pb = (q1.x - p0.x) * (p1.y - p0.y);
  const float s2 = pa - pb;
  const float s3 = fmaf(p0.x - q0.x, q1.y - q0.y, -((q1.x - q0.x) * (p0.y - q0.y)));
  const float s4 = fmaf(q1.x - q0.x, p1.y - q0.y, -((p1.x - q0.x) * (q1.y - q0.y)));
  if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
  const float s5 = pb - pa;
  if (fabsf(s5 - s1) > IOU_EPS) {
    ans->x = fmaf(s5, q0.x, -(s1 * q1.x)) / (s5 - s1);
    ans->y = fmaf(s5, q0.y, -(s1 * q1.y)) / (s5 - s1);
  } else {
    const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = fmaf(p0.x, p1.y, -(p1.x * p0.y));
    const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = fmaf(q0.x, q1.y, -(q1.x * q0.y));
    const float D = fmaf(a0, b1, -(a1 * b0));
    ans->x = fmaf(b0, c1, -(b1 * c0)) / D;
    ans->y = fmaf(a1, c0, -(a0 * c1)) / D;
  }
  return 1;
}

/* iou3d_kernel.cu:108-212 box_overlap: rotated-rectangle intersection area */
static float box_overlap(const float *box_a, const float *box_b) {
  const float a_x1 = box_a[0], a_y1 = box_a[1], a_x2 = box_a[2], a_y2 = box_a[3], a_ang = box_a[4];
  const float b_x1 = box_b[0], b_y1 = box_b[1], b_x2 = box_b[2], b_y2 = box_b[3], b_ang = box_b[4];
  const float acx = (a_x1 + a_x2) / 2, acy = (a_y1 + a_y2) / 2;
  const float bcx = (b_x1 + b_x2) / 2, bcy = (b_y1 + b_y2) / 2;
  const float ac = cosf(a_ang), as = sinf(a_ang), bc = cosf(b_ang), bs = sinf(b_ang);
  pt2 A[5], Bc[5];
  A[0] = rot_about(a_x1, a_y1, acx, acy, ac, as);
  A[1] = rot_about(a_x2, a_y1, acx, acy, ac, as);
  A[2] = rot_about(a_x2, a_y2, acx, acy, ac, as);
  A[3] = rot_about(a_x1, a_y2, acx, acy, ac, as);
  Bc[0] = rot_about(b_x1, b_y1, bcx, bcy, bc, bs);
  Bc[1] = rot_about(b_x2, b_y1, bcx, bcy, bc, bs);
  Bc[2] = rot_about(b_x2, b_y2, bcx, bcy, bc, bs);
  Bc[3] = rot_about(b_x1, b_y2, bcx, bcy, bc, bs);
  A[4] = A[0];
  Bc[4] = Bc[0];

  pt2 cross_points[16];
  pt2 center = {0.f, 0.f};
  int cnt = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      if (seg_intersection(A[i + 1], A[i], Bc[j + 1], Bc[j], &cross_points[cnt])) {
        center.x = center.x + cross_points[cnt].x;
        center.y = center.y + cross_points[cnt].y;
        ++cnt;
      }
    }
  for (int k = 0; k < 4; ++k) {
    if (in_box2d(box_a, Bc[k])) {
      center.x = center.x + Bc[k].x;
      center.y = center.y + Bc[k].y;
      cross_points[cnt++] = Bc[k];
    }
    if (in_box2d(box_b, A[k])) {
      center.x = center.x + A[k].x;
      center.y = center.y + A[k].y;
      cross_points[cnt++] = A[k];
    }
  }
  center.x /= (float)cnt; /* cnt == 0 -> NaN centre, harmless: area stays 0 */
  center.y /= (float)cnt;
  /* bubble sort, comparator atan2f(a - c) > atan2f(b - c)  (:104-106,188-196) */
  for (int j = 0; j < cnt - 1; ++j)
    for (int i = 0; i < cnt - j - 1; ++i) {
      const float ta = atan2f(cross_points[i].y - center.y, cross_points[i].x - center.x);
      const float tb = atan2f(cross_points[i + 1].y - center.y, cross_points[i + 1].x - center.x);
      if (ta > tb) {
        pt2 t = cross_points[i];
        cross_points[i] = cross_points[i + 1];
        cross_points[i + 1] = t;
      }
    }
  float area = 0.f;
  for (int k = 0; k < cnt - 1; ++k) {
    const float ax = cross_points[k].x - cross_points[0].x, ay = cross_points[k].y - cross_points[0].y;
    const float bx = cross_points[k + 1].x - cross_points[0].x, by = cross_points[k + 1].y - cross_points[0].y;
    area = area + fmaf(ax, by, -(ay * bx));
  }
  return fabsf(area) * 0.5f;
}

/* iou3d_kernel.cu:214-221 iou_bev; sa + sb contracted as fma(sa-factors, rn(sb)) */
static float iou_bev(const float *a, const float *b) {
  const float sb = (b[2] - b[0]) * (b[3] - b[1]);
  const float sab = fmaf(a[2] - a[0], a[3] - a[1], sb);
  const float s = box_overlap(a, b);
  return s / fmaxf(sab - s, IOU_EPS);
}

/* iou3d_kernel.cu:295-303 iou_normal; in the nms_normal kernel Sa is hoisted out
 * of the column loop (rounded) and Sb is the fused product.                   */
static float iou_normal(const float *a, const float *b) {
  const float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  const float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  const float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
  const float interS = width * height;
  const float Sa = (a[2] - a[0]) * (a[3] - a[1]);
  const float Sab = fmaf(b[2] - b[0], b[3] - b[1], Sa);
  return interS / fmaxf(Sab - interS, IOU_EPS);
}

/* iou3d_kernel.cu:223-234 + iou3d.cpp:31-50 */
ORACLE_API void oracle_boxes_overlap_bev(int num_a, const float *boxes_a, int num_b,
                                         const float *boxes_b, float *ans) {
#pragma omp parallel for schedule(dynamic, 4)
  for (int i = 0; i < num_a; ++i)
    for (int j = 0; j < num_b; ++j)
      ans[(size_t)i * num_b + j] = box_overlap(boxes_a + i * 5, boxes_b + j * 5);
}

/* iou3d_kernel.cu:236-248 + iou3d.cpp:52-71 */
ORACLE_API void oracle_boxes_iou_bev(int num_a, const float *boxes_a, int num_b,
                                     const float *boxes_b, float *ans) {
#pragma omp parallel for schedule(dynamic, 4)
  for (int i = 0; i < num_a; ++i)
    for (int j = 0; j < num_b; ++j)
      ans[(size_t)i * num_b + j] = iou_bev(boxes_a + i * 5, boxes_b + j * 5);
}

/* nms_kernel / nms_normal_kernel (iou3d_kernel.cu:250-292,306-348) build a
 * (N, ceil(N/64)) uint64 suppression mask -- box i marks j (j > i inside the
 * diagonal tile, every j in the other tiles) when iou > thresh -- and
 * iou3d.cpp:101-116 scans it greedily: box i is kept unless an earlier kept box
 * marked it; the scan only ORs column words >= i/64, so marks in lower-triangle
 * tiles are never read.  Returns the number kept; keep[] gets their indices.  */
static int nms_impl(const float *boxes, int n, float thresh, int64_t *keep, int rotated) {
  const int col_blocks = (n + 63) / 64;
  uint64_t *mask = (uint64_t *)calloc((size_t)n * col_blocks, sizeof(uint64_t));
#pragma omp parallel for schedule(dynamic, 8)
  for (int i = 0; i < n; ++i) {
    const int rb = i / 64;
    for (int cb = rb; cb < col_blocks; ++cb) {
      const int col_size = (n - cb * 64 < 64) ? n - cb * 64 : 64;
      uint64_t t = 0;
      const int start = (cb == rb) ? (i % 64) + 1 : 0;
      for (int jj = start; jj < col_size; ++jj) {
        const float *bj = boxes + (size_t)(cb * 64 + jj) * 5;
        const float v = rotated ? iou_bev(boxes + (size_t)i * 5, bj) : iou_normal(boxes + (size_t)i * 5, bj);
        if (v > thresh) t |= 1ULL << jj;
      }
      mask[(size_t)i * col_blocks + cb] = t;
    }
  }
  uint64_t *remv = (uint64_t *)calloc(col_blocks, sizeof(uint64_t));
  int num_to_keep = 0;
  for (int i = 0; i < n; ++i) {
    const int nblock = i / 64, inblock = i % 64;
    if (!(remv[nblock] & (1ULL << inblock))) {
      keep[num_to_keep++] = i;
      const uint64_t *p = mask + (size_t)i * col_blocks;
      for (int j = nblock; j < col_blocks; ++j) remv[j] |= p[j];
    }
  }
  free(remv);
  free(mask);
  return num_to_keep;
}

ORACLE_API int oracle_nms(const float *boxes, int n, float thresh, int64_t *keep) {
  return nms_impl(boxes, n, thresh, keep, 1);
}

ORACLE_API int oracle_nms_normal(const float *boxes, int n, float thresh, int64_t *keep) {
  return nms_impl(boxes, n, thresh, keep, 0);
}

/* ------------------------------------------------------------------------ */
/* lib/utils/roipool3d/src/roipool3d_kernel.cu                               */

/* roipool3d_kernel.cu:14-28 pt_in_box3d (GPU flavour).  cy and the y-slab test
 * go through double exactly as written there; the rotation is contracted as
 * x_rot = fma(dx,cos,-rn(dz*sin)), z_rot = fma(dz,cos, rn(dx*sin)) (SASS).    */
static inline int pt_in_box3d_gpu(float x, float y, float z, float cx, float bottom_y, float cz,
                                  float h, float w, float l, float angle, float max_dis) {
  const float cy = (float)((double)bottom_y - (double)h / 2.0);
  if ((fabsf(x - cx) > max_dis) || ((double)fabsf(y - cy) > (double)h / 2.0) ||
      (fabsf(z - cz) > max_dis))
    return 0;
  const float cosa = cosf(angle), sina = sinf(angle);
  const float dx = x - cx, dz = z - cz;
  const float x_rot = fmaf(dx, cosa, -(dz * sina));
  const float z_rot = fmaf(dz, cosa, dx * sina);
  return ((double)x_rot >= -(double)l / 2.0) & ((double)x_rot <= (double)l / 2.0) &
         ((double)z_rot >= -(double)w / 2.0) & ((double)z_rot <= (double)w / 2.0);
}

/* roipool3d.cpp:82-95 pt_in_box3d_cpu (host flavour: g++ -O2 on x86-64 baseline
 * has no FMA, so nothing is contracted).                                      */
static inline int pt_in_box3d_host(float x, float y, float z, float cx, float bottom_y, float cz,
                                   float h, float w, float l, float angle) {
  const float max_dis = 10.0f;
  const float cy = (float)((double)bottom_y - (double)h / 2.0);
  if ((fabsf(x - cx) > max_dis) || ((double)fabsf(y - cy) > (double)h / 2.0) ||
      (fabsf(z - cz) > max_dis))
    return 0;
  const float cosa = cosf(angle), sina = sinf(angle);
  const float x_rot = (x - cx) * cosa + (z - cz) * (-sina);
  const float z_rot = (x - cx) * sina + (z - cz) * cosa;
  return ((double)x_rot >= -(double)l / 2.0) & ((double)x_rot <= (double)l / 2.0) &
         ((double)z_rot >= -(double)w / 2.0) & ((double)z_rot <= (double)w / 2.0);
}

/* roipool3d_kernel.cu:97-194,209-237 (assign_pts_to_box3d + get_pooled_idx +
 * roipool3d_forward): per box the first S inside points in index order, wrapped
 * (k % cnt) when fewer, flag=1 and untouched (caller-zeroed) rows when none.
 * xyz (B,N,3), boxes3d (B,M,7), pts_feature (B,N,C) -> pooled (B,M,S,3+C),
 * flag (B,M) int32.                                                           */
ORACLE_API void oracle_roipool3d(int batch, int n, int m, int c, int s, const float *xyz,
                                 const float *boxes3d, const float *pts_feature,
                                 float *pooled, int *empty_flag) {
#pragma omp parallel for collapse(2) schedule(dynamic, 16)
  for (int bi = 0; bi < batch; ++bi)
    for (int j = 0; j < m; ++j) {
      const float *bx = boxes3d + ((size_t)bi * m + j) * 7;
      const float *p = xyz + (size_t)bi * n * 3;
      const float *f = pts_feature + (size_t)bi * n * c;
      float *o = pooled + ((size_t)bi * m + j) * s * (3 + c);
      int *sel = (int *)malloc(sizeof(int) * (s > 0 ? s : 1));
      int cnt = 0;
      for (int k = 0; k < n && cnt < s; ++k)
        if (pt_in_box3d_gpu(p[k * 3], p[k * 3 + 1], p[k * 3 + 2], bx[0], bx[1], bx[2], bx[3],
                            bx[4], bx[5], bx[6], 10.0f))
          sel[cnt++] = k;
      if (cnt == 0) {
        empty_flag[(size_t)bi * m + j] = 1;
      } else {
        for (int k = cnt; k < s; ++k) sel[k] = sel[k % cnt];
        for (int k = 0; k < s; ++k) {
          float *row = o + (size_t)k * (3 + c);
          for (int t = 0; t < 3; ++t) row[t] = p[sel[k] * 3 + t];
          for (int t = 0; t < c; ++t) row[3 + t] = f[(size_t)sel[k] * c + t];
        }
      }
      free(sel);
    }
}

/* roipool3d.cpp:97-125 pts_in_boxes3d_cpu: flag (M,N) int64 */
ORACLE_API void oracle_pts_in_boxes3d_cpu(int64_t *flag, const float *pts, const float *boxes3d,
                                          int boxes_num, int pts_num) {
  for (int i = 0; i < boxes_num; ++i)
    for (int j = 0; j < pts_num; ++j)
      flag[(size_t)i * pts_num + j] =
          pt_in_box3d_host(pts[j * 3], pts[j * 3 + 1], pts[j * 3 + 2], boxes3d[i * 7],
                           boxes3d[i * 7 + 1], boxes3d[i * 7 + 2], boxes3d[i * 7 + 3],
                           boxes3d[i * 7 + 4], boxes3d[i * 7 + 5], boxes3d[i * 7 + 6]);
}

/* roipool3d.cpp:127-197 roipool3d_cpu: pooled_pts (M,S,3), pooled_features (M,S,C),
 * flag (M) int64 (zeroed here like the reference's memset).                   */
ORACLE_API void oracle_roipool3d_cpu(const float *pts, const float *boxes3d,
                                     const float *pts_feature, float *pooled_pts,
                                     float *pooled_features, int64_t *empty_flag, int boxes_num,
                                     int pts_num, int feature_len, int sampled) {
  memset(empty_flag, 0, sizeof(int64_t) * boxes_num);
  for (int i = 0; i < boxes_num; ++i) {
    int cnt = 0;
    for (int j = 0; j < pts_num; ++j) {
      if (!pt_in_box3d_host(pts[j * 3], pts[j * 3 + 1], pts[j * 3 + 2], boxes3d[i * 7],
                            boxes3d[i * 7 + 1], boxes3d[i * 7 + 2], boxes3d[i * 7 + 3],
                            boxes3d[i * 7 + 4], boxes3d[i * 7 + 5], boxes3d[i * 7 + 6]))
        continue;
      if (cnt >= sampled) break;
      memcpy(pooled_pts + ((size_t)i * sampled + cnt) * 3, pts + j * 3, sizeof(float) * 3);
      memcpy(pooled_features + ((size_t)i * sampled + cnt) * feature_len,
             pts_feature + (size_t)j * feature_len, sizeof(float) * feature_len);
      ++cnt;
    }
    if (cnt == 0) {
      empty_flag[i] = 1;
    } else {
      for (int j = cnt; j < sampled; ++j) {
        memcpy(pooled_pts + ((size_t)i * sampled + j) * 3,
               pooled_pts + ((size_t)i * sampled + j % cnt) * 3, sizeof(float) * 3);
        memcpy(pooled_features + ((size_t)i * sampled + j) * feature_len,
               pooled_features + ((size_t)i * sampled + j % cnt) * feature_len,
               sizeof(float) * feature_len);
      }
    }
  }
}

/* ------------------------------------------------------------------------ */
/* SURVEY.md section 8 "next" rows f2 / f3 (host-side PyTorch code in the      */
/* reference; restated with one IEEE operation per torch elementwise kernel).  */

/* lib/utils/kitti_utils.py:134-147 boxes3d_to_bev_torch: [x,y,z,h,w,l,ry] -> [x1,y1,x2,y2,ry];
 * half extents are l/2, w/2 (exact), the corners one rounded add / subtract each. */
static void box3d_to_bev(const float *b, float *o) {
  const float half_l = b[5] / 2.0f, half_w = b[4] / 2.0f;
  o[0] = b[0] - half_l; o[1] = b[2] - half_w;
  o[2] = b[0] + half_l; o[3] = b[2] + half_w;
  o[4] = b[6];
}

static inline float clamp_min(float v, float lo) { return v < lo ? lo : v; } /* torch.clamp(min=): NaN stays NaN */

/* The DIAGONAL of lib/utils/iou3d/iou3d_utils.py:21-56 boxes_iou3d_gpu(a, b): what the live Stage-2 training
 * keeps of the fg x fg matrix (lib/net/train_functions.py:258-260, :287-289).  boxes (n,7) each. */
ORACLE_API void oracle_boxes_iou3d_aligned(int n, const float *boxes_a, const float *boxes_b, float *iou2d,
                                           float *iou3d) {
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = 0; i < n; ++i) {
    const float *a = boxes_a + (size_t)i * 7, *b = boxes_b + (size_t)i * 7;
    float abev[5], bbev[5];
    box3d_to_bev(a, abev);
    box3d_to_bev(b, bbev);
    const float ov = box_overlap(abev, bbev);                       /* :31-32 */
    const float a_hmin = a[1] - a[3], b_hmin = b[1] - b[3];         /* :35-38 */
    const float max_of_min = a_hmin > b_hmin ? a_hmin : b_hmin;     /* :40 */
    const float min_of_max = a[1] < b[1] ? a[1] : b[1];             /* :41 */
    const float ov_h = clamp_min(min_of_max - max_of_min, 0.0f);    /* :42 */
    const float s_a = a[4] * a[5], s_b = b[4] * b[5];               /* :44-45 */
    const float ssum = s_a + s_b;
    iou2d[i] = ov / clamp_min(ssum - ov, 1e-7f);                    /* :46 */
    const float ov3 = ov * ov_h;                                    /* :48 */
    const float ha = a[3] * a[4], hb = b[3] * b[4];
    const float vol_a = ha * a[5], vol_b = hb * b[5];               /* :50-51 */
    const float vsum = vol_a + vol_b;
    iou3d[i] = ov3 / clamp_min(vsum - ov3, 1e-7f);                  /* :53 */
  }
}

/* lib/utils/distance.py:3 distance_2 on (x, z) pairs: sqrt(sum((a - b) ** 2)) in float32, one rounding per
 * torch kernel (subtract, square, two-term sum, sqrt). */
static inline float bev_dist(float ax, float az, float bx, float bz) {
  const float dx = ax - bx, dz = az - bz;
  const float sx = dx * dx, sz = dz * dz;
  return sqrtf(sx + sz);
}

/* tools/eval_auto.py:263-279 "radius NMS": centres already sorted by descending score; candidate i is kept when
 * its distance to every centre kept so far is > radius (min(...) > 0.3 in the script; a NaN distance fails the
 * comparison, so the candidate is dropped).  The first candidate is always kept.  Returns the number kept. */
ORACLE_API int oracle_radius_nms(const float *centers, int n, float radius, int64_t *keep) {
  int num = 0;
  for (int i = 0; i < n; ++i) {
    int ok = 1;
    if (i > 0) {
      float mn = INFINITY;
      int nan = 0;
      for (int k = 0; k < num; ++k) {
        const int j = (int)keep[k];
        /* prop_prop_distance[keep_id, i] = distance_2(rois, rois)[j, i] = |rois[i] - rois[j]| */
        const float d = bev_dist(centers[2 * i], centers[2 * i + 1], centers[2 * j], centers[2 * j + 1]);
        if (d != d) nan = 1;
        if (d < mn) mn = d;
      }
      ok = !nan && (mn > radius);   /* torch.min propagates NaN; NaN > r is False */
    }
    if (ok) keep[num++] = i;
  }
  return num;
}

/* tools/eval_auto.py:289-291 and :327-343: point_center_distance = distance_2(rpn_center, inputs[:, [0, 2]]),
 * a point belongs to proposal c when that distance is < radius (4.0 in the script); the per-proposal crops keep
 * the points in index order.  idx (m, cap) receives the first `cap` members of every proposal (rest untouched),
 * cnt (m) the full member count, any (n) = 1 where the point is in at least one cylinder (:291). */
ORACLE_API void oracle_cylinder_query(int n, int m, int cap, float radius, const float *pts, const float *centers,
                                      int *idx, int *cnt, unsigned char *any) {
  for (int i = 0; i < n; ++i) any[i] = 0;
  for (int c = 0; c < m; ++c) {
    int k = 0;
    for (int i = 0; i < n; ++i) {
      const float d = bev_dist(centers[2 * c], centers[2 * c + 1], pts[3 * i], pts[3 * i + 2]);
      if (d < radius) {
        if (k < cap) idx[(size_t)c * cap + k] = i;
        ++k;
        any[i] = 1;
      }
    }
    cnt[c] = k;
  }
}

/* Row f4: KittiRCNNDataset.generate_gaussian_training_labels (lib/datasets/kitti_rcnn_dataset.py:529-573) for one
 * scene.  pts (n,3), boxes (g,7).  float32 arithmetic as numpy runs it -- np.power(v, 2) rounds like v * v, the Python
 * float constants meet float32 arrays as float32 -- and the Gaussian in double:
 * multivariate_normal.pdf(d, 0, cov) / (1 / sqrt(2 pi cov)) = exp(-d^2 / (2 cov)).  cls (n) double, reg (n,3) float. */
ORACLE_API void oracle_gaussian_rpn_labels(int n, int g, const float *pts, const float *boxes, float gauss_height,
                                           float gauss_status, double gauss_cov, float fg_radius, double *cls, float *reg) {
  for (int i = 0; i < n; ++i) {
    const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    const float hy = y * gauss_height;
    const float hy2 = hy * hy;
    float centre_dist = 100.0f, best = 0.0f;
    int target = -1;
    for (int k = 0; k < g; ++k) {
      const float dx = x - boxes[7 * k], dz = z - boxes[7 * k + 2];
      const float s = (dx * dx + hy2) + dz * dz;                       /* :545-548 */
      const float d = sqrtf(s);
      float c = d - gauss_status;                                      /* :550 np.clip(d - status, 0, 100) */
      c = c < 0.0f ? 0.0f : (c > 100.0f ? 100.0f : c);
      if (c < centre_dist) centre_dist = c;                            /* np.minimum */
      if (target < 0 || d < best) { best = d; target = k; }            /* :561-562 min / first argmin */
    }
    cls[i] = 0.0;
    reg[3 * i] = reg[3 * i + 1] = reg[3 * i + 2] = 0.0f;
    if (g > 0) {
      cls[i] = exp(-0.5 * (double)centre_dist * (double)centre_dist / gauss_cov);   /* :563-564 */
      if (best < fg_radius) {                                          /* :568-572 */
        reg[3 * i] = boxes[7 * target] - x;
        reg[3 * i + 2] = boxes[7 * target + 2] - z;
      }
    }
  }
}

/* Row f2, loss-side box math.  lib/utils/kitti_utils.py:104-131 boxes3d_to_corners3d_torch for one box:
 * x_corners = +-l/2, y_corners = 0 / -h, z_corners = +-w/2 (:117-119); R = [[cos,0,sin],[0,1,0],[-sin,0,cos]] (:122-126);
 * corners_rotated = R x corners + centre (:128-129).  The float32 batched matmul accumulates k = 0,1,2 in order with
 * FMAs from zero (its middle term multiplies an exact 0 or 1), then the centre is added with one rounding.  `flip`
 * adds (float)pi to ry first (:113-114).  Host libm cosf / sinf against libdevice: last-ulp differences possible. */
static void box_corners(const float *b, int flip, float *out /* 8 x 3 */) {
  const float ry = flip ? b[6] + 3.14159265358979323846f : b[6];
  const float cosa = cosf(ry), sina = sinf(ry);
  const float hl = b[5] / 2.0f, hw = b[4] / 2.0f;
  const float xs[8] = {hl, hl, -hl, -hl, hl, hl, -hl, -hl};
  const float zs[8] = {hw, -hw, -hw, hw, hw, -hw, -hw, hw};
  for (int k = 0; k < 8; ++k) {
    const float yc = k < 4 ? 0.0f : -b[3];
    out[3 * k] = fmaf(sina, zs[k], cosa * xs[k]) + b[0];
    out[3 * k + 1] = yc + b[1];
    out[3 * k + 2] = fmaf(cosa, zs[k], (-sina) * xs[k]) + b[2];
  }
}

ORACLE_API void oracle_boxes3d_to_corners3d(int n, const float *boxes, int flip, float *corners) {
  for (int i = 0; i < n; ++i) box_corners(boxes + (size_t)i * 7, flip, corners + (size_t)i * 24);
}

/* lib/net/train_functions.py:266-271: corner_dist = min(norm(pred_corner - gt_corner), norm(pred_corner - gt_flip_corner))
 * per corner; torch.norm(dim=-1) of a 3-vector = sqrt((dx*dx + dy*dy) + dz*dz), one rounding per step. */
ORACLE_API void oracle_corner_distance(int n, const float *pred, const float *gt, float *dist) {
  for (int i = 0; i < n; ++i) {
    float p[24], g[24], gf[24];
    box_corners(pred + (size_t)i * 7, 0, p);
    box_corners(gt + (size_t)i * 7, 0, g);
    box_corners(gt + (size_t)i * 7, 1, gf);
    for (int k = 0; k < 8; ++k) {
      float d[2];
      const float *q[2] = {g, gf};
      for (int s = 0; s < 2; ++s) {
        const float dx = p[3 * k] - q[s][3 * k], dy = p[3 * k + 1] - q[s][3 * k + 1], dz = p[3 * k + 2] - q[s][3 * k + 2];
        d[s] = sqrtf((dx * dx + dy * dy) + dz * dz);
      }
      dist[(size_t)i * 8 + k] = d[0] < d[1] ? d[0] : d[1];
    }
  }
}
