// Host (CPU) entry points of the reference's roipool3d_cuda module: pts_in_boxes3d_cpu and
// roipool3d_cpu (lib/utils/roipool3d/src/roipool3d.cpp:82-197).  They are part of the module's
// public surface (the dataset code and the annotation tool call them), so the drop-in provides
// them natively.  Compiled with -ffp-contract=off: the reference's host build (g++ -O2, x86-64
// baseline) contains no fused multiply-adds.
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../../include/ws3d_ops.h"

#define WS3D_API extern "C" __attribute__((visibility("default")))

namespace {
inline int pt_in_box3d_host(float x, float y, float z, float cx, float bottom_y, float cz, float h, float w, float l,
                            float angle) {
  const float max_dis = 10.0f;
  const float cy = (float)((double)bottom_y - (double)h / 2.0);
  if ((std::fabs(x - cx) > max_dis) || ((double)std::fabs(y - cy) > (double)h / 2.0) || (std::fabs(z - cz) > max_dis))
    return 0;
  const float cosa = std::cos(angle), sina = std::sin(angle);
  const float x_rot = (x - cx) * cosa + (z - cz) * (-sina);
  const float z_rot = (x - cx) * sina + (z - cz) * cosa;
  return ((double)x_rot >= -(double)l / 2.0) & ((double)x_rot <= (double)l / 2.0) &
         ((double)z_rot >= -(double)w / 2.0) & ((double)z_rot <= (double)w / 2.0);
}
}  // namespace

// roipool3d.cpp:97-125.  pts_flag (M,N) int64 <- 0/1; pts (N,3); boxes3d (M,7).  All HOST memory.
WS3D_API int ws3d_pts_in_boxes3d_cpu(int64_t *pts_flag_host, const float *pts_host, const float *boxes3d_host,
                                     int boxes_num, int pts_num) {
  if (boxes_num < 0 || pts_num < 0) return 1;
  for (int i = 0; i < boxes_num; ++i) {
    const float *b = boxes3d_host + (size_t)i * 7;
    for (int j = 0; j < pts_num; ++j) {
      const float *p = pts_host + (size_t)j * 3;
      pts_flag_host[(size_t)i * pts_num + j] = pt_in_box3d_host(p[0], p[1], p[2], b[0], b[1], b[2], b[3], b[4], b[5], b[6]);
    }
  }
  return 0;
}

// roipool3d.cpp:127-197.  pooled_pts (M,S,3), pooled_features (M,S,C), pooled_empty_flag (M) int64.
WS3D_API int ws3d_roipool3d_cpu(const float *pts_host, const float *boxes3d_host, const float *pts_feature_host,
                                float *pooled_pts_host, float *pooled_features_host, int64_t *pooled_empty_flag_host,
                                int boxes_num, int pts_num, int feature_len, int sampled_pts_num) {
  if (boxes_num < 0 || pts_num < 0 || feature_len < 0 || sampled_pts_num < 0) return 1;
  const size_t S = (size_t)sampled_pts_num, C = (size_t)feature_len;
  std::memset(pooled_empty_flag_host, 0, sizeof(int64_t) * (size_t)boxes_num);
  for (int i = 0; i < boxes_num; ++i) {
    const float *b = boxes3d_host + (size_t)i * 7;
    size_t cnt = 0;
    for (int j = 0; j < pts_num; ++j) {
      const float *p = pts_host + (size_t)j * 3;
      if (!pt_in_box3d_host(p[0], p[1], p[2], b[0], b[1], b[2], b[3], b[4], b[5], b[6])) continue;
      if (cnt >= S) break;
      std::memcpy(pooled_pts_host + ((size_t)i * S + cnt) * 3, p, sizeof(float) * 3);
      std::memcpy(pooled_features_host + ((size_t)i * S + cnt) * C, pts_feature_host + (size_t)j * C, sizeof(float) * C);
      ++cnt;
    }
    if (cnt == 0) {
      pooled_empty_flag_host[i] = 1;
      continue;
    }
    for (size_t j = cnt; j < S; ++j) {
      std::memcpy(pooled_pts_host + ((size_t)i * S + j) * 3, pooled_pts_host + ((size_t)i * S + j % cnt) * 3, sizeof(float) * 3);
      std::memcpy(pooled_features_host + ((size_t)i * S + j) * C, pooled_features_host + ((size_t)i * S + j % cnt) * C,
                  sizeof(float) * C);
    }
  }
  return 0;
}
