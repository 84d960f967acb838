// Uniform cell grid over one cloud (counting sort by cell), shared by ball_query_grid.cu and three_nn_grid.cu.
// Everything here has internal linkage: each translation unit gets its own copy of the build kernel.
#pragma once
#include <math.h>

#include "common.cuh"

namespace ws3d {
namespace {

constexpr int kMaxCells = 16384;   // 64 KB of shared-memory counters in the build kernel
constexpr int kMaxDim = 1024;      // per-axis cells (bounds the rounding error of the cell coordinate)

struct GridHdr {
  float ox, oy, oz, inv;
  int dx, dy, dz, ncell;
  float cell, pad0, pad1, pad2;   // cell edge
};

__device__ __forceinline__ bool finite3(float x, float y, float z) {
  return isfinite(x) && isfinite(y) && isfinite(z);
}
// cell coordinate along one axis, clamped to [0, dim-1]; monotone in v
__device__ __forceinline__ int cell_axis(float v, float o, float inv, int dim) {
  float t = __fmul_rn(__fsub_rn(v, o), inv);
  t = fminf(fmaxf(t, 0.f), (float)(dim - 1));
  return (int)t;
}

__global__ void __launch_bounds__(1024, 1) grid_build_kernel(int n, float r_max, int target_cells, const float *__restrict__ xyz,
                                                             GridHdr *__restrict__ hdrs, int *__restrict__ cell_start,
                                                             float4 *__restrict__ sorted) {
  extern __shared__ int s_cnt[];  // kMaxCells
  __shared__ float s_red[6][32];
  __shared__ GridHdr s_hdr;
  __shared__ int s_warp_sum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t cloud = blockIdx.x;
  const float *pts = xyz + cloud * (size_t)n * 3;
  int *cstart = cell_start + cloud * (size_t)(kMaxCells + 1);
  float4 *out = sorted + cloud * (size_t)n;

  // ---- 1. bounding box of the finite points
  const float kInf = __int_as_float(0x7f800000);
  float lo[3] = {kInf, kInf, kInf}, hi[3] = {-kInf, -kInf, -kInf};
  for (int k = tid; k < n; k += 1024) {
    const float x = __ldg(pts + (size_t)k * 3), y = __ldg(pts + (size_t)k * 3 + 1), z = __ldg(pts + (size_t)k * 3 + 2);
    if (finite3(x, y, z)) {
      lo[0] = fminf(lo[0], x); hi[0] = fmaxf(hi[0], x);
      lo[1] = fminf(lo[1], y); hi[1] = fmaxf(hi[1], y);
      lo[2] = fminf(lo[2], z); hi[2] = fmaxf(hi[2], z);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], o));
    }
    if (lane == 0) { s_red[a][warp] = lo[a]; s_red[3 + a][warp] = hi[a]; }
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float l = s_red[a][lane], h = s_red[3 + a][lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        l = fminf(l, __shfl_xor_sync(0xFFFFFFFFu, l, o));
        h = fmaxf(h, __shfl_xor_sync(0xFFFFFFFFu, h, o));
      }
      lo[a] = l; hi[a] = h;
    }
    if (lane == 0) {
      GridHdr h;
      if (!(lo[0] <= hi[0])) {  // no finite point at all
        h.ox = h.oy = h.oz = 0.f; h.inv = 0.f; h.dx = h.dy = h.dz = 1; h.ncell = 1; h.cell = 0.f;
      } else {
        const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
        // r_max > 0: cells at least 1.01 r_max wide (ball query); otherwise as fine as `target_cells` allows
        const int cap = (target_cells > 0 && target_cells < kMaxCells) ? target_cells : kMaxCells;
        float s = r_max > 0.f ? r_max * 1.01f : fmaxf(fmaxf(ex, ey), ez) / (float)kMaxDim;
        s = fmaxf(s, 1e-30f);
        if (!isfinite(s)) s = 3.0e38f;
        int dx, dy, dz;
        for (int iter = 0; iter < 200; ++iter) {
          dx = (int)fminf(ex / s, (float)(kMaxDim - 1)) + 1;
          dy = (int)fminf(ey / s, (float)(kMaxDim - 1)) + 1;
          dz = (int)fminf(ez / s, (float)(kMaxDim - 1)) + 1;
          const bool clipped = ex / s >= (float)kMaxDim || ey / s >= (float)kMaxDim || ez / s >= (float)kMaxDim;
          if (!clipped && (long long)dx * dy * dz <= cap) break;
          s *= 1.26f;  // doubles the cell volume
          if (!isfinite(s)) { s = 3.0e38f; }
        }
        // a per-axis clip means the last cell is longer than s: still a valid (coarser) cell
        if ((long long)dx * dy * dz > kMaxCells) { dx = dy = dz = 1; }
        h.ox = lo[0]; h.oy = lo[1]; h.oz = lo[2];
        h.inv = 1.f / s;
        h.cell = s;
        h.dx = dx; h.dy = dy; h.dz = dz; h.ncell = dx * dy * dz;
      }
      s_hdr = h;
      hdrs[cloud] = h;
    }
  }
  __syncthreads();
  const GridHdr h = s_hdr;

  // ---- 2. histogram
  for (int c = tid; c < h.ncell; c += 1024) s_cnt[c] = 0;
  __syncthreads();
  for (int k = tid; k < n; k += 1024) {
    const float x = __ldg(pts + (size_t)k * 3), y = __ldg(pts + (size_t)k * 3 + 1), z = __ldg(pts + (size_t)k * 3 + 2);
    if (finite3(x, y, z)) {
      const int c = (cell_axis(z, h.oz, h.inv, h.dz) * h.dy + cell_axis(y, h.oy, h.inv, h.dy)) * h.dx +
                    cell_axis(x, h.ox, h.inv, h.dx);
      atomicAdd(&s_cnt[c], 1);
    }
  }
  __syncthreads();

  // ---- 3. exclusive scan of the counters (each thread owns `per` consecutive cells)
  const int per = ceil_div(h.ncell, 1024);
  const int c0 = min(tid * per, h.ncell), c1 = min(c0 + per, h.ncell);
  int local = 0;
  for (int c = c0; c < c1; ++c) local += s_cnt[c];
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp_sum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = s_warp_sum[lane], w = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xFFFFFFFFu, w, o);
      if (lane >= o) w += u;
    }
    s_warp_sum[lane] = w - v;  // exclusive
  }
  __syncthreads();
  int run = s_warp_sum[warp] + incl - local;
  for (int c = c0; c < c1; ++c) {
    const int v = s_cnt[c];
    s_cnt[c] = run;      // becomes the scatter cursor
    cstart[c] = run;
    run += v;
  }
  if (c1 == h.ncell && c0 < c1) cstart[h.ncell] = run;
  if (h.ncell == 0 && tid == 0) cstart[0] = 0;
  __syncthreads();

  // ---- 4. scatter (order inside a cell is arbitrary; the query orders its hits by original index)
  for (int k = tid; k < n; k += 1024) {
    const float x = __ldg(pts + (size_t)k * 3), y = __ldg(pts + (size_t)k * 3 + 1), z = __ldg(pts + (size_t)k * 3 + 2);
    if (finite3(x, y, z)) {
      const int c = (cell_axis(z, h.oz, h.inv, h.dz) * h.dy + cell_axis(y, h.oy, h.inv, h.dy)) * h.dx +
                    cell_axis(x, h.ox, h.inv, h.dx);
      const int pos = atomicAdd(&s_cnt[c], 1);
      out[pos] = make_float4(x, y, z, __int_as_float(k));
    }
  }
}


}  // namespace
}  // namespace ws3d
