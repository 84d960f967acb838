// three_nn through a cell grid over the KNOWN points (exact: same indices and distances as the full scan).
//
// The reference (pointnet2_lib/pointnet2/src/interpolate_gpu.cu:26-52) scans all m known points for every
// unknown point (FP0: 16384 x 4096 = 67 M distance tests per cloud) and keeps the three nearest with a strict
// '<' cascade in index order, i.e. the three smallest (distance, index) pairs.  Here the known points are
// binned into cells (cell_grid.cuh, ~1 point per cell), and each unknown point visits the cells round its
// own in rings of growing Chebyshev radius r.  Once ring r is done every unvisited point is at least
// (r - eps) cell edges away, so the search stops as soon as the third best squared distance is strictly
// below that bound: no unvisited point can tie or beat it.  Distances use the reference arithmetic and the
// (distance, index) order is applied explicitly, so the result is bit-identical whatever the visiting order.
#include <limits.h>
#include <math.h>
#include <stdlib.h>

#include "cell_grid.cuh"
#include "common.cuh"

namespace ws3d {
namespace {

constexpr int kThreads = 128;

struct Best3 {
  float d1, d2, d3;
  int i1, i2, i3;
};

// (d, k) is inserted if it precedes a kept pair in (distance, index) order
__device__ __forceinline__ void insert3(Best3 &b, float d, int k) {
  // cheap reject first (one compare for the common case; NaN fails it): half of this kernel's instructions were this cascade
  if (d <= b.d3 && (d < b.d3 || k < b.i3)) {
    if (d < b.d1 || (d == b.d1 && k < b.i1)) {
      b.d3 = b.d2; b.i3 = b.i2; b.d2 = b.d1; b.i2 = b.i1; b.d1 = d; b.i1 = k;
    } else if (d < b.d2 || (d == b.d2 && k < b.i2)) {
      b.d3 = b.d2; b.i3 = b.i2; b.d2 = d; b.i2 = k;
    } else {
      b.d3 = d; b.i3 = k;
    }
  }
}

__global__ void __launch_bounds__(kThreads) three_nn_grid_kernel(int n, int m, const float *__restrict__ unknown,
                                                                  const GridHdr *__restrict__ hdrs,
                                                                  const int *__restrict__ cell_start,
                                                                  const float4 *__restrict__ sorted,
                                                                  float *__restrict__ dist2, int *__restrict__ idx,
                                                                  float *__restrict__ weight) {
  const size_t cloud = blockIdx.y;
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const GridHdr h = hdrs[cloud];
  const int *cstart = cell_start + cloud * (size_t)(kMaxCells + 1);
  const float4 *pts = sorted + cloud * (size_t)m;
  const float *u = unknown + (cloud * (size_t)n + i) * 3;
  const float ux = __ldg(u), uy = __ldg(u + 1), uz = __ldg(u + 2);

  const float kInf = __int_as_float(0x7f800000);
  Best3 b;
  b.d1 = b.d2 = b.d3 = kInf;  // the reference's (float)1e40
  b.i1 = b.i2 = b.i3 = 0;

  if (isfinite(ux) && isfinite(uy) && isfinite(uz)) {  // otherwise every distance is inf/NaN: nothing is ever kept
    const int cx = cell_axis(ux, h.ox, h.inv, h.dx), cy = cell_axis(uy, h.oy, h.inv, h.dy), cz = cell_axis(uz, h.oz, h.inv, h.dz);
    const int rmax = max(max(max(cx, h.dx - 1 - cx), max(cy, h.dy - 1 - cy)), max(cz, h.dz - 1 - cz));
    for (int r = 0; r <= rmax; ++r) {
      const int z0 = max(cz - r, 0), z1 = min(cz + r, h.dz - 1);
      const int y0 = max(cy - r, 0), y1 = min(cy + r, h.dy - 1);
      for (int zz = z0; zz <= z1; ++zz) {
        const bool zface = (zz == cz - r) || (zz == cz + r);
        for (int yy = y0; yy <= y1; ++yy) {
          const bool face = zface || (yy == cy - r) || (yy == cy + r);
          const int rowbase = (zz * h.dy + yy) * h.dx;
          // on a y/z face of the ring the whole x run belongs to it; elsewhere only its two end cells
          const int nseg = face ? 1 : (r == 0 ? 1 : 2);
          for (int sgm = 0; sgm < nseg; ++sgm) {
            int xa, xb;
            if (face) { xa = cx - r; xb = cx + r; }
            else if (sgm == 0) { xa = xb = cx - r; }
            else { xa = xb = cx + r; }
            xa = max(xa, 0); xb = min(xb, h.dx - 1);
            if (xa > xb) continue;
            if (!face && ((sgm == 0 && cx - r < 0) || (sgm == 1 && cx + r > h.dx - 1))) continue;  // end cell outside the grid
            const int p0 = __ldg(cstart + rowbase + xa), p1 = __ldg(cstart + rowbase + xb + 1);
            for (int p = p0; p < p1; ++p) {
              const float4 q = __ldg(pts + p);
              insert3(b, sqdist_ref(ux - q.x, uy - q.y, uz - q.z), __float_as_int(q.w));
            }
          }
        }
      }
      // every point not yet visited lies at least (r - 0.01) cell edges away (0.01 covers the rounding of the cell
      // coordinate); strict '<' with a relative margin: such a point can neither beat nor tie the third best
      if (r >= 1) {
        const float reach = ((float)r - 0.01f) * h.cell;
        if (b.d3 < reach * reach * 0.9999f) break;
      }
    }
  }
  store_three_nn(dist2, idx, weight, cloud * (size_t)n + i, b.d1, b.d2, b.d3, b.i1, b.i2, b.i3);
}

}  // namespace

bool three_nn_grid_applicable(int b, int n, int m) {
  static const int enabled = []() { const char *e = getenv("WS3D_NN_GRID"); return (e && *e) ? atoi(e) : 1; }();
  return enabled && m >= 512 && m <= 65536 && n >= 1024 && b >= 1 && b <= 65535;
}

int three_nn_grid(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, float *weight,
                  cudaStream_t stream) {
  const size_t hdr_bytes = ((size_t)b * sizeof(GridHdr) + 255) & ~(size_t)255;
  const size_t start_bytes = ((size_t)b * (kMaxCells + 1) * sizeof(int) + 255) & ~(size_t)255;
  const size_t sorted_bytes = (size_t)b * m * sizeof(float4);
  char *ws = (char *)scratch(hdr_bytes + start_bytes + sorted_bytes, 5);
  if (!ws) return (int)cudaErrorMemoryAllocation;
  GridHdr *hdrs = (GridHdr *)ws;
  int *cell_start = (int *)(ws + hdr_bytes);
  float4 *sorted = (float4 *)(ws + hdr_bytes + start_bytes);
  cudaError_t e = cudaFuncSetAttribute(grid_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxCells * (int)sizeof(int));
  if (e != cudaSuccess) { set_error("three_nn grid: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  // about two cells per known point: a 3x3x3 block then holds the three nearest in the common case
  int target = 2 * m;
  if (target > kMaxCells) target = kMaxCells;
  grid_build_kernel<<<b, 1024, kMaxCells * sizeof(int), stream>>>(m, 0.f, target, known, hdrs, cell_start, sorted);
  int rc = check_launch("three_nn (grid build)");
  if (rc) return rc;
  dim3 grid((unsigned)ceil_div(n, kThreads), (unsigned)b);
  three_nn_grid_kernel<<<grid, kThreads, 0, stream>>>(n, m, unknown, hdrs, cell_start, sorted, dist2, idx, weight);
  return check_launch("three_nn (grid)");
}

}  // namespace ws3d
