// ball_query through a uniform cell grid (exact: same hits, same order as the reference scan).
//
// The reference (pointnet2_lib/pointnet2/src/ball_query_gpu.cu:23-44) tests every point of the cloud
// against every centre: 67 M distance tests per 16384-point cloud, of which a few dozen per centre can
// succeed.  For clouds that are large against the ball (SA1: n = 16384, r <= 0.5 m in an 80 x 70 m
// scene; SA2: n = 4096) this file prunes the scan with a cell list and keeps the result bit-identical:
//
//   build (one 1024-thread CTA per cloud, everything in shared memory):
//     bounding box -> cell edge s >= 1.01 * r_max, enlarged until the grid has <= kMaxCells cells ->
//     counting sort of the points by cell into `sorted` (x, y, z, original index).
//   query (one warp per centre): the 3 x 3 x 3 block of cells round the centre is 9 contiguous runs of
//     `sorted` (x is the fastest cell axis).  Lanes stride over the concatenated runs with 16-byte
//     coalesced loads, apply the reference's test (same rounding order, strict <), compact the hits by
//     ballot into a per-warp buffer, and finally order them by original index (rank sort): the
//     reference returns the first `nsample` hits in index order and pads with the first hit.
//     A ball holding more than kHitCap points takes the exact early-exit scan over the original order
//     instead (it terminates quickly precisely because the ball is dense).
//
// Why the pruning is exact: a point with computed d2 < r2 has |dx| <= r (1 + 2e-6) per axis; the cell
// coordinate t(x) = fl(fl(x - origin) * inv) is monotone in x with absolute error <= dim * 2^-22
// <= 2.5e-4 (dim <= 1024), and s >= 1.01 r leaves a margin of 1e-2 cell: the point's cell is at most
// one cell away from the centre's on every axis.  Non-finite points can never pass the test and are left
// out of the grid; non-finite centres find nothing (as in the reference, where NaN/inf compare false).
#include <math.h>

#include "cell_grid.cuh"
#include "common.cuh"

namespace ws3d {
namespace {

constexpr int kHitCap = 128;       // buffered hits per (warp, radius)
constexpr int kQueryWarps = 8;

template <int NR>
struct GridQueryParams {
  int n, m;
  float r2[NR];
  int nsample[NR];
  const float *new_xyz;
  const float *xyz;
  const GridHdr *hdrs;
  const int *cell_start;
  const float4 *sorted;
  int *idx[NR];
};

// Exact early-exit scan over the original point order (used when a ball overflows the hit buffer).
__device__ __forceinline__ void scan_in_order(const float *__restrict__ pts, int n, float qx, float qy, float qz, float r2,
                                              int K, int *__restrict__ row, int lane) {
  const uint32_t lt_mask = (1u << lane) - 1u;
  int cnt = 0, first = 0;
  for (int base = 0; base < n && cnt < K; base += 32) {
    const int k = base + lane;
    bool hit = false;
    if (k < n) {
      const float x = __ldg(pts + (size_t)k * 3), y = __ldg(pts + (size_t)k * 3 + 1), z = __ldg(pts + (size_t)k * 3 + 2);
      hit = sqdist_ref(qx - x, qy - y, qz - z) < r2;
    }
    const uint32_t hits = __ballot_sync(0xFFFFFFFFu, hit);
    if (hits) {
      if (cnt == 0) first = base + __ffs(hits) - 1;
      const int pos = cnt + __popc(hits & lt_mask);
      if (hit && pos < K) row[pos] = k;
      cnt += __popc(hits);
    }
  }
  if (cnt > 0)
    for (int s = cnt + lane; s < K; s += 32) row[s] = first;
}

template <int NR>
__global__ void __launch_bounds__(kQueryWarps * 32) grid_query_kernel(GridQueryParams<NR> prm) {
  __shared__ int s_hits[kQueryWarps][NR][kHitCap];
  __shared__ int s_run_start[kQueryWarps][9], s_run_pref[kQueryWarps][10];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t cloud = blockIdx.y;
  const int n = prm.n, m = prm.m;
  const GridHdr h = prm.hdrs[cloud];
  const int *cstart = prm.cell_start + cloud * (size_t)(kMaxCells + 1);
  const float4 *sorted = prm.sorted + cloud * (size_t)n;
  const float *pts = prm.xyz + cloud * (size_t)n * 3;
  const uint32_t lt_mask = (1u << lane) - 1u;

  for (int q = blockIdx.x * kQueryWarps + warp; q < m; q += gridDim.x * kQueryWarps) {
    const float *qp = prm.new_xyz + (cloud * (size_t)m + q) * 3;
    const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
    if (!finite3(qx, qy, qz)) {  // nothing can be inside the ball: zero rows (what the reference's caller pre-fills)
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        int *row = prm.idx[r] + (cloud * (size_t)m + q) * prm.nsample[r];
        for (int s = lane; s < prm.nsample[r]; s += 32) row[s] = 0;
      }
      continue;
    }
    // ---- the 9 runs of cells (x-1..x+1 contiguous) round the centre's cell
    {
      // unclamped float cell coordinates, limited so that the int conversion cannot overflow
      const float tx = fminf(fmaxf(__fmul_rn(__fsub_rn(qx, h.ox), h.inv), -2.f), (float)h.dx + 1.f);
      const float ty = fminf(fmaxf(__fmul_rn(__fsub_rn(qy, h.oy), h.inv), -2.f), (float)h.dy + 1.f);
      const float tz = fminf(fmaxf(__fmul_rn(__fsub_rn(qz, h.oz), h.inv), -2.f), (float)h.dz + 1.f);
      // points are binned with clamping to [0, dim-1]; do the same to the centre (a centre outside the box by
      // more than a cell has no neighbour, and the clamped block is then merely a superset)
      const int cx = min(max((int)floorf(tx), 0), h.dx - 1);
      const int cy = min(max((int)floorf(ty), 0), h.dy - 1);
      const int cz = min(max((int)floorf(tz), 0), h.dz - 1);
      int len = 0, start = 0;
      if (lane < 9) {
        const int yy = cy + (lane % 3) - 1, zz = cz + (lane / 3) - 1;
        if (yy >= 0 && yy < h.dy && zz >= 0 && zz < h.dz) {
          const int x0 = max(cx - 1, 0), x1 = min(cx + 1, h.dx - 1);
          const int rowbase = (zz * h.dy + yy) * h.dx;
          start = __ldg(cstart + rowbase + x0);
          len = __ldg(cstart + rowbase + x1 + 1) - start;
        }
      }
      int incl = len;
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
      }
      __syncwarp();
      if (lane < 9) { s_run_start[warp][lane] = start; s_run_pref[warp][lane + 1] = incl; }
      if (lane == 0) s_run_pref[warp][0] = 0;
      __syncwarp();
    }
    const int total = s_run_pref[warp][9];

    int cnt[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) cnt[r] = 0;
    for (int j0 = 0; j0 < total; j0 += 32) {
      const int j = j0 + lane;
      float d2 = __int_as_float(0x7f800000);
      int k = 0;
      if (j < total) {
        int run = 0;
#pragma unroll
        for (int t = 1; t < 9; ++t) run += (j >= s_run_pref[warp][t]) ? 1 : 0;
        const float4 p = __ldg(sorted + s_run_start[warp][run] + (j - s_run_pref[warp][run]));
        d2 = sqdist_ref(qx - p.x, qy - p.y, qz - p.z);
        k = __float_as_int(p.w);
      }
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const bool hit = d2 < prm.r2[r];
        const uint32_t hits = __ballot_sync(0xFFFFFFFFu, hit);
        const int pos = cnt[r] + __popc(hits & lt_mask);
        if (hit && pos < kHitCap) s_hits[warp][r][pos] = k;
        cnt[r] += __popc(hits);
      }
    }
    __syncwarp();

#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int K = prm.nsample[r];
      int *row = prm.idx[r] + (cloud * (size_t)m + q) * K;
      if (cnt[r] == 0) {
        for (int s = lane; s < K; s += 32) row[s] = 0;
        continue;
      }
      if (cnt[r] > kHitCap) {
        scan_in_order(pts, n, qx, qy, qz, prm.r2[r], K, row, lane);
        continue;
      }
      // rank sort by original index; the first K in index order are the reference's answer
      const int c = cnt[r];
      int first = INT_MAX;
      for (int h0 = 0; h0 < c; h0 += 32) {
        const int hh = h0 + lane;
        const int mine = hh < c ? s_hits[warp][r][hh] : INT_MAX;
        int rank = 0;
        for (int t = 0; t < c; ++t) rank += (s_hits[warp][r][t] < mine) ? 1 : 0;
        if (hh < c && rank < K) row[rank] = mine;
        first = min(first, mine);
      }
      first = __reduce_min_sync(0xFFFFFFFFu, first);
      for (int s = c + lane; s < K; s += 32) row[s] = first;
    }
    __syncwarp();
  }
}

template <int NR>
int launch_grid(int b, int n, int m, const float *radius, const int *nsample, const float *new_xyz, const float *xyz,
                int *const *idx, cudaStream_t stream) {
  const size_t hdr_bytes = ((size_t)b * sizeof(GridHdr) + 255) & ~(size_t)255;
  const size_t start_bytes = ((size_t)b * (kMaxCells + 1) * sizeof(int) + 255) & ~(size_t)255;
  const size_t sorted_bytes = (size_t)b * n * sizeof(float4);
  char *ws = (char *)scratch(hdr_bytes + start_bytes + sorted_bytes, 4);
  if (!ws) return (int)cudaErrorMemoryAllocation;
  GridHdr *hdrs = (GridHdr *)ws;
  int *cell_start = (int *)(ws + hdr_bytes);
  float4 *sorted = (float4 *)(ws + hdr_bytes + start_bytes);

  float r_max = 0.f;
  for (int r = 0; r < NR; ++r) r_max = fmaxf(r_max, fabsf(radius[r]));
  {
    cudaError_t e = cudaFuncSetAttribute(grid_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxCells * (int)sizeof(int));
    if (e != cudaSuccess) { set_error("ball_query grid: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  }
  grid_build_kernel<<<b, 1024, kMaxCells * sizeof(int), stream>>>(n, r_max, 0, xyz, hdrs, cell_start, sorted);
  int rc = check_launch("ball_query (grid build)");
  if (rc) return rc;

  GridQueryParams<NR> prm;
  prm.n = n; prm.m = m; prm.new_xyz = new_xyz; prm.xyz = xyz; prm.hdrs = hdrs; prm.cell_start = cell_start; prm.sorted = sorted;
  for (int r = 0; r < NR; ++r) {
    prm.r2[r] = radius[r] * radius[r];  // f32 product, as ball_query_gpu.cu:23
    prm.nsample[r] = nsample[r];
    prm.idx[r] = idx[r];
  }
  int gx = ceil_div(m, kQueryWarps);
  const int cap = ceil_div(8 * num_sms(), b);  // ~8 CTAs per SM in flight
  if (gx > cap) gx = cap;
  grid_query_kernel<NR><<<dim3((unsigned)gx, (unsigned)b), kQueryWarps * 32, 0, stream>>>(prm);
  return check_launch("ball_query (grid query)");
}

}  // namespace

// Clouds for which the cell grid pays: enough points that the full scan dominates, few enough that the
// 1024-thread build CTA per cloud is short.  Every radius must be positive and finite.
bool ball_query_grid_applicable(int nr, int b, int n, int m, const float *radius) {
  if (n < 2048 || n > 65536 || m < 1 || b < 1 || b > 65535) return false;
  for (int r = 0; r < nr; ++r)
    if (!(radius[r] > 0.f) || !isfinite(radius[r])) return false;
  return true;
}

int ball_query_grid(int nr, int b, int n, int m, const float *radius, const int *nsample, const float *new_xyz,
                    const float *xyz, int *const *idx, cudaStream_t stream) {
  if (nr == 1) return launch_grid<1>(b, n, m, radius, nsample, new_xyz, xyz, idx, stream);
  if (nr == 2) return launch_grid<2>(b, n, m, radius, nsample, new_xyz, xyz, idx, stream);
  return fail_arg("ball_query (1 or 2 radii)");
}

}  // namespace ws3d
