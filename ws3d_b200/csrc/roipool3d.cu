// roipool3d for B200: per-proposal point crop + feature gather in one kernel.
//
// Replaces lib/utils/roipool3d/src/roipool3d_kernel.cu:97-237: the reference writes a dense
// (B,N,M) int32 inside-flag tensor (1 GiB at N = M = 16384) with stride-M stores, re-scans it
// serially per box, and gathers output rows with one thread per (box, sample) so that
// neighbouring threads write rows 4*(3+C) bytes apart; it cudaMallocs/cudaFrees both scratch
// tensors on every call and runs on the legacy default stream.
//
// Design: the cloud's xyz is staged once per CTA into shared memory by TMA bulk copies; one WARP
// owns one box and scans the points 32 at a time (packed xyz, stride-3 words = conflict free),
// compacting the inside lanes in point order with ballot + popc until S are found; the selected
// indices stay in shared memory and the same warp then streams the (S, 3+C) output rows with
// coalesced stores (each row is one contiguous 4*(3+C)-byte line, wrap-around duplicates are
// re-reads that hit L2).  No flag tensor, no scratch allocation, caller's stream.
//
// Exactness: pt_in_box3d (roipool3d_kernel.cu:14-28) mixes float and double.  cy is computed in
// double exactly as written there (once per box); the remaining double comparisons compare a float
// against h/2, l/2, w/2, which are exact in float, so float comparisons give identical results.
// The rotation uses the reference build's FMA shape: x_rot = fma(dx,cos,-rn(dz*sin)),
// z_rot = fma(dz,cos,rn(dx*sin)).
//
// Two-kernel path (clouds of 1024..65536 points, S <= 1024: every WS3D call):
//   select  -- the points are binned into a uniform cell grid (cell_grid.cuh, shared with ball_query / three_nn); one warp
//              per box visits only the cells under the box's padded axis-aligned bound (a car-sized box: ~300 candidate
//              points instead of 16384), applies the reference predicate, orders the hits by original index (the
//              reference keeps the first S in index order) and writes the S source indices -- wrap-around duplicates
//              `k % cnt` expanded -- to an L2-resident scratch (B, M, S).  Boxes with more hits than the buffer holds
//              take the exact early-exit scan in index order (it ends quickly precisely because the box is dense).
//   write   -- ALL warps stream the (B, M, S, 3+C) tensor: a box's block is S*(3+C) contiguous floats that starts on a
//              16-byte boundary, so every thread assembles four consecutive floats (crossing row ends where needed) and
//              issues one 16-byte streaming store; a warp instruction writes 512 contiguous bytes whatever 3+C is.
// The one-kernel scan below remains for shapes outside that range.
#include <stdlib.h>

#include "cell_grid.cuh"
#include "common.cuh"
#include "tma.cuh"

namespace ws3d {
namespace {

constexpr int kChunkPts = 16384;  // points staged per pass (192 KB); larger clouds loop
constexpr int kMaxWarps = 16;
constexpr int kMaxSel = 512;      // selected-index slots per warp kept in shared memory

struct RoiParams {
  int n, m, c, s;
  const float *xyz;       // (B,N,3)
  const float *boxes3d;   // (B,M,7)
  const float *feat;      // (B,N,C)
  float *pooled;          // (B,M,S,3+C)
  int *empty_flag;        // (B,M)
  int *sel_g;             // global spill for selected indices when S > kMaxSel (B,M,S) or null
  int boxes_per_cta;
};

__global__ void __launch_bounds__(kMaxWarps * 32) roipool3d_kernel(RoiParams prm) {
  extern __shared__ __align__(16) float s_dyn[];
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ int s_next;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int n = prm.n, m = prm.m, c = prm.c, S = prm.s;
  const size_t cloud = blockIdx.y;
  const int j0 = blockIdx.x * prm.boxes_per_cta, j1 = min(m, j0 + prm.boxes_per_cta);
  if (j0 >= j1) return;
  const int chunk_cap = min(n, kChunkPts);
  float *s_xyz = s_dyn;                                                   // chunk_cap*3 floats
  int *s_sel = reinterpret_cast<int *>(s_dyn + (size_t)((chunk_cap * 3 + 3) & ~3));  // nwarps*min(S,kMaxSel)
  const int sel_cap = min(S, kMaxSel);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&s_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_next = 0;
  }
  __syncthreads();
  const float *pts = prm.xyz + cloud * (size_t)n * 3;
  const bool single = n <= kChunkPts;
  if (single) stage_floats(s_xyz, pts, n * 3, &s_bar, 0);
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int row_len = 3 + c;
  uint32_t stage_parity = 0;

  // Boxes are handed out one group (one box per warp) at a time so that multi-chunk clouds can
  // re-stage the points between groups with CTA-wide barriers.
  for (int g0 = j0; g0 < j1; g0 += nwarps) {
    const int j = g0 + warp;
    const bool active = j < j1;
    float cx = 0, cy = 0, cz = 0, hh = 0, hw = 0, hl = 0, cosa = 1, sina = 0;
    if (active) {
      const float *bx = prm.boxes3d + (cloud * (size_t)m + j) * 7;
      const float bot = __ldg(bx + 1), h = __ldg(bx + 3), w = __ldg(bx + 4), l = __ldg(bx + 5), ang = __ldg(bx + 6);
      cx = __ldg(bx); cz = __ldg(bx + 2);
      cy = (float)((double)bot - (double)h / 2.0);  // roipool3d_kernel.cu:18
      hh = __fmul_rn(h, 0.5f); hw = __fmul_rn(w, 0.5f); hl = __fmul_rn(l, 0.5f);
      cosa = cosf(ang); sina = sinf(ang);
    }
    int *sel = prm.sel_g ? prm.sel_g + (cloud * (size_t)m + (active ? j : 0)) * S : s_sel + warp * sel_cap;
    int cnt = 0;
    for (int c0 = 0; c0 < n; c0 += kChunkPts) {
      const int cn = min(kChunkPts, n - c0);
      if (!single) {
        __syncthreads();  // everyone is done with the previous chunk
        stage_floats(s_xyz, pts + (size_t)c0 * 3, cn * 3, &s_bar, stage_parity);
        stage_parity ^= 1u;
      }
      if (active) {
        for (int base = 0; base < cn && cnt < S; base += 32) {
          const int k = base + lane;
          bool in = false;
          if (k < cn) {
            const float x = s_xyz[k * 3], y = s_xyz[k * 3 + 1], z = s_xyz[k * 3 + 2];
            const float dx = __fsub_rn(x, cx), dz = __fsub_rn(z, cz);
            const bool pre = !(fabsf(dx) > 10.0f) && !(fabsf(__fsub_rn(y, cy)) > hh) && !(fabsf(dz) > 10.0f);
            const float x_rot = __fmaf_rn(dx, cosa, -__fmul_rn(dz, sina));
            const float z_rot = __fmaf_rn(dz, cosa, __fmul_rn(dx, sina));
            in = pre && (x_rot >= -hl) && (x_rot <= hl) && (z_rot >= -hw) && (z_rot <= hw);
          }
          const uint32_t hits = __ballot_sync(0xFFFFFFFFu, in);
          if (hits) {
            const int pos = cnt + __popc(hits & lt_mask);
            if (in && pos < S) sel[pos] = c0 + k;
            cnt += __popc(hits);
          }
        }
      }
    }
    if (active) {
      __syncwarp();
      if (cnt == 0) {
        if (lane == 0) prm.empty_flag[cloud * (size_t)m + j] = 1;
      } else {
        const int have = min(cnt, S);
        float *dst = prm.pooled + (cloud * (size_t)m + j) * (size_t)S * row_len;
        const float *feat = prm.feat + cloud * (size_t)n * c;
        if (row_len >= 32) {
          // one row per step: lanes stride over the 3+C channels
          for (int k = 0; k < S; ++k) {
            const int src = sel[k < have ? k : k % have];
            float *row = dst + (size_t)k * row_len;
            for (int t = lane; t < row_len; t += 32)
              __stcs(row + t, t < 3 ? __ldg(pts + (size_t)src * 3 + t) : __ldg(feat + (size_t)src * c + (t - 3)));
          }
        } else {
          // short rows: flatten (sample, channel) over the lanes
          const int total = S * row_len;
          for (int e = lane; e < total; e += 32) {
            const int k = e / row_len, t = e - k * row_len;
            const int src = sel[k < have ? k : k % have];
            __stcs(dst + e, t < 3 ? __ldg(pts + (size_t)src * 3 + t) : __ldg(feat + (size_t)src * c + (t - 3)));
          }
        }
      }
    }
  }
}

// ---- two-kernel path ------------------------------------------------------------------------------
constexpr int kSelWarps = 8;
constexpr int kSelCap = 1024;     // buffered hits per warp (4 KB of shared memory each)

struct BoxGeom { float cx, cy, cz, hh, hw, hl, cosa, sina; };

__device__ __forceinline__ BoxGeom box_geom(const float *__restrict__ bx) {
  BoxGeom g;
  const float bot = __ldg(bx + 1), h = __ldg(bx + 3), w = __ldg(bx + 4), l = __ldg(bx + 5), ang = __ldg(bx + 6);
  g.cx = __ldg(bx); g.cz = __ldg(bx + 2);
  g.cy = (float)((double)bot - (double)h / 2.0);  // roipool3d_kernel.cu:18
  g.hh = __fmul_rn(h, 0.5f); g.hw = __fmul_rn(w, 0.5f); g.hl = __fmul_rn(l, 0.5f);
  g.cosa = cosf(ang); g.sina = sinf(ang);
  return g;
}
// pt_in_box3d (roipool3d_kernel.cu:14-28) with the reference build's FMA shape
__device__ __forceinline__ bool in_box(const BoxGeom &g, float x, float y, float z) {
  const float dx = __fsub_rn(x, g.cx), dz = __fsub_rn(z, g.cz);
  const bool pre = !(fabsf(dx) > 10.0f) && !(fabsf(__fsub_rn(y, g.cy)) > g.hh) && !(fabsf(dz) > 10.0f);
  const float x_rot = __fmaf_rn(dx, g.cosa, -__fmul_rn(dz, g.sina));
  const float z_rot = __fmaf_rn(dz, g.cosa, __fmul_rn(dx, g.sina));
  return pre && (x_rot >= -g.hl) && (x_rot <= g.hl) && (z_rot >= -g.hw) && (z_rot <= g.hw);
}

__global__ void __launch_bounds__(kSelWarps * 32) roipool_select_kernel(int n, int m, int S, const float *__restrict__ xyz,
                                                                        const float *__restrict__ boxes3d,
                                                                        const GridHdr *__restrict__ hdrs,
                                                                        const int *__restrict__ cell_start,
                                                                        const float4 *__restrict__ sorted,
                                                                        int *__restrict__ sel_out, int *__restrict__ empty_flag) {
  __shared__ int s_hits[kSelWarps][kSelCap];
  __shared__ int s_row_start[kSelWarps][32], s_row_pref[kSelWarps][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t cloud = blockIdx.y;
  const GridHdr h = hdrs[cloud];
  const int *cstart = cell_start + cloud * (size_t)(kMaxCells + 1);
  const float4 *spts = sorted + cloud * (size_t)n;
  const float *pts = xyz + cloud * (size_t)n * 3;
  const uint32_t lt_mask = (1u << lane) - 1u;
  int *hits = s_hits[warp];

  for (int j = blockIdx.x * kSelWarps + warp; j < m; j += gridDim.x * kSelWarps) {
    const BoxGeom g = box_geom(boxes3d + (cloud * (size_t)m + j) * 7);
    int *sel = sel_out + (cloud * (size_t)m + j) * (size_t)S;
    // Padded axis-aligned bound of the points that can pass the predicate: |dx| <= hl|cos| + hw|sin| (and <= 10),
    // |dz| <= hl|sin| + hw|cos| (and <= 10), |y - cy| <= hh, each widened against the rounding of the float predicate.
    // cell_axis is monotone and is what binned the points, so the cell range of the bound covers them exactly.
    const float ex = fminf(fabsf(g.cosa) * g.hl + fabsf(g.sina) * g.hw, 10.0f) * 1.0001f + 1e-3f;
    const float ez = fminf(fabsf(g.sina) * g.hl + fabsf(g.cosa) * g.hw, 10.0f) * 1.0001f + 1e-3f;
    const float ey = g.hh * 1.0001f + 1e-3f;
    int cnt = 0;
    bool overflow = false;
    // a NaN anywhere in the bound: no point can pass the predicate (every comparison with NaN is false)
    if (ex == ex && ez == ez && ey == ey && g.cx == g.cx && g.cy == g.cy && g.cz == g.cz && h.ncell > 0 && h.inv > 0.f) {
      const int x0 = cell_axis(g.cx - ex, h.ox, h.inv, h.dx), x1 = cell_axis(g.cx + ex, h.ox, h.inv, h.dx);
      const int y0 = cell_axis(g.cy - ey, h.oy, h.inv, h.dy), y1 = cell_axis(g.cy + ey, h.oy, h.inv, h.dy);
      const int z0 = cell_axis(g.cz - ez, h.oz, h.inv, h.dz), z1 = cell_axis(g.cz + ez, h.oz, h.inv, h.dz);
      const int ny = y1 - y0 + 1, nrows = ny * (z1 - z0 + 1);
      for (int r0 = 0; r0 < nrows && !overflow; r0 += 32) {
        // up to 32 (y, z) rows of cells at a time: each is one contiguous run of `sorted` (x is the fastest cell axis)
        const int r = r0 + lane;
        int start = 0, len = 0;
        if (r < nrows) {
          const int rowbase = ((z0 + r / ny) * h.dy + (y0 + r % ny)) * h.dx;
          start = __ldg(cstart + rowbase + x0);
          len = __ldg(cstart + rowbase + x1 + 1) - start;
        }
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
          if (lane >= o) incl += v;
        }
        __syncwarp();
        s_row_start[warp][lane] = start;
        s_row_pref[warp][lane + 1] = incl;
        if (lane == 0) s_row_pref[warp][0] = 0;
        __syncwarp();
        const int total = s_row_pref[warp][32];
        for (int c0 = 0; c0 < total; c0 += 32) {
          const int c = c0 + lane;
          bool in = false;
          int k = 0;
          if (c < total) {
            int lo = 0;                                  // largest row with pref[row] <= c (5-step binary search)
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1)
              if (s_row_pref[warp][lo + step] <= c) lo += step;
            const float4 p = __ldg(spts + s_row_start[warp][lo] + (c - s_row_pref[warp][lo]));
            in = in_box(g, p.x, p.y, p.z);
            k = __float_as_int(p.w);
          }
          const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, in);
          const int pos = cnt + __popc(ballot & lt_mask);
          if (in && pos < kSelCap) hits[pos] = k;
          cnt += __popc(ballot);
        }
        if (cnt > kSelCap) overflow = true;
      }
    }
    __syncwarp();
    if (overflow) {
      // dense box: exact early-exit scan over the original order (ends after ~S / density points)
      cnt = 0;
      for (int base = 0; base < n && cnt < S; base += 32) {
        const int k = base + lane;
        bool in = false;
        if (k < n) in = in_box(g, __ldg(pts + (size_t)k * 3), __ldg(pts + (size_t)k * 3 + 1), __ldg(pts + (size_t)k * 3 + 2));
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, in);
        const int pos = cnt + __popc(ballot & lt_mask);
        if (in && pos < S) sel[pos] = k;
        cnt += __popc(ballot);
      }
      __syncwarp();
      const int have = min(cnt, S);
      for (int k = have + lane; k < S; k += 32) sel[k] = sel[k % have];   // (have >= 1: the box overflowed the buffer)
    } else if (cnt > 0) {
      // rank of every hit among the hits = its slot in index order; the first S are the reference's selection
      const int have = min(cnt, S);
      for (int h0 = 0; h0 < cnt; h0 += 32) {
        const int hh = h0 + lane;
        const int mine = hh < cnt ? hits[hh] : 0x7FFFFFFF;
        int rank = 0;
        for (int t = 0; t < cnt; ++t) rank += (hits[t] < mine) ? 1 : 0;
        if (hh < cnt && rank < S) sel[rank] = mine;
      }
      __syncwarp();
      for (int k = have + lane; k < S; k += 32) sel[k] = sel[k % have];   // wrap-around duplicates (roipool3d_kernel.cu:150-158)
    }
    if (lane == 0 && cnt == 0) {
      empty_flag[cloud * (size_t)m + j] = 1;   // roipool3d_kernel.cu:146-148; non-empty boxes leave the caller's fill alone
      sel[0] = -1;                             // tells the write kernel to skip the box (its rows keep the caller's zeros)
    }
    __syncwarp();
  }
}

// One CTA per (box, cloud): the box's S*(3+C) output floats as float4 stores.
constexpr int kWrThreads = 256;
template <bool kVec>
__global__ void __launch_bounds__(kWrThreads) roipool_write_kernel(int n, int m, int c, int S, const float *__restrict__ xyz,
                                                                    const float *__restrict__ feat,
                                                                    const int *__restrict__ sel_in,
                                                                    float *__restrict__ pooled) {
  const size_t cloud = blockIdx.y;
  const int j = blockIdx.x;
  const int row_len = 3 + c;
  const int *sel = sel_in + (cloud * (size_t)m + j) * (size_t)S;
  if (__ldg(sel) < 0) return;                           // empty box: its rows keep the caller's zero fill (roipool3d_kernel.cu:146-148)
  const float *pts = xyz + cloud * (size_t)n * 3;
  const float *fts = feat + cloud * (size_t)n * c;
  float *dst = pooled + (cloud * (size_t)m + j) * (size_t)S * row_len;
  const int total = S * row_len;
  if (kVec) {
    // element e = 4 * v: row k = e / row_len, column t = e % row_len, advanced incrementally (no division in the loop)
    const int stride = 4 * kWrThreads;
    const int dk = stride / row_len, dt = stride - dk * row_len;
    int e = 4 * (int)threadIdx.x;
    int k = e / row_len, t = e - k * row_len;
    for (; e < total; e += stride) {
      float v[4];
      int kk = k, tt = t;
      int src = __ldg(sel + kk);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v[q] = tt < 3 ? __ldg(pts + (size_t)src * 3 + tt) : __ldg(fts + (size_t)src * c + (tt - 3));
        if (++tt == row_len) { tt = 0; ++kk; if (q < 3 && kk < S) src = __ldg(sel + kk); }
      }
      __stcs(reinterpret_cast<float4 *>(dst + e), make_float4(v[0], v[1], v[2], v[3]));
      k += dk; t += dt;
      if (t >= row_len) { t -= row_len; ++k; }
    }
  } else {
    for (int e = threadIdx.x; e < total; e += kWrThreads) {
      const int k = e / row_len, t = e - k * row_len;
      const int src = __ldg(sel + k);
      __stcs(dst + e, t < 3 ? __ldg(pts + (size_t)src * 3 + t) : __ldg(fts + (size_t)src * c + (t - 3)));
    }
  }
}

bool roipool_grid_applicable(int batch, int n, int m, int s) {
  static const int enabled = []() { const char *e = getenv("WS3D_ROIPOOL_GRID"); return (e && *e) ? atoi(e) : 1; }();
  return enabled && n >= 1024 && n <= 65536 && s >= 1 && s <= kSelCap && m >= 1 && m <= 0x7FFFFFFF / kWrThreads && batch <= 65535;
}

int roipool_grid(int batch, int n, int m, int c, int s, const float *xyz, const float *boxes3d, const float *feat, float *pooled,
                 int *flag, cudaStream_t stream) {
  const size_t hdr_bytes = ((size_t)batch * sizeof(GridHdr) + 255) & ~(size_t)255;
  const size_t start_bytes = ((size_t)batch * (kMaxCells + 1) * sizeof(int) + 255) & ~(size_t)255;
  const size_t sorted_bytes = ((size_t)batch * n * sizeof(float4) + 255) & ~(size_t)255;
  const size_t sel_bytes = (size_t)batch * m * s * sizeof(int);
  char *ws = (char *)scratch(hdr_bytes + start_bytes + sorted_bytes + sel_bytes, 1);
  if (!ws) return (int)cudaErrorMemoryAllocation;
  GridHdr *hdrs = (GridHdr *)ws;
  int *cell_start = (int *)(ws + hdr_bytes);
  float4 *sorted = (float4 *)(ws + hdr_bytes + start_bytes);
  int *sel = (int *)(ws + hdr_bytes + start_bytes + sorted_bytes);
  cudaError_t e = cudaFuncSetAttribute(grid_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxCells * (int)sizeof(int));
  if (e != cudaSuccess) { set_error("roipool3d grid: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  grid_build_kernel<<<batch, 1024, kMaxCells * sizeof(int), stream>>>(n, 0.f, kMaxCells, xyz, hdrs, cell_start, sorted);
  int rc = check_launch("roipool3d (grid build)");
  if (rc) return rc;
  int gx = ceil_div(m, kSelWarps);
  const int cap = ceil_div(8 * num_sms(), batch);
  if (gx > cap) gx = cap;
  roipool_select_kernel<<<dim3((unsigned)gx, (unsigned)batch), kSelWarps * 32, 0, stream>>>(n, m, s, xyz, boxes3d, hdrs, cell_start,
                                                                                          sorted, sel, flag);
  rc = check_launch("roipool3d (select)");
  if (rc) return rc;
  const bool vec = ((size_t)s * (3 + c)) % 4 == 0 && (reinterpret_cast<uintptr_t>(pooled) & 15u) == 0;
  const dim3 grid((unsigned)m, (unsigned)batch);
  if (vec) roipool_write_kernel<true><<<grid, kWrThreads, 0, stream>>>(n, m, c, s, xyz, feat, sel, pooled);
  else roipool_write_kernel<false><<<grid, kWrThreads, 0, stream>>>(n, m, c, s, xyz, feat, sel, pooled);
  return check_launch("roipool3d (write)");
}

int roipool_dispatch(int batch, int n, int m, int c, int s, const float *xyz, const float *boxes3d, const float *feat,
                     float *pooled, int *flag, cudaStream_t stream) {
  const char *what = "roipool3d";
  if (batch < 0 || n < 0 || m < 0 || c < 0 || s < 0) return fail_arg(what);
  if (batch == 0 || m == 0) return 0;
  if (!boxes3d || !flag || (n > 0 && !xyz) || (s > 0 && !pooled) || (n > 0 && c > 0 && !feat)) return fail_arg(what);
  if (batch > 65535) return fail_arg(what);
  if (roipool_grid_applicable(batch, n, m, s)) return roipool_grid(batch, n, m, c, s, xyz, boxes3d, feat, pooled, flag, stream);
  RoiParams prm;
  prm.n = n; prm.m = m; prm.c = c; prm.s = s;
  prm.xyz = xyz; prm.boxes3d = boxes3d; prm.feat = feat; prm.pooled = pooled; prm.empty_flag = flag;
  prm.sel_g = nullptr;
  if (s > kMaxSel) {
    prm.sel_g = (int *)scratch((size_t)batch * m * s * sizeof(int), 1);
    if (!prm.sel_g) return (int)cudaErrorMemoryAllocation;
  }
  const int chunk_cap = n < kChunkPts ? n : kChunkPts;
  const int threads = kMaxWarps * 32;
  const size_t smem = (size_t)((chunk_cap * 3 + 3) & ~3) * sizeof(float) +
                      (prm.sel_g ? 0 : (size_t)kMaxWarps * (s < kMaxSel ? s : kMaxSel) * sizeof(int)) + 16;
  const int resident = smem > 110 * 1024 ? 1 : (smem > 70 * 1024 ? 2 : 4);  // CTAs per SM by shared memory
  int ctas_per_cloud = (num_sms() * resident) / batch;
  if (ctas_per_cloud < 1) ctas_per_cloud = 1;
  int boxes_per_cta = ceil_div(m, ctas_per_cloud);
  boxes_per_cta = ceil_div(boxes_per_cta, kMaxWarps) * kMaxWarps;  // whole groups of one box per warp
  prm.boxes_per_cta = boxes_per_cta;
  cudaError_t e = cudaFuncSetAttribute(roipool3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("roipool3d: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; }
  dim3 grid((unsigned)ceil_div(m, boxes_per_cta), (unsigned)batch);
  roipool3d_kernel<<<grid, threads, smem, stream>>>(prm);
  return check_launch(what);
}

}  // namespace
}  // namespace ws3d

WS3D_API int ws3d_roipool3d(int batch_size, int pts_num, int boxes_num, int feature_in_len, int sampled_pts_num,
                            const float *xyz, const float *boxes3d, const float *pts_feature, float *pooled_features,
                            int *pooled_empty_flag, ws3d_stream_t stream) {
  return ws3d::roipool_dispatch(batch_size, pts_num, boxes_num, feature_in_len, sampled_pts_num, xyz, boxes3d,
                                pts_feature, pooled_features, pooled_empty_flag, ws3d::to_stream(stream));
}
