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

int roipool_dispatch(int batch, int n, int m, int c, int s, const float *xyz, const float *boxes3d, const float *feat,
                     float *pooled, int *flag, cudaStream_t stream) {
  const char *what = "roipool3d";
  if (batch < 0 || n < 0 || m < 0 || c < 0 || s < 0) return fail_arg(what);
  if (batch == 0 || m == 0) return 0;
  if (!boxes3d || !flag || (n > 0 && !xyz) || (s > 0 && !pooled) || (n > 0 && c > 0 && !feat)) return fail_arg(what);
  if (batch > 65535) return fail_arg(what);
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
