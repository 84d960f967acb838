// ball_query, group/gather (+grad), three_nn, three_interpolate (+grad) for B200.
//
// Replaces pointnet2_lib/pointnet2/src/{ball_query_gpu,group_points_gpu,interpolate_gpu}.cu and
// the gather kernels of sampling_gpu.cu.  The reference kernels are thread-per-query scans that
// stream the whole cloud through L1/L2 per thread and thread-per-output-element gathers that
// re-read idx once per channel.  Here:
//   * ball_query: the cloud (<= 18432 points, 216 KB) is staged ONCE per CTA into shared memory by
//     TMA bulk copies (cp.async.bulk + mbarrier); one WARP per query scans it 32 points per step
//     (packed xyz, stride-3 words: bank-conflict free), hit lanes are compacted in index order by
//     ballot + popc, and the warp stops at the K-th hit.  Queries are handed to warps dynamically.
//     Up to two radii are scanned in one pass (multi-scale grouping shares the distance).
//   * group/gather/interpolate: idx (and weights) are loaded once per thread and reused across
//     all channels; stores are 16-byte wide and coalesced.
//   * three_nn: known points staged in shared memory as float4, thread per unknown point.
// Index-producing ops are bit-exact with the reference (same rounding order, same tie rules).
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"

namespace ws3d {
// interp_grad.cu: 0 = done, 1 = shape outside the gather path's limits, else a cudaError code
int three_interpolate_grad_gather(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight,
                                  float *grad_points, cudaStream_t stream);

// ball_query_grid.cu: cell-list pruned scan for clouds that are large against the ball
bool ball_query_grid_applicable(int nr, int b, int n, int m, const float *radius);
int ball_query_grid(int nr, int b, int n, int m, const float *radius, const int *nsample, const float *new_xyz,
                    const float *xyz, int *const *idx, cudaStream_t stream);

// three_nn_grid.cu: ring search over a cell grid of the known points
bool three_nn_grid_applicable(int b, int n, int m);
int three_nn_grid(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, float *weight,
                  cudaStream_t stream);

namespace {

// ---------------------------------------------------------------------------------------------
// ball_query
constexpr int kBqMaxChunkPts = 18432;  // 216 KB of packed xyz

template <int NR>
struct BqParams {
  int n, m;
  float r2[NR];
  int nsample[NR];
  const float *new_xyz;  // (B,M,3)
  const float *xyz;      // (B,N,3)
  int *idx[NR];          // (B,M,nsample[r]) zero-filled by the caller
  int q_per_cta;
};

template <int NR>
__global__ void __launch_bounds__(1024, 1) ball_query_kernel(BqParams<NR> prm) {
  extern __shared__ __align__(16) float s_xyz[];  // n*3 packed
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ int s_next;
  const int lane = threadIdx.x & 31;
  const int n = prm.n, m = prm.m;
  const size_t cloud = blockIdx.y;
  const int q0 = blockIdx.x * prm.q_per_cta;
  const int q1 = min(m, q0 + prm.q_per_cta);
  if (q0 >= q1) return;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&s_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_next = 0;
  }
  __syncthreads();
  stage_floats(s_xyz, prm.xyz + cloud * (size_t)n * 3, n * 3, &s_bar, 0);

  const uint32_t lt_mask = (1u << lane) - 1u;
  for (;;) {
    int q = 0;
    if (lane == 0) q = q0 + atomicAdd(&s_next, 1);
    q = __shfl_sync(0xFFFFFFFFu, q, 0);
    if (q >= q1) break;
    const float *qp = prm.new_xyz + (cloud * (size_t)m + q) * 3;
    const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
    int cnt[NR], first[NR];
    int *row[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      cnt[r] = 0; first[r] = 0;
      row[r] = prm.idx[r] + (cloud * (size_t)m + q) * prm.nsample[r];
    }
    for (int base = 0; base < n; base += 128) {
      float d2[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = base + u * 32 + lane;
        if (k < n) {
          const float x = s_xyz[k * 3], y = s_xyz[k * 3 + 1], z = s_xyz[k * 3 + 2];
          d2[u] = sqdist_ref(qx - x, qy - y, qz - z);
        } else {
          d2[u] = __int_as_float(0x7f800000);  // +inf: never inside
        }
      }
      bool all_full = true;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int K = prm.nsample[r];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (cnt[r] < K) {
            const bool hit = d2[u] < prm.r2[r];
            const uint32_t hits = __ballot_sync(0xFFFFFFFFu, hit);
            if (hits) {
              const int k = base + u * 32 + lane;
              if (cnt[r] == 0) first[r] = base + u * 32 + __ffs(hits) - 1;
              const int pos = cnt[r] + __popc(hits & lt_mask);
              if (hit && pos < K) row[r][pos] = k;
              cnt[r] += __popc(hits);
            }
          }
        }
        all_full = all_full && (cnt[r] >= K);
      }
      if (all_full) break;
    }
    // the first hit fills every slot that found no neighbour of its own (ball_query_gpu.cu:35-39)
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int K = prm.nsample[r];
      // a centre without any neighbour gets a zero row: what the caller's zero fill (pointnet2_utils.py:218) leaves
      // there in the reference -- written here so that callers of this library need no fill pass
      const int fill = cnt[r] > 0 ? first[r] : 0;
      for (int s = cnt[r] + lane; s < K; s += 32) row[r][s] = fill;
    }
  }
}

// Any-size fallback (n > kBqMaxChunkPts): thread per query over global memory.
__global__ void ball_query_generic_kernel(int n, int m, float r2, int K, const float *__restrict__ new_xyz,
                                          const float *__restrict__ xyz, int *__restrict__ idx) {
  const size_t cloud = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m) return;
  const float *qp = new_xyz + (cloud * (size_t)m + q) * 3;
  const float *p = xyz + cloud * (size_t)n * 3;
  int *row = idx + (cloud * (size_t)m + q) * K;
  const float qx = qp[0], qy = qp[1], qz = qp[2];
  int cnt = 0, first = 0;
  for (int k = 0; k < n && cnt < K; ++k) {
    const float d2 = sqdist_ref(qx - p[(size_t)k * 3], qy - p[(size_t)k * 3 + 1], qz - p[(size_t)k * 3 + 2]);
    if (d2 < r2) {
      if (cnt == 0) first = k;
      row[cnt++] = k;
    }
  }
  for (int s = cnt; s < K; ++s) row[s] = cnt > 0 ? first : 0;   // no neighbour: zero row (see ball_query_kernel)
}

template <int NR>
int launch_ball_query(int b, int n, int m, const float *radius, const int *nsample, const float *new_xyz,
                      const float *xyz, int *const *idx, cudaStream_t stream) {
  BqParams<NR> prm;
  prm.n = n; prm.m = m; prm.new_xyz = new_xyz; prm.xyz = xyz;
  for (int r = 0; r < NR; ++r) {
    prm.r2[r] = radius[r] * radius[r];  // f32 product, as ball_query_gpu.cu:23
    prm.nsample[r] = nsample[r];
    prm.idx[r] = idx[r];
  }
  const int threads = n > 8192 ? 1024 : (n > 1024 ? 512 : 256);
  const int warps = threads / 32;
  const size_t smem = (size_t)n * 3 * sizeof(float);
  int ctas_per_cloud, q_per_cta;
  if (smem > 100 * 1024) {
    // one resident CTA per SM: exactly one wave, queries handed out dynamically inside the CTA
    ctas_per_cloud = num_sms() / b > 0 ? num_sms() / b : 1;
    q_per_cta = ceil_div(m, ctas_per_cloud);
  } else {
    ctas_per_cloud = ceil_div(2 * num_sms(), b);
    q_per_cta = ceil_div(m, ctas_per_cloud);
    if (q_per_cta < warps) q_per_cta = warps;  // do not reload the cloud for less than a query per warp
  }
  prm.q_per_cta = q_per_cta;
  auto kern = ball_query_kernel<NR>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("ball_query: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; }
  dim3 grid((unsigned)ceil_div(m, q_per_cta), (unsigned)b);
  kern<<<grid, threads, smem, stream>>>(prm);
  return check_launch("ball_query");
}

int ball_query_dispatch(int nr, int b, int n, int m, const float *radius, const int *nsample, const float *new_xyz,
                        const float *xyz, int *const *idx, cudaStream_t stream) {
  if (b < 0 || n < 0 || m < 0) return fail_arg("ball_query");
  for (int r = 0; r < nr; ++r)
    if (nsample[r] < 0) return fail_arg("ball_query (nsample)");
  if (b == 0 || m == 0) return 0;
  if (n == 0) {  // no points: every row is a no-neighbour row
    for (int r = 0; r < nr; ++r) {
      if (!idx[r] || nsample[r] == 0) continue;
      cudaError_t e = cudaMemsetAsync(idx[r], 0, (size_t)b * m * nsample[r] * sizeof(int), stream);
      if (e != cudaSuccess) { set_error("ball_query: memset: %s", cudaGetErrorString(e)); return (int)e; }
    }
    return 0;
  }
  if (!new_xyz || !xyz) return fail_arg("ball_query (null pointer)");
  if (b > 65535) return fail_arg("ball_query (batch > 65535)");
  {
    static const int use_grid = []() { const char *e = getenv("WS3D_BQ_GRID"); return (e && *e) ? atoi(e) : 1; }();
    if (use_grid && ball_query_grid_applicable(nr, b, n, m, radius))
      return ball_query_grid(nr, b, n, m, radius, nsample, new_xyz, xyz, idx, stream);
  }
  if (n > kBqMaxChunkPts) {
    for (int r = 0; r < nr; ++r) {
      if (nsample[r] == 0) continue;
      dim3 grid((unsigned)ceil_div(m, 128), (unsigned)b);
      ball_query_generic_kernel<<<grid, 128, 0, stream>>>(n, m, radius[r] * radius[r], nsample[r], new_xyz, xyz, idx[r]);
      int rc = check_launch("ball_query (generic)");
      if (rc) return rc;
    }
    return 0;
  }
  if (nr == 1) return launch_ball_query<1>(b, n, m, radius, nsample, new_xyz, xyz, idx, stream);
  if (nr == 2) return launch_ball_query<2>(b, n, m, radius, nsample, new_xyz, xyz, idx, stream);
  return fail_arg("ball_query (1 or 2 radii)");
}

// ---------------------------------------------------------------------------------------------
// group_points / gather_points:  out[b,c,e] = points[b,c,idx[b,e]],  e over npoints*nsample
constexpr int kGatherThreads = 256;

template <bool VEC>
__global__ void __launch_bounds__(kGatherThreads) group_points_kernel(int c, int n, long long per_cloud, int c_per_cta,
                                                                        const float *__restrict__ points,
                                                                        const int *__restrict__ idx,
                                                                        float *__restrict__ out) {
  const size_t cloud = blockIdx.y;
  const long long e0 = ((long long)blockIdx.x * kGatherThreads + threadIdx.x) * (VEC ? 4 : 1);
  if (e0 >= per_cloud) return;
  const int cb = blockIdx.z * c_per_cta, ce = min(c, cb + c_per_cta);
  const int *ip = idx + cloud * per_cloud + e0;
  const float *src = points + cloud * (size_t)c * n;
  float *dst = out + cloud * (size_t)c * per_cloud + e0;
  if (VEC) {
    const int4 id = __ldg(reinterpret_cast<const int4 *>(ip));
#pragma unroll 4
    for (int ci = cb; ci < ce; ++ci) {
      const float *row = src + (size_t)ci * n;
      float4 v;
      v.x = __ldg(row + id.x); v.y = __ldg(row + id.y); v.z = __ldg(row + id.z); v.w = __ldg(row + id.w);
      __stcs(reinterpret_cast<float4 *>(dst + (size_t)ci * per_cloud), v);
    }
  } else {
    const int id = __ldg(ip);
#pragma unroll 4
    for (int ci = cb; ci < ce; ++ci) dst[(size_t)ci * per_cloud] = __ldg(src + (size_t)ci * n + id);
  }
}

template <bool VEC>
__global__ void __launch_bounds__(kGatherThreads) group_points_grad_kernel(int c, int n, long long per_cloud, int c_per_cta,
                                                                             long long grad_cloud_stride,
                                                                             const float *__restrict__ grad_out,
                                                                             const int *__restrict__ idx,
                                                                             float *__restrict__ grad_points) {
  const size_t cloud = blockIdx.y;
  const long long e0 = ((long long)blockIdx.x * kGatherThreads + threadIdx.x) * (VEC ? 4 : 1);
  if (e0 >= per_cloud) return;
  const int cb = blockIdx.z * c_per_cta, ce = min(c, cb + c_per_cta);
  const int *ip = idx + cloud * per_cloud + e0;
  const float *g = grad_out + cloud * (size_t)grad_cloud_stride + e0;   // a cloud's block may hold leading channels that carry no gradient
  float *dst = grad_points + cloud * (size_t)c * n;
  if (VEC) {
    const int4 id = __ldg(reinterpret_cast<const int4 *>(ip));
#pragma unroll 2
    for (int ci = cb; ci < ce; ++ci) {
      const float4 v = __ldcs(reinterpret_cast<const float4 *>(g + (size_t)ci * per_cloud));
      float *row = dst + (size_t)ci * n;
      atomicAdd(row + id.x, v.x); atomicAdd(row + id.y, v.y); atomicAdd(row + id.z, v.z); atomicAdd(row + id.w, v.w);
    }
  } else {
    const int id = __ldg(ip);
    for (int ci = cb; ci < ce; ++ci) atomicAdd(dst + (size_t)ci * n + id, g[(size_t)ci * per_cloud]);
  }
}

// Channels handled by one CTA: all of them when the (element tile x cloud) grid already fills the
// GPU, otherwise split so that at least ~4 CTAs per SM exist (idx is re-read once per split only).
int channels_per_cta(long long ctas_without_split, int c) {
  const long long want = 4LL * num_sms();
  if (ctas_without_split >= want || c <= 8) return c > 0 ? c : 1;
  long long splits = (want + ctas_without_split - 1) / ctas_without_split;
  int per = (int)((c + splits - 1) / splits);
  if (per < 8) per = 8;
  return per;
}

int group_dispatch(bool grad, int b, int c, int n, long long per_cloud, const float *a, const int *idx, float *o,
                   cudaStream_t stream, const char *what, long long grad_cloud_stride = -1) {
  if (grad_cloud_stride < 0) grad_cloud_stride = (long long)c * per_cloud;
  if (b < 0 || c < 0 || n < 0 || per_cloud < 0) return fail_arg(what);
  if (b == 0 || c == 0 || per_cloud == 0) return 0;
  if (!a || !idx || !o) return fail_arg(what);
  if (b > 65535) return fail_arg(what);
  const bool vec = (per_cloud % 4 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15u) == 0) &&
                   ((reinterpret_cast<uintptr_t>(grad ? (const void *)a : (const void *)o) & 15u) == 0);
  const long long per_thread = vec ? 4 : 1;
  const long long gx = (per_cloud + kGatherThreads * per_thread - 1) / (kGatherThreads * per_thread);
  const int c_per_cta = channels_per_cta(gx * b, c);
  dim3 grid((unsigned)gx, (unsigned)b, (unsigned)ceil_div(c, c_per_cta));
  if (!grad) {
    if (vec) group_points_kernel<true><<<grid, kGatherThreads, 0, stream>>>(c, n, per_cloud, c_per_cta, a, idx, o);
    else group_points_kernel<false><<<grid, kGatherThreads, 0, stream>>>(c, n, per_cloud, c_per_cta, a, idx, o);
  } else {
    if (vec) group_points_grad_kernel<true><<<grid, kGatherThreads, 0, stream>>>(c, n, per_cloud, c_per_cta, grad_cloud_stride, a, idx, o);
    else group_points_grad_kernel<false><<<grid, kGatherThreads, 0, stream>>>(c, n, per_cloud, c_per_cta, grad_cloud_stride, a, idx, o);
  }
  return check_launch(what);
}

// ---------------------------------------------------------------------------------------------
// three_nn: thread per unknown point, known points staged as float4 in shared memory
constexpr int kNnThreads = 256;
constexpr int kNnChunk = 4096;  // known points per stage (64 KB)

__global__ void __launch_bounds__(kNnThreads) three_nn_kernel(int n, int m, const float *__restrict__ unknown,
                                                               const float *__restrict__ known,
                                                               float *__restrict__ dist2, int *__restrict__ idx,
                                                               float *__restrict__ weight) {
  extern __shared__ __align__(16) float4 s_known[];
  const size_t cloud = blockIdx.y;
  const int i = blockIdx.x * kNnThreads + threadIdx.x;
  const float *kn = known + cloud * (size_t)m * 3;
  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (i < n) {
    const float *u = unknown + (cloud * (size_t)n + i) * 3;
    ux = __ldg(u); uy = __ldg(u + 1); uz = __ldg(u + 2);
  }
  // The reference keeps doubles initialised to 1e40 and stores (float)1e40 == +inf for slots it
  // never fills (interpolate_gpu.cu:30,50); a float +inf start is equivalent: "d < best" is false
  // for d == +inf in both.
  const float kInf = __int_as_float(0x7f800000);
  float b1 = kInf, b2 = kInf, b3 = kInf;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int c0 = 0; c0 < m; c0 += kNnChunk) {
    const int cn = min(kNnChunk, m - c0);
    __syncthreads();
    for (int k = threadIdx.x; k < cn; k += kNnThreads) {
      const float *p = kn + (size_t)(c0 + k) * 3;
      s_known[k] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < cn; ++k) {
      const float4 p = s_known[k];
      const float d = sqdist_ref(ux - p.x, uy - p.y, uz - p.z);
      if (d < b3) {
        const int kk = c0 + k;
        if (d < b1) {
          b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = kk;
        } else if (d < b2) {
          b3 = b2; i3 = i2; b2 = d; i2 = kk;
        } else {
          b3 = d; i3 = kk;
        }
      }
    }
  }
  if (i < n) store_three_nn(dist2, idx, weight, cloud * (size_t)n + i, b1, b2, b3, i1, i2, i3);
}

// three_interpolate: out[b,c,i] = fma(w2,p2, fma(w0,p0, rn(w1*p1)))  (reference SASS order)
constexpr int kInterpThreads = 256;
constexpr int kInterpChannels = 32;  // channels per CTA (idx/weight loaded once for all of them)

__global__ void __launch_bounds__(kInterpThreads) three_interpolate_kernel(int c, int m, int n,
                                                                            const float *__restrict__ points,
                                                                            const int *__restrict__ idx,
                                                                            const float *__restrict__ weight,
                                                                            float *__restrict__ out) {
  const size_t cloud = blockIdx.z;
  const int i = blockIdx.x * kInterpThreads + threadIdx.x;
  if (i >= n) return;
  const int c0 = blockIdx.y * kInterpChannels, c1 = min(c, c0 + kInterpChannels);
  const int *ip = idx + (cloud * (size_t)n + i) * 3;
  const float *wp = weight + (cloud * (size_t)n + i) * 3;
  const int i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2);
  const float w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
#pragma unroll 4
  for (int ci = c0; ci < c1; ++ci) {
    const float *row = points + (cloud * (size_t)c + ci) * m;
    float t = __fmul_rn(w1, __ldg(row + i1));
    t = __fmaf_rn(w0, __ldg(row + i0), t);
    t = __fmaf_rn(w2, __ldg(row + i2), t);
    __stcs(out + (cloud * (size_t)c + ci) * n + i, t);
  }
}

// Shared-memory variant for the common case (m*4 bytes per channel row fits): the CB feature rows
// of a channel chunk are staged once per CTA (one contiguous TMA bulk copy: rows of consecutive
// channels are adjacent in (B,C,M)), so the three random reads per output become shared-memory reads
// (a few-way bank conflict) instead of 32 L1 sector look-ups per warp instruction.
constexpr int kInterpSmemThreads = 1024;  // launch bound; 512 threads are used when a CTA has few points
__global__ void __launch_bounds__(kInterpSmemThreads) three_interpolate_smem_kernel(
    int c, int m, int n, int cb, int n_per_cta, const float *__restrict__ points, const int *__restrict__ idx,
    const float *__restrict__ weight, float *__restrict__ out) {
  extern __shared__ __align__(16) float s_rows[];  // cb * m
  __shared__ __align__(8) unsigned long long s_bar;
  const size_t cloud = blockIdx.z;
  const int c0 = blockIdx.x * cb, cn = min(cb, c - c0);
  const int i_begin = blockIdx.y * n_per_cta, i_end = min(n, i_begin + n_per_cta);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&s_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  stage_floats(s_rows, points + (cloud * (size_t)c + c0) * m, cn * m, &s_bar, 0);
  for (int i = i_begin + (int)threadIdx.x; i < i_end; i += (int)blockDim.x) {
    const int *ip = idx + (cloud * (size_t)n + i) * 3;
    const float *wp = weight + (cloud * (size_t)n + i) * 3;
    const int i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2);
    const float w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
    float *o = out + (cloud * (size_t)c + c0) * n + i;
#pragma unroll 4
    for (int cc = 0; cc < cn; ++cc) {
      const float *row = s_rows + cc * m;
      float t = __fmul_rn(w1, row[i1]);
      t = __fmaf_rn(w0, row[i0], t);
      t = __fmaf_rn(w2, row[i2], t);
      __stcs(o + (size_t)cc * n, t);
    }
  }
}

// Same, with four channels interleaved per shared-memory slot ([k][4] floats): one 16-byte read per neighbour
// serves four channels.  The gathers are random in k, so a 4-byte read per lane costs ~3.5 bank-conflict
// wavefronts per warp instruction; a 16-byte read is issued per quarter warp and costs ~1.6 per 8 lanes,
// i.e. less than half the shared-memory time per output (the kernel's bottleneck in ncu: 22 % DRAM).
// Optional epilogue (EPI): out = act(interp + scale1[c] * row1[b, i] + shift[c]), flags bit 0 = ReLU, bit 1 = round the
// result to TF32.  It lets a feature-propagation layer apply the interpolated part of its first 1x1 convolution to the
// KNOWN points (4x fewer columns) and interpolate the product: conv(interp(f)) = interp(conv(f)) because the stencil
// weights do not depend on the channel (ws3d_three_interpolate_affine).
struct InterpEpilogue {
  const float *scale1;   // (c) or null
  const float *row1;     // (B, n) or null
  const float *shift;    // (c) or null
  int flags;
};

// MODE: bit 0 = epilogue present, bit 1 = ReLU, bit 2 = round to TF32 (compile-time: ncu showed 46 instructions per output
// element with run-time flag tests and per-element loads of the per-channel terms)
template <int MODE>
__global__ void __launch_bounds__((MODE & 1) ? 512 : kInterpSmemThreads) three_interpolate_smem4_kernel(
    int c, int m, int n, int cb, int n_per_cta, const float *__restrict__ points, const int *__restrict__ idx,
    const float *__restrict__ weight, float *__restrict__ out, InterpEpilogue epi) {
  extern __shared__ __align__(16) float s_rows[];  // (cb/4) groups x m x 4
  __shared__ float4 s_scale[16], s_shift[16];      // per-channel epilogue terms of this CTA's channels (cb <= 64)
  constexpr bool EPI = (MODE & 1) != 0;
  const size_t cloud = blockIdx.z;
  const int c0 = blockIdx.x * cb, cn = min(cb, c - c0);  // cn % 4 == 0 (host guarantees c % 4 == 0 and cb % 4 == 0)
  const int i_begin = blockIdx.y * n_per_cta, i_end = min(n, i_begin + n_per_cta);
  const float *src = points + (cloud * (size_t)c + c0) * m;
  if (EPI && threadIdx.x < cn) {
    reinterpret_cast<float *>(s_scale)[threadIdx.x] = epi.scale1 ? __ldg(epi.scale1 + c0 + threadIdx.x) : 0.f;
    reinterpret_cast<float *>(s_shift)[threadIdx.x] = epi.shift ? __ldg(epi.shift + c0 + threadIdx.x) : 0.f;
  }
  // The stencils (3 indices + 3 weights per point) do not depend on shared memory: the first two points' loads are
  // issued BEFORE the rows are staged, later ones two points ahead of their use (ncu: the kernel waited on these
  // global loads more than on anything else; one point of look-ahead left most of the latency exposed).
  constexpr int PB = 2;
  const int stride = (int)blockDim.x;
  int si[PB][3];
  float sw[PB][3], sr[PB];
  auto load_stencil = [&](int q, int i) {
    if (i < i_end) {
      const int *ip = idx + (cloud * (size_t)n + i) * 3;
      const float *wp = weight + (cloud * (size_t)n + i) * 3;
      si[q][0] = __ldg(ip); si[q][1] = __ldg(ip + 1); si[q][2] = __ldg(ip + 2);
      sw[q][0] = __ldg(wp); sw[q][1] = __ldg(wp + 1); sw[q][2] = __ldg(wp + 2);
      sr[q] = (EPI && epi.row1) ? __ldg(epi.row1 + cloud * (size_t)n + i) : 0.f;
    } else {
      si[q][0] = si[q][1] = si[q][2] = 0;
      sw[q][0] = sw[q][1] = sw[q][2] = 0.f;
      sr[q] = 0.f;
    }
  };
  int i = i_begin + (int)threadIdx.x;
#pragma unroll
  for (int q = 0; q < PB; ++q) load_stencil(q, i + q * stride);
  // stage the rows: the four channels of a group at one position k are four coalesced loads and ONE conflict-free
  // 16-byte shared-memory store ([group][k][4] layout)
  {
    float4 *d4 = reinterpret_cast<float4 *>(s_rows);
    for (int g = 0; g < (cn >> 2); ++g) {
      const float *r0 = src + (size_t)(g * 4) * m;
#pragma unroll 2
      for (int k = threadIdx.x; k < m; k += stride)
        d4[(size_t)g * m + k] = make_float4(__ldg(r0 + k), __ldg(r0 + m + k), __ldg(r0 + 2 * (size_t)m + k), __ldg(r0 + 3 * (size_t)m + k));
    }
  }
  __syncthreads();
  const float4 *s4 = reinterpret_cast<const float4 *>(s_rows);
  for (; i < i_end; i += PB * stride) {
    int ci[PB][3];
    float cw[PB][3], cr[PB];
#pragma unroll
    for (int q = 0; q < PB; ++q) {
#pragma unroll
      for (int a = 0; a < 3; ++a) { ci[q][a] = si[q][a]; cw[q][a] = sw[q][a]; }
      cr[q] = sr[q];
    }
#pragma unroll
    for (int q = 0; q < PB; ++q) load_stencil(q, i + (PB + q) * stride);   // two points ahead
#pragma unroll
    for (int q = 0; q < PB; ++q) {
      const int iq = i + q * stride;
      if (iq >= i_end) break;
      const int i0 = ci[q][0], i1 = ci[q][1], i2 = ci[q][2];
      const float w0 = cw[q][0], w1 = cw[q][1], w2 = cw[q][2], r1 = cr[q];
      float *o = out + (cloud * (size_t)c + c0) * n + iq;
      const float4 *row = s4;
      const int ng = cn >> 2;
#pragma unroll 2
      for (int g = 0; g < ng; ++g, row += m, o += 4 * (size_t)n) {
        const float4 p0 = row[i0], p1 = row[i1], p2 = row[i2];
        float v[4] = {__fmaf_rn(w2, p2.x, __fmaf_rn(w0, p0.x, __fmul_rn(w1, p1.x))),
                      __fmaf_rn(w2, p2.y, __fmaf_rn(w0, p0.y, __fmul_rn(w1, p1.y))),
                      __fmaf_rn(w2, p2.z, __fmaf_rn(w0, p0.z, __fmul_rn(w1, p1.z))),
                      __fmaf_rn(w2, p2.w, __fmaf_rn(w0, p0.w, __fmul_rn(w1, p1.w)))};
        if (EPI) {
          const float4 sc = s_scale[g], sh = s_shift[g];
          v[0] = fmaf(sc.x, r1, v[0]) + sh.x; v[1] = fmaf(sc.y, r1, v[1]) + sh.y;
          v[2] = fmaf(sc.z, r1, v[2]) + sh.z; v[3] = fmaf(sc.w, r1, v[3]) + sh.w;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (MODE & 2) v[u] = fmaxf(v[u], 0.f);
            if (MODE & 4) v[u] = __uint_as_float((__float_as_uint(v[u]) + 0x1000u) & 0xFFFFE000u);
          }
        }
        __stcs(o, v[0]);
        __stcs(o + n, v[1]);
        __stcs(o + 2 * (size_t)n, v[2]);
        __stcs(o + 3 * (size_t)n, v[3]);
      }
    }
  }
}

__global__ void __launch_bounds__(kInterpThreads) three_interpolate_grad_kernel(int c, int n, int m,
                                                                                 const float *__restrict__ grad_out,
                                                                                 const int *__restrict__ idx,
                                                                                 const float *__restrict__ weight,
                                                                                 float *__restrict__ grad_points) {
  const size_t cloud = blockIdx.z;
  const int i = blockIdx.x * kInterpThreads + threadIdx.x;
  if (i >= n) return;
  const int c0 = blockIdx.y * kInterpChannels, c1 = min(c, c0 + kInterpChannels);
  const int *ip = idx + (cloud * (size_t)n + i) * 3;
  const float *wp = weight + (cloud * (size_t)n + i) * 3;
  const int i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2);
  const float w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
  for (int ci = c0; ci < c1; ++ci) {
    const float g = __ldcs(grad_out + (cloud * (size_t)c + ci) * n + i);
    float *row = grad_points + (cloud * (size_t)c + ci) * m;
    atomicAdd(row + i0, __fmul_rn(g, w0));
    atomicAdd(row + i1, __fmul_rn(g, w1));
    atomicAdd(row + i2, __fmul_rn(g, w2));
  }
}

}  // namespace

// used by query_and_group.cu
int ball_query_multi(int nr, int b, int n, int m, const float *radius, const int *nsample, const float *new_xyz,
                     const float *xyz, int *const *idx, cudaStream_t stream) {
  return ball_query_dispatch(nr, b, n, m, radius, nsample, new_xyz, xyz, idx, stream);
}

}  // namespace ws3d

using namespace ws3d;

WS3D_API int ws3d_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                             int *idx, ws3d_stream_t stream) {
  if (nsample > 0 && !idx) return fail_arg("ball_query (idx)");
  int *rows[1] = {idx};
  return ball_query_dispatch(1, b, n, m, &radius, &nsample, new_xyz, xyz, rows, to_stream(stream));
}

WS3D_API int ws3d_ball_query2(int b, int n, int m, float radius0, int nsample0, float radius1, int nsample1,
                              const float *new_xyz, const float *xyz, int *idx0, int *idx1, ws3d_stream_t stream) {
  const float rr[2] = {radius0, radius1};
  const int kk[2] = {nsample0, nsample1};
  int *rows[2] = {idx0, idx1};
  if ((nsample0 > 0 && !idx0) || (nsample1 > 0 && !idx1)) return fail_arg("ball_query2 (idx)");
  return ball_query_dispatch(2, b, n, m, rr, kk, new_xyz, xyz, rows, to_stream(stream));
}

WS3D_API int ws3d_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx,
                               float *out, ws3d_stream_t stream) {
  return group_dispatch(false, b, c, n, (long long)npoints * nsample, points, idx, out, to_stream(stream), "group_points");
}

WS3D_API int ws3d_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                                    const int *idx, float *grad_points, ws3d_stream_t stream) {
  return group_dispatch(true, b, c, n, (long long)npoints * nsample, grad_out, idx, grad_points, to_stream(stream),
                        "group_points_grad");
}

WS3D_API int ws3d_group_concat_grad(int b, int n, int m, int c, int nsample, int use_xyz, const float *grad_out, const int *idx,
                                    float *grad_features, ws3d_stream_t stream) {
  if (b < 0 || n < 0 || m < 0 || c < 0 || nsample < 0) return fail_arg("group_concat_grad");
  const long long per_cloud = (long long)m * nsample;
  const int lead = use_xyz ? 3 : 0;     // the coordinate channels of the grouped tensor carry no gradient
  return group_dispatch(true, b, c, n, per_cloud, grad_out ? grad_out + (size_t)lead * per_cloud : nullptr, idx, grad_features,
                        to_stream(stream), "group_concat_grad", (long long)(lead + c) * per_cloud);
}

WS3D_API int ws3d_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out,
                                ws3d_stream_t stream) {
  return group_dispatch(false, b, c, n, (long long)npoints, points, idx, out, to_stream(stream), "gather_points");
}

WS3D_API int ws3d_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx,
                                     float *grad_points, ws3d_stream_t stream) {
  return group_dispatch(true, b, c, n, (long long)npoints, grad_out, idx, grad_points, to_stream(stream),
                        "gather_points_grad");
}

namespace ws3d {
namespace {
int three_nn_dispatch(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, float *weight,
                      cudaStream_t stream) {
  if (b < 0 || n < 0 || m < 0) return fail_arg("three_nn");
  if (b == 0 || n == 0) return 0;
  if (!unknown || (!dist2 && !weight) || !idx || (m > 0 && !known)) return fail_arg("three_nn (null pointer)");
  if (b > 65535) return fail_arg("three_nn (batch > 65535)");
  if (three_nn_grid_applicable(b, n, m)) return three_nn_grid(b, n, m, unknown, known, dist2, idx, weight, stream);
  const size_t smem = (size_t)(m < kNnChunk ? (m > 0 ? m : 1) : kNnChunk) * sizeof(float4);
  cudaError_t e = cudaFuncSetAttribute(three_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kNnChunk * sizeof(float4)));
  if (e != cudaSuccess) { set_error("three_nn: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  dim3 grid((unsigned)ceil_div(n, kNnThreads), (unsigned)b);
  three_nn_kernel<<<grid, kNnThreads, smem, stream>>>(n, m, unknown, known, dist2, idx, weight);
  return check_launch("three_nn");
}
}  // namespace
}  // namespace ws3d

WS3D_API int ws3d_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                           ws3d_stream_t stream) {
  if (b > 0 && n > 0 && !dist2) return fail_arg("three_nn (null pointer)");
  return three_nn_dispatch(b, n, m, unknown, known, dist2, idx, nullptr, to_stream(stream));
}

WS3D_API int ws3d_three_nn_weights(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                                   float *weight, ws3d_stream_t stream) {
  if (b > 0 && n > 0 && !weight) return fail_arg("three_nn_weights (null pointer)");
  return three_nn_dispatch(b, n, m, unknown, known, dist2, idx, weight, to_stream(stream));
}

static int three_interpolate_impl(int b, int c, int m, int n, const float *points, const int *idx, const float *weight,
                                  float *out, const InterpEpilogue *epi, ws3d_stream_t stream) {
  if (b < 0 || c < 0 || m < 0 || n < 0) return fail_arg("three_interpolate");
  if (b == 0 || c == 0 || n == 0) return 0;
  if (!points || !idx || !weight || !out) return fail_arg("three_interpolate (null pointer)");
  if (b > 65535 || ceil_div(c, kInterpChannels) > 65535) return fail_arg("three_interpolate (grid too large)");
  if (m > 0 && (size_t)m * 4 <= 32 * 1024 && n >= 256) {
    // channel rows per CTA: as many as fit in ~128 KB, at least 4 CTAs per SM worth of work overall
    int cb = (int)((128 * 1024) / ((size_t)m * 4));
    if (cb > 64) cb = 64;
    if (cb > c) cb = c;
    while (cb > 8 && (long long)ceil_div(c, cb) * b < 2LL * num_sms()) cb >>= 1;
    if (c % 4 == 0 && cb % 4 != 0) cb = cb > 4 ? (cb & ~3) : 4;   // keep the four-channel interleaved kernel available
    const int chunks = ceil_div(c, cb);
    int nsplit = 1;
    while ((long long)chunks * b * nsplit < 2LL * num_sms() && ceil_div(n, nsplit * 2) >= 2 * 512) nsplit <<= 1;
    const int n_per_cta = ceil_div(n, nsplit);
    const int threads = n_per_cta >= 8192 ? 1024 : 512;  // more warps hide the stencil loads when there is enough work
    const size_t smem = (size_t)cb * m * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(three_interpolate_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("three_interpolate: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    if (chunks <= 65535 * 32 && nsplit <= 65535 && c % 4 == 0 && cb % 4 == 0) {
      const int mode = epi ? (1 | ((epi->flags & 1) ? 2 : 0) | ((epi->flags & 2) ? 4 : 0)) : 0;
      const int threads_k = mode ? 512 : threads;   // the epilogue variants need more than the 64 registers of a 1024-thread CTA
      auto kern = mode == 0 ? three_interpolate_smem4_kernel<0> : mode == 1 ? three_interpolate_smem4_kernel<1>
                  : mode == 3 ? three_interpolate_smem4_kernel<3> : mode == 5 ? three_interpolate_smem4_kernel<5>
                                                                              : three_interpolate_smem4_kernel<7>;
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { set_error("three_interpolate: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
      dim3 grid((unsigned)chunks, (unsigned)nsplit, (unsigned)b);
      const InterpEpilogue none = {nullptr, nullptr, nullptr, 0};
      kern<<<grid, threads_k, smem, to_stream(stream)>>>(c, m, n, cb, n_per_cta, points, idx, weight, out, epi ? *epi : none);
      return check_launch("three_interpolate");
    }
    if (epi) return fail_arg("three_interpolate_affine (needs c % 4 == 0)");
    if (chunks <= 65535 * 32 && nsplit <= 65535) {
      dim3 grid((unsigned)chunks, (unsigned)nsplit, (unsigned)b);
      three_interpolate_smem_kernel<<<grid, threads, smem, to_stream(stream)>>>(c, m, n, cb, n_per_cta, points, idx,
                                                                                         weight, out);
      return check_launch("three_interpolate");
    }
  }
  if (epi) return fail_arg("three_interpolate_affine (shape outside the shared-memory kernel: m <= 8192, n >= 256, c % 4 == 0)");
  dim3 grid((unsigned)ceil_div(n, kInterpThreads), (unsigned)ceil_div(c, kInterpChannels), (unsigned)b);
  three_interpolate_kernel<<<grid, kInterpThreads, 0, to_stream(stream)>>>(c, m, n, points, idx, weight, out);
  return check_launch("three_interpolate");
}

WS3D_API int ws3d_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                                    const float *weight, float *out, ws3d_stream_t stream) {
  return three_interpolate_impl(b, c, m, n, points, idx, weight, out, nullptr, stream);
}

// Extension: three_interpolate with an affine epilogue,
//   out[b,c,i] = act( sum_k weight[b,i,k] * points[b,c,idx[b,i,k]] + scale1[c] * row1[b,i] + shift[c] ),
// scale1 / row1 / shift optional; flags bit 0 = ReLU, bit 1 = round to TF32.  Shapes served: m <= 8192, n >= 256, c % 4 == 0.
WS3D_API int ws3d_three_interpolate_affine(int b, int c, int m, int n, const float *points, const int *idx,
                                           const float *weight, const float *scale1, const float *row1, const float *shift,
                                           int flags, float *out, ws3d_stream_t stream) {
  if ((scale1 == nullptr) != (row1 == nullptr)) return fail_arg("three_interpolate_affine (scale1 and row1 go together)");
  const InterpEpilogue epi = {scale1, row1, shift, flags};
  return three_interpolate_impl(b, c, m, n, points, idx, weight, out, &epi, stream);
}

WS3D_API int ws3d_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                         const float *weight, float *grad_points, ws3d_stream_t stream) {
  if (b < 0 || c < 0 || m < 0 || n < 0) return fail_arg("three_interpolate_grad");
  if (b == 0 || c == 0 || n == 0) return 0;
  if (!grad_out || !idx || !weight || !grad_points) return fail_arg("three_interpolate_grad (null pointer)");
  // gather over the inverse stencil (interp_grad.cu): no atomics, fixed summation order; 1 = shape outside its limits
  static const bool force_atomic = [] { const char *e = getenv("WS3D_INTERP_GRAD_ATOMIC"); return e && e[0] == '1'; }();
  if (!force_atomic) {
    const int rc = three_interpolate_grad_gather(b, c, n, m, grad_out, idx, weight, grad_points, to_stream(stream));
    if (rc != 1) return rc;
  }
  if (b > 65535 || ceil_div(c, kInterpChannels) > 65535) return fail_arg("three_interpolate_grad (grid too large)");
  dim3 grid((unsigned)ceil_div(n, kInterpThreads), (unsigned)ceil_div(c, kInterpChannels), (unsigned)b);
  three_interpolate_grad_kernel<<<grid, kInterpThreads, 0, to_stream(stream)>>>(c, n, m, grad_out, idx, weight, grad_points);
  return check_launch("three_interpolate_grad");
}
