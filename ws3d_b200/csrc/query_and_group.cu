// QueryAndGroup.forward (pointnet2_lib/pointnet2/pointnet2_utils.py:241-264) as two launches:
// ball_query (pointnet2_ops.cu) and ONE grouping pass that gathers the neighbour coordinates,
// subtracts the ball centre, gathers every feature channel and writes the concatenated
// (B, 3+C, M, K) tensor once -- the reference runs transpose + group(xyz) + subtract +
// group(features) + cat, i.e. writes the big tensor about twice and re-reads idx per channel.
#include "common.cuh"

namespace ws3d {

int ball_query_multi(int nr, int b, int n, int m, const float *radius, const int *nsample, const float *new_xyz,
                     const float *xyz, int *const *idx, cudaStream_t stream);

namespace {

constexpr int kThreads = 256;

// Each thread owns EPT consecutive neighbours of one centre (K % EPT == 0).
template <int EPT>
__global__ void __launch_bounds__(kThreads) group_concat_kernel(int c, int n, int m, int K, int use_xyz, int c_per_cta,
                                                                 const float *__restrict__ xyz,
                                                                 const float *__restrict__ new_xyz,
                                                                 const float *__restrict__ features,
                                                                 const int *__restrict__ idx, float *__restrict__ out) {
  const size_t cloud = blockIdx.y;
  const long long per_cloud = (long long)m * K;
  const long long e0 = ((long long)blockIdx.x * kThreads + threadIdx.x) * EPT;
  if (e0 >= per_cloud) return;
  const int j = (int)(e0 / K);
  int id[EPT];
  const int *ip = idx + cloud * per_cloud + e0;
  if (EPT == 4) {
    const int4 v = __ldg(reinterpret_cast<const int4 *>(ip));
    id[0] = v.x; id[1] = v.y; id[2] = v.z; id[3] = v.w;
  } else {
#pragma unroll
    for (int t = 0; t < EPT; ++t) id[t] = __ldg(ip + t);
  }
  const int cout = c + (use_xyz ? 3 : 0);
  float *dst = out + cloud * (size_t)cout * per_cloud + e0;
  const int cb = blockIdx.z * c_per_cta, ce = min(c, cb + c_per_cta);
  if (use_xyz && blockIdx.z == 0) {
    const float *ctr = new_xyz + (cloud * (size_t)m + j) * 3;
    const float *pts = xyz + cloud * (size_t)n * 3;
    float v[3][EPT];
#pragma unroll
    for (int t = 0; t < EPT; ++t) {
      const float *p = pts + (size_t)id[t] * 3;
      v[0][t] = __ldg(p); v[1][t] = __ldg(p + 1); v[2][t] = __ldg(p + 2);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float cc = __ldg(ctr + a);
#pragma unroll
      for (int t = 0; t < EPT; ++t) v[a][t] = __fsub_rn(v[a][t], cc);
      if (EPT == 4) __stcs(reinterpret_cast<float4 *>(dst), make_float4(v[a][0], v[a][1], v[a][2], v[a][3]));
      else {
#pragma unroll
        for (int t = 0; t < EPT; ++t) dst[t] = v[a][t];
      }
      dst += per_cloud;
    }
  } else if (use_xyz) {
    dst += 3 * per_cloud;
  }
  if (cb < ce) {
    const float *src = features + cloud * (size_t)c * n;
    dst += (size_t)cb * per_cloud;
#pragma unroll 4
    for (int ci = cb; ci < ce; ++ci) {
      const float *row = src + (size_t)ci * n;
      if (EPT == 4) {
        float4 f;
        f.x = __ldg(row + id[0]); f.y = __ldg(row + id[1]); f.z = __ldg(row + id[2]); f.w = __ldg(row + id[3]);
        __stcs(reinterpret_cast<float4 *>(dst), f);
      } else {
#pragma unroll
        for (int t = 0; t < EPT; ++t) dst[t] = __ldg(row + id[t]);
      }
      dst += per_cloud;
    }
  }
}

// Gather with an affine epilogue:
//   out[b,c,j,s] = act( P[b,c,i] + wx[c] . (xyz[b,i] - new_xyz[b,j]) + shift[c] ),  i = idx[b,j,s];
// flags bit 0 = ReLU, bit 1 = round to TF32.  It is the FIRST LAYER of a set-abstraction MLP with the feature part of
// the 1x1 convolution moved in front of the grouping: W [xyz[i] - centre ; f[i]] = (W_f f)[i] + W_x (xyz[i] - centre),
// so P = W_f f is formed on the n source points (instead of a GEMM over the m * nsample grouped columns), the three
// coordinate channels are applied here in FP32 (they must not go through TF32: the difference of two large coordinates
// would lose its low bits), and the grouped (3 + C)-channel tensor is never written.  Thread mapping of group_concat_kernel.
template <int EPT>
__global__ void __launch_bounds__(kThreads) group_affine_kernel(int c, int n, int m, int K, int c_per_cta, int flags,
                                                                 const float *__restrict__ P, const float *__restrict__ xyz,
                                                                 const float *__restrict__ new_xyz,
                                                                 const float *__restrict__ wx, const float *__restrict__ shift,
                                                                 const int *__restrict__ idx, float *__restrict__ out) {
  const size_t cloud = blockIdx.y;
  const long long per_cloud = (long long)m * K;
  const long long e0 = ((long long)blockIdx.x * kThreads + threadIdx.x) * EPT;
  if (e0 >= per_cloud) return;
  const int j = (int)(e0 / K);
  int id[EPT];
  const int *ip = idx + cloud * per_cloud + e0;
  if (EPT == 4) {
    const int4 v = __ldg(reinterpret_cast<const int4 *>(ip));
    id[0] = v.x; id[1] = v.y; id[2] = v.z; id[3] = v.w;
  } else {
#pragma unroll
    for (int t = 0; t < EPT; ++t) id[t] = __ldg(ip + t);
  }
  float d[3][EPT];
  {
    const float *ctr = new_xyz + (cloud * (size_t)m + j) * 3;
    const float *pts = xyz + cloud * (size_t)n * 3;
    const float c0 = __ldg(ctr), c1 = __ldg(ctr + 1), c2 = __ldg(ctr + 2);
#pragma unroll
    for (int t = 0; t < EPT; ++t) {
      const float *p = pts + (size_t)id[t] * 3;
      d[0][t] = __fsub_rn(__ldg(p), c0); d[1][t] = __fsub_rn(__ldg(p + 1), c1); d[2][t] = __fsub_rn(__ldg(p + 2), c2);
    }
  }
  const int cb = blockIdx.z * c_per_cta, ce = min(c, cb + c_per_cta);
  const float *src = P + cloud * (size_t)c * n;
  float *dst = out + (cloud * (size_t)c + cb) * per_cloud + e0;
  const bool relu = (flags & 1) != 0, rnd = (flags & 2) != 0;
#pragma unroll 4
  for (int ci = cb; ci < ce; ++ci) {
    const float *row = src + (size_t)ci * n;
    const float w0 = __ldg(wx + ci * 3), w1 = __ldg(wx + ci * 3 + 1), w2 = __ldg(wx + ci * 3 + 2), sh = __ldg(shift + ci);
    float v[EPT];
#pragma unroll
    for (int t = 0; t < EPT; ++t) {
      v[t] = __ldg(row + id[t]) + fmaf(w2, d[2][t], fmaf(w1, d[1][t], fmaf(w0, d[0][t], sh)));
      if (relu) v[t] = fmaxf(v[t], 0.f);
      if (rnd) v[t] = __uint_as_float((__float_as_uint(v[t]) + 0x1000u) & 0xFFFFE000u);
    }
    if (EPT == 4) __stcs(reinterpret_cast<float4 *>(dst), make_float4(v[0], v[1], v[2], v[3]));
    else {
#pragma unroll
      for (int t = 0; t < EPT; ++t) dst[t] = v[t];
    }
    dst += per_cloud;
  }
}

}  // namespace

int group_concat(int b, int n, int m, int c, int K, int use_xyz, const float *xyz, const float *new_xyz,
                 const float *features, const int *idx, float *out, cudaStream_t stream) {
  if (b == 0 || m == 0 || K == 0) return 0;
  const long long per_cloud = (long long)m * K;
  const bool vec = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15u) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
  const long long gx = (per_cloud + kThreads * (vec ? 4 : 1) - 1) / (kThreads * (vec ? 4 : 1));
  int c_per_cta = c > 0 ? c : 1;
  if (gx * b < 4LL * num_sms() && c > 8) {
    const long long splits = (4LL * num_sms() + gx * b - 1) / (gx * b);
    c_per_cta = (int)((c + splits - 1) / splits);
    if (c_per_cta < 8) c_per_cta = 8;
  }
  dim3 grid((unsigned)gx, (unsigned)b, (unsigned)(c > 0 ? ceil_div(c, c_per_cta) : 1));
  if (vec) group_concat_kernel<4><<<grid, kThreads, 0, stream>>>(c, n, m, K, use_xyz, c_per_cta, xyz, new_xyz, features, idx, out);
  else group_concat_kernel<1><<<grid, kThreads, 0, stream>>>(c, n, m, K, use_xyz, c_per_cta, xyz, new_xyz, features, idx, out);
  return check_launch("query_and_group (group)");
}

}  // namespace ws3d

using namespace ws3d;

WS3D_API int ws3d_group_concat(int b, int n, int m, int c, int nsample, int use_xyz, const float *xyz,
                               const float *new_xyz, const float *features, const int *idx, float *out,
                               ws3d_stream_t stream) {
  if (b < 0 || n < 0 || m < 0 || c < 0 || nsample < 0) return fail_arg("group_concat");
  if (b > 65535) return fail_arg("group_concat (batch > 65535)");
  if (b && m && nsample && (!idx || !out || (use_xyz && (!xyz || !new_xyz)) || (c > 0 && !features)))
    return fail_arg("group_concat (null pointer)");
  if (!use_xyz && c == 0) return fail_arg("group_concat (no channels)");
  return group_concat(b, n, m, c, nsample, use_xyz, xyz, new_xyz, features, idx, out, to_stream(stream));
}

WS3D_API int ws3d_query_and_group(int b, int n, int m, int c, float radius, int nsample, int use_xyz,
                                  const float *xyz, const float *new_xyz, const float *features, float *out,
                                  int *idx_out, ws3d_stream_t stream) {
  if (b < 0 || n < 0 || m < 0 || c < 0 || nsample < 0) return fail_arg("query_and_group");
  if (b == 0 || m == 0 || nsample == 0) return 0;
  int *idx = idx_out;
  if (!idx) {
    const size_t bytes = (size_t)b * m * nsample * sizeof(int);
    idx = (int *)scratch(bytes, 1);
    if (!idx) return (int)cudaErrorMemoryAllocation;
    cudaError_t e = cudaMemsetAsync(idx, 0, bytes, to_stream(stream));
    if (e != cudaSuccess) { set_error("query_and_group: memset: %s", cudaGetErrorString(e)); return (int)e; }
  }
  int *rows[1] = {idx};
  int rc = ball_query_multi(1, b, n, m, &radius, &nsample, new_xyz, xyz, rows, to_stream(stream));
  if (rc) return rc;
  return ws3d_group_concat(b, n, m, c, nsample, use_xyz, xyz, new_xyz, features, idx, out, stream);
}

// P (B,c,n), xyz (B,n,3), new_xyz (B,m,3), wx (c,3), shift (c), idx (B,m,nsample) -> out (B,c,m,nsample); see group_affine_kernel.
WS3D_API int ws3d_group_affine(int b, int n, int m, int c, int nsample, const float *P, const float *xyz, const float *new_xyz,
                               const float *wx, const float *shift, const int *idx, int flags, float *out, ws3d_stream_t stream) {
  if (b < 0 || n < 0 || m < 0 || c <= 0 || nsample < 0 || b > 65535) return fail_arg("group_affine");
  if (b == 0 || m == 0 || nsample == 0) return 0;
  if (!P || !xyz || !new_xyz || !wx || !shift || !idx || !out) return fail_arg("group_affine (null pointer)");
  const long long per_cloud = (long long)m * nsample;
  const bool vec = (nsample % 4 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15u) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
  const long long gx = (per_cloud + kThreads * (vec ? 4 : 1) - 1) / (kThreads * (vec ? 4 : 1));
  int c_per_cta = c;
  if (gx * b < 4LL * num_sms() && c > 8) {
    const long long splits = (4LL * num_sms() + gx * b - 1) / (gx * b);
    c_per_cta = (int)((c + splits - 1) / splits);
    if (c_per_cta < 8) c_per_cta = 8;
  }
  dim3 grid((unsigned)gx, (unsigned)b, (unsigned)ceil_div(c, c_per_cta));
  if (vec) group_affine_kernel<4><<<grid, kThreads, 0, to_stream(stream)>>>(c, n, m, nsample, c_per_cta, flags, P, xyz, new_xyz, wx, shift, idx, out);
  else group_affine_kernel<1><<<grid, kThreads, 0, to_stream(stream)>>>(c, n, m, nsample, c_per_cta, flags, P, xyz, new_xyz, wx, shift, idx, out);
  return check_launch("group_affine");
}

namespace ws3d {
namespace {
// pc (B,N,3+C) -> xyz (B,N,3), features (B,C,N): thread per point (lib/net/pointnet2_msg.py:52-60)
__global__ void __launch_bounds__(256) split_pointcloud_kernel(int n, int c, const float *__restrict__ pc, float *__restrict__ xyz,
                                                                float *__restrict__ feat) {
  const size_t cloud = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float *p = pc + (cloud * (size_t)n + i) * (size_t)(3 + c);
  float *x = xyz + (cloud * (size_t)n + i) * 3;
  x[0] = __ldg(p); x[1] = __ldg(p + 1); x[2] = __ldg(p + 2);
  for (int k = 0; k < c; ++k) feat[(cloud * (size_t)c + k) * n + i] = __ldg(p + 3 + k);
}

// The inverse direction: xyz (B,N,3) + channel-major features (B,C,N) -> point-major rows (B,N,ld) = [xyz | features | zeros], the
// operand layout the fused set-abstraction kernel gathers from (one contiguous row per neighbour instead of one 32-byte sector per
// channel).  32 x 32 tiles through shared memory: reads are coalesced along the points, writes along the row.
__global__ void __launch_bounds__(256) pack_rows_kernel(int n, int c, int ld, const float *__restrict__ xyz, const float *__restrict__ feat,
                                                        float *__restrict__ rows) {
  __shared__ float tile[32][33];
  const size_t cloud = blockIdx.z;
  const int i0 = blockIdx.x * 32, q0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    const int q = q0 + ty + r, i = i0 + tx;     // column q of the row of point i
    float v = 0.f;
    if (i < n) {
      if (q < 3) v = __ldg(xyz + (cloud * (size_t)n + i) * 3 + q);
      else if (q < 3 + c) v = __ldg(feat + (cloud * (size_t)c + (q - 3)) * n + i);
    }
    tile[ty + r][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    const int i = i0 + ty + r, q = q0 + tx;
    if (i < n && q < ld) rows[(cloud * (size_t)n + i) * ld + q] = tile[tx][ty + r];
  }
}
}  // namespace
}  // namespace ws3d

WS3D_API int ws3d_pack_rows(int b, int n, int c, int ld, const float *xyz, const float *features, float *rows, ws3d_stream_t stream) {
  if (b < 0 || n < 0 || c < 0 || ld < 3 + c || b > 65535) return fail_arg("pack_rows");
  if (b == 0 || n == 0) return 0;
  if (!xyz || !rows || (c > 0 && !features)) return fail_arg("pack_rows (null pointer)");
  pack_rows_kernel<<<dim3((unsigned)ceil_div(n, 32), (unsigned)ceil_div(ld, 32), (unsigned)b), 256, 0, to_stream(stream)>>>(n, c, ld, xyz,
                                                                                                                        features, rows);
  return check_launch("pack_rows");
}

WS3D_API int ws3d_split_pointcloud(int b, int n, int c, const float *pc, float *xyz, float *features, ws3d_stream_t stream) {
  if (b < 0 || n < 0 || c < 0 || b > 65535) return fail_arg("split_pointcloud");
  if (b == 0 || n == 0) return 0;
  if (!pc || !xyz || (c > 0 && !features)) return fail_arg("split_pointcloud (null pointer)");
  split_pointcloud_kernel<<<dim3((unsigned)ceil_div(n, 256), (unsigned)b), 256, 0, to_stream(stream)>>>(n, c, pc, xyz, features);
  return check_launch("split_pointcloud");
}
