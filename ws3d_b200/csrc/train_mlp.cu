// Training-mode shared MLP (BASELINE config 3):
//
//     conv1x1 (no bias under BN) -> BatchNorm with BATCH statistics -> ReLU [-> max-pool over nsample]
//
// forward and backward, on the channel-major (B, C, cols) activations every other kernel of this library uses.  The
// reference runs each layer as cuDNN conv -> BN kernel -> ReLU kernel (-> max-pool kernel) plus autograd's mirror
// images (pytorch_utils.py:5-32, pointnet2_modules.py:40-44); round 1 of this repo kept that path and only changed
// the memory format.  Here:
//
//   forward   Y = W X                 ws3d_mlp_layer_stats: the tcgen05 layer kernel (mlp_tc.cu) with the per-channel
//                                     sum / sum of squares of its accumulator tile reduced in the epilogue
//             mean, invstd, running   bn_finalize_kernel (one thread per channel)
//             Z = relu(Y*s + t)       bn_relu_apply_kernel; the pooled variant also emits the arg-max of every group
//   backward  s1 = sum dA, s2 = sum dA*xhat   bn_relu_bwd_reduce_kernel (dA = dZ under the ReLU mask; pooled: dZ lands
//                                     on the arg-max column only, so only that column is read)
//             dY = s*(dA - s1/N - xhat*s2/N)   bn_relu_bwd_apply_kernel (TF32-rounded: it feeds two tensor-core GEMMs)
//             dX = W^T dY             ws3d_mlp_layer on the transposed weights
//             dW = dY X^T             mlp_wgrad_kernel below: tcgen05, both operands K-major (the contraction runs over
//                                     the contiguous column axis), split-K over the CTAs, FP32 atomics into dW
//
// Numerics: TF32 operands with FP32 accumulation for the three GEMMs -- the class PyTorch's cuDNN convolutions use by
// default, forward and backward (torch.backends.cudnn.allow_tf32); statistics are accumulated in double.
#include <cuda.h>
#include <math.h>

#include "common.cuh"
#include "tma.cuh"

namespace ws3d {
namespace {

constexpr int kEwThreads = 256;

// ---- BatchNorm statistics -> per-channel affine ------------------------------------------------------
// stats = [sum(c) | sumsq(c)] in double.  scale = gamma * invstd, shift = beta - mean * scale; running statistics as
// torch.nn.BatchNorm does (momentum update, UNBIASED variance for running_var).
__global__ void bn_finalize_kernel(int c, double count, const double *__restrict__ stats, const float *__restrict__ gamma,
                                   const float *__restrict__ beta, float eps, float momentum, float *__restrict__ running_mean,
                                   float *__restrict__ running_var, float *__restrict__ scale, float *__restrict__ shift,
                                   float *__restrict__ mean_out, float *__restrict__ invstd_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const double mean = stats[i] / count;
  double var = stats[c + i] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[i] : 1.f, b = beta ? beta[i] : 0.f;
  const float sc = g * invstd;
  scale[i] = sc;
  shift[i] = b - (float)mean * sc;
  mean_out[i] = (float)mean;
  invstd_out[i] = invstd;
  if (running_mean) running_mean[i] = (1.f - momentum) * running_mean[i] + momentum * (float)mean;
  if (running_var) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[i] = (1.f - momentum) * running_var[i] + momentum * (float)unbiased;
  }
}

// ---- forward: Z = act(Y * scale[c] + shift[c]) -------------------------------------------------------
// grid (col blocks, c, b); a thread owns four consecutive columns.
__global__ void __launch_bounds__(kEwThreads) bn_relu_apply_kernel(int c, int cols, int relu, int round_out, const float *__restrict__ y,
                                                                    const float *__restrict__ scale, const float *__restrict__ shift,
                                                                    float *__restrict__ z) {
  const int ch = blockIdx.y;
  const size_t row = ((size_t)blockIdx.z * c + ch) * (size_t)cols;
  const float s = __ldg(scale + ch), t = __ldg(shift + ch);
  for (int e = 4 * (blockIdx.x * kEwThreads + threadIdx.x); e < cols; e += 4 * kEwThreads * gridDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4 *>(y + row + e));
    v.x = fmaf(v.x, s, t); v.y = fmaf(v.y, s, t); v.z = fmaf(v.z, s, t); v.w = fmaf(v.w, s, t);
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (round_out) {   // nearest TF32: the next layer's tensor core would otherwise truncate
      v.x = __uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xFFFFE000u);
      v.y = __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xFFFFE000u);
      v.z = __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xFFFFE000u);
      v.w = __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xFFFFE000u);
    }
    *reinterpret_cast<float4 *>(z + row + e) = v;
  }
}

// Pooled variant: Zp[b, c, g] = max over the `pool` columns of group g of act(Y*s + t), arg[b, c, g] = the FIRST column
// of the group that attains it (F.max_pool2d keeps the first maximum).  Lanes read consecutive float4s (coalesced) and
// the pool / 4 lanes of a group combine (value, index) by butterfly.  pool in {4, 8, ..., 128}.
__global__ void __launch_bounds__(kEwThreads) bn_relu_pool_kernel(int c, int cols, int pool, int relu, const float *__restrict__ y,
                                                                   const float *__restrict__ scale, const float *__restrict__ shift,
                                                                   float *__restrict__ zp, unsigned char *__restrict__ arg) {
  const int ch = blockIdx.y, lane = threadIdx.x & 31;
  const size_t row = ((size_t)blockIdx.z * c + ch) * (size_t)cols;
  const size_t prow = ((size_t)blockIdx.z * c + ch) * (size_t)(cols / pool);
  const float s = __ldg(scale + ch), t = __ldg(shift + ch);
  const int lanes_per_group = pool >> 2;                 // 1 .. 32
  const int iters = (cols / 4 + kEwThreads * gridDim.x - 1) / (kEwThreads * gridDim.x);
  for (int it = 0; it < iters; ++it) {                   // every lane of a warp runs the same number of rounds (shuffles)
    const int q = (it * gridDim.x + blockIdx.x) * kEwThreads + threadIdx.x;   // float4 index in the row
    const int e = 4 * q;
    float best = -INFINITY;
    int bi = 0;
    if (e < cols) {
      const float4 v4 = __ldg(reinterpret_cast<const float4 *>(y + row + e));
      float v[4] = {fmaf(v4.x, s, t), fmaf(v4.y, s, t), fmaf(v4.z, s, t), fmaf(v4.w, s, t)};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (relu) v[k] = fmaxf(v[k], 0.f);
        if (v[k] > best) { best = v[k]; bi = (e + k) & (pool - 1); }
      }
    }
    for (int o = 1; o < lanes_per_group; o <<= 1) {
      const float ov = __shfl_xor_sync(0xFFFFFFFFu, best, o);
      const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (e < cols && (lane & (lanes_per_group - 1)) == 0) {
      zp[prow + e / pool] = best;
      arg[prow + e / pool] = (unsigned char)bi;
    }
  }
}

// ---- backward -----------------------------------------------------------------------------------------
// sums[c] += sum dA, sums[C + c] += sum dA * xhat, dA = dZ where act'(.) != 0, xhat = (Y - mean) * invstd.
// grid (splits, c): CTA (split, ch) walks its share of the (b, column) pairs of channel ch.
__global__ void __launch_bounds__(kEwThreads) bn_relu_bwd_reduce_kernel(int b, int c, int cols, int relu, const float *__restrict__ y,
                                                                         const float *__restrict__ dz, const float *__restrict__ scale,
                                                                         const float *__restrict__ shift, const float *__restrict__ mean,
                                                                         const float *__restrict__ invstd, double *__restrict__ sums) {
  __shared__ double s_a[kEwThreads / 32], s_b[kEwThreads / 32];
  const int ch = blockIdx.y;
  const float s = __ldg(scale + ch), t = __ldg(shift + ch), mu = __ldg(mean + ch), is = __ldg(invstd + ch);
  const int q_per_row = cols >> 2;
  const long long total = (long long)b * q_per_row;
  float a1 = 0.f, a2 = 0.f;
  for (long long q = (long long)blockIdx.x * kEwThreads + threadIdx.x; q < total; q += (long long)kEwThreads * gridDim.x) {
    const int bb = (int)(q / q_per_row), e = 4 * (int)(q - (long long)bb * q_per_row);
    const size_t off = ((size_t)bb * c + ch) * (size_t)cols + e;
    const float4 yv = __ldg(reinterpret_cast<const float4 *>(y + off));
    const float4 gv = __ldg(reinterpret_cast<const float4 *>(dz + off));
    const float ys[4] = {yv.x, yv.y, yv.z, yv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float da = (!relu || fmaf(ys[k], s, t) > 0.f) ? gs[k] : 0.f;
      a1 += da;
      a2 = fmaf(da, (ys[k] - mu) * is, a2);
    }
  }
  double d1 = a1, d2 = a2;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    d1 += __shfl_xor_sync(0xFFFFFFFFu, d1, o);
    d2 += __shfl_xor_sync(0xFFFFFFFFu, d2, o);
  }
  if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = d1; s_b[threadIdx.x >> 5] = d2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0, t2 = 0;
    for (int w = 0; w < kEwThreads / 32; ++w) { t1 += s_a[w]; t2 += s_b[w]; }
    atomicAdd(sums + ch, t1);
    atomicAdd(sums + c + ch, t2);
  }
}

// Pooled variant: dZp (b, c, cols / pool) lands on column arg of every group.
__global__ void __launch_bounds__(kEwThreads) bn_relu_pool_bwd_reduce_kernel(int b, int c, int cols, int pool, int relu,
                                                                              const float *__restrict__ y, const float *__restrict__ dzp,
                                                                              const unsigned char *__restrict__ arg,
                                                                              const float *__restrict__ scale, const float *__restrict__ shift,
                                                                              const float *__restrict__ mean, const float *__restrict__ invstd,
                                                                              double *__restrict__ sums) {
  __shared__ double s_a[kEwThreads / 32], s_b[kEwThreads / 32];
  const int ch = blockIdx.y;
  const float s = __ldg(scale + ch), t = __ldg(shift + ch), mu = __ldg(mean + ch), is = __ldg(invstd + ch);
  const int groups = cols / pool;
  const long long total = (long long)b * groups;
  float a1 = 0.f, a2 = 0.f;
  for (long long q = (long long)blockIdx.x * kEwThreads + threadIdx.x; q < total; q += (long long)kEwThreads * gridDim.x) {
    const int bb = (int)(q / groups), g = (int)(q - (long long)bb * groups);
    const size_t prow = ((size_t)bb * c + ch) * (size_t)groups + g;
    const float yv = __ldg(y + ((size_t)bb * c + ch) * (size_t)cols + (size_t)g * pool + __ldg(arg + prow));
    const float da = (!relu || fmaf(yv, s, t) > 0.f) ? __ldg(dzp + prow) : 0.f;
    a1 += da;
    a2 = fmaf(da, (yv - mu) * is, a2);
  }
  double d1 = a1, d2 = a2;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    d1 += __shfl_xor_sync(0xFFFFFFFFu, d1, o);
    d2 += __shfl_xor_sync(0xFFFFFFFFu, d2, o);
  }
  if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = d1; s_b[threadIdx.x >> 5] = d2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0, t2 = 0;
    for (int w = 0; w < kEwThreads / 32; ++w) { t1 += s_a[w]; t2 += s_b[w]; }
    atomicAdd(sums + ch, t1);
    atomicAdd(sums + c + ch, t2);
  }
}

// dY = scale * (dA - s1 / N - xhat * s2 / N)   (inv_count = 0: no batch-norm coupling -- a plain bias layer: dY = dA)
__global__ void __launch_bounds__(kEwThreads) bn_relu_bwd_apply_kernel(int c, int cols, int pool, int relu, float inv_count,
                                                                        const float *__restrict__ y, const float *__restrict__ dz,
                                                                        const unsigned char *__restrict__ arg,
                                                                        const float *__restrict__ scale, const float *__restrict__ shift,
                                                                        const float *__restrict__ mean, const float *__restrict__ invstd,
                                                                        const double *__restrict__ sums, float *__restrict__ dy) {
  const int ch = blockIdx.y;
  const size_t row = ((size_t)blockIdx.z * c + ch) * (size_t)cols;
  const size_t prow = pool ? ((size_t)blockIdx.z * c + ch) * (size_t)(cols / pool) : 0;
  const float s = __ldg(scale + ch), t = __ldg(shift + ch), mu = __ldg(mean + ch), is = __ldg(invstd + ch);
  const float m1 = (float)(sums[ch] * (double)inv_count), m2 = (float)(sums[c + ch] * (double)inv_count);
  for (int e = 4 * (blockIdx.x * kEwThreads + threadIdx.x); e < cols; e += 4 * kEwThreads * gridDim.x) {
    const float4 yv = __ldg(reinterpret_cast<const float4 *>(y + row + e));
    const float ys[4] = {yv.x, yv.y, yv.z, yv.w};
    float gs[4];
    if (pool) {
      const int g = e / pool;                               // pool % 4 == 0: the four columns share a group
      const int a = __ldg(arg + prow + g) - (e & (pool - 1));
      const float gz = __ldg(dz + prow + g);
#pragma unroll
      for (int k = 0; k < 4; ++k) gs[k] = (a == k) ? gz : 0.f;
    } else {
      const float4 gv = __ldg(reinterpret_cast<const float4 *>(dz + row + e));
      gs[0] = gv.x; gs[1] = gv.y; gs[2] = gv.z; gs[3] = gv.w;
    }
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float da = (!relu || fmaf(ys[k], s, t) > 0.f) ? gs[k] : 0.f;
      const float v = s * (da - m1 - (ys[k] - mu) * is * m2);
      o[k] = (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u;   // nearest TF32 (operand of the two gradient GEMMs)
    }
    *reinterpret_cast<uint4 *>(dy + row + e) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---- dW = dY X^T on the tensor cores -------------------------------------------------------------------
// D[128 x NT] (TMEM) += A[128 x 32] (dY rows = output channels) * B[NT x 32]^T (X rows = input channels) per 32-column
// chunk; both operands K-major with 128-byte swizzle (one chunk row = 32 floats = 128 bytes).  The contraction axis
// (b x cols, up to 2 M) is split over the CTAs of a tile; every CTA adds its partial tile into dW with red.global.add.
constexpr int kWgTileM = 128, kWgTileN = 128, kWgChunk = 32, kWgStages = 6, kWgThreads = 192;

struct WgradParams {
  int c_out, c_in, ldw;
  int chunks_per_cloud, total_chunks, chunks_per_cta;
  int n_m_tiles, n_n_tiles;
  float *dw;
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint64_t smem_desc_k128(uint32_t addr) {   // K-major, SWIZZLE_128B: 8-row groups 1 KB apart
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)(16u >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit_arrive(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1) mlp_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy,
                                                                   const __grid_constant__ CUtensorMap map_x, const WgradParams prm) {
  constexpr uint32_t kABytes = kWgTileM * kWgChunk * 4, kBBytes = kWgTileN * kWgChunk * 4, kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kWgTileN >> 3) << 17) | ((uint32_t)(kWgTileM >> 4) << 24);
  extern __shared__ uint8_t s_raw[];
  __shared__ __align__(8) unsigned long long s_full[kWgStages], s_empty[kWgStages], s_done;
  __shared__ uint32_t s_tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t stage_base = (smem_u32(s_raw) + 1023u) & ~1023u;
  const int tile = blockIdx.y, mt = tile % prm.n_m_tiles, nt = tile / prm.n_m_tiles;
  const int m0 = mt * kWgTileM, n0 = nt * kWgTileN;
  const int k_begin = blockIdx.x * prm.chunks_per_cta;
  const int k_end = min(prm.total_chunks, k_begin + prm.chunks_per_cta);
  const int nk = k_end - k_begin;      // >= 1 by construction of the grid

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(smem_u32(&s_full[s]), 1); mbar_init(smem_u32(&s_empty[s]), 1); }
    mbar_init(smem_u32(&s_done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(kWgTileN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nk; ++i) {
        const int s = i % kWgStages;
        if (i >= kWgStages) mbar_wait(smem_u32(&s_empty[s]), (uint32_t)((i / kWgStages) - 1) & 1u);
        const int chunk = k_begin + i, cloud = chunk / prm.chunks_per_cloud, col0 = (chunk - cloud * prm.chunks_per_cloud) * kWgChunk;
        const uint32_t bar = smem_u32(&s_full[s]);
        const uint32_t a_dst = stage_base + (uint32_t)s * kStageBytes, b_dst = a_dst + kABytes;
        mbar_expect_tx(bar, kStageBytes);               // out-of-range rows / columns are zero-filled and still counted
        tma_load_3d(a_dst, &map_dy, col0, m0, cloud, bar);
        tma_load_3d(b_dst, &map_x, col0, n0, cloud, bar);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < nk; ++i) {
        const int s = i % kWgStages;
        mbar_wait(smem_u32(&s_full[s]), (uint32_t)(i / kWgStages) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_base = stage_base + (uint32_t)s * kStageBytes, b_base = a_base + kABytes;
#pragma unroll
        for (int kk = 0; kk < kWgChunk / 8; ++kk)
          umma_tf32_ss(tmem_base, smem_desc_k128(a_base + (uint32_t)kk * 32u), smem_desc_k128(b_base + (uint32_t)kk * 32u), kIdesc,
                       (i | kk) != 0 ? 1u : 0u);
        umma_commit_arrive(smem_u32(&s_empty[s]));
      }
      umma_commit_arrive(smem_u32(&s_done));
    }
  } else {
    // epilogue: warp w drains TMEM lanes 32 * (w % 4) .. + 31 = output channels m0 + that; columns = input channels
    const int quarter = warp & 3;
    mbar_wait(smem_u32(&s_done), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int co = m0 + quarter * 32 + lane;
    const bool warp_live = m0 + quarter * 32 < prm.c_out;
    if (warp_live) {
      for (int cbase = 0; cbase < kWgTileN && n0 + cbase < prm.c_in; cbase += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)cbase, r);
        if (co < prm.c_out) {
          float *dst = prm.dw + (size_t)co * prm.ldw + n0 + cbase;
#pragma unroll
          for (int t = 0; t < 32; ++t)
            if (n0 + cbase + t < prm.c_in) atomicAdd(dst + t, __uint_as_float(r[t]));
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kWgTileN) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// (B, rows, cols) tensor, box = 32 columns x `box_rows` rows of one cloud, 128-byte swizzle
bool rows_map(CUtensorMap *m, const float *base, int b, int rows, int cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("mlp_wgrad: cuTensorMapEncodeTiled unavailable"); return false; }
  const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)b};
  const cuuint64_t strides[2] = {(cuuint64_t)cols * 4, (cuuint64_t)cols * rows * 4};
  const cuuint32_t box[3] = {(cuuint32_t)kWgChunk, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("mlp_wgrad: cuTensorMapEncodeTiled failed (%d)", (int)r); return false; }
  return true;
}

inline int ew_col_blocks(int cols, int c, int b) {
  int gx = ceil_div(cols / 4, kEwThreads);
  const long long others = (long long)c * b;
  const long long want = 8LL * num_sms();               // enough CTAs to fill the machine, no more column splits than that
  if ((long long)gx * others > want) {
    gx = (int)((want + others - 1) / others);
    if (gx < 1) gx = 1;
  }
  return gx;
}

}  // namespace
}  // namespace ws3d

using namespace ws3d;

WS3D_API int ws3d_bn_finalize(int c, double count, const double *stats, const float *gamma, const float *beta, float eps,
                              float momentum, float *running_mean, float *running_var, float *scale, float *shift, float *mean,
                              float *invstd, ws3d_stream_t stream) {
  if (c < 0 || !(count > 0.0)) return fail_arg("bn_finalize");
  if (c == 0) return 0;
  if (!stats || !scale || !shift || !mean || !invstd) return fail_arg("bn_finalize (null pointer)");
  bn_finalize_kernel<<<ceil_div(c, 128), 128, 0, to_stream(stream)>>>(c, count, stats, gamma, beta, eps, momentum, running_mean, running_var,
                                                                     scale, shift, mean, invstd);
  return check_launch("bn_finalize");
}

// y (B,c,cols) -> z (B,c,cols), or with pool > 0: z (B,c,cols/pool) and arg (B,c,cols/pool) uint8.  flags: 1 ReLU, 2 round z to TF32.
WS3D_API int ws3d_bn_relu_apply(int b, int c, int cols, int pool, const float *y, const float *scale, const float *shift, int flags,
                                float *z, unsigned char *arg, ws3d_stream_t stream) {
  const char *what = "bn_relu_apply";
  if (b < 0 || c < 0 || cols < 0 || cols % 4 || b > 65535 || c > 65535) return fail_arg(what);
  if (pool < 0 || (pool > 0 && (pool < 4 || pool > 128 || (pool & (pool - 1)) || cols % pool))) return fail_arg("bn_relu_apply (pool)");
  if (b == 0 || c == 0 || cols == 0) return 0;
  if (!y || !scale || !shift || !z || (pool > 0 && !arg)) return fail_arg("bn_relu_apply (null pointer)");
  const dim3 grid((unsigned)ew_col_blocks(cols, c, b), (unsigned)c, (unsigned)b);
  if (pool) bn_relu_pool_kernel<<<grid, kEwThreads, 0, to_stream(stream)>>>(c, cols, pool, flags & 1, y, scale, shift, z, arg);
  else bn_relu_apply_kernel<<<grid, kEwThreads, 0, to_stream(stream)>>>(c, cols, flags & 1, (flags >> 1) & 1, y, scale, shift, z);
  return check_launch(what);
}

// sums (2c doubles, zeroed by the caller) += [sum dA | sum dA * xhat]; dz is (B,c,cols) or, with pool > 0, (B,c,cols/pool) + arg.
WS3D_API int ws3d_bn_relu_bwd_reduce(int b, int c, int cols, int pool, const float *y, const float *dz, const unsigned char *arg,
                                     const float *scale, const float *shift, const float *mean, const float *invstd, int flags,
                                     double *sums, ws3d_stream_t stream) {
  const char *what = "bn_relu_bwd_reduce";
  if (b < 0 || c < 0 || cols < 0 || cols % 4 || c > 65535) return fail_arg(what);
  if (pool < 0 || (pool > 0 && (pool < 4 || pool > 128 || (pool & (pool - 1)) || cols % pool))) return fail_arg("bn_relu_bwd_reduce (pool)");
  if (b == 0 || c == 0 || cols == 0) return 0;
  if (!y || !dz || !scale || !shift || !mean || !invstd || !sums || (pool > 0 && !arg)) return fail_arg("bn_relu_bwd_reduce (null pointer)");
  const long long units = pool ? (long long)b * (cols / pool) : (long long)b * (cols / 4);
  long long splits = (4LL * num_sms() + c - 1) / c;
  const long long max_splits = (units + kEwThreads - 1) / kEwThreads;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const dim3 grid((unsigned)splits, (unsigned)c);
  if (pool) bn_relu_pool_bwd_reduce_kernel<<<grid, kEwThreads, 0, to_stream(stream)>>>(b, c, cols, pool, flags & 1, y, dz, arg, scale, shift, mean,
                                                                                      invstd, sums);
  else bn_relu_bwd_reduce_kernel<<<grid, kEwThreads, 0, to_stream(stream)>>>(b, c, cols, flags & 1, y, dz, scale, shift, mean, invstd, sums);
  return check_launch(what);
}

// dy (B,c,cols) = scale * (dA - sums[c] / count - xhat * sums[C + c] / count); count <= 0: no batch-norm coupling (dy = scale * dA).
WS3D_API int ws3d_bn_relu_bwd_apply(int b, int c, int cols, int pool, const float *y, const float *dz, const unsigned char *arg,
                                    const float *scale, const float *shift, const float *mean, const float *invstd, int flags,
                                    const double *sums, double count, float *dy, ws3d_stream_t stream) {
  const char *what = "bn_relu_bwd_apply";
  if (b < 0 || c < 0 || cols < 0 || cols % 4 || b > 65535 || c > 65535) return fail_arg(what);
  if (pool < 0 || (pool > 0 && (pool < 4 || pool > 128 || (pool & (pool - 1)) || cols % pool))) return fail_arg("bn_relu_bwd_apply (pool)");
  if (b == 0 || c == 0 || cols == 0) return 0;
  if (!y || !dz || !scale || !shift || !mean || !invstd || !sums || !dy || (pool > 0 && !arg)) return fail_arg("bn_relu_bwd_apply (null pointer)");
  const dim3 grid((unsigned)ew_col_blocks(cols, c, b), (unsigned)c, (unsigned)b);
  bn_relu_bwd_apply_kernel<<<grid, kEwThreads, 0, to_stream(stream)>>>(c, cols, pool, flags & 1, count > 0.0 ? (float)(1.0 / count) : 0.f, y, dz,
                                                                      arg, scale, shift, mean, invstd, sums, dy);
  return check_launch(what);
}

// dw (c_out x c_in, row stride ldw floats, zeroed or holding a running sum) += sum_b dy[b] (c_out x cols) * x[b]^T (cols x c_in)
WS3D_API int ws3d_mlp_wgrad(int b, int c_out, int c_in, int cols, const float *dy, const float *x, float *dw, int ldw,
                            ws3d_stream_t stream) {
  const char *what = "mlp_wgrad";
  if (b < 0 || c_out <= 0 || c_in <= 0 || cols < 0 || cols % 4 || ldw < c_in || b > 65535) return fail_arg(what);
  if (b == 0 || cols == 0) return 0;
  if (!dy || !x || !dw) return fail_arg("mlp_wgrad (null pointer)");
  CUtensorMap mdy, mx;
  if (!rows_map(&mdy, dy, b, c_out, cols, kWgTileM) || !rows_map(&mx, x, b, c_in, cols, kWgTileN)) return (int)cudaErrorInvalidValue;
  WgradParams prm;
  prm.c_out = c_out; prm.c_in = c_in; prm.ldw = ldw; prm.dw = dw;
  prm.n_m_tiles = ceil_div(c_out, kWgTileM);
  prm.n_n_tiles = ceil_div(c_in, kWgTileN);
  prm.chunks_per_cloud = ceil_div(cols, kWgChunk);
  const long long total = (long long)prm.chunks_per_cloud * b;
  if (total > 0x7FFFFFFFLL) return fail_arg("mlp_wgrad (too many chunks)");
  prm.total_chunks = (int)total;
  const int tiles = prm.n_m_tiles * prm.n_n_tiles;
  int splits = persistent_ctas(1) / tiles;
  if (splits < 1) splits = 1;
  if ((long long)splits > total) splits = (int)total;
  prm.chunks_per_cta = (int)((total + splits - 1) / splits);
  splits = (int)((total + prm.chunks_per_cta - 1) / prm.chunks_per_cta);   // no empty CTA
  const size_t smem = (size_t)kWgStages * (kWgTileM + kWgTileN) * kWgChunk * 4 + 1024;
  cudaError_t e = cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("mlp_wgrad: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  mlp_wgrad_kernel<<<dim3((unsigned)splits, (unsigned)tiles), kWgThreads, smem, to_stream(stream)>>>(mdy, mx, prm);
  return check_launch(what);
}
