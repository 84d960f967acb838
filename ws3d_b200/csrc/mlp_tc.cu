// Shared-MLP layer on the 5th-generation tensor cores (tcgen05, TF32 inputs, FP32 accumulate).
//
// The only dense contraction on the hot path: the 1x1-conv +
// BatchNorm(eval) + ReLU layers that follow every grouping (pointnet2_modules.py:40-44, :154;
// pytorch_utils.py:5-32) and, for the last layer of a set-abstraction scale, the max-pool over the
// nsample neighbours.  The reference runs them as cuDNN conv -> BN kernel -> ReLU kernel -> max-pool
// kernel, i.e. four full passes over the (B, C, npoint, nsample) tensor per layer.
//
//   Y[b, co, e] = relu( sum_ci W'[co, ci] * X[b, ci, e] + shift[co] )          (W' = BN-folded weight)
//   pooled:  Yp[b, co, j] = max_{s < nsample} Y[b, co, j*nsample + s]
//
// Orientation: output channels on the MMA M dimension (TMEM lanes), grouped points on N (TMEM
// columns).  That keeps the reference's channel-major (B, C, e) layout on both sides with no
// transposes -- W' is the K-major A operand, X is an MN-major B operand -- and makes the nsample
// max-pool a per-thread reduction over consecutive TMEM columns.
//
// Persistent CTAs; a tile = 128 output channels x NT columns of one cloud:
//   warp 0    TMA producer   cp.async.bulk.tensor (W: 128B swizzle; X: 128B swizzle with 32B atoms) of 32-deep K chunks, 4-stage ring
//   warp 1    MMA issuer     4 x tcgen05.mma.kind::tf32 (M=128, N=NT, K=8) per chunk, tcgen05.commit
//   warps 2-5 epilogue       tcgen05.ld 32x32b.x32 -> +shift, ReLU, (max over nsample) -> 128-byte stores
// A second input tensor can supply the tail of the K range (the skip features of a feature-propagation
// layer), so the torch.cat of pointnet2_modules.py:149 is never materialised.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"

namespace ws3d {
namespace {

constexpr int kChunkK = 32;        // fp32 elements per K chunk = one 128-byte swizzle row
constexpr int kTileM = 128;        // output channels per CTA (UMMA M)
constexpr int kThreads = 192;

struct MlpParams {
  int c_out;        // real output channels
  int cols;         // grouped points per cloud (npoint * nsample, or n)
  int nk1, nk2;     // K chunks taken from input 1 / input 2
  int pool;         // 0: write (B, c_out, cols); else nsample: write (B, c_out, cols / nsample)
  int flags;        // bit 0: ReLU, bit 1: round the stored output to TF32, bits 4-5: log2(copies of the weight rows)
  const float *shift;  // (c_out_padded)
  float *out;
  int out_ctot, out_coff;   // out is (B, out_ctot, cols or cols / pool); this layer writes channels [out_coff, out_coff + c_out)
  double *stats;            // training: [sum(c_out) | sum of squares(c_out)] of the RAW accumulator, added per tile; or null
  int n_col_tiles, n_m_tiles, n_tiles;   // tile = (cloud, 128-channel block, NT-column block), column block fastest
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
               : "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout).
//   layout 2 = SWIZZLE_128B (16-byte chunks XOR row%8): the K-major weight tile;
//   layout 1 = SWIZZLE_128B_BASE32B (32-byte chunks XOR row%4): the ONLY layout the tensor core accepts for an
//              MN-major 32-bit operand (measured with tools/umma_probe.cu: plain SWIZZLE_128B silently yields 0).
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw128Base32 = 1;
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFFu);              // start address, 16-byte units
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;   // leading-dimension byte offset
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;   // stride-dimension byte offset
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;                         // layout type
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
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

// Max-pool epilogue of one warp: 32 channels (one per lane) x the columns [cbeg, cend) of the accumulator.
// NS = nsample when it is 16 or 32 (the backbone's values: everything folds at compile time), NS = 0 takes any
// power of two `ns_rt` <= 128 that divides the warp's column share.  cols % ns == 0, so a started group is complete.
template <int NS>
__device__ __forceinline__ void pooled_chunks(uint32_t trow, int cbeg, int cend, int col0, int cols, bool live, float shift,
                                              bool relu, float *dst, int ns_rt = 0) {
  const int ns = NS ? NS : ns_rt;
  const int lg = __ffs(ns) - 1;
  float run = -INFINITY;
  for (int c = cbeg; c < cend && col0 + c < cols; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32(trow + (uint32_t)c, r);
    if (!live) continue;
    // The RAW accumulators are pooled; shift and ReLU touch the pooled value only: x -> fl(x + shift) and ReLU are monotone, so
    // max_t relu(x_t + shift) == relu(max_t x_t + shift) bit for bit (two to three instructions less per accumulator value).
    float v[32];
#pragma unroll
    for (int t = 0; t < 32; ++t) v[t] = __uint_as_float(r[t]);
#pragma unroll
    for (int w = 1; w < 32; w <<= 1) {
      if (w < ns) {
#pragma unroll
        for (int t = 0; t < 32; t += 2 * w) v[t] = fmaxf(v[t], v[t + w]);
      }
    }
    auto post = [&](float x) { x += shift; return relu ? fmaxf(x, 0.f) : x; };
    if (NS == 32) {
      dst[c >> 5] = post(v[0]);
    } else if (NS == 16) {
      float *o = dst + (c >> 4);
      const bool second = col0 + c + 16 < cols;   // the row may end in the middle of this 32-column chunk
      if (second && (reinterpret_cast<uintptr_t>(o) & 7u) == 0) *reinterpret_cast<float2 *>(o) = make_float2(post(v[0]), post(v[16]));
      else { o[0] = post(v[0]); if (second) o[1] = post(v[16]); }
    } else if (ns > 32) {   // a group spans several 32-column chunks (cbeg % ns == 0: the share starts on a group)
      run = fmaxf(run, v[0]);
      if (((c + 32) & (ns - 1)) == 0) {
        dst[(c + 32 - ns) >> lg] = post(run);
        run = -INFINITY;
      }
    } else {
#pragma unroll
      for (int g = 0; g < 32; ++g)
        if ((g & (ns - 1)) == 0 && col0 + c + g < cols) dst[(c + g) >> lg] = post(v[g]);
    }
  }
}

template <int NT, int kStages, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) mlp_layer_kernel(const __grid_constant__ CUtensorMap map_w,
                                                                const __grid_constant__ CUtensorMap map_x1,
                                                                const __grid_constant__ CUtensorMap map_x2,
                                                                const MlpParams prm) {
  constexpr uint32_t kABytes = kTileM * kChunkK * 4;           // 16 KB
  constexpr uint32_t kBBytes = NT * kChunkK * 4;               // NT columns x 32 k
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = 2 * NT;                       // two accumulators: the epilogue of tile j overlaps the MMAs of tile j+1
  // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, A K-major, B MN-major
  constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) |
                              ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

  extern __shared__ uint8_t s_raw[];
  __shared__ __align__(8) unsigned long long s_full[kStages], s_empty[kStages], s_tmem_full[2], s_tmem_empty[2];
  __shared__ uint32_t s_tmem_base;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t stage_base = (smem_u32(s_raw) + 1023u) & ~1023u;  // 128B-swizzle atoms need 1 KB alignment
  const int nk = prm.nk1 + prm.nk2;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_empty[s]), 1);
    }
    for (int q = 0; q < 2; ++q) {
      mbar_init(smem_u32(&s_tmem_full[q]), 1);
      mbar_init(smem_u32(&s_tmem_empty[q]), 4);   // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x1) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem_base;

  // Persistent CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ...  The three roles walk the same tile list and
  // are coupled only through the mbarrier rings (smem stages: full/empty; accumulators: tmem_full/tmem_empty).
  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer
      int it = 0;
      for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
        const int ct = tile % prm.n_col_tiles, rest = tile / prm.n_col_tiles;
        const int col0 = ct * NT, m0 = (rest % prm.n_m_tiles) * kTileM, cloud = rest / prm.n_m_tiles;
        for (int i = 0; i < nk; ++i, ++it) {
          const int s = it % kStages;
          if (it >= kStages) mbar_wait(smem_u32(&s_empty[s]), (uint32_t)((it / kStages) - 1) & 1u);
          const uint32_t bar = smem_u32(&s_full[s]);
          const uint32_t a_dst = stage_base + (uint32_t)s * kStageBytes, b_dst = a_dst + kABytes;
          mbar_expect_tx(bar, kStageBytes);
          tma_load_2d(a_dst, &map_w, i * kChunkK, m0, bar);
          const bool second = i >= prm.nk1;
          const CUtensorMap *mx = second ? &map_x2 : &map_x1;
          const int k0 = (second ? i - prm.nk1 : i) * kChunkK;
#pragma unroll
          for (int nb = 0; nb < NT / 32; ++nb) tma_load_3d(b_dst + (uint32_t)nb * 4096u, mx, col0 + nb * 32, k0, cloud, bar);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- MMA issuer
      int it = 0, j = 0;
      for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x, ++j) {
        const int buf = j & 1;
        if (j >= 2) {  // the epilogue must have drained this accumulator (tile j-2)
          mbar_wait(smem_u32(&s_tmem_empty[buf]), (uint32_t)((j >> 1) - 1) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t acc = tmem_base + (uint32_t)(buf * NT);
        for (int i = 0; i < nk; ++i, ++it) {
          const int s = it % kStages;
          mbar_wait(smem_u32(&s_full[s]), (uint32_t)(it / kStages) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_base = stage_base + (uint32_t)s * kStageBytes, b_base = a_base + kABytes;
#pragma unroll
          for (int kk = 0; kk < kChunkK / 8; ++kk) {
            // A: K-major, rows 128 B apart, 8-row groups 1 KB apart; advance 32 B per K=8 step
            const uint64_t da = smem_desc(a_base + (uint32_t)kk * 32u, 16u, 1024u, kLayoutSw128);
            // B: MN-major, 32-column blocks 4 KB apart (LBO), 4-deep K atoms 512 B apart (SBO), 8 k-rows per MMA
            const uint64_t db = smem_desc(b_base + (uint32_t)kk * 1024u, 4096u, 512u, kLayoutSw128Base32);
            umma_tf32(acc, da, db, kIdesc, (i | kk) != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&s_empty[s]));  // frees the stage once these MMAs have read it
        }
        umma_commit(smem_u32(&s_tmem_full[buf]));   // accumulator complete
      }
    }
  } else {
   int j = 0;
   for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x, ++j) {
    const int buf = j & 1;
    const int ct = tile % prm.n_col_tiles, rest = tile / prm.n_col_tiles;
    const int col0 = ct * NT, m0 = (rest % prm.n_m_tiles) * kTileM, cloud = rest / prm.n_m_tiles;
    // ---- epilogue: warp w may only touch TMEM lanes 32*(w%4) .. +31
    // With few output channels the weight rows are REPLICATED over the 128 MMA rows (rep = 2 or 4 copies, made
    // by the host): every TMEM quarter then holds real channels and the four epilogue warps split the columns,
    // instead of one warp draining the whole tile.
    const int quarter = warp & 3;
    const int rep_log2 = (prm.flags >> 4) & 3;
    const int rows_per_copy = kTileM >> rep_log2;                 // 128, 64 or 32
    const int co = m0 + ((quarter * 32 + lane) & (rows_per_copy - 1));
    const int part = (quarter * 32) / rows_per_copy;              // which column share this warp drains
    const int cbeg = part * (NT >> rep_log2), cend = cbeg + (NT >> rep_log2);
    mbar_wait(smem_u32(&s_tmem_full[buf]), (uint32_t)(j >> 1) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const bool live = co < prm.c_out;
    const float shift = live ? __ldg(prm.shift + co) : 0.f;
    const bool relu = (prm.flags & 1) != 0, round_out = (prm.flags & 2) != 0;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * NT);
    const bool warp_live = m0 + ((quarter * 32) & (rows_per_copy - 1)) < prm.c_out;  // a quarter with no real channel has nothing to read
    if (warp_live && prm.pool == 0) {
      // Each lane holds one channel row (32 consecutive columns per TMEM load).  Storing that directly would touch
      // 32 different cache lines with 16 bytes each per instruction; the chunk is transposed through shared memory
      // (a 4.5 KB staging area per epilogue warp behind the operand ring) so that every store instruction writes
      // four full 128-byte lines.
      float *stage = reinterpret_cast<float *>(s_raw + (stage_base - smem_u32(s_raw)) + kStages * kStageBytes) + quarter * (32 * 36);
      const int co_base = m0 + ((quarter * 32) & (rows_per_copy - 1));
      const int srow = lane >> 3, scol = (lane & 7) * 4;
      const uint32_t rnd_add = round_out ? 0x1000u : 0u, rnd_mask = round_out ? 0xFFFFE000u : 0xFFFFFFFFu;
      const int rows_live = prm.c_out - co_base;       // channel rows of this warp that exist (may be <= 0 or > 32)
      float *out_row0 = prm.out + ((size_t)cloud * prm.out_ctot + prm.out_coff + co_base + srow) * prm.cols + col0 + scol;
      float st_sum = 0.f, st_sq = 0.f;
      for (int c = cbeg; c < cend && col0 + c < prm.cols; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(trow + (uint32_t)c, r);
        if (prm.stats) {   // columns beyond the row are zero-filled by TMA: they add nothing
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            const float a = __uint_as_float(r[t]);
            st_sum += a;
            st_sq = fmaf(a, a, st_sq);
          }
        }
        // + shift (packed FP32 adds), ReLU and rounding to the nearest TF32 (so that the next layer's tensor-core truncation is
        // exact) as ONE fused integer add-max per value: max(bits + h, h) = bits(max(v, 0)) + h for every non-NaN v, h = half a
        // TF32 ulp or 0 (negative floats have negative bit patterns); the flags are warp-uniform and tested once per chunk
        // (ncu: the per-value flag tests and the two-instruction ReLU + add were a third of this kernel's instructions)
        const float2 sh2 = make_float2(shift, shift);
        if (relu) {
#pragma unroll
          for (int t = 0; t < 32; t += 2) {
            const float2 v = __fadd2_rn(make_float2(__uint_as_float(r[t]), __uint_as_float(r[t + 1])), sh2);
            r[t] = (uint32_t)__viaddmax_s32(__float_as_int(v.x), (int)rnd_add, (int)rnd_add) & rnd_mask;
            r[t + 1] = (uint32_t)__viaddmax_s32(__float_as_int(v.y), (int)rnd_add, (int)rnd_add) & rnd_mask;
          }
        } else {
#pragma unroll
          for (int t = 0; t < 32; t += 2) {
            const float2 v = __fadd2_rn(make_float2(__uint_as_float(r[t]), __uint_as_float(r[t + 1])), sh2);
            r[t] = (__float_as_uint(v.x) + rnd_add) & rnd_mask;
            r[t + 1] = (__float_as_uint(v.y) + rnd_add) & rnd_mask;
          }
        }
#pragma unroll
        for (int t = 0; t < 32; t += 4)
          *reinterpret_cast<uint4 *>(stage + lane * 36 + t) = make_uint4(r[t], r[t + 1], r[t + 2], r[t + 3]);
        __syncwarp();
        if (col0 + c + scol < prm.cols) {
          float *dst = out_row0 + c;                    // row srow of this warp's 32 channels, this lane's four columns
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int row = k * 4 + srow;
            const uint4 val = *reinterpret_cast<const uint4 *>(stage + row * 36 + scol);
            if (row < rows_live) __stcs(reinterpret_cast<uint4 *>(dst + (size_t)(k * 4) * prm.cols), val);
          }
        }
        __syncwarp();
      }
      if (prm.stats && live) {
        atomicAdd(prm.stats + co, (double)st_sum);
        atomicAdd(prm.stats + prm.c_out + co, (double)st_sq);
      }
    } else if (warp_live) {
      const int ns = prm.pool;                  // power of two dividing the warp's column share
      const int lg = __ffs(ns) - 1;
      float *dst = prm.out + ((size_t)cloud * prm.out_ctot + prm.out_coff + co) * (prm.cols >> lg) + (col0 >> lg);
      if (ns == 16)
        pooled_chunks<16>(trow, cbeg, cend, col0, prm.cols, live, shift, relu, dst);
      else if (ns == 32)
        pooled_chunks<32>(trow, cbeg, cend, col0, prm.cols, live, shift, relu, dst);
      else
        pooled_chunks<0>(trow, cbeg, cend, col0, prm.cols, live, shift, relu, dst, ns);
    }
      // this accumulator may be overwritten by the MMAs of tile j+2
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_tmem_empty[buf])) : "memory");
   }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

bool make_map(CUtensorMap *m, const void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides_bytes,
              const cuuint32_t *box, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("mlp_layer: cuTensorMapEncodeTiled unavailable"); return false; }
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void *>(base), dims, strides_bytes, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("mlp_layer: cuTensorMapEncodeTiled failed (%d)", (int)r); return false; }
  return true;
}

}  // namespace
}  // namespace ws3d

using namespace ws3d;

// w: (c_out_pad, k_pad) row-major device matrix, c_out_pad % 128 == 0, k_pad = 32*(nk1+nk2): BN-folded weights,
//    zero padded; columns [0, 32*nk1) multiply x1's channels, the rest x2's.
// x1: (B, c1, cols), x2: (B, c2, cols) or NULL.  shift: (c_out_pad).  cols % 4 == 0.
// out: (B, c_out, cols) or, when pool > 0, (B, c_out, cols / pool) with the max over each run of `pool` columns.
// out_ctot / out_coff: the layer writes channels [out_coff, out_coff + c_out) of an (B, out_ctot, .) tensor (the slot of one
//    scale in the concatenated multi-scale output of pointnet2_modules.py:55: no torch.cat pass).
static int mlp_layer_impl(int b, int c_out, int c_out_pad, int c1, int c2, int cols, const float *w, const float *shift,
                          const float *x1, const float *x2, float *out, int out_ctot, int out_coff, int relu, int pool,
                          double *stats, ws3d_stream_t stream) {
  const char *what = "mlp_layer";
  if (stats && pool != 0) return fail_arg("mlp_layer (statistics are taken on unpooled layers)");
  if (out_coff < 0 || out_coff + c_out > out_ctot) return fail_arg("mlp_layer (output channel slot)");
  if (b < 0 || c_out <= 0 || c1 <= 0 || c2 < 0 || cols < 0 || c_out_pad % kTileM || c_out_pad < c_out) return fail_arg(what);
  if (b == 0 || cols == 0) return 0;
  if (!w || !shift || !x1 || !out || (c2 > 0 && !x2)) return fail_arg(what);
  if (cols % 4 || b > 65535) return fail_arg("mlp_layer (cols % 4 != 0 or batch > 65535)");
  {
    const int rep_log2 = (relu >> 4) & 3;
    if (rep_log2 > 2 || (rep_log2 > 0 && (c_out_pad != kTileM || c_out > (kTileM >> rep_log2))))
      return fail_arg("mlp_layer (row replication needs c_out_pad == 128 and c_out <= 128 / copies)");
    if (pool > 128 || (pool > 0 && ((128 >> rep_log2) % pool) != 0)) return fail_arg("mlp_layer (pool must be <= 128 and divide the per-warp column share)");
  }
  // tile configuration: (columns per tile, smem stages, CTAs per SM).  Several small CTAs per SM overlap one
  // CTA's TMA / MMA phase with another's epilogue; WS3D_MLP_CFG overrides for experiments.
  static const int cfg_env = []() { const char *e = getenv("WS3D_MLP_CFG"); return (e && *e) ? atoi(e) : -1; }();
  // measured over the backbone's layers (profiles/r1_mlp_bench_v4.json): 128-column tiles with two CTAs per SM win
  // when the K loop is short (<= 64 input channels) or the layer pools over 16 columns; 256-column tiles otherwise
  int cfg = cfg_env >= 0 ? cfg_env : ((c1 + c2 <= 64 || (pool > 0 && pool <= 16)) ? 3 : 0);
  const int max_nt = cfg <= 1 ? 256 : 128;
  if (pool < 0 || (pool > 0 && (cols % pool != 0 || max_nt % pool != 0 || (pool & (pool - 1)) != 0)))
    return fail_arg("mlp_layer (pool must be a power of two dividing the column tile and cols)");
  MlpParams prm;
  prm.c_out = c_out; prm.cols = cols; prm.pool = pool; prm.flags = relu; prm.shift = shift; prm.out = out;
  prm.out_ctot = out_ctot; prm.out_coff = out_coff;
  prm.stats = stats;
  prm.nk1 = ceil_div(c1, kChunkK);
  prm.nk2 = c2 > 0 ? ceil_div(c2, kChunkK) : 0;
  prm.n_m_tiles = c_out_pad / kTileM;
  const int k_pad = (prm.nk1 + prm.nk2) * kChunkK;
  CUtensorMap mw, m1, m2;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)k_pad, (cuuint64_t)c_out_pad};
    const cuuint64_t strides[1] = {(cuuint64_t)k_pad * 4};
    const cuuint32_t box[2] = {kChunkK, kTileM};
    if (!make_map(&mw, w, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return (int)cudaErrorInvalidValue;
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)c1, (cuuint64_t)b};
    const cuuint64_t strides[2] = {(cuuint64_t)cols * 4, (cuuint64_t)cols * c1 * 4};
    const cuuint32_t box[3] = {32, kChunkK, 1};
    if (!make_map(&m1, x1, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return (int)cudaErrorInvalidValue;
  }
  if (c2 > 0) {
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)c2, (cuuint64_t)b};
    const cuuint64_t strides[2] = {(cuuint64_t)cols * 4, (cuuint64_t)cols * c2 * 4};
    const cuuint32_t box[3] = {32, kChunkK, 1};
    if (!make_map(&m2, x2, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return (int)cudaErrorInvalidValue;
  } else {
    m2 = m1;
  }
  cudaError_t e = cudaSuccess;
#define WS3D_MLP_LAUNCH(NT_, ST_, MB_)                                                                              \
  {                                                                                                                 \
    const size_t smem = (size_t)(ST_) * (kTileM * kChunkK * 4 + (NT_) * kChunkK * 4) + 4 * 32 * 36 * 4 + 1024;      \
    auto kern = mlp_layer_kernel<NT_, ST_, MB_>;                                                                    \
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                         \
    if (e != cudaSuccess) { set_error("mlp_layer: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }     \
    prm.n_col_tiles = ceil_div(cols, NT_);                                                                          \
    const long long tiles = (long long)prm.n_col_tiles * prm.n_m_tiles * b;                                         \
    if (tiles > 0x7FFFFFFFLL) return fail_arg("mlp_layer (too many tiles)");                                        \
    prm.n_tiles = (int)tiles;                                                                                       \
    const int pc = persistent_ctas(MB_);                                                                            \
    const int ctas = (int)(tiles < (long long)pc ? tiles : (long long)pc);                                                          \
    kern<<<ctas, kThreads, smem, to_stream(stream)>>>(mw, m1, m2, prm);                                             \
  }
  switch (cfg) {
    case 0: WS3D_MLP_LAUNCH(256, 4, 1) break;   // one persistent CTA per SM, 256-column tiles (2 x 256 TMEM columns), deep ring
    case 1: WS3D_MLP_LAUNCH(256, 3, 1) break;
    case 2: WS3D_MLP_LAUNCH(128, 3, 2) break;   // two persistent CTAs per SM, 128-column tiles (2 x 128 TMEM columns each)
    default: WS3D_MLP_LAUNCH(128, 2, 2) break;
  }
#undef WS3D_MLP_LAUNCH
  return check_launch(what);
}

WS3D_API int ws3d_mlp_layer_into(int b, int c_out, int c_out_pad, int c1, int c2, int cols, const float *w, const float *shift,
                                 const float *x1, const float *x2, float *out, int out_ctot, int out_coff, int relu, int pool,
                                 ws3d_stream_t stream) {
  return mlp_layer_impl(b, c_out, c_out_pad, c1, c2, cols, w, shift, x1, x2, out, out_ctot, out_coff, relu, pool, nullptr, stream);
}

WS3D_API int ws3d_mlp_layer(int b, int c_out, int c_out_pad, int c1, int c2, int cols, const float *w, const float *shift,
                            const float *x1, const float *x2, float *out, int relu, int pool, ws3d_stream_t stream) {
  return mlp_layer_impl(b, c_out, c_out_pad, c1, c2, cols, w, shift, x1, x2, out, c_out, 0, relu, pool, nullptr, stream);
}

// Training forward: y (B, c_out, cols) = W [x1 ; x2] (raw accumulator: no shift, no activation) and the per-channel
// sum / sum of squares of y added into stats (2 * c_out doubles, zeroed by the caller) from the epilogue.
WS3D_API int ws3d_mlp_layer_stats(int b, int c_out, int c_out_pad, int c1, int c2, int cols, const float *w, const float *zero_shift,
                                  const float *x1, const float *x2, float *y, double *stats, ws3d_stream_t stream) {
  if (!stats) return fail_arg("mlp_layer_stats (null pointer)");
  return mlp_layer_impl(b, c_out, c_out_pad, c1, c2, cols, w, zero_shift, x1, x2, y, c_out, 0, 0, 0, stats, stream);
}
