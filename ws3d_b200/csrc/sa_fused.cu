// One set-abstraction scale in ONE kernel:
//
//     ball-query indices -> gather [xyz - centre ; features] -> (conv1x1 + BN + ReLU) x 3 -> max over nsample
//
// i.e. QueryAndGroup's grouping (pointnet2_utils.py:241-264), the three SharedMLP layers (pytorch_utils.py:5-32) and
// the max-pool of pointnet2_modules.py:40-44, without ever materialising the (B, 3+C, npoint, nsample) grouped tensor
// or the two intermediate activations in HBM.  The per-layer path (mlp_tc.cu) moves them through HBM five times;
// at the Stage-1 shapes that is 80-95 % of the traffic of SA1 and SA2.
//
// Orientation (transposed with respect to mlp_tc.cu): the 128 grouped points of a tile sit on the MMA M dimension
// = the 128 TMEM lanes, channels run along TMEM columns.
//   * The gathered input tile is written straight into TMEM (tcgen05.st) by the thread that owns the row: thread r
//     fetches idx, the centre, xyz[idx] - centre and features[:, idx] and stores them along lane r.
//   * Every layer is  D[128 x N] = A[128 x K] * W^T  with A READ FROM TMEM (tcgen05.mma, A in tensor memory) and
//     W (N x K, K-major, BN folded, TF32-rounded) resident in shared memory for the whole kernel (TMA, 128B swizzle).
//   * The accumulator of layer l has exactly the layout layer l+1 wants for its A operand, so the epilogue is
//     tcgen05.ld -> + shift, ReLU, round to TF32 -> tcgen05.st in place.  Activations never leave TMEM.
//   * The last epilogue max-pools over nsample consecutive rows = lanes: the values are >= 0 after the ReLU, so their
//     bit patterns order like unsigned integers and one redux.sync.max.u32 per channel pools a whole warp.
// TMEM columns are reused: layer 2's accumulator overlays the (dead) input tile, layer 3's overlays layer 1's.
//
// 128 threads: warp w owns TMEM lanes 32w..32w+31 (gather + the three epilogues); after a CTA barrier thread 0 issues
// the layer's MMAs and everybody waits for the tcgen05.commit.  Persistent CTAs, up to four per SM (TMEM columns,
// shared memory), so one CTA's gather latency and MMA hand-offs are covered by the others' epilogues.
#include <cuda.h>

#include "common.cuh"
#include "tma.cuh"

namespace ws3d {
namespace {

constexpr int kRows = 128;          // grouped points per tile (UMMA M)
constexpr int kThreads = 128;      // per half: a CTA has kThreads x kHalves threads (see the kernel)

struct SaFusedParams {
  int n, m, ns, c_feat;          // points per cloud, centres per cloud, nsample, feature channels
  int k0;                        // 3 + c_feat rounded up to 8: K of layer 1
  int n1, n2, n3;                // layer widths rounded up to 16 (UMMA N; K of the next layer)
  int c3;                        // real output channels
  int nk1, nk2, nk3;             // 32-deep K chunks of the resident weights
  int tm_a0, tm_r1, tm_r2, tm_r3, tmem_cols;
  int tiles_per_cloud, n_tiles;
  const float *xyz, *new_xyz, *feat;
  const int *idx;
  const float *shift1, *shift2, *shift3;   // padded to n1 / n2 / n3
  float *out;
  int out_ctot, out_coff;        // out is (B, out_ctot, m); this scale writes channels [out_coff, out_coff + c3)
  // Point-major operand rows (optional): rows (B, n, ld) = [x, y, z, features..., zero padding], ld % 4 == 0, so that a
  // thread gathers its grouped point with ld / 4 16-byte loads of ONE contiguous row instead of 3 + c_feat 4-byte loads
  // that each touch their own 32-byte sector of a channel-major tensor (at SA2 the channel-major gather moves 1.7 GB
  // of sectors through L2 for 0.2 GB of data).  out_pm (B, m, ld_pm) receives this scale's pooled channels in the same
  // layout for the NEXT level; the scale with pm_xyz set also writes the centre coordinates and the zero padding.
  const float *rows;
  int ld;
  float *out_pm;
  int ld_pm, pm_xyz;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}

// K-major, 128-byte-swizzled operand tile: rows 128 B apart, 8-row groups 1 KB apart (cute::UMMA::SmemDescriptor).
__device__ __forceinline__ uint64_t smem_desc_k128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)(16u >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

// D[tmem] (+)= A[tmem] * B[smem]: TF32 inputs, FP32 accumulate, M = 128
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t round_tf32(float v) { return (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u; }

// Epilogue of layers 1 and 2 for four accumulator values: + shift, ReLU, round to the nearest TF32 value (the tensor core ignores
// the low 13 bits, so adding half a TF32 ulp to the bit pattern is the rounding).  Two packed FP32 adds (add.rn.f32x2) and one
// fused integer add-and-max per value: max(bits + 0x1000, 0x1000) equals bits(max(v, 0)) + 0x1000 for every non-NaN v (negative
// values have negative bit patterns); a NaN stays a NaN instead of becoming 0, as in the reference's ReLU.  Six instructions for
// the twelve of FADD + FMNMX + IADD per value: the two epilogues were 20-30 % of this kernel's instructions.
__device__ __forceinline__ void shift_relu_round4(uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d, const float4 s4) {
  const float2 p = __fadd2_rn(make_float2(__uint_as_float(a), __uint_as_float(b)), make_float2(s4.x, s4.y));
  const float2 q = __fadd2_rn(make_float2(__uint_as_float(c), __uint_as_float(d)), make_float2(s4.z, s4.w));
  a = (uint32_t)__viaddmax_s32(__float_as_int(p.x), 0x1000, 0x1000);
  b = (uint32_t)__viaddmax_s32(__float_as_int(p.y), 0x1000, 0x1000);
  c = (uint32_t)__viaddmax_s32(__float_as_int(q.x), 0x1000, 0x1000);
  d = (uint32_t)__viaddmax_s32(__float_as_int(q.y), 0x1000, 0x1000);
}

// Butterfly max-reduction over the G lanes of a pooling group: every lane enters with G channel values of its own
// row and leaves with ONE channel (index = lane % G) maximised over the G rows -- a reduce-scatter, G - 1 shuffles
// for G channels instead of G full-warp reductions.
template <int G>
__device__ __forceinline__ float pool_scatter(float (&v)[G], int lane) {
#pragma unroll
  for (int h = G / 2; h >= 1; h >>= 1) {
    const bool upper = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = upper ? v[i] : v[i + h];
      const float keep = upper ? v[i + h] : v[i];
      v[i] = fmaxf(keep, __shfl_xor_sync(0xFFFFFFFFu, send, h));
    }
  }
  return v[0];
}

// The same over the EIGHT lanes that share lane % 4 (they differ in lane bits 2..4): eight values in, one out; the lane with
// q = lane / 4 leaves with value index q.
__device__ __forceinline__ float pool8_q(float (&v)[8], int lane) {
#pragma unroll
  for (int h = 4; h >= 1; h >>= 1) {
    const bool upper = (lane & (h << 2)) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = upper ? v[i] : v[i + h];
      const float keep = upper ? v[i + h] : v[i];
      v[i] = fmaxf(keep, __shfl_xor_sync(0xFFFFFFFFu, send, h << 2));
    }
  }
  return v[0];
}

// kBatch feature channels are fetched per thread before they are stored to TMEM (independent loads in flight);
// the wide variant (64) is for the scales whose TMEM / shared-memory footprint allows two CTAs per SM anyway.
// kHalves = 2 / 4 (256 / 512 threads) is for the shapes whose resident weights leave ONE CTA per SM (the Stage-2 widths: 208 KB):
// with four warps every scheduler of the SM would own a single warp and idle through each of its stalls (ncu: 6 % of the warp
// slots active, 18 % issue-active).  Warps w, w + 4, ... share a TMEM lane quarter and split every per-row phase by columns -- the
// row's gather, the two in-place epilogues, the pooling groups, the output loops -- so each phase has kHalves times the warps to
// issue from and to hide latency behind; barriers and the MMA issue are unchanged.  Measured on the Stage-2 stack (512 proposals):
// 2.84 ms with 128 threads, 2.52 with 256, 2.42 with 512.  Tried on top and not kept: issuing every layer as two halves of its output
// columns with separate commits, so that the epilogue of the first half overlaps the MMAs of the second -- 2.99 ms: the N = 64
// instructions take as long as the N = 128 ones, so the tensor-core time doubles and eats the overlap.
template <int kBatch, int kMinCtas, int kHalves>
__global__ void __launch_bounds__(kThreads * kHalves, kMinCtas) sa_mlp_fused_kernel(const __grid_constant__ CUtensorMap map_w1,
                                                                   const __grid_constant__ CUtensorMap map_w2,
                                                                   const __grid_constant__ CUtensorMap map_w3,
                                                                   const SaFusedParams prm) {
  extern __shared__ uint8_t s_raw[];
  __shared__ __align__(8) unsigned long long s_bar_w, s_bar_d;
  __shared__ uint32_t s_tmem_base;

  constexpr int kT = kThreads * kHalves;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = kHalves > 1 ? (int)(threadIdx.x >> 7) : 0;       // which of the kHalves column shares this thread takes
  const uint32_t w_base = (smem_u32(s_raw) + 1023u) & ~1023u;      // 128B-swizzle atoms need 1 KB alignment
  const uint32_t w1 = w_base, w2 = w1 + (uint32_t)prm.nk1 * prm.n1 * 128u, w3 = w2 + (uint32_t)prm.nk2 * prm.n2 * 128u;
  const uint32_t w_end = w3 + (uint32_t)prm.nk3 * prm.n3 * 128u;
  // behind the weights: the three shift vectors (n1 + n2 + n3 floats), then the pooled staging (c3 x centres per tile)
  float *s_shift = reinterpret_cast<float *>(s_raw + (w_end - smem_u32(s_raw)));
  unsigned int *s_pool = reinterpret_cast<unsigned int *>(s_shift + prm.n1 + prm.n2 + prm.n3);
  const uint32_t bar_w = smem_u32(&s_bar_w), bar_d = smem_u32(&s_bar_d);
  const int lg_ns = __ffs(prm.ns) - 1;
  const int cpt = kRows >> lg_ns;                     // centres per tile (>= 1)

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_d, 1);    // tcgen05.commit
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < prm.n1; i += kT) s_shift[i] = __ldg(prm.shift1 + i);
  for (int i = threadIdx.x; i < prm.n2; i += kT) s_shift[prm.n1 + i] = __ldg(prm.shift2 + i);
  for (int i = threadIdx.x; i < prm.n3; i += kT) s_shift[prm.n1 + prm.n2 + i] = __ldg(prm.shift3 + i);
  for (int i = threadIdx.x; i < prm.c3 * cpt; i += kT) s_pool[i] = 0u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(prm.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem_base;

  if (threadIdx.x == 0) {
    // ---- resident weights: every 32-deep K chunk of the three layers, once per CTA
    const uint32_t bytes = ((uint32_t)prm.nk1 * prm.n1 + (uint32_t)prm.nk2 * prm.n2 + (uint32_t)prm.nk3 * prm.n3) * 128u;
    mbar_expect_tx(bar_w, bytes);
    for (int i = 0; i < prm.nk1; ++i) tma_load_2d(w1 + (uint32_t)i * prm.n1 * 128u, &map_w1, i * 32, 0, bar_w);
    for (int i = 0; i < prm.nk2; ++i) tma_load_2d(w2 + (uint32_t)i * prm.n2 * 128u, &map_w2, i * 32, 0, bar_w);
    for (int i = 0; i < prm.nk3; ++i) tma_load_2d(w3 + (uint32_t)i * prm.n3 * 128u, &map_w3, i * 32, 0, bar_w);
    mbar_wait(bar_w, 0);
  }

  // One layer's MMAs, issued by thread 0 once all 128 threads have written their rows of the A operand:
  // D[128 x N] = A[128 x K] (TMEM) * W^T (shared memory), K = 8 per instruction.
  auto issue_layer = [&](uint32_t wl, int nl, int kl, uint32_t a_col, uint32_t d_col) {
    tmem_st_wait();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = nl
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kRows >> 4) << 24) | ((uint32_t)(nl >> 3) << 17);
      const int ksteps = kl >> 3;
      for (int ks = 0; ks < ksteps; ++ks) {
        // weights: chunk ks/4 (N rows x 128 B), 32 B further per K = 8 step inside the swizzled row
        const uint64_t db = smem_desc_k128(wl + (uint32_t)(ks >> 2) * (uint32_t)nl * 128u + (uint32_t)(ks & 3) * 32u);
        umma_tf32_ts(tmem_base + d_col, tmem_base + a_col + (uint32_t)ks * 8u, db, idesc, ks != 0 ? 1u : 0u);
      }
      umma_commit(bar_d);
    }
  };

  // ---- thread = one grouped point (TMEM lane)
  const int row = threadIdx.x & (kRows - 1);          // 0..127 (both halves)
  const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const int ns = prm.ns;
  const int cols = prm.m * ns;                        // grouped points per cloud
  const int n = prm.n, c_feat = prm.c_feat;
  uint32_t phase_d = 0;
  // the neighbour index of this thread's row is fetched one tile ahead (it heads the gather's dependency chain)
  // (cloud, tile inside the cloud) of the current tile and of the one whose index is being prefetched are carried along
  // instead of divided out per tile: the two integer divisions were 7 % of the SA1 launches' stall samples
  auto load_idx = [&](int cl, int ti) -> int {
    const int fl = ti * kRows + row;
    return fl < cols ? __ldg(prm.idx + (size_t)cl * cols + fl) : 0;
  };
  const int tpc = prm.tiles_per_cloud;
  const int step_cl = (int)gridDim.x / tpc, step_ti = (int)gridDim.x - step_cl * tpc;     // one stride of the tile loop
  int cloud = (int)blockIdx.x / tpc, t_in = (int)blockIdx.x - cloud * tpc;
  int p_next = blockIdx.x < prm.n_tiles ? load_idx(cloud, t_in) : 0;
  for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
    const int flat = t_in * kRows + row;
    const bool valid = flat < cols;
    const int p = p_next;
    int cloud_n = cloud + step_cl, t_in_n = t_in + step_ti;
    if (t_in_n >= tpc) { t_in_n -= tpc; ++cloud_n; }
    if (tile + (int)gridDim.x < prm.n_tiles) p_next = load_idx(cloud_n, t_in_n);
    // ---- gather: [xyz[idx] - centre ; features[:, idx]] along this thread's TMEM lane.  Rows beyond the cloud
    // (last tile only) recompute point 0 against centre 0: finite values that never reach an output.
    if (prm.rows) {
      // point-major source: the whole operand row is one contiguous run of ld floats
      const int j = valid ? flat >> lg_ns : 0;
      const float4 *rp = reinterpret_cast<const float4 *>(prm.rows + ((size_t)cloud * n + p) * prm.ld);
      const float *pc = prm.new_xyz + ((size_t)cloud * prm.m + j) * 3;
      const float cx = __ldg(pc), cy = __ldg(pc + 1), cz = __ldg(pc + 2);
      for (int c0 = half * kBatch; c0 < prm.k0; c0 += kBatch * kHalves) {     // the halves take alternate column batches
        uint32_t v[kBatch];
#pragma unroll
        for (int t = 0; t < kBatch; t += 4) {
          if (c0 + t < prm.ld) {                      // CTA-uniform; columns beyond the row are zero (k0 may exceed ld)
            const float4 q = __ldg(rp + ((c0 + t) >> 2));
            v[t] = __float_as_uint(q.x); v[t + 1] = __float_as_uint(q.y); v[t + 2] = __float_as_uint(q.z); v[t + 3] = __float_as_uint(q.w);
          } else {
            v[t] = v[t + 1] = v[t + 2] = v[t + 3] = 0u;
          }
        }
        if (c0 == 0) {
          v[0] = round_tf32(__fsub_rn(__uint_as_float(v[0]), cx));
          v[1] = round_tf32(__fsub_rn(__uint_as_float(v[1]), cy));
          v[2] = round_tf32(__fsub_rn(__uint_as_float(v[2]), cz));
        }
#pragma unroll
        for (int g = 0; g < kBatch / 8; ++g) {
          if (c0 + g * 8 < prm.k0) {                  // CTA-uniform
            uint32_t q[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) q[t] = v[g * 8 + t];
            tmem_st8(lane_addr + (uint32_t)(prm.tm_a0 + c0 + g * 8), q);
          }
        }
      }
    } else {
      const int j = valid ? flat >> lg_ns : 0;
      const float *px = prm.xyz + ((size_t)cloud * n + p) * 3;
      const float *pc = prm.new_xyz + ((size_t)cloud * prm.m + j) * 3;
      const float *pf = prm.feat + (size_t)cloud * c_feat * n;     // not dereferenced when c_feat == 0
      if (half == 0) {
        uint32_t r[8];
        unsigned off = (unsigned)p;                   // index of (channel, point p) in this cloud's features
        // first group: 3 coordinates + channels 0..4 (raw FP32 features: the tensor core truncates to TF32)
#pragma unroll
        for (int t = 0; t < 5; ++t) {
          r[3 + t] = t < c_feat ? __float_as_uint(__ldg(pf + off)) : 0u;
          off += (unsigned)n;
        }
        r[0] = round_tf32(__fsub_rn(__ldg(px), __ldg(pc)));
        r[1] = round_tf32(__fsub_rn(__ldg(px + 1), __ldg(pc + 1)));
        r[2] = round_tf32(__fsub_rn(__ldg(px + 2), __ldg(pc + 2)));
        tmem_st8(lane_addr + (uint32_t)prm.tm_a0, r);
      }
      // then kBatch channels (independent loads) at a time; the last batch is predicated / zero padded; the halves alternate
      for (int ch = 5 + half * kBatch; ch + 3 < prm.k0; ch += kBatch * kHalves) {
        unsigned off = (unsigned)p + (unsigned)ch * (unsigned)n;
        uint32_t v[kBatch];
        if (ch + kBatch <= c_feat) {
#pragma unroll
          for (int t = 0; t < kBatch; ++t) { v[t] = __float_as_uint(__ldg(pf + off)); off += (unsigned)n; }
        } else {
#pragma unroll
          for (int t = 0; t < kBatch; ++t) { v[t] = ch + t < c_feat ? __float_as_uint(__ldg(pf + off)) : 0u; off += (unsigned)n; }
        }
#pragma unroll
        for (int g = 0; g < kBatch / 8; ++g) {
          if (ch + 3 + g * 8 < prm.k0) {              // CTA-uniform
            uint32_t q[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) q[t] = v[g * 8 + t];
            tmem_st8(lane_addr + (uint32_t)(prm.tm_a0 + ch + 3 + g * 8), q);
          }
        }
      }
    }
    issue_layer(w1, prm.n1, prm.k0, (uint32_t)prm.tm_a0, (uint32_t)prm.tm_r1);
    // ---- layers 1 and 2: accumulator -> + shift, ReLU, round to TF32 -> A operand of the next layer, in place.
    // (Adding half a TF32 ulp is the rounding: the tensor core ignores the low 13 bits.)
#pragma unroll 1
    for (int l = 0; l < 2; ++l) {
      const int nl = l == 0 ? prm.n1 : prm.n2;
      const uint32_t acc = lane_addr + (uint32_t)(l == 0 ? prm.tm_r1 : prm.tm_r2);
      const float4 *sh = reinterpret_cast<const float4 *>(s_shift + (l == 0 ? 0 : prm.n1));
      mbar_wait(bar_d, phase_d & 1u);
      ++phase_d;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t ra[16], rb[16];
      if (kHalves > 1) {
        // kHalves warps per lane quarter: each takes every kHalves-th group of 16 columns (the second warp of the scheduler covers the
        // TMEM round trip of the first, so no software pipelining inside a thread)
        for (int c0 = 16 * half; c0 < nl; c0 += 16 * kHalves) {
          tmem_ld16(acc + (uint32_t)c0, ra);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 16; t += 4) {
            const float4 s4 = sh[(c0 + t) >> 2];
            shift_relu_round4(ra[t], ra[t + 1], ra[t + 2], ra[t + 3], s4);
          }
          tmem_st16(acc + (uint32_t)c0, ra);
        }
      } else {
      tmem_ld16(acc, ra);
      for (int c0 = 0; c0 < nl; c0 += 32) {
        tmem_ld_wait();
        if (c0 + 16 < nl) tmem_ld16(acc + (uint32_t)(c0 + 16), rb);      // in flight while ra is processed
#pragma unroll
        for (int t = 0; t < 16; t += 4) {
          const float4 s4 = sh[(c0 + t) >> 2];
          shift_relu_round4(ra[t], ra[t + 1], ra[t + 2], ra[t + 3], s4);
        }
        tmem_st16(acc + (uint32_t)c0, ra);
        if (c0 + 16 < nl) {
          tmem_ld_wait();
          if (c0 + 32 < nl) tmem_ld16(acc + (uint32_t)(c0 + 32), ra);
#pragma unroll
          for (int t = 0; t < 16; t += 4) {
            const float4 s4 = sh[(c0 + 16 + t) >> 2];
            shift_relu_round4(rb[t], rb[t + 1], rb[t + 2], rb[t + 3], s4);
          }
          tmem_st16(acc + (uint32_t)(c0 + 16), rb);
        }
      }
      }
      if (l == 0) issue_layer(w2, prm.n2, prm.n1, (uint32_t)prm.tm_r1, (uint32_t)prm.tm_r2);
      else issue_layer(w3, prm.n3, prm.n2, (uint32_t)prm.tm_r2, (uint32_t)prm.tm_r3);
    }
    // ---- layer 3: max over the nsample rows of every centre, then + shift and ReLU on the pooled value (the shift
    // is per channel and ReLU is monotone, so they commute with the max).  A butterfly reduce-scatter leaves lane t
    // of a pooling group with channel c0 + t.
    {
      const uint32_t acc = lane_addr + (uint32_t)prm.tm_r3;
      const float *sh = s_shift + prm.n1 + prm.n2;
      mbar_wait(bar_d, phase_d & 1u);
      ++phase_d;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int centre_local = row >> lg_ns;          // centre of this row inside the tile
      int c0 = 0;
      // 32 channels at a time through the 16x256b load shape (thread t receives rows {q, q + 8} of a 16-row half, q = t / 4, and
      // the column pairs 8 g + 2 (t % 4) + {0, 1}, g < 4 -- probed with tools/tmem_ld_probe.cu): after two loads a thread holds
      // FOUR rows of eight channels, so the first two levels of the max tree are plain FMNMX on its own registers and only the
      // eight lanes that share t % 4 are left to combine -- 7 exchanges per 32 channels instead of 31 (52 instructions for
      // 124; the butterfly was a quarter of this kernel's instructions).
      {
        const uint32_t warp_acc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)prm.tm_r3;
        const int q = lane >> 2;
        const int ch_in_block = 8 * (q >> 1) + 2 * (lane & 3) + (q & 1);     // the channel this lane ends up with
        const int warp_row0 = (warp & 3) * 32;
        for (; c0 + 32 <= prm.n3; c0 += 32) {
          if (kHalves > 1 && ((c0 >> 5) & (kHalves - 1)) != half) continue;     // warp-uniform: the shares take alternate channel groups
          uint32_t ra[16], rb[16];
          tmem_ld_16x256b_x4(warp_acc + (uint32_t)c0, ra);                       // rows 0..15 of this warp's lane quarter
          tmem_ld_16x256b_x4(warp_acc + (16u << 16) + (uint32_t)c0, rb);         // rows 16..31
          tmem_ld_wait();
          const int c = c0 + ch_in_block;
          if (ns >= 32) {
            float v[8];
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
              for (int j = 0; j < 2; ++j)
                v[2 * g + j] = fmaxf(fmaxf(__uint_as_float(ra[4 * g + j]), __uint_as_float(ra[4 * g + 2 + j])),
                                     fmaxf(__uint_as_float(rb[4 * g + j]), __uint_as_float(rb[4 * g + 2 + j])));
            const float mx = pool8_q(v, lane);
            if (c < prm.c3) {
              const int centre = warp_row0 >> lg_ns;
              const unsigned u = __float_as_uint(fmaxf(mx + sh[c], 0.f));
              if (ns == 32) s_pool[c * cpt + centre] = u;
              else atomicMax(&s_pool[c * cpt + centre], u);   // >= 0: unsigned order == float order
            }
          } else {   // ns == 16: rows 0..15 and 16..31 are two centres
            float va[8], vb[8];
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                va[2 * g + j] = fmaxf(__uint_as_float(ra[4 * g + j]), __uint_as_float(ra[4 * g + 2 + j]));
                vb[2 * g + j] = fmaxf(__uint_as_float(rb[4 * g + j]), __uint_as_float(rb[4 * g + 2 + j]));
              }
            const float ma = pool8_q(va, lane), mb = pool8_q(vb, lane);
            if (c < prm.c3) {
              const int centre = warp_row0 >> 4;
              s_pool[c * cpt + centre] = __float_as_uint(fmaxf(ma + sh[c], 0.f));
              s_pool[c * cpt + centre + 1] = __float_as_uint(fmaxf(mb + sh[c], 0.f));
            }
          }
        }
      }
      for (; c0 < prm.n3; c0 += 16) {
        if (kHalves > 1 && ((c0 >> 4) & (kHalves - 1)) != half) continue;
        uint32_t ra[16];
        tmem_ld16(acc + (uint32_t)c0, ra);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          v[t] = __uint_as_float(ra[t]);
          if (ns >= 32) v[t] = fmaxf(v[t], __shfl_xor_sync(0xFFFFFFFFu, v[t], 16));   // both half-warps belong to one centre
        }
        const float mx = pool_scatter<16>(v, lane);
        const int c = c0 + (lane & 15);
        if (c < prm.c3 && (ns < 32 || lane < 16)) {
          const unsigned u = __float_as_uint(fmaxf(mx + sh[c], 0.f));
          if (ns <= 32) s_pool[c * cpt + centre_local] = u;
          else atomicMax(&s_pool[c * cpt + centre_local], u);
        }
      }
      // the accumulator has been read: the next tile's gather may overwrite TMEM
      __syncthreads();
      const int centre0 = t_in * cpt;
      if (prm.out_pm) {
        // the same pooled values once more, point-major, for the next level's gather (channels fastest: coalesced)
        float *base = prm.out_pm + ((size_t)cloud * prm.m + centre0) * prm.ld_pm;
        // (shift / mask instead of an integer division per element when c3 is a power of two: the division was 8 % of the
        // instructions of the SA1 launches)
        const int sh3 = (prm.c3 & (prm.c3 - 1)) == 0 ? __ffs(prm.c3) - 1 : -1;
        for (int i = (int)threadIdx.x; i < cpt * prm.c3; i += kT) {
          int k, c;
          if (sh3 >= 0) { k = i >> sh3; c = i & (prm.c3 - 1); }
          else { k = i / prm.c3; c = i - k * prm.c3; }
          if (centre0 + k < prm.m) base[(size_t)k * prm.ld_pm + 3 + prm.out_coff + c] = __uint_as_float(s_pool[c * cpt + k]);
        }
        if (prm.pm_xyz) {   // this scale also writes the centre coordinates and the zero padding behind the channels
          const int tail0 = 3 + prm.out_ctot, per = 3 + (prm.ld_pm - tail0);
          for (int i = (int)threadIdx.x; i < cpt * per; i += kT) {
            const int k = i / per, q = i - k * per;
            if (centre0 + k < prm.m) {
              if (q < 3) base[(size_t)k * prm.ld_pm + q] = __ldg(prm.new_xyz + ((size_t)cloud * prm.m + centre0 + k) * 3 + q);
              else base[(size_t)k * prm.ld_pm + tail0 + (q - 3)] = 0.f;
            }
          }
        }
        if (ns > 32) __syncthreads();   // the loop below clears s_pool rows other threads have just read
      }
      for (int c = (int)threadIdx.x; c < prm.c3; c += kT) {
        float *dst = prm.out + ((size_t)cloud * prm.out_ctot + prm.out_coff + c) * prm.m + centre0;
        unsigned int *src = s_pool + c * cpt;
        if (cpt % 4 == 0 && centre0 + cpt <= prm.m && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
          for (int k = 0; k < cpt; k += 4) {
            *reinterpret_cast<uint4 *>(dst + k) = *reinterpret_cast<const uint4 *>(src + k);
            if (ns > 32) *reinterpret_cast<uint4 *>(src + k) = make_uint4(0u, 0u, 0u, 0u);
          }
        } else {
          for (int k = 0; k < cpt; ++k) {
            if (centre0 + k < prm.m) dst[k] = __uint_as_float(src[k]);
            if (ns > 32) src[k] = 0u;
          }
        }
      }
      // (the next write to s_pool comes after three more __syncthreads)
    }
    cloud = cloud_n; t_in = t_in_n;
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(prm.tmem_cols) : "memory");
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

// weights (rows x k_pad) row-major, box = 32 k x `rows`
bool weight_map(CUtensorMap *m, const float *w, int rows, int k_pad) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("sa_mlp_fused: cuTensorMapEncodeTiled unavailable"); return false; }
  const cuuint64_t dims[2] = {(cuuint64_t)k_pad, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)k_pad * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(w), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("sa_mlp_fused: cuTensorMapEncodeTiled failed (%d)", (int)r); return false; }
  return true;
}

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

struct Plan { int k0, n1, n2, n3, nk1, nk2, nk3, a0, r1, r2, r3, tmem_cols; size_t smem; int ctas_per_sm; bool ok; };

Plan make_plan(int c_feat, int nsample, int c1, int c2, int c3) {
  Plan p = {};
  p.k0 = round_up(3 + c_feat, 8);
  p.n1 = round_up(c1, 16); p.n2 = round_up(c2, 16); p.n3 = round_up(c3, 16);
  p.nk1 = ceil_div(p.k0, 32); p.nk2 = ceil_div(p.n1, 32); p.nk3 = ceil_div(p.n2, 32);
  // TMEM layouts (a region may overlay one that is dead by the time it is written; an accumulator must never
  // overlap the A operand of its own layer):
  //   sequential : a0 | r1 | r2 | r3
  //   A          : r2 over a0, r3 over r1            (needs n2 <= k0)
  //   C          : r3 over a0, then r1, r2           (the input tile is dead long before layer 3)
  int total = p.k0 + p.n1 + p.n2 + p.n3;
  p.a0 = 0; p.r1 = p.k0; p.r2 = p.k0 + p.n1; p.r3 = p.r2 + p.n2;
  if (p.n2 <= p.k0) {
    const int t = p.k0 + (p.n1 > p.n3 ? p.n1 : p.n3);
    if (t < total) { total = t; p.a0 = 0; p.r1 = p.k0; p.r2 = 0; p.r3 = p.k0; }
  }
  {
    const int head = p.n3 > p.k0 ? p.n3 : p.k0;
    const int t = head + p.n1 + p.n2;
    if (t < total) { total = t; p.a0 = 0; p.r3 = 0; p.r1 = head; p.r2 = head + p.n1; }
  }
  p.tmem_cols = total <= 32 ? 32 : total <= 64 ? 64 : total <= 128 ? 128 : total <= 256 ? 256 : 512;
  const int cpt = nsample > 0 && nsample <= 128 ? 128 / nsample : 1;
  p.smem = ((size_t)p.nk1 * p.n1 + (size_t)p.nk2 * p.n2 + (size_t)p.nk3 * p.n3) * 128 + 1024   // weights + alignment slack
           + (size_t)(p.n1 + p.n2 + p.n3) * 4 + (size_t)c3 * cpt * 4 + 16;
  p.ok = total <= 512 && p.n1 <= 256 && p.n2 <= 256 && p.n3 <= 256 && p.smem <= 232448 - 1024;
  const int by_tmem = 512 / p.tmem_cols;
  const int by_smem = (int)((233472 - 2048) / (p.smem + 1024 + 256));   // 228 KB per SM, 1 KB reserved per CTA
  p.ctas_per_sm = by_tmem < by_smem ? by_tmem : by_smem;
  if (p.ctas_per_sm > 4) p.ctas_per_sm = 4;   // registers: 4 x 128 threads x <= 128
  if (p.ctas_per_sm < 1) p.ok = false;
  return p;
}

}  // namespace
}  // namespace ws3d

using namespace ws3d;

WS3D_API int ws3d_sa_mlp_fused_supported(int c_feat, int nsample, int c1, int c2, int c3) {
  if (c_feat < 0 || c1 <= 0 || c2 <= 0 || c3 <= 0 || nsample < 16 || nsample > kRows || (nsample & (nsample - 1))) return 0;
  return make_plan(c_feat, nsample, c1, c2, c3).ok ? 1 : 0;
}

// One set-abstraction scale: grouping + 3 x (conv1x1 + BN(eval, folded) + ReLU) + max over nsample.
//   xyz (B,n,3), new_xyz (B,m,3), features (B,c_feat,n) or NULL (c_feat = 0), idx (B,m,nsample) from ball_query;
//   w1 (n1 x 32*ceil(k0/32)), w2 (n2 x 32*ceil(n1/32)), w3 (n3 x 32*ceil(n2/32)): BN-folded, zero padded, TF32-rounded,
//     row-major; k0 = roundup(3 + c_feat, 8), n_l = roundup(c_l, 16); w1's columns are [dx,dy,dz, features...];
//   shift_l (n_l) zero padded;  out (B, out_ctot, m): channels [out_coff, out_coff + c3) are written.
static int sa_mlp_fused_impl(int b, int n, int m, int nsample, int c_feat, const float *xyz, const float *new_xyz,
                             const float *features, const int *idx, int c1, int c2, int c3, const float *w1,
                             const float *shift1, const float *w2, const float *shift2, const float *w3,
                             const float *shift3, float *out, int out_ctot, int out_coff, const float *rows, int ld,
                             float *out_pm, int ld_pm, int pm_xyz, ws3d_stream_t stream) {
  const char *what = "sa_mlp_fused";
  if (b < 0 || n <= 0 || m < 0 || out_coff < 0 || out_coff + c3 > out_ctot) return fail_arg(what);
  if (!ws3d_sa_mlp_fused_supported(c_feat, nsample, c1, c2, c3)) return fail_arg("sa_mlp_fused (unsupported shape)");
  if (b == 0 || m == 0) return 0;
  if (!new_xyz || !idx || !w1 || !w2 || !w3 || !shift1 || !shift2 || !shift3 || !out) return fail_arg(what);
  if (rows) {
    if (ld < 3 + c_feat || ld % 4 || (reinterpret_cast<uintptr_t>(rows) & 15u)) return fail_arg("sa_mlp_fused (rows: ld % 4, ld >= 3 + c_feat, 16-byte aligned)");
  } else if (!xyz || (c_feat > 0 && !features)) {
    return fail_arg(what);
  }
  if (out_pm && (ld_pm < 3 + out_ctot || ld_pm % 4)) return fail_arg("sa_mlp_fused (out_pm: ld_pm % 4, ld_pm >= 3 + out_ctot)");
  const Plan pl = make_plan(c_feat, nsample, c1, c2, c3);
  SaFusedParams prm;
  prm.n = n; prm.m = m; prm.ns = nsample; prm.c_feat = c_feat;
  prm.k0 = pl.k0; prm.n1 = pl.n1; prm.n2 = pl.n2; prm.n3 = pl.n3; prm.c3 = c3;
  prm.nk1 = pl.nk1; prm.nk2 = pl.nk2; prm.nk3 = pl.nk3;
  prm.tm_a0 = pl.a0; prm.tm_r1 = pl.r1; prm.tm_r2 = pl.r2; prm.tm_r3 = pl.r3; prm.tmem_cols = pl.tmem_cols;
  const long long cols = (long long)m * nsample;
  prm.tiles_per_cloud = (int)((cols + kRows - 1) / kRows);
  const long long tiles = (long long)prm.tiles_per_cloud * b;
  if (tiles > 0x7FFFFFFFLL || cols > 0x7FFFFFFFLL) return fail_arg("sa_mlp_fused (too many tiles)");
  prm.n_tiles = (int)tiles;
  prm.xyz = xyz; prm.new_xyz = new_xyz; prm.feat = c_feat > 0 ? features : nullptr; prm.idx = idx;
  prm.shift1 = shift1; prm.shift2 = shift2; prm.shift3 = shift3;
  prm.out = out; prm.out_ctot = out_ctot; prm.out_coff = out_coff;
  prm.rows = rows; prm.ld = ld; prm.out_pm = out_pm; prm.ld_pm = ld_pm; prm.pm_xyz = pm_xyz;
  CUtensorMap m1, m2, m3;
  if (!weight_map(&m1, w1, pl.n1, pl.nk1 * 32) || !weight_map(&m2, w2, pl.n2, pl.nk2 * 32) ||
      !weight_map(&m3, w3, pl.n3, pl.nk3 * 32))
    return (int)cudaErrorInvalidValue;
  const bool wide = pl.ctas_per_sm <= 2 && c_feat > 37;
  // one CTA per SM (resident weights > half the shared memory): four column shares, 512 threads; WS3D_SA_HALVES=1|2|4 overrides (A/B runs)
  static const int env_halves = [] { const char *e = getenv("WS3D_SA_HALVES"); return e && *e ? atoi(e) : -1; }();
  int halves = pl.ctas_per_sm == 1 ? 4 : 1;
  if (pl.ctas_per_sm == 1 && (env_halves == 1 || env_halves == 2 || env_halves == 4)) halves = env_halves;
  // (two CTAs per SM x 256 threads was measured for the SA2 shapes: 0.305 ms either way -- they stay at 128 threads)
  auto kern = halves == 4 ? sa_mlp_fused_kernel<32, 1, 4>
              : halves == 2 ? sa_mlp_fused_kernel<64, 1, 2>
              : wide        ? sa_mlp_fused_kernel<64, 2, 1>
                            : sa_mlp_fused_kernel<32, 4, 1>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) { set_error("sa_mlp_fused: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  const int pc = persistent_ctas(pl.ctas_per_sm);
  const int ctas = (int)(tiles < (long long)pc ? tiles : (long long)pc);
  kern<<<ctas, kThreads * halves, pl.smem, to_stream(stream)>>>(m1, m2, m3, prm);
  return check_launch(what);
}

WS3D_API int ws3d_sa_mlp_fused(int b, int n, int m, int nsample, int c_feat, const float *xyz, const float *new_xyz,
                               const float *features, const int *idx, int c1, int c2, int c3, const float *w1,
                               const float *shift1, const float *w2, const float *shift2, const float *w3,
                               const float *shift3, float *out, int out_ctot, int out_coff, ws3d_stream_t stream) {
  return sa_mlp_fused_impl(b, n, m, nsample, c_feat, xyz, new_xyz, features, idx, c1, c2, c3, w1, shift1, w2, shift2, w3, shift3, out,
                           out_ctot, out_coff, nullptr, 0, nullptr, 0, 0, stream);
}

// Same with the grouped points gathered from point-major rows (B, n, ld) = [x, y, z, features..., zeros] (see SaFusedParams)
// and, optionally, this scale's pooled channels also written point-major into out_pm (B, m, ld_pm) for the next level.
WS3D_API int ws3d_sa_mlp_fused_rows(int b, int n, int m, int nsample, int c_feat, const float *rows, int ld, const float *new_xyz,
                                    const int *idx, int c1, int c2, int c3, const float *w1, const float *shift1, const float *w2,
                                    const float *shift2, const float *w3, const float *shift3, float *out, int out_ctot, int out_coff,
                                    float *out_pm, int ld_pm, int pm_xyz, ws3d_stream_t stream) {
  if (!rows) return fail_arg("sa_mlp_fused_rows (null pointer)");
  return sa_mlp_fused_impl(b, n, m, nsample, c_feat, nullptr, new_xyz, nullptr, idx, c1, c2, c3, w1, shift1, w2, shift2, w3, shift3, out,
                           out_ctot, out_coff, rows, ld, out_pm, ld_pm, pm_xyz, stream);
}
