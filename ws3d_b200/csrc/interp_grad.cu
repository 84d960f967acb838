// three_interpolate_grad without atomics (reference: interpolate_gpu.cu:112-137, one float atomicAdd per
// (unknown point, channel, neighbour) into grad_points, order-nondeterministic).
//
// The stencil (idx, weight) does not depend on the channel, so its inverse is built ONCE per call -- for every known point j the
// list of (unknown point i, weight) that reference it, a CSR over the 3n stencil entries of a cloud, sorted by (i, k) -- and the
// gradient becomes a GATHER:   grad_points[b, c, j] += sum_{(i, w) in list(j)} w * grad_out[b, c, i]
// The lists are very uneven on LiDAR-shaped clouds (a known point in the dense near field is the neighbour of hundreds of unknown
// points, one in the far field of three): one thread sums one list, and the builder also emits the known points ORDERED BY LIST
// LENGTH, so the 32 lists of a warp are equally long and no lane idles behind a long neighbour.
// The rows grad_out[b, c, :] of a channel chunk are staged in shared memory with one TMA bulk copy (they are contiguous in
// (B, C, n)), so the random reads are shared-memory reads, every grad_points element is written once by one thread with a
// coalesced store, and the summation order is fixed: the result is bit-reproducible run to run, which the atomic version (and the
// reference) is not.  Measured at the four FP levels of the Stage-1 training step (32 scenes): profiles/r2_interp_grad_bench.json.
// Tried and not kept (measured, B = 32, FP0): a warp per block of 16 lists with lane = entry and a segmented shuffle scan -- perfectly
// balanced, but 127 instructions per 32 entries x 3 channels (17 SHFL + 15 FADD + the segment bookkeeping) made it issue-bound
// (80 % issue-active, 1.05 ms against 0.89 ms for the unsorted one-thread-per-list form at 5 of 32 lanes active).
// Also measured: group_points_grad through the same machinery (lists = the grouped slots of every source point, weight 1).  Its rows
// are npoint x nsample floats (128 KB at SA2), so one channel fills a CTA's shared memory and the lists are re-read per channel:
// 1.76 ms against 0.70 ms for the atomic scatter over the six launches of a training step -- the grouping gradient keeps atomics.
//
// Clouds whose inverse does not fit the one-CTA-per-cloud builder (m > kMaxKnown) or whose rows do not fit shared memory
// (n > kMaxRow) keep the atomic kernel in pointnet2_ops.cu.
#include "common.cuh"
#include "tma.cuh"

namespace ws3d {
namespace {

constexpr int kBuildThreads = 1024;
constexpr int kMaxGatherThreads = 1024;
constexpr int kMaxSmem = 224 * 1024;     // rows + list offsets of one CTA
constexpr int kMaxKnown = 24576;          // counters + cursors of a cloud in shared memory (2 x 4 B x m)
constexpr int kMaxRow = 49152;            // floats of one grad_out row staged in shared memory (192 KB)
constexpr int kStageBudget = 96 * 1024;   // bytes of staged rows per CTA when several channels fit: two CTAs per SM overlap copy and sums

struct __align__(8) Entry { int key; float w; };   // key = 3 i + k in the unsorted buffer, i in the sorted one

// One CTA per cloud: count -> exclusive scan -> fill -> per-list sort by (i, k).
// `total` entries per cloud; entry e refers to row element e / div (div = 3: the three neighbours of an unknown point; div = 1:
// one grouped slot); weight == nullptr means weight 1 (grouping).
__global__ void __launch_bounds__(kBuildThreads) stencil_inverse_kernel(int total, int m, int div, const int *__restrict__ idx,
                                                                        const float *__restrict__ weight, int *__restrict__ off_g,
                                                                        Entry *__restrict__ ent_g, Entry *__restrict__ out_g,
                                                                        unsigned short *__restrict__ order_g) {
  extern __shared__ int sm_i[];
  int *cnt = sm_i;             // m + 1
  int *cur = sm_i + (m + 1);   // m
  __shared__ int warp_tot[kBuildThreads / 32];
  const size_t cloud = blockIdx.x;
  const int *ip = idx + cloud * (size_t)total;
  const float *wp = weight ? weight + cloud * (size_t)total : nullptr;
  int *off = off_g + cloud * (size_t)(m + 1);
  Entry *ent = ent_g + cloud * (size_t)total;
  const int t = threadIdx.x;
  for (int j = t; j <= m; j += kBuildThreads) cnt[j] = 0;
  __syncthreads();
  for (int e = t; e < total; e += kBuildThreads) {
    const int j = __ldg(ip + e);
    if (j >= 0 && j < m) atomicAdd(&cnt[j], 1);
  }
  __syncthreads();
  // exclusive scan of cnt[0..m) in chunks of kBuildThreads, running base carried in cnt[m]
  int base = 0;
  for (int j0 = 0; j0 < m; j0 += kBuildThreads) {
    const int j = j0 + t;
    const int v = j < m ? cnt[j] : 0;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, d);
      if ((t & 31) >= d) x += y;
    }
    if ((t & 31) == 31) warp_tot[t >> 5] = x;
    __syncthreads();
    if (t < 32) {
      int w = warp_tot[t];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, d);
        if (t >= d) w += y;
      }
      warp_tot[t] = w;
    }
    __syncthreads();
    const int before = base + ((t >> 5) ? warp_tot[(t >> 5) - 1] : 0) + x - v;
    if (j < m) {
      cnt[j] = before;
      cur[j] = before;
      off[j] = before;
    }
    base += warp_tot[kBuildThreads / 32 - 1];
    __syncthreads();
  }
  if (t == 0) { cnt[m] = base; off[m] = base; }
  __syncthreads();
  for (int e = t; e < total; e += kBuildThreads) {
    const int j = __ldg(ip + e);
    if (j >= 0 && j < m) {
      const int slot = atomicAdd(&cur[j], 1);
      Entry en;
      en.key = e;
      en.w = wp ? __ldg(wp + e) : 1.0f;
      ent[slot] = en;
    }
  }
  __syncthreads();   // the CTA's own global writes are visible to the CTA after the barrier
  // Every list ordered by e = 3 i + k (unique), so that the summation order of the gather is fixed: a warp per list ranks each
  // entry by counting the smaller keys (lists are short on average; broadcast reads of the list hit L1) and writes it to its place
  // in the second buffer as {i, w}.
  // known points by descending list length (bins 255 .. 0, longer lists share bin 255): which thread of the gather sums which
  // list; the order inside a bin is irrelevant to the result
  __shared__ int hist[256];
  if (t < 256) hist[t] = 0;
  __syncthreads();
  for (int j = t; j < m; j += kBuildThreads) atomicAdd(&hist[255 - min(cnt[j + 1] - cnt[j], 255)], 1);
  __syncthreads();
  if (t < 32) {
    int run = 0;
    for (int k = 0; k < 8; ++k) run += hist[t * 8 + k];
    int x = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, d);
      if (t >= d) x += y;
    }
    int before = x - run;
    for (int k = 0; k < 8; ++k) {
      const int h = hist[t * 8 + k];
      hist[t * 8 + k] = before;
      before += h;
    }
  }
  __syncthreads();
  for (int j = t; j < m; j += kBuildThreads) {
    const int pos = atomicAdd(&hist[255 - min(cnt[j + 1] - cnt[j], 255)], 1);
    order_g[cloud * (size_t)m + pos] = (unsigned short)j;
  }
  const int lane = t & 31;
  for (int j = t >> 5; j < m; j += kBuildThreads / 32) {
    const int b = cnt[j], L = cnt[j + 1] - b;
    for (int p = lane; p < L; p += 32) {
      const Entry x = ent[b + p];
      int rank = 0;
      for (int q = 0; q < L; ++q) rank += (ent[b + q].key < x.key) ? 1 : 0;
      Entry y;
      y.key = x.key / div;
      y.w = x.w;
      out_g[cloud * (size_t)total + b + rank] = y;
    }
  }
}

__device__ __forceinline__ Entry load_entry(const Entry *p) {
  const int2 raw = __ldg(reinterpret_cast<const int2 *>(p));
  Entry en;
  en.key = raw.x;
  en.w = __int_as_float(raw.y);
  return en;
}

// grad_points[j] += sum.  Every element is the target of exactly ONE such operation per call, so the result is the same as a plain
// read-add-write -- but a reduction without a return value does not put an L2 round trip on the thread's dependency chain.
__device__ __forceinline__ void red_add(float *p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// grid (channel chunks, clouds).  G channels per work item are accumulated in registers; an item = (list, channel group), lists in
// the builder's length order.  The chain of an item is list offsets -> entries (L2) -> staged rows (shared memory) -> FMA: the
// offsets and the order of the cloud sit in shared memory next to the rows, and the entries of a list are fetched four at a time
// so that one L2 latency covers four of them.
template <int G>
__global__ void __launch_bounds__(kMaxGatherThreads) interp_grad_gather_kernel(int c, int n, int m, int chunk, int total,
                                                                               long long grad_cloud_stride,
                                                                               const float *__restrict__ grad_out,
                                                                               const int *__restrict__ off_g,
                                                                               const Entry *__restrict__ ent_g,
                                                                               const unsigned short *__restrict__ order_g,
                                                                               float *__restrict__ grad_points) {
  extern __shared__ __align__(16) float rows[];   // chunk x n floats, m + 1 list offsets, m list numbers
  __shared__ __align__(8) unsigned long long bar;
  const size_t cloud = blockIdx.y;
  const int c0 = blockIdx.x * chunk, chs = min(chunk, c - c0);
  int *off = reinterpret_cast<int *>(rows + (size_t)chunk * n);
  unsigned short *order = reinterpret_cast<unsigned short *>(off + m + 1);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    const int *og = off_g + cloud * (size_t)(m + 1);
    for (int j = threadIdx.x; j <= m; j += blockDim.x) off[j] = __ldg(og + j);
    const unsigned short *rg = order_g + cloud * (size_t)m;
    for (int j = threadIdx.x; j < m; j += blockDim.x) order[j] = __ldg(rg + j);
  }
  __syncthreads();
  stage_floats(rows, grad_out + cloud * (size_t)grad_cloud_stride + (size_t)c0 * n, chs * n, &bar, 0);
  const Entry *ent = ent_g + cloud * (size_t)total;
  float *out = grad_points + (cloud * (size_t)c + c0) * m;
  const int groups = (chs + G - 1) / G;
  for (int item = threadIdx.x; item < groups * m; item += blockDim.x) {
    const int g = item / m, j = order[item - g * m];
    const int ch0 = g * G, live = min(G, chs - ch0);
    const int b = off[j], e = off[j + 1];
    if (b == e) continue;
    float acc[G];
#pragma unroll
    for (int u = 0; u < G; ++u) acc[u] = 0.f;
    const float *r0 = rows + (size_t)ch0 * n;
    int a = b;
    for (; a + 4 <= e; a += 4) {
      Entry en[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) en[q] = load_entry(ent + a + q);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int u = 0; u < G; ++u)
          if (u < live) acc[u] = __fmaf_rn(en[q].w, r0[u * n + en[q].key], acc[u]);
      }
    }
    for (; a < e; ++a) {
      const Entry en = load_entry(ent + a);
#pragma unroll
      for (int u = 0; u < G; ++u)
        if (u < live) acc[u] = __fmaf_rn(en.w, r0[u * n + en.key], acc[u]);
    }
#pragma unroll
    for (int u = 0; u < G; ++u)
      if (u < live) red_add(out + (size_t)(ch0 + u) * m + j, acc[u]);
  }
}

}  // namespace

// Returns 0 when the gather path ran, 1 when the shape is outside its limits (the caller falls back to the atomic kernel),
// or a cudaError code.
// Generic form: out[b, ch, idx[b, e]] += weight[b, e] * grad_out[b, ch, e / div] for e < total, rows of n = total / div floats,
// m targets per cloud; a cloud's rows start grad_cloud_stride floats apart (rows of one cloud are contiguous).
int scatter_as_gather(int b, int c, int n, int total, int div, int m, const float *grad_out, long long grad_cloud_stride, const int *idx,
                      const float *weight, float *grad_points, cudaStream_t stream, const char *what) {
  static_assert(kMaxKnown <= 65536, "the length order stores list numbers in 16 bits");
  if (m < 1 || m > kMaxKnown || n < 1 || n > kMaxRow) return 1;
  // channels per CTA and per work item
  const size_t off_smem = (size_t)(m + 1) * sizeof(int) + (size_t)((m + 1) & ~1) * sizeof(unsigned short);
  if (off_smem + (size_t)n * 4 > (size_t)kMaxSmem) return 1;
  int chunk, G, threads;
  if ((size_t)n * 4 * 2 > (size_t)kStageBudget) {
    // long rows (FP0: 16384 floats = 64 KB): as many as fit one CTA per SM, the stencil entries are read once per chunk;
    // 1024 threads, because that one CTA is all the latency hiding the SM has
    chunk = (int)(((size_t)kMaxSmem - off_smem) / ((size_t)n * 4));
    if (chunk > 4) chunk = 4;
    if (chunk > c) chunk = c;
    G = chunk;
    threads = 1024;
  } else {
    chunk = (int)(kStageBudget / ((size_t)n * 4));
    if (chunk > 32) chunk = 32;
    if (chunk > c) chunk = c;
    G = chunk >= 4 ? 4 : chunk;
    chunk = chunk / G * G;
    threads = 512;
  }
  if (b > 65535 || ceil_div(c, chunk) > 0x7fffffff) return 1;
  const size_t off_bytes = ((size_t)b * (m + 1) * sizeof(int) + 255) & ~(size_t)255;
  const size_t ent_bytes = (size_t)b * total * sizeof(Entry);
  const size_t order_bytes = ((size_t)b * m * sizeof(unsigned short) + 255) & ~(size_t)255;
  char *ws = (char *)scratch(off_bytes + 2 * ent_bytes + order_bytes, 7);
  if (!ws) return (int)cudaErrorMemoryAllocation;
  int *off = (int *)ws;
  Entry *unsorted = (Entry *)(ws + off_bytes);
  Entry *ent = (Entry *)(ws + off_bytes + ent_bytes);
  unsigned short *order = (unsigned short *)(ws + off_bytes + 2 * ent_bytes);
  const size_t build_smem = (size_t)(2 * m + 1) * sizeof(int);
  static bool attr_done_dev[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  bool &attr_done = attr_done_dev[dev & 15];
  if (!attr_done) {
    cudaFuncSetAttribute(stencil_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (2 * kMaxKnown + 1) * (int)sizeof(int));
    cudaFuncSetAttribute(interp_grad_gather_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(interp_grad_gather_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(interp_grad_gather_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(interp_grad_gather_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    attr_done = true;
  }
  stencil_inverse_kernel<<<b, kBuildThreads, build_smem, stream>>>(total, m, div, idx, weight, off, unsorted, ent, order);
  int rc = check_launch(what);
  if (rc) return rc;
  dim3 grid((unsigned)ceil_div(c, chunk), (unsigned)b);
  const size_t smem = (size_t)chunk * n * sizeof(float) + off_smem;
  switch (G) {
    case 1: interp_grad_gather_kernel<1><<<grid, threads, smem, stream>>>(c, n, m, chunk, total, grad_cloud_stride, grad_out, off, ent, order, grad_points); break;
    case 2: interp_grad_gather_kernel<2><<<grid, threads, smem, stream>>>(c, n, m, chunk, total, grad_cloud_stride, grad_out, off, ent, order, grad_points); break;
    case 3: interp_grad_gather_kernel<3><<<grid, threads, smem, stream>>>(c, n, m, chunk, total, grad_cloud_stride, grad_out, off, ent, order, grad_points); break;
    default: interp_grad_gather_kernel<4><<<grid, threads, smem, stream>>>(c, n, m, chunk, total, grad_cloud_stride, grad_out, off, ent, order, grad_points); break;
  }
  return check_launch(what);
}

int three_interpolate_grad_gather(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight,
                                  float *grad_points, cudaStream_t stream) {
  if ((long long)n * 3 > 0x3fffffffLL) return 1;
  return scatter_as_gather(b, c, n, 3 * n, 3, m, grad_out, (long long)c * n, idx, weight, grad_points, stream, "three_interpolate_grad (gather)");
}

}  // namespace ws3d
