// Rotated-box BEV overlap / IoU matrices and rotated / axis-aligned NMS for B200.
//
// Replaces lib/utils/iou3d/src/iou3d_kernel.cu (thread-per-pair kernels with 208 B of local
// stack, per-pair cosf/sinf, an atan2f bubble sort whose comparator recomputes atan2f, a full
// N^2/64 mask including the lower triangle) and the host tail of iou3d.cpp:73-171 (cudaMalloc
// per call, 8*N*ceil(N/64) bytes D2H, single-threaded greedy scan on the CPU).
//
// Design:
//   * per-box work (cosf/sinf of +-angle, rotated corners, margins, AABB) is done once per box per
//     tile and kept in shared memory; a pair first meets an AABB test on the rotated corners and
//     only pairs whose padded AABBs touch run the exact polygon clipping;
//   * NMS computes only upper-triangle 64x64 tiles, keeps the mask on the device (it lives in L2)
//     and runs the greedy suppression in one CTA, 64 boxes per step: the diagonal word chain is
//     resolved by one thread, the surviving rows are OR-ed into the removal bitmap by the CTA.
//
// Exactness: every float operation of the per-pair routine is spelled with the rounding order
// the reference object code uses (read from the SASS of oracle/_ref/obj/iou3d_kernel.o): e.g.
// "a*b - c*d" is fma(a, b, -rn(c*d)) except for s2/s5 of the segment test, whose two products are
// shared and therefore both rounded.  The AABB shortcut returns exactly what the reference returns
// for such pairs (no crossing, no corner inside => cnt == 0 => area 0); the pad covers the 1e-5
// margin of check_in_box2d plus rounding.
#include "common.cuh"

namespace ws3d {
namespace {

constexpr float kEps = 1e-8f;
constexpr float kMargin = 1e-5f;

struct BoxPre {
  float x1, y1, x2, y2;  // raw extents (iou3d_kernel.cu box layout [x1,y1,x2,y2,ry])
  float cx, cy;          // centre
  float px[4], py[4];    // corners rotated by +angle about the centre
  float cn, sn;          // cosf(-angle), sinf(-angle)  (check_in_box2d)
  float lox, hix, loy, hiy;  // AABB of the rotated corners, already padded
};

__device__ __forceinline__ void rot_about(float px, float py, float cx, float cy, float c, float s, float &ox,
                                          float &oy) {
  const float dx = __fsub_rn(px, cx), dy = __fsub_rn(py, cy);
  ox = __fadd_rn(__fmaf_rn(dx, c, __fmul_rn(dy, s)), cx);
  oy = __fadd_rn(__fmaf_rn(dy, c, -__fmul_rn(dx, s)), cy);
}

__device__ __forceinline__ void precompute(const float *__restrict__ box, BoxPre &o) {
  const float x1 = box[0], y1 = box[1], x2 = box[2], y2 = box[3], ang = box[4];
  o.x1 = x1; o.y1 = y1; o.x2 = x2; o.y2 = y2;
  o.cx = __fmul_rn(__fadd_rn(x1, x2), 0.5f);
  o.cy = __fmul_rn(__fadd_rn(y1, y2), 0.5f);
  const float c = cosf(ang), s = sinf(ang);
  rot_about(x1, y1, o.cx, o.cy, c, s, o.px[0], o.py[0]);
  rot_about(x2, y1, o.cx, o.cy, c, s, o.px[1], o.py[1]);
  rot_about(x2, y2, o.cx, o.cy, c, s, o.px[2], o.py[2]);
  rot_about(x1, y2, o.cx, o.cy, c, s, o.px[3], o.py[3]);
  o.cn = cosf(-ang);
  o.sn = sinf(-ang);
  float lox = fminf(fminf(o.px[0], o.px[1]), fminf(o.px[2], o.px[3]));
  float hix = fmaxf(fmaxf(o.px[0], o.px[1]), fmaxf(o.px[2], o.px[3]));
  float loy = fminf(fminf(o.py[0], o.py[1]), fminf(o.py[2], o.py[3]));
  float hiy = fmaxf(fmaxf(o.py[0], o.py[1]), fmaxf(o.py[2], o.py[3]));
  const float mag = fmaxf(fmaxf(fabsf(lox), fabsf(hix)), fmaxf(fabsf(loy), fabsf(hiy)));
  const float pad = 1e-4f + 1e-5f * mag;
  o.lox = lox - pad; o.hix = hix + pad; o.loy = loy - pad; o.hiy = hiy + pad;
}

__device__ __forceinline__ bool aabb_apart(const BoxPre &a, const BoxPre &b) {
  // written so that NaN extents fall through to the exact path
  return (a.lox > b.hix) || (b.lox > a.hix) || (a.loy > b.hiy) || (b.loy > a.hiy);
}

// check_in_box2d (iou3d_kernel.cu:50-65)
__device__ __forceinline__ bool in_box2d(const BoxPre &b, float px, float py) {
  float rx, ry;
  rot_about(px, py, b.cx, b.cy, b.cn, b.sn, rx, ry);
  return rx > __fsub_rn(b.x1, kMargin) && rx < __fadd_rn(b.x2, kMargin) && ry > __fsub_rn(b.y1, kMargin) &&
         ry < __fadd_rn(b.y2, kMargin);
}

// intersection (iou3d_kernel.cu:67-96) incl. check_rect_cross (:42-48)
__device__ __forceinline__ bool seg_intersection(float p1x, float p1y, float p0x, float p0y, float q1x, float q1y,
                                                 float q0x, float q0y, float &ax, float &ay) {
  if (!(fminf(p0x, p1x) <= fmaxf(q0x, q1x) && fminf(q0x, q1x) <= fmaxf(p0x, p1x) &&
        fminf(p0y, p1y) <= fmaxf(q0y, q1y) && fminf(q0y, q1y) <= fmaxf(p0y, p1y)))
    return false;
  const float s1 = __fmaf_rn(__fsub_rn(q0x, p0x), __fsub_rn(p1y, p0y),
                             -__fmul_rn(__fsub_rn(p1x, p0x), __fsub_rn(q0y, p0y)));
  const float pa = __fmul_rn(__fsub_rn(p1x, p0x), __fsub_rn(q1y, p0y));
  const float pb = __fmul_rn(__fsub_rn(q1x, p0x), __fsub_rn(p1y, p0y));
  const float s2 = __fsub_rn(pa, pb);
  const float s3 = __fmaf_rn(__fsub_rn(p0x, q0x), __fsub_rn(q1y, q0y),
                             -__fmul_rn(__fsub_rn(q1x, q0x), __fsub_rn(p0y, q0y)));
  const float s4 = __fmaf_rn(__fsub_rn(q1x, q0x), __fsub_rn(p1y, q0y),
                             -__fmul_rn(__fsub_rn(p1x, q0x), __fsub_rn(q1y, q0y)));
  if (!(__fmul_rn(s1, s2) > 0.f && __fmul_rn(s3, s4) > 0.f)) return false;
  const float s5 = __fsub_rn(pb, pa);
  const float den = __fsub_rn(s5, s1);
  if (fabsf(den) > kEps) {
    ax = __fdiv_rn(__fmaf_rn(s5, q0x, -__fmul_rn(s1, q1x)), den);
    ay = __fdiv_rn(__fmaf_rn(s5, q0y, -__fmul_rn(s1, q1y)), den);
  } else {
    const float a0 = __fsub_rn(p0y, p1y), b0 = __fsub_rn(p1x, p0x), c0 = __fmaf_rn(p0x, p1y, -__fmul_rn(p1x, p0y));
    const float a1 = __fsub_rn(q0y, q1y), b1 = __fsub_rn(q1x, q0x), c1 = __fmaf_rn(q0x, q1y, -__fmul_rn(q1x, q0y));
    const float D = __fmaf_rn(a0, b1, -__fmul_rn(a1, b0));
    ax = __fdiv_rn(__fmaf_rn(b0, c1, -__fmul_rn(b1, c0)), D);
    ay = __fdiv_rn(__fmaf_rn(a1, c0, -__fmul_rn(a0, c1)), D);
  }
  return true;
}

// box_overlap (iou3d_kernel.cu:108-212) on precomputed boxes; exact path.
__device__ __noinline__ float overlap_exact(const BoxPre &a, const BoxPre &b) {
  float qx[16], qy[16], qa[16];
  float sx = 0.f, sy = 0.f;
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int i1 = (i + 1) & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int j1 = (j + 1) & 3;
      float ix, iy;
      if (seg_intersection(a.px[i1], a.py[i1], a.px[i], a.py[i], b.px[j1], b.py[j1], b.px[j], b.py[j], ix, iy)) {
        sx = __fadd_rn(sx, ix);
        sy = __fadd_rn(sy, iy);
        qx[cnt] = ix; qy[cnt] = iy;
        ++cnt;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (in_box2d(a, b.px[k], b.py[k])) {
      sx = __fadd_rn(sx, b.px[k]);
      sy = __fadd_rn(sy, b.py[k]);
      qx[cnt] = b.px[k]; qy[cnt] = b.py[k];
      ++cnt;
    }
    if (in_box2d(b, a.px[k], a.py[k])) {
      sx = __fadd_rn(sx, a.px[k]);
      sy = __fadd_rn(sy, a.py[k]);
      qx[cnt] = a.px[k]; qy[cnt] = a.py[k];
      ++cnt;
    }
  }
  if (cnt < 3) return 0.f;  // 0, 1 or 2 points: the shoelace sum below is exactly 0 in the reference too
  const float fc = (float)cnt;
  const float mx = __fdiv_rn(sx, fc), my = __fdiv_rn(sy, fc);
  for (int i = 0; i < cnt; ++i) qa[i] = atan2f(__fsub_rn(qy[i], my), __fsub_rn(qx[i], mx));
  // bubble sort with the reference's comparator: swap when angle[i] > angle[i+1]
  for (int j = 0; j < cnt - 1; ++j)
    for (int i = 0; i < cnt - j - 1; ++i)
      if (qa[i] > qa[i + 1]) {
        float t;
        t = qa[i]; qa[i] = qa[i + 1]; qa[i + 1] = t;
        t = qx[i]; qx[i] = qx[i + 1]; qx[i + 1] = t;
        t = qy[i]; qy[i] = qy[i + 1]; qy[i + 1] = t;
      }
  float area = 0.f;
  for (int k = 0; k < cnt - 1; ++k) {
    const float ax = __fsub_rn(qx[k], qx[0]), ay = __fsub_rn(qy[k], qy[0]);
    const float bx = __fsub_rn(qx[k + 1], qx[0]), by = __fsub_rn(qy[k + 1], qy[0]);
    area = __fadd_rn(area, __fmaf_rn(ax, by, -__fmul_rn(ay, bx)));
  }
  return __fmul_rn(fabsf(area), 0.5f);
}

__device__ __forceinline__ float overlap_pair(const BoxPre &a, const BoxPre &b) {
  if (aabb_apart(a, b)) return 0.f;
  return overlap_exact(a, b);
}

// iou_bev (iou3d_kernel.cu:214-221): sa + sb is fma(sa factors, rn(sb)) in the reference build
__device__ __forceinline__ float iou_from_overlap(const BoxPre &a, const BoxPre &b, float s) {
  const float sb = __fmul_rn(__fsub_rn(b.x2, b.x1), __fsub_rn(b.y2, b.y1));
  const float sab = __fmaf_rn(__fsub_rn(a.x2, a.x1), __fsub_rn(a.y2, a.y1), sb);
  return __fdiv_rn(s, fmaxf(__fsub_rn(sab, s), kEps));
}

// iou_normal (iou3d_kernel.cu:295-303); a = row box (its area is the hoisted, rounded product)
__device__ __forceinline__ float iou_normal(const float *a, const float *b) {
  const float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  const float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  const float width = fmaxf(__fsub_rn(right, left), 0.f), height = fmaxf(__fsub_rn(bottom, top), 0.f);
  const float interS = __fmul_rn(width, height);
  const float Sa = __fmul_rn(__fsub_rn(a[2], a[0]), __fsub_rn(a[3], a[1]));
  const float Sab = __fmaf_rn(__fsub_rn(b[2], b[0]), __fsub_rn(b[3], b[1]), Sa);
  return __fdiv_rn(interS, fmaxf(__fsub_rn(Sab, interS), kEps));
}

// ---------------------------------------------------------------------------------------------
// all-pairs matrices: 64 x 64 tile per CTA, 256 threads, thread (r, cidx) covers 16 rows
constexpr int kTile = 64;

template <bool IOU>
__global__ void __launch_bounds__(256) pair_matrix_kernel(int num_a, const float *__restrict__ boxes_a, int num_b,
                                                           const float *__restrict__ boxes_b, float *__restrict__ ans) {
  __shared__ BoxPre sa[kTile], sb[kTile];
  __shared__ unsigned short s_queue[kTile * kTile];  // pairs that need the exact clipping: r * 64 + c
  __shared__ int s_count;
  const int a0 = blockIdx.y * kTile, b0 = blockIdx.x * kTile;
  const int na = min(kTile, num_a - a0), nb = min(kTile, num_b - b0);
  if (threadIdx.x == 0) s_count = 0;
  if (threadIdx.x < 2 * kTile) {
    const int t = threadIdx.x & (kTile - 1);
    if (threadIdx.x < kTile) {
      if (t < na) precompute(boxes_a + (size_t)(a0 + t) * 5, sa[t]);
    } else {
      if (t < nb) precompute(boxes_b + (size_t)(b0 + t) * 5, sb[t]);
    }
  }
  __syncthreads();
  // phase 1: AABB screen.  Apart => the result is exactly 0 (overlap and IoU alike): coalesced store.
  const int cb = threadIdx.x & (kTile - 1);
  if (cb < nb) {
    const BoxPre &bb = sb[cb];
    for (int r = threadIdx.x >> 6; r < na; r += 4) {
      if (aabb_apart(sa[r], bb)) __stcs(ans + (size_t)(a0 + r) * num_b + (b0 + cb), 0.f);
      else s_queue[atomicAdd(&s_count, 1)] = (unsigned short)(r * kTile + cb);
    }
  }
  __syncthreads();
  // phase 2: the surviving pairs, densely packed over the threads (no divergence against cheap pairs)
  const int count = s_count;
  for (int q = threadIdx.x; q < count; q += 256) {
    const int r = s_queue[q] >> 6, c = s_queue[q] & (kTile - 1);
    const BoxPre &aa = sa[r];
    const BoxPre &bb = sb[c];
    float s = overlap_exact(aa, bb);
    if (IOU) s = iou_from_overlap(aa, bb, s);
    ans[(size_t)(a0 + r) * num_b + (b0 + c)] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// NMS mask: upper-triangle 64x64 tiles only; 64 threads per tile, thread = row box.
// BEV centre distance exactly as lib/utils/distance.py:3 evaluates it in float32 torch kernels
// (subtract, square, two-term sum, sqrt: one rounding each).
__device__ __forceinline__ float bev_dist(float ax, float az, float bx, float bz) {
  const float dx = __fsub_rn(ax, bx), dz = __fsub_rn(az, bz);
  return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dz, dz)));
}

constexpr int kModeNormal = 0, kModeRotated = 1, kModeRadius = 2;

template <int MODE>
__global__ void __launch_bounds__(64) nms_mask_kernel(int n, float thresh, const float *__restrict__ boxes,
                                                      unsigned long long *__restrict__ mask) {
  const int col_blocks = ceil_div(n, 64);
  // linear upper-triangle tile id -> (row_block, col_block), row_block <= col_block
  int rb = 0;
  {
    long long t = blockIdx.x;
    // row rb owns (col_blocks - rb) tiles; solve by a short search (col_blocks <= a few thousand)
    int lo = 0, hi = col_blocks - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      const long long before = (long long)mid * col_blocks - (long long)mid * (mid - 1) / 2;
      if (before <= t) lo = mid; else hi = mid - 1;
    }
    rb = lo;
    t -= (long long)rb * col_blocks - (long long)rb * (rb - 1) / 2;
    // t is now the offset inside row rb
    const int cbk = rb + (int)t;
    const int row_size = min(64, n - rb * 64), col_size = min(64, n - cbk * 64);
    const int tid = threadIdx.x;
    if (MODE == kModeRotated) {
      __shared__ BoxPre scol[64], srow[64];
      __shared__ unsigned short s_queue[64 * 64];
      __shared__ unsigned int s_bits[64][2];
      __shared__ int s_count;
      if (tid == 0) s_count = 0;
      s_bits[tid][0] = 0u; s_bits[tid][1] = 0u;
      if (tid < col_size) precompute(boxes + (size_t)(cbk * 64 + tid) * 5, scol[tid]);
      if (tid < row_size) precompute(boxes + (size_t)(rb * 64 + tid) * 5, srow[tid]);
      __syncthreads();
      if (tid < row_size) {
        const BoxPre &me = srow[tid];
        const int start = (rb == cbk) ? tid + 1 : 0;
        if (thresh < 0.f) {
          // degenerate threshold: zero-overlap pairs are suppressed too, so nothing can be screened out
          for (int j = start; j < col_size; ++j) s_queue[atomicAdd(&s_count, 1)] = (unsigned short)(tid * 64 + j);
        } else {
          // AABB apart => overlap 0 => IoU 0 => not "> thresh"
          for (int j = start; j < col_size; ++j)
            if (!aabb_apart(me, scol[j])) s_queue[atomicAdd(&s_count, 1)] = (unsigned short)(tid * 64 + j);
        }
      }
      __syncthreads();
      const int count = s_count;
      for (int q = tid; q < count; q += 64) {
        const int r = s_queue[q] >> 6, j = s_queue[q] & 63;
        const float s = overlap_pair(srow[r], scol[j]);
        if (iou_from_overlap(srow[r], scol[j], s) > thresh) atomicOr(&s_bits[r][j >> 5], 1u << (j & 31));
      }
      __syncthreads();
      if (tid < row_size)
        mask[(size_t)(rb * 64 + tid) * col_blocks + cbk] =
            (unsigned long long)s_bits[tid][0] | ((unsigned long long)s_bits[tid][1] << 32);
    } else if (MODE == kModeRadius) {
      // "radius NMS" of tools/eval_auto.py:272-279: `boxes` = (n, 2) BEV centres, thresh = radius; a kept centre
      // suppresses every later one that is NOT farther than the radius (a NaN distance suppresses, as there)
      __shared__ float2 scol[64];
      if (tid < col_size) scol[tid] = reinterpret_cast<const float2 *>(boxes)[cbk * 64 + tid];
      __syncthreads();
      if (tid < row_size) {
        const float2 me = reinterpret_cast<const float2 *>(boxes)[rb * 64 + tid];
        unsigned long long bits = 0;
        const int start = (rb == cbk) ? tid + 1 : 0;
        for (int j = start; j < col_size; ++j)
          if (!(bev_dist(scol[j].x, scol[j].y, me.x, me.y) > thresh)) bits |= 1ULL << j;
        mask[(size_t)(rb * 64 + tid) * col_blocks + cbk] = bits;
      }
    } else {
      __shared__ float scol[64 * 5];
      if (tid < col_size)
        for (int q = 0; q < 5; ++q) scol[tid * 5 + q] = boxes[(size_t)(cbk * 64 + tid) * 5 + q];
      __syncthreads();
      if (tid < row_size) {
        float me[5];
        for (int q = 0; q < 5; ++q) me[q] = boxes[(size_t)(rb * 64 + tid) * 5 + q];
        unsigned long long bits = 0;
        const int start = (rb == cbk) ? tid + 1 : 0;
        for (int j = start; j < col_size; ++j)
          if (iou_normal(me, scol + j * 5) > thresh) bits |= 1ULL << j;
        mask[(size_t)(rb * 64 + tid) * col_blocks + cbk] = bits;
      }
    }
  }
}

// Greedy suppression over the device mask (iou3d.cpp:101-116), one CTA.
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads) nms_scan_kernel(int n, const unsigned long long *__restrict__ mask,
                                                                 long long *__restrict__ keep, int *__restrict__ num_keep,
                                                                 unsigned long long *__restrict__ remv_g) {
  extern __shared__ unsigned long long s_remv[];  // col_blocks words, or 0 when spilled to global
  __shared__ unsigned long long s_diag[64];
  __shared__ unsigned long long s_keepbits;
  __shared__ int s_total;
  const int col_blocks = ceil_div(n, 64);
  unsigned long long *remv = remv_g ? remv_g : s_remv;
  for (int j = threadIdx.x; j < col_blocks; j += kScanThreads) remv[j] = 0;
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  for (int c = 0; c < col_blocks; ++c) {
    const int rows = min(64, n - c * 64);
    if (threadIdx.x < 64)
      s_diag[threadIdx.x] = threadIdx.x < rows ? mask[(size_t)(c * 64 + threadIdx.x) * col_blocks + c] : 0ULL;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long rem = remv[c], kb = 0;
#pragma unroll 8
      for (int r = 0; r < 64; ++r) {
        const unsigned long long d = s_diag[r];  // rows beyond `rows` hold 0 and are masked below
        if (!((rem >> r) & 1ULL)) { kb |= 1ULL << r; rem |= d; }
      }
      if (rows < 64) kb &= (1ULL << rows) - 1ULL;
      s_keepbits = kb;
    }
    __syncthreads();
    const unsigned long long kb = s_keepbits;
    const int base = s_total;
    if (threadIdx.x < 64 && ((kb >> threadIdx.x) & 1ULL))
      keep[base + __popcll(kb & ((1ULL << threadIdx.x) - 1ULL))] = (long long)c * 64 + threadIdx.x;
    // OR the kept rows into the removal words of the later column blocks: 16 threads per row, all
    // loads independent (the mask lives in L2), merged with 32-bit shared/global atomics
    {
      const int r = threadIdx.x >> 4, sub = threadIdx.x & 15;
      if ((kb >> r) & 1ULL) {
        const unsigned long long *row = mask + (size_t)(c * 64 + r) * col_blocks;
        unsigned int *remv32 = reinterpret_cast<unsigned int *>(remv);
#pragma unroll 4
        for (int j = c + 1 + sub; j < col_blocks; j += 16) {
          const unsigned long long v = row[j];
          if ((unsigned int)v) atomicOr(remv32 + 2 * j, (unsigned int)v);
          if ((unsigned int)(v >> 32)) atomicOr(remv32 + 2 * j + 1, (unsigned int)(v >> 32));
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_total = base + __popcll(kb);
  }
  __syncthreads();
  if (threadIdx.x == 0) *num_keep = s_total;
}

// ---------------------------------------------------------------------------------------------
// Aligned pairs: the DIAGONAL of boxes_iou3d_gpu (lib/utils/iou3d/iou3d_utils.py:21-56),
// which is all that the Stage-2 losses keep of their fg x fg matrices (lib/net/train_functions.py:258-260,:287-289).
// One thread per pair: BEV conversion (kitti_utils.py:134-147), the exact rotated overlap above, then the height /
// area / volume arithmetic of :35-53 with one IEEE operation per torch elementwise kernel.
__device__ __forceinline__ float clamp_min(float v, float lo) { return v < lo ? lo : v; }  // torch.clamp(min=): NaN stays NaN

__device__ __forceinline__ void box3d_to_bev(const float *__restrict__ b, float *o) {
  const float half_l = __fmul_rn(b[5], 0.5f), half_w = __fmul_rn(b[4], 0.5f);   // x / 2 is exact
  o[0] = __fsub_rn(b[0], half_l); o[1] = __fsub_rn(b[2], half_w);
  o[2] = __fadd_rn(b[0], half_l); o[3] = __fadd_rn(b[2], half_w);
  o[4] = b[6];
}

__global__ void __launch_bounds__(128) iou3d_aligned_kernel(int n, const float *__restrict__ boxes_a,
                                                             const float *__restrict__ boxes_b, float *__restrict__ iou2d,
                                                             float *__restrict__ iou3d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a[7], b[7], abev[5], bbev[5];
#pragma unroll
  for (int q = 0; q < 7; ++q) { a[q] = boxes_a[(size_t)i * 7 + q]; b[q] = boxes_b[(size_t)i * 7 + q]; }
  box3d_to_bev(a, abev);
  box3d_to_bev(b, bbev);
  BoxPre pa, pb;
  precompute(abev, pa);
  precompute(bbev, pb);
  const float ov = overlap_pair(pa, pb);
  const float a_hmin = __fsub_rn(a[1], a[3]), b_hmin = __fsub_rn(b[1], b[3]);
  const float max_of_min = (a_hmin != a_hmin || b_hmin != b_hmin) ? __fadd_rn(a_hmin, b_hmin) : fmaxf(a_hmin, b_hmin);
  const float min_of_max = (a[1] != a[1] || b[1] != b[1]) ? __fadd_rn(a[1], b[1]) : fminf(a[1], b[1]);  // torch.max/min propagate NaN
  const float ov_h = clamp_min(__fsub_rn(min_of_max, max_of_min), 0.f);
  const float s_a = __fmul_rn(a[4], a[5]), s_b = __fmul_rn(b[4], b[5]);
  if (iou2d) iou2d[i] = __fdiv_rn(ov, clamp_min(__fsub_rn(__fadd_rn(s_a, s_b), ov), 1e-7f));
  const float ov3 = __fmul_rn(ov, ov_h);
  const float vol_a = __fmul_rn(__fmul_rn(a[3], a[4]), a[5]), vol_b = __fmul_rn(__fmul_rn(b[3], b[4]), b[5]);
  if (iou3d) iou3d[i] = __fdiv_rn(ov3, clamp_min(__fsub_rn(__fadd_rn(vol_a, vol_b), ov3), 1e-7f));
}

int matrix_dispatch(bool iou, int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *ans,
                    cudaStream_t stream) {
  const char *what = iou ? "boxes_iou_bev" : "boxes_overlap_bev";
  if (num_a < 0 || num_b < 0) return fail_arg(what);
  if (num_a == 0 || num_b == 0) return 0;
  if (!boxes_a || !boxes_b || !ans) return fail_arg(what);
  dim3 grid((unsigned)ceil_div(num_b, kTile), (unsigned)ceil_div(num_a, kTile));
  if (grid.y > 65535) return fail_arg(what);
  if (iou) pair_matrix_kernel<true><<<grid, 256, 0, stream>>>(num_a, boxes_a, num_b, boxes_b, ans);
  else pair_matrix_kernel<false><<<grid, 256, 0, stream>>>(num_a, boxes_a, num_b, boxes_b, ans);
  return check_launch(what);
}

size_t nms_ws_bytes(int n) {
  const size_t cb = (size_t)ceil_div(n, 64);
  return (size_t)n * cb * 8 + cb * 8 + 256;
}

int nms_dispatch(int mode, const float *boxes, int n, float thresh, int64_t *keep, int *num_keep, void *workspace,
                 cudaStream_t stream) {
  const char *what = mode == kModeRotated ? "nms" : mode == kModeRadius ? "radius_nms" : "nms_normal";
  if (n < 0) return fail_arg(what);
  if (!num_keep) return fail_arg(what);
  if (n == 0) {
    cudaError_t e = cudaMemsetAsync(num_keep, 0, sizeof(int), stream);
    if (e != cudaSuccess) { set_error("%s: memset: %s", what, cudaGetErrorString(e)); return (int)e; }
    return 0;
  }
  if (!boxes || !keep) return fail_arg(what);
  if (!workspace) workspace = scratch(nms_ws_bytes(n), 2);
  if (!workspace) return (int)cudaErrorMemoryAllocation;
  const int cb = ceil_div(n, 64);
  unsigned long long *mask = reinterpret_cast<unsigned long long *>(workspace);
  unsigned long long *remv_g = mask + (size_t)n * cb;
  const long long tiles = (long long)cb * (cb + 1) / 2;
  if (tiles > 2147483647LL) return fail_arg(what);
  if (mode == kModeRotated) nms_mask_kernel<kModeRotated><<<(unsigned)tiles, 64, 0, stream>>>(n, thresh, boxes, mask);
  else if (mode == kModeRadius) nms_mask_kernel<kModeRadius><<<(unsigned)tiles, 64, 0, stream>>>(n, thresh, boxes, mask);
  else nms_mask_kernel<kModeNormal><<<(unsigned)tiles, 64, 0, stream>>>(n, thresh, boxes, mask);
  int rc = check_launch(what);
  if (rc) return rc;
  const size_t smem = (size_t)cb * 8;
  const bool spill = smem > 40 * 1024;
  nms_scan_kernel<<<1, kScanThreads, spill ? 0 : smem, stream>>>(n, mask, reinterpret_cast<long long *>(keep), num_keep,
                                                                 spill ? remv_g : nullptr);
  return check_launch("nms scan");
}

int nms_host(bool rotated, const float *boxes, int n, float thresh, int64_t *keep_host, cudaStream_t stream) {
  if (n < 0 || (n > 0 && (!boxes || !keep_host))) return -fail_arg("nms_host");
  if (n == 0) return 0;
  const size_t out_bytes = (size_t)n * 8 + 64;
  char *dev = (char *)scratch(out_bytes, 3);
  if (!dev) return -(int)cudaErrorMemoryAllocation;
  int64_t *keep_dev = reinterpret_cast<int64_t *>(dev + 64);
  int *num_dev = reinterpret_cast<int *>(dev);
  int rc = nms_dispatch(rotated ? kModeRotated : kModeNormal, boxes, n, thresh, keep_dev, num_dev, nullptr, stream);
  if (rc) return -rc;
  int num = 0;
  cudaError_t e = cudaMemcpyAsync(&num, num_dev, sizeof(int), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e == cudaSuccess && num > 0) {
    e = cudaMemcpyAsync(keep_host, keep_dev, (size_t)num * 8, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  }
  if (e != cudaSuccess) { set_error("nms_host: %s", cudaGetErrorString(e)); return -(int)e; }
  return num;
}

}  // namespace
}  // namespace ws3d

using namespace ws3d;

WS3D_API int ws3d_boxes_overlap_bev(int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *ans,
                                    ws3d_stream_t stream) {
  return matrix_dispatch(false, num_a, boxes_a, num_b, boxes_b, ans, to_stream(stream));
}
WS3D_API int ws3d_boxes_iou_bev(int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *ans,
                                ws3d_stream_t stream) {
  return matrix_dispatch(true, num_a, boxes_a, num_b, boxes_b, ans, to_stream(stream));
}
WS3D_API size_t ws3d_nms_workspace_bytes(int boxes_num) { return boxes_num > 0 ? nms_ws_bytes(boxes_num) : 0; }
WS3D_API int ws3d_nms(const float *boxes, int boxes_num, float thresh, int64_t *keep, int *num_keep, void *workspace,
                      ws3d_stream_t stream) {
  return nms_dispatch(kModeRotated, boxes, boxes_num, thresh, keep, num_keep, workspace, to_stream(stream));
}
WS3D_API int ws3d_nms_normal(const float *boxes, int boxes_num, float thresh, int64_t *keep, int *num_keep,
                             void *workspace, ws3d_stream_t stream) {
  return nms_dispatch(kModeNormal, boxes, boxes_num, thresh, keep, num_keep, workspace, to_stream(stream));
}
WS3D_API int ws3d_nms_host(const float *boxes, int boxes_num, float thresh, int64_t *keep_host, ws3d_stream_t stream) {
  return nms_host(true, boxes, boxes_num, thresh, keep_host, to_stream(stream));
}
WS3D_API int ws3d_nms_normal_host(const float *boxes, int boxes_num, float thresh, int64_t *keep_host,
                                  ws3d_stream_t stream) {
  return nms_host(false, boxes, boxes_num, thresh, keep_host, to_stream(stream));
}

WS3D_API int ws3d_boxes_iou3d_aligned(int n, const float *boxes_a, const float *boxes_b, float *iou2d, float *iou3d,
                                      ws3d_stream_t stream) {
  const char *what = "boxes_iou3d_aligned";
  if (n < 0) return fail_arg(what);
  if (n == 0) return 0;
  if (!boxes_a || !boxes_b || (!iou2d && !iou3d)) return fail_arg(what);
  iou3d_aligned_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, to_stream(stream)>>>(n, boxes_a, boxes_b, iou2d, iou3d);
  return check_launch(what);
}
WS3D_API int ws3d_radius_nms(const float *centers, int n, float radius, int64_t *keep, int *num_keep, void *workspace,
                             ws3d_stream_t stream) {
  return nms_dispatch(kModeRadius, centers, n, radius, keep, num_keep, workspace, to_stream(stream));
}
