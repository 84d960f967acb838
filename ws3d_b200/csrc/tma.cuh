// mbarrier + 1-D TMA bulk copy (cp.async.bulk) helpers shared by the staging kernels.
#pragma once
#include "common.cuh"

namespace ws3d {

// TMA bulk copy helpers (1-D cp.async.bulk global -> shared, completion on an mbarrier)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes or ~the hint (ns) has passed, so a
// waiter comes back once instead of several times (the default time limit is short: ncu counted four try_wait rounds per wait and
// 6 % of the fused SA kernel's executed instructions in the polling loop, most of them the clock reads of the watchdog)
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  unsigned spins = 0;
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {
    // never hang the GPU on a lost copy: the watchdog reads the clock every 64th round only
    if ((++spins & 63u) == 0u && clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// Stage `count` packed floats (16-byte multiple when the TMA path is taken) into shared memory.
// All threads of the CTA call this; on return the data is visible to all of them.
__device__ __forceinline__ void stage_floats(float *s_dst, const float *g_src, int count, unsigned long long *s_bar,
                                             uint32_t parity) {
  const bool tma_ok = ((reinterpret_cast<uintptr_t>(g_src) & 15u) == 0) && count >= 4;
  const int vec = tma_ok ? (count & ~3) : 0;  // floats moved by TMA
  if (tma_ok) {
    if (threadIdx.x == 0) {
      const uint32_t bar = smem_u32(s_bar);
      mbar_expect_tx(bar, (uint32_t)vec * 4u);
      const char *src = reinterpret_cast<const char *>(g_src);
      uint32_t dst = smem_u32(s_dst);
      uint32_t left = (uint32_t)vec * 4u;
      while (left) {
        const uint32_t step = left > 65536u ? 65536u : left;
        tma_bulk_g2s(dst, src, step, bar);
        dst += step; src += step; left -= step;
      }
    }
  }
  for (int i = vec + (int)threadIdx.x; i < count; i += (int)blockDim.x) s_dst[i] = __ldg(g_src + i);
  if (tma_ok) mbar_wait(smem_u32(s_bar), parity);
  __syncthreads();
}


}  // namespace ws3d
