// Shared helpers for libws3d_ops (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ws3d_ops.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libws3d_ops is written for sm_100a (B200) only"
#endif

#define WS3D_API extern "C" __attribute__((visibility("default")))

namespace ws3d {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
// cached per-device scratch (grown on demand, never shrunk); nullptr + error on failure
void *scratch(size_t bytes, int slot);
// grid size of a persistent kernel with `per_sm` CTAs per SM, honouring ws3d_set_sm_budget()
int persistent_ctas(int per_sm);
// ws3d_set_fps_mode() of the calling thread: 0 auto, 1 throughput (one SM per cloud), 2 latency (clusters)
int fps_mode();

inline cudaStream_t to_stream(ws3d_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Checks the launch that has just been issued; returns 0 or the cudaError code.
inline int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  count_launch();
  return 0;
}

inline int fail_arg(const char *what) {
  set_error("%s: invalid argument", what);
  return (int)cudaErrorInvalidValue;
}

constexpr int kNumSMs = 148;  // B200

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Squared distance with the rounding order the reference kernels compile to
// (SASS of ball_query/three_nn/FPS: t = rn(dy*dy); t = fma(dx,dx,t); d = fma(dz,dz,t)).
__device__ __forceinline__ float sqdist_ref(float dx, float dy, float dz) {
  float t = __fmul_rn(dy, dy);
  t = __fmaf_rn(dx, dx, t);
  return __fmaf_rn(dz, dz, t);
}

// Two squared distances at once with the packed FP32 pipe (FADD2 / FMUL2 / FFMA2 on sm_100): every lane of a
// packed operation is an IEEE round-to-nearest operation, so each result is bit-identical to sqdist_ref.
__device__ __forceinline__ void sqdist_ref_x2(float x0, float x1, float y0, float y1, float z0, float z1, float cx, float cy,
                                              float cz, float &d0, float &d1) {
  unsigned long long X, Y, Z, CX, CY, CZ, T;
  asm("mov.b64 %0, {%1, %2};" : "=l"(X) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(Y) : "f"(y0), "f"(y1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(Z) : "f"(z0), "f"(z1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(CX) : "f"(cx));
  asm("mov.b64 %0, {%1, %1};" : "=l"(CY) : "f"(cy));
  asm("mov.b64 %0, {%1, %1};" : "=l"(CZ) : "f"(cz));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(X) : "l"(X), "l"(CX));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(Y) : "l"(Y), "l"(CY));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(Z) : "l"(Z), "l"(CZ));
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(T) : "l"(Y));
  asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(T) : "l"(X), "l"(T));
  asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(T) : "l"(Z), "l"(T));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(T));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

}  // namespace ws3d
