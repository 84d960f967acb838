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
// cached per-device scratch (grown on demand; outgrown buffers are retired until ws3d_release_scratch());
// nullptr + error on failure
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

// SM count of the current device (cudaDevAttrMultiProcessorCount, cached per device; 148 on a full B200, fewer on
// a MIG slice or an SM-limited context).  Grid sizes of persistent kernels and the FPS decomposition derive from it.
int num_sms();

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

// Output stage shared by the two three_nn kernels: squared distances and indices as the reference writes them
// (interpolate_gpu.cu:49-51) and / or the normalised inverse-distance weights PointnetFPModule.forward derives from
// them with five elementwise torch kernels (pointnet2_modules.py:139-144): dist = sqrt(d2); r = 1 / (dist + 1e-8);
// w = r / (r0 + r1 + r2) -- every step one IEEE round-to-nearest float32 operation, as torch evaluates it.
__device__ __forceinline__ void store_three_nn(float *__restrict__ dist2, int *__restrict__ idx, float *__restrict__ weight,
                                               size_t row, float d1, float d2, float d3, int i1, int i2, int i3) {
  int *oi = idx + row * 3;
  oi[0] = i1; oi[1] = i2; oi[2] = i3;
  if (dist2) {
    float *od = dist2 + row * 3;
    od[0] = d1; od[1] = d2; od[2] = d3;
  }
  if (weight) {
    const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(d1), 1e-8f));
    const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(d2), 1e-8f));
    const float r3 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(d3), 1e-8f));
    const float norm = __fadd_rn(__fadd_rn(r1, r2), r3);
    float *ow = weight + row * 3;
    ow[0] = __fdiv_rn(r1, norm); ow[1] = __fdiv_rn(r2, norm); ow[2] = __fdiv_rn(r3, norm);
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

}  // namespace ws3d
