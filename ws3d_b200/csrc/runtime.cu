// Library-wide state: error text, launch counter, cached device scratch.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace ws3d {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

namespace {
constexpr int kMaxDev = 16, kSlots = 8;
struct Slot { void *p = nullptr; size_t cap = 0; };
Slot g_slots[kMaxDev][kSlots];
std::mutex g_mu;
}  // namespace

void *scratch(size_t bytes, int slot) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev || slot < 0 || slot >= kSlots) {
    set_error("scratch: bad device/slot");
    return nullptr;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  Slot &s = g_slots[dev][slot];
  if (s.cap < bytes) {
    if (s.p) {
      cudaDeviceSynchronize();  // earlier work may still be using the old buffer
      cudaFree(s.p);
      s.p = nullptr;
      s.cap = 0;
    }
    size_t want = bytes + (bytes >> 2);
    cudaError_t e = cudaMalloc(&s.p, want);
    if (e != cudaSuccess) {
      set_error("scratch: cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
      s.p = nullptr;
      return nullptr;
    }
    s.cap = want;
  }
  return s.p;
}

}  // namespace ws3d

WS3D_API int ws3d_abi_version(void) { return WS3D_ABI_VERSION; }
WS3D_API const char *ws3d_last_error(void) { return ws3d::g_err; }
WS3D_API uint64_t ws3d_launch_count(void) { return ws3d::g_launches.load(); }
