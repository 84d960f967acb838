// Library-wide state: error text, launch counter, cached device scratch.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "fps_common.cuh"

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
constexpr int kMaxDev = 16, kSlots = 12, kArenas = 8;
struct Slot { void *p = nullptr; size_t cap = 0; };
Slot g_slots[kMaxDev][kArenas][kSlots];
std::vector<Slot> g_retired[kMaxDev];   // outgrown buffers: captured graphs / queued launches may still hold them
std::mutex g_mu;
std::atomic<int> g_sms[kMaxDev];        // 0 = not queried yet
// Which set of cached scratch buffers the calling thread's launches use.  Two forward passes that are in
// flight at the same time (two CUDA graphs replayed on different streams) must not share cell grids.
thread_local int g_arena = 0;
std::atomic<int> g_sm_budget{0};
thread_local int g_fps_mode = 0;
}  // namespace

int fps_mode() { return g_fps_mode; }

int num_sms() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return 148;
  int v = g_sms[dev].load(std::memory_order_relaxed);
  if (v <= 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    g_sms[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

int persistent_ctas(int per_sm) {
  const int b = g_sm_budget.load(std::memory_order_relaxed), sms = num_sms();
  return (b > 0 && b < sms ? b : sms) * per_sm;
}

void *scratch(size_t bytes, int slot) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev || slot < 0 || slot >= kSlots) {
    set_error("scratch: bad device/slot");
    return nullptr;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  Slot &s = g_slots[dev][g_arena][slot];
  if (s.cap < bytes) {
    // A buffer that is outgrown is RETIRED, not freed: launches already queued and -- more importantly -- captured
    // CUDA graphs (graphs.py) hold its address for as long as they live.  Buffers grow geometrically, so the retired
    // ones add up to less than the live one; ws3d_release_scratch() frees them once the caller knows nothing refers
    // to them any more.
    const size_t old_cap = s.cap;
    if (s.p) g_retired[dev].push_back(s);
    s.p = nullptr;
    s.cap = 0;
    size_t want = bytes + (bytes >> 2);
    if (want < 2 * old_cap) want = 2 * old_cap;
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

WS3D_API int ws3d_set_workspace_arena(int arena) {
  const int prev = ws3d::g_arena;
  if (arena >= 0 && arena < ws3d::kArenas) ws3d::g_arena = arena;
  return prev;
}
WS3D_API size_t ws3d_scratch_bytes(int retired_only) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= ws3d::kMaxDev) return 0;
  std::lock_guard<std::mutex> lk(ws3d::g_mu);
  size_t total = 0;
  for (const auto &r : ws3d::g_retired[dev]) total += r.cap;
  if (!retired_only)
    for (int a = 0; a < ws3d::kArenas; ++a)
      for (int k = 0; k < ws3d::kSlots; ++k) total += ws3d::g_slots[dev][a][k].cap;
  return total;
}
WS3D_API int ws3d_release_scratch(int all) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= ws3d::kMaxDev) { ws3d::set_error("release_scratch: bad device"); return (int)cudaErrorInvalidDevice; }
  e = cudaDeviceSynchronize();   // nothing queued may still be using a buffer that is about to go
  if (e != cudaSuccess) { ws3d::set_error("release_scratch: %s", cudaGetErrorString(e)); return (int)e; }
  std::lock_guard<std::mutex> lk(ws3d::g_mu);
  for (auto &r : ws3d::g_retired[dev]) cudaFree(r.p);
  ws3d::g_retired[dev].clear();
  if (all)
    for (int a = 0; a < ws3d::kArenas; ++a)
      for (int k = 0; k < ws3d::kSlots; ++k) {
        ws3d::Slot &s = ws3d::g_slots[dev][a][k];
        if (s.p) cudaFree(s.p);
        s.p = nullptr;
        s.cap = 0;
      }
  return 0;
}
WS3D_API int ws3d_num_arenas(void) { return ws3d::kArenas; }
WS3D_API int ws3d_set_sm_budget(int sms) { return ws3d::g_sm_budget.exchange(sms < 0 ? 0 : sms); }
WS3D_API int ws3d_fps_clouds_per_cta(int b, int n) { return ws3d::fps_smem_clouds_per_cta(b, n); }
WS3D_API int ws3d_set_fps_mode(int mode) {
  const int prev = ws3d::g_fps_mode;
  if (mode >= 0 && mode <= 2) ws3d::g_fps_mode = mode;
  return prev;
}
