// Shared helpers for the stgcn_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <mutex>
#include <cstdarg>
#include <cstdlib>
#include <utility>
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace stgcn {

// thread-local last-error message (the library is re-entrant across devices/streams)
inline char *err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
inline int fail(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return 1;
}

#define STGCN_CUDA_OK(expr)                                                               \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return ::stgcn::fail("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
  } while (0)

// ---- launch accounting / optional per-kernel-class event timing ---------------------
// Classes used by bench.py's roofline leg (stgcn_profile_*).
enum KernelClass {
  KC_LAYOUT = 0,   // NCTV <-> NTVC transposes
  KC_GEMM_1X1,     // feature transform / residual 1x1 GEMMs
  KC_GEMM_TCN,     // Gamma x 1 temporal implicit GEMM
  KC_FRAME,        // adjacency + norm + residual epilogue kernels
  KC_BN,           // batch-statistics kernels
  KC_EMBED,        // norm_in + fcn_in
  KC_POOL,         // pooling + classifier
  KC_MISC,         // weight packing, CSR build, counters
  KC_COUNT
};

// One profiler per device: replicas run one host thread per GPU, each thread brackets and reads
// the launches of its own device; slots are handed out under the profiler's mutex, so several threads may
// also drive one device.  The launch counter and the "any profiler on" flag are process-wide atomics, so a
// launch with profiling off costs one relaxed load and one add (no cudaGetDevice on the launch path).
struct Profiler {
  std::atomic<bool> on{false};
  std::mutex mu;
  static constexpr int kMax = 4096;
  cudaEvent_t ev[kMax][2];
  int cls[kMax];
  int n = 0;       // events in flight
  int created = 0; // events created so far
  float ms[KC_COUNT] = {0};
  long long count[KC_COUNT] = {0};
};
struct ProfGlobal {
  std::atomic<long long> launches{0};
  std::atomic<int> active{0};      // devices with profiling on
};
inline ProfGlobal &prof_global() {
  static ProfGlobal g;
  return g;
}
constexpr int kMaxDevices = 32;
inline Profiler *prof_table() {
  static Profiler p[kMaxDevices];
  return p;
}
inline Profiler &prof() {
  int dev = 0;
  cudaGetDevice(&dev);
  return prof_table()[dev >= 0 && dev < kMaxDevices ? dev : 0];
}
// RAII: brackets one kernel launch with events when profiling is on.
struct ProfScope {
  cudaStream_t st;
  cudaEvent_t stop = nullptr;
  ProfScope(int cls, cudaStream_t s) : st(s) {
    if (prof_global().active.load(std::memory_order_relaxed) == 0) return;
    Profiler &p = prof();
    if (!p.on.load(std::memory_order_relaxed)) return;
    cudaEvent_t start;
    {
      std::lock_guard<std::mutex> g(p.mu);
      if (p.n >= Profiler::kMax) return;
      const int slot = p.n++;
      if (slot >= p.created) {
        cudaEventCreate(&p.ev[slot][0]);
        cudaEventCreate(&p.ev[slot][1]);
        p.created = slot + 1;
      }
      p.cls[slot] = cls;
      start = p.ev[slot][0];
      stop = p.ev[slot][1];
    }
    cudaEventRecord(start, st);
  }
  ~ProfScope() {
    if (stop) cudaEventRecord(stop, st);
  }
};

#define STGCN_LAUNCH_OK()                                                                 \
  do {                                                                                    \
    ::stgcn::prof_global().launches.fetch_add(1, std::memory_order_relaxed);              \
    cudaError_t e_ = cudaGetLastError();                                                  \
    if (e_ != cudaSuccess)                                                                \
      return ::stgcn::fail("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
  } while (0)

#define STGCN_REQUIRE(cond, ...)                                                          \
  do {                                                                                    \
    if (!(cond)) return ::stgcn::fail(__VA_ARGS__);                                       \
  } while (0)

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------
// A kernel launched with launch_pdl may be scheduled while the previous kernel of the stream is still running: its
// CTAs become resident and run their prologue (barrier init, TMEM allocation, static tables, weight prefetch), and
// block in griddep_wait() until the previous kernel has completed and its writes are visible.  Rules kept by every
// kernel that is launched this way: (1) no global access to data another kernel of the chain writes before
// griddep_wait(); (2) griddep_launch() only AFTER griddep_wait(), so at most two kernels of a chain are in flight
// and code before a wait can only race with the immediate predecessor.  Launched without the attribute (or with
// STGCN_PDL=0) both instructions are no-ops.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("STGCN_PDL");
    on = e ? atoi(e) != 0 : 1;
  }
  return on != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Bump allocator over a caller-owned workspace.  With base == nullptr it only
// measures, so *_workspace_bytes() and the forward share one code path.
struct Bump {
  char *base;
  size_t cap;
  size_t off = 0;
  size_t peak = 0;
  bool overflow = false;
  Bump(void *b, size_t c) : base(static_cast<char *>(b)), cap(c) {}
  template <typename T>
  T *take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    size_t at = off;
    off += bytes;
    if (off > peak) peak = off;
    if (!base) return nullptr;
    if (off > cap) {
      overflow = true;
      return nullptr;
    }
    return reinterpret_cast<T *>(base + at);
  }
  bool measuring() const { return base == nullptr; }
  size_t mark() const { return off; }
  void release(size_t m) { off = m; }
};

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace stgcn
