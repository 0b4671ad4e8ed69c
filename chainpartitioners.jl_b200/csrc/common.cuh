// common.cuh -- shared infrastructure of libchainb200 (sm_100a): error handling, the library
// context (device, stream, stream-ordered memory pool), RAII device buffers, launch accounting and
// the CUDA-event phase profiler used by bench.py.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace cpb {

using i64 = long long;
using u32 = uint32_t;
using u64 = unsigned long long;

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CPB_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      throw ::cpb::Error(-3, std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

#define CPB_REQUIRE(cond, msg)                        \
  do {                                                \
    if (!(cond)) throw ::cpb::Error(-1, (msg));       \
  } while (0)

struct ProfEntry {
  double ms = 0;
  i64 launches = 0;
  double bytes = 0;
};

struct Context {
  bool ready = false;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  i64 launches = 0;
  bool profiling = false;
  std::vector<std::pair<std::string, ProfEntry>> prof;
  ProfEntry& prof_entry(const std::string& name) {
    for (auto& p : prof)
      if (p.first == name) return p.second;
    prof.push_back({name, ProfEntry()});
    return prof.back().second;
  }
};

Context& ctx();          // capi.cu
void ensure_context();   // throws Error(-3) when no usable device

// Every kernel launch goes through this macro: counts launches (gpu_launches in bench.py) and
// surfaces launch-configuration errors immediately.
#define CPB_LAUNCH(kernel, grid, block, smem, ...)                                 \
  do {                                                                             \
    kernel<<<(grid), (block), (smem), ::cpb::ctx().stream>>>(__VA_ARGS__);         \
    ::cpb::ctx().launches += 1;                                                    \
    CPB_CUDA(cudaPeekAtLastError());                                               \
  } while (0)

// Large buffers (>= 64 MiB) bypass the driver's stream-ordered pool: at the GB sizes of the big configurations the pool
// occasionally re-stitches its virtual ranges inside cudaMallocAsync, which showed up as 50-700 ms stalls on the device
// timeline of a 37 ms solve.  The library keeps its own free list of exact-size blocks instead (every solve asks for the
// same sizes again); all work runs on the one library stream, so handing a freed block to the next request is ordered.
static constexpr size_t CPB_BIG_BYTES = (size_t)64 << 20;
void* big_alloc(size_t bytes);           // capi.cu
void big_free(void* p, size_t bytes);    // capi.cu
void big_trim();                         // returns every cached block to the driver

// Stream-ordered device buffer (cudaMallocAsync on the library stream; large blocks from the library's own free list).
template <class T> struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  DBuf() {}
  explicit DBuf(size_t count) { alloc(count); }
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DBuf& operator=(DBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    const size_t bytes = (count ? count : 1) * sizeof(T);
    if (bytes >= CPB_BIG_BYTES) p = (T*)big_alloc(bytes);
    else CPB_CUDA(cudaMallocAsync((void**)&p, bytes, ctx().stream));
  }
  void zero() { if (p) CPB_CUDA(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), ctx().stream)); }
  void release() {
    if (p) {
      const size_t bytes = (n ? n : 1) * sizeof(T);
      if (bytes >= CPB_BIG_BYTES) big_free(p, bytes); else cudaFreeAsync(p, ctx().stream);
      p = nullptr;
      n = 0;
    }
  }
  T* get() const { return p; }
  operator T*() const { return p; }
};

// RAII phase timer: CUDA events on the library stream, resolved lazily at cpb_profile_get().
struct ProfScope {
  std::string name;
  double bytes;
  i64 launches0;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  bool on;
  ProfScope(const char* nm, double algorithmic_bytes = 0);
  ~ProfScope();
};

// CPB_TRACE=1: host wall-clock marks (each after a stream synchronise) printed to stderr by the next cpb_* entry point that
// flushes them -- for attributing time that no kernel scope accounts for.  Off: one predictable branch.
void trace_mark(const char* label);  // capi.cu
void trace_flush(const char* what);

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace cpb
