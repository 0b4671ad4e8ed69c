// hostpack.cpp -- host side of the pageable-memory upload (plain C++, compiled by g++: target_clones + a worker pool).
//
// A Julia Array is pageable memory.  The driver would stage it through one internal buffer on one thread; instead a small
// persistent pool of host threads copies it chunk by chunk into pinned staging buffers while the previous chunks are in
// flight over PCIe -- and, for Int64 index arrays, NARROWS while it copies: the device keeps 32-bit indices anyway
// (engine.cuh), so only the low halves cross the bus (half the PCIe bytes, half the staging writes; the high halves are
// OR-ed together so that an index >= 2^32 is still reported).
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace cpb {

// dst[i] = low 32 bits of src[i]; returns the OR of all src[i] (the caller tests the high half)
__attribute__((target_clones("avx512f", "avx2", "default"))) uint64_t pack_i64_u32(const int64_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                                                   size_t n) {
  uint64_t acc = 0;
  for (size_t i = 0; i < n; ++i) {
    const uint64_t v = (uint64_t)src[i];
    acc |= v;
    dst[i] = (uint32_t)v;
  }
  return acc;
}

// A fixed pool of worker threads that run `fn(part, parts)` for part = 0..parts-1 and return; the calling thread takes
// part 0.  Workers sleep on a condition variable between uploads (no spinning while the library is idle).
class HostPool {
 public:
  explicit HostPool(int workers) {
    for (int t = 0; t < workers; ++t) threads_.emplace_back([this, t] { loop(t + 1); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++epoch_;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  int parts() const { return (int)threads_.size() + 1; }
  void run(const std::function<void(int, int)>& fn) {
    const int P = parts();
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      pending_ = P - 1;
      ++epoch_;
    }
    cv_.notify_all();
    fn(0, P);
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void loop(int part) {
    uint64_t seen = 0;
    while (true) {
      const std::function<void(int, int)>* fn = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (stop_) return;
        fn = fn_;
      }
      if (fn) (*fn)(part, parts());
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_cv_.notify_one();
      }
    }
  }
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(int, int)>* fn_ = nullptr;
  int pending_ = 0;
  uint64_t epoch_ = 0;
  bool stop_ = false;
};

static int g_pool_share = 1;   // processes of this library sharing the host (set by cpb_comm_init before the first upload)
static HostPool* g_pool = nullptr;
static int g_pool_threads = 0;
static int wanted_threads() {
  const unsigned hc = std::max(1u, std::thread::hardware_concurrency());
  const char* env = std::getenv("CPB_H2D_THREADS");
  int T = env ? std::atoi(env) : (int)std::min(12u, std::max(2u, ((hc * 3) / 4) / (unsigned)std::max(g_pool_share, 1)));
  return T < 1 ? 1 : T;
}
static HostPool& pool() {
  const int T = wanted_threads();
  if (!g_pool || g_pool_threads != T) {  // (re)built between uploads only: run() is never in flight here
    delete g_pool;
    g_pool = new HostPool(T - 1);  // lives for the process
    g_pool_threads = T;
  }
  return *g_pool;
}
void host_pool_share(int ranks) { g_pool_share = ranks < 1 ? 1 : ranks; }

// parallel memcpy of `bytes` bytes
void host_copy(void* dst, const void* src, size_t bytes) {
  if (bytes < ((size_t)1 << 20)) {
    std::memcpy(dst, src, bytes);
    return;
  }
  std::function<void(int, int)> fn = [&](int part, int parts) {
    const size_t slice = ((bytes + parts - 1) / parts + 63) & ~(size_t)63;
    const size_t o = (size_t)part * slice;
    if (o < bytes) std::memcpy((char*)dst + o, (const char*)src + o, std::min(slice, bytes - o));
  };
  pool().run(fn);
}

// parallel narrowing copy of n Int64 values; returns the OR of the values
uint64_t host_pack(uint32_t* dst, const int64_t* src, size_t n) {
  if (n < ((size_t)1 << 16)) return pack_i64_u32(src, dst, n);
  std::atomic<uint64_t> acc{0};
  std::function<void(int, int)> fn = [&](int part, int parts) {
    const size_t slice = ((n + parts - 1) / parts + 15) & ~(size_t)15;
    const size_t o = (size_t)part * slice;
    if (o < n) acc.fetch_or(pack_i64_u32(src + o, dst + o, std::min(slice, n - o)), std::memory_order_relaxed);
  };
  pool().run(fn);
  return acc.load();
}

}  // namespace cpb
