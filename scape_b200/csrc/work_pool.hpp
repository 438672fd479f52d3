// Persistent host worker pool of the wave scheduler.
//
// The host side of a wave is three short parallel regions (RNG replay of the initial chains, one
// task per RNG stream).  Workers sleep on a condition variable between regions: with one process
// per GPU on a shared host (torchrun, 8 ranks) spinning runtimes oversubscribe the cores and the
// regions get 20x slower, and a sleeping pool costs ~10 us per region against 13 ms of kernels.
#pragma once
#include <sched.h>
#include <sys/resource.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace scape {

// CPUs this process may use, divided by the ranks that share the host (torchrun's LOCAL_WORLD_SIZE).
inline int default_host_threads() {
  int n = 0;
  cpu_set_t set;
  CPU_ZERO(&set);
  if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
  if (n <= 0) n = int(std::thread::hardware_concurrency());
  if (n <= 0) n = 1;
  if (const char* s = getenv("LOCAL_WORLD_SIZE")) {
    const int w = atoi(s);
    if (w > 1) n = std::max(1, n / w);
  }
  return n;
}

class WorkPool {
 public:
  // `background`: workers run at nice 10 (Linux niceness is per thread), so a pool that works ahead
  // (the RNG-free pre-pass) yields the cores to the pool on the critical path when both want them.
  explicit WorkPool(int threads, bool background = false) : n_workers_(std::max(0, threads - 1)) {
    for (int i = 0; i < n_workers_; i++)
      workers_.emplace_back([this, background]() {
        if (background) lower_priority();
        loop();
      });
  }
  static void lower_priority() { setpriority(PRIO_PROCESS, id_t(syscall(SYS_gettid)), 10); }
  ~WorkPool() {
    {
      std::lock_guard<std::mutex> g(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int threads() const { return n_workers_ + 1; }

  // fn(i) for i in [0, n), dynamic schedule; the caller works too and returns when all are done
  void run(int64_t n, const std::function<void(int64_t)>& fn) {
    if (n <= 0) return;
    if (n_workers_ == 0 || n == 1) {
      for (int64_t i = 0; i < n; i++) fn(i);
      return;
    }
    {
      std::lock_guard<std::mutex> g(m_);
      fn_ = &fn;
      n_ = n;
      next_.store(0, std::memory_order_relaxed);
      busy_ = n_workers_;
      gen_++;
    }
    cv_.notify_all();
    drain();
    std::unique_lock<std::mutex> g(m_);
    done_.wait(g, [this]() { return busy_ == 0; });
    fn_ = nullptr;
  }

 private:
  void drain() {
    for (;;) {
      const int64_t i = next_.fetch_add(1, std::memory_order_relaxed);
      if (i >= n_) break;
      (*fn_)(i);
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&]() { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
      }
      drain();
      {
        std::lock_guard<std::mutex> g(m_);
        if (--busy_ == 0) done_.notify_one();
      }
    }
  }
  int n_workers_;
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(int64_t)>* fn_ = nullptr;
  int64_t n_ = 0;
  std::atomic<int64_t> next_{0};
  int busy_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

}  // namespace scape
