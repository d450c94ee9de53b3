// Persistent host worker pool for the planner's data-parallel passes (enumeration, program
// construction, result assembly).  Spawning std::threads per pass cost ~0.4 ms per pass on 16
// cores -- as much as the pass itself at 256 taxa; the pool's workers sleep on a condition
// variable between passes.  run(n, fn) calls fn(i, n) for i in [0, n) and returns when all are done.
// A slice that throws (PLG_FAIL / PLG_CUDA / bad_alloc, on the caller or on a worker) does not tear the
// pool down: the first exception is kept, every slice still runs to its end, the pool is left idle,
// and the exception is rethrown on the caller -- so it reaches the C-ABI's status code, never
// std::terminate in a worker thread.
#pragma once
#include <condition_variable>
#include <exception>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

extern "C" const char *plg_last_error(void);

namespace plg {

void set_error(const char *fmt, ...);

class HostPool {
 public:
  static HostPool &get() {
    static HostPool *pool = new HostPool();  // never destroyed: workers must outlive static teardown
    return *pool;
  }
  int size() const { return (int)workers_.size() + 1; }

  template <typename F>
  void run(int n, F fn) {
    if (n <= 1) { fn(0, 1); return; }
    std::lock_guard<std::mutex> one_at_a_time(run_mu_);
    std::function<void(int, int)> f = [this, &fn](int i, int n) {
      try {
        fn(i, n);
      } catch (...) {
        std::lock_guard<std::mutex> lk(mu_);
        if (!err_) {
          err_ = std::current_exception();
          err_msg_ = plg_last_error();  // the message lives in the FAILING thread's storage: carry it over
        }
      }
    };
    {
      std::lock_guard<std::mutex> lk(mu_);
      err_ = nullptr;
      fn_ = &f;
      n_ = n;
      next_ = 1;      // slice 0 runs on the caller
      pending_ = n - 1;
      gen_++;
    }
    cv_.notify_all();
    f(0, n);
    for (;;) {  // the caller helps with slices nobody has taken yet
      int i;
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (next_ >= n_) break;
        i = next_++;
      }
      f(i, n);
      std::lock_guard<std::mutex> lk(mu_);
      pending_--;
    }
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
    if (err_) {
      std::exception_ptr e = err_;
      err_ = nullptr;
      set_error("%s", err_msg_.c_str());
      lk.unlock();
      std::rethrow_exception(e);
    }
  }

 private:
  HostPool() {
    const int n = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u)) - 1;
    for (int i = 0; i < n; i++) workers_.emplace_back([this] { loop(); });
    for (auto &w : workers_) w.detach();
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return gen_ != seen && fn_ && next_ < n_; });
      seen = gen_;
      while (fn_ && next_ < n_) {
        const int i = next_++, n = n_;
        auto *f = fn_;
        lk.unlock();
        (*f)(i, n);
        lk.lock();
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  std::mutex mu_, run_mu_;
  std::condition_variable cv_, done_;
  std::vector<std::thread> workers_;
  const std::function<void(int, int)> *fn_ = nullptr;
  std::exception_ptr err_;
  std::string err_msg_;
  int n_ = 0, next_ = 0, pending_ = 0;
  uint64_t gen_ = 0;
};

}  // namespace plg
