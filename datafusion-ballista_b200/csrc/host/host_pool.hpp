// Small persistent fork-join pool for the host side of ingest (b200_engine_register_batch).
//
// The reference hands the engine Arrow batches in host memory (`ExecutionPlan::execute` streams of
// RecordBatch, ballista/core/src/execution_plans/shuffle_writer.rs:218); getting them into HBM is a
// PCIe copy, and for Decimal128 columns whose values fit 32 / 64 bits three quarters / half of those
// bytes are sign extension.  The pool lets idle host cores squeeze them out before the copy
// (engine.cpp: import_batch); the device widens them back, bit-exactly.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace b200 {

class HostPool {
 public:
  explicit HostPool(int n_threads) {
    if (n_threads < 1) n_threads = 1;
    for (int i = 0; i < n_threads - 1; i++) workers_.emplace_back([this] { worker(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return (int)workers_.size() + 1; }

  // run fn(i) for i in [0, n) on the pool (the caller takes part); returns when all are done
  void parallel_for(int n, const std::function<void(int)>& fn) {
    if (n <= 0) return;
    auto job = std::make_shared<Job>();
    job->fn = &fn;
    job->n = n;
    job->pending.store(n);
    {
      std::lock_guard<std::mutex> g(mu_);
      job_ = job;
      gen_++;
    }
    cv_.notify_all();
    drain(*job);
    std::unique_lock<std::mutex> l(mu_);
    done_cv_.wait(l, [&] { return job->pending.load() == 0; });
    job_.reset();
  }

 private:
  struct Job {  // one parallel_for; late workers only ever touch the job they woke up for
    const std::function<void(int)>* fn = nullptr;
    int n = 0;
    std::atomic<int> next{0}, pending{0};
  };
  void drain(Job& job) {
    for (;;) {
      const int i = job.next.fetch_add(1);
      if (i >= job.n) break;
      (*job.fn)(i);
      if (job.pending.fetch_sub(1) == 1) {
        std::lock_guard<std::mutex> g(mu_);
        done_cv_.notify_all();
      }
    }
  }
  void worker() {
    uint64_t seen = 0;
    for (;;) {
      std::shared_ptr<Job> job;
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        job = job_;
      }
      if (job) drain(*job);
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  std::shared_ptr<Job> job_;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

// Decimal128 (16-byte little-endian two's complement) -> int32 / int64 when every value fits
// (host_narrow.cpp: AVX2 with a scalar fallback).  Return true on success; on failure `out` holds
// garbage and the caller falls back to the next width.
bool narrow_i128_to_i32(const int64_t* p, int64_t n, int32_t* out);
bool narrow_i128_to_i64(const int64_t* p, int64_t n, int64_t* out);

}  // namespace b200
