// CPU check of the ingest host pool (csrc/host/host_pool.hpp): every index of every parallel_for runs exactly
// once, across many back-to-back generations and pool sizes, including n smaller than the pool and n == 0.
#include <atomic>
#include <cstdio>
#include <vector>
#include "../../datafusion-ballista_b200/csrc/host/host_pool.hpp"

int main() {
  int fails = 0;
  for (int threads : {1, 2, 3, 8, 33}) {
    b200::HostPool pool(threads);
    if (pool.size() != threads) { fails++; std::printf("size mismatch %d vs %d\n", pool.size(), threads); }
    for (int gen = 0; gen < 300; gen++) {
      const int n = (gen * 7919) % 97;  // 0 .. 96, includes 0 and values below the pool size
      std::vector<std::atomic<int>> hits(n > 0 ? n : 1);
      for (auto& h : hits) h.store(0);
      std::atomic<long long> sum{0};
      pool.parallel_for(n, [&](int i) {
        hits[i].fetch_add(1);
        sum.fetch_add(i);
      });
      long long want = (long long)n * (n - 1) / 2;
      if (n > 0 && sum.load() != want) { fails++; std::printf("sum mismatch threads=%d n=%d\n", threads, n); }
      for (int i = 0; i < n; i++)
        if (hits[i].load() != 1) { fails++; std::printf("index %d ran %d times (threads=%d n=%d)\n", i, hits[i].load(), threads, n); break; }
    }
  }
  std::printf("fails=%d\n", fails);
  return fails != 0;
}
