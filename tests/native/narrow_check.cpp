// CPU check of the ingest narrowing kernels (csrc/host/host_narrow.cpp: AVX2 + scalar): flags and values
// against the definition, on random inputs with negative, 32-bit-boundary, 64-bit and wide values.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <random>
namespace b200 { bool narrow_i128_to_i32(const int64_t*, int64_t, int32_t*); bool narrow_i128_to_i64(const int64_t*, int64_t, int64_t*); }
int main() {
  std::mt19937_64 rng(1);
  int fails = 0;
  for (int t = 0; t < 2000; t++) {
    int64_t n = 1 + rng() % 70;
    std::vector<int64_t> v(2 * n);
    bool fit32 = true, fit64 = true;
    for (int64_t i = 0; i < n; i++) {
      int64_t lo = (int64_t)(rng() % 4000000000ull) - 2000000000ll; int64_t hi = lo >> 63;
      int k = rng() % 200;
      if (k == 0) lo = (int64_t)rng();             // 64-bit range
      if (k == 1) { lo = (int64_t)rng(); hi = (int64_t)(rng() % 5) - 2; }  // wide
      if (k == 2) lo = 2147483648ll; if (k == 3) lo = -2147483649ll; if (k == 4) lo = 2147483647ll; if (k == 5) lo = -2147483648ll;
      if (k != 1) hi = lo >> 63;
      v[2*i] = lo; v[2*i+1] = hi;
      if (hi != (lo >> 63)) fit64 = fit32 = false;
      if (lo != (int64_t)(int32_t)lo) fit32 = false;
    }
    std::vector<int32_t> o32(n); std::vector<int64_t> o64(n);
    bool r32 = b200::narrow_i128_to_i32(v.data(), n, o32.data());
    bool r64 = b200::narrow_i128_to_i64(v.data(), n, o64.data());
    if (r32 != fit32 || r64 != fit64) { fails++; printf("flag mismatch n=%ld r32=%d fit32=%d r64=%d fit64=%d\n", (long)n, r32, fit32, r64, fit64); }
    if (r32) for (int64_t i = 0; i < n; i++) if (o32[i] != (int32_t)v[2*i]) { fails++; printf("v32 mismatch\n"); break; }
    if (r64) for (int64_t i = 0; i < n; i++) if (o64[i] != v[2*i]) { fails++; printf("v64 mismatch at %ld of %ld\n", (long)i, (long)n); break; }
  }
  printf("fails=%d\n", fails);
  return fails != 0;
}
