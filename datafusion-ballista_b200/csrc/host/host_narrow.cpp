// Host-side narrowing of Decimal128 values for ingest (see host_pool.hpp): plain C++ translation unit,
// compiled by g++ (not nvcc) so that the AVX2 path can use intrinsics with a function-level target.
#include <cstdint>
#include <cstdlib>
#include <immintrin.h>

namespace b200 {

static bool narrow32_scalar(const int64_t* p, int64_t n, int32_t* out) {
  uint64_t bad = 0;
  for (int64_t i = 0; i < n; i++) {
    const int64_t lo = p[2 * i], hi = p[2 * i + 1];
    bad |= (uint64_t)(hi ^ (lo >> 63)) | (uint64_t)(lo ^ (int64_t)(int32_t)lo);
    out[i] = (int32_t)lo;
  }
  return bad == 0;
}
static bool narrow64_scalar(const int64_t* p, int64_t n, int64_t* out) {
  uint64_t bad = 0;
  for (int64_t i = 0; i < n; i++) {
    const int64_t lo = p[2 * i], hi = p[2 * i + 1];
    bad |= (uint64_t)(hi ^ (lo >> 63));
    out[i] = lo;
  }
  return bad == 0;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) static bool narrow32_avx2(const int64_t* p, int64_t n, int32_t* out) {
  __m256i bad = _mm256_setzero_si256();
  const __m256i zero = _mm256_setzero_si256();
  const __m256i hi_mask = _mm256_set1_epi64x((long long)0xFFFFFFFF00000000ull);
  const __m256i pick = _mm256_setr_epi32(0, 4, 2, 6, 0, 4, 2, 6);
  int64_t i = 0;
  for (; i + 4 <= n; i += 4) {
    const __m256i v0 = _mm256_loadu_si256((const __m256i*)(p + 2 * i));      // lo0 hi0 lo1 hi1
    const __m256i v1 = _mm256_loadu_si256((const __m256i*)(p + 2 * i + 4));  // lo2 hi2 lo3 hi3
    const __m256i lo = _mm256_unpacklo_epi64(v0, v1);                        // lo0 lo2 lo1 lo3
    const __m256i hi = _mm256_unpackhi_epi64(v0, v1);                        // hi0 hi2 hi1 hi3
    const __m256i sgn = _mm256_cmpgt_epi64(zero, lo);                        // all ones where lo < 0
    bad = _mm256_or_si256(bad, _mm256_xor_si256(hi, sgn));                   // hi must be the sign extension
    const __m256i s32 = _mm256_shuffle_epi32(_mm256_srai_epi32(lo, 31), 0xA0);  // sign of the low dword, in both dwords
    bad = _mm256_or_si256(bad, _mm256_and_si256(_mm256_xor_si256(lo, s32), hi_mask));  // upper dword == that sign
    const __m256i packed = _mm256_permutevar8x32_epi32(lo, pick);            // lo0 lo1 lo2 lo3 (low dwords)
    _mm_storeu_si128((__m128i*)(out + i), _mm256_castsi256_si128(packed));
  }
  bool ok = _mm256_testz_si256(bad, bad) != 0;
  if (i < n) ok &= narrow32_scalar(p + 2 * i, n - i, out + i);
  return ok;
}
__attribute__((target("avx2"))) static bool narrow64_avx2(const int64_t* p, int64_t n, int64_t* out) {
  __m256i bad = _mm256_setzero_si256();
  const __m256i zero = _mm256_setzero_si256();
  int64_t i = 0;
  for (; i + 4 <= n; i += 4) {
    const __m256i v0 = _mm256_loadu_si256((const __m256i*)(p + 2 * i));
    const __m256i v1 = _mm256_loadu_si256((const __m256i*)(p + 2 * i + 4));
    const __m256i lo = _mm256_unpacklo_epi64(v0, v1);  // lo0 lo2 lo1 lo3
    const __m256i hi = _mm256_unpackhi_epi64(v0, v1);
    bad = _mm256_or_si256(bad, _mm256_xor_si256(hi, _mm256_cmpgt_epi64(zero, lo)));
    _mm256_storeu_si256((__m256i*)(out + i), _mm256_permute4x64_epi64(lo, 0xD8));  // -> lo0 lo1 lo2 lo3
  }
  bool ok = _mm256_testz_si256(bad, bad) != 0;
  if (i < n) ok &= narrow64_scalar(p + 2 * i, n - i, out + i);
  return ok;
}
#endif

// AVX-512: 8 values (128 source bytes) per iteration -- two full-width loads, one two-source permute per word half, one
// truncating down-convert; the checks fold into mask registers
__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq"))) static bool narrow32_avx512(const int64_t* p, int64_t n, int32_t* out) {
  const __m512i even = _mm512_setr_epi64(0, 2, 4, 6, 8, 10, 12, 14), odd = _mm512_setr_epi64(1, 3, 5, 7, 9, 11, 13, 15);
  __mmask8 bad = 0;
  int64_t i = 0;
  for (; i + 8 <= n; i += 8) {
    const __m512i v0 = _mm512_loadu_si512((const void*)(p + 2 * i));
    const __m512i v1 = _mm512_loadu_si512((const void*)(p + 2 * i + 8));
    const __m512i lo = _mm512_permutex2var_epi64(v0, even, v1);
    const __m512i hi = _mm512_permutex2var_epi64(v0, odd, v1);
    const __m256i t = _mm512_cvtepi64_epi32(lo);                                   // low dwords
    bad |= _mm512_cmpneq_epi64_mask(hi, _mm512_srai_epi64(lo, 63));                  // high word == sign extension
    bad |= _mm512_cmpneq_epi64_mask(lo, _mm512_cvtepi32_epi64(t));                   // value == its own 32-bit image
    _mm256_storeu_si256((__m256i*)(out + i), t);
  }
  bool ok = bad == 0;
  if (i < n) ok &= narrow32_scalar(p + 2 * i, n - i, out + i);
  return ok;
}
__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq"))) static bool narrow64_avx512(const int64_t* p, int64_t n, int64_t* out) {
  const __m512i even = _mm512_setr_epi64(0, 2, 4, 6, 8, 10, 12, 14), odd = _mm512_setr_epi64(1, 3, 5, 7, 9, 11, 13, 15);
  __mmask8 bad = 0;
  int64_t i = 0;
  for (; i + 8 <= n; i += 8) {
    const __m512i v0 = _mm512_loadu_si512((const void*)(p + 2 * i));
    const __m512i v1 = _mm512_loadu_si512((const void*)(p + 2 * i + 8));
    const __m512i lo = _mm512_permutex2var_epi64(v0, even, v1);
    const __m512i hi = _mm512_permutex2var_epi64(v0, odd, v1);
    bad |= _mm512_cmpneq_epi64_mask(hi, _mm512_srai_epi64(lo, 63));
    _mm512_storeu_si512((void*)(out + i), lo);
  }
  bool ok = bad == 0;
  if (i < n) ok &= narrow64_scalar(p + 2 * i, n - i, out + i);
  return ok;
}
static bool have_avx512() {
  static const bool v = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
                        __builtin_cpu_supports("avx512dq") && !getenv("B200_NO_AVX512");
  return v;
}

static bool have_avx2() {
#if defined(__x86_64__)
  static const bool v = __builtin_cpu_supports("avx2");
  return v;
#else
  return false;
#endif
}

// Decimal128 (16-byte little-endian two's complement) -> int32 / int64 when every value fits.
// Return true on success; on failure `out` holds garbage and the caller falls back to the next width.
bool narrow_i128_to_i32(const int64_t* p, int64_t n, int32_t* out) {
#if defined(__x86_64__)
  if (have_avx512()) return narrow32_avx512(p, n, out);
  if (have_avx2()) return narrow32_avx2(p, n, out);
#endif
  return narrow32_scalar(p, n, out);
}
bool narrow_i128_to_i64(const int64_t* p, int64_t n, int64_t* out) {
#if defined(__x86_64__)
  if (have_avx512()) return narrow64_avx512(p, n, out);
  if (have_avx2()) return narrow64_avx2(p, n, out);
#endif
  return narrow64_scalar(p, n, out);
}

}  // namespace b200
