// Fused pipeline kernel for sm_100a:  scan -> [FilterExec | ProjectionExec]* -> sink
// (sink = materialise/compact | partial-or-final hash aggregate).
//
// Reference operators replaced: DataFusion FilterExec, ProjectionExec and
// AggregateExec pulled by the shuffle writers at ballista/core/src/execution_plans/
// shuffle_writer.rs:218 and sort_shuffle/writer.rs:214.  The CPU path evaluates each PhysicalExpr
// column-at-a-time over 8192-row batches and materialises every intermediate array; here one
// persistent CTA per SM streams column tiles HBM -> shared memory with 1-D TMA bulk copies
// (cp.async.bulk + mbarrier, multi-stage ring), evaluates the whole expression program on the
// resident tile (values never leave the SM) and folds rows straight into the sink.
//
// HBM traffic per row == the Arrow bytes of the referenced columns, once (SURVEY.md 8(d)
// "fused scan->filter->project->partial-agg": N*w_referenced + G*(w_keys+w_state)).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../common/hash.hpp"
#include "kernels.h"
#include "program.h"

namespace b200 {

typedef __int128 i128;
typedef unsigned __int128 u128;

// ------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + 1-D bulk async copy (TMA without a tensor map; SASS: UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// 128-bit helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ i128 make_i128(uint64_t lo, uint64_t hi) { return (i128)(((u128)hi << 64) | (u128)lo); }
__device__ __forceinline__ uint64_t lo64(i128 v) { return (uint64_t)v; }
__device__ __forceinline__ uint64_t hi64(i128 v) { return (uint64_t)((u128)v >> 64); }
__device__ __forceinline__ bool fits_i64(i128 v) { return (i128)(int64_t)v == v; }

// returns true on overflow
__device__ __forceinline__ bool mul_i128_checked(i128 a, i128 b, i128* out) {
  if (fits_i64(a) && fits_i64(b)) {  // 64x64 -> 128 never overflows
    int64_t x = (int64_t)a, y = (int64_t)b;
    uint64_t lo = (uint64_t)x * (uint64_t)y;
    int64_t hi = __mul64hi(x, y);
    *out = make_i128(lo, (uint64_t)hi);
    return false;
  }
  bool neg = (a < 0) != (b < 0);
  u128 ua = a < 0 ? (u128)0 - (u128)a : (u128)a;
  u128 ub = b < 0 ? (u128)0 - (u128)b : (u128)b;
  uint64_t al = (uint64_t)ua, ah = (uint64_t)(ua >> 64), bl = (uint64_t)ub, bh = (uint64_t)(ub >> 64);
  if (ah && bh) return true;
  uint64_t ch = ah ? ah : bh, cl = ah ? bl : al;  // cross term (at most one is non-zero)
  uint64_t cross_lo = ch * cl, cross_hi = __umul64hi(ch, cl);
  if (cross_hi) return true;
  uint64_t lo = al * bl, hi = __umul64hi(al, bl);
  uint64_t hi2 = hi + cross_lo;
  if (hi2 < hi) return true;
  u128 r = ((u128)hi2 << 64) | lo;
  if (neg) {
    if (r > ((u128)1 << 127)) return true;
    *out = (i128)((u128)0 - r);
  } else {
    if (r >> 127) return true;
    *out = (i128)r;
  }
  return false;
}
__device__ __forceinline__ bool add_i128_checked(i128 a, i128 b, i128* out) {
  i128 r = (i128)((u128)a + (u128)b);
  *out = r;
  return ((a < 0) == (b < 0)) && ((r < 0) != (a < 0));
}
__device__ __forceinline__ bool sub_i128_checked(i128 a, i128 b, i128* out) {
  i128 r = (i128)((u128)a - (u128)b);
  *out = r;
  return ((a < 0) != (b < 0)) && ((r < 0) != (a < 0));
}
__device__ i128 pow10_dev(int n) {
  i128 r = 1;
  for (int i = 0; i < n; i++) r *= 10;
  return r;
}
__device__ __forceinline__ int total_cmp_f64(double a, double b) {
  long long x = __double_as_longlong(a), y = __double_as_longlong(b);
  x ^= (long long)((unsigned long long)(x >> 63) >> 1);
  y ^= (long long)((unsigned long long)(y >> 63) >> 1);
  return x < y ? -1 : (x > y ? 1 : 0);
}
__device__ __forceinline__ long long f64_order_key(double a) {
  long long x = __double_as_longlong(a);
  return x ^ (long long)((unsigned long long)(x >> 63) >> 1);
}
__device__ __forceinline__ double f64_from_order_key(long long k) {
  long long x = k ^ (long long)((unsigned long long)(k >> 63) >> 1);
  return __longlong_as_double(x);
}

struct StrRef {
  const uint8_t* p;
  uint32_t len;
};

__device__ __forceinline__ int str_cmp(StrRef a, StrRef b) {
  uint32_t n = a.len < b.len ? a.len : b.len;
  for (uint32_t i = 0; i < n; i++) {
    uint8_t x = a.p[i], y = b.p[i];
    if (x != y) return x < y ? -1 : 1;
  }
  return a.len < b.len ? -1 : (a.len > b.len ? 1 : 0);
}
__device__ __forceinline__ bool str_eq(StrRef a, StrRef b) {
  if (a.len != b.len) return false;
  for (uint32_t i = 0; i < a.len; i++)
    if (a.p[i] != b.p[i]) return false;
  return true;
}
__device__ bool like_match_dev(const uint8_t* s, uint32_t sn, const uint8_t* p, uint32_t pn) {
  uint32_t si = 0, pi = 0, star_p = 0xFFFFFFFFu, star_s = 0;
  while (si < sn) {
    if (pi < pn && p[pi] != '%' && (p[pi] == '_' || p[pi] == s[si])) {
      si++;
      pi++;
      continue;
    }
    if (pi < pn && p[pi] == '%') {
      star_p = pi++;
      star_s = si;
      continue;
    }
    if (star_p != 0xFFFFFFFFu) {
      pi = star_p + 1;
      si = ++star_s;
      continue;
    }
    return false;
  }
  while (pi < pn && p[pi] == '%') pi++;
  return pi == pn;
}
__device__ __forceinline__ int64_t year_of_days_dev(int64_t z) {
  z += 719468;
  int64_t era = (z >= 0 ? z : z - 146096) / 146097;
  int64_t doe = z - era * 146097;
  int64_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  int64_t y = yoe + era * 400;
  int64_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  int64_t mp = (5 * doy + 2) / 153;
  int64_t m = mp < 10 ? mp + 3 : mp - 9;
  return y + (m <= 2);
}

// ------------------------------------------------------------------------------------------------
// Per-thread execution context
// ------------------------------------------------------------------------------------------------
struct Ctx {
  const Program* P;
  const uint8_t* stage;  // current stage buffer (source tile)
  uint8_t* regs;         // VM register file
  int tid;
  int B;                 // blockDim.x
  int64_t tile_base;     // first row of the tile
  uint32_t active;       // bit r: row (r*B + tid) of the tile is live
};

#define FOR_R for (int r = 0; r < VM_R; r++)

__device__ __forceinline__ uint32_t fetch_valid(const Ctx& c, Operand o) {
  const Program& P = *c.P;
  if (o.kind == OPD_COL) {
    const ColDesc& cd = P.cols[o.idx];
    if (!cd.valid) return 0xFFFFFFFFu;
    const uint8_t* v = c.stage + cd.valid_smem_off;
    uint32_t m = 0;
#pragma unroll
    FOR_R m |= (v[r * c.B + c.tid] ? 1u : 0u) << r;
    return m;
  }
  if (o.kind == OPD_REG) {
    const RegDesc& rd = P.regs[o.idx];
    if (rd.valid_off == 0xFFFFFFFFu) return 0xFFFFFFFFu;
    return ((const uint32_t*)(c.regs + rd.valid_off))[c.tid];
  }
  if (o.kind == OPD_IMM) return P.imms[o.idx].is_null ? 0u : 0xFFFFFFFFu;
  return 0xFFFFFFFFu;
}

__device__ __forceinline__ void store_valid(const Ctx& c, Operand dst, uint32_t m) {
  const RegDesc& rd = c.P->regs[dst.idx];
  if (rd.valid_off != 0xFFFFFFFFu) ((uint32_t*)(c.regs + rd.valid_off))[c.tid] = m;
}

__device__ __forceinline__ void fetch_i64(const Ctx& c, Operand o, int64_t v[VM_R]) {
  const Program& P = *c.P;
  if (o.kind == OPD_COL) {
    const ColDesc& cd = P.cols[o.idx];
    const uint8_t* base = c.stage + cd.smem_off;
    switch (cd.phys) {
      case PH_I32:
#pragma unroll
        FOR_R v[r] = ((const int32_t*)base)[r * c.B + c.tid];
        break;
      case PH_I64:
      case PH_U64:
#pragma unroll
        FOR_R v[r] = ((const int64_t*)base)[r * c.B + c.tid];
        break;
      case PH_U32:
#pragma unroll
        FOR_R v[r] = ((const uint32_t*)base)[r * c.B + c.tid];
        break;
      case PH_I16:
#pragma unroll
        FOR_R v[r] = ((const int16_t*)base)[r * c.B + c.tid];
        break;
      case PH_U16:
#pragma unroll
        FOR_R v[r] = ((const uint16_t*)base)[r * c.B + c.tid];
        break;
      case PH_I8:
#pragma unroll
        FOR_R v[r] = ((const int8_t*)base)[r * c.B + c.tid];
        break;
      case PH_DEC128:  // low word of a decimal known to fit 64 bits
#pragma unroll
        FOR_R v[r] = (int64_t)((const ulonglong2*)base)[r * c.B + c.tid].x;
        break;
      default:  // PH_U8 / PH_BOOL8
#pragma unroll
        FOR_R v[r] = ((const uint8_t*)base)[r * c.B + c.tid];
        break;
    }
  } else if (o.kind == OPD_REG) {
    const RegDesc& rd = P.regs[o.idx];
    if (rd.vk == VK_BOOL) {
      uint32_t m = ((const uint32_t*)(c.regs + rd.smem_off))[c.tid];
#pragma unroll
      FOR_R v[r] = (m >> r) & 1;
    } else if (rd.vk == VK_I128) {
      const ulonglong2* p = (const ulonglong2*)(c.regs + rd.smem_off);
#pragma unroll
      FOR_R v[r] = (int64_t)p[r * c.B + c.tid].x;
    } else {
      const int64_t* p = (const int64_t*)(c.regs + rd.smem_off);
#pragma unroll
      FOR_R v[r] = p[r * c.B + c.tid];
    }
  } else {
    int64_t x = (int64_t)P.imms[o.idx].lo;
#pragma unroll
    FOR_R v[r] = x;
  }
}

__device__ __forceinline__ void fetch_f64(const Ctx& c, Operand o, double v[VM_R]) {
  const Program& P = *c.P;
  if (o.kind == OPD_COL) {
    const ColDesc& cd = P.cols[o.idx];
    const uint8_t* base = c.stage + cd.smem_off;
    if (cd.phys == PH_F32) {
#pragma unroll
      FOR_R v[r] = (double)((const float*)base)[r * c.B + c.tid];
    } else {
#pragma unroll
      FOR_R v[r] = ((const double*)base)[r * c.B + c.tid];
    }
  } else if (o.kind == OPD_REG) {
    const double* p = (const double*)(c.regs + P.regs[o.idx].smem_off);
#pragma unroll
    FOR_R v[r] = p[r * c.B + c.tid];
  } else {
    double x = __longlong_as_double((long long)P.imms[o.idx].lo);
#pragma unroll
    FOR_R v[r] = x;
  }
}

__device__ __forceinline__ void fetch_i128(const Ctx& c, Operand o, i128 v[VM_R]) {
  const Program& P = *c.P;
  if (o.vk != VK_I128) {  // integer operand used in a decimal context: sign-extend
    int64_t t[VM_R];
    fetch_i64(c, o, t);
#pragma unroll
    FOR_R v[r] = (i128)t[r];
    return;
  }
  if (o.kind == OPD_COL) {
    const ulonglong2* p = (const ulonglong2*)(c.stage + P.cols[o.idx].smem_off);
#pragma unroll
    FOR_R {
      ulonglong2 x = p[r * c.B + c.tid];
      v[r] = make_i128(x.x, x.y);
    }
  } else if (o.kind == OPD_REG) {
    const ulonglong2* p = (const ulonglong2*)(c.regs + P.regs[o.idx].smem_off);
#pragma unroll
    FOR_R {
      ulonglong2 x = p[r * c.B + c.tid];
      v[r] = make_i128(x.x, x.y);
    }
  } else {
    i128 x = make_i128(P.imms[o.idx].lo, P.imms[o.idx].hi);
#pragma unroll
    FOR_R v[r] = x;
  }
}

__device__ __forceinline__ void fetch_str(const Ctx& c, Operand o, StrRef v[VM_R]) {
  const Program& P = *c.P;
  if (o.kind == OPD_COL) {
    const ColDesc& cd = P.cols[o.idx];
    if (cd.phys == PH_UTF8) {
      const int32_t* off = (const int32_t*)(c.stage + cd.smem_off);
#pragma unroll
      FOR_R {
        int32_t o0 = off[r * c.B + c.tid], o1 = off[r * c.B + c.tid + 1];
        v[r].p = cd.chars + o0;
        v[r].len = (uint32_t)(o1 - o0);
      }
    } else {
      const ulonglong2* p = (const ulonglong2*)(c.stage + cd.smem_off);
#pragma unroll
      FOR_R {
        ulonglong2 x = p[r * c.B + c.tid];
        v[r].p = (const uint8_t*)x.x;
        v[r].len = (uint32_t)x.y;
      }
    }
  } else if (o.kind == OPD_REG) {
    const ulonglong2* p = (const ulonglong2*)(c.regs + P.regs[o.idx].smem_off);
#pragma unroll
    FOR_R {
      ulonglong2 x = p[r * c.B + c.tid];
      v[r].p = (const uint8_t*)x.x;
      v[r].len = (uint32_t)x.y;
    }
  } else {
    StrRef s;
    s.p = (const uint8_t*)P.imms[o.idx].lo;
    s.len = (uint32_t)P.imms[o.idx].hi;
#pragma unroll
    FOR_R v[r] = s;
  }
}

__device__ __forceinline__ uint32_t fetch_bool(const Ctx& c, Operand o) {
  const Program& P = *c.P;
  if (o.kind == OPD_REG && P.regs[o.idx].vk == VK_BOOL) return ((const uint32_t*)(c.regs + P.regs[o.idx].smem_off))[c.tid];
  int64_t t[VM_R];
  fetch_i64(c, o, t);
  uint32_t m = 0;
#pragma unroll
  FOR_R m |= (t[r] != 0 ? 1u : 0u) << r;
  return m;
}

__device__ __forceinline__ void store_i64(const Ctx& c, Operand dst, const int64_t v[VM_R]) {
  int64_t* p = (int64_t*)(c.regs + c.P->regs[dst.idx].smem_off);
#pragma unroll
  FOR_R p[r * c.B + c.tid] = v[r];
}
__device__ __forceinline__ void store_f64(const Ctx& c, Operand dst, const double v[VM_R]) {
  double* p = (double*)(c.regs + c.P->regs[dst.idx].smem_off);
#pragma unroll
  FOR_R p[r * c.B + c.tid] = v[r];
}
__device__ __forceinline__ void store_i128(const Ctx& c, Operand dst, const i128 v[VM_R]) {
  ulonglong2* p = (ulonglong2*)(c.regs + c.P->regs[dst.idx].smem_off);
#pragma unroll
  FOR_R p[r * c.B + c.tid] = make_ulonglong2(lo64(v[r]), hi64(v[r]));
}
__device__ __forceinline__ void store_str(const Ctx& c, Operand dst, const StrRef v[VM_R]) {
  ulonglong2* p = (ulonglong2*)(c.regs + c.P->regs[dst.idx].smem_off);
#pragma unroll
  FOR_R p[r * c.B + c.tid] = make_ulonglong2((unsigned long long)v[r].p, (unsigned long long)v[r].len);
}
__device__ __forceinline__ void store_bool(const Ctx& c, Operand dst, uint32_t m) {
  ((uint32_t*)(c.regs + c.P->regs[dst.idx].smem_off))[c.tid] = m;
}

__device__ __forceinline__ void raise(const Program& P, unsigned int code) { atomicMax(&P.status->error, code); }

__device__ __forceinline__ bool cmp_result(int op, int c) {
  switch (op) {
    case OP_CMP_EQ: return c == 0;
    case OP_CMP_NE: return c != 0;
    case OP_CMP_LT: return c < 0;
    case OP_CMP_LE: return c <= 0;
    case OP_CMP_GT: return c > 0;
    default: return c >= 0;
  }
}

// ------------------------------------------------------------------------------------------------
// The interpreter: one pass over the expression program for the R rows this thread owns.
// All branches are warp-uniform (driven by the program, not by data).
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ void run_program(Ctx& c) {
  const Program& P = *c.P;
  for (int pc = 0; pc < P.n_instr; pc++) {
    const VInstr ins = P.code[pc];
    uint32_t va = 0xFFFFFFFFu, vb = 0xFFFFFFFFu;
    if (ins.flags & IF_NULLCHK) {
      va = fetch_valid(c, ins.a);
      if (ins.b.kind != OPD_NONE) vb = fetch_valid(c, ins.b);
    }
    const uint32_t live = c.active & va & vb;  // rows whose errors count
    switch (ins.op) {
      case OP_ADD:
      case OP_SUB:
      case OP_MUL:
      case OP_DIV:
      case OP_MOD: {
        uint32_t vout = va & vb;
        if (ins.t == VK_I64) {
          int64_t a[VM_R], b[VM_R], o[VM_R];
          fetch_i64(c, ins.a, a);
          fetch_i64(c, ins.b, b);
#pragma unroll
          FOR_R {
            uint64_t x = (uint64_t)a[r], y = (uint64_t)b[r];
            switch (ins.op) {
              case OP_ADD: o[r] = (int64_t)(x + y); break;
              case OP_SUB: o[r] = (int64_t)(x - y); break;
              case OP_MUL: o[r] = (int64_t)(x * y); break;
              default: {
                if (b[r] == 0) {
                  if ((live >> r) & 1) raise(P, 2);
                  o[r] = 0;
                } else if (a[r] == INT64_MIN && b[r] == -1) {
                  if (ins.op == OP_DIV && ((live >> r) & 1)) raise(P, 1);
                  o[r] = 0;
                } else {
                  o[r] = ins.op == OP_DIV ? a[r] / b[r] : a[r] % b[r];
                }
              }
            }
          }
          store_i64(c, ins.dst, o);
        } else if (ins.t == VK_F64) {
          double a[VM_R], b[VM_R], o[VM_R];
          fetch_f64(c, ins.a, a);
          fetch_f64(c, ins.b, b);
#pragma unroll
          FOR_R {
            switch (ins.op) {
              case OP_ADD: o[r] = a[r] + b[r]; break;
              case OP_SUB: o[r] = a[r] - b[r]; break;
              case OP_MUL: o[r] = a[r] * b[r]; break;
              case OP_DIV: o[r] = a[r] / b[r]; break;
              default: o[r] = fmod(a[r], b[r]);
            }
            if (ins.aux == PH_F32) o[r] = (double)(float)o[r];
          }
          store_f64(c, ins.dst, o);
        } else {
          i128 a[VM_R], b[VM_R], o[VM_R];
          fetch_i128(c, ins.a, a);
          fetch_i128(c, ins.b, b);
#pragma unroll
          FOR_R {
            bool ovf = false;
            switch (ins.op) {
              case OP_ADD: ovf = add_i128_checked(a[r], b[r], &o[r]); break;
              case OP_SUB: ovf = sub_i128_checked(a[r], b[r], &o[r]); break;
              case OP_MUL: ovf = mul_i128_checked(a[r], b[r], &o[r]); break;
              default: {
                // DIV: a * 10^imm / b (truncating); MOD: a % b (operands pre-scaled by the lowering)
                if (b[r] == 0) {
                  if ((live >> r) & 1) raise(P, 2);
                  o[r] = 0;
                } else if (ins.op == OP_DIV) {
                  i128 num;
                  ovf = mul_i128_checked(a[r], pow10_dev(ins.imm), &num);
                  o[r] = ovf ? 0 : num / b[r];
                } else {
                  o[r] = a[r] % b[r];
                }
              }
            }
            if (ovf && ((live >> r) & 1)) raise(P, 1);
          }
          store_i128(c, ins.dst, o);
        }
        store_valid(c, ins.dst, vout);
        break;
      }
      case OP_DEC_MUL_LIT_MINUS:
      case OP_DEC_MUL_LIT_PLUS: {
        // dst = a * (imm +/- b): the TPC-H revenue shape l_extendedprice * (1 - l_discount)
        i128 a[VM_R], b[VM_R], o[VM_R];
        fetch_i128(c, ins.a, a);
        fetch_i128(c, ins.b, b);
        i128 lit = make_i128(P.imms[ins.imm].lo, P.imms[ins.imm].hi);
#pragma unroll
        FOR_R {
          i128 t;
          bool ovf = ins.op == OP_DEC_MUL_LIT_MINUS ? sub_i128_checked(lit, b[r], &t) : add_i128_checked(lit, b[r], &t);
          ovf |= mul_i128_checked(a[r], t, &o[r]);
          if (ovf && ((live >> r) & 1)) raise(P, 1);
        }
        store_i128(c, ins.dst, o);
        store_valid(c, ins.dst, va & vb);
        break;
      }
      case OP_NEG: {
        if (ins.t == VK_F64) {
          double a[VM_R];
          fetch_f64(c, ins.a, a);
#pragma unroll
          FOR_R a[r] = -a[r];
          store_f64(c, ins.dst, a);
        } else if (ins.t == VK_I128) {
          i128 a[VM_R];
          fetch_i128(c, ins.a, a);
#pragma unroll
          FOR_R a[r] = (i128)((u128)0 - (u128)a[r]);
          store_i128(c, ins.dst, a);
        } else {
          int64_t a[VM_R];
          fetch_i64(c, ins.a, a);
#pragma unroll
          FOR_R a[r] = (int64_t)(0 - (uint64_t)a[r]);
          store_i64(c, ins.dst, a);
        }
        store_valid(c, ins.dst, va);
        break;
      }
      case OP_CMP_EQ:
      case OP_CMP_NE:
      case OP_CMP_LT:
      case OP_CMP_LE:
      case OP_CMP_GT:
      case OP_CMP_GE: {
        uint32_t m = 0;
        if (ins.t == VK_I64 || ins.t == VK_BOOL) {
          int64_t a[VM_R], b[VM_R];
          fetch_i64(c, ins.a, a);
          fetch_i64(c, ins.b, b);
          if (ins.aux == PH_U64) {
#pragma unroll
            FOR_R m |= (cmp_result(ins.op, (uint64_t)a[r] < (uint64_t)b[r] ? -1 : ((uint64_t)a[r] > (uint64_t)b[r] ? 1 : 0)) ? 1u : 0u) << r;
          } else {
#pragma unroll
            FOR_R m |= (cmp_result(ins.op, a[r] < b[r] ? -1 : (a[r] > b[r] ? 1 : 0)) ? 1u : 0u) << r;
          }
        } else if (ins.t == VK_F64) {
          double a[VM_R], b[VM_R];
          fetch_f64(c, ins.a, a);
          fetch_f64(c, ins.b, b);
#pragma unroll
          FOR_R m |= (cmp_result(ins.op, total_cmp_f64(a[r], b[r])) ? 1u : 0u) << r;
        } else if (ins.t == VK_I128) {
          i128 a[VM_R], b[VM_R];
          fetch_i128(c, ins.a, a);
          fetch_i128(c, ins.b, b);
#pragma unroll
          FOR_R m |= (cmp_result(ins.op, a[r] < b[r] ? -1 : (a[r] > b[r] ? 1 : 0)) ? 1u : 0u) << r;
        } else {
          StrRef a[VM_R], b[VM_R];
          fetch_str(c, ins.a, a);
          fetch_str(c, ins.b, b);
#pragma unroll
          FOR_R {
            bool ok = ((c.active & va & vb) >> r) & 1;  // never chase pointers of dead rows
            int cm = ok ? ((ins.op == OP_CMP_EQ || ins.op == OP_CMP_NE) ? (str_eq(a[r], b[r]) ? 0 : 1) : str_cmp(a[r], b[r])) : 0;
            m |= (cmp_result(ins.op, cm) ? 1u : 0u) << r;
          }
        }
        store_bool(c, ins.dst, m);
        store_valid(c, ins.dst, va & vb);
        break;
      }
      case OP_AND:
      case OP_OR: {  // Kleene logic
        uint32_t a = fetch_bool(c, ins.a), b = fetch_bool(c, ins.b);
        uint32_t ta = a & va, tb = b & vb;    // definitely true
        uint32_t fa = ~a & va, fb = ~b & vb;  // definitely false
        uint32_t val, vld;
        if (ins.op == OP_AND) {
          val = ta & tb;
          vld = (va & vb) | fa | fb;
        } else {
          val = ta | tb;
          vld = (va & vb) | ta | tb;
        }
        store_bool(c, ins.dst, val);
        store_valid(c, ins.dst, vld);
        break;
      }
      case OP_NOT: {
        store_bool(c, ins.dst, ~fetch_bool(c, ins.a));
        store_valid(c, ins.dst, va);
        break;
      }
      case OP_IS_NULL:
      case OP_IS_NOT_NULL: {
        uint32_t v = fetch_valid(c, ins.a);
        store_bool(c, ins.dst, ins.op == OP_IS_NULL ? ~v : v);
        store_valid(c, ins.dst, 0xFFFFFFFFu);
        break;
      }
      case OP_CAST_I64_F64: {
        int64_t a[VM_R];
        double o[VM_R];
        fetch_i64(c, ins.a, a);
#pragma unroll
        FOR_R {
          o[r] = ins.aux == PH_U64 ? (double)(uint64_t)a[r] : (double)a[r];
          if (ins.imm == 1) o[r] = (double)(float)o[r];
        }
        store_f64(c, ins.dst, o);
        store_valid(c, ins.dst, va);
        break;
      }
      case OP_CAST_I64_I128:
      case OP_CAST_I128_I128_UP: {
        i128 a[VM_R];
        fetch_i128(c, ins.a, a);
        i128 mul = pow10_dev(ins.imm);
        uint32_t vout = va;
#pragma unroll
        FOR_R {
          i128 o;
          bool ovf = mul_i128_checked(a[r], mul, &o);
          a[r] = o;
          if (ovf && ((live >> r) & 1)) raise(P, 1);
        }
        store_i128(c, ins.dst, a);
        store_valid(c, ins.dst, vout);
        break;
      }
      case OP_CAST_I128_I128_DOWN: {
        i128 a[VM_R];
        fetch_i128(c, ins.a, a);
        i128 div = pow10_dev(ins.imm), half = div / 2;
#pragma unroll
        FOR_R {
          i128 q = a[r] / div, rem = a[r] % div;
          if (a[r] >= 0 && rem >= half) q += 1;
          else if (a[r] < 0 && rem <= -half) q -= 1;
          a[r] = q;
        }
        store_i128(c, ins.dst, a);
        store_valid(c, ins.dst, va);
        break;
      }
      case OP_CHECK_PRECISION: {  // |v| >= 10^aux -> NULL
        i128 a[VM_R];
        fetch_i128(c, ins.a, a);
        i128 lim = pow10_dev(ins.aux);
        uint32_t vout = va;
#pragma unroll
        FOR_R if (a[r] >= lim || a[r] <= -lim) {
          if (ins.flags & IF_CHECKED) {
            if ((live >> r) & 1) raise(P, 1);
          } else {
            vout &= ~(1u << r);
          }
        }
        store_i128(c, ins.dst, a);
        store_valid(c, ins.dst, vout);
        break;
      }
      case OP_CAST_I128_F64: {
        i128 a[VM_R];
        double o[VM_R];
        fetch_i128(c, ins.a, a);
        double div = pow(10.0, (double)ins.imm);
#pragma unroll
        FOR_R o[r] = (double)a[r] / div;
        store_f64(c, ins.dst, o);
        store_valid(c, ins.dst, va);
        break;
      }
      case OP_CAST_F64_I64: {
        double a[VM_R];
        int64_t o[VM_R];
        fetch_f64(c, ins.a, a);
        uint32_t vout = va;
#pragma unroll
        FOR_R {
          double t = trunc(a[r]);
          if (!(t >= -9.2233720368547758e18 && t < 9.2233720368547758e18)) {
            vout &= ~(1u << r);
            o[r] = 0;
          } else {
            o[r] = (int64_t)t;
          }
        }
        store_i64(c, ins.dst, o);
        store_valid(c, ins.dst, vout);
        break;
      }
      case OP_CAST_I128_I64: {
        i128 a[VM_R];
        int64_t o[VM_R];
        fetch_i128(c, ins.a, a);
        i128 div = pow10_dev(ins.imm);
        uint32_t vout = va;
#pragma unroll
        FOR_R {
          i128 q = a[r] / div;
          if (!fits_i64(q)) vout &= ~(1u << r);
          o[r] = (int64_t)q;
        }
        store_i64(c, ins.dst, o);
        store_valid(c, ins.dst, vout);
        break;
      }
      case OP_CAST_F64_I128: {
        double a[VM_R];
        i128 o[VM_R];
        fetch_f64(c, ins.a, a);
        double mul = pow(10.0, (double)ins.imm);
        uint32_t vout = va;
#pragma unroll
        FOR_R {
          double t = round(a[r] * mul);
          if (!(fabs(t) < 1.7e38)) {
            vout &= ~(1u << r);
            o[r] = 0;
          } else {
            o[r] = (i128)t;
          }
        }
        store_i128(c, ins.dst, o);
        store_valid(c, ins.dst, vout);
        break;
      }
      case OP_WRAP_I64:
      case OP_NARROW_I64: {
        int64_t a[VM_R];
        fetch_i64(c, ins.a, a);
        uint32_t vout = va;
#pragma unroll
        FOR_R {
          int64_t w;
          switch (ins.aux) {
            case PH_I8: w = (int8_t)a[r]; break;
            case PH_I16: w = (int16_t)a[r]; break;
            case PH_I32: w = (int32_t)a[r]; break;
            case PH_U8: w = (uint8_t)a[r]; break;
            case PH_U16: w = (uint16_t)a[r]; break;
            case PH_U32: w = (uint32_t)a[r]; break;
            default: w = a[r];
          }
          if (ins.op == OP_NARROW_I64 && w != a[r]) vout &= ~(1u << r);
          a[r] = w;
        }
        store_i64(c, ins.dst, a);
        store_valid(c, ins.dst, vout);
        break;
      }
      case OP_SELECT: {  // dst = (cond true) ? b : dst
        uint32_t cond = fetch_bool(c, ins.a) & fetch_valid(c, ins.a);
        uint32_t vbv = fetch_valid(c, ins.b);
        uint32_t vd = fetch_valid(c, ins.dst);
        if (ins.t == VK_I128) {
          i128 b[VM_R], d[VM_R];
          fetch_i128(c, ins.b, b);
          fetch_i128(c, ins.dst, d);
#pragma unroll
          FOR_R if ((cond >> r) & 1) d[r] = b[r];
          store_i128(c, ins.dst, d);
        } else if (ins.t == VK_F64) {
          double b[VM_R], d[VM_R];
          fetch_f64(c, ins.b, b);
          fetch_f64(c, ins.dst, d);
#pragma unroll
          FOR_R if ((cond >> r) & 1) d[r] = b[r];
          store_f64(c, ins.dst, d);
        } else if (ins.t == VK_STR) {
          StrRef b[VM_R], d[VM_R];
          fetch_str(c, ins.b, b);
          fetch_str(c, ins.dst, d);
#pragma unroll
          FOR_R if ((cond >> r) & 1) d[r] = b[r];
          store_str(c, ins.dst, d);
        } else if (ins.t == VK_BOOL) {
          uint32_t b = fetch_bool(c, ins.b), d = fetch_bool(c, ins.dst);
          store_bool(c, ins.dst, (d & ~cond) | (b & cond));
        } else {
          int64_t b[VM_R], d[VM_R];
          fetch_i64(c, ins.b, b);
          fetch_i64(c, ins.dst, d);
#pragma unroll
          FOR_R if ((cond >> r) & 1) d[r] = b[r];
          store_i64(c, ins.dst, d);
        }
        store_valid(c, ins.dst, (vd & ~cond) | (vbv & cond));
        break;
      }
      case OP_MOV: {
        uint32_t v = fetch_valid(c, ins.a);
        if (ins.t == VK_I128) {
          i128 a[VM_R];
          fetch_i128(c, ins.a, a);
          store_i128(c, ins.dst, a);
        } else if (ins.t == VK_F64) {
          double a[VM_R];
          fetch_f64(c, ins.a, a);
          store_f64(c, ins.dst, a);
        } else if (ins.t == VK_STR) {
          StrRef a[VM_R];
          fetch_str(c, ins.a, a);
          store_str(c, ins.dst, a);
        } else if (ins.t == VK_BOOL) {
          store_bool(c, ins.dst, fetch_bool(c, ins.a));
        } else {
          int64_t a[VM_R];
          fetch_i64(c, ins.a, a);
          store_i64(c, ins.dst, a);
        }
        store_valid(c, ins.dst, v);
        break;
      }
      case OP_LIKE: {
        StrRef a[VM_R];
        fetch_str(c, ins.a, a);
        const uint8_t* pat = (const uint8_t*)P.imms[ins.imm].lo;
        uint32_t pn = (uint32_t)P.imms[ins.imm].hi;
        uint32_t m = 0;
#pragma unroll
        FOR_R {
          bool ok = ((c.active & va) >> r) & 1;
          bool hit = ok && like_match_dev(a[r].p, a[r].len, pat, pn);
          m |= ((ins.aux ? !hit : hit) ? 1u : 0u) << r;
        }
        store_bool(c, ins.dst, m);
        store_valid(c, ins.dst, va);
        break;
      }
      case OP_YEAR: {
        int64_t a[VM_R];
        fetch_i64(c, ins.a, a);
#pragma unroll
        FOR_R a[r] = year_of_days_dev(a[r]);
        store_i64(c, ins.dst, a);
        store_valid(c, ins.dst, va);
        break;
      }
      case OP_SUBSTR: {
        StrRef a[VM_R];
        int64_t st[VM_R], ln[VM_R];
        fetch_str(c, ins.a, a);
        fetch_i64(c, ins.b, st);
        bool has_len = ins.imm >= 0;
        if (has_len) {
          Operand lo;
          lo.kind = OPD_IMM;
          lo.vk = VK_I64;
          lo.idx = (uint16_t)ins.imm;
          fetch_i64(c, lo, ln);
        }
#pragma unroll
        FOR_R {
          int64_t s0 = st[r] - 1, e0 = has_len ? s0 + ln[r] : (int64_t)a[r].len;
          if (has_len && ln[r] < 0 && ((live >> r) & 1)) raise(P, 3);
          if (s0 < 0) s0 = 0;
          if (e0 > (int64_t)a[r].len) e0 = a[r].len;
          if (e0 > s0) {
            a[r].p += s0;
            a[r].len = (uint32_t)(e0 - s0);
          } else {
            a[r].len = 0;
          }
        }
        store_str(c, ins.dst, a);
        store_valid(c, ins.dst, va & vb);
        break;
      }
      case OP_HASH:
      case OP_HASH_COMBINE: {
        uint32_t v = fetch_valid(c, ins.a);
        uint64_t h[VM_R];
        if (ins.t == VK_F64) {
          double a[VM_R];
          fetch_f64(c, ins.a, a);
#pragma unroll
          FOR_R h[r] = hash_f64(a[r]);
        } else if (ins.t == VK_I128) {
          i128 a[VM_R];
          fetch_i128(c, ins.a, a);
#pragma unroll
          FOR_R h[r] = hash_i128(lo64(a[r]), hi64(a[r]));
        } else if (ins.t == VK_STR) {
          StrRef a[VM_R];
          fetch_str(c, ins.a, a);
#pragma unroll
          FOR_R h[r] = ((c.active & v) >> r) & 1 ? hash_bytes(a[r].p, a[r].len) : 0;
        } else {
          int64_t a[VM_R];
          fetch_i64(c, ins.a, a);
#pragma unroll
          FOR_R h[r] = hash_i64(a[r]);
        }
        int64_t o[VM_R];
        if (ins.op == OP_HASH) {
#pragma unroll
          FOR_R o[r] = ((v >> r) & 1) ? (int64_t)h[r] : 0;
        } else {
          fetch_i64(c, ins.dst, o);
#pragma unroll
          FOR_R if ((v >> r) & 1) o[r] = (int64_t)combine_hashes(h[r], (uint64_t)o[r]);
        }
        store_i64(c, ins.dst, o);
        break;
      }
      case OP_MOD_U64: {
        int64_t a[VM_R];
        fetch_i64(c, ins.a, a);
        uint64_t m = P.imms[ins.imm].lo;
#pragma unroll
        FOR_R a[r] = (int64_t)((uint64_t)a[r] % m);
        store_i64(c, ins.dst, a);
        break;
      }
      case OP_FILTER: {
        uint32_t m = fetch_bool(c, ins.a) & fetch_valid(c, ins.a);
        c.active &= m;
        break;
      }
      default: break;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tile loading
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t col_tile_bytes(const ColDesc& cd, int tile_rows) {
  uint32_t b = (uint32_t)tile_rows * cd.width;
  if (cd.phys == PH_UTF8) b += 16;  // one extra offset (+ padding to a 16-byte multiple)
  return b;
}

// TMA path: one elected thread issues a bulk copy per staged column; completion is signalled on
// the stage's mbarrier through complete_tx.
__device__ __forceinline__ void issue_tile_tma(const Program& P, uint8_t* stage, uint64_t* bar, int64_t row0, int tile_rows) {
  uint32_t total = 0;
  for (int i = 0; i < P.n_cols; i++) {
    const ColDesc& cd = P.cols[i];
    total += col_tile_bytes(cd, tile_rows);
    if (cd.valid) total += (uint32_t)tile_rows;
  }
  mbar_expect_tx(bar, total);
  for (int i = 0; i < P.n_cols; i++) {
    const ColDesc& cd = P.cols[i];
    bulk_g2s(stage + cd.smem_off, (const uint8_t*)cd.data + row0 * cd.width, col_tile_bytes(cd, tile_rows), bar);
    if (cd.valid) bulk_g2s(stage + cd.valid_smem_off, cd.valid + row0, (uint32_t)tile_rows, bar);
  }
}

// Fallback path (ragged last tile, unaligned slices): cooperative loads, zero fill past the end.
__device__ void load_tile_coop(const Program& P, uint8_t* stage, int64_t row0, int rows, int tile_rows, int tid, int B) {
  for (int i = 0; i < P.n_cols; i++) {
    const ColDesc& cd = P.cols[i];
    uint8_t* dst = stage + cd.smem_off;
    if (cd.phys == PH_UTF8) {
      const int32_t* src = (const int32_t*)cd.data + row0;
      int32_t* d = (int32_t*)dst;
      for (int k = tid; k <= tile_rows; k += B) d[k] = src[k <= rows ? k : rows];
    } else {
      const uint8_t* src = (const uint8_t*)cd.data + row0 * cd.width;
      uint32_t nb = (uint32_t)rows * cd.width, tb = (uint32_t)tile_rows * cd.width;
      if ((cd.width & 3) == 0 && (((uintptr_t)src) & 3) == 0) {
        const uint32_t* s4 = (const uint32_t*)src;
        uint32_t* d4 = (uint32_t*)dst;
        for (uint32_t k = tid; k < tb / 4; k += B) d4[k] = (k * 4 < nb) ? s4[k] : 0u;
      } else {
        for (uint32_t k = tid; k < tb; k += B) dst[k] = (k < nb) ? src[k] : 0;
      }
    }
    if (cd.valid) {
      uint8_t* dv = stage + cd.valid_smem_off;
      for (int k = tid; k < tile_rows; k += B) dv[k] = (k < rows) ? cd.valid[row0 + k] : 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Sink: materialise (FilterExec compaction + ProjectionExec outputs)
// ------------------------------------------------------------------------------------------------
__device__ void sink_materialize(Ctx& c, uint32_t* warp_tot /*[VM_R][32]*/, unsigned long long* tile_base_sh) {
  const Program& P = *c.P;
  const int lane = c.tid & 31, warp = c.tid >> 5, nwarps = c.B >> 5;
  uint32_t lane_pre[VM_R];
#pragma unroll
  FOR_R {
    uint32_t m = __ballot_sync(0xFFFFFFFFu, (c.active >> r) & 1);
    lane_pre[r] = __popc(m & ((1u << lane) - 1));
    if (lane == 0) warp_tot[r * 32 + warp] = __popc(m);
  }
  __syncthreads();
  // position of row (r, warp, lane) = sum of all rows with smaller (r, warp) + lane_pre
  uint32_t before[VM_R];
  uint32_t run = 0;
#pragma unroll
  FOR_R {
    uint32_t mine = 0;
    for (int w = 0; w < nwarps; w++) {
      if (w == warp) mine = run;
      run += warp_tot[r * 32 + w];
    }
    before[r] = mine;
  }
  if (c.tid == 0) *tile_base_sh = run ? atomicAdd(&P.status->out_rows, (unsigned long long)run) : 0ull;
  __syncthreads();
  const unsigned long long base = *tile_base_sh;
  if (run == 0) return;
  for (int j = 0; j < P.n_out; j++) {
    const OutCol oc = P.out[j];
    uint32_t v = oc.valid ? fetch_valid(c, oc.src) : 0xFFFFFFFFu;
    switch (oc.src.vk) {
      case VK_I128: {
        i128 a[VM_R];
        fetch_i128(c, oc.src, a);
#pragma unroll
        FOR_R if ((c.active >> r) & 1) ((ulonglong2*)oc.data)[base + before[r] + lane_pre[r]] = make_ulonglong2(lo64(a[r]), hi64(a[r]));
        break;
      }
      case VK_F64: {
        double a[VM_R];
        fetch_f64(c, oc.src, a);
#pragma unroll
        FOR_R if ((c.active >> r) & 1) {
          unsigned long long pos = base + before[r] + lane_pre[r];
          if (oc.phys == PH_F32) ((float*)oc.data)[pos] = (float)a[r];
          else ((double*)oc.data)[pos] = a[r];
        }
        break;
      }
      case VK_STR: {
        StrRef a[VM_R];
        fetch_str(c, oc.src, a);
#pragma unroll
        FOR_R if ((c.active >> r) & 1) {
          bool ok = (v >> r) & 1;
          ((ulonglong2*)oc.data)[base + before[r] + lane_pre[r]] = make_ulonglong2(ok ? (unsigned long long)a[r].p : 0ull, ok ? (unsigned long long)a[r].len : 0ull);
        }
        break;
      }
      default: {
        int64_t a[VM_R];
        if (oc.src.vk == VK_BOOL) {
          uint32_t m = fetch_bool(c, oc.src);
#pragma unroll
          FOR_R a[r] = (m >> r) & 1;
        } else {
          fetch_i64(c, oc.src, a);
        }
#pragma unroll
        FOR_R if ((c.active >> r) & 1) {
          unsigned long long pos = base + before[r] + lane_pre[r];
          switch (oc.phys) {
            case PH_I8:
            case PH_U8:
            case PH_BOOL8: ((int8_t*)oc.data)[pos] = (int8_t)a[r]; break;
            case PH_I16:
            case PH_U16: ((int16_t*)oc.data)[pos] = (int16_t)a[r]; break;
            case PH_I32:
            case PH_U32: ((int32_t*)oc.data)[pos] = (int32_t)a[r]; break;
            default: ((int64_t*)oc.data)[pos] = a[r];
          }
        }
      }
    }
    if (oc.valid) {
#pragma unroll
      FOR_R if ((c.active >> r) & 1) oc.valid[base + before[r] + lane_pre[r]] = (v >> r) & 1;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Aggregate hash table (global memory, open addressing, linear probing)
// ------------------------------------------------------------------------------------------------
struct KeyVal {
  unsigned long long w0, w1;
  unsigned char valid;
  unsigned char vk;
};

__device__ __forceinline__ bool key_equal(const KeyVal& a, unsigned long long w0, unsigned long long w1, unsigned char valid) {
  if (a.valid != valid) return false;
  if (!valid) return true;
  if (a.vk == VK_STR) {
    StrRef x{(const uint8_t*)a.w0, (uint32_t)a.w1}, y{(const uint8_t*)w0, (uint32_t)w1};
    return str_eq(x, y);
  }
  if (a.vk == VK_I128) return a.w0 == w0 && a.w1 == w1;
  if (a.vk == VK_F64) {
    double p = __longlong_as_double((long long)a.w0), q = __longlong_as_double((long long)w0);
    return (p == q) || (p != p && q != q);
  }
  return a.w0 == w0;
}

// find-or-insert; returns slot or ~0ull on overflow
__device__ unsigned long long table_upsert(const AggTable& T, int n_keys, unsigned long long h, const KeyVal* kv) {
  const unsigned long long mask = T.cap - 1;
  unsigned long long slot = h & mask;
  for (unsigned long long probes = 0; probes < T.cap; probes++) {
    unsigned int st = *(volatile unsigned int*)&T.state[slot];
    if (st == 0) {
      unsigned int old = atomicCAS(&T.state[slot], 0u, 1u);
      if (old == 0) {
        T.hash[slot] = h;
        for (int k = 0; k < n_keys; k++) {
          T.keys[((unsigned long long)k * T.cap + slot) * 2 + 0] = kv[k].w0;
          T.keys[((unsigned long long)k * T.cap + slot) * 2 + 1] = kv[k].w1;
          T.key_valid[(unsigned long long)k * T.cap + slot] = kv[k].valid;
        }
        __threadfence();
        atomicExch(&T.state[slot], 2u);
        unsigned int ng = atomicAdd(T.n_groups, 1u);
        if ((unsigned long long)ng * 4 > T.cap * 3) return ~0ull;  // load factor > 0.75
        return slot;
      }
      st = old;
    }
    while (st == 1) st = *(volatile unsigned int*)&T.state[slot];
    __threadfence();
    if (*(volatile unsigned long long*)&T.hash[slot] == h) {
      bool eq = true;
      for (int k = 0; k < n_keys && eq; k++) {
        unsigned long long w0 = *(volatile unsigned long long*)&T.keys[((unsigned long long)k * T.cap + slot) * 2 + 0];
        unsigned long long w1 = *(volatile unsigned long long*)&T.keys[((unsigned long long)k * T.cap + slot) * 2 + 1];
        unsigned char vl = *(volatile unsigned char*)&T.key_valid[(unsigned long long)k * T.cap + slot];
        eq = key_equal(kv[k], w0, w1, vl);
      }
      if (eq) return slot;
    }
    slot = (slot + 1) & mask;
  }
  return ~0ull;
}

__device__ __forceinline__ void acc_add_i128_atomic(unsigned long long* cell, i128 v) {
  unsigned long long lo = lo64(v), hi = hi64(v);
  unsigned long long old = atomicAdd(&cell[0], lo);
  unsigned long long carry = (old + lo) < old ? 1ull : 0ull;
  if (hi + carry) atomicAdd(&cell[1], hi + carry);
}

__device__ __forceinline__ void table_lock(const AggTable& T, unsigned long long slot) {
  while (atomicCAS(&T.lock[slot], 0u, 1u) != 0u) {
  }
  __threadfence();
}
__device__ __forceinline__ void table_unlock(const AggTable& T, unsigned long long slot) {
  __threadfence();
  atomicExch(&T.lock[slot], 0u);
}

// value of an accumulator source operand for the R rows of this thread, as 128-bit / f64 / order key
__device__ __forceinline__ void fetch_acc_vals(const Ctx& c, const AccDesc& ad, i128 vi[VM_R], double vf[VM_R]) {
  if (ad.kind == ACC_COUNT_STAR || ad.kind == ACC_COUNT) return;
  if (ad.kind == ACC_SUM_F64 || ad.kind == ACC_MIN_F64 || ad.kind == ACC_MAX_F64) fetch_f64(c, ad.src, vf);
  else fetch_i128(c, ad.src, vi);
}

// per-row path (high cardinality): every live row upserts its group and updates with atomics
__device__ void sink_agg_global(Ctx& c) {
  const Program& P = *c.P;
  const AggTable& T = P.table;
  if (*(volatile unsigned int*)&P.status->overflow) return;
  unsigned long long slots[VM_R];
  {
    int64_t h[VM_R];
    if (P.n_keys) fetch_i64(c, P.key_hash, h);
    KeyVal kv[VM_R][VM_MAX_KEYS];
    for (int k = 0; k < P.n_keys; k++) {
      Operand ko = P.keys[k];
      uint32_t v = fetch_valid(c, ko);
      if (ko.vk == VK_STR) {
        StrRef s[VM_R];
        fetch_str(c, ko, s);
#pragma unroll
        FOR_R {
          kv[r][k].w0 = (unsigned long long)s[r].p;
          kv[r][k].w1 = s[r].len;
        }
      } else if (ko.vk == VK_I128) {
        i128 a[VM_R];
        fetch_i128(c, ko, a);
#pragma unroll
        FOR_R {
          kv[r][k].w0 = lo64(a[r]);
          kv[r][k].w1 = hi64(a[r]);
        }
      } else if (ko.vk == VK_F64) {
        double a[VM_R];
        fetch_f64(c, ko, a);
#pragma unroll
        FOR_R {
          kv[r][k].w0 = (unsigned long long)__double_as_longlong(a[r] == 0.0 ? 0.0 : a[r]);
          kv[r][k].w1 = 0;
        }
      } else {
        int64_t a[VM_R];
        if (ko.vk == VK_BOOL) {
          uint32_t m = fetch_bool(c, ko);
#pragma unroll
          FOR_R a[r] = (m >> r) & 1;
        } else {
          fetch_i64(c, ko, a);
        }
#pragma unroll
        FOR_R {
          kv[r][k].w0 = (unsigned long long)a[r];
          kv[r][k].w1 = 0;
        }
      }
#pragma unroll
      FOR_R {
        kv[r][k].valid = (v >> r) & 1;
        kv[r][k].vk = ko.vk;
        if (!kv[r][k].valid) kv[r][k].w0 = kv[r][k].w1 = 0;
      }
    }
#pragma unroll
    FOR_R {
      slots[r] = 0;
      if ((c.active >> r) & 1) {
        unsigned long long s = table_upsert(T, P.n_keys, P.n_keys ? (unsigned long long)h[r] : 0ull, kv[r]);
        if (s == ~0ull) {
          atomicExch(&P.status->overflow, 1u);
          c.active &= ~(1u << r);
        } else {
          slots[r] = s;
        }
      }
    }
  }
  for (int a = 0; a < P.n_acc; a++) {
    const AccDesc ad = P.acc[a];
    unsigned long long* col = T.acc + (unsigned long long)a * T.cap * 2;
    uint32_t v = (ad.kind == ACC_COUNT_STAR) ? 0xFFFFFFFFu : (ad.nullable ? fetch_valid(c, ad.src) : 0xFFFFFFFFu);
    i128 vi[VM_R];
    double vf[VM_R];
    fetch_acc_vals(c, ad, vi, vf);
#pragma unroll
    FOR_R {
      if (!(((c.active & v) >> r) & 1)) continue;
      unsigned long long* cell = col + slots[r] * 2;
      switch (ad.kind) {
        case ACC_COUNT_STAR:
        case ACC_COUNT: atomicAdd(&cell[0], 1ull); break;
        case ACC_SUM_I128: acc_add_i128_atomic(cell, vi[r]); break;
        case ACC_SUM_F64: atomicAdd((double*)&cell[0], vf[r]); break;
        case ACC_MIN_F64: atomicMin((long long*)&cell[0], f64_order_key(vf[r])); break;
        case ACC_MAX_F64: atomicMax((long long*)&cell[0], f64_order_key(vf[r])); break;
        default: {  // 128-bit min/max under the slot lock
          table_lock(T, slots[r]);
          i128 cur = make_i128(cell[0], cell[1]);
          bool take = ad.kind == ACC_MIN_I128 ? vi[r] < cur : vi[r] > cur;
          if (take) {
            cell[0] = lo64(vi[r]);
            cell[1] = hi64(vi[r]);
          }
          table_unlock(T, slots[r]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Register-resident aggregate sink: <= VM_REG_GROUPS groups, <= VM_REG_ACC accumulators.
// Every thread keeps the full (group x accumulator) matrix in registers: no atomics, no shared
// memory traffic in the per-row path.  Group ids are dense per CTA (tiny shared-memory key table).
// ------------------------------------------------------------------------------------------------
struct RegGroupTable {  // shared memory
  unsigned long long hash[VM_REG_GROUPS];
  unsigned int state[VM_REG_GROUPS];
  unsigned long long key_w0[VM_REG_GROUPS][VM_MAX_KEYS];
  unsigned long long key_w1[VM_REG_GROUPS][VM_MAX_KEYS];
  unsigned char key_valid[VM_REG_GROUPS][VM_MAX_KEYS];
  unsigned int n_groups;
};

struct Acc128 {
  uint64_t lo, hi;
};

template <int G>
struct RegAggState {
  Acc128 acc[G][VM_REG_ACC];
};

__device__ __forceinline__ void acc128_add(Acc128& a, i128 v) {
  uint64_t lo = a.lo + lo64(v);
  a.hi += hi64(v) + (lo < a.lo ? 1 : 0);
  a.lo = lo;
}

template <int G>
__device__ __forceinline__ void reg_agg_init(const Program& P, RegAggState<G>& S) {
#pragma unroll
  for (int g = 0; g < G; g++)
#pragma unroll
    for (int a = 0; a < VM_REG_ACC; a++) {
      uint64_t lo = 0, hi = 0;
      if (a < P.n_acc) {
        switch (P.acc[a].kind) {
          case ACC_MIN_I128: lo = ~0ull; hi = 0x7FFFFFFFFFFFFFFFull; break;
          case ACC_MAX_I128: lo = 0; hi = 0x8000000000000000ull; break;
          case ACC_MIN_F64: lo = 0x7FFFFFFFFFFFFFFFull; break;
          case ACC_MAX_F64: lo = 0x8000000000000000ull; break;
          default: break;
        }
      }
      S.acc[g][a].lo = lo;
      S.acc[g][a].hi = hi;
    }
}

template <int G>
__device__ __forceinline__ void sink_agg_reg(Ctx& c, RegAggState<G>& S, RegGroupTable* gt) {
  const Program& P = *c.P;
  uint32_t gid[VM_R];
#pragma unroll
  FOR_R gid[r] = 0;
  if (G > 1) {
    int64_t h[VM_R];
    fetch_i64(c, P.key_hash, h);
    // gather this thread's key values once (needed for verification / insertion)
    KeyVal kv[VM_R][VM_MAX_KEYS];
    for (int k = 0; k < P.n_keys; k++) {
      Operand ko = P.keys[k];
      uint32_t v = fetch_valid(c, ko);
      if (ko.vk == VK_STR) {
        StrRef s[VM_R];
        fetch_str(c, ko, s);
#pragma unroll
        FOR_R {
          kv[r][k].w0 = (unsigned long long)s[r].p;
          kv[r][k].w1 = s[r].len;
        }
      } else if (ko.vk == VK_I128) {
        i128 a[VM_R];
        fetch_i128(c, ko, a);
#pragma unroll
        FOR_R {
          kv[r][k].w0 = lo64(a[r]);
          kv[r][k].w1 = hi64(a[r]);
        }
      } else if (ko.vk == VK_F64) {
        double a[VM_R];
        fetch_f64(c, ko, a);
#pragma unroll
        FOR_R {
          kv[r][k].w0 = (unsigned long long)__double_as_longlong(a[r] == 0.0 ? 0.0 : a[r]);
          kv[r][k].w1 = 0;
        }
      } else {
        int64_t a[VM_R];
        if (ko.vk == VK_BOOL) {
          uint32_t m = fetch_bool(c, ko);
#pragma unroll
          FOR_R a[r] = (m >> r) & 1;
        } else {
          fetch_i64(c, ko, a);
        }
#pragma unroll
        FOR_R {
          kv[r][k].w0 = (unsigned long long)a[r];
          kv[r][k].w1 = 0;
        }
      }
#pragma unroll
      FOR_R {
        kv[r][k].valid = (v >> r) & 1;
        kv[r][k].vk = ko.vk;
        if (!kv[r][k].valid) kv[r][k].w0 = kv[r][k].w1 = 0;
      }
    }
#pragma unroll
    FOR_R {
      if (!((c.active >> r) & 1)) continue;
      const unsigned long long hh = (unsigned long long)h[r];
      int found = -1;
      for (int g = 0; g < G && found < 0; g++) {
        unsigned int st = *(volatile unsigned int*)&gt->state[g];
        if (st == 0) {
          unsigned int old = atomicCAS(&gt->state[g], 0u, 1u);
          if (old == 0) {
            gt->hash[g] = hh;
            for (int k = 0; k < P.n_keys; k++) {
              gt->key_w0[g][k] = kv[r][k].w0;
              gt->key_w1[g][k] = kv[r][k].w1;
              gt->key_valid[g][k] = kv[r][k].valid;
            }
            __threadfence_block();
            atomicExch(&gt->state[g], 2u);
            atomicAdd(&gt->n_groups, 1u);
            found = g;
            break;
          }
          st = old;
        }
        while (st == 1) st = *(volatile unsigned int*)&gt->state[g];
        __threadfence_block();
        if (*(volatile unsigned long long*)&gt->hash[g] == hh) {
          bool eq = true;
          for (int k = 0; k < P.n_keys && eq; k++)
            eq = key_equal(kv[r][k], *(volatile unsigned long long*)&gt->key_w0[g][k], *(volatile unsigned long long*)&gt->key_w1[g][k],
                           *(volatile unsigned char*)&gt->key_valid[g][k]);
          if (eq) found = g;
        }
      }
      if (found < 0) {
        atomicExch(&P.status->overflow, 1u);  // a (G+1)-th group: the host re-runs with the global sink
        c.active &= ~(1u << r);
      } else {
        gid[r] = (uint32_t)found;
      }
    }
  }
  // accumulate: static register indexing only; the kind switch is hoisted out of the row/group loops
#pragma unroll
  for (int a = 0; a < VM_REG_ACC; a++) {
    if (a >= P.n_acc) break;
    const AccDesc ad = P.acc[a];
    uint32_t v = (ad.kind == ACC_COUNT_STAR) ? 0xFFFFFFFFu : (ad.nullable ? fetch_valid(c, ad.src) : 0xFFFFFFFFu);
    v &= c.active;
    if (ad.kind == ACC_SUM_I128 || ad.kind == ACC_COUNT || ad.kind == ACC_COUNT_STAR) {
      i128 vi[VM_R];
      if (ad.kind == ACC_SUM_I128) {
        fetch_i128(c, ad.src, vi);
      } else {
#pragma unroll
        FOR_R vi[r] = 1;
      }
#pragma unroll
      FOR_R {
#pragma unroll
        for (int g = 0; g < G; g++)
          if (((v >> r) & 1) && (G == 1 || gid[r] == (uint32_t)g)) acc128_add(S.acc[g][a], vi[r]);
      }
    } else if (ad.kind == ACC_SUM_F64) {
      double vf[VM_R];
      fetch_f64(c, ad.src, vf);
#pragma unroll
      FOR_R {
#pragma unroll
        for (int g = 0; g < G; g++)
          if (((v >> r) & 1) && (G == 1 || gid[r] == (uint32_t)g))
            S.acc[g][a].lo = (uint64_t)__double_as_longlong(__longlong_as_double((long long)S.acc[g][a].lo) + vf[r]);
      }
    } else if (ad.kind == ACC_MIN_F64 || ad.kind == ACC_MAX_F64) {
      double vf[VM_R];
      fetch_f64(c, ad.src, vf);
      const bool is_min = ad.kind == ACC_MIN_F64;
#pragma unroll
      FOR_R {
        long long k = f64_order_key(vf[r]);
#pragma unroll
        for (int g = 0; g < G; g++)
          if (((v >> r) & 1) && (G == 1 || gid[r] == (uint32_t)g)) {
            long long cur = (long long)S.acc[g][a].lo;
            if (is_min ? k < cur : k > cur) S.acc[g][a].lo = (uint64_t)k;
          }
      }
    } else {
      i128 vi[VM_R];
      fetch_i128(c, ad.src, vi);
      const bool is_min = ad.kind == ACC_MIN_I128;
#pragma unroll
      FOR_R {
#pragma unroll
        for (int g = 0; g < G; g++)
          if (((v >> r) & 1) && (G == 1 || gid[r] == (uint32_t)g)) {
            i128 cur = make_i128(S.acc[g][a].lo, S.acc[g][a].hi);
            if (is_min ? vi[r] < cur : vi[r] > cur) {
              S.acc[g][a].lo = lo64(vi[r]);
              S.acc[g][a].hi = hi64(vi[r]);
            }
          }
      }
    }
  }
}

__device__ __forceinline__ Acc128 acc_combine(int kind, Acc128 x, Acc128 y) {
  switch (kind) {
    case ACC_COUNT_STAR:
    case ACC_COUNT: x.lo += y.lo; return x;
    case ACC_SUM_I128: {
      uint64_t lo = x.lo + y.lo;
      x.hi += y.hi + (lo < x.lo ? 1 : 0);
      x.lo = lo;
      return x;
    }
    case ACC_SUM_F64: x.lo = (uint64_t)__double_as_longlong(__longlong_as_double((long long)x.lo) + __longlong_as_double((long long)y.lo)); return x;
    case ACC_MIN_F64: return (long long)y.lo < (long long)x.lo ? y : x;
    case ACC_MAX_F64: return (long long)y.lo > (long long)x.lo ? y : x;
    case ACC_MIN_I128: return make_i128(y.lo, y.hi) < make_i128(x.lo, x.hi) ? y : x;
    default: return make_i128(y.lo, y.hi) > make_i128(x.lo, x.hi) ? y : x;
  }
}

// End of kernel: reduce the per-thread matrices over the CTA and merge them into the global table.
template <int G>
__device__ __forceinline__ void reg_agg_flush(const Program& P, RegAggState<G>& S, RegGroupTable* gt, Acc128* scratch /*[nwarps][G][VM_REG_ACC]*/, int tid, int B) {
  const int lane = tid & 31, warp = tid >> 5, nwarps = B >> 5;
#pragma unroll
  for (int g = 0; g < G; g++)
#pragma unroll
    for (int a = 0; a < VM_REG_ACC; a++) {
      if (a >= P.n_acc) break;
      Acc128 x = S.acc[g][a];
      const int kind = P.acc[a].kind;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        Acc128 y;
        y.lo = __shfl_xor_sync(0xFFFFFFFFu, x.lo, off);
        y.hi = __shfl_xor_sync(0xFFFFFFFFu, x.hi, off);
        x = acc_combine(kind, x, y);
      }
      if (lane == 0) scratch[(warp * G + g) * VM_REG_ACC + a] = x;
    }
  __syncthreads();
  const AggTable& T = P.table;
  const unsigned int ng = (G == 1) ? 1u : gt->n_groups;
  // one thread per (group): merge under the slot lock
  if (tid < (int)ng && tid < G) {
    const int g = tid;
    KeyVal kv[VM_MAX_KEYS];
    unsigned long long h = 0;
    if (G > 1) {
      h = gt->hash[g];
      for (int k = 0; k < P.n_keys; k++) {
        kv[k].w0 = gt->key_w0[g][k];
        kv[k].w1 = gt->key_w1[g][k];
        kv[k].valid = gt->key_valid[g][k];
        kv[k].vk = P.keys[k].vk;
      }
    }
    // a scalar aggregate (no keys) with zero live rows still owns its single output group
    unsigned long long slot = table_upsert(T, P.n_keys, h, kv);
    if (slot == ~0ull) {
      atomicExch(&P.status->overflow, 1u);
      return;
    }
    table_lock(T, slot);
    for (int a = 0; a < P.n_acc; a++) {
      Acc128 x = scratch[(0 * G + g) * VM_REG_ACC + a];
      for (int w = 1; w < nwarps; w++) x = acc_combine(P.acc[a].kind, x, scratch[(w * G + g) * VM_REG_ACC + a]);
      unsigned long long* cell = T.acc + ((unsigned long long)a * T.cap + slot) * 2;
      Acc128 cur;
      cur.lo = cell[0];
      cur.hi = cell[1];
      cur = acc_combine(P.acc[a].kind, cur, x);
      cell[0] = cur.lo;
      cell[1] = cur.hi;
    }
    table_unlock(T, slot);
  }
}

// ------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------
template <int SINK, int G>
__global__ void __launch_bounds__(256, 1) pipeline_kernel(const __grid_constant__ Program P) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[VM_MAX_STAGES];
  __shared__ uint32_t warp_tot[VM_R * 32];
  __shared__ unsigned long long tile_base_sh;
  __shared__ RegGroupTable gtable;

  const int tid = threadIdx.x, B = blockDim.x;
  const int TILE = B * VM_R;
  const int64_t n_rows = P.n_rows;
  const int64_t n_tiles = (n_rows + TILE - 1) / TILE;
  const int S = (int)P.n_stages;
  uint8_t* stage0 = smem;
  uint8_t* regs = smem + (size_t)S * P.stage_bytes;

  if (tid == 0) {
    for (int s = 0; s < S; s++) mbar_init(&full_bar[s], 1);
    mbar_fence_init();
  }
  if (SINK == SINK_AGG_REG && tid < VM_REG_GROUPS) {
    gtable.state[tid] = 0;
    gtable.hash[tid] = 0;
    if (tid == 0) gtable.n_groups = 0;
  }
  __syncthreads();

  RegAggState<G> S_reg;
  if (SINK == SINK_AGG_REG) reg_agg_init<G>(P, S_reg);

  Ctx c;
  c.P = &P;
  c.regs = regs;
  c.tid = tid;
  c.B = B;

  // tiles are dealt round-robin: tile(k) = blockIdx.x + k * gridDim.x
  auto tile_of = [&](int64_t k) { return (int64_t)blockIdx.x + k * (int64_t)gridDim.x; };
  auto tile_is_tma = [&](int64_t t) { return P.use_tma && (t + 1) * (int64_t)TILE <= n_rows; };

  if (tid == 0) {
    for (int k = 0; k < S - 1; k++) {
      int64_t t = tile_of(k);
      if (t < n_tiles && tile_is_tma(t)) issue_tile_tma(P, stage0 + (size_t)(k % S) * P.stage_bytes, &full_bar[k % S], t * TILE, TILE);
    }
  }
  uint32_t phase_bits = 0;
  for (int64_t k = 0;; k++) {
    const int64_t t = tile_of(k);
    if (t >= n_tiles) break;
    const int s = (int)(k % S);
    uint8_t* stage = stage0 + (size_t)s * P.stage_bytes;
    // prefetch tile k+S-1 into the buffer released at the end of iteration k-1
    if (tid == 0) {
      const int64_t kn = k + S - 1, tn = tile_of(kn);
      if (tn < n_tiles && tile_is_tma(tn)) issue_tile_tma(P, stage0 + (size_t)(kn % S) * P.stage_bytes, &full_bar[kn % S], tn * TILE, TILE);
    }
    const int64_t row0 = t * TILE;
    const int rows = (int)((n_rows - row0) < TILE ? (n_rows - row0) : TILE);
    if (tile_is_tma(t)) {
      mbar_wait(&full_bar[s], (phase_bits >> s) & 1);
      phase_bits ^= 1u << s;
    } else {
      load_tile_coop(P, stage, row0, rows, TILE, tid, B);
      __syncthreads();
    }
    c.stage = stage;
    c.tile_base = row0;
    c.active = 0;
#pragma unroll
    FOR_R if (r * B + tid < rows) c.active |= 1u << r;
    run_program(c);
    if (SINK == SINK_MATERIALIZE) {
      sink_materialize(c, warp_tot, &tile_base_sh);
    } else if (SINK == SINK_AGG_GLOBAL) {
      sink_agg_global(c);
    } else {
      sink_agg_reg<G>(c, S_reg, &gtable);
    }
    if (SINK != SINK_MATERIALIZE) {
      uint32_t cnt = __popc(c.active);
      cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
      if ((tid & 31) == 0 && cnt) atomicAdd(&P.status->in_active, (unsigned long long)cnt);
    }
    // everyone is done with this stage buffer (and the VM registers); the sink-overflow flag is
    // sampled CTA-uniformly so that all threads leave the loop together
    if (__syncthreads_or(SINK != SINK_MATERIALIZE && *(volatile unsigned int*)&P.status->overflow != 0)) {
      // drain bulk copies that are still in flight before the CTA may exit
      for (int64_t kk = k + 1; kk < k + S; kk++) {
        const int64_t tt = tile_of(kk);
        if (tt < n_tiles && tile_is_tma(tt)) {
          mbar_wait(&full_bar[kk % S], (phase_bits >> (kk % S)) & 1);
          phase_bits ^= 1u << (kk % S);
        }
      }
      break;
    }
  }
  if (SINK == SINK_AGG_GLOBAL && P.n_keys == 0 && blockIdx.x == 0 && tid == 0) {
    // a scalar aggregate owns exactly one output group even if no row survived
    KeyVal none[1];
    if (table_upsert(P.table, 0, 0ull, none) == ~0ull) atomicExch(&P.status->overflow, 1u);
  }
  if (SINK == SINK_AGG_REG) {
    __syncthreads();
    // scalar aggregates emit their single group even when no CTA saw a row: CTA 0 always flushes
    bool has_rows = tile_of(0) < n_tiles;
    if (has_rows || (G == 1 && blockIdx.x == 0)) reg_agg_flush<G>(P, S_reg, &gtable, (Acc128*)regs, tid, B);
  }
}

// ------------------------------------------------------------------------------------------------
// Host launcher
// ------------------------------------------------------------------------------------------------
template <int SINK, int G>
static cudaError_t launch_one(const Program& P, int grid, int block, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(pipeline_kernel<SINK, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  pipeline_kernel<SINK, G><<<grid, block, smem, st>>>(P);
  return cudaGetLastError();
}

cudaError_t launch_pipeline(const Program& P, int reg_groups, int grid, int block, size_t smem, cudaStream_t st) {
  switch (P.sink) {
    case SINK_MATERIALIZE: return launch_one<SINK_MATERIALIZE, 1>(P, grid, block, smem, st);
    case SINK_AGG_GLOBAL: return launch_one<SINK_AGG_GLOBAL, 1>(P, grid, block, smem, st);
    default:
      if (reg_groups <= 1) return launch_one<SINK_AGG_REG, 1>(P, grid, block, smem, st);
      return launch_one<SINK_AGG_REG, VM_REG_GROUPS>(P, grid, block, smem, st);
  }
}

}  // namespace b200
