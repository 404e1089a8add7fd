// Fused scan -> FilterExec -> ProjectionExec (decimal products) -> AggregateExec(Partial) kernel.
//
// Included by pipeline.cu (shares its __constant__ program, PTX wrappers and the group-table code).
// Reference operators replaced: the per-partition pipeline DataFusion builds for TPC-H q1/q6 inside
// ShuffleWriterExec's input (ballista/core/src/execution_plans/shuffle_writer.rs:150-260 drives it;
// the operators themselves are [EXT] datafusion FilterExec/ProjectionExec/AggregateExec).
//
// Design
//   * one CTA per SM, 16 warps; every WARP owns a private ring of S stage buffers in shared memory
//     and streams its own tiles (32*R rows) with 1-D TMA bulk copies signalled on per-warp
//     mbarriers.  There is no CTA-wide barrier on the data path: a warp only ever waits for its
//     own bytes.
//   * all arithmetic is in registers; aggregate partials are exact int64 per (group, accumulator)
//     per thread (|addend| < 2^46 checked per tile, rare large addends go straight to the global
//     table), reduced over the CTA and merged into the global table once at the end.
//   * the code is compiled per FusedShape (program.h): widths, compare operators, product kinds and
//     accumulator sources are template constants for the listed shapes; shape (0,0) reads them from
//     constant memory so that any matching program still runs.
#pragma once

namespace b200 {

template <uint64_t SA, uint64_t SB>
struct FusedX {
  static constexpr bool ST = (SA >> 63) != 0;
  static __device__ __forceinline__ int nf() { return ST ? (int)(SA & 7) : c_fused.n_filters; }
  static __device__ __forceinline__ uint32_t fw(int i) { return ST ? fused_wbytes((SA >> (3 + 5 * i)) & 3) : (uint32_t)c_fused.f[i].w; }
  static __device__ __forceinline__ int fop(int i) { return ST ? (int)((SA >> (5 + 5 * i)) & 7) : (int)c_fused.f[i].op - (int)OP_CMP_EQ; }
  static __device__ __forceinline__ int nk() { return ST ? (int)((SA >> 33) & 3) : c_fused.n_keys; }
  static __device__ __forceinline__ int kkind(int k) { return ST ? (int)((SA >> (35 + 3 * k)) & 1) : (int)c_fused.k[k].kind; }
  static __device__ __forceinline__ uint32_t kw(int k) { return ST ? fused_wbytes((SA >> (36 + 3 * k)) & 3) : (uint32_t)c_fused.k[k].w; }
  static __device__ __forceinline__ bool combine() { return ST ? ((SA >> 41) & 1) != 0 : c_fused.combine != 0; }
  static __device__ __forceinline__ int np() { return ST ? (int)((SA >> 42) & 3) : c_fused.n_prod; }
  static __device__ __forceinline__ int pkind(int j) { return ST ? (int)((SA >> (44 + 7 * j)) & 3) : (int)c_fused.p[j].kind; }
  static __device__ __forceinline__ int pasrc(int j) { return ST ? (int)((SA >> (46 + 7 * j)) & 1) : (int)c_fused.p[j].a_src; }
  static __device__ __forceinline__ uint32_t paw(int j) { return ST ? fused_wbytes((SA >> (47 + 7 * j)) & 3) : (uint32_t)c_fused.p[j].a_w; }
  static __device__ __forceinline__ uint32_t pbw(int j) { return ST ? fused_wbytes((SA >> (49 + 7 * j)) & 3) : (uint32_t)c_fused.p[j].b_w; }
  static __device__ __forceinline__ int na() { return ST ? (int)(SB & 7) : c_fused.n_acc; }
  static __device__ __forceinline__ int asrc(int i) { return ST ? (int)((SB >> (3 + 4 * i)) & 3) : (int)c_fused.a[i].src; }
  static __device__ __forceinline__ uint32_t aw(int i) { return ST ? fused_wbytes((SB >> (5 + 4 * i)) & 3) : (uint32_t)c_fused.a[i].w; }
};

// one tile's bulk copies (issued by lane 0 of the owning warp)
__device__ __noinline__ void fused_issue(uint8_t* stage, uint64_t* bar, int64_t row0) {
  const FusedSpec& F = c_fused;
  mbar_expect_tx(bar, F.tile_tx);
  const int n = F.n_cols;
  for (int c = 0; c < n; c++) {
    const FusedCol& fc = F.cols[c];
    bulk_g2s(stage + fc.off, (const uint8_t*)fc.data + row0 * fc.width, fc.tile_bytes, bar);
  }
}

// ragged last tile / unaligned slices: the warp loads its tile itself, zero (or empty-string) fill
__device__ __noinline__ void fused_load_coop(uint8_t* stage, int64_t row0, int rows, int tile_rows, int lane) {
  const FusedSpec& F = c_fused;
  const int n = F.n_cols;
  for (int c = 0; c < n; c++) {
    const FusedCol& fc = F.cols[c];
    uint8_t* dst = stage + fc.off;
    if (fc.utf8) {
      const int32_t* src = (const int32_t*)fc.data + row0;
      int32_t* d = (int32_t*)dst;
      for (int k = lane; k <= tile_rows; k += 32) d[k] = src[k <= rows ? k : rows];
    } else {
      const uint8_t* src = (const uint8_t*)fc.data + row0 * fc.width;
      const uint32_t nb = (uint32_t)rows * fc.width, tb = (uint32_t)tile_rows * fc.width;
      if ((fc.width & 3) == 0 && (((uintptr_t)src) & 3) == 0) {
        const uint32_t* s4 = (const uint32_t*)src;
        uint32_t* d4 = (uint32_t*)dst;
        for (uint32_t k = lane; k < tb / 4; k += 32) d4[k] = (k * 4 < nb) ? s4[k] : 0u;
      } else {
        for (uint32_t k = lane; k < tb; k += 32) dst[k] = (k < nb) ? src[k] : 0;
      }
    }
  }
}

__device__ __forceinline__ bool fused_cmp(int fop, int64_t v, int64_t imm) {
  switch (fop) {
    case 0: return v == imm;
    case 1: return v != imm;
    case 2: return v < imm;
    case 3: return v <= imm;
    case 4: return v > imm;
    default: return v >= imm;
  }
}

// |v| < 2^46 test, accumulated: returns bits that are non-zero iff (lo, hi) is NOT a small value
__device__ __forceinline__ uint64_t fused_range_bits(uint64_t lo, uint64_t hi) {
  const uint64_t s = (uint64_t)((int64_t)hi >> 63);
  return (hi ^ s) | ((lo ^ s) >> 46);
}

template <int G, int R, class X>
__device__ __forceinline__ uint32_t fused_rows(const uint8_t* __restrict__ stage, const int lane, uint32_t active, RegAggState<G>& S, RegGroupTable* gt,
                                               unsigned long long (&dir)[G], uint32_t& dir_n) {
  const FusedSpec& F = c_fused;
  // ---- key images, phase 1: offsets from the tile, then the dependent chars loads issued back to
  // ---- back (aligned 8-byte words; allocations carry slack) so their latency overlaps the rest
  uint32_t klen[2][R], ksh[2][R];
  uint64_t kw0[2][R], kw1[2][R];
  uint32_t key_too_long = 0;
#pragma unroll
  for (int k = 0; k < 2; k++) {
#pragma unroll
    for (int r = 0; r < R; r++) {
      klen[k][r] = 0;
      ksh[k][r] = 0;
      kw0[k][r] = kw1[k][r] = 0;
    }
  }
  if (G > 1) {
#pragma unroll
    for (int k = 0; k < 2; k++) {
      if (k >= X::nk()) break;
      const FusedKey& fk = F.k[k];
      if (X::kkind(k) == 1) {
        const int32_t* off = (const int32_t*)(stage + fk.off);
        const uint8_t* chars = fk.chars;
        const uint32_t max_len = fk.max_len;
#pragma unroll
        for (int r = 0; r < R; r++) {
          const int e = lane + 32 * r;
          const int32_t o0 = off[e];
          const uint32_t len = (uint32_t)(off[e + 1] - o0);
          key_too_long |= (len > max_len ? 1u : 0u) << r;
          const uint8_t* p = chars + o0;
          const uint64_t* base = (const uint64_t*)((uintptr_t)p & ~(uintptr_t)7);
          const uint32_t l = len > max_len ? 0u : len;
          const uint32_t sh = (uint32_t)((uintptr_t)p & 7) * 8;
          klen[k][r] = l;
          ksh[k][r] = sh;
          kw0[k][r] = l ? base[0] : 0ull;
          kw1[k][r] = (sh + l * 8 > 64) ? base[1] : 0ull;
        }
      } else {
        const uint8_t* p = stage + fk.off;
        const uint32_t w = X::kw(k);
#pragma unroll
        for (int r = 0; r < R; r++) kw0[k][r] = (uint64_t)ld_w(p, w, lane + 32 * r);
      }
    }
  }
  // ---- filters
#pragma unroll
  for (int i = 0; i < FUSED_MAX_FILTERS; i++) {
    if (i >= X::nf()) break;
    const uint8_t* p = stage + F.f[i].off;
    const uint32_t w = X::fw(i);
    const int fop = X::fop(i);
    const int64_t imm = F.f[i].imm;
    uint32_t pass = 0;
#pragma unroll
    for (int r = 0; r < R; r++) pass |= (fused_cmp(fop, ld_w(p, w, lane + 32 * r), imm) ? 1u : 0u) << r;
    active &= pass;
  }
  // ---- products (checked decimal arithmetic, in registers)
  uint64_t p0lo[R], p0hi[R], p1lo[R], p1hi[R];
#pragma unroll
  for (int r = 0; r < R; r++) p0lo[r] = p0hi[r] = p1lo[r] = p1hi[r] = 0;
  uint32_t ovf = 0;
  if (X::np() >= 1) {
    const FusedProd& q = F.p[0];
    const uint8_t* pa = stage + q.a_off;
    const uint8_t* pb = stage + q.b_off;
    const uint32_t aw = X::paw(0), bw = X::pbw(0);
    const int kind = X::pkind(0);
    const uint64_t llo = q.lit_lo, lhi = q.lit_hi;
#pragma unroll
    for (int r = 0; r < R; r++) {
      uint64_t alo, ahi, blo, bhi;
      ld_raw128(pa, aw, lane + 32 * r, alo, ahi);
      ld_raw128(pb, bw, lane + 32 * r, blo, bhi);
      if (kind != 2) ovf |= addsub128(kind == 0, llo, lhi, blo, bhi) << r;
      const Prod128 pr = mul128_fast_val(alo, ahi, blo, bhi);
      p0lo[r] = pr.lo;
      p0hi[r] = pr.hi;
      ovf |= pr.ovf << r;
    }
  }
  if (X::np() >= 2) {
    const FusedProd& q = F.p[1];
    const uint8_t* pa = stage + q.a_off;
    const uint8_t* pb = stage + q.b_off;
    const uint32_t aw = X::paw(1), bw = X::pbw(1);
    const int kind = X::pkind(1), a_src = X::pasrc(1);
    const uint64_t llo = q.lit_lo, lhi = q.lit_hi;
#pragma unroll
    for (int r = 0; r < R; r++) {
      uint64_t alo = p0lo[r], ahi = p0hi[r], blo, bhi;
      if (!a_src) ld_raw128(pa, aw, lane + 32 * r, alo, ahi);
      ld_raw128(pb, bw, lane + 32 * r, blo, bhi);
      if (kind != 2) ovf |= addsub128(kind == 0, llo, lhi, blo, bhi) << r;
      const Prod128 pr = mul128_fast_val(alo, ahi, blo, bhi);
      p1lo[r] = pr.lo;
      p1hi[r] = pr.hi;
      ovf |= pr.ovf << r;
    }
  }
  if (ovf & active) raise(1);
  // ---- group resolution: one-hot membership oh[r][g] (0/1) against the register-cached directory
  uint32_t oh[R][G];
  if (G == 1) {
#pragma unroll
    for (int r = 0; r < R; r++) oh[r][0] = (active >> r) & 1;
  } else {
    if (key_too_long & active) atomicExch(&PROG.status->pack_overflow, 1u);
    unsigned long long kv0[R], kv1[R], ck[R];
    // key images, phase 2: finish the packing now that the chars words have arrived
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const bool packed = k < X::nk() && X::kkind(k) == 1;
      const int shift = F.k[k].shift;
#pragma unroll
      for (int r = 0; r < R; r++) {
        unsigned long long v = kw0[k][r];
        if (packed) {
          const uint32_t len = klen[k][r], sh = ksh[k][r];
          unsigned long long w = kw0[k][r] >> sh;
          if (sh) w |= kw1[k][r] << ((64 - sh) & 63);
          w &= (len >= 8) ? ~0ull : ((1ull << (len * 8)) - 1);
          v = len ? (w | ((unsigned long long)len << shift)) : 0ull;
        }
        if (k == 0) kv0[r] = v;
        else kv1[r] = v;
      }
    }
    const unsigned long long bias0 = (unsigned long long)F.k[0].bias, bias1 = (unsigned long long)F.k[1].bias;
    const bool combine = X::combine();
#pragma unroll
    for (int r = 0; r < R; r++) ck[r] = combine ? ((kv0[r] + bias0) + (kv1[r] + bias1) * 4294967296ull) : kv0[r];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const bool on = (active >> r) & 1;
      uint32_t hit = 0;
#pragma unroll
      for (int q = 0; q < G; q++) {
        const uint32_t m = (on && q < (int)dir_n && dir[q] == ck[r]) ? 1u : 0u;
        oh[r][q] = m;
        hit |= m;
      }
      if (on && !hit) {  // rare: a key this thread has not seen yet
        const int g = fused_resolve_slow(gt, G, X::nk(), ck[r], kv0[r], kv1[r]);
        uint32_t pub = 0;
#pragma unroll
        for (int q = 0; q < G; q++) {
          const bool ok = (q == (int)pub) && (*(volatile unsigned int*)&gt->state[q] == 2u);
          if (ok) {
            dir[q] = *(volatile unsigned long long*)&gt->hash[q];
            pub++;
          }
        }
        dir_n = pub;
        if (g < 0) {
          atomicExch(&PROG.status->overflow, 1u);
          active &= ~(1u << r);
        } else {
#pragma unroll
          for (int q = 0; q < G; q++) oh[r][q] = (q == g) ? 1u : 0u;
        }
      }
    }
  }
  // ---- accumulate (exact int64 partials)
#pragma unroll
  for (int a = 0; a < VM_REG_ACC; a++) {
    if (a >= X::na()) break;
    const int src = X::asrc(a);
    if (src == 3) {  // COUNT
#pragma unroll
      for (int r = 0; r < R; r++) {
#pragma unroll
        for (int g = 0; g < G; g++) S.lo[g][a] += (uint64_t)oh[r][g];
      }
      continue;
    }
    uint64_t vlo[R], vhi[R];
    const uint32_t w = (src == 0) ? X::aw(a) : 16u;
    if (src == 1) {
#pragma unroll
      for (int r = 0; r < R; r++) {
        vlo[r] = p0lo[r];
        vhi[r] = p0hi[r];
      }
    } else if (src == 2) {
#pragma unroll
      for (int r = 0; r < R; r++) {
        vlo[r] = p1lo[r];
        vhi[r] = p1hi[r];
      }
    } else {
      const uint8_t* p = stage + F.a[a].off;
#pragma unroll
      for (int r = 0; r < R; r++) ld_raw128(p, w, lane + 32 * r, vlo[r], vhi[r]);
    }
    uint64_t big = 0;
    if (w != 4) {
#pragma unroll
      for (int r = 0; r < R; r++) big |= fused_range_bits(vlo[r], vhi[r]);
    }
    if (big) {  // rare: some addend of this thread is >= 2^46 in magnitude -> straight to the global table
#pragma unroll
      for (int r = 0; r < R; r++) {
        uint32_t any = 0, gsel = 0;
#pragma unroll
        for (int g = 0; g < G; g++) {
          any |= oh[r][g];
          if (oh[r][g]) gsel = (uint32_t)g;
        }
        if (any && fused_range_bits(vlo[r], vhi[r])) {
          reg_merge_big(gt, G, (int)gsel, a, make_i128(vlo[r], vhi[r]));
          vlo[r] = 0;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
#pragma unroll
      for (int g = 0; g < G; g++) S.lo[g][a] += vlo[r] * (uint64_t)oh[r][g];
    }
  }
  return active;
}

template <int G, int R, int BT, uint64_t SA, uint64_t SB>
__global__ void __launch_bounds__(BT, 1) fused_kernel() {
  typedef FusedX<SA, SB> X;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[FUSED_MAX_WARPS * FUSED_MAX_STAGES];
  __shared__ RegGroupTable gtable;
  const FusedSpec& F = c_fused;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, B = blockDim.x, NW = B >> 5;
  constexpr int TR = 32 * R;
  const int64_t n_rows = PROG.n_rows;
  const int64_t n_tiles = (n_rows + TR - 1) / TR;
  const int S = F.n_stages;
  const uint32_t stage_bytes = F.stage_bytes;
  const bool use_tma = F.use_tma != 0;
  uint8_t* ring = smem + (size_t)warp * S * stage_bytes;
  uint64_t* bar = bars + warp * FUSED_MAX_STAGES;
  if (lane == 0) {
    for (int s = 0; s < S; s++) mbar_init(&bar[s], 1);
    mbar_fence_init();
  }
  if (tid < VM_REG_GROUPS) {
    gtable.state[tid] = 0;
    gtable.hash[tid] = 0;
    if (tid == 0) gtable.n_groups = 0;
  }
  __syncthreads();

  RegAggState<G> S_reg;
#pragma unroll
  for (int g = 0; g < G; g++) {
#pragma unroll
    for (int a = 0; a < VM_REG_ACC; a++) S_reg.lo[g][a] = 0;
  }
  unsigned long long dir[G];
  uint32_t dir_n = 0;
#pragma unroll
  for (int q = 0; q < G; q++) dir[q] = 0xFFFFFFFFFFFFFFFFull;
  uint32_t live_rows = 0;

  // warp tiles are dealt round-robin over all warps of the grid: adjacent warps read adjacent rows
  const int64_t gw = (int64_t)blockIdx.x * NW + warp, stride = (int64_t)gridDim.x * NW;
  auto tile_is_tma = [&](int64_t t) { return use_tma && (t + 1) * (int64_t)TR <= n_rows; };
  if (lane == 0) {
    for (int k = 0; k < S - 1; k++) {
      const int64_t t = gw + k * stride;
      if (t < n_tiles && tile_is_tma(t)) fused_issue(ring + (size_t)k * stage_bytes, &bar[k], t * TR);
    }
  }
  uint32_t phase_bits = 0;
  int s = 0;
  for (int64_t t = gw; t < n_tiles; t += stride, s = (s + 1 == S) ? 0 : s + 1) {
    uint8_t* stage = ring + (size_t)s * stage_bytes;
    unsigned int stop = 0;
    if (lane == 0) {
      // refill the buffer this warp released at the end of the previous iteration
      const int64_t tn = t + (int64_t)(S - 1) * stride;
      const int sn = (s == 0) ? S - 1 : s - 1;
      if (tn < n_tiles && tile_is_tma(tn)) fused_issue(ring + (size_t)sn * stage_bytes, &bar[sn], tn * TR);
      stop = *(volatile unsigned int*)&PROG.status->overflow;  // sampled early, consumed at the end of the tile
    }
    const int64_t row0 = t * TR;
    const int rows = (int)((n_rows - row0) < TR ? (n_rows - row0) : TR);
    if (tile_is_tma(t)) {
      mbar_wait(&bar[s], (phase_bits >> s) & 1);
      phase_bits ^= 1u << s;
    } else {
      fused_load_coop(stage, row0, rows, TR, lane);
      __syncwarp();
    }
    uint32_t active = 0;
#pragma unroll
    for (int r = 0; r < R; r++)
      if (lane + 32 * r < rows) active |= 1u << r;
    active = fused_rows<G, R, X>(stage, lane, active, S_reg, &gtable, dir, dir_n);
    live_rows += __popc(active);
    // warp-uniform exit test; also the point after which lane 0 may overwrite this stage
    if (__any_sync(0xFFFFFFFFu, stop != 0)) {
      for (int j = 1; j < S; j++) {  // drain bulk copies still in flight before the CTA may exit
        const int64_t tt = t + j * stride;
        const int sj = (s + j) % S;
        if (tt < n_tiles && tile_is_tma(tt)) mbar_wait(&bar[sj], (phase_bits >> sj) & 1);
      }
      break;
    }
  }
  live_rows = __reduce_add_sync(0xFFFFFFFFu, live_rows);
  if (lane == 0 && live_rows) atomicAdd(&PROG.status->in_active, (unsigned long long)live_rows);
  __syncthreads();
  // scalar aggregates emit their single group even when no CTA saw a row: CTA 0 always flushes
  const bool has_rows = (int64_t)blockIdx.x * NW < n_tiles;
  if (has_rows || (G == 1 && blockIdx.x == 0)) reg_agg_flush<G, true>(S_reg, nullptr, &gtable, (Acc128*)smem, tid, B);
}

template <int G, int R, int BT, uint64_t SA, uint64_t SB>
static cudaError_t launch_fused_one(int grid, int block, size_t smem, cudaStream_t st) {
  if (block > BT) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(fused_kernel<G, R, BT, SA, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  fused_kernel<G, R, BT, SA, SB><<<grid, block, smem, st>>>();
  return cudaGetLastError();
}

// ---- shapes compiled ahead of time -----------------------------------------------------------------
// TPC-H q1 (benchmarks/queries/q1.sql): 1 date filter, 2 packed single-character keys,
// disc_price = price*(1-disc), charge = disc_price*(1+tax); count + 4 decimal sums + sum(disc)
constexpr FusedShapeDesc kShapeQ1 = {1, {4}, {3 /*LE*/}, 2, {1, 1}, {0, 0}, 1, 2, {0, 1}, {0, 1}, {16, 0}, {16, 16}, 6,
                                     {3, 0, 0, 1, 2, 0}, {0, 16, 16, 0, 0, 16}};
// TPC-H q6 (benchmarks/queries/q6.sql): date range, discount BETWEEN, quantity <; count + sum(price*disc)
constexpr FusedShapeDesc kShapeQ6 = {5, {4, 4, 16, 16, 16}, {5 /*GE*/, 2 /*LT*/, 5, 3 /*LE*/, 2}, 0, {0, 0}, {0, 0}, 0, 1, {2, 0}, {0, 0}, {16, 0}, {16, 0}, 2,
                                     {3, 1}, {0, 0}};
constexpr FusedShape kQ1 = fused_shape_encode(kShapeQ1);
constexpr FusedShape kQ6 = fused_shape_encode(kShapeQ6);

// (rows per thread, launch bound) variants compiled for each shape
template <int G, uint64_t SA, uint64_t SB>
static cudaError_t launch_fused_variant(int R, int grid, int block, size_t smem, cudaStream_t st) {
  if (R == 2) {
    if (block <= 256) return launch_fused_one<G, 2, 256, SA, SB>(grid, block, smem, st);
    if (block <= 384) return launch_fused_one<G, 2, 384, SA, SB>(grid, block, smem, st);
    return launch_fused_one<G, 2, 512, SA, SB>(grid, block, smem, st);
  }
  if (R == 4) {
    if (block <= 256) return launch_fused_one<G, 4, 256, SA, SB>(grid, block, smem, st);
    return launch_fused_one<G, 4, 384, SA, SB>(grid, block, smem, st);
  }
  return cudaErrorInvalidValue;
}

// *is_static tells whether an ahead-of-time shape ran (else the run-time-described variant)
static cudaError_t launch_fused(const FusedSpec& F, FusedShape shape, int reg_groups, int grid, int block, size_t smem, cudaStream_t st, int* is_static) {
  const int R = F.rows_per_thread;
  *is_static = 1;
  if (reg_groups > 1 && shape.a == kQ1.a && shape.b == kQ1.b) return launch_fused_variant<VM_REG_GROUPS, kQ1.a, kQ1.b>(R, grid, block, smem, st);
  if (reg_groups <= 1 && shape.a == kQ6.a && shape.b == kQ6.b) return launch_fused_variant<1, kQ6.a, kQ6.b>(R, grid, block, smem, st);
  *is_static = 0;
  if (reg_groups > 1) return launch_fused_variant<VM_REG_GROUPS, 0, 0>(R, grid, block, smem, st);
  return launch_fused_variant<1, 0, 0>(R, grid, block, smem, st);
}

}  // namespace b200
