// Shuffle-side CUDA kernels (sm_100a): one-pass stable radix partition (ShuffleWriterExec /
// SortShuffleWriterExec), the exchange message packer, and the small-result export packer.
//
// Reference behaviour being reproduced:
//   * BatchPartitioner::partition / compute_partition_indices -- rows go to partition hash % P and keep
//     their input order inside a partition (ballista/core/src/execution_plans/sort_shuffle/writer.rs:729-749,
//     shuffle_writer.rs:291-343); here: per-tile histogram -> exclusive scan in partition-major order ->
//     per-tile stable ranks (warp match + per-warp running counters) -> every column scattered in the same
//     kernel.  HBM bound: 2 * N * w_row algorithmic bytes (+ 4 B/row of partition ids read twice).
//   * the per-partition IPC writers (shuffle_writer.rs:317-328): replaced by packing the column slices of
//     every (destination, partition) into one contiguous message (exchange_pack_kernel).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../common/hash.hpp"
#include "kernels.h"
#include "keyimg.cuh"

namespace b200 {

static const int PT_BLOCK = 256;                      // 8 warps
static const int PT_WARPS = PT_BLOCK / 32;
static const int PT_ROUNDS = 8;                       // rows per thread
static const int PT_TILE = PT_BLOCK * PT_ROUNDS;      // 2048 rows per tile; warp w owns rows [w*256, (w+1)*256)

__device__ __forceinline__ uint32_t pid_of(const PidSrc& ps, int64_t i, uint32_t P) {
  if (ps.pid) return ps.pid[i];
  if (ps.n_keys == 0) return 0u;
  uint64_t h = jkey_valid(ps.keys[0], i) ? hash_i64((int64_t)jkey_image(ps.keys[0], i)) : 0ull;
  for (int k = 1; k < ps.n_keys; k++)
    if (jkey_valid(ps.keys[k], i)) h = combine_hashes(hash_i64((int64_t)jkey_image(ps.keys[k], i)), h);
  return (uint32_t)(h % (uint64_t)P);
}

__device__ __forceinline__ uint32_t str_len_of(const PartStrCol& c, int64_t i) {
  if (c.valid && !c.valid[i]) return 0u;
  if (c.is_view) return (uint32_t)((const unsigned long long*)c.data)[2 * i + 1];
  const int32_t* o = (const int32_t*)c.data;
  return (uint32_t)(o[i + 1] - o[i]);
}

// tile_hist[p * n_tiles + tile] = rows of partition p in the tile; counts[p] += the same;
// str_bytes[c * P + p] += string bytes of column c that go to partition p (ShuffleWritePartition.num_bytes)
__global__ void __launch_bounds__(PT_BLOCK) part_tile_hist_kernel(const PidSrc pid, int64_t n, uint32_t P, uint32_t n_tiles, uint32_t* __restrict__ tile_hist,
                                                                 unsigned long long* __restrict__ counts, PartStrCols sc, unsigned long long* __restrict__ str_bytes) {
  extern __shared__ unsigned int sh[];  // [P] counts, then [n_str][P] byte sums (as 2 x u32: lo/hi not needed: < 2^32 per tile)
  const int n_str = sc.n;
  for (uint32_t b = threadIdx.x; b < P * (1 + n_str); b += PT_BLOCK) sh[b] = 0;
  __syncthreads();
  const int64_t t0 = (int64_t)blockIdx.x * PT_TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < PT_ROUNDS; r++) {
    const int64_t i = t0 + warp * (PT_TILE / PT_WARPS) + r * 32 + lane;
    if (i < n) {
      const uint32_t p = pid_of(pid, i, P);
      atomicAdd(&sh[p], 1u);
      for (int c = 0; c < n_str; c++) {
        const uint32_t len = str_len_of(sc.c[c], i);
        if (len) atomicAdd(&sh[P * (1 + c) + p], len);
      }
    }
  }
  __syncthreads();
  for (uint32_t p = threadIdx.x; p < P; p += PT_BLOCK) {
    const unsigned int v = sh[p];
    if (tile_hist) tile_hist[(size_t)p * n_tiles + blockIdx.x] = v;
    if (v) atomicAdd(&counts[p], (unsigned long long)v);
    for (int c = 0; c < n_str; c++) {
      const unsigned int bts = sh[P * (1 + c) + p];
      if (bts) atomicAdd(&str_bytes[(size_t)c * P + p], (unsigned long long)bts);
    }
  }
}

template <typename T>
__device__ __forceinline__ void scatter_rows(const GatherCol& c, const int64_t (&row)[PT_ROUNDS], const uint32_t (&dst)[PT_ROUNDS], const uint32_t (&p)[PT_ROUNDS],
                                             int64_t n) {
  const T* __restrict__ in = (const T*)c.in;
  T* __restrict__ out = (T*)c.out;
  T v[PT_ROUNDS];
#pragma unroll
  for (int r = 0; r < PT_ROUNDS; r++)
    if (row[r] < n) v[r] = in[row[r]];
  if (c.part_base) {
    // per-partition destination bases (possibly peer memory): generic stores, NVLink carries the remote ones
#pragma unroll
    for (int r = 0; r < PT_ROUNDS; r++)
      if (row[r] < n) ((T*)c.part_base[p[r]])[dst[r]] = v[r];
    return;
  }
#pragma unroll
  for (int r = 0; r < PT_ROUNDS; r++)
    if (row[r] < n) out[dst[r]] = v[r];
}

// offsets[p * n_tiles + tile] = first output row of (partition p, tile); the rank of a row inside its
// (p, tile) group is its stable position: rows of earlier warps, earlier rounds, lower lanes first.
__global__ void __launch_bounds__(PT_BLOCK) part_tile_scatter_kernel(const PidSrc pid, int64_t n, uint32_t P, uint32_t n_tiles,
                                                                    const uint64_t* __restrict__ offsets, GatherCols cols, uint32_t* __restrict__ dest_out) {
  extern __shared__ unsigned int sh[];  // [PT_WARPS][P] per-warp counts -> per-warp bases
  for (uint32_t b = threadIdx.x; b < P * PT_WARPS; b += PT_BLOCK) sh[b] = 0;
  __syncthreads();
  const int64_t t0 = (int64_t)blockIdx.x * PT_TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned int* mine = sh + (size_t)warp * P;
  uint32_t p[PT_ROUNDS];
  int64_t row[PT_ROUNDS];
  uint32_t rank[PT_ROUNDS];
  const uint32_t lt = (1u << lane) - 1u;
  // pass 1: rank inside the warp's 256 rows (round-major, lane-minor == input order) and per-warp counts
#pragma unroll
  for (int r = 0; r < PT_ROUNDS; r++) {
    row[r] = t0 + warp * (PT_TILE / PT_WARPS) + r * 32 + lane;
    const bool live = row[r] < n;
    p[r] = live ? pid_of(pid, row[r], P) : 0xFFFFFFFFu;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, p[r]);
    uint32_t base = 0;
    if (live) base = mine[p[r]];
    __syncwarp();
    rank[r] = base + __popc(peers & lt);
    if (live && (peers & lt) == 0) mine[p[r]] = base + __popc(peers);  // the lowest lane of each group advances the counter
    __syncwarp();
  }
  __syncthreads();
  // per partition: exclusive scan of the warp counts, seeded with the (partition, tile) output offset
  for (uint32_t q = threadIdx.x; q < P; q += PT_BLOCK) {
    uint64_t run = offsets[(size_t)q * n_tiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < PT_WARPS; w++) {
      const unsigned int c = sh[(size_t)w * P + q];
      sh[(size_t)w * P + q] = (unsigned int)run;  // < 2^32 rows per task (checked by the host)
      run += c;
    }
  }
  __syncthreads();
  uint32_t dst[PT_ROUNDS];
#pragma unroll
  for (int r = 0; r < PT_ROUNDS; r++) dst[r] = row[r] < n ? mine[p[r]] + rank[r] : 0u;
  if (dest_out) {
#pragma unroll
    for (int r = 0; r < PT_ROUNDS; r++)
      if (row[r] < n) dest_out[row[r]] = dst[r];
  }
  for (int c = 0; c < cols.n; c++) {
    const GatherCol& gc = cols.c[c];
    switch (gc.width) {
      case 1: scatter_rows<uint8_t>(gc, row, dst, p, n); break;
      case 2: scatter_rows<uint16_t>(gc, row, dst, p, n); break;
      case 4: scatter_rows<uint32_t>(gc, row, dst, p, n); break;
      case 8: scatter_rows<uint64_t>(gc, row, dst, p, n); break;
      default: scatter_rows<ulonglong2>(gc, row, dst, p, n); break;
    }
  }
}

size_t partition_scatter_smem(uint32_t P) { return (size_t)P * PT_WARPS * sizeof(unsigned int); }
uint32_t partition_n_tiles(int64_t n) { return (uint32_t)((n + PT_TILE - 1) / PT_TILE); }

cudaError_t launch_partition_hist(const PidSrc& pid, int64_t n, uint32_t P, uint32_t* tile_hist, unsigned long long* counts, const PartStrCols& sc,
                                  unsigned long long* str_bytes, cudaStream_t st) {
  const uint32_t nt = partition_n_tiles(n);
  if (nt == 0) return cudaSuccess;
  const size_t sm = (size_t)P * (1 + sc.n) * sizeof(unsigned int);
  if (sm > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(part_tile_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
  }
  part_tile_hist_kernel<<<nt, PT_BLOCK, sm, st>>>(pid, n, P, nt, tile_hist, counts, sc, str_bytes);
  return cudaGetLastError();
}

cudaError_t launch_partition_scatter(const PidSrc& pid, int64_t n, uint32_t P, const uint64_t* offsets, const GatherCols& cols, uint32_t* dest_out,
                                     cudaStream_t st) {
  const uint32_t nt = partition_n_tiles(n);
  if (nt == 0) return cudaSuccess;
  const size_t sm = partition_scatter_smem(P);
  if (sm > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(part_tile_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
  }
  part_tile_scatter_kernel<<<nt, PT_BLOCK, sm, st>>>(pid, n, P, nt, offsets, cols, dest_out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Pack jobs: copy / convert column slices into a contiguous destination (exchange messages, small exports)
// ------------------------------------------------------------------------------------------------
// One CTA per job.  PK_COPY: plain bytes.  PK_BITMAP: validity bytes -> bitmap + null count.  PK_STR_VIEWS / PK_STR_UTF8: write `rows + 1` int32 offsets starting
// at 0 to dst and the character bytes to dst2 (block-wide running prefix sum over chunks of 256 rows).
__global__ void __launch_bounds__(256) pack_jobs_kernel(const PackJob* __restrict__ jobs, int n_jobs) {
  const int j = blockIdx.x;
  if (j >= n_jobs) return;
  const PackJob J = jobs[j];
  const int tid = threadIdx.x;
  if (J.kind == PK_COPY) {
    const uint8_t* s = (const uint8_t*)J.src;
    uint8_t* d = (uint8_t*)J.dst;
    if ((((uintptr_t)s | (uintptr_t)d) & 15) == 0) {
      const uint64_t n16 = J.bytes / 16;
      for (uint64_t k = tid; k < n16; k += 256) ((uint4*)d)[k] = ((const uint4*)s)[k];
      for (uint64_t k = n16 * 16 + tid; k < J.bytes; k += 256) d[k] = s[k];
    } else {
      for (uint64_t k = tid; k < J.bytes; k += 256) d[k] = s[k];
    }
    return;
  }
  if (J.kind == PK_UTF8_VIEWS) {
    // Arrow offsets + characters -> 16-byte views {pointer, length} (the layout intermediate batches carry)
    const int32_t* o = (const int32_t*)J.src;
    unsigned long long* v = (unsigned long long*)J.dst;
    for (int64_t i = tid; i < J.rows; i += 256) {
      const int32_t o0 = o[i], o1 = o[i + 1];
      v[2 * i] = (unsigned long long)(J.chars + o0);
      v[2 * i + 1] = (unsigned long long)(uint32_t)(o1 - o0);
    }
    return;
  }
  if (J.kind == PK_BITMAP) {
    // byte-per-value -> Arrow bitmap (+ number of zero values at dst2, if asked for)
    const uint8_t* s = (const uint8_t*)J.src;
    uint8_t* d = (uint8_t*)J.dst;
    const int64_t nbytes = (J.rows + 7) / 8;
    unsigned int zeros = 0;
    for (int64_t b = tid; b < nbytes; b += 256) {
      uint8_t v = 0;
      for (int k = 0; k < 8; k++) {
        const int64_t i = b * 8 + k;
        if (i < J.rows) {
          if (s[i]) v |= (uint8_t)(1u << k);
          else zeros++;
        }
      }
      d[b] = v;
    }
    if (J.dst2) {
      __shared__ unsigned int zsum;
      if (tid == 0) zsum = 0;
      __syncthreads();
      if (zeros) atomicAdd(&zsum, zeros);
      __syncthreads();
      if (tid == 0) *(unsigned long long*)J.dst2 = zsum;
    }
    return;
  }
  __shared__ uint32_t warp_sum[8];
  __shared__ uint32_t carry_sh;
  if (tid == 0) carry_sh = 0;
  __syncthreads();
  int32_t* off_out = (int32_t*)J.dst;
  uint8_t* ch_out = (uint8_t*)J.dst2;
  const int lane = tid & 31, warp = tid >> 5;
  for (int64_t r0 = 0; r0 < J.rows; r0 += 256) {
    const int64_t i = r0 + tid;
    uint32_t len = 0;
    const uint8_t* sp = nullptr;
    if (i < J.rows && !(J.valid && !J.valid[i])) {
      if (J.kind == PK_STR_VIEWS) {
        sp = (const uint8_t*)((const unsigned long long*)J.src)[2 * i];
        len = (uint32_t)((const unsigned long long*)J.src)[2 * i + 1];
      } else {
        const int32_t* o = (const int32_t*)J.src;
        sp = J.chars + o[i];
        len = (uint32_t)(o[i + 1] - o[i]);
      }
    }
    uint32_t inc = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    uint32_t base = carry_sh;
    for (int w = 0; w < warp; w++) base += warp_sum[w];
    const uint32_t excl = base + inc - len;
    if (i < J.rows) {
      off_out[i] = (int32_t)excl;
      if ((uint64_t)excl + len <= J.bytes)  // J.bytes = capacity of the character area (the host re-checks the total)
        for (uint32_t k = 0; k < len; k++) ch_out[excl + k] = sp[k];
    }
    __syncthreads();
    if (tid == 255) carry_sh = base + inc;
    __syncthreads();
  }
  if (tid == 0) off_out[J.rows] = (int32_t)carry_sh;
}

void launch_pack_jobs(const PackJob* jobs_dev, int n_jobs, cudaStream_t st) {
  if (n_jobs <= 0) return;
  pack_jobs_kernel<<<n_jobs, 256, 0, st>>>(jobs_dev, n_jobs);
}

}  // namespace b200
