// Device helpers and types shared by the Golomb encoders (coding.cu: tile walks with global prefix arrays; coding2.cu: wide
// tiles, fused scans, register-assembled codewords).
#pragma once
#include "bic_internal.cuh"

#define TILE_THREADS 256
#define TILE_WORDS_PER_THREAD 4
#define TILE_WORDS (TILE_THREADS * TILE_WORDS_PER_THREAD)
#define GOL_SMEM_WORDS 2560  // staged output range of one tile: 80 Kbit for 32 Kbit of input

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// k after t samples whose sum is acc (uint32 arithmetic, GolombCoder.cpp:33; the search is
// capped at 31 because k >= 32 trips the reference's own assert, GolombCoder.cpp:14)
__device__ __forceinline__ uint32_t golomb_k(uint64_t t64, uint64_t bits_consumed) {
  if (t64 == 0) return 1;  // Golomb.h:18
  const uint32_t t = (uint32_t)t64;
  const uint32_t acc = (uint32_t)(bits_consumed - t64);  // sum of the first t samples, mod 2^32
  if (acc <= t) return 0;
  if (acc < 0x80000000u && t != 0) {
    // no shift below the answer can wrap: t << k has the bit length of acc
    int k = __clz(t) - __clz(acc);
    if ((t << k) < acc) k++;
    return (uint32_t)k;
  }
  uint32_t k = 0;
  while (k < 31 && (uint32_t)(t << k) < acc) k++;
  return k;
}

// Is k the same for every sample of a 128-bit stretch? Sample ranks run over [t, t + 127] and the sums
// over [acc, acc + 128] (acc = consumed - t never decreases and grows by the zeros passed). If
// (t << k) >= acc + 128 and ((t + 127) << (k - 1)) < acc, with no 32-bit wrap anywhere, every sample of the
// stretch takes the k of the first: the coder's adaptation is far slower than one thread's four words once
// a few thousand samples are in, so the per-sample search runs once per thread instead of once per sample.
__device__ __forceinline__ bool golomb_k_stable(uint64_t t, uint64_t consumed, uint32_t* kout) {
  if (t == 0 || t + 128 >= (1ull << 31)) return false;
  const uint64_t acc = consumed - t;
  if (acc + 128 >= (1ull << 31)) return false;
  const uint32_t k = golomb_k(t, consumed);
  if (((t + 127) << k) >= (1ull << 32)) return false;
  if ((t << k) < acc + 128) return false;
  if (k > 0 && ((t + 127) << (k - 1)) >= acc) return false;
  *kout = k;
  return true;
}

__device__ __forceinline__ unsigned long long block_excl_scan_u64(unsigned long long v, unsigned long long* total,
                                                                  unsigned long long* s_warp /* 8 */) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  unsigned long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  __syncthreads();
  if (lane == 31) s_warp[wib] = inc;
  __syncthreads();
  unsigned long long base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < TILE_THREADS / 32; ++w) {
    const unsigned long long x = s_warp[w];
    if (w < wib) base += x;
    tot += x;
  }
  if (total) *total = tot;
  return base + inc - v;
}

// exclusive prefix max of "position of my last one" (-1 = none)
__device__ __forceinline__ long long block_excl_scan_max(long long v, long long* total, long long* s_warp /* 8 */) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc = inc > y ? inc : y;
  }
  long long excl = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) excl = -1;
  __syncthreads();
  if (lane == 31) s_warp[wib] = inc;
  __syncthreads();
  long long base = -1, tot = -1;
#pragma unroll
  for (int w = 0; w < TILE_THREADS / 32; ++w) {
    const long long x = s_warp[w];
    if (w < wib) base = base > x ? base : x;
    tot = tot > x ? tot : x;
  }
  if (total) *total = tot;
  return base > excl ? base : excl;
}

__device__ __forceinline__ void put_bits(uint32_t* __restrict__ out, unsigned long long o, uint32_t value, uint32_t nb) {
  if (nb == 0 || value == 0) return;
  const unsigned long long wi = o >> 5;
  const unsigned off = (unsigned)(o & 31);
  const unsigned long long v64 = (unsigned long long)value << (64 - off - nb);
  const uint32_t hi = (uint32_t)(v64 >> 32), lo = (uint32_t)v64;
  if (hi) atomicOr(out + wi, bswap32(hi));
  if (lo) atomicOr(out + wi + 1, bswap32(lo));
}
__device__ __forceinline__ void put_one(uint32_t* __restrict__ out, unsigned long long o) {
  atomicOr(out + (o >> 5), bswap32(0x80000000u >> (unsigned)(o & 31)));
}

// MODE 0: code lengths per tile; MODE 1: scatter codewords (+ chunk index)
// Where a (shard of a) matrix sits in the global sample stream. Single GPU: all zero / -1. Row-sharded
// coding (bic_dist_golomb_encode): the shard's codewords are an exact substring of the one global stream.
struct GolBase {
  unsigned long long t0;      // ones before this shard (global rank of its first sample)
  long long pos0;             // global stream position of the shard's first bit
  long long prev0;            // global position of the last one before the shard, -1 if none
  unsigned long long out0;    // bit offset of the shard's first codeword inside its own output buffer (code0 & 31)
  unsigned long long code0;   // global code-bit offset of the shard's first codeword
  unsigned long long chunk0;  // first chunk-index entry this shard stores
  int closing;                // this shard writes the run closed by the virtual one
  unsigned long long close_t, close_consumed, close_off;  // its sample rank, bits consumed before it, local bit offset
  unsigned long long close_n;                             // bits of the whole (global) matrix: where the virtual one sits (coding2.cu)
};

