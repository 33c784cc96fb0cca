// Parallel Golomb / EG coding of a bit matrix and the chunk-parallel decoders.
// Reference semantics: Golomb state src/Golomb.h:12-29; GolombCoder::codeSample
// src/GolombCoder.cpp:13-34 (bit layout from its commented writer calls :22-25 and from
// GolombDecoder::binaryDecode src/GolombDecoder.cpp:15-23); EGCoder::codeRun src/eg.cpp:20-37.
//
// Golomb. The samples are the zero-run lengths of the matrix read row-major; a virtual one after
// the last bit closes the last run, so there are popcount+1 samples. With pos_t the stream
// position of the t-th one, sample x_t = pos_t - pos_{t-1} - 1 and the coder state before sample
// t is a closed form of (t, pos_{t-1}): samples = t, accumulatedError = pos_{t-1} + 1 - t
// (mod 2^32), k_t = min{k : (t << k) >= accumulatedError} (k_0 = 1). So the serial adaptive coder
// becomes: rank of every one (prefix sum of word popcounts) + position of the previous one
// (prefix max) -> k_t and codeword length per one -> prefix sum of lengths = bit offsets ->
// scatter of the k remainder bits and the closing one of each codeword into a zeroed buffer.
#include "bic_internal.cuh"

#include "gol_common.cuh"

// ------------------------------------------------------------------ dense row-major bit stream
// When cols is not a multiple of 32 the rows carry pad bits; the coders work on a compacted copy.
__global__ void k_compact_rows(const uint32_t* __restrict__ M, uint64_t cols, uint64_t wpr, uint64_t N,
                               uint32_t* __restrict__ S, uint64_t T) {
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < T; t += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t s = t * 32;
    const uint64_t send = (s + 32 < N) ? s + 32 : N;
    uint32_t out = 0;
    unsigned filled = 0;
    while (s < send) {
      const uint64_t r = s / cols, cc = s - r * cols;
      uint64_t len = cols - cc;
      if (len > send - s) len = send - s;
      const uint64_t wi = cc >> 5;
      const unsigned off = (unsigned)(cc & 31);
      const uint32_t* row = M + r * wpr;
      const uint32_t hi = row[wi];
      const uint32_t lo = (off + len > 32 && wi + 1 < wpr) ? row[wi + 1] : 0u;
      const uint32_t v = __funnelshift_l(lo, hi, off) >> (32 - (unsigned)len);
      out |= v << (32 - filled - (unsigned)len);
      filled += (unsigned)len;
      s += len;
    }
    S[t] = out;
  }
}

__global__ void k_expand_rows(const uint32_t* __restrict__ S, uint64_t cols, uint64_t wpr, uint64_t rows,
                              uint32_t* __restrict__ M) {
  const uint64_t total = rows * wpr;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = i / wpr, w = i - r * wpr;
    const uint64_t c0 = w * 32;
    const unsigned len = (unsigned)((cols - c0 < 32) ? cols - c0 : 32);
    const uint64_t s = r * cols + c0;
    const uint64_t wi = s >> 5;
    const unsigned off = (unsigned)(s & 31);
    const uint32_t hi = S[wi];
    const uint32_t lo = (off + len > 32) ? S[wi + 1] : 0u;
    M[i] = (__funnelshift_l(lo, hi, off) >> (32 - len)) << (32 - len);
  }
}

// ------------------------------------------------------------------ Golomb encoder
struct GolTile {           // per 1024-word tile
  uint32_t* ones;          // ones in the tile
  long long* last;         // stream position of the tile's last one, -1 if none
  unsigned long long* ones_before;  // exclusive prefix of ones
  long long* last_before;  // last one before the tile, -1 if none
  unsigned long long* bits;         // code bits produced by the tile's ones
  unsigned long long* bits_before;  // exclusive prefix
  unsigned long long* tbits;        // per thread (4 words of input): code bits of its ones, written by the length pass
};

__device__ __forceinline__ void load_tile_words(const uint32_t* __restrict__ S, uint64_t T, uint64_t w0, uint32_t (&v)[4]) {
  if (w0 + 4 <= T) {
    const uint4 q = *reinterpret_cast<const uint4*>(S + w0);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (w0 + i < T) ? S[w0 + i] : 0u;
  }
}

__global__ void __launch_bounds__(TILE_THREADS) k_gol_tile_counts(const uint32_t* __restrict__ S, uint64_t T, GolTile g) {
  __shared__ unsigned long long s_a[8];
  __shared__ long long s_b[8];
  const uint64_t w0 = (uint64_t)blockIdx.x * TILE_WORDS + threadIdx.x * TILE_WORDS_PER_THREAD;
  uint32_t v[4];
  load_tile_words(S, T, w0, v);
  unsigned long long c = 0;
  long long last = -1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    c += __popc(v[i]);
    if (v[i]) last = (long long)((w0 + i) * 32 + (32 - __ffs(v[i])));
  }
  unsigned long long tot;
  long long tlast;
  block_excl_scan_u64(c, &tot, s_a);
  block_excl_scan_max(last, &tlast, s_b);
  if (threadIdx.x == 0) { g.ones[blockIdx.x] = (uint32_t)tot; g.last[blockIdx.x] = tlast; }
}

// single CTA of 1024 threads: exclusive scans over the tiles. Each thread owns a contiguous run of tiles
// (serial over its run, one block-wide scan over the 1024 run totals), so 65536 tiles (2^31 input bits)
// cost one pass instead of 256 dependent block scans.
#define SCAN_THREADS 1024
__device__ __forceinline__ void scan1024_sum_max(unsigned long long& sum, long long& mx, unsigned long long* s_sum, long long* s_max,
                                                 unsigned long long* tot_sum, long long* tot_max) {
  // exclusive prefix over threads of (sum, max); totals returned to everyone
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  unsigned long long inc = sum;
  long long incm = mx;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
    const long long ym = __shfl_up_sync(0xffffffffu, incm, o);
    if (lane >= o) { inc += y; incm = incm > ym ? incm : ym; }
  }
  if (lane == 31) { s_sum[wib] = inc; s_max[wib] = incm; }
  __syncthreads();
  if (wib == 0) {
    unsigned long long v = s_sum[lane];
    long long vm = s_max[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, v, o);
      const long long ym = __shfl_up_sync(0xffffffffu, vm, o);
      if (lane >= o) { v += y; vm = vm > ym ? vm : ym; }
    }
    s_sum[lane] = v;   // inclusive over warps
    s_max[lane] = vm;
  }
  __syncthreads();
  const unsigned long long wbase = wib ? s_sum[wib - 1] : 0ull;
  const long long wbasem = wib ? s_max[wib - 1] : -1;
  unsigned long long ex = __shfl_up_sync(0xffffffffu, inc, 1);
  long long exm = __shfl_up_sync(0xffffffffu, incm, 1);
  if (lane == 0) { ex = 0; exm = -1; }
  *tot_sum = s_sum[31];
  *tot_max = s_max[31];
  sum = wbase + ex;
  mx = wbasem > exm ? wbasem : exm;
  __syncthreads();
}

__global__ void __launch_bounds__(SCAN_THREADS) k_gol_scan_tiles_a(GolTile g, uint64_t ntiles) {
  __shared__ unsigned long long s_sum[32];
  __shared__ long long s_max[32];
  const uint64_t per = div_up_u64(ntiles, SCAN_THREADS);
  const uint64_t t0 = threadIdx.x * per, t1 = (t0 + per < ntiles) ? t0 + per : ntiles;
  unsigned long long sum = 0;
  long long mx = -1;
  for (uint64_t i = t0; i < t1; ++i) { sum += g.ones[i]; const long long l = g.last[i]; mx = mx > l ? mx : l; }
  unsigned long long tot;
  long long totm;
  scan1024_sum_max(sum, mx, s_sum, s_max, &tot, &totm);  // now exclusive prefixes of this thread's run
  for (uint64_t i = t0; i < t1; ++i) {
    g.ones_before[i] = sum;
    g.last_before[i] = mx;
    sum += g.ones[i];
    const long long l = g.last[i];
    mx = mx > l ? mx : l;
  }
}

// scalars: [0] bitcount, [1] nsamples, [2] offset of the closing sample, [3] last one + 1
// cap_bits != 0: scalars[4] = 1 if the code does not fit a buffer of cap_bits (the asynchronous encoder sizes its buffer
// before it knows the bit count; the scatter pass then does nothing and the caller re-encodes with an exact allocation)
__global__ void __launch_bounds__(SCAN_THREADS) k_gol_scan_tiles_b(GolTile g, uint64_t ntiles, uint64_t N,
                                                                   unsigned long long* scalars, unsigned long long cap_bits) {
  __shared__ unsigned long long s_sum[32];
  __shared__ long long s_max[32];
  const uint64_t per = div_up_u64(ntiles, SCAN_THREADS);
  const uint64_t t0 = threadIdx.x * per, t1 = (t0 + per < ntiles) ? t0 + per : ntiles;
  unsigned long long sum = 0;
  long long dummy = -1;
  for (uint64_t i = t0; i < t1; ++i) sum += g.bits[i];
  unsigned long long carry;
  long long totm;
  scan1024_sum_max(sum, dummy, s_sum, s_max, &carry, &totm);
  for (uint64_t i = t0; i < t1; ++i) { g.bits_before[i] = sum; sum += g.bits[i]; }
  if (threadIdx.x == 0) {
    unsigned long long ones = 0;
    long long last = -1;
    if (ntiles) {
      ones = g.ones_before[ntiles - 1] + g.ones[ntiles - 1];
      last = g.last_before[ntiles - 1] > g.last[ntiles - 1] ? g.last_before[ntiles - 1] : g.last[ntiles - 1];
    }
    // the run closed by the virtual one at position N
    const unsigned long long consumed = (unsigned long long)(last + 1);
    const unsigned long long x = N - consumed;
    const uint32_t k = golomb_k(ones, consumed);
    scalars[0] = carry + k + (x >> k) + 1;
    scalars[1] = ones + 1;
    scalars[2] = carry;
    scalars[3] = consumed;
    scalars[4] = (cap_bits && scalars[0] + 64 > cap_bits) ? 1ull : 0ull;
  }
}

// zero the words the code will occupy (the scatter ORs into a zeroed buffer); the bit count is only known on the device
__global__ void k_gol_zero_code(uint32_t* __restrict__ out, const unsigned long long* __restrict__ scalars) {
  if (scalars[4]) return;
  const unsigned long long nw = ((scalars[0] + 31) >> 5) + 4;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < nw; i += (unsigned long long)gridDim.x * blockDim.x)
    out[i] = 0u;
}

template <int MODE>
__global__ void __launch_bounds__(TILE_THREADS) k_gol_walk(const uint32_t* __restrict__ S, uint64_t T, uint64_t N, GolTile g,
                                                           uint32_t* __restrict__ out, unsigned long long* __restrict__ index,
                                                           uint32_t chunk, GolBase gb, const unsigned long long* __restrict__ dyn) {
  // dyn (MODE 1, asynchronous encoder): k_gol_scan_tiles_b's scalars, still on the device -- the closing sample's rank,
  // offset and position come from there instead of gb, and a code that does not fit the buffer is not written at all
  if (MODE == 1 && dyn && dyn[4]) return;
  __shared__ unsigned long long s_a[8];
  __shared__ long long s_b[8];
  const uint64_t w0 = (uint64_t)blockIdx.x * TILE_WORDS + threadIdx.x * TILE_WORDS_PER_THREAD;
  uint32_t v[4];
  load_tile_words(S, T, w0, v);
  unsigned long long c = 0;
  long long last = -1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    c += __popc(v[i]);
    if (v[i]) last = (long long)((w0 + i) * 32 + (32 - __ffs(v[i])));
  }
  const unsigned long long rank0 = gb.t0 + g.ones_before[blockIdx.x] + block_excl_scan_u64(c, nullptr, s_a);
  const long long pl = block_excl_scan_max(last, nullptr, s_b);
  const long long lb = g.last_before[blockIdx.x];
  const long long pv_local = pl > lb ? pl : lb;                       // previous one inside this shard, -1 if none
  long long prev = pv_local >= 0 ? pv_local + gb.pos0 : gb.prev0;     // GLOBAL position of the previous one, -1 if none

  // pass 1: my code bits -- computed by the length pass (MODE 0) and kept per thread; the scatter pass
  // (MODE 1) reads them back instead of walking the samples twice
  unsigned long long mybits = 0;
  const uint64_t tslot = (uint64_t)blockIdx.x * TILE_THREADS + threadIdx.x;
  const long long tb = (long long)(w0 * 32) + gb.pos0;  // global position of this thread's first bit
  if (MODE == 0) {
    unsigned long long t = rank0;
    long long pv = prev;
    bool first = true, fast = false;
    uint32_t kc = 0, lpv = 0, fastbits = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t b = v[i];
      while (b) {
        const int p = __clz(b);
        b &= ~(0x80000000u >> p);
        const uint32_t lp = (uint32_t)(i * 32 + p);
        if (fast) {                       // k is kc for the rest of this thread's samples (golomb_k_stable)
          const uint32_t x = lp - lpv - 1;
          fastbits += kc + 1 + (x >> kc);
          lpv = lp;
          continue;
        }
        const long long pos = tb + lp;
        const unsigned long long x = (unsigned long long)(pos - pv - 1);
        const uint32_t k = golomb_k(t, (unsigned long long)(pv + 1));
        mybits += k + (x >> k) + 1;
        pv = pos;
        ++t;
        if (first) {
          first = false;
          fast = golomb_k_stable(t, (unsigned long long)(pv + 1), &kc);
          lpv = lp;
        }
      }
    }
    mybits += fastbits;
    g.tbits[tslot] = mybits;
  } else {
    mybits = g.tbits[tslot];
  }
  unsigned long long tot;
  const unsigned long long ex = block_excl_scan_u64(mybits, &tot, s_a);
  if (MODE == 0) {
    if (threadIdx.x == 0) g.bits[blockIdx.x] = tot;
    return;
  }
  // pass 2: scatter. The tile's codewords are one contiguous bit range [o0, o0 + tot) of the stream.
  // When it fits, the range is assembled in shared memory (shared-memory atomics, no L2 round trips)
  // and flushed with coalesced stores; only its first and last word are shared with the neighbouring
  // tiles and need a global atomic. Otherwise (very long unary runs) codewords go straight to global.
  __shared__ uint32_t s_out[GOL_SMEM_WORDS];
  const unsigned long long o0 = g.bits_before[blockIdx.x] + gb.out0;
  const unsigned long long base = o0 & ~31ull;                       // bit position of s_out[0]
  const unsigned long long span_words = ((o0 - base) + tot + 31) >> 5;
  const bool staged = span_words <= GOL_SMEM_WORDS;                  // uniform over the CTA
  if (staged) {
    for (unsigned i = threadIdx.x; i < (unsigned)span_words; i += blockDim.x) s_out[i] = 0;
    __syncthreads();
  }
  unsigned long long o = o0 + ex;
  unsigned long long t = rank0;
  const unsigned long long cmask = (unsigned long long)chunk - 1;  // chunk is a power of two
  const int clog = 31 - __clz(chunk);
  bool first = true, fast = false;
  uint32_t kc = 0, lpv = 0, ob = 0, tl = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t b = v[i];
    while (b) {
      const int p = __clz(b);
      b &= ~(0x80000000u >> p);
      const uint32_t lp = (uint32_t)(i * 32 + p);
      if (fast) {
        // staged tile, k = kc for the rest of this thread's samples: 32-bit arithmetic relative to the
        // thread's first bit (lp, lpv) and to the staged range (ob)
        const uint32_t x = lp - lpv - 1;
        if ((tl & (uint32_t)cmask) == 0) {
          const unsigned long long slot = (t >> clog) - gb.chunk0;
          index[2 * slot] = base + ob - gb.out0 + gb.code0;
          index[2 * slot + 1] = (unsigned long long)(tb + lpv + 1);
        }
        const uint32_t rem = x & ((1u << kc) - 1);
        if (rem) {
          const unsigned long long v64 = (unsigned long long)rem << (64 - (ob & 31) - kc);
          const uint32_t hi = (uint32_t)(v64 >> 32), lo = (uint32_t)v64;
          if (hi) atomicOr(&s_out[ob >> 5], hi);
          if (lo) atomicOr(&s_out[(ob >> 5) + 1], lo);
        }
        const uint32_t sb = ob + kc + (x >> kc);
        atomicOr(&s_out[sb >> 5], 0x80000000u >> (sb & 31));
        ob = sb + 1;
        lpv = lp;
        ++tl;
        ++t;
        continue;
      }
      const long long pos = tb + lp;
      const unsigned long long x = (unsigned long long)(pos - prev - 1);
      const uint32_t k = golomb_k(t, (unsigned long long)(prev + 1));
      if ((t & cmask) == 0) {
        const unsigned long long slot = (t >> clog) - gb.chunk0;
        index[2 * slot] = o - gb.out0 + gb.code0;  // global code-bit offset
        index[2 * slot + 1] = (unsigned long long)(prev + 1);
      }
      const uint32_t rem = (uint32_t)(x & ((1ull << k) - 1));        // k-bit remainder, MSB first
      const unsigned long long stop = o + k + (x >> k);              // x>>k zeros (the buffer is zeroed), then a one
      if (staged) {
        if (rem) {
          const unsigned lo_bit = (unsigned)(o - base);
          const unsigned long long v64 = (unsigned long long)rem << (64 - (lo_bit & 31) - k);
          const uint32_t hi = (uint32_t)(v64 >> 32), lo = (uint32_t)v64;
          if (hi) atomicOr(&s_out[lo_bit >> 5], hi);
          if (lo) atomicOr(&s_out[(lo_bit >> 5) + 1], lo);
        }
        const unsigned sb = (unsigned)(stop - base);
        atomicOr(&s_out[sb >> 5], 0x80000000u >> (sb & 31));
      } else {
        put_bits(out, o, rem, k);
        put_one(out, stop);
      }
      o = stop + 1;
      prev = pos;
      ++t;
      if (first) {
        first = false;
        if (staged) {
          fast = golomb_k_stable(t, (unsigned long long)(prev + 1), &kc);
          lpv = lp;
          ob = (uint32_t)(o - base);
          tl = (uint32_t)t;
        }
      }
    }
  }
  if (staged) {
    __syncthreads();
    uint32_t* gout = out + (base >> 5);
    for (unsigned i = threadIdx.x; i < (unsigned)span_words; i += blockDim.x) {
      const uint32_t w = s_out[i];
      if (!w) continue;
      if (i == 0 || i == (unsigned)span_words - 1) atomicOr(gout + i, bswap32(w));
      else gout[i] = bswap32(w);
    }
  }
  if (gb.closing && blockIdx.x == 0 && threadIdx.x == 0) {  // the run closed by the virtual one (N = global bit count)
    const unsigned long long tt = dyn ? dyn[1] - 1 : gb.close_t, consumed = dyn ? dyn[3] : gb.close_consumed;
    unsigned long long oo = (dyn ? dyn[2] : gb.close_off) + gb.out0;
    const unsigned long long x = N - consumed;
    const uint32_t k = golomb_k(tt, consumed);
    if ((tt & cmask) == 0) { index[2 * ((tt >> clog) - gb.chunk0)] = oo - gb.out0 + gb.code0; index[2 * ((tt >> clog) - gb.chunk0) + 1] = consumed; }
    put_bits(out, oo, (uint32_t)(x & ((1ull << k) - 1)), k);
    oo += k + (x >> k);
    put_one(out, oo);
  }
}

// ------------------------------------------------------------------ single-pass encoder
// The three-kernel pipeline above reads the input three times and needs the bit count on the host before
// the scatter. When the output buffer is already large enough (stream objects are sized for 1.25 code bits
// per input bit) everything runs in ONE kernel: tiles take tickets in order and obtain their two prefixes
// -- (ones, last one) before the tile, then code bits before the tile -- by decoupled look-back over
// per-tile status words; the last tile also writes the closing sample and the totals. If the code would not
// fit, nothing past the capacity is written, the overflow flag is set and the caller falls back to the
// three-kernel path with an exact allocation.
struct GolLook {  // per tile, two chained scans
  unsigned long long agg_ones, pre_ones;
  long long agg_last, pre_last;
  unsigned long long agg_bits, pre_bits;
  volatile unsigned int flag_a;  // 0 none, 1 aggregate, 2 inclusive prefix  (ones/last)
  volatile unsigned int flag_b;  // same for bits
};

__global__ void __launch_bounds__(TILE_THREADS) k_gol_onepass(const uint32_t* __restrict__ S, uint64_t T, uint64_t N, uint64_t ntiles,
                                                              GolLook* __restrict__ look, unsigned int* __restrict__ ticket,
                                                              uint32_t* __restrict__ out, unsigned long long cap_bits,
                                                              unsigned long long* __restrict__ index, uint32_t chunk,
                                                              unsigned long long* __restrict__ scalars /* [0] bits [1] samples [4] overflow */) {
  __shared__ unsigned long long s_a[8];
  __shared__ long long s_b[8];
  __shared__ unsigned int s_tile;
  __shared__ unsigned long long s_ones_before, s_bits_before;
  __shared__ long long s_last_before;
  __shared__ uint32_t s_out[GOL_SMEM_WORDS];
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint64_t tile = s_tile;
  const uint64_t w0 = tile * TILE_WORDS + threadIdx.x * TILE_WORDS_PER_THREAD;
  uint32_t v[4];
  load_tile_words(S, T, w0, v);
  unsigned long long c = 0;
  long long last = -1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    c += __popc(v[i]);
    if (v[i]) last = (long long)((w0 + i) * 32 + (32 - __ffs(v[i])));
  }
  unsigned long long tot_c;
  long long tot_last;
  const unsigned long long ex_c = block_excl_scan_u64(c, &tot_c, s_a);
  const long long pl = block_excl_scan_max(last, &tot_last, s_b);
  if (threadIdx.x == 0) {  // look-back 1: ones and last one before this tile
    GolLook* me = look + tile;
    me->agg_ones = tot_c;
    me->agg_last = tot_last;
    __threadfence();
    me->flag_a = 1;
    unsigned long long ones = 0;
    long long lb = -1;
    for (long long j = (long long)tile - 1; j >= 0; --j) {
      GolLook* q = look + j;
      unsigned int f;
      while ((f = q->flag_a) == 0) {}
      __threadfence();
      if (f == 2) {
        ones += __ldcg(&q->pre_ones);
        const long long pq = __ldcg(&q->pre_last);
        lb = lb > pq ? lb : pq;
        break;
      }
      ones += __ldcg(&q->agg_ones);
      const long long aq = __ldcg(&q->agg_last);
      lb = lb > aq ? lb : aq;
    }
    me->pre_ones = ones + tot_c;
    me->pre_last = lb > tot_last ? lb : tot_last;
    __threadfence();
    me->flag_a = 2;
    s_ones_before = ones;
    s_last_before = lb;
  }
  __syncthreads();
  const unsigned long long rank0 = s_ones_before + ex_c;
  long long prev = pl > s_last_before ? pl : s_last_before;
  unsigned long long mybits = 0;
  {
    unsigned long long t = rank0;
    long long pv = prev;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t b = v[i];
      while (b) {
        const int p = __clz(b);
        b &= ~(0x80000000u >> p);
        const long long pos = (long long)((w0 + i) * 32 + p);
        const unsigned long long x = (unsigned long long)(pos - pv - 1);
        const uint32_t k = golomb_k(t, (unsigned long long)(pv + 1));
        mybits += k + (x >> k) + 1;
        pv = pos;
        ++t;
      }
    }
  }
  unsigned long long tot;
  const unsigned long long ex = block_excl_scan_u64(mybits, &tot, s_a);
  if (threadIdx.x == 0) {  // look-back 2: code bits before this tile
    GolLook* me = look + tile;
    me->agg_bits = tot;
    __threadfence();
    me->flag_b = 1;
    unsigned long long bits = 0;
    for (long long j = (long long)tile - 1; j >= 0; --j) {
      GolLook* q = look + j;
      unsigned int f;
      while ((f = q->flag_b) == 0) {}
      __threadfence();
      if (f == 2) { bits += __ldcg(&q->pre_bits); break; }
      bits += __ldcg(&q->agg_bits);
    }
    me->pre_bits = bits + tot;
    __threadfence();
    me->flag_b = 2;
    s_bits_before = bits;
  }
  __syncthreads();
  const unsigned long long o0 = s_bits_before;
  const bool fits = (o0 + tot + 64 <= cap_bits);  // uniform over the CTA
  if (!fits && threadIdx.x == 0) scalars[4] = 1;
  const unsigned long long base = o0 & ~31ull;
  const unsigned long long span_words = ((o0 - base) + tot + 31) >> 5;
  const bool staged = span_words <= GOL_SMEM_WORDS;
  if (fits) {
    if (staged) {
      for (unsigned i = threadIdx.x; i < (unsigned)span_words; i += blockDim.x) s_out[i] = 0;
      __syncthreads();
    }
    unsigned long long o = o0 + ex;
    unsigned long long t = rank0;
    const unsigned long long cmask = (unsigned long long)chunk - 1;
    const int clog = 31 - __clz(chunk);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t b = v[i];
      while (b) {
        const int p = __clz(b);
        b &= ~(0x80000000u >> p);
        const long long pos = (long long)((w0 + i) * 32 + p);
        const unsigned long long x = (unsigned long long)(pos - prev - 1);
        const uint32_t k = golomb_k(t, (unsigned long long)(prev + 1));
        if ((t & cmask) == 0) { index[2 * (t >> clog)] = o; index[2 * (t >> clog) + 1] = (unsigned long long)(prev + 1); }
        const uint32_t rem = (uint32_t)(x & ((1ull << k) - 1));
        const unsigned long long stop = o + k + (x >> k);
        if (staged) {
          if (rem) {
            const unsigned lo_bit = (unsigned)(o - base);
            const unsigned long long v64 = (unsigned long long)rem << (64 - (lo_bit & 31) - k);
            const uint32_t hi = (uint32_t)(v64 >> 32), lo = (uint32_t)v64;
            if (hi) atomicOr(&s_out[lo_bit >> 5], hi);
            if (lo) atomicOr(&s_out[(lo_bit >> 5) + 1], lo);
          }
          const unsigned sb = (unsigned)(stop - base);
          atomicOr(&s_out[sb >> 5], 0x80000000u >> (sb & 31));
        } else {
          put_bits(out, o, rem, k);
          put_one(out, stop);
        }
        o = stop + 1;
        prev = pos;
        ++t;
      }
    }
    if (staged) {
      __syncthreads();
      uint32_t* gout = out + (base >> 5);
      for (unsigned i = threadIdx.x; i < (unsigned)span_words; i += blockDim.x) {
        const uint32_t w = s_out[i];
        if (!w) continue;
        if (i == 0 || i == (unsigned)span_words - 1) atomicOr(gout + i, bswap32(w));
        else gout[i] = bswap32(w);
      }
    }
  }
  if (tile == ntiles - 1 && threadIdx.x == 0) {  // totals and the run closed by the virtual one
    const unsigned long long ones = s_ones_before + tot_c;
    const long long lastg = s_last_before > tot_last ? s_last_before : tot_last;
    const unsigned long long consumed = (unsigned long long)(lastg + 1);
    unsigned long long oo = o0 + tot;
    const unsigned long long x = N - consumed;
    const uint32_t k = golomb_k(ones, consumed);
    const unsigned long long total = oo + k + (x >> k) + 1;
    scalars[0] = total;
    scalars[1] = ones + 1;
    if (total + 64 <= cap_bits) {
      const unsigned long long cmask = (unsigned long long)chunk - 1;
      const int clog = 31 - __clz(chunk);
      if ((ones & cmask) == 0) { index[2 * (ones >> clog)] = oo; index[2 * (ones >> clog) + 1] = consumed; }
      put_bits(out, oo, (uint32_t)(x & ((1ull << k) - 1)), k);
      oo += k + (x >> k);
      put_one(out, oo);
    } else {
      scalars[4] = 1;
    }
  }
}

// ------------------------------------------------------------------ Golomb decoder: a thread per chunk
__device__ __forceinline__ uint32_t peek32(const uint32_t* __restrict__ in, unsigned long long o) {
  const unsigned long long wi = o >> 5;
  const unsigned off = (unsigned)(o & 31);
  const uint32_t a = bswap32(__ldg(in + wi)), b = bswap32(__ldg(in + wi + 1));
  return __funnelshift_l(b, a, off);
}

__global__ void k_gol_decode(const uint32_t* __restrict__ in, unsigned long long bitcount,
                             const unsigned long long* __restrict__ index, uint64_t nchunks, uint32_t chunk,
                             uint64_t nsamples, uint64_t N, uint32_t* __restrict__ S, unsigned long long* __restrict__ err) {
  for (uint64_t ch = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; ch < nchunks; ch += (uint64_t)gridDim.x * blockDim.x) {
    unsigned long long o = index[2 * ch], pos = index[2 * ch + 1];
    unsigned long long t = ch * chunk;
    const unsigned long long tend = (t + chunk < nsamples) ? t + chunk : nsamples;
    bool bad = (o >= bitcount) || (pos > N);  // every chunk holds at least one codeword (>= 1 bit)
    for (; t < tend && !bad; ++t) {
      const uint32_t k = golomb_k(t, pos);
      const uint32_t rem = k ? (peek32(in, o) >> (32 - k)) : 0u;  // readBits(k), GolombDecoder.cpp:17
      o += k;
      unsigned long long unary = 0;                               // countZeros(), :18
      for (;;) {
        if (o >= bitcount) { bad = true; break; }
        const uint32_t w = peek32(in, o);
        if (w == 0) { unary += 32; o += 32; continue; }
        const int z = __clz(w);
        unary += z;
        o += z + 1;                                               // the closing one, :19
        break;
      }
      if (bad) break;
      pos += (unary << k) | rem;                                  // :21
      if (pos < N) {
        atomicOr(S + (pos >> 5), 0x80000000u >> (unsigned)(pos & 31));
        pos++;
      } else if (pos > N || t + 1 != nsamples) {
        bad = true;
      }
    }
    if (!bad && tend == nsamples && pos != N) bad = true;  // the closing run must end exactly at N
    if (bad) atomicAdd(err, 1ull);
  }
}

// ------------------------------------------------------------------ EG
// codeRun (eg.cpp:20-37) with the coder's block growth disabled (:25) keeps blockSize = 1, and g
// drops from 1 to 0 after the first run that ends in a one. The stream is therefore: a one per
// zero of the input, a zero per one of the input, a one at every end of row, and a single extra
// zero (the one-bit remainder, g = 1) right after the first input one's zero.
__global__ void k_first_one(const uint32_t* __restrict__ S, uint64_t T, unsigned long long* __restrict__ first) {
  unsigned long long best = ~0ull;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < T; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t v = S[t];
    if (v) { const unsigned long long p = t * 32 + __clz(v); best = p < best ? p : best; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long y = __shfl_xor_sync(0xffffffffu, best, o);
    best = y < best ? y : best;
  }
  if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(first, best);
}

__global__ void k_fill_ones(uint32_t* __restrict__ out, unsigned long long nbits) {
  const unsigned long long nw = (nbits + 31) >> 5;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < nw;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    uint32_t v = 0xFFFFFFFFu;
    if (i == nw - 1 && (nbits & 31)) v <<= (32 - (unsigned)(nbits & 31));
    out[i] = bswap32(v);
  }
}

__global__ void k_eg_encode(const uint32_t* __restrict__ S, uint64_t T, uint64_t cols, const unsigned long long* __restrict__ first,
                            uint32_t* __restrict__ out) {
  const unsigned long long f = *first;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < T; t += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t b = S[t];
    while (b) {
      const int p = __clz(b);
      b &= ~(0x80000000u >> p);
      const unsigned long long s = t * 32 + p;
      const unsigned long long o = s + s / cols + (s > f ? 1 : 0);
      atomicAnd(out + (o >> 5), ~bswap32(0x80000000u >> (unsigned)(o & 31)));
      if (s == f) atomicAnd(out + ((o + 1) >> 5), ~bswap32(0x80000000u >> (unsigned)((o + 1) & 31)));
    }
  }
}

// first zero of the code = the first input one's terminator
__global__ void k_first_zero(const uint32_t* __restrict__ in, unsigned long long nbits, unsigned long long* __restrict__ first) {
  const unsigned long long nw = (nbits + 31) >> 5;
  unsigned long long best = ~0ull;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < nw;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t v = ~bswap32(in[i]);
    if (v) { const unsigned long long p = i * 32 + __clz(v); if (p < nbits && p < best) best = p; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long y = __shfl_xor_sync(0xffffffffu, best, o);
    best = y < best ? y : best;
  }
  if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(first, best);
}

__global__ void k_eg_decode(const uint32_t* __restrict__ in, uint64_t rows, uint64_t cols, uint64_t wpr,
                            const unsigned long long* __restrict__ first_zero, uint32_t* __restrict__ M,
                            unsigned long long bitcount, unsigned long long* __restrict__ err) {
  const unsigned long long z = *first_zero;  // code position; ~0 if the matrix is all zero
  const uint64_t total = rows * wpr;
  // what the coder always writes must be there: the extra zero right after the first input one's zero, the one that
  // ends every row, and a length that says whether the extra bit exists
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    bool bad = (z == ~0ull) ? (bitcount != rows * (cols + 1)) : (bitcount != rows * (cols + 1) + 1);
    if (!bad && z != ~0ull) {
      const unsigned long long o = z + 1;
      bad = (bswap32(__ldg(in + (o >> 5))) >> (31 - (unsigned)(o & 31))) & 1u;
    }
    if (bad) atomicAdd(err, 1ull);
  }
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = i / wpr, w = i - r * wpr;
    if (w == wpr - 1) {  // end-of-row one
      unsigned long long o = r * (cols + 1) + cols;
      if (o > z) o += 1;
      if (o >= bitcount || !((bswap32(__ldg(in + (o >> 5))) >> (31 - (unsigned)(o & 31))) & 1u)) atomicAdd(err, 1ull);
    }
    uint32_t v = 0;
    for (int b = 0; b < 32; ++b) {
      const uint64_t cc = w * 32 + b;
      if (cc >= cols) break;
      unsigned long long o = r * (cols + 1) + cc;  // position without the extra bit
      if (o > z) o += 1;
      const uint32_t word = bswap32(__ldg(in + (o >> 5)));
      if (!((word >> (31 - (unsigned)(o & 31))) & 1u)) v |= 0x80000000u >> b;
    }
    M[i] = v;
  }
}

// ------------------------------------------------------------------ host side
extern "C" bic_status bic_stream_create(bic_ctx* c, bic_stream** out) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !out) return BIC_ERR_INVALID;
  *out = new (std::nothrow) bic_stream();
  return *out ? BIC_OK : BIC_ERR_NOMEM;
}

extern "C" bic_status bic_stream_destroy(bic_ctx* c, bic_stream* s) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !s) return BIC_ERR_INVALID;
  BIC_CUDA(c, bic_wait_stream(c));
  if (s->d_bytes) cudaFree(s->d_bytes);
  if (s->d_index) cudaFree(s->d_index);
  delete s;
  return BIC_OK;
}

extern "C" bic_status bic_stream_get_info(const bic_stream* s, bic_stream_info* info) {
  if (!s || !info) return BIC_ERR_INVALID;
  *info = s->info;
  return BIC_OK;
}

static bic_status stream_reserve(bic_ctx* c, bic_stream* s, uint64_t bitcount, uint64_t nchunks, uint64_t src_bits = 0) {
  if (bitcount > (1ull << 45) || nchunks > (1ull << 40)) return bic_fail(c, BIC_ERR_INVALID, "stream size overflows");
  // whole 32-bit words plus slack so peek32 / put_bits may touch one word past the end
  size_t need = (size_t)(div_up_u64(bitcount, 32) * 4 + 16);
  // cudaMalloc/cudaFree stall every stream of the device, so a stream object is sized once for
  // what its source can plausibly produce (1.25 bits per input bit) instead of growing plane by plane
  const size_t typical = (size_t)(src_bits / 8 + src_bits / 32) + 4096;
  if (need > s->cap_bytes && typical > need) need = typical;
  if (need > s->cap_bytes) {
    BIC_CUDA(c, bic_wait_stream(c));
    bic_free_device(c, s->d_bytes);
    s->d_bytes = nullptr; s->cap_bytes = 0;
    const size_t want = (need + (need >> 3) + 255) & ~(size_t)255;
    if (cudaMalloc(&s->d_bytes, want) != cudaSuccess) { cudaGetLastError(); c->err = "stream allocation failed"; return BIC_ERR_NOMEM; }
    s->cap_bytes = want;
  }
  size_t needi = (size_t)(nchunks ? nchunks : 1) * 2;
  const size_t typicali = (size_t)(src_bits / 256 + 64) * 2;  // one chunk per 256 samples, <= 1 sample per bit... /2 on average
  if (needi > s->cap_index && typicali > needi) needi = typicali;
  if (needi > s->cap_index) {
    BIC_CUDA(c, bic_wait_stream(c));
    bic_free_device(c, s->d_index);
    s->d_index = nullptr; s->cap_index = 0;
    if (cudaMalloc(&s->d_index, needi * 8) != cudaSuccess) { cudaGetLastError(); c->err = "index allocation failed"; return BIC_ERR_NOMEM; }
    s->cap_index = needi;
  }
  return BIC_OK;
}

// dense stream view of a matrix: the matrix itself when cols % 32 == 0, else a compacted copy
static bic_status dense_stream(bic_ctx* c, const bic_mat* M, const uint32_t** S, uint64_t* T) {
  const uint64_t N = M->rows * M->cols;
  *T = div_up_u64(N, 32);
  if ((M->cols & 31) == 0) { *S = M->d; return BIC_OK; }
  BIC_TRY(bic_scratch_reserve(c, &c->work[4], (size_t)(*T) * 4 + 16));
  if (*T) {
    BIC_PROF(c, KID_COMPACT_ROWS);
    k_compact_rows<<<bic_grid_for(c, *T, 256, 8), 256, 0, c->stream>>>(M->d, M->cols, M->wpr, N, (uint32_t*)c->work[4].p, *T);
    BIC_LAUNCH_CHECK(c);
  }
  *S = (const uint32_t*)c->work[4].p;
  return BIC_OK;
}

bic_status bic_stream_reserve_for(bic_ctx* c, bic_stream* s, uint64_t bitcount, uint64_t nchunks, uint64_t src_bits) {
  return stream_reserve(c, s, bitcount, nchunks, src_bits);
}

// dense stream of M for coding2.cu: the matrix itself, or a compacted copy written to `scratch` (16-byte aligned)
bic_status bic_dense_stream_into(bic_ctx* c, const bic_mat* M, uint32_t* scratch, const uint32_t** S, uint64_t* T) {
  const uint64_t N = M->rows * M->cols;
  *T = div_up_u64(N, 32);
  if ((M->cols & 31) == 0) { *S = M->d; return BIC_OK; }
  if (*T) {
    BIC_PROF(c, KID_COMPACT_ROWS);
    k_compact_rows<<<bic_grid_for(c, *T, 256, 8), 256, 0, c->stream>>>(M->d, M->cols, M->wpr, N, scratch, *T);
    BIC_LAUNCH_CHECK(c);
  }
  *S = scratch;
  return BIC_OK;
}

bic_status bic_k_golomb_encode_multi(bic_ctx* c, const bic_mat* const* mats, int nmat, uint32_t chunk_samples, bic_stream* const* outs,
                                     unsigned long long* d_info);

static GolBase gol_base_single() {
  GolBase b;
  memset(&b, 0, sizeof(b));
  b.prev0 = -1;
  b.closing = 1;
  return b;
}

struct GolWork {
  const uint32_t* S;
  uint64_t T, N, ntiles;
  GolTile g;
};

// per-tile counts and their scans. Afterwards h_scalars[1] = ones + 1 and h_scalars[3] = position after the
// matrix's last one (0 if none); [0] and [2] are not meaningful yet.
static bic_status golomb_counts(bic_ctx* c, const bic_mat* M, GolWork* w) {
  w->N = M->rows * M->cols;
  BIC_TRY(dense_stream(c, M, &w->S, &w->T));
  w->ntiles = div_up_u64(w->T, TILE_WORDS);
  const size_t per = (size_t)(w->ntiles ? w->ntiles : 1);  // work[5]: per-tile arrays
  BIC_TRY(bic_scratch_reserve(c, &c->work[5], per * (8 * 5 + 8) + per * TILE_THREADS * 8));
  uint8_t* p = (uint8_t*)c->work[5].p;
  GolTile* g = &w->g;
  g->last = (long long*)p;
  g->ones_before = (unsigned long long*)(p + per * 8);
  g->last_before = (long long*)(p + per * 16);
  g->bits = (unsigned long long*)(p + per * 24);
  g->bits_before = (unsigned long long*)(p + per * 32);
  g->ones = (uint32_t*)(p + per * 40);
  g->tbits = (unsigned long long*)(p + per * 48);
  BIC_CUDA(c, cudaMemsetAsync(g->bits, 0, per * 8, c->stream));
  if (!w->ntiles) BIC_CUDA(c, cudaMemsetAsync(g->tbits, 0, per * TILE_THREADS * 8, c->stream));  // the lone CTA of an empty matrix reads them
  if (w->ntiles) {
    BIC_PROF(c, KID_GOL_TILE_COUNTS);
    k_gol_tile_counts<<<(unsigned)w->ntiles, TILE_THREADS, 0, c->stream>>>(w->S, w->T, *g);
    BIC_LAUNCH_CHECK(c);
    BIC_PROF(c, KID_GOL_SCAN_A);
    k_gol_scan_tiles_a<<<1, SCAN_THREADS, 0, c->stream>>>(*g, w->ntiles);
    BIC_LAUNCH_CHECK(c);
  }
  return BIC_OK;
}

// code length of every tile under `base`, then their scan. Afterwards h_scalars: [0] bit count including the
// closing sample (only meaningful for the single-stream base), [1] ones + 1, [2] code bits of the matrix's own
// ones, [3] position after its last one.
static bic_status golomb_lengths(bic_ctx* c, GolWork* w, const GolBase& base) {
  if (w->ntiles) {
    BIC_PROF(c, KID_GOL_LENGTHS);
    k_gol_walk<0><<<(unsigned)w->ntiles, TILE_THREADS, 0, c->stream>>>(w->S, w->T, w->N, w->g, nullptr, nullptr, 1, base, nullptr);
    BIC_LAUNCH_CHECK(c);
  }
  BIC_PROF(c, KID_GOL_SCAN_B);
  k_gol_scan_tiles_b<<<1, SCAN_THREADS, 0, c->stream>>>(w->g, w->ntiles, w->N, (unsigned long long*)c->d_scalars, 0ull);
  BIC_LAUNCH_CHECK(c);
  return bic_read_scalars(c, 4);
}

static bic_status golomb_scatter(bic_ctx* c, GolWork* w, const GolBase& base, uint64_t N_global, uint32_t chunk, bic_stream* out,
                                 const unsigned long long* dyn = nullptr) {
  // with no tile (empty matrix) one CTA still has to write the closing sample
  const unsigned grid = (unsigned)(w->ntiles ? w->ntiles : 1);
  BIC_PROF(c, KID_GOL_SCATTER);
  k_gol_walk<1><<<grid, TILE_THREADS, 0, c->stream>>>(w->S, w->ntiles ? w->T : 0, N_global, w->g, (uint32_t*)out->d_bytes,
                                                     (unsigned long long*)out->d_index, chunk, base, dyn);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

// The encoder with nothing waiting for the host (pipeline.cu): counts, lengths, scans, clear and scatter are queued back to
// back. The output buffer is sized for what its source can plausibly produce (stream_reserve: 1.25 code bits per input bit);
// d_info (device, 8 u64) receives [0] bit count, [1] samples, [2] offset of the closing sample, [3] position after the last
// one, [4] 1 if the code did not fit (nothing was written: re-encode with bic_golomb_encode). out->info is NOT filled in:
// the caller completes it from d_info once it has been copied to the host (bic_golomb_async_finish).
bic_status bic_k_golomb_encode_async(bic_ctx* c, const bic_mat* M, uint32_t chunk_samples, bic_stream* out, unsigned long long* d_info) {
  BIC_RANGE("bic:golomb_encode(async)");
  const uint64_t N = M->rows * M->cols;
  BIC_TRY(stream_reserve(c, out, N + N / 4 + 32768, div_up_u64(N + 1, chunk_samples), N));  // index: worst case, every bit a sample
  const uint64_t cap_bits = (uint64_t)(out->cap_bytes - 32) * 8;
  GolWork w;
  BIC_TRY(golomb_counts(c, M, &w));
  GolBase base = gol_base_single();
  if (w.ntiles) {
    BIC_PROF(c, KID_GOL_LENGTHS);
    k_gol_walk<0><<<(unsigned)w.ntiles, TILE_THREADS, 0, c->stream>>>(w.S, w.T, w.N, w.g, nullptr, nullptr, 1, base, nullptr);
    BIC_LAUNCH_CHECK(c);
  }
  BIC_PROF(c, KID_GOL_SCAN_B);
  k_gol_scan_tiles_b<<<1, SCAN_THREADS, 0, c->stream>>>(w.g, w.ntiles, w.N, d_info, cap_bits);
  BIC_LAUNCH_CHECK(c);
  BIC_PROF(c, KID_GOL_SCAN_B);
  k_gol_zero_code<<<bic_grid_for(c, div_up_u64(N, 32) + 4, 256, 4), 256, 0, c->stream>>>((uint32_t*)out->d_bytes, d_info);
  BIC_LAUNCH_CHECK(c);
  BIC_TRY(golomb_scatter(c, &w, base, w.N, chunk_samples, out, d_info));
  out->info.coder = BIC_CODER_GOLOMB;
  out->info.chunk_samples = chunk_samples;
  out->info.rows = M->rows;
  out->info.cols = M->cols;
  out->info.bitcount = out->info.nsamples = out->info.nchunks = 0;
  return BIC_OK;
}

// host_info: the 8 u64 of d_info after they reached the host. Returns BIC_ERR_CAPACITY if the code did not fit.
bic_status bic_golomb_async_finish(bic_stream* out, const uint64_t* host_info) {
  if (host_info[4]) return BIC_ERR_CAPACITY;
  out->info.bitcount = host_info[0];
  out->info.nsamples = host_info[1];
  out->info.nchunks = div_up_u64(host_info[1], out->info.chunk_samples);
  return BIC_OK;
}

extern "C" bic_status bic_golomb_bitcount(bic_ctx* c, const bic_mat* M, uint64_t* bitcount, uint64_t* nsamples) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !M) return BIC_ERR_INVALID;
  GolWork w;
  BIC_TRY(golomb_counts(c, M, &w));
  BIC_TRY(golomb_lengths(c, &w, gol_base_single()));
  if (bitcount) *bitcount = c->h_scalars[0];
  if (nsamples) *nsamples = c->h_scalars[1];
  return BIC_OK;
}

bic_status bic_golomb_async_finish(bic_stream* out, const uint64_t* host_info);

extern "C" bic_status bic_golomb_encode(bic_ctx* c, const bic_mat* M, uint32_t chunk_samples, bic_stream* out) {
  BIC_RANGE("bic:golomb_encode");
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !M || !out) return BIC_ERR_INVALID;
  if (chunk_samples == 0) chunk_samples = 256;
  while (chunk_samples & (chunk_samples - 1)) chunk_samples++;  // the kernels index chunks with shifts: round up to a power of two
  GolWork w;
  if (c->gol_algo == 2 && M->rows * M->cols > 0 && !c->gol_onepass) {
    // wide-tile encoder (coding2.cu) into a pre-sized buffer; the exact-size path below takes over if the code does not fit
    unsigned long long* d_info = (unsigned long long*)(c->d_scalars + 56);
    const bic_mat* mats[1] = {M};
    bic_stream* outs[1] = {out};
    BIC_TRY(bic_k_golomb_encode_multi(c, mats, 1, chunk_samples, outs, d_info));
    BIC_CUDA(c, cudaMemcpyAsync(c->h_scalars + 56, d_info, 8 * 8, cudaMemcpyDeviceToHost, c->stream));
    BIC_CUDA(c, bic_wait_stream(c));
    if (bic_golomb_async_finish(out, c->h_scalars + 56) == BIC_OK) return BIC_OK;
  }
  {  // single pass into the pre-sized buffer (the common case); falls through to the exact path on overflow
    const uint64_t N = M->rows * M->cols;
    const uint64_t T = div_up_u64(N, 32), ntiles = div_up_u64(T, TILE_WORDS);
    BIC_TRY(stream_reserve(c, out, 0, div_up_u64(N + 1, chunk_samples), N));  // worst case: every bit a sample
    const uint64_t cap_bits = (uint64_t)(out->cap_bytes - 16) * 8;
    if (ntiles && c->gol_onepass) {
      const uint32_t* S;
      uint64_t T2;
      BIC_TRY(dense_stream(c, M, &S, &T2));
      BIC_TRY(bic_scratch_reserve(c, &c->work[5], ntiles * sizeof(GolLook) + 64));
      GolLook* look = (GolLook*)((uint8_t*)c->work[5].p + 64);
      unsigned int* ticket = (unsigned int*)c->work[5].p;
      BIC_CUDA(c, cudaMemsetAsync(c->work[5].p, 0, ntiles * sizeof(GolLook) + 64, c->stream));
      BIC_CUDA(c, cudaMemsetAsync(c->d_scalars, 0, 5 * sizeof(uint64_t), c->stream));
      BIC_CUDA(c, cudaMemsetAsync(out->d_bytes, 0, out->cap_bytes, c->stream));
      BIC_PROF(c, KID_GOL_SCATTER);
      k_gol_onepass<<<(unsigned)ntiles, TILE_THREADS, 0, c->stream>>>(S, T2, N, ntiles, look, ticket, (uint32_t*)out->d_bytes, cap_bits,
                                                                    (unsigned long long*)out->d_index, chunk_samples,
                                                                    (unsigned long long*)c->d_scalars);
      BIC_LAUNCH_CHECK(c);
      BIC_TRY(bic_read_scalars(c, 5));
      if (!c->h_scalars[4]) {
        out->info.coder = BIC_CODER_GOLOMB;
        out->info.chunk_samples = chunk_samples;
        out->info.rows = M->rows;
        out->info.cols = M->cols;
        out->info.bitcount = c->h_scalars[0];
        out->info.nsamples = c->h_scalars[1];
        out->info.nchunks = div_up_u64(c->h_scalars[1], chunk_samples);
        return BIC_OK;
      }
    }
  }
  BIC_TRY(golomb_counts(c, M, &w));
  GolBase base = gol_base_single();
  BIC_TRY(golomb_lengths(c, &w, base));
  const uint64_t bitcount = c->h_scalars[0], nsamples = c->h_scalars[1];
  base.close_t = nsamples - 1;
  base.close_off = c->h_scalars[2];
  base.close_consumed = c->h_scalars[3];
  const uint64_t nchunks = div_up_u64(nsamples, chunk_samples);
  BIC_TRY(stream_reserve(c, out, bitcount, nchunks, w.N));
  BIC_CUDA(c, cudaMemsetAsync(out->d_bytes, 0, (size_t)(div_up_u64(bitcount, 32) * 4 + 16), c->stream));
  BIC_TRY(golomb_scatter(c, &w, base, w.N, chunk_samples, out));
  out->info.coder = BIC_CODER_GOLOMB;
  out->info.chunk_samples = chunk_samples;
  out->info.rows = M->rows;
  out->info.cols = M->cols;
  out->info.bitcount = bitcount;
  out->info.nsamples = nsamples;
  out->info.nchunks = nchunks;
  return BIC_OK;
}

// ------------------------------------------------------------------ a shard through the wide-tile encoder (coding2.cu)
bic_status bic_g2_plan(bic_ctx* c, const bic_mat* M, void* plan_out, size_t plan_bytes);
bic_status bic_g2_count(bic_ctx* c, void* plan, unsigned long long* d_tot);
bic_status bic_g2_lengths(bic_ctx* c, void* plan, const GolBase* gb, unsigned long long** d_info);
bic_status bic_g2_scatter(bic_ctx* c, void* plan, const GolBase* gb, uint32_t chunk, bic_stream* out);

// what the caller knows about the rest of the matrix: called with this shard's own numbers, fills in the global ones
struct ShardPrefix {   // after pass 1
  uint64_t ones_before, bits_before, ones_global, bits_global;
  long long last_before, last_global;   // global positions of the last one before this shard / of the whole matrix, -1 if none
  int is_last;
};
typedef bic_status (*shard_prefix_fn)(void* user, bic_ctx* c, uint64_t ones, uint64_t lastpos1, uint64_t nbits, ShardPrefix* out);
typedef bic_status (*shard_code0_fn)(void* user, bic_ctx* c, uint64_t my_code_bits, uint64_t* code_bits_before, uint64_t* code_bits_total);

static uint32_t golomb_k_host(uint64_t t64, uint64_t bits_consumed);

static bic_status golomb_shard_g2(bic_ctx* c, const bic_mat* M, uint32_t chunk_samples, bic_stream* out, bic_shard_info* shard,
                                  shard_prefix_fn prefix, shard_code0_fn code0fn, void* user) {
  alignas(16) unsigned char plan[4096];
  BIC_TRY(bic_g2_plan(c, M, plan, sizeof(plan)));
  unsigned long long* d_tot = (unsigned long long*)(c->d_scalars + 58);
  BIC_TRY(bic_g2_count(c, plan, d_tot));
  BIC_CUDA(c, cudaMemcpyAsync(c->h_scalars + 58, d_tot, 16, cudaMemcpyDeviceToHost, c->stream));
  BIC_CUDA(c, bic_wait_stream(c));
  const uint64_t ones = c->h_scalars[58], lastpos1 = c->h_scalars[59], N = M->rows * M->cols;
  ShardPrefix px;
  BIC_TRY(prefix(user, c, ones, lastpos1, N, &px));
  GolBase base = gol_base_single();
  base.closing = 0;
  base.t0 = px.ones_before; base.pos0 = (long long)px.bits_before; base.prev0 = px.last_before;
  unsigned long long* d_info = nullptr;
  BIC_TRY(bic_g2_lengths(c, plan, &base, &d_info));
  BIC_CUDA(c, cudaMemcpyAsync(c->h_scalars + 56, d_info + 2, 8, cudaMemcpyDeviceToHost, c->stream));
  BIC_CUDA(c, bic_wait_stream(c));
  const uint64_t bits_mine = c->h_scalars[56];
  uint64_t code0 = 0, code_total = 0;
  BIC_TRY(code0fn(user, c, bits_mine, &code0, &code_total));
  const uint64_t consumed = (uint64_t)(px.last_global + 1);
  const uint32_t kc = golomb_k_host(px.ones_global, consumed);
  const uint64_t closing_bits = kc + ((px.bits_global - consumed) >> kc) + 1;
  base.code0 = code0;
  base.out0 = code0 & 31;
  base.chunk0 = div_up_u64(base.t0, chunk_samples);
  uint64_t local_bits = bits_mine, local_samples = ones;
  if (px.is_last) {  // the last shard also writes the run closed by the virtual one
    base.closing = 1;
    base.close_t = px.ones_global;
    base.close_consumed = consumed;
    base.close_off = bits_mine;
    base.close_n = px.bits_global;
    local_bits += closing_bits;
    local_samples += 1;
  }
  const uint64_t nchunks = div_up_u64(base.t0 + local_samples, chunk_samples) - base.chunk0;
  if (shard) {
    // code_total == ~0: the caller does not know the other shards' lengths; then only the last shard can tell the global count
    shard->global_bitcount = code_total != ~0ull ? code_total + closing_bits : (px.is_last ? code0 + local_bits : 0);
    shard->global_nsamples = px.ones_global + 1;
    shard->code_bit_offset = code0;
    shard->local_code_bits = local_bits;
    shard->first_chunk = base.chunk0;
    shard->local_chunks = nchunks;
  }
  if (!out) return BIC_OK;
  BIC_TRY(stream_reserve(c, out, base.out0 + local_bits, nchunks, N));
  BIC_CUDA(c, cudaMemsetAsync(out->d_bytes, 0, (size_t)(div_up_u64(base.out0 + local_bits, 32) * 4 + 16), c->stream));
  BIC_TRY(bic_g2_scatter(c, plan, &base, chunk_samples, out));
  out->info.coder = BIC_CODER_GOLOMB;
  out->info.chunk_samples = chunk_samples;
  out->info.rows = M->rows;
  out->info.cols = M->cols;
  out->info.bitcount = base.out0 + local_bits;  // bits of the local buffer, incl. the (code0 & 31) leading pad
  out->info.nsamples = local_samples;
  out->info.nchunks = nchunks;
  return BIC_OK;
}

// ------------------------------------------------------------------ row-sharded coding (several GPUs)
struct bic_comm;
bic_status bic_comm_allgather_u64(bic_ctx* c, bic_comm* m, const uint64_t* mine, int count, uint64_t* all);
int bic_comm_rank(const bic_comm* m);
int bic_comm_size(const bic_comm* m);

static uint32_t golomb_k_host(uint64_t t64, uint64_t bits_consumed) {  // twin of the device golomb_k
  if (t64 == 0) return 1;
  const uint32_t t = (uint32_t)t64, acc = (uint32_t)(bits_consumed - t64);
  if (acc <= t) return 0;
  uint32_t k = 0;
  while (k < 31 && (uint32_t)(t << k) < acc) k++;
  return k;
}

// Each rank holds a contiguous block of rows of ONE matrix (rank order = row order). Every rank codes its rows
// as the exact substring of the single global Golomb stream: the coder state at a shard boundary is a closed
// form of (ones before the shard, position of the last one before it), which two tiny allgathers provide.
// The shard's bits start at bit (code_bit_offset & 31) of its own buffer, so the global stream is the
// word-wise OR of the shards placed at word (code_bit_offset >> 5).
extern "C" bic_status bic_dist_golomb_encode(bic_ctx* c, bic_comm* m, const bic_mat* M, uint32_t chunk_samples, bic_stream* out,
                                             bic_shard_info* shard) {
  if (!c || !m || !M || !out) return BIC_ERR_INVALID;
  cudaSetDevice(c->device);
  if (chunk_samples == 0) chunk_samples = 256;
  while (chunk_samples & (chunk_samples - 1)) chunk_samples++;
  const int rank = bic_comm_rank(m), nr = bic_comm_size(m);
  if (c->gol_algo == 2 && M->rows * M->cols > 0) {
    struct U { bic_comm* m; int rank, nr; } u{m, rank, nr};
    auto prefix = [](void* user, bic_ctx* cc, uint64_t ones, uint64_t lastpos1, uint64_t nbits, ShardPrefix* px) -> bic_status {
      U* uu = (U*)user;
      uint64_t mine[3] = {ones, lastpos1, nbits}, all[8 * 3];
      BIC_TRY(bic_comm_allgather_u64(cc, uu->m, mine, 3, all));
      memset(px, 0, sizeof(*px));
      px->last_before = -1; px->last_global = -1;
      uint64_t pos = 0, og = 0;
      for (int r = 0; r < uu->nr; ++r) {
        if (r == uu->rank) { px->ones_before = og; px->bits_before = pos; px->last_before = px->last_global; }
        if (all[r * 3 + 1]) px->last_global = (long long)(pos + all[r * 3 + 1] - 1);
        og += all[r * 3];
        pos += all[r * 3 + 2];
      }
      px->ones_global = og; px->bits_global = pos; px->is_last = uu->rank == uu->nr - 1;
      return BIC_OK;
    };
    auto code0fn = [](void* user, bic_ctx* cc, uint64_t mybits, uint64_t* before, uint64_t* total) -> bic_status {
      U* uu = (U*)user;
      uint64_t all[8];
      BIC_TRY(bic_comm_allgather_u64(cc, uu->m, &mybits, 1, all));
      *before = 0; *total = 0;
      for (int r = 0; r < uu->nr; ++r) { if (r < uu->rank) *before += all[r]; *total += all[r]; }
      return BIC_OK;
    };
    return golomb_shard_g2(c, M, chunk_samples, out, shard, prefix, code0fn, &u);
  }
  GolWork w;
  BIC_TRY(golomb_counts(c, M, &w));
  BIC_PROF(c, KID_GOL_SCAN_B);
  k_gol_scan_tiles_b<<<1, SCAN_THREADS, 0, c->stream>>>(w.g, w.ntiles, w.N, (unsigned long long*)c->d_scalars, 0ull);
  BIC_LAUNCH_CHECK(c);
  BIC_TRY(bic_read_scalars(c, 4));
  uint64_t mine[3] = {c->h_scalars[1] - 1, c->h_scalars[3], w.N};  // ones, position after the last one (0 = none), bits
  uint64_t all[8 * 3];
  BIC_TRY(bic_comm_allgather_u64(c, m, mine, 3, all));
  GolBase base = gol_base_single();
  base.closing = 0;
  uint64_t N_global = 0, ones_global = 0;
  long long last_global = -1;
  {
    uint64_t pos = 0;
    for (int r = 0; r < nr; ++r) {
      if (r == rank) { base.t0 = ones_global; base.pos0 = (long long)pos; base.prev0 = last_global; }
      if (all[r * 3 + 1]) last_global = (long long)(pos + all[r * 3 + 1] - 1);
      ones_global += all[r * 3];
      pos += all[r * 3 + 2];
    }
    N_global = pos;
  }
  BIC_TRY(golomb_lengths(c, &w, base));
  uint64_t bits_mine = c->h_scalars[2], bits_all[8];
  BIC_TRY(bic_comm_allgather_u64(c, m, &bits_mine, 1, bits_all));
  uint64_t code0 = 0, code_total = 0;
  for (int r = 0; r < nr; ++r) { if (r < rank) code0 += bits_all[r]; code_total += bits_all[r]; }
  const uint64_t consumed = (uint64_t)(last_global + 1);
  const uint32_t kc = golomb_k_host(ones_global, consumed);
  const uint64_t closing_bits = kc + ((N_global - consumed) >> kc) + 1;
  base.code0 = code0;
  base.out0 = code0 & 31;
  base.chunk0 = div_up_u64(base.t0, chunk_samples);
  uint64_t local_bits = bits_mine, local_samples = mine[0];
  if (rank == nr - 1) {  // the last rank also writes the run closed by the virtual one
    base.closing = 1;
    base.close_t = ones_global;
    base.close_consumed = consumed;
    base.close_off = bits_mine;
    local_bits += closing_bits;
    local_samples += 1;
  }
  const uint64_t chunk_end = div_up_u64(base.t0 + local_samples, chunk_samples);  // first chunk the NEXT shard stores
  const uint64_t nchunks = chunk_end - base.chunk0;
  BIC_TRY(stream_reserve(c, out, base.out0 + local_bits, nchunks, w.N));
  BIC_CUDA(c, cudaMemsetAsync(out->d_bytes, 0, (size_t)(div_up_u64(base.out0 + local_bits, 32) * 4 + 16), c->stream));
  BIC_TRY(golomb_scatter(c, &w, base, N_global, chunk_samples, out));
  out->info.coder = BIC_CODER_GOLOMB;
  out->info.chunk_samples = chunk_samples;
  out->info.rows = M->rows;
  out->info.cols = M->cols;
  out->info.bitcount = base.out0 + local_bits;  // bits of the local buffer, incl. the (code0 & 31) leading pad
  out->info.nsamples = local_samples;
  out->info.nchunks = nchunks;
  if (shard) {
    shard->global_bitcount = code_total + closing_bits;
    shard->global_nsamples = ones_global + 1;
    shard->code_bit_offset = code0;
    shard->local_code_bits = local_bits;
    shard->first_chunk = base.chunk0;
    shard->local_chunks = nchunks;
  }
  return BIC_OK;
}

// The same for a caller that does its own plumbing (MPI, files, a second process ...): the prefix state of the shard is an
// explicit input. ones_before / bits_before / last_one_before describe the rows that precede M in the one global matrix,
// code_bits_before the code bits their ones produced (0 is fine when only the shard's own length is wanted: pass out = NULL).
extern "C" bic_status bic_golomb_encode_shard(bic_ctx* c, const bic_mat* M, uint32_t chunk_samples, uint64_t ones_before,
                                              uint64_t bits_before, int64_t last_one_before, uint64_t code_bits_before, int closing,
                                              uint64_t total_bits, bic_stream* out, bic_shard_info* shard) {
  if (!c || !M) return BIC_ERR_INVALID;
  cudaSetDevice(c->device);
  if (last_one_before < -1 || (last_one_before >= 0 && (uint64_t)last_one_before >= bits_before) ||
      (closing && total_bits < bits_before + M->rows * M->cols))
    return bic_fail(c, BIC_ERR_INVALID, "golomb_encode_shard: inconsistent prefix state");
  if (chunk_samples == 0) chunk_samples = 256;
  while (chunk_samples & (chunk_samples - 1)) chunk_samples++;
  if (c->gol_algo == 2 && M->rows * M->cols > 0) {
    struct U { uint64_t ones_before, bits_before, code_before, total_bits; long long last_before; int closing; }
        u{ones_before, bits_before, code_bits_before, total_bits, (long long)last_one_before, closing};
    auto prefix = [](void* user, bic_ctx*, uint64_t ones, uint64_t lastpos1, uint64_t nbits, ShardPrefix* px) -> bic_status {
      U* uu = (U*)user;
      memset(px, 0, sizeof(*px));
      px->ones_before = uu->ones_before; px->bits_before = uu->bits_before; px->last_before = uu->last_before;
      px->ones_global = uu->ones_before + ones;          // meaningful for the closing shard only (nothing follows it)
      px->bits_global = uu->closing ? uu->total_bits : uu->bits_before + nbits;
      px->last_global = lastpos1 ? (long long)(uu->bits_before + lastpos1 - 1) : uu->last_before;
      px->is_last = uu->closing != 0;
      return BIC_OK;
    };
    auto code0fn = [](void* user, bic_ctx*, uint64_t, uint64_t* before, uint64_t* total) -> bic_status {
      U* uu = (U*)user;
      *before = uu->code_before; *total = ~0ull;         // the total is only known to whoever holds every shard
      return BIC_OK;
    };
    bic_shard_info si;
    memset(&si, 0, sizeof(si));
    BIC_TRY(golomb_shard_g2(c, M, chunk_samples, out, &si, prefix, code0fn, &u));
    if (!closing) { si.global_bitcount = 0; si.global_nsamples = 0; }   // only the last shard knows them
    if (shard) *shard = si;
    return BIC_OK;
  }
  GolWork w;
  BIC_TRY(golomb_counts(c, M, &w));
  GolBase base = gol_base_single();
  base.closing = 0;
  base.t0 = ones_before; base.pos0 = (long long)bits_before; base.prev0 = (long long)last_one_before;
  BIC_TRY(golomb_lengths(c, &w, base));
  const uint64_t ones = c->h_scalars[1] - 1, bits_mine = c->h_scalars[2];
  const long long last_global = c->h_scalars[3] ? (long long)(bits_before + c->h_scalars[3] - 1) : (long long)last_one_before;
  uint64_t local_bits = bits_mine, local_samples = ones, closing_bits = 0;
  base.code0 = code_bits_before;
  base.out0 = code_bits_before & 31;
  base.chunk0 = div_up_u64(base.t0, chunk_samples);
  if (closing) {
    const uint64_t consumed = (uint64_t)(last_global + 1);
    const uint32_t kc = golomb_k_host(ones_before + ones, consumed);
    closing_bits = kc + ((total_bits - consumed) >> kc) + 1;
    base.closing = 1;
    base.close_t = ones_before + ones;
    base.close_consumed = consumed;
    base.close_off = bits_mine;
    local_bits += closing_bits;
    local_samples += 1;
  }
  const uint64_t nchunks = div_up_u64(base.t0 + local_samples, chunk_samples) - base.chunk0;
  if (shard) {
    memset(shard, 0, sizeof(*shard));
    shard->global_bitcount = closing ? code_bits_before + local_bits : 0;   // only the last shard knows it
    shard->global_nsamples = closing ? ones_before + ones + 1 : 0;
    shard->code_bit_offset = code_bits_before;
    shard->local_code_bits = local_bits;
    shard->first_chunk = base.chunk0;
    shard->local_chunks = nchunks;
  }
  if (!out) return BIC_OK;
  BIC_TRY(stream_reserve(c, out, base.out0 + local_bits, nchunks, w.N));
  BIC_CUDA(c, cudaMemsetAsync(out->d_bytes, 0, (size_t)(div_up_u64(base.out0 + local_bits, 32) * 4 + 16), c->stream));
  BIC_TRY(golomb_scatter(c, &w, base, closing ? total_bits : 0, chunk_samples, out));
  out->info.coder = BIC_CODER_GOLOMB;
  out->info.chunk_samples = chunk_samples;
  out->info.rows = M->rows;
  out->info.cols = M->cols;
  out->info.bitcount = base.out0 + local_bits;
  out->info.nsamples = local_samples;
  out->info.nchunks = nchunks;
  return BIC_OK;
}

extern "C" bic_status bic_golomb_decode(bic_ctx* c, const bic_stream* s, bic_mat* M) {
  BIC_RANGE("bic:golomb_decode");
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !s || !M) return BIC_ERR_INVALID;
  if (s->info.coder != BIC_CODER_GOLOMB) return bic_fail(c, BIC_ERR_INVALID, "golomb_decode: not a Golomb stream");
  if (M->rows != s->info.rows || M->cols != s->info.cols) return bic_fail(c, BIC_ERR_INVALID, "golomb_decode: shape mismatch");
  const uint64_t N = M->rows * M->cols, T = div_up_u64(N, 32);
  if (s->info.chunk_samples == 0 || s->info.nchunks != div_up_u64(s->info.nsamples, s->info.chunk_samples))
    return bic_fail(c, BIC_ERR_CORRUPT, "golomb_decode: inconsistent chunk index");
  uint32_t* S;
  const bool dense = (M->cols & 31) == 0;
  if (dense) {
    S = M->d;
    BIC_TRY(bic_mat_clear(c, M));
  } else {
    BIC_TRY(bic_scratch_reserve(c, &c->work[4], (size_t)T * 4 + 16));
    S = (uint32_t*)c->work[4].p;
    BIC_CUDA(c, cudaMemsetAsync(S, 0, (size_t)T * 4 + 16, c->stream));
  }
  BIC_CUDA(c, cudaMemsetAsync(c->d_scalars, 0, 8, c->stream));
  BIC_PROF(c, KID_GOL_DECODE);
  k_gol_decode<<<bic_grid_for(c, s->info.nchunks, 128, 16), 128, 0, c->stream>>>(
      (const uint32_t*)s->d_bytes, s->info.bitcount, (const unsigned long long*)s->d_index, s->info.nchunks,
      s->info.chunk_samples, s->info.nsamples, N, S, (unsigned long long*)c->d_scalars);
  BIC_LAUNCH_CHECK(c);
  if (!dense && M->words()) {
    BIC_PROF(c, KID_EXPAND_ROWS);
    k_expand_rows<<<bic_grid_for(c, M->words(), 256, 8), 256, 0, c->stream>>>(S, M->cols, M->wpr, M->rows, M->d);
    BIC_LAUNCH_CHECK(c);
  }
  BIC_TRY(bic_read_scalars(c, 1));
  if (c->h_scalars[0]) return bic_fail(c, BIC_ERR_CORRUPT, "golomb_decode: stream does not decode to the stated shape");
  return BIC_OK;
}

extern "C" bic_status bic_eg_encode(bic_ctx* c, const bic_mat* M, bic_stream* out) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !M || !out) return BIC_ERR_INVALID;
  const uint64_t N = M->rows * M->cols;
  const uint32_t* S; uint64_t T;
  BIC_TRY(dense_stream(c, M, &S, &T));
  BIC_CUDA(c, cudaMemsetAsync(c->d_scalars, 0xFF, 8, c->stream));
  if (T) {
    BIC_PROF(c, KID_EG_FIRST);
    k_first_one<<<bic_grid_for(c, T, 256, 8), 256, 0, c->stream>>>(S, T, (unsigned long long*)c->d_scalars);
    BIC_LAUNCH_CHECK(c);
  }
  BIC_TRY(bic_read_scalars(c, 1));
  const bool any = c->h_scalars[0] != ~0ull;
  const uint64_t bitcount = N + M->rows + (any ? 1 : 0);
  BIC_TRY(stream_reserve(c, out, bitcount, 0));
  BIC_CUDA(c, cudaMemsetAsync(out->d_bytes, 0, out->cap_bytes, c->stream));
  if (bitcount) {
    BIC_PROF(c, KID_EG_FILL);
    k_fill_ones<<<bic_grid_for(c, div_up_u64(bitcount, 32), 256, 8), 256, 0, c->stream>>>((uint32_t*)out->d_bytes, bitcount);
    BIC_LAUNCH_CHECK(c);
  }
  if (any) {
    BIC_PROF(c, KID_EG_ENCODE);
    k_eg_encode<<<bic_grid_for(c, T, 256, 8), 256, 0, c->stream>>>(S, T, M->cols, (const unsigned long long*)c->d_scalars,
                                                                  (uint32_t*)out->d_bytes);
    BIC_LAUNCH_CHECK(c);
  }
  out->info.coder = BIC_CODER_EG;
  out->info.chunk_samples = 0;
  out->info.rows = M->rows;
  out->info.cols = M->cols;
  out->info.bitcount = bitcount;
  out->info.nsamples = 0;
  out->info.nchunks = 0;
  return BIC_OK;
}

extern "C" bic_status bic_eg_decode(bic_ctx* c, const bic_stream* s, bic_mat* M) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !s || !M) return BIC_ERR_INVALID;
  if (s->info.coder != BIC_CODER_EG) return bic_fail(c, BIC_ERR_INVALID, "eg_decode: not an EG stream");
  if (M->rows != s->info.rows || M->cols != s->info.cols) return bic_fail(c, BIC_ERR_INVALID, "eg_decode: shape mismatch");
  const uint64_t N = M->rows * M->cols;
  if (s->info.bitcount != N + M->rows && s->info.bitcount != N + M->rows + 1)
    return bic_fail(c, BIC_ERR_CORRUPT, "eg_decode: bit count does not match the shape");
  BIC_CUDA(c, cudaMemsetAsync(c->d_scalars, 0xFF, 8, c->stream));
  BIC_CUDA(c, cudaMemsetAsync(c->d_scalars + 1, 0, 8, c->stream));
  if (s->info.bitcount) {
    BIC_PROF(c, KID_EG_FIRST);
    k_first_zero<<<bic_grid_for(c, div_up_u64(s->info.bitcount, 32), 256, 8), 256, 0, c->stream>>>(
        (const uint32_t*)s->d_bytes, s->info.bitcount, (unsigned long long*)c->d_scalars);
    BIC_LAUNCH_CHECK(c);
  }
  if (M->words()) {
    BIC_PROF(c, KID_EG_DECODE);
    k_eg_decode<<<bic_grid_for(c, M->words(), 256, 8), 256, 0, c->stream>>>((const uint32_t*)s->d_bytes, M->rows, M->cols,
                                                                          M->wpr, (const unsigned long long*)c->d_scalars, M->d,
                                                                          s->info.bitcount, (unsigned long long*)c->d_scalars + 1);
    BIC_LAUNCH_CHECK(c);
  }
  BIC_TRY(bic_read_scalars(c, 2));
  if (c->h_scalars[1]) return bic_fail(c, BIC_ERR_CORRUPT, "eg_decode: stream is not an EG code of the stated shape");
  return BIC_OK;
}

extern "C" bic_status bic_stream_download(bic_ctx* c, const bic_stream* s, uint8_t* bytes, uint64_t cap_bytes,
                                          uint64_t* index, uint64_t cap_index_entries) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !s) return BIC_ERR_INVALID;
  const uint64_t nb = div_up_u64(s->info.bitcount, 8);
  if (bytes) {
    if (cap_bytes < nb) return BIC_ERR_CAPACITY;
    if (nb) BIC_CUDA(c, cudaMemcpyAsync(bytes, s->d_bytes, nb, cudaMemcpyDeviceToHost, c->stream));
  }
  if (index) {
    if (cap_index_entries < s->info.nchunks) return BIC_ERR_CAPACITY;
    if (s->info.nchunks)
      BIC_CUDA(c, cudaMemcpyAsync(index, s->d_index, s->info.nchunks * 16, cudaMemcpyDeviceToHost, c->stream));
  }
  BIC_CUDA(c, bic_wait_stream(c));
  return BIC_OK;
}

extern "C" bic_status bic_stream_upload(bic_ctx* c, bic_stream* s, const bic_stream_info* info, const uint8_t* bytes,
                                        const uint64_t* index) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !s || !info) return BIC_ERR_INVALID;
  const uint64_t nb = div_up_u64(info->bitcount, 8);
  if ((nb && !bytes) || (info->nchunks && !index)) return BIC_ERR_INVALID;
  BIC_TRY(stream_reserve(c, s, info->bitcount, info->nchunks));
  BIC_CUDA(c, cudaMemsetAsync(s->d_bytes, 0, s->cap_bytes, c->stream));
  if (nb) BIC_CUDA(c, cudaMemcpyAsync(s->d_bytes, bytes, nb, cudaMemcpyHostToDevice, c->stream));
  if (info->nchunks) BIC_CUDA(c, cudaMemcpyAsync(s->d_index, index, info->nchunks * 16, cudaMemcpyHostToDevice, c->stream));
  s->info = *info;
  return BIC_OK;
}
