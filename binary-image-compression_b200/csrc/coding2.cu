// Golomb encoder, second formulation: the same bytes as coding.cu's (and as the serial coder implied by
// src/GolombCoder.cpp:13-34), reorganised around what limited the first one (profiles/r1_coder_sweep.md):
//
//   * wide tiles: a thread takes 16 consecutive words (four independent 128-bit loads in flight), a CTA 4096 words, so the fixed
//     cost per tile (two block scans, the tile's prefix) is spread over four times the input and the loads cover the latency;
//   * no scan kernels: the LAST CTA to finish a pass (an arrival counter per stream) scans the per-tile totals itself, so a stream
//     costs three launches (count, lengths, scatter) + one clear instead of seven, and nothing spins;
//   * up to three streams per launch (the D, A and E of one raster): blockIdx selects the stream, each with its own coder state;
//   * the scatter assembles every thread's codewords in registers (a thread's code is one contiguous bit range) and stores whole
//     words into the tile's staged output: plain shared-memory stores for the words the thread owns outright, an atomic OR only
//     for its first and last partial word, instead of two or three shared-memory atomics per codeword;
//   * the output buffer is sized before the bit count is known (1.25 code bits per input bit + slack); the bit count, the sample
//     count and an overflow flag stay on the device (info[]), so the encoder never waits for the host. A stream that does not fit
//     is reported through the flag and re-encoded by the exact-size path of coding.cu.
//
// Coder state as in coding.cu: before sample t (the t-th one, at position pos_t, previous one at pos_{t-1}) the reference's
// coder has samples = t, accumulatedError = pos_{t-1} + 1 - t (mod 2^32), k_t = min{k : (t << k) >= accumulatedError}.
#include "gol_common.cuh"

#define G2_THREADS 256
// words per thread: 16 for long streams (four 128-bit loads in flight per thread, 4096-word tiles), 4 for streams so short that
// 4096-word tiles would not fill the SMs for several waves (a thread's codewords are a serial chain: few fat CTAs end in a tail)
#define G2_TILE_WORDS(WPT) (G2_THREADS * (WPT))
#define G2_TILE_BITS(WPT) (G2_TILE_WORDS(WPT) * 32)
#define G2_STAGE_WORDS(WPT) (G2_TILE_WORDS(WPT) * 3 / 2)   // staged code words of one tile (1.5 x its input)
#define G2_MAXSEG 3
// Sparse tiles (at most one bit in 64 set): the count pass also writes the positions of the tile's ones as a list, and the length
// and scatter passes of such a tile work from the list -- a thread per few SAMPLES instead of a thread per 16 words, and the
// words are not read again. Decided tile by tile from the tile's own count, so a stream may mix both kinds.
#define G2_LIST_CAP(WPT) (G2_TILE_WORDS(WPT) / 2)

struct G2Seg {
  const uint32_t* S;                 // dense bit stream of the matrix, MSB first
  uint64_t T, N;                     // words, bits
  uint32_t tile0, ntiles;            // block index of the stream's first tile, tiles
  uint32_t* ones;                    // per tile: ones
  uint32_t* list;                    // per tile: G2_LIST_CAP entries, bit index inside the tile of each one (tiles with <= CAP ones); null = off
  long long* last;                   // per tile: position of the last one, -1 if none
  unsigned long long* ones_before;   // exclusive prefixes, written by the last CTA of the count pass
  long long* last_before;
  unsigned long long* bits;          // per tile: code bits of its ones
  unsigned long long* bits_before;   // exclusive prefix, written by the last CTA of the length pass
  unsigned long long* tbits;         // per thread: code bits of its 16 words
  unsigned int* done;                // [0] tiles that finished the count pass, [1] the length pass
  unsigned long long* info;          // [0] bit count [1] samples [2] code bits of the ones [3] position after the last one [4] overflow
  uint32_t* out;
  unsigned long long* index;
  unsigned long long cap_bits;
  uint32_t chunk;                    // samples per chunk-index entry, a power of two
  // where the matrix sits in the ONE global sample stream when it is a row shard of a bigger matrix (coding.cu: GolBase; all zero
  // / prev0 = -1 for a whole matrix). closing: 0 = another shard writes the run closed by the virtual one, 1 = this one does, with
  // the rank / position / offset in gb.close_* (known on the host), 2 = this one does, from info[] (the totals are still on the device)
  GolBase gb;
  const GolBase* gbp;                // non-null: the base is still being computed on the device (pipeline.cu's sharded jobs): read it from here
};
struct G2Params {
  G2Seg s[G2_MAXSEG];
  uint32_t nseg;
  uint32_t sep_scan;                 // 1: the scans over the tiles run as their own launch (k_g2_scan_tiles), not in the passes' last CTA
};

// Byte-at-a-time fast path. Once a thread's k is known to be constant (g2_k_stable) and small (<= 3: dense data, where a byte holds
// several ones), the ones of a byte are not walked one by one: the byte's FIRST one closes a run that began earlier (its codeword
// is computed from the run length as usual), and the codewords of its other ones depend on the byte alone -- they come out of a
// table as one ready-made bit pattern (gen_gol_lut.py). ncu of the per-one walk on a 34 % dense residual: 74 instructions per
// one in the scatter, 29 in the length pass, a third of the lanes idle from divergence; a byte costs about 35 / 8.
__device__ const uint2 g_gol_lut[4 * 256] = {
#include "gol_lut.inc"
};
#define G2_LUT_F(m) ((m) & 7u)
#define G2_LUT_TZ(m) (((m) >> 3) & 7u)
#define G2_LUT_NBITS(m) (((m) >> 6) & 31u)
#define G2_LUT_NONES(m) (((m) >> 11) & 15u)
// (read with __ldg straight from global memory: the 8 KB stay in L1; a per-CTA copy to shared memory cost more than the walk saved)

__device__ __forceinline__ const G2Seg& g2_segment(const G2Params& P, uint32_t* tile) {
  uint32_t i = 0;
  while (i + 1 < P.nseg && blockIdx.x >= P.s[i].tile0 + P.s[i].ntiles) ++i;
  *tile = blockIdx.x - P.s[i].tile0;
  return P.s[i];
}

// the thread's 16 words; words past the end of the stream read as zero
template <int WPT>
__device__ __forceinline__ void g2_load(const G2Seg& g, uint64_t w0, uint32_t (&v)[WPT]) {
#pragma unroll
  for (int j = 0; j < WPT / 4; ++j) {
    const uint64_t w = w0 + 4 * j;
    if (w + 4 <= g.T) {
      const uint4 q = *reinterpret_cast<const uint4*>(g.S + w);
      v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[4 * j + i] = (w + i < g.T) ? g.S[w + i] : 0u;
    }
  }
}

// ones of the 16 words and the bit index (inside the tile) of the last one, -1 if none
template <int WPT>
__device__ __forceinline__ void g2_count(const uint32_t (&v)[WPT], uint32_t* c, int* last) {
  uint32_t n = 0, lw = 0;
  int li = 0;
#pragma unroll
  for (int i = 0; i < WPT; ++i) {
    n += __popc(v[i]);
    if (v[i]) { lw = v[i]; li = i; }                 // two selects per word; the bit position is worked out once, below
  }
  *c = n;
  *last = lw ? (int)(threadIdx.x * (WPT * 32) + li * 32 + (32 - __ffs(lw))) : -1;
}

// the count pass's own scan: exclusive prefix of the ones (list offsets) and only the MAXIMUM of the last-one positions
__device__ __forceinline__ void g2_scan_count(uint32_t c, int last, uint32_t* ex_c, uint32_t* tot_c, int* tot_last, int* s_w) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  const int wmax = __reduce_max_sync(0xffffffffu, last);
  if (lane == 31) { s_w[wib] = (int)inc; s_w[8 + wib] = wmax; }
  __syncthreads();
  uint32_t base = 0, tot = 0;
  int totm = -1;
#pragma unroll
  for (int w = 0; w < G2_THREADS / 32; ++w) {
    const uint32_t x = (uint32_t)s_w[w];
    if (w < wib) base += x;
    tot += x;
    totm = max(totm, s_w[8 + w]);
  }
  *ex_c = base + inc - c;
  *tot_c = tot;
  *tot_last = totm;
}

// exclusive prefix sum and exclusive prefix max over the CTA's 256 threads (s_w: 16 words of shared memory)
__device__ __forceinline__ void g2_scan(uint32_t c, int last, uint32_t* ex_c, int* ex_last, uint32_t* tot_c, int* tot_last, int* s_w) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t inc = c;
  int incm = last;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
    const int ym = __shfl_up_sync(0xffffffffu, incm, o);
    if (lane >= o) { inc += y; incm = max(incm, ym); }
  }
  int exm = __shfl_up_sync(0xffffffffu, incm, 1);
  if (lane == 0) exm = -1;
  __syncthreads();                     // s_w may still be read from the previous use
  if (lane == 31) { s_w[wib] = (int)inc; s_w[8 + wib] = incm; }
  __syncthreads();
  uint32_t base = 0, tot = 0;
  int basem = -1, totm = -1;
#pragma unroll
  for (int w = 0; w < G2_THREADS / 32; ++w) {
    const uint32_t x = (uint32_t)s_w[w];
    const int xm = s_w[8 + w];
    if (w < wib) { base += x; basem = max(basem, xm); }
    tot += x;
    totm = max(totm, xm);
  }
  *ex_c = base + inc - c;
  *ex_last = max(basem, exm);
  *tot_c = tot;
  *tot_last = totm;
}

__device__ __forceinline__ unsigned long long g2_excl_sum_u64(unsigned long long v, unsigned long long* total, unsigned long long* s_w /* 8 */) {
  return block_excl_scan_u64(v, total, s_w);
}

// k is the same for every sample of a stretch of `span` input bits that starts with sample rank t and consumed bits `consumed`
// (coding.cu: golomb_k_stable, with the thread's stretch as a parameter)
__device__ __forceinline__ bool g2_k_stable(uint64_t t, uint64_t consumed, uint32_t span, uint32_t* kout) {
  if (t == 0 || t + span >= (1ull << 31)) return false;
  const uint64_t acc = consumed - t;
  if (acc + span >= (1ull << 31)) return false;
  const uint32_t k = golomb_k(t, consumed);
  if (((t + span - 1) << k) >= (1ull << 32)) return false;
  if ((t << k) < acc + span) return false;
  if (k > 0 && ((t + span - 1) << (k - 1)) >= acc) return false;
  *kout = k;
  return true;
}

// ------------------------------------------------------------------ pass 1: ones and last one per tile; the last CTA scans the tiles
template <int WPT>
__global__ void __launch_bounds__(G2_THREADS) k_g2_count(G2Params P) {
  __shared__ int s_w[16];
  __shared__ unsigned long long s_a[8];
  __shared__ long long s_b[8];
  __shared__ int s_last_cta;
  uint32_t tile;
  const G2Seg& g = g2_segment(P, &tile);
  uint32_t v[WPT];
  g2_load(g, (uint64_t)tile * G2_TILE_WORDS(WPT) + threadIdx.x * WPT, v);
  uint32_t c, ex_c, tot_c;
  int last, tot_last;
  g2_count(v, &c, &last);
  g2_scan_count(c, last, &ex_c, &tot_c, &tot_last, s_w);
  if (g.list && c && tot_c <= G2_LIST_CAP(WPT)) {                       // a sparse tile: my ones go into the tile's list
    uint32_t* L = g.list + (uint64_t)tile * G2_LIST_CAP(WPT) + ex_c;
#pragma unroll
    for (int i = 0; i < WPT; ++i) {
      uint32_t b = v[i];
      while (b) {
        const int p = __clz(b);
        b &= ~(0x80000000u >> p);
        *L++ = threadIdx.x * (WPT * 32) + i * 32 + p;
      }
    }
  }
  if (threadIdx.x == 0) {
    g.ones[tile] = tot_c;
    g.last[tile] = tot_last >= 0 ? (long long)tile * G2_TILE_BITS(WPT) + tot_last : -1;
    if (!P.sep_scan) {
      __threadfence();
      s_last_cta = (atomicAdd(g.done, 1u) == g.ntiles - 1);
    }
  }
  if (P.sep_scan) return;
  __syncthreads();
  if (!s_last_cta) return;
  __threadfence();
  // exclusive scans over the stream's tiles: a contiguous run of tiles per thread, one block scan over the run totals. The loads
  // of a run are issued eight at a time (restrict + unroll): this CTA is the kernel's tail and a dependent L2 round trip per tile
  // would be most of a sparse stream's count pass
  const uint32_t* __restrict__ ones = g.ones;
  const long long* __restrict__ lastp = g.last;
  unsigned long long* __restrict__ ob = g.ones_before;
  long long* __restrict__ lbp = g.last_before;
  const uint64_t per = div_up_u64(g.ntiles, G2_THREADS);
  const uint64_t t0 = threadIdx.x * per, t1 = (t0 + per < g.ntiles) ? t0 + per : g.ntiles;
  unsigned long long sum = 0;
  long long mx = -1;
#pragma unroll 8
  for (uint64_t i = t0; i < t1; ++i) { sum += __ldcg(ones + i); const long long l = __ldcg(lastp + i); mx = mx > l ? mx : l; }
  unsigned long long run = block_excl_scan_u64(sum, nullptr, s_a);
  long long runm = block_excl_scan_max(mx, nullptr, s_b);
#pragma unroll 8
  for (uint64_t i = t0; i < t1; ++i) {
    ob[i] = run;
    lbp[i] = runm;
    run += __ldcg(ones + i);
    const long long l = __ldcg(lastp + i);
    runm = runm > l ? runm : l;
  }
}

// ------------------------------------------------------------------ sparse tiles: a thread's samples from the tile's list
// Thread j of the CTA owns samples [i0, i1) of the tile (consecutive, so its codewords are one contiguous bit range, like a
// thread's 16 words in the dense route). Returns their code bits; *t / *prev: global rank of sample i0 / global position of the
// one before it (the state the first codeword is computed from).
template <int WPT>
__device__ __forceinline__ unsigned long long g2_list_span(const G2Seg& g, const GolBase& gb, uint32_t tile, uint32_t nones, long long lb,
                                                           uint32_t* i0o, uint32_t* i1o, unsigned long long* t0o, long long* prev0o) {
  const uint32_t q = (nones + G2_THREADS - 1) / G2_THREADS;
  const uint32_t i0 = min(threadIdx.x * q, nones), i1 = min(i0 + q, nones);
  const uint32_t* L = g.list + (uint64_t)tile * G2_LIST_CAP(WPT);
  const long long tileb = (long long)tile * G2_TILE_BITS(WPT) + gb.pos0;
  long long prev = i0 ? tileb + (long long)__ldcg(L + i0 - 1) : (lb >= 0 ? lb + gb.pos0 : gb.prev0);
  unsigned long long t = gb.t0 + g.ones_before[tile] + i0;
  *i0o = i0; *i1o = i1; *t0o = t; *prev0o = prev;
  unsigned long long bits = 0;
  for (uint32_t i = i0; i < i1; ++i) {
    const long long pos = tileb + (long long)__ldcg(L + i);
    const unsigned long long x = (unsigned long long)(pos - prev - 1);
    const uint32_t k = golomb_k(t, (unsigned long long)(prev + 1));
    bits += k + (x >> k) + 1;
    prev = pos;
    ++t;
  }
  return bits;
}

// ------------------------------------------------------------------ pass 2: code bits per thread and tile; the last CTA scans and totals
template <int WPT>
__global__ void __launch_bounds__(G2_THREADS) k_g2_lengths(G2Params P) {
  __shared__ int s_w[16];
  __shared__ unsigned long long s_a[8];
  __shared__ int s_last_cta;
  uint32_t tile;
  const G2Seg& g = g2_segment(P, &tile);
  const GolBase gb = g.gbp ? *g.gbp : g.gb;
  const long long lb = g.last_before[tile];
  unsigned long long mybits = 0;
  const uint32_t nones = g.list ? g.ones[tile] : 0xffffffffu;
  if (nones <= G2_LIST_CAP(WPT)) {                                                     // sparse tile (uniform over the CTA): from the list
    uint32_t i0, i1;
    unsigned long long t;
    long long prev;
    mybits = g2_list_span<WPT>(g, gb, tile, nones, lb, &i0, &i1, &t, &prev);
  } else {
  uint32_t v[WPT];
  const uint64_t w0 = (uint64_t)tile * G2_TILE_WORDS(WPT) + threadIdx.x * WPT;
  g2_load(g, w0, v);
  uint32_t c, ex_c, tot_c;
  int last, ex_last, tot_last;
  g2_count(v, &c, &last);
  g2_scan(c, last, &ex_c, &ex_last, &tot_c, &tot_last, s_w);
  unsigned long long t = gb.t0 + g.ones_before[tile] + ex_c;                          // global rank of my first sample
  const long long pvl = ex_last >= 0 ? (long long)tile * G2_TILE_BITS(WPT) + ex_last : lb;  // the one before my stretch inside this matrix, -1 if none
  long long prev = pvl >= 0 ? pvl + gb.pos0 : gb.prev0;                                // ... as a GLOBAL position
  const long long tb = (long long)(w0 * 32) + gb.pos0;                                 // global position of my first bit
  if (c) {
    // my first one closes a run that began before my stretch: full-width arithmetic, and the coder state it leaves decides how
    // the rest of the stretch is walked (the three loops below are each uniform, so a warp only diverges where lanes differ in mode)
    uint32_t fw = 0, fi = 0;
#pragma unroll
    for (int i = WPT - 1; i >= 0; --i) if (v[i]) { fw = v[i]; fi = (uint32_t)i; }
    const uint32_t p0 = (uint32_t)__clz(fw);
    uint32_t lpv = fi * 32 + p0;
#pragma unroll
    for (int i = 0; i < WPT; ++i) if ((uint32_t)i == fi) v[i] &= ~(0x80000000u >> p0);
    {
      const long long pos = tb + lpv;
      const unsigned long long x = (unsigned long long)(pos - prev - 1);
      const uint32_t k = golomb_k(t, (unsigned long long)(prev + 1));
      mybits += k + (x >> k) + 1;
      prev = pos;
      ++t;
    }
    uint32_t kc = 0, fastbits = 0;
    const bool fast = g2_k_stable(t, (unsigned long long)(prev + 1), WPT * 32, &kc);
    if (fast && kc <= 3) {
      const uint2* lutk = g_gol_lut + kc * 256;
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        uint32_t wv = v[i];
        for (uint32_t bb = i * 32; wv; bb += 8, wv <<= 8) {        // byte by byte, MSB first: the whole byte from the table
          const uint32_t B = wv >> 24;
          if (!B) continue;
          const uint32_t meta = __ldg(&lutk[B].y);
          fastbits += kc + 1 + ((bb + G2_LUT_F(meta) - lpv - 1) >> kc) + G2_LUT_NBITS(meta);
          lpv = bb + 7 - G2_LUT_TZ(meta);
        }
      }
    } else if (fast) {
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        uint32_t b = v[i];
        while (b) {
          const int p = __clz(b);
          b &= ~(0x80000000u >> p);
          const uint32_t lp = (uint32_t)(i * 32 + p);
          fastbits += kc + 1 + ((lp - lpv - 1) >> kc);
          lpv = lp;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        uint32_t b = v[i];
        while (b) {
          const int p = __clz(b);
          b &= ~(0x80000000u >> p);
          const long long pos = tb + i * 32 + p;
          const unsigned long long x = (unsigned long long)(pos - prev - 1);
          const uint32_t k = golomb_k(t, (unsigned long long)(prev + 1));
          mybits += k + (x >> k) + 1;
          prev = pos;
          ++t;
        }
      }
    }
    mybits += fastbits;
  }
  g.tbits[(uint64_t)tile * G2_THREADS + threadIdx.x] = mybits;
  }
  unsigned long long tot;
  g2_excl_sum_u64(mybits, &tot, s_a);
  if (threadIdx.x == 0) {
    g.bits[tile] = tot;
    if (!P.sep_scan) {
      __threadfence();
      s_last_cta = (atomicAdd(g.done + 1, 1u) == g.ntiles - 1);
    }
  }
  if (P.sep_scan) return;
  __syncthreads();
  if (!s_last_cta) return;
  __threadfence();
  const uint64_t per = div_up_u64(g.ntiles, G2_THREADS);
  const uint64_t t0 = threadIdx.x * per, t1 = (t0 + per < g.ntiles) ? t0 + per : g.ntiles;
  const unsigned long long* __restrict__ bitsp = g.bits;
  unsigned long long* __restrict__ bbp = g.bits_before;
  unsigned long long sum = 0;
#pragma unroll 8
  for (uint64_t i = t0; i < t1; ++i) sum += __ldcg(bitsp + i);
  unsigned long long carry;
  unsigned long long run = block_excl_scan_u64(sum, &carry, s_a);
#pragma unroll 8
  for (uint64_t i = t0; i < t1; ++i) { bbp[i] = run; run += __ldcg(bitsp + i); }
  if (threadIdx.x == 0) {
    const uint64_t lt = g.ntiles - 1;
    const unsigned long long ones = g.ones_before[lt] + g.ones[lt];
    const long long lastg = g.last_before[lt] > g.last[lt] ? g.last_before[lt] : g.last[lt];
    const unsigned long long consumed = (unsigned long long)(lastg + 1);
    const unsigned long long x = g.N - consumed;                                        // the run closed by the virtual one
    const uint32_t k = golomb_k(ones, consumed);
    const unsigned long long total = carry + k + (x >> k) + 1;
    if (g.gbp) {            // a shard: the stream's totals are the business of k_g2_shard_base2; this rank's code bits are all it needs
      g.info[2] = carry;
    } else {
      g.info[0] = total;
      g.info[1] = ones + 1;
      g.info[2] = carry;
      g.info[3] = consumed;
      g.info[4] = (total + 64 > g.cap_bits) ? 1ull : 0ull;
    }
  }
}

// ------------------------------------------------------------------ the scans over the tiles as their own launch (long streams)
// One CTA of 1024 threads per stream. PASS 0 (after the count pass): ones / last -> ones_before / last_before. PASS 1 (after the
// length pass): bits -> bits_before and the stream's totals in info[]. In the passes' last CTA the same scans are a serial tail
// of 256 threads (ncu, 2^31 bits at 0.1 % ones: the SMs idle for 44 % of the count pass and 35 % of the length pass behind it).
#define G2_SCAN_THREADS 1024
template <typename T, bool MAX>
__device__ __forceinline__ T g2_block_excl_1024(T v, T ident, T* total, T* s_w /* 32 */) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  T inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const T y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc = MAX ? (inc > y ? inc : y) : inc + y;
  }
  T excl = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) excl = ident;
  __syncthreads();
  if (lane == 31) s_w[wib] = inc;
  __syncthreads();
  T base = ident, tot = ident;
  for (int w = 0; w < G2_SCAN_THREADS / 32; ++w) {
    const T x = s_w[w];
    if (w < wib) base = MAX ? (base > x ? base : x) : base + x;
    tot = MAX ? (tot > x ? tot : x) : tot + x;
  }
  if (total) *total = tot;
  return MAX ? (base > excl ? base : excl) : base + excl;
}

template <int PASS>
__global__ void __launch_bounds__(G2_SCAN_THREADS) k_g2_scan_tiles(G2Params P) {
  __shared__ unsigned long long s_a[32];
  __shared__ long long s_b[32];
  const G2Seg& g = P.s[blockIdx.x];
  const uint64_t per = div_up_u64(g.ntiles, G2_SCAN_THREADS);
  const uint64_t t0 = min((uint64_t)threadIdx.x * per, (uint64_t)g.ntiles), t1 = min(t0 + per, (uint64_t)g.ntiles);
  if (PASS == 0) {
    const uint32_t* __restrict__ ones = g.ones;
    const long long* __restrict__ lastp = g.last;
    unsigned long long* __restrict__ ob = g.ones_before;
    long long* __restrict__ lbp = g.last_before;
    unsigned long long sum = 0;
    long long mx = -1;
#pragma unroll 4
    for (uint64_t i = t0; i < t1; ++i) { sum += ones[i]; const long long l = lastp[i]; mx = mx > l ? mx : l; }
    unsigned long long run = g2_block_excl_1024<unsigned long long, false>(sum, 0ull, nullptr, s_a);
    long long runm = g2_block_excl_1024<long long, true>(mx, -1ll, nullptr, s_b);
#pragma unroll 4
    for (uint64_t i = t0; i < t1; ++i) {
      ob[i] = run;
      lbp[i] = runm;
      run += ones[i];
      const long long l = lastp[i];
      runm = runm > l ? runm : l;
    }
  } else {
    const unsigned long long* __restrict__ bitsp = g.bits;
    unsigned long long* __restrict__ bbp = g.bits_before;
    unsigned long long sum = 0;
#pragma unroll 4
    for (uint64_t i = t0; i < t1; ++i) sum += bitsp[i];
    unsigned long long carry;
    unsigned long long run = g2_block_excl_1024<unsigned long long, false>(sum, 0ull, &carry, s_a);
#pragma unroll 4
    for (uint64_t i = t0; i < t1; ++i) { bbp[i] = run; run += bitsp[i]; }
    if (threadIdx.x == 0) {            // the stream's totals, as in the length pass's last CTA
      const uint64_t lt = g.ntiles - 1;
      const unsigned long long ones = g.ones_before[lt] + g.ones[lt];
      const long long lastg = g.last_before[lt] > g.last[lt] ? g.last_before[lt] : g.last[lt];
      const unsigned long long consumed = (unsigned long long)(lastg + 1);
      const unsigned long long x = g.N - consumed;
      const uint32_t k = golomb_k(ones, consumed);
      const unsigned long long total = carry + k + (x >> k) + 1;
      g.info[0] = total;
      g.info[1] = ones + 1;
      g.info[2] = carry;
      g.info[3] = consumed;
      g.info[4] = (total + 64 > g.cap_bits) ? 1ull : 0ull;
    }
  }
}

// ------------------------------------------------------------------ clear: the words the codes will occupy
__global__ void __launch_bounds__(256) k_g2_clear(G2Params P) {
  for (uint32_t si = 0; si < P.nseg; ++si) {
    const G2Seg& g = P.s[si];
    if (g.info[4]) continue;
    const unsigned long long nw = ((g.info[0] + 31) >> 5) + 4;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < nw; i += (unsigned long long)gridDim.x * blockDim.x)
      g.out[i] = 0u;
  }
}

// ------------------------------------------------------------------ pass 3: scatter
// A thread's codewords are the contiguous bit range [o, o + mybits) of the stream. They are built in a register (cur: the word
// being filled, `fill` bits used from the top) and stored word by word into the tile's staged range.
struct G2Out {
  uint32_t* s_out;
  uint32_t wp, fill, cur;
  uint32_t first_wp, first_val;  // my first word when my range starts inside it: shared with the previous thread, so its value is
  bool first_pending;            // kept in a register and OR-ed in at the end (finish) instead of being stored
  uint32_t limit;                // words of the staged range (debug build: every store is checked against it)
  // Everything below is branch-free on purpose: lanes complete their words at different times, and a branch around the store
  // would be taken by a few lanes on almost every codeword (ncu: the store block ran with 2.5 of 32 lanes active).
  __device__ __forceinline__ void store_if(bool doit, uint32_t val) {
    BIC_DCHECK(!doit || !val || wp < limit);
    const bool cap = doit && first_pending;
    first_val = cap ? val : first_val;
    if (doit && !first_pending) s_out[wp] = val;                       // one predicated store: the word is mine alone
    first_pending = first_pending && !doit;
    wp += doit ? 1u : 0u;
  }
  __device__ __forceinline__ void put(uint32_t cw, uint32_t L) {        // L <= 32 bits, right aligned in cw
    const unsigned long long a = ((unsigned long long)cur << 32) | ((unsigned long long)cw << (64 - fill - L));
    fill += L;
    const bool full = fill >= 32;
    store_if(full, (uint32_t)(a >> 32));
    cur = full ? (uint32_t)a : (uint32_t)(a >> 32);
    fill -= full ? 32u : 0u;
  }
  __device__ __forceinline__ void zeros(unsigned long long u) {         // u zero bits (the buffer is zeroed: just move on)
    const unsigned long long total = fill + u;
    if (total >= 32) {
      store_if(true, cur);
      wp += (uint32_t)(total >> 5) - 1;
      cur = 0;
      fill = (uint32_t)(total & 31);
    } else {
      fill = (uint32_t)total;
    }
  }
  __device__ __forceinline__ void finish() {
    if (first_pending) {            // never completed a word: everything I wrote is in `cur`, inside my (shared) first word
      BIC_DCHECK(!cur || wp < limit);
      if (cur) atomicOr(&s_out[wp], cur);
      return;
    }
    if (first_val) atomicOr(&s_out[first_wp], first_val);
    BIC_DCHECK(!(fill && cur) || wp < limit);
    if (fill && cur) atomicOr(&s_out[wp], cur);                         // my last partial word is shared with the next thread
  }
};

// a tile coded from its words (the dense route): thread j owns words [16 j, 16 j + 16) of the tile
template <int WPT>
__device__ __forceinline__ void g2_scatter_word_tile(const G2Seg& g, const GolBase& gb, uint32_t tile, int* s_w, unsigned long long* s_a, uint32_t* s_out) {
  uint32_t v[WPT];
  const uint64_t w0 = (uint64_t)tile * G2_TILE_WORDS(WPT) + threadIdx.x * WPT;
  g2_load(g, w0, v);
  uint32_t c, ex_c, tot_c;
  int last, ex_last, tot_last;
  g2_count(v, &c, &last);
  g2_scan(c, last, &ex_c, &ex_last, &tot_c, &tot_last, s_w);
  unsigned long long t = gb.t0 + g.ones_before[tile] + ex_c;
  const long long lb = g.last_before[tile];
  const long long pvl = ex_last >= 0 ? (long long)tile * G2_TILE_BITS(WPT) + ex_last : lb;
  long long prev = pvl >= 0 ? pvl + gb.pos0 : gb.prev0;
  const long long tb = (long long)(w0 * 32) + gb.pos0;
  const unsigned long long mybits = g.tbits[(uint64_t)tile * G2_THREADS + threadIdx.x];
  unsigned long long tot;
  const unsigned long long ex = g2_excl_sum_u64(mybits, &tot, s_a);
  const unsigned long long o0 = g.bits_before[tile] + gb.out0;         // bit offset inside this shard's own buffer
  const unsigned long long base = o0 & ~31ull;                           // buffer bit position of s_out[0]
  const unsigned long long goff = gb.code0 - gb.out0;                // buffer bit offset -> global code bit offset (chunk index)
  const unsigned long long span_words = ((o0 - base) + tot + 31) >> 5;
  const bool staged = span_words <= G2_STAGE_WORDS(WPT);                      // uniform over the CTA
  const unsigned long long cmask = (unsigned long long)g.chunk - 1;
  const int clog = 31 - __clz(g.chunk);
  if (staged) {
    for (unsigned i = threadIdx.x; i < (unsigned)span_words; i += blockDim.x) s_out[i] = 0;
    __syncthreads();
    if (c) {
      const uint32_t ob = (uint32_t)(o0 + ex - base);
      G2Out w;
      w.s_out = s_out; w.wp = ob >> 5; w.fill = ob & 31; w.cur = 0;
      w.first_wp = w.wp; w.first_val = 0; w.first_pending = (ob & 31) != 0;   // a range that starts on a word boundary shares nothing there
      w.limit = (uint32_t)span_words;
      const uint32_t cmask32 = g.chunk - 1;
      // my first one (see the length pass): codeword from the full-width run length, then the mode of the rest
      uint32_t fw = 0, fi = 0;
#pragma unroll
      for (int i = WPT - 1; i >= 0; --i) if (v[i]) { fw = v[i]; fi = (uint32_t)i; }
      const uint32_t p0 = (uint32_t)__clz(fw);
      uint32_t lpv = fi * 32 + p0;
#pragma unroll
      for (int i = 0; i < WPT; ++i) if ((uint32_t)i == fi) v[i] &= ~(0x80000000u >> p0);
      {
        const uint32_t k = golomb_k(t, (unsigned long long)(prev + 1));
        const unsigned long long x = (unsigned long long)(tb + lpv - prev - 1);
        if ((t & cmask) == 0) {                                          // chunk index: where this sample's codeword and run start
          const unsigned long long slot = (t >> clog) - gb.chunk0;
          BIC_DCHECK(slot <= (g.N >> clog) + 1);
          g.index[2 * slot] = base + (unsigned long long)w.wp * 32 + w.fill + goff;
          g.index[2 * slot + 1] = (unsigned long long)(prev + 1);
        }
        const unsigned long long u = x >> k;
        const uint32_t rem = (uint32_t)(x & ((1ull << k) - 1));
        if (k + u + 1 <= 32) {
          w.put((rem << (uint32_t)(u + 1)) | 1u, k + (uint32_t)u + 1);    // k remainder bits, u zeros, a one
        } else {
          if (k) w.put(rem, k);
          w.zeros(u);
          w.put(1u, 1);
        }
        ++t;
        prev = tb + lpv;
      }
      uint32_t kc = 0;
      const bool fast = g2_k_stable(t, (unsigned long long)(prev + 1), WPT * 32, &kc);
      const uint32_t kmask = (1u << kc) - 1u;
      uint32_t tl = (uint32_t)t;                                           // low bits of the running sample rank (fast modes)
      if (fast && kc > 3) {
        // constant k, sparse data (at most a one or two per word): one sample at a time, word by word
#pragma unroll
        for (int i = 0; i < WPT; ++i) {
          uint32_t b = v[i];
          while (b) {
            const int p = __clz(b);
            b &= ~(0x80000000u >> p);
            const uint32_t lp = (uint32_t)(i * 32 + p);
            const uint32_t x = lp - lpv - 1, u = x >> kc, rem = x & kmask;
            if ((tl & cmask32) == 0) {
              const unsigned long long slot = ((t + (tl - (uint32_t)t)) >> clog) - gb.chunk0;
              BIC_DCHECK(slot <= (g.N >> clog) + 1);
              g.index[2 * slot] = base + (unsigned long long)w.wp * 32 + w.fill + goff;
              g.index[2 * slot + 1] = (unsigned long long)(tb + lpv + 1);
            }
            if (kc + u + 1 <= 32) {
              w.put((rem << (u + 1)) | 1u, kc + u + 1);
            } else {
              w.put(rem, kc);
              w.zeros(u);
              w.put(1u, 1);
            }
            lpv = lp;
            ++tl;
          }
        }
      } else if (fast) {
        const uint2* lutk = g_gol_lut + kc * 256;
#pragma unroll
        for (int i = 0; i < WPT; ++i) {
          uint32_t wv = v[i];
          for (uint32_t bb = i * 32; wv; bb += 8, wv <<= 8) {            // byte by byte, MSB first
            const uint32_t B = wv >> 24;
            if (!B) continue;
            {
              // constant small k: the byte's first one closes the run in progress, its other ones come from the table as one bit
              // pattern -- unless one of the byte's samples starts a chunk (its code position goes into the index): then the byte
              // is walked one by one below
              const uint2 en = __ldg(&lutk[B]);
              const uint32_t no = G2_LUT_NONES(en.y);
              if (((tl + no - 1) >> clog) == ((tl - 1) >> clog)) {
                const uint32_t x = bb + G2_LUT_F(en.y) - lpv - 1, u = x >> kc, rem = x & kmask;
                if (kc + u + 1 <= 32) {
                  w.put((rem << (u + 1)) | 1u, kc + u + 1);
                } else {
                  if (kc) w.put(rem, kc);
                  w.zeros(u);
                  w.put(1u, 1);
                }
                const uint32_t nb = G2_LUT_NBITS(en.y);
                if (nb) w.put(en.x, nb);
                lpv = bb + 7 - G2_LUT_TZ(en.y);
                tl += no;
                continue;
              }
            }
            uint32_t b = B << 24;
            while (b) {                                                  // k = kc, one sample at a time
              const int p = __clz(b);
              b &= ~(0x80000000u >> p);
              const uint32_t lp = bb + p;
              const uint32_t x = lp - lpv - 1, u = x >> kc, rem = x & kmask;
              if ((tl & cmask32) == 0) {
                const unsigned long long slot = ((t + (tl - (uint32_t)t)) >> clog) - gb.chunk0;
                BIC_DCHECK(slot <= (g.N >> clog) + 1);
                g.index[2 * slot] = base + (unsigned long long)w.wp * 32 + w.fill + goff;
                g.index[2 * slot + 1] = (unsigned long long)(tb + lpv + 1);
              }
              if (kc + u + 1 <= 32) {
                w.put((rem << (u + 1)) | 1u, kc + u + 1);
              } else {
                if (kc) w.put(rem, kc);
                w.zeros(u);
                w.put(1u, 1);
              }
              lpv = lp;
              ++tl;
            }
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < WPT; ++i) {
          uint32_t b = v[i];
          while (b) {                                                    // k re-derived for every sample
            const int p = __clz(b);
            b &= ~(0x80000000u >> p);
            const long long pos = tb + i * 32 + p;
            const uint32_t k = golomb_k(t, (unsigned long long)(prev + 1));
            const unsigned long long x = (unsigned long long)(pos - prev - 1);
            if ((t & cmask) == 0) {
              const unsigned long long slot = (t >> clog) - gb.chunk0;
              BIC_DCHECK(slot <= (g.N >> clog) + 1);
              g.index[2 * slot] = base + (unsigned long long)w.wp * 32 + w.fill + goff;
              g.index[2 * slot + 1] = (unsigned long long)(prev + 1);
            }
            const unsigned long long u = x >> k;
            const uint32_t rem = (uint32_t)(x & ((1ull << k) - 1));
            if (k + u + 1 <= 32) {
              w.put((rem << (uint32_t)(u + 1)) | 1u, k + (uint32_t)u + 1);
            } else {
              if (k) w.put(rem, k);
              w.zeros(u);
              w.put(1u, 1);
            }
            ++t;
            prev = pos;
          }
        }
      }
      w.finish();
      BIC_DCHECK((unsigned long long)w.wp * 32 + w.fill == (o0 + ex - base) + mybits);   // I wrote exactly the bits the length pass counted
    }
    __syncthreads();
    uint32_t* gout = g.out + (base >> 5);
    for (unsigned i = threadIdx.x; i < (unsigned)span_words; i += blockDim.x) {
      const uint32_t wv = s_out[i];
      if (!wv) continue;
      if (i == 0 || i == (unsigned)span_words - 1) atomicOr(gout + i, bswap32(wv));   // shared with the neighbouring tiles
      else gout[i] = bswap32(wv);
    }
  } else if (c) {
    // a tile whose code is far longer than its input (long unary parts): codewords go straight to global memory
    unsigned long long o = o0 + ex;
#pragma unroll
    for (int i = 0; i < WPT; ++i) {
      uint32_t b = v[i];
      while (b) {
        const int p = __clz(b);
        b &= ~(0x80000000u >> p);
        const long long pos = tb + i * 32 + p;
        const unsigned long long x = (unsigned long long)(pos - prev - 1);
        const uint32_t k = golomb_k(t, (unsigned long long)(prev + 1));
        if ((t & cmask) == 0) { g.index[2 * ((t >> clog) - gb.chunk0)] = o + goff; g.index[2 * ((t >> clog) - gb.chunk0) + 1] = (unsigned long long)(prev + 1); }
        put_bits(g.out, o, (uint32_t)(x & ((1ull << k) - 1)), k);
        const unsigned long long stop = o + k + (x >> k);
        put_one(g.out, stop);
        o = stop + 1;
        prev = pos;
        ++t;
      }
    }
  }
}

// a sparse tile coded from the list of its ones: thread j owns samples [i0, i1) (g2_list_span), k re-derived for every sample
template <int WPT>
__device__ __forceinline__ void g2_scatter_list_tile(const G2Seg& g, const GolBase& gb, uint32_t tile, uint32_t nones, unsigned long long* s_a, uint32_t* s_out) {
  const long long lb = g.last_before[tile];
  uint32_t li0, li1;
  unsigned long long t;
  long long prev;
  const unsigned long long mybits = g2_list_span<WPT>(g, gb, tile, nones, lb, &li0, &li1, &t, &prev);
  const uint32_t* const lst = g.list + (uint64_t)tile * G2_LIST_CAP(WPT);
  const long long tileb = (long long)tile * G2_TILE_BITS(WPT) + gb.pos0;
  unsigned long long tot;
  const unsigned long long ex = g2_excl_sum_u64(mybits, &tot, s_a);
  const unsigned long long o0 = g.bits_before[tile] + gb.out0;
  const unsigned long long base = o0 & ~31ull;
  const unsigned long long goff = gb.code0 - gb.out0;
  const unsigned long long span_words = ((o0 - base) + tot + 31) >> 5;
  const bool staged = span_words <= G2_STAGE_WORDS(WPT);
  const unsigned long long cmask = (unsigned long long)g.chunk - 1;
  const int clog = 31 - __clz(g.chunk);
  if (staged) {
    for (unsigned i = threadIdx.x; i < (unsigned)span_words; i += blockDim.x) s_out[i] = 0;
    __syncthreads();
    if (li1 > li0) {
      const uint32_t ob = (uint32_t)(o0 + ex - base);
      G2Out w;
      w.s_out = s_out; w.wp = ob >> 5; w.fill = ob & 31; w.cur = 0;
      w.first_wp = w.wp; w.first_val = 0; w.first_pending = (ob & 31) != 0;
      w.limit = (uint32_t)span_words;
      for (uint32_t i = li0; i < li1; ++i) {
        const long long pos = tileb + (long long)__ldcg(lst + i);
        const uint32_t k = golomb_k(t, (unsigned long long)(prev + 1));
        const unsigned long long x = (unsigned long long)(pos - prev - 1);
        if ((t & cmask) == 0) {                                          // chunk index: where this sample's codeword and run start
          const unsigned long long slot = (t >> clog) - gb.chunk0;
          BIC_DCHECK(slot <= (g.N >> clog) + 1);
          g.index[2 * slot] = base + (unsigned long long)w.wp * 32 + w.fill + goff;
          g.index[2 * slot + 1] = (unsigned long long)(prev + 1);
        }
        const unsigned long long u = x >> k;
        const uint32_t rem = (uint32_t)(x & ((1ull << k) - 1));
        if (k + u + 1 <= 32) {
          w.put((rem << (uint32_t)(u + 1)) | 1u, k + (uint32_t)u + 1);
        } else {
          if (k) w.put(rem, k);
          w.zeros(u);
          w.put(1u, 1);
        }
        ++t;
        prev = pos;
      }
      w.finish();
      BIC_DCHECK((unsigned long long)w.wp * 32 + w.fill == (o0 + ex - base) + mybits);
    }
    __syncthreads();
    uint32_t* gout = g.out + (base >> 5);
    for (unsigned i = threadIdx.x; i < (unsigned)span_words; i += blockDim.x) {
      const uint32_t wv = s_out[i];
      if (!wv) continue;
      if (i == 0 || i == (unsigned)span_words - 1) atomicOr(gout + i, bswap32(wv));
      else gout[i] = bswap32(wv);
    }
  } else {
    unsigned long long o = o0 + ex;                                      // codes far longer than the input: straight to global memory
    for (uint32_t i = li0; i < li1; ++i) {
      const long long pos = tileb + (long long)__ldcg(lst + i);
      const unsigned long long x = (unsigned long long)(pos - prev - 1);
      const uint32_t k = golomb_k(t, (unsigned long long)(prev + 1));
      if ((t & cmask) == 0) { g.index[2 * ((t >> clog) - gb.chunk0)] = o + goff; g.index[2 * ((t >> clog) - gb.chunk0) + 1] = (unsigned long long)(prev + 1); }
      put_bits(g.out, o, (uint32_t)(x & ((1ull << k) - 1)), k);
      const unsigned long long stop = o + k + (x >> k);
      put_one(g.out, stop);
      o = stop + 1;
      prev = pos;
      ++t;
    }
  }
}

// the run closed by the virtual one after the matrix's last bit (one thread of the stream's first tile)
__device__ __forceinline__ void g2_scatter_closing(const G2Seg& g, const GolBase& gb, uint32_t tile) {
  const unsigned long long goff = gb.code0 - gb.out0;
  const unsigned long long cmask = (unsigned long long)g.chunk - 1;
  const int clog = 31 - __clz(g.chunk);
  if (gb.closing && tile == 0 && threadIdx.x == 0) {                   // the run closed by the virtual one
    const bool dyn = gb.closing == 2;
    const unsigned long long tt = dyn ? g.info[1] - 1 : gb.close_t, consumed = dyn ? g.info[3] : gb.close_consumed;
    unsigned long long oo = (dyn ? g.info[2] : gb.close_off) + gb.out0;
    const unsigned long long x = (dyn ? g.N : gb.close_n) - consumed;
    const uint32_t k = golomb_k(tt, consumed);
    if ((tt & cmask) == 0) { g.index[2 * ((tt >> clog) - gb.chunk0)] = oo + goff; g.index[2 * ((tt >> clog) - gb.chunk0) + 1] = consumed; }
    put_bits(g.out, oo, (uint32_t)(x & ((1ull << k) - 1)), k);
    oo += k + (x >> k);
    put_one(g.out, oo);
  }
}

// Two launches: the tiles coded from their words here (the closing run too), the tiles coded from their lists in
// k_g2_scatter_list below; a CTA whose tile belongs to the other kernel leaves at once. (One kernel with both routes cost the
// word route a quarter of its speed on dense wide-tile streams: 4.9 against 3.9 ms at 2^31 bits, 20 % ones.)
template <int WPT>
__global__ void __launch_bounds__(G2_THREADS) k_g2_scatter(G2Params P) {
  __shared__ int s_w[16];
  __shared__ unsigned long long s_a[8];
  __shared__ uint32_t s_out[G2_STAGE_WORDS(WPT)];
  uint32_t tile;
  const G2Seg& g = g2_segment(P, &tile);
  if (g.info[4]) return;                                                 // the code does not fit: nothing is written
  const GolBase gb = g.gbp ? *g.gbp : g.gb;
  if (!(g.list && g.ones[tile] <= G2_LIST_CAP(WPT))) g2_scatter_word_tile<WPT>(g, gb, tile, s_w, s_a, s_out);   // uniform over the CTA
  g2_scatter_closing(g, gb, tile);
}

template <int WPT>
__global__ void __launch_bounds__(G2_THREADS) k_g2_scatter_list(G2Params P) {
  __shared__ unsigned long long s_a[8];
  __shared__ uint32_t s_out[G2_STAGE_WORDS(WPT)];
  uint32_t tile;
  const G2Seg& g = g2_segment(P, &tile);
  if (g.info[4] || !g.list) return;
  const uint32_t nones = g.ones[tile];
  if (nones > G2_LIST_CAP(WPT)) return;
  const GolBase gb = g.gbp ? *g.gbp : g.gb;
  g2_scatter_list_tile<WPT>(g, gb, tile, nones, s_a, s_out);
}

// ------------------------------------------------------------------ host side
bic_status bic_stream_reserve_for(bic_ctx* c, bic_stream* s, uint64_t bitcount, uint64_t nchunks, uint64_t src_bits);
bic_status bic_dense_stream_into(bic_ctx* c, const bic_mat* M, uint32_t* scratch, const uint32_t** S, uint64_t* T);

// Up to three matrices in one set of launches. outs[i] is sized here; d_info + 8 * i receives stream i's info words. Nothing
// waits for the host. Every matrix must have at least one bit (the callers route empty ones to coding.cu).
bic_status bic_k_golomb_encode_multi(bic_ctx* c, const bic_mat* const* mats, int nmat, uint32_t chunk_samples, bic_stream* const* outs,
                                     unsigned long long* d_info) {
  if (nmat < 1 || nmat > G2_MAXSEG) return BIC_ERR_INVALID;
  G2Params P;
  memset(&P, 0, sizeof(P));
  P.nseg = (uint32_t)nmat;
  // scratch: per stream, the compacted copy (only when cols is not a multiple of 32) and the per-tile arrays
  uint64_t words_compact = 0, ntiles_total = 0, words_total = 0;
  uint64_t Ts[G2_MAXSEG], nts[G2_MAXSEG];
  for (int i = 0; i < nmat; ++i) {
    const uint64_t N = mats[i]->rows * mats[i]->cols;
    if (N == 0) return BIC_ERR_INVALID;
    Ts[i] = div_up_u64(N, 32);
    words_total += Ts[i];
    if (mats[i]->cols & 31) words_compact += (Ts[i] + 7) & ~(uint64_t)3;
  }
  // wide tiles only when they still give every SM a few waves of CTAs
  const int WPT = (words_total / G2_TILE_WORDS(16) >= (uint64_t)c->sm_count * 16) ? 16 : 4;
  for (int i = 0; i < nmat; ++i) { nts[i] = div_up_u64(Ts[i], (uint64_t)G2_TILE_WORDS(WPT)); ntiles_total += nts[i]; }
  if (ntiles_total >= (1ull << 31)) return bic_fail(c, BIC_ERR_UNSUPPORTED, "golomb: too many tiles for one launch");
  if (words_compact) BIC_TRY(bic_scratch_reserve(c, &c->work[4], words_compact * 4 + 64));
  const bool use_list = c->gol_list == 2 || (c->gol_list == 1 && WPT == 16);
  const size_t list_tile = use_list ? (size_t)G2_TILE_WORDS(WPT) / 2 * 4 : 0;
  const size_t per_tile = 8 * 5 + 8 + (size_t)G2_THREADS * 8 + list_tile;   // last, ones_before, last_before, bits, bits_before | ones (padded to 8) | tbits | list
  BIC_TRY(bic_scratch_reserve(c, &c->work[5], ntiles_total * per_tile + 64 * G2_MAXSEG + 64));
  uint8_t* p = (uint8_t*)c->work[5].p;
  unsigned int* done = (unsigned int*)p;                         // 2 counters per stream, 64 bytes apart
  BIC_CUDA(c, cudaMemsetAsync(done, 0, 64 * G2_MAXSEG, c->stream));
  p += 64 * G2_MAXSEG;
  uint32_t* compact = (uint32_t*)c->work[4].p;
  uint32_t tile0 = 0;
  for (int i = 0; i < nmat; ++i) {
    G2Seg& g = P.s[i];
    const bic_mat* M = mats[i];
    g.N = M->rows * M->cols;
    // pre-sized for gol_presize_pct (default 125) % of the input bits + slack; the stream object's buffer only ever grows
    const uint64_t want_bits = g.N / 100 * (uint64_t)c->gol_presize_pct + 32768;
    BIC_TRY(bic_stream_reserve_for(c, outs[i], want_bits, div_up_u64(g.N + 1, chunk_samples), 0));
    if (M->cols & 31) {
      BIC_TRY(bic_dense_stream_into(c, M, compact, &g.S, &g.T));
      compact += (Ts[i] + 7) & ~(uint64_t)3;
    } else {
      g.S = M->d; g.T = Ts[i];
    }
    g.tile0 = tile0; g.ntiles = (uint32_t)nts[i];
    tile0 += g.ntiles;
    const size_t nt = nts[i];
    g.last = (long long*)p;                         p += nt * 8;
    g.ones_before = (unsigned long long*)p;         p += nt * 8;
    g.last_before = (long long*)p;                  p += nt * 8;
    g.bits = (unsigned long long*)p;                p += nt * 8;
    g.bits_before = (unsigned long long*)p;         p += nt * 8;
    g.ones = (uint32_t*)p;                          p += nt * 8;
    g.tbits = (unsigned long long*)p;               p += nt * (size_t)G2_THREADS * 8;
    g.list = use_list ? (uint32_t*)p : nullptr;     p += nt * list_tile;
    g.done = done + 16 * i;
    g.info = d_info + 8 * i;
    g.out = (uint32_t*)outs[i]->d_bytes;
    g.index = (unsigned long long*)outs[i]->d_index;
    g.cap_bits = (uint64_t)(outs[i]->cap_bytes - 32) * 8;
    if (c->gol_presize_pct < 100 && g.cap_bits > want_bits) g.cap_bits = want_bits;   // a test setting: pretend the buffer is that small
    g.chunk = chunk_samples;
    memset(&g.gb, 0, sizeof(g.gb));
    g.gb.prev0 = -1;
    g.gb.closing = 2;                                            // a whole matrix: the closing sample's state comes from info[]
    g.gbp = nullptr;
    outs[i]->info.coder = BIC_CODER_GOLOMB;
    outs[i]->info.chunk_samples = chunk_samples;
    outs[i]->info.rows = M->rows;
    outs[i]->info.cols = M->cols;
    outs[i]->info.bitcount = outs[i]->info.nsamples = outs[i]->info.nchunks = 0;
  }
  // gol_scan: 0 scans in the passes' last CTA, 2 always as their own launch, 1 (default) their own launch for long streams
  P.sep_scan = (c->gol_scan == 2 || (c->gol_scan == 1 && ntiles_total >= 8192)) ? 1u : 0u;
  BIC_PROF(c, KID_GOL_TILE_COUNTS);
  if (WPT == 16) k_g2_count<16><<<tile0, G2_THREADS, 0, c->stream>>>(P);
  else k_g2_count<4><<<tile0, G2_THREADS, 0, c->stream>>>(P);
  BIC_LAUNCH_CHECK(c);
  if (P.sep_scan) { k_g2_scan_tiles<0><<<nmat, G2_SCAN_THREADS, 0, c->stream>>>(P); BIC_LAUNCH_CHECK(c); }
  BIC_PROF(c, KID_GOL_LENGTHS);
  if (WPT == 16) k_g2_lengths<16><<<tile0, G2_THREADS, 0, c->stream>>>(P);
  else k_g2_lengths<4><<<tile0, G2_THREADS, 0, c->stream>>>(P);
  BIC_LAUNCH_CHECK(c);
  if (P.sep_scan) { k_g2_scan_tiles<1><<<nmat, G2_SCAN_THREADS, 0, c->stream>>>(P); BIC_LAUNCH_CHECK(c); }
  uint64_t maxN = 0;
  for (int i = 0; i < nmat; ++i) maxN = P.s[i].N > maxN ? P.s[i].N : maxN;
  BIC_PROF(c, KID_GOL_SCAN_B);
  k_g2_clear<<<bic_grid_for(c, div_up_u64(maxN, 32) + 4, 256, 4), 256, 0, c->stream>>>(P);
  BIC_LAUNCH_CHECK(c);
  BIC_PROF(c, KID_GOL_SCATTER);
  if (WPT == 16) k_g2_scatter<16><<<tile0, G2_THREADS, 0, c->stream>>>(P);
  else k_g2_scatter<4><<<tile0, G2_THREADS, 0, c->stream>>>(P);
  BIC_LAUNCH_CHECK(c);
  if (use_list) {
    if (WPT == 16) k_g2_scatter_list<16><<<tile0, G2_THREADS, 0, c->stream>>>(P);
    else k_g2_scatter_list<4><<<tile0, G2_THREADS, 0, c->stream>>>(P);
    BIC_LAUNCH_CHECK(c);
  }
  return BIC_OK;
}


// ---- the three passes as separate calls, for the row-sharded coders of coding.cu: the prefix state of a shard (GolBase) only
// exists after the ranks have exchanged their counts, and its code offset after they have exchanged their code lengths.
struct G2Plan {
  G2Params P;
  int WPT;
  uint32_t ntiles;
};
static G2Plan g_unused_plan;  // (keeps the type complete for the declarations in coding.cu)

bic_status bic_g2_plan(bic_ctx* c, const bic_mat* M, void* plan_out, size_t plan_bytes) {
  if (plan_bytes < sizeof(G2Plan)) return BIC_ERR_INVALID;
  (void)g_unused_plan;
  G2Plan* pl = (G2Plan*)plan_out;
  memset(pl, 0, sizeof(*pl));
  const uint64_t N = M->rows * M->cols;
  if (N == 0) return BIC_ERR_INVALID;
  const uint64_t T = div_up_u64(N, 32);
  pl->WPT = (T / G2_TILE_WORDS(16) >= (uint64_t)c->sm_count * 16) ? 16 : 4;
  const uint64_t nt = div_up_u64(T, (uint64_t)G2_TILE_WORDS(pl->WPT));
  if (nt >= (1ull << 31)) return bic_fail(c, BIC_ERR_UNSUPPORTED, "golomb: too many tiles for one launch");
  if (M->cols & 31) BIC_TRY(bic_scratch_reserve(c, &c->work[4], ((T + 7) & ~(uint64_t)3) * 4 + 64));
  const bool use_list = c->gol_list == 2 || (c->gol_list == 1 && pl->WPT == 16);
  const size_t list_tile = use_list ? (size_t)G2_TILE_WORDS(pl->WPT) / 2 * 4 : 0;
  const size_t per_tile = 8 * 5 + 8 + (size_t)G2_THREADS * 8 + list_tile;
  BIC_TRY(bic_scratch_reserve(c, &c->work[5], nt * per_tile + 64 * G2_MAXSEG + 64 + 64));
  uint8_t* p = (uint8_t*)c->work[5].p;
  unsigned int* done = (unsigned int*)p;
  BIC_CUDA(c, cudaMemsetAsync(done, 0, 64 * G2_MAXSEG, c->stream));
  p += 64 * G2_MAXSEG;
  G2Seg& g = pl->P.s[0];
  pl->P.nseg = 1;
  g.N = N;
  if (M->cols & 31) BIC_TRY(bic_dense_stream_into(c, M, (uint32_t*)c->work[4].p, &g.S, &g.T));
  else { g.S = M->d; g.T = T; }
  g.tile0 = 0; g.ntiles = (uint32_t)nt;
  pl->ntiles = (uint32_t)nt;
  g.last = (long long*)p;                         p += nt * 8;
  g.ones_before = (unsigned long long*)p;         p += nt * 8;
  g.last_before = (long long*)p;                  p += nt * 8;
  g.bits = (unsigned long long*)p;                p += nt * 8;
  g.bits_before = (unsigned long long*)p;         p += nt * 8;
  g.ones = (uint32_t*)p;                          p += nt * 8;
  g.tbits = (unsigned long long*)p;               p += nt * (size_t)G2_THREADS * 8;
  g.list = use_list ? (uint32_t*)p : nullptr;     p += nt * list_tile;
  g.info = (unsigned long long*)p;                 // 8 words after the per-tile arrays
  g.done = done;
  g.cap_bits = ~0ull >> 1;
  g.chunk = 256;
  memset(&g.gb, 0, sizeof(g.gb));
  g.gb.prev0 = -1;
  g.gbp = nullptr;
  return BIC_OK;
}

// pass 1; afterwards (stream order) the per-tile counts and their prefixes exist. ones / last-one position are returned through
// d_tot (device, 2 u64: ones, position after the last one or 0), computed by a tiny kernel from the tile arrays.
__global__ void k_g2_totals(G2Seg g, unsigned long long* tot) {
  const uint64_t lt = g.ntiles - 1;
  const long long lastg = g.last_before[lt] > g.last[lt] ? g.last_before[lt] : g.last[lt];
  tot[0] = g.ones_before[lt] + g.ones[lt];
  tot[1] = (unsigned long long)(lastg + 1);
}

bic_status bic_g2_count(bic_ctx* c, void* plan, unsigned long long* d_tot) {
  G2Plan* pl = (G2Plan*)plan;
  BIC_PROF(c, KID_GOL_TILE_COUNTS);
  if (pl->WPT == 16) k_g2_count<16><<<pl->ntiles, G2_THREADS, 0, c->stream>>>(pl->P);
  else k_g2_count<4><<<pl->ntiles, G2_THREADS, 0, c->stream>>>(pl->P);
  BIC_LAUNCH_CHECK(c);
  k_g2_totals<<<1, 1, 0, c->stream>>>(pl->P.s[0], d_tot);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

// pass 2 under the shard's prefix state; the shard's own code bits land in info[2] (device) = *d_bits
bic_status bic_g2_lengths(bic_ctx* c, void* plan, const GolBase* gb, unsigned long long** d_info) {
  G2Plan* pl = (G2Plan*)plan;
  pl->P.s[0].gb = *gb;
  BIC_PROF(c, KID_GOL_LENGTHS);
  if (pl->WPT == 16) k_g2_lengths<16><<<pl->ntiles, G2_THREADS, 0, c->stream>>>(pl->P);
  else k_g2_lengths<4><<<pl->ntiles, G2_THREADS, 0, c->stream>>>(pl->P);
  BIC_LAUNCH_CHECK(c);
  *d_info = pl->P.s[0].info;
  return BIC_OK;
}

// pass 3 into `out` (already zeroed by the caller), with the complete base (code0 / out0 / chunk0 / closing now known)
bic_status bic_g2_scatter(bic_ctx* c, void* plan, const GolBase* gb, uint32_t chunk, bic_stream* out) {
  G2Plan* pl = (G2Plan*)plan;
  G2Seg& g = pl->P.s[0];
  g.gb = *gb;
  g.chunk = chunk;
  g.out = (uint32_t*)out->d_bytes;
  g.index = (unsigned long long*)out->d_index;
  BIC_CUDA(c, cudaMemsetAsync(g.info + 4, 0, 8, c->stream));   // the overflow flag of the single-stream path: the buffer is exact here
  BIC_PROF(c, KID_GOL_SCATTER);
  if (pl->WPT == 16) k_g2_scatter<16><<<pl->ntiles, G2_THREADS, 0, c->stream>>>(pl->P);
  else k_g2_scatter<4><<<pl->ntiles, G2_THREADS, 0, c->stream>>>(pl->P);
  BIC_LAUNCH_CHECK(c);
  if (g.list) {
    if (pl->WPT == 16) k_g2_scatter_list<16><<<pl->ntiles, G2_THREADS, 0, c->stream>>>(pl->P);
    else k_g2_scatter_list<4><<<pl->ntiles, G2_THREADS, 0, c->stream>>>(pl->P);
    BIC_LAUNCH_CHECK(c);
  }
  return BIC_OK;
}


// ------------------------------------------------------------------ row shards with nothing waiting for the host (pipeline.cu)
// The prefix state of a shard is computed ON THE DEVICE from what the ranks exchange (two small all-gathers queued on the stream by
// the caller's hook): base1 after the count pass -- ones / last one / bits of every rank -> t0, pos0, prev0 and the global totals;
// base2 after the length pass -- every rank's code bits -> code0, out0, chunk0, the closing sample, this shard's info[] and
// bic_shard_info. all1: [rank][matrix][3] = ones, position after the last one (0: none), bits. all2: [rank][matrix] code bits.
struct G2ShardAux { unsigned long long ones_global, bits_global, consumed_global, ones_local; };

__global__ void k_g2_shard_base1(const unsigned long long* __restrict__ all1, uint32_t nr, uint32_t rank, uint32_t nmat,
                                 GolBase* __restrict__ base, G2ShardAux* __restrict__ aux) {
  const uint32_t i = threadIdx.x;
  if (i >= nmat) return;
  GolBase b;
  memset(&b, 0, sizeof(b));
  b.prev0 = -1;
  unsigned long long pos = 0, og = 0;
  long long lastg = -1;
  for (uint32_t r = 0; r < nr; ++r) {
    const unsigned long long* e = all1 + ((size_t)r * nmat + i) * 3;
    if (r == rank) { b.t0 = og; b.pos0 = (long long)pos; b.prev0 = lastg; aux[i].ones_local = e[0]; }
    if (e[1]) lastg = (long long)(pos + e[1] - 1);
    og += e[0];
    pos += e[2];
  }
  aux[i].ones_global = og; aux[i].bits_global = pos; aux[i].consumed_global = (unsigned long long)(lastg + 1);
  base[i] = b;
}

struct G2ShardOut { unsigned long long* info[G2_MAXSEG]; unsigned long long cap_bits[G2_MAXSEG]; };

__global__ void k_g2_shard_base2(const unsigned long long* __restrict__ all2, uint32_t nr, uint32_t rank, uint32_t nmat, uint32_t chunk,
                                 GolBase* __restrict__ base, const G2ShardAux* __restrict__ aux, G2ShardOut o,
                                 unsigned long long* __restrict__ shard /* 6 per matrix: bic_shard_info */) {
  const uint32_t i = threadIdx.x;
  if (i >= nmat) return;
  GolBase b = base[i];
  unsigned long long code0 = 0, total = 0;
  for (uint32_t r = 0; r < nr; ++r) { const unsigned long long v = all2[(size_t)r * nmat + i]; if (r < rank) code0 += v; total += v; }
  const unsigned long long mine = all2[(size_t)rank * nmat + i];
  const unsigned long long consumed = aux[i].consumed_global;
  const uint32_t kc = golomb_k(aux[i].ones_global, consumed);
  const unsigned long long closing_bits = kc + ((aux[i].bits_global - consumed) >> kc) + 1;
  b.code0 = code0;
  b.out0 = code0 & 31;
  b.chunk0 = (b.t0 + chunk - 1) / chunk;
  unsigned long long local_bits = mine, local_samples = aux[i].ones_local;
  b.closing = 0;
  if (rank == nr - 1) {
    b.closing = 1;
    b.close_t = aux[i].ones_global;
    b.close_consumed = consumed;
    b.close_off = mine;
    b.close_n = aux[i].bits_global;
    local_bits += closing_bits;
    local_samples += 1;
  }
  base[i] = b;
  const unsigned long long nchunks = (b.t0 + local_samples + chunk - 1) / chunk - b.chunk0;
  unsigned long long* info = o.info[i];
  info[0] = b.out0 + local_bits;       // bits of the local buffer, incl. the (code0 & 31) leading pad
  info[1] = local_samples;
  info[3] = nchunks;                   // (a shard's chunk count is not ceil(samples / chunk): the caller takes it from here)
  info[4] = (info[0] + 64 > o.cap_bits[i]) ? 1ull : 0ull;
  unsigned long long* sh = shard + 6 * i;
  sh[0] = total + closing_bits;        // global_bitcount
  sh[1] = aux[i].ones_global + 1;      // global_nsamples
  sh[2] = code0;                       // code_bit_offset
  sh[3] = local_bits;                  // local_code_bits
  sh[4] = b.chunk0;                    // first_chunk
  sh[5] = nchunks;                     // local_chunks
}

__global__ void k_g2_pack(const unsigned long long* a0, const unsigned long long* a1, const unsigned long long* a2, unsigned long long* dst, uint32_t n,
                          uint32_t stride, uint32_t off, unsigned long long n0, unsigned long long n1, unsigned long long n2) {
  // dst[i * stride + off] = a_i[0]; with stride 3 (pass 1) also the matrix's bit count at + 2 and a_i[1] at + 1
  const unsigned long long* a[3] = {a0, a1, a2};
  const unsigned long long nb[3] = {n0, n1, n2};
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    if (stride == 3) { dst[i * 3] = a[i][0]; dst[i * 3 + 1] = a[i][1]; dst[i * 3 + 2] = nb[i]; }
    else dst[i] = a[i][off];
  }
}

bic_status bic_k_golomb_encode_multi_sharded(bic_ctx* c, const bic_mat* const* mats, int nmat, uint32_t chunk_samples, bic_stream* const* outs,
                                             unsigned long long* d_info, unsigned long long* d_shard, uint32_t nranks, uint32_t rank,
                                             bic_status (*allgather)(void* user, bic_ctx* c, const unsigned long long* d_src, int count,
                                                                     unsigned long long* d_dst),
                                             void* user) {
  if (nmat < 1 || nmat > G2_MAXSEG || nranks < 1 || nranks > 64) return BIC_ERR_INVALID;
  G2Params P;
  memset(&P, 0, sizeof(P));
  P.nseg = (uint32_t)nmat;
  uint64_t words_compact = 0, ntiles_total = 0, words_total = 0;
  uint64_t Ts[G2_MAXSEG], nts[G2_MAXSEG];
  for (int i = 0; i < nmat; ++i) {
    const uint64_t N = mats[i]->rows * mats[i]->cols;
    if (N == 0) return BIC_ERR_INVALID;
    Ts[i] = div_up_u64(N, 32);
    words_total += Ts[i];
    if (mats[i]->cols & 31) words_compact += (Ts[i] + 7) & ~(uint64_t)3;
  }
  const int WPT = (words_total / G2_TILE_WORDS(16) >= (uint64_t)c->sm_count * 16) ? 16 : 4;
  for (int i = 0; i < nmat; ++i) { nts[i] = div_up_u64(Ts[i], (uint64_t)G2_TILE_WORDS(WPT)); ntiles_total += nts[i]; }
  if (ntiles_total >= (1ull << 31)) return bic_fail(c, BIC_ERR_UNSUPPORTED, "golomb: too many tiles for one launch");
  if (words_compact) BIC_TRY(bic_scratch_reserve(c, &c->work[4], words_compact * 4 + 64));
  const bool use_list = c->gol_list == 2 || (c->gol_list == 1 && WPT == 16);
  const size_t list_tile = use_list ? (size_t)G2_TILE_WORDS(WPT) / 2 * 4 : 0;
  const size_t per_tile = 8 * 5 + 8 + (size_t)G2_THREADS * 8 + list_tile;
  // tail of the scratch: counters | base[3] | aux[3] | tot[3][2] | pack1 (3*3) | all1 (nranks*3*3) | pack2 (3) | all2 (nranks*3)
  const size_t tail = 64 * G2_MAXSEG + sizeof(GolBase) * G2_MAXSEG + sizeof(G2ShardAux) * G2_MAXSEG + 8 * (6 + 9 + 3) + 8 * (size_t)nranks * (9 + 3) + 256;
  BIC_TRY(bic_scratch_reserve(c, &c->work[5], ntiles_total * per_tile + tail));
  uint8_t* p = (uint8_t*)c->work[5].p;
  unsigned int* done = (unsigned int*)p;                     p += 64 * G2_MAXSEG;
  BIC_CUDA(c, cudaMemsetAsync(done, 0, 64 * G2_MAXSEG, c->stream));
  GolBase* d_base = (GolBase*)p;                             p += (sizeof(GolBase) * G2_MAXSEG + 15) & ~(size_t)15;
  G2ShardAux* d_aux = (G2ShardAux*)p;                        p += sizeof(G2ShardAux) * G2_MAXSEG;
  unsigned long long* d_tot = (unsigned long long*)p;        p += 8 * 6;
  unsigned long long* d_pack1 = (unsigned long long*)p;      p += 8 * 9;
  unsigned long long* d_all1 = (unsigned long long*)p;       p += 8 * (size_t)nranks * 9;
  unsigned long long* d_pack2 = (unsigned long long*)p;      p += 8 * 3;
  unsigned long long* d_all2 = (unsigned long long*)p;       p += 8 * (size_t)nranks * 3;
  p = (uint8_t*)(((uintptr_t)p + 63) & ~(uintptr_t)63);
  uint32_t* compact = (uint32_t*)c->work[4].p;
  uint32_t tile0 = 0;
  G2ShardOut so;
  memset(&so, 0, sizeof(so));
  for (int i = 0; i < nmat; ++i) {
    G2Seg& g = P.s[i];
    const bic_mat* M = mats[i];
    g.N = M->rows * M->cols;
    const uint64_t want_bits = g.N / 100 * (uint64_t)c->gol_presize_pct + 32768;
    BIC_TRY(bic_stream_reserve_for(c, outs[i], want_bits, div_up_u64(g.N + 1, chunk_samples) + 2, 0));
    if (M->cols & 31) {
      BIC_TRY(bic_dense_stream_into(c, M, compact, &g.S, &g.T));
      compact += (Ts[i] + 7) & ~(uint64_t)3;
    } else {
      g.S = M->d; g.T = Ts[i];
    }
    g.tile0 = tile0; g.ntiles = (uint32_t)nts[i];
    tile0 += g.ntiles;
    const size_t nt = nts[i];
    g.last = (long long*)p;                         p += nt * 8;
    g.ones_before = (unsigned long long*)p;         p += nt * 8;
    g.last_before = (long long*)p;                  p += nt * 8;
    g.bits = (unsigned long long*)p;                p += nt * 8;
    g.bits_before = (unsigned long long*)p;         p += nt * 8;
    g.ones = (uint32_t*)p;                          p += nt * 8;
    g.tbits = (unsigned long long*)p;               p += nt * (size_t)G2_THREADS * 8;
    g.list = use_list ? (uint32_t*)p : nullptr;     p += nt * list_tile;
    g.done = done + 16 * i;
    g.info = d_info + 8 * i;
    g.out = (uint32_t*)outs[i]->d_bytes;
    g.index = (unsigned long long*)outs[i]->d_index;
    g.cap_bits = (uint64_t)(outs[i]->cap_bytes - 32) * 8;
    if (c->gol_presize_pct < 100 && g.cap_bits > want_bits) g.cap_bits = want_bits;
    g.chunk = chunk_samples;
    memset(&g.gb, 0, sizeof(g.gb));
    g.gbp = d_base + i;
    so.info[i] = g.info;
    so.cap_bits[i] = g.cap_bits;
    outs[i]->info.coder = BIC_CODER_GOLOMB;
    outs[i]->info.chunk_samples = chunk_samples;
    outs[i]->info.rows = M->rows;
    outs[i]->info.cols = M->cols;
    outs[i]->info.bitcount = outs[i]->info.nsamples = outs[i]->info.nchunks = 0;
  }
  BIC_CUDA(c, cudaMemsetAsync(d_info, 0, 8 * 8 * (size_t)nmat, c->stream));
  BIC_PROF(c, KID_GOL_TILE_COUNTS);
  if (WPT == 16) k_g2_count<16><<<tile0, G2_THREADS, 0, c->stream>>>(P);
  else k_g2_count<4><<<tile0, G2_THREADS, 0, c->stream>>>(P);
  BIC_LAUNCH_CHECK(c);
  for (int i = 0; i < nmat; ++i) {
    k_g2_totals<<<1, 1, 0, c->stream>>>(P.s[i], d_tot + 2 * i);
    BIC_LAUNCH_CHECK(c);
  }
  k_g2_pack<<<1, 32, 0, c->stream>>>(d_tot, d_tot + 2, d_tot + 4, d_pack1, (uint32_t)nmat, 3, 0, P.s[0].N, P.s[1].N, P.s[2].N);
  BIC_LAUNCH_CHECK(c);
  BIC_TRY(allgather(user, c, d_pack1, 3 * nmat, d_all1));
  k_g2_shard_base1<<<1, 32, 0, c->stream>>>(d_all1, nranks, rank, (uint32_t)nmat, d_base, d_aux);
  BIC_LAUNCH_CHECK(c);
  BIC_PROF(c, KID_GOL_LENGTHS);
  if (WPT == 16) k_g2_lengths<16><<<tile0, G2_THREADS, 0, c->stream>>>(P);
  else k_g2_lengths<4><<<tile0, G2_THREADS, 0, c->stream>>>(P);
  BIC_LAUNCH_CHECK(c);
  k_g2_pack<<<1, 32, 0, c->stream>>>(d_info, d_info + 8, d_info + 16, d_pack2, (uint32_t)nmat, 1, 2, 0, 0, 0);
  BIC_LAUNCH_CHECK(c);
  BIC_TRY(allgather(user, c, d_pack2, nmat, d_all2));
  k_g2_shard_base2<<<1, 32, 0, c->stream>>>(d_all2, nranks, rank, (uint32_t)nmat, chunk_samples, d_base, d_aux, so, d_shard);
  BIC_LAUNCH_CHECK(c);
  uint64_t maxN = 0;
  for (int i = 0; i < nmat; ++i) maxN = P.s[i].N > maxN ? P.s[i].N : maxN;
  BIC_PROF(c, KID_GOL_SCAN_B);
  k_g2_clear<<<bic_grid_for(c, div_up_u64(maxN, 32) + 4, 256, 4), 256, 0, c->stream>>>(P);
  BIC_LAUNCH_CHECK(c);
  BIC_PROF(c, KID_GOL_SCATTER);
  if (WPT == 16) k_g2_scatter<16><<<tile0, G2_THREADS, 0, c->stream>>>(P);
  else k_g2_scatter<4><<<tile0, G2_THREADS, 0, c->stream>>>(P);
  BIC_LAUNCH_CHECK(c);
  if (use_list) {
    if (WPT == 16) k_g2_scatter_list<16><<<tile0, G2_THREADS, 0, c->stream>>>(P);
    else k_g2_scatter_list<4><<<tile0, G2_THREADS, 0, c->stream>>>(P);
    BIC_LAUNCH_CHECK(c);
  }
  return BIC_OK;
}
