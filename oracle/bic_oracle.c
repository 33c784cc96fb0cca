/*
 * TEST INFRASTRUCTURE -- the CPU oracle. NOT part of the product. See bic_oracle.h for the
 * pinning statement and the layout convention. Every function cites the reference
 * file:line (relative to /root/reference/) whose behaviour it restates.
 */
#include "bic_oracle.h"

#include <stdlib.h>
#include <string.h>

#define BO_MSB ((bo_word)1 << 63)

uint64_t bo_wpr(uint64_t cols) { return (cols + 63) / 64; } /* src/binmat.cpp:143 */

static inline int popc64(bo_word v) { return __builtin_popcountll(v); } /* = block_weight, src/binmat.cpp:22-37 */

static inline int get_bit(const bo_word* row, uint64_t j) { /* src/binmat.h:114-116 */
  return (int)((row[j >> 6] >> (63 - (j & 63))) & 1u);
}
static inline void put_bit(bo_word* row, uint64_t j, int v) { /* src/binmat.h:121-141 */
  const bo_word mask = BO_MSB >> (j & 63);
  if (v) row[j >> 6] |= mask; else row[j >> 6] &= ~mask;
}

/* src/binmat.cpp:57-67 (pad bits are zero by convention, so no trail mask is needed) */
uint64_t bo_weight(const bo_word* M, uint64_t rows, uint64_t cols) {
  const uint64_t total = rows * bo_wpr(cols);
  uint64_t w = 0;
  for (uint64_t i = 0; i < total; ++i) w += (uint64_t)popc64(M[i]);
  return w;
}

/* ------------------------------------------------------------------------------------------
 * Patch extraction. src/bsvd_test.cpp:80-99.
 *
 * copy_submatrix_to (src/binmat.cpp:267-298) reads the raster through LINEAR word indices
 * k = i0*bpr + j0/64 + di*bpr + dj (and k+1 for unaligned tiles), substituting 0 only when
 * k >= data_blocks. So pixel (r, c) is the bit at linear position r*bpr*64 + c of the padded
 * raster: columns cols..bpr*64-1 read the (zero) pad, and a column >= bpr*64 -- which only an
 * edge tile with W not dividing 64 can ask for -- falls through into the next raster row.
 * That is reproduced here on purpose. Tile rows are then masked to W bits (get_block,
 * src/binmat.h:188-190) and concatenated row-major, MSB first (copy_vectorized_to, :306-320;
 * well defined for W < 64, which is all the reference's drivers use).
 * ------------------------------------------------------------------------------------------ */
static inline int raster_bit(const bo_word* I, uint64_t data_blocks, uint64_t bpr, uint64_t r, uint64_t c) {
  const uint64_t k = r * bpr + (c >> 6);
  if (k >= data_blocks) return 0;
  return (int)((I[k] >> (63 - (c & 63))) & 1u);
}

void bo_extract_patches(const bo_word* I, uint64_t rows, uint64_t cols, uint64_t W, bo_word* X) {
  const uint64_t bpr = bo_wpr(cols);
  const uint64_t data_blocks = bpr * rows;
  const uint64_t Ny = (W - 1 + rows) / W, Nx = (W - 1 + cols) / W; /* bsvd_test.cpp:82-83 */
  const uint64_t m = W * W, xw = bo_wpr(m);
  memset(X, 0, sizeof(bo_word) * Nx * Ny * xw);
  uint64_t li = 0;
  for (uint64_t i = 0; i < Ny; ++i) {
    for (uint64_t j = 0; j < Nx; ++j, ++li) { /* row-major patch order, bsvd_test.cpp:92-93 */
      bo_word* xr = X + li * xw;
      for (uint64_t pr = 0; pr < W; ++pr)
        for (uint64_t pc = 0; pc < W; ++pc)
          if (raster_bit(I, data_blocks, bpr, i * W + pr, j * W + pc)) put_bit(xr, pr * W + pc, 1);
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * RNG: GSL's rand48 and gsl_rng_uniform_int (un-vendored dependency; call sites
 * src/bsvd.cpp:8-15, :241). Restated from GSL's published rng/rand48.c and rng/rng.c.
 * ------------------------------------------------------------------------------------------ */
void bo_rand48_seed(bo_rand48* r, unsigned long s) {
  if (s == 0) { r->x0 = 0x330E; r->x1 = 0xABCD; r->x2 = 0x1234; }
  else { r->x0 = 0x330E; r->x1 = (uint16_t)(s & 0xFFFF); r->x2 = (uint16_t)((s >> 16) & 0xFFFF); }
}

uint32_t bo_rand48_next(bo_rand48* r) {
  /* one 48-bit multiply-add instead of GSL's three 16-bit limbs; same recurrence */
  uint64_t x = ((uint64_t)r->x2 << 32) | ((uint64_t)r->x1 << 16) | (uint64_t)r->x0;
  x = (x * 0x5DEECE66DULL + 0xBULL) & 0xFFFFFFFFFFFFULL;
  r->x0 = (uint16_t)(x & 0xFFFF);
  r->x1 = (uint16_t)((x >> 16) & 0xFFFF);
  r->x2 = (uint16_t)((x >> 32) & 0xFFFF);
  return (uint32_t)(x >> 16);
}

uint64_t bo_uniform_int(bo_rand48* r, uint64_t n) {
  const uint64_t range = 0xFFFFFFFFULL;
  if (n == 0 || n > range) return 0;
  const uint64_t scale = range / n;
  uint64_t k;
  do { k = (uint64_t)bo_rand48_next(r) / scale; } while (k >= n);
  return k;
}

/* ------------------------------------------------------------------------------------------
 * initialize_model_neighbor. src/bsvd.cpp:227-267.
 * ------------------------------------------------------------------------------------------ */
uint64_t bo_draw_pivots(const bo_word* X, uint64_t n, uint64_t m, uint64_t p, bo_rand48* rng,
                        uint64_t* pivots) {
  const uint64_t wpr = bo_wpr(m);
  uint64_t draws = 0;
  for (uint64_t k = 0; k < p;) {            /* :239 */
    const uint64_t i = bo_uniform_int(rng, n); /* :241 */
    ++draws;
    const bo_word* Ei = X + i * wpr;
    uint64_t w = 0;
    for (uint64_t b = 0; b < wpr; ++b) w += (uint64_t)popc64(Ei[b]);
    if (w == 0) continue;                   /* :243  rejected draw is consumed */
    /* a non-zero pivot intersects itself, so u > 0 (:258) always holds and k advances */
    pivots[k++] = i;
  }
  return draws;
}

void bo_init_neighbor_pivots(const bo_word* X, uint64_t n, uint64_t m, uint64_t p,
                             const uint64_t* pivots, bo_word* D, bo_word* A) {
  const uint64_t wpr = bo_wpr(m), apr = bo_wpr(p);
  memset(A, 0, sizeof(bo_word) * n * apr); /* A.clear(), :237 */
  memset(D, 0, sizeof(bo_word) * p * wpr); /* D.clear(), :238 */
  uint64_t* s = (uint64_t*)malloc(sizeof(uint64_t) * (m ? m : 1));
  bo_word* Ej = (bo_word*)malloc(sizeof(bo_word) * (wpr ? wpr : 1));
  for (uint64_t k = 0; k < p; ++k) {
    const bo_word* Ei = X + pivots[k] * wpr;
    uint64_t u = 0;
    for (uint64_t b = 0; b < m; ++b) s[b] = 0; /* :246 */
    for (uint64_t j = 0; j < n; ++j) {         /* :247 */
      uint64_t w = 0;
      for (uint64_t b = 0; b < wpr; ++b) {     /* Ej = E[j] & Ei, :248-249 */
        Ej[b] = X[j * wpr + b] & Ei[b];
        w += (uint64_t)popc64(Ej[b]);
      }
      if (w > 0) {                             /* :250 */
        u++;
        for (uint64_t b = 0; b < m; ++b)       /* counts the AND, not E[j]; :252-255 */
          if (get_bit(Ej, b)) s[b]++;
      }
    }
    if (u > 0) {                               /* :258 */
      for (uint64_t b = 0; b < m; ++b) put_bit(D + k * wpr, b, s[b] >= u / 2); /* ">=" and integer u/2, :259-260 */
    }
  }
  free(s);
  free(Ej);
}

void bo_init_neighbor(const bo_word* X, uint64_t n, uint64_t m, uint64_t p, bo_rand48* rng,
                      bo_word* D, bo_word* A) {
  uint64_t* pivots = (uint64_t*)malloc(sizeof(uint64_t) * (p ? p : 1));
  bo_draw_pivots(X, n, m, p, rng, pivots);
  bo_init_neighbor_pivots(X, n, m, p, pivots, D, A);
  free(pivots);
}

/* ------------------------------------------------------------------------------------------
 * update_coefficients_omp (src/bsvd.cpp:1029-1107) == update_coefficients_basic (:399-460):
 * greedy matching pursuit over GF(2), one row at a time, rows independent given D.
 * ------------------------------------------------------------------------------------------ */
uint64_t bo_update_coefficients(bo_word* E, const bo_word* D, bo_word* A,
                                uint64_t n, uint64_t m, uint64_t p) {
  const uint64_t wpr = bo_wpr(m), apr = bo_wpr(p);
  uint64_t changed = 0;
  if (p == 0) return 0;
  for (uint64_t i = 0; i < n; ++i) {
    bo_word* Ei = E + i * wpr;
    bo_word* Ai = A + i * apr;
    int ichanged = 0;
    for (;;) {
      uint64_t w = 0;                                   /* w = Ei.weight(), :1065 */
      for (uint64_t b = 0; b < wpr; ++b) w += (uint64_t)popc64(Ei[b]);
      uint64_t bestk = 0, bestd = 0;                    /* :1067 */
      for (uint64_t b = 0; b < wpr; ++b) bestd += (uint64_t)popc64(Ei[b] ^ D[b]);
      for (uint64_t k = 1; k < p; ++k) {                /* :1071-1082 */
        uint64_t dk = 0;
        for (uint64_t b = 0; b < wpr; ++b) dk += (uint64_t)popc64(Ei[b] ^ D[k * wpr + b]);
        if (dk < bestd) { bestd = dk; bestk = k; }      /* strict <: lowest k wins ties */
      }
      if (bestd < w) {                                  /* strict <, :1084 */
        Ai[bestk >> 6] ^= BO_MSB >> (bestk & 63);       /* Ai.flip(0,bestk), :1086 */
        for (uint64_t b = 0; b < wpr; ++b) Ei[b] ^= D[bestk * wpr + b]; /* :1087 */
        ichanged = 1;
      } else {
        break;                                          /* :1090-1092 */
      }
    }
    if (ichanged) changed++;                            /* :1095-1096 */
  }
  return changed;
}

/* ------------------------------------------------------------------------------------------
 * update_dictionary_steepest. src/bsvd.cpp:463-527. Atoms strictly in order; E is patched
 * after each changed atom, so atom k+1 sees atom k's outcome.
 * ------------------------------------------------------------------------------------------ */
uint64_t bo_update_dictionary(bo_word* E, bo_word* D, const bo_word* A,
                              uint64_t n, uint64_t m, uint64_t p) {
  const uint64_t wpr = bo_wpr(m), apr = bo_wpr(p);
  uint64_t* weights = (uint64_t*)malloc(sizeof(uint64_t) * (m ? m : 1));
  bo_word* newDk = (bo_word*)malloc(sizeof(bo_word) * (wpr ? wpr : 1));
  uint64_t changed = 0;
  for (uint64_t k = 0; k < p; ++k) {
    bo_word* Dk = D + k * wpr;
    uint64_t usage = 0;
    for (uint64_t j = 0; j < m; ++j) weights[j] = 0;
    for (uint64_t i = 0; i < n; ++i) {                  /* :486-498 */
      if (!get_bit(A + i * apr, k)) continue;
      usage++;
      for (uint64_t j = 0; j < m; ++j)                  /* bits of E[i] ^ Dk */
        if (get_bit(E + i * wpr, j) ^ get_bit(Dk, j)) weights[j]++;
    }
    if (!usage) continue;                               /* :499-500 */
    const uint64_t u = usage / 2;                       /* :502 */
    memcpy(newDk, Dk, sizeof(bo_word) * wpr);           /* :503 */
    for (uint64_t j = 0; j < m; ++j) put_bit(newDk, j, weights[j] > u); /* strict >, :504-506 */
    uint64_t d = 0;
    for (uint64_t b = 0; b < wpr; ++b) d += (uint64_t)popc64(newDk[b] ^ Dk[b]);
    if (d > 0) {                                        /* :507 */
      changed++;
      for (uint64_t i = 0; i < n; ++i) {                /* :512-520 */
        if (!get_bit(A + i * apr, k)) continue;
        for (uint64_t b = 0; b < wpr; ++b) E[i * wpr + b] ^= Dk[b] ^ newDk[b];
      }
      memcpy(Dk, newDk, sizeof(bo_word) * wpr);         /* D.set_row(k,newDk), :510 */
    }
  }
  free(weights);
  free(newDk);
  return changed;
}

/* mul_AB (src/binmat.cpp:516-543) then add (:463-478) */
void bo_residual(const bo_word* X, const bo_word* A, const bo_word* D, bo_word* E,
                 uint64_t n, uint64_t m, uint64_t p) {
  const uint64_t wpr = bo_wpr(m), apr = bo_wpr(p);
  for (uint64_t i = 0; i < n; ++i) {
    bo_word* Ei = E + i * wpr;
    for (uint64_t b = 0; b < wpr; ++b) Ei[b] = 0;
    for (uint64_t k = 0; k < p; ++k)
      if (get_bit(A + i * apr, k))
        for (uint64_t b = 0; b < wpr; ++b) Ei[b] ^= D[k * wpr + b];
    for (uint64_t b = 0; b < wpr; ++b) Ei[b] ^= X[i * wpr + b];
  }
}

/* src/bsvd.cpp:1215-1244 */
uint64_t bo_learn_traditional(const bo_word* X, bo_word* E, bo_word* D, bo_word* A,
                              uint64_t n, uint64_t m, uint64_t p,
                              uint64_t* trace, uint64_t trace_cap) {
  bo_residual(X, A, D, E, n, m, p);                     /* :1219-1220 */
  uint64_t changed = 1, iter = 0;
  while (changed > 0) {                                 /* :1227 */
    iter++;
    const uint64_t cc = bo_update_coefficients(E, D, A, n, m, p); /* :1229 */
    const uint64_t ca = bo_update_dictionary(E, D, A, n, m, p);   /* :1235 */
    changed = cc + ca;
    if (trace && iter <= trace_cap) { trace[2 * (iter - 1)] = cc; trace[2 * (iter - 1) + 1] = ca; }
  }
  return iter;
}

/* ------------------------------------------------------------------------------------------
 * Bit I/O
 * ------------------------------------------------------------------------------------------ */
static void bw_put_bit(bo_bitwriter* bw, int bit) {
  if (bw->nbits < bw->cap_bits) {
    if (bit) bw->buf[bw->nbits >> 3] |= (uint8_t)(0x80u >> (bw->nbits & 7));
  }
  bw->nbits++;
}
static void bw_put_bits(bo_bitwriter* bw, uint32_t value, unsigned nbits) { /* writeBits(value, nbits): MSB of the field first */
  for (unsigned i = nbits; i-- > 0;) bw_put_bit(bw, (int)((value >> i) & 1u));
}
static void bw_put_zeros(bo_bitwriter* bw, uint64_t n) { bw->nbits += n; } /* buffer is pre-zeroed */

static int br_get_bit(bo_bitreader* br) {
  int b = 0;
  if (br->pos < br->nbits) b = (br->buf[br->pos >> 3] >> (7 - (br->pos & 7))) & 1;
  br->pos++;
  return b;
}
static uint32_t br_get_bits(bo_bitreader* br, unsigned nbits) {
  uint32_t v = 0;
  for (unsigned i = 0; i < nbits; ++i) v = (v << 1) | (uint32_t)br_get_bit(br);
  return v;
}

/* ------------------------------------------------------------------------------------------
 * Golomb. src/Golomb.h:12-29, src/GolombCoder.cpp:13-34, src/GolombDecoder.cpp:15-23.
 * ------------------------------------------------------------------------------------------ */
void bo_golomb_init(bo_golomb* g) { g->accumulatedError = 0; g->samples = 0; g->k = 1; g->bitcount = 0; } /* Golomb.h:14-19 */

static void golomb_adapt(bo_golomb* g, uint32_t sample) {
  g->samples++;                                         /* GolombCoder.cpp:31 */
  g->accumulatedError += sample;                        /* :32, uint32 wrap-around */
  uint32_t k;
  /* :33  for(k=0; (samples<<k) < accumulatedError; k++);  in unsigned 32-bit arithmetic.
   * k >= 32 would be an out-of-range shift there and trips the assert at :14 on the next
   * sample; the search stops at 31 so the state stays inside the reference's domain. */
  for (k = 0; k < 31 && (uint32_t)(g->samples << k) < g->accumulatedError; k++) {}
  g->k = k;
}

void bo_golomb_code_sample(bo_golomb* g, bo_bitwriter* bw, uint32_t sample) {
  const uint32_t k = g->k;
  const uint32_t unary = sample >> k;                   /* GolombCoder.cpp:19 */
  if (bw) {
    const uint32_t binary = k ? (sample & (0xFFFFFFFFu >> (32 - k))) : 0u; /* :18 */
    bw_put_bits(bw, binary, k);                         /* :22  file->writeBits(binary, k)   */
    bw_put_zeros(bw, unary);                            /* :24  file->writeZeros(unary)      */
    bw_put_bit(bw, 1);                                  /* :25  file->writeBits(1, 1)        */
  }
  g->bitcount += (int64_t)k + (int64_t)unary + 1;       /* :26 */
  golomb_adapt(g, sample);
}

uint32_t bo_golomb_decode_sample(bo_golomb* g, bo_bitreader* br) {
  const uint32_t k = g->k;
  const uint32_t binary = br_get_bits(br, k);           /* GolombDecoder.cpp:17 readBits(k)   */
  uint32_t unary = 0;                                   /* :18 countZeros()                   */
  while (br->pos < br->nbits && !br_get_bit(br)) unary++; /* consumes the terminating one (:19) */
  const uint32_t sample = (unary << k) | binary;        /* :21 */
  g->bitcount += (int64_t)k + (int64_t)unary + 1;
  golomb_adapt(g, sample);                              /* :36-37 */
  return sample;
}

uint64_t bo_zero_runs(const bo_word* M, uint64_t rows, uint64_t cols, uint32_t* samples, uint64_t cap) {
  const uint64_t wpr = bo_wpr(cols);
  uint64_t count = 0, run = 0;
  for (uint64_t i = 0; i < rows; ++i)
    for (uint64_t j = 0; j < cols; ++j) {
      if (get_bit(M + i * wpr, j)) {
        if (samples && count < cap) samples[count] = (uint32_t)run;
        count++;
        run = 0;
      } else {
        run++;
      }
    }
  if (samples && count < cap) samples[count] = (uint32_t)run; /* run closed by the virtual one */
  count++;
  return count;
}

uint64_t bo_golomb_encode_matrix(const bo_word* M, uint64_t rows, uint64_t cols,
                                 uint8_t* out, uint64_t cap_bytes, uint64_t* nsamples) {
  const uint64_t wpr = bo_wpr(cols);
  bo_golomb g;
  bo_golomb_init(&g);
  bo_bitwriter bw = {out, cap_bytes * 8, 0};
  if (out) memset(out, 0, cap_bytes);
  uint64_t count = 0, run = 0;
  for (uint64_t i = 0; i < rows; ++i)
    for (uint64_t j = 0; j < cols; ++j) {
      if (get_bit(M + i * wpr, j)) {
        bo_golomb_code_sample(&g, out ? &bw : NULL, (uint32_t)run);
        count++;
        run = 0;
      } else {
        run++;
      }
    }
  bo_golomb_code_sample(&g, out ? &bw : NULL, (uint32_t)run);
  count++;
  if (nsamples) *nsamples = count;
  return (uint64_t)g.bitcount;
}

/* The serial coder started in the middle of a stream: the rows of M are a contiguous block of a bigger matrix.
 * State of GolombCoder (Golomb.h:12-29) after the `ones_before` samples that precede the block: samples = ones_before,
 * accumulatedError = sum of those samples = (last_one_before + 1) - ones_before (both uint32, wrapping as the reference's
 * unsigned members do), k re-derived by GolombCoder.cpp:33 (k = 1 before the first sample, Golomb.h:18). The run in
 * progress at the block's first bit is bits_before - (last_one_before + 1) zeros long. Codes the block's ones; with
 * `closing` also the run closed by the virtual one at total_bits. Returns the bits written; out may be NULL. */
uint64_t bo_golomb_encode_shard(const bo_word* M, uint64_t rows, uint64_t cols, uint64_t ones_before, uint64_t bits_before,
                                int64_t last_one_before, int closing, uint64_t total_bits, uint8_t* out, uint64_t cap_bytes,
                                uint64_t* nsamples) {
  const uint64_t wpr = bo_wpr(cols);
  bo_golomb g;
  bo_golomb_init(&g);
  if (ones_before) {
    g.samples = (uint32_t)ones_before;
    g.accumulatedError = (uint32_t)((uint64_t)(last_one_before + 1) - ones_before);
    uint32_t k;
    for (k = 0; k < 31 && (uint32_t)(g.samples << k) < g.accumulatedError; k++) {} /* GolombCoder.cpp:33 */
    g.k = k;
  }
  bo_bitwriter bw = {out, cap_bytes * 8, 0};
  if (out) memset(out, 0, cap_bytes);
  uint64_t count = 0, run = bits_before - (uint64_t)(last_one_before + 1);
  for (uint64_t i = 0; i < rows; ++i)
    for (uint64_t j = 0; j < cols; ++j) {
      if (get_bit(M + i * wpr, j)) {
        bo_golomb_code_sample(&g, out ? &bw : NULL, (uint32_t)run);
        count++;
        run = 0;
      } else {
        run++;
      }
    }
  if (closing) {
    run += total_bits - (bits_before + rows * cols);
    bo_golomb_code_sample(&g, out ? &bw : NULL, (uint32_t)run);
    count++;
  }
  if (nsamples) *nsamples = count;
  return (uint64_t)g.bitcount;
}

int bo_golomb_decode_matrix(const uint8_t* in, uint64_t nbits_in, uint64_t rows, uint64_t cols, bo_word* M) {
  const uint64_t wpr = bo_wpr(cols), N = rows * cols;
  memset(M, 0, sizeof(bo_word) * rows * wpr);
  bo_golomb g;
  bo_golomb_init(&g);
  bo_bitreader br = {in, nbits_in, 0};
  uint64_t pos = 0; /* next undecided bit of the row-major stream */
  while (pos <= N) {
    if (br.pos >= br.nbits) return -1;
    const uint64_t x = bo_golomb_decode_sample(&g, &br);
    pos += x;
    if (pos > N) return -2;
    if (pos == N) break; /* the virtual terminating one */
    put_bit(M + (pos / cols) * wpr, pos % cols, 1);
    pos++;
  }
  return br.pos == nbits_in ? 0 : -3;
}

/* ------------------------------------------------------------------------------------------
 * EG. src/eg.h:6-27, src/eg.cpp:2-37 (coder), :41-55 (decoder sketch, `#if 0` there).
 * ------------------------------------------------------------------------------------------ */
static const short BO_EGLUT[32] = {0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3,
                                   4, 4, 5, 5, 6, 6, 7, 7, 8, 9, 10, 11, 12, 13, 14, 15}; /* eg.cpp:2 */

void bo_eg_init(bo_eg* e) { e->g = 1; e->blockSize = 1; e->lutIndex = 0; e->bitcount = 0; } /* eg.h:9 */

static void eg_dec_block(bo_eg* e) { /* eg.cpp:12-18 */
  if (e->lutIndex > 0) e->lutIndex--;
  e->g = (uint32_t)BO_EGLUT[e->lutIndex];
  e->blockSize = 1u << e->g;
}

void bo_eg_code_run(bo_eg* e, bo_bitwriter* bw, int len, int eol) {
  /* eg.cpp:22 compares int len with unsigned blockSize: len is converted to unsigned */
  while ((unsigned)len >= e->blockSize) {               /* :22-27 (incBlockSize disabled, :25) */
    len -= (int)e->blockSize;
    if (bw) bw_put_bit(bw, 1);                          /* :24 */
    e->bitcount++;
  }
  if (eol) {                                            /* :28-30 */
    if (bw) bw_put_bit(bw, 1);
    e->bitcount++;
  } else {                                              /* :31-36 */
    if (bw) { bw_put_bit(bw, 0); bw_put_bits(bw, (uint32_t)len, e->g); }
    e->bitcount += e->g + 1;
    eg_dec_block(e);
  }
}

uint64_t bo_eg_encode_matrix(const bo_word* M, uint64_t rows, uint64_t cols, uint8_t* out, uint64_t cap_bytes) {
  const uint64_t wpr = bo_wpr(cols);
  bo_eg e;
  bo_eg_init(&e);
  bo_bitwriter bw = {out, cap_bytes * 8, 0};
  if (out) memset(out, 0, cap_bytes);
  for (uint64_t i = 0; i < rows; ++i) {
    int run = 0;
    for (uint64_t j = 0; j < cols; ++j) {
      if (get_bit(M + i * wpr, j)) { bo_eg_code_run(&e, out ? &bw : NULL, run, 0); run = 0; }
      else run++;
    }
    bo_eg_code_run(&e, out ? &bw : NULL, run, 1);
  }
  return e.bitcount;
}

int bo_eg_decode_matrix(const uint8_t* in, uint64_t nbits_in, uint64_t rows, uint64_t cols, bo_word* M) {
  const uint64_t wpr = bo_wpr(cols);
  memset(M, 0, sizeof(bo_word) * rows * wpr);
  bo_eg e;
  bo_eg_init(&e);
  bo_bitreader br = {in, nbits_in, 0};
  for (uint64_t i = 0; i < rows; ++i) {
    uint64_t col = 0;
    for (;;) {
      const uint64_t maxlen = cols - col;
      uint64_t len = 0;
      int eol = 0;
      /* eg.cpp:44-50: a one per full block; running past maxlen means the row ended */
      while (br_get_bit(&br)) {
        len += e.blockSize;
        if (len > maxlen) { eol = 1; break; }
        if (br.pos > br.nbits) return -1;
      }
      if (eol) break;
      len += br_get_bits(&br, e.g);                     /* :52 */
      eg_dec_block(&e);                                 /* :53 */
      col += len;
      if (col >= cols) return -2;
      put_bit(M + i * wpr, col, 1);
      col++;
    }
  }
  return br.pos == nbits_in ? 0 : -3;
}

/* ------------------------------------------------------------------------------------------
 * MDL model selection: model_codelength src/bsvd.cpp:1438-1461 over
 * universal_codelength src/coding.cpp:24-32; learn_model_mdl_forward_selection :1463-1546,
 * learn_model_mdl_backward_selection :1548-1660, learn_model_mdl_full_search :1662-1717.
 * The inner learner is learn_model_traditional and the initialiser initialize_model_neighbor
 * (the reference's defaults, learn_model_setup(0,0,0,*,0)). The only floating point on the
 * path: double log2 on the host, truncated into idx_t at every accumulation, as the reference
 * does it.
 * ------------------------------------------------------------------------------------------ */
#include <math.h>

/* src/coding.cpp:24-32; the parameters are `unsigned` there, so 64-bit counts wrap to 32 bits */
double bo_universal_codelength(unsigned n, unsigned r) {
  const double p1 = (double)r / (double)n;
  if ((r > 0) && (r < n)) {
    return (double)n * (-p1 * log2(p1) - (1.0 - p1) * log2(1.0 - p1)) + 0.5 * log2(n);
  } else {
    return 0.5 * log2(n);
  }
}

static uint64_t col_weight(const bo_word* A, uint64_t n, uint64_t p, uint64_t k) {
  const uint64_t apr = bo_wpr(p);
  uint64_t w = 0;
  for (uint64_t i = 0; i < n; ++i) w += (A[i * apr + (k >> 6)] >> (63 - (k & 63))) & 1u;
  return w;
}

/* src/bsvd.cpp:1438-1461. LE, LD, LA are idx_t: `LD += double` converts the sum back to an
 * integer (truncation) after every atom. */
uint64_t bo_model_codelength(const bo_word* E, const bo_word* D, const bo_word* A,
                             uint64_t n, uint64_t m, uint64_t p) {
  const uint64_t wpr = bo_wpr(m);
  uint64_t LE = (uint64_t)bo_universal_codelength((unsigned)(n * m), (unsigned)bo_weight(E, n, m));  /* :1449 */
  uint64_t LD = 0, LA = 0;
  for (uint64_t k = 0; k < p; ++k) {                                                                  /* :1451-1456 */
    LD = (uint64_t)((double)LD + bo_universal_codelength((unsigned)m, (unsigned)bo_weight(D + k * wpr, 1, m)));
    LA = (uint64_t)((double)LA + bo_universal_codelength((unsigned)n, (unsigned)col_weight(A, n, p, k)));
  }
  return LE + LD + LA;
}

struct bo_mdl_result { uint64_t p, n, m, bestL; bo_word *D, *A; };

static bo_word* mat_alloc(uint64_t rows, uint64_t cols) {
  const uint64_t w = rows * bo_wpr(cols);
  return (bo_word*)calloc(w ? w : 1, sizeof(bo_word));
}
static bo_word* mat_dup(const bo_word* M, uint64_t rows, uint64_t cols) {
  bo_word* r = mat_alloc(rows, cols);
  memcpy(r, M, rows * bo_wpr(cols) * sizeof(bo_word));
  return r;
}
static int mget(const bo_word* M, uint64_t cols, uint64_t i, uint64_t j) {
  return (int)((M[i * bo_wpr(cols) + (j >> 6)] >> (63 - (j & 63))) & 1u);
}
static void mset(bo_word* M, uint64_t cols, uint64_t i, uint64_t j, int b) {
  const bo_word mask = (bo_word)1 << (63 - (j & 63));
  bo_word* w = M + i * bo_wpr(cols) + (j >> 6);
  *w = b ? (*w | mask) : (*w & ~mask);
}
/* A (n x p) -> n x (p+1) with `col` (n x 1) appended (set_submatrix(0,0,A); set_submatrix(0,p,col), :1513-1516) */
static bo_word* append_col(const bo_word* A, uint64_t n, uint64_t p, const bo_word* col) {
  bo_word* R = mat_alloc(n, p + 1);
  for (uint64_t i = 0; i < n; ++i) {
    for (uint64_t j = 0; j < p; ++j) if (mget(A, p, i, j)) mset(R, p + 1, i, j, 1);
    if (mget(col, 1, i, 0)) mset(R, p + 1, i, p, 1);
  }
  return R;
}
/* column k removed (:1605-1616) */
static bo_word* delete_col(const bo_word* A, uint64_t n, uint64_t p, uint64_t k) {
  bo_word* R = mat_alloc(n, p - 1);
  for (uint64_t i = 0; i < n; ++i)
    for (uint64_t j = 0, o = 0; j < p; ++j) {
      if (j == k) continue;
      if (mget(A, p, i, j)) mset(R, p - 1, i, o, 1);
      ++o;
    }
  return R;
}

static struct bo_mdl_result* mdl_result(uint64_t n, uint64_t m, uint64_t p, uint64_t bestL, bo_word* D, bo_word* A) {
  struct bo_mdl_result* r = (struct bo_mdl_result*)malloc(sizeof(*r));
  r->p = p; r->n = n; r->m = m; r->bestL = bestL; r->D = D; r->A = A;
  return r;
}
void bo_mdl_result_info(const struct bo_mdl_result* r, uint64_t* p, uint64_t* bestL) { *p = r->p; *bestL = r->bestL; }
void bo_mdl_result_copy(const struct bo_mdl_result* r, bo_word* D, bo_word* A) {
  if (r->p) {
    memcpy(D, r->D, r->p * bo_wpr(r->m) * sizeof(bo_word));
    memcpy(A, r->A, r->n * bo_wpr(r->p) * sizeof(bo_word));
  }
}
void bo_mdl_result_free(struct bo_mdl_result* r) { free(r->D); free(r->A); free(r); }

/* learn_model_mdl_forward_selection, src/bsvd.cpp:1463-1546. D, A: the initialised model with p atoms (left
 * untouched; the result carries the selected model); E receives the residual of the selected model. */
struct bo_mdl_result* bo_learn_mdl_forward(const bo_word* X, bo_word* E, const bo_word* D_in, const bo_word* A_in,
                                           uint64_t n, uint64_t m, uint64_t p, bo_rand48* rng) {
  const uint64_t wpr = bo_wpr(m);
  uint64_t K = p;
  bo_word* D = mat_dup(D_in, K, m);
  bo_word* A = mat_dup(A_in, n, K);
  bo_learn_traditional(X, E, D, A, n, m, K, NULL, 0);                 /* :1470 */
  bo_word* nextAtom = mat_alloc(1, m);
  bo_word* nextCoefs = mat_alloc(n, 1);
  bo_word *currD = mat_dup(D, K, m), *currA = mat_dup(A, n, K), *currE = mat_dup(E, n, m);  /* :1473 */
  uint64_t bestK = K;
  uint64_t bestL = bo_model_codelength(E, D, A, n, m, K);             /* :1476 */
  uint64_t stuck = 0, sumStuck = 0, allStuck = 0;
  do {
    const int dev = allStuck > 0 ? (int)(sumStuck / allStuck) : 0;    /* :1486 */
    bo_init_neighbor(currE, n, m, 1, rng, nextAtom, nextCoefs);       /* initialize_model(currE,nextAtom,nextCoefs), :1488 */
    bo_word* nD = mat_alloc(K + 1, m);                                /* :1499-1506 */
    memcpy(nD, currD, K * wpr * sizeof(bo_word));
    memcpy(nD + K * wpr, nextAtom, wpr * sizeof(bo_word));
    free(currD);
    currD = nD;
    bo_word* nA = append_col(currA, n, K, nextCoefs);                 /* :1508-1516 */
    free(currA);
    currA = nA;
    bo_learn_traditional(X, currE, currD, currA, n, m, K + 1, NULL, 0);  /* :1518 */
    const uint64_t currL = bo_model_codelength(currE, currD, currA, n, m, K + 1);
    if ((currL + (uint64_t)(int64_t)dev) < bestL) {                   /* :1520 */
      stuck = 0;
      bestL = currL;
      free(D); free(A);
      D = mat_dup(currD, K + 1, m);
      A = mat_dup(currA, n, K + 1);
      memcpy(E, currE, n * wpr * sizeof(bo_word));
      bestK = K + 1;
    } else {
      stuck++;
      allStuck++;
      sumStuck += (currL - bestL);
      if (stuck >= 10) break;                                         /* :1532-1535 */
    }
    K++;
  } while (stuck < 10);
  free(currD); free(currA); free(currE); free(nextAtom); free(nextCoefs);
  return mdl_result(n, m, bestK, bestL, D, A);
}

/* learn_model_mdl_backward_selection, src/bsvd.cpp:1548-1660 */
struct bo_mdl_result* bo_learn_mdl_backward(const bo_word* X, bo_word* E, const bo_word* D_in, const bo_word* A_in,
                                            uint64_t n, uint64_t m, uint64_t p) {
  const uint64_t wpr = bo_wpr(m);
  uint64_t K = p;
  bo_word* D = mat_dup(D_in, K, m);
  bo_word* A = mat_dup(A_in, n, K);
  uint64_t outK = K;                                                  /* atoms of (D, A) as handed back */
  bo_learn_traditional(X, E, D, A, n, m, K, NULL, 0);                 /* :1555 */
  uint64_t bestL = bo_model_codelength(E, D, A, n, m, K);
  uint64_t currL = bestL;
  bo_word *currD = mat_dup(D, K, m), *currA = mat_dup(A, n, K);
  bo_word *nextD = NULL, *nextA = NULL;
  bo_word* nextE = mat_alloc(n, m);
  uint64_t stuck = 0, sumStuck = 0, allStuck = 0;
  (void)currL;
  for (; K > 0; K--) {
    const int dev = allStuck > 0 ? (int)(sumStuck / allStuck) : 0;    /* :1575 */
    uint64_t nextk = 0;
    uint64_t nextL = ~(1UL << (sizeof(uint64_t) - 1));                /* :1578 */
    for (uint64_t k = 0; k < K; k++) {                                /* :1579-1592 */
      const bo_word* Dk = currD + k * wpr;
      for (uint64_t i = 0; i < n; ++i) {                              /* nextE = Ak' * Dk xor E */
        const int a = mget(currA, K, i, k);
        for (uint64_t b = 0; b < wpr; ++b) nextE[i * wpr + b] = E[i * wpr + b] ^ (a ? Dk[b] : 0);
      }
      uint64_t tmpL = bo_model_codelength(nextE, currD, currA, n, m, K);
      tmpL = (uint64_t)((double)tmpL - bo_universal_codelength((unsigned)m, (unsigned)bo_weight(Dk, 1, m)));
      tmpL = (uint64_t)((double)tmpL - bo_universal_codelength((unsigned)n, (unsigned)col_weight(currA, n, K, k)));
      if (tmpL < nextL) { nextL = tmpL; nextk = k; }
    }
    free(nextD); free(nextA);
    nextD = NULL; nextA = NULL;
    if (K > 1) {                                                      /* :1598-1616 */
      nextD = mat_alloc(K - 1, m);
      for (uint64_t k = 0, o = 0; k < K; ++k) {
        if (k == nextk) continue;
        memcpy(nextD + o * wpr, currD + k * wpr, wpr * sizeof(bo_word));
        ++o;
      }
      nextA = delete_col(currA, n, K, nextk);
      bo_learn_traditional(X, nextE, nextD, nextA, n, m, K - 1, NULL, 0);
      nextL = bo_model_codelength(nextE, nextD, nextA, n, m, K - 1);
    } else {
      nextL = bo_model_codelength(nextE, NULL, NULL, n, m, 0);        /* :1618, destroyed (0 x 0) D and A */
    }
    if (nextL + (uint64_t)(int64_t)dev < bestL) {                     /* :1621 */
      if (K == 1) {                                                   /* "Resulted in empty model!", :1623-1629 */
        free(D); free(A);
        D = NULL; A = NULL;
        outK = 0;
        memcpy(E, X, n * wpr * sizeof(bo_word));
        break;
      }
      stuck = 0;
      bestL = nextL;
      free(D); free(A);
      D = mat_dup(nextD, K - 1, m);
      A = mat_dup(nextA, n, K - 1);
      outK = K - 1;
      memcpy(E, nextE, n * wpr * sizeof(bo_word));
    } else {
      stuck++;
      allStuck++;
      sumStuck += (nextL - bestL);
      if (stuck >= 10) break;
    }
    free(currD); free(currA);                                         /* :1649-1655 */
    currD = (K > 1) ? mat_dup(nextD, K - 1, m) : mat_alloc(0, m);
    currA = (K > 1) ? mat_dup(nextA, n, K - 1) : mat_alloc(n, 0);
    currL = nextL;
  }
  free(currD); free(currA); free(nextD); free(nextA); free(nextE);
  return mdl_result(n, m, outK, bestL, D, A);
}

/* learn_model_mdl_full_search, src/bsvd.cpp:1662-1717: dictionary sizes 20, 40, ... <= Kmax, eleven fits each (the
 * RNG stream simply continues: random_seed is rewritten at :1680 but the generator was seeded on first use). The
 * matrices kept for a size are those of its LAST fit, the length recorded is the minimum over the last ten. */
struct bo_mdl_result* bo_learn_mdl_full_search(const bo_word* X, bo_word* E, uint64_t n, uint64_t m, uint64_t Kmax,
                                               bo_rand48* rng) {
  const uint64_t wpr = bo_wpr(m);
  bo_word* candE = mat_alloc(n, m);
  uint64_t bestL = 1UL << 30, bestk = 0;
  bo_word *D = NULL, *A = NULL;
  for (uint64_t k = 20; k <= Kmax; k += 20) {
    bo_word *candD = mat_alloc(k, m), *candA = mat_alloc(n, k);
    bo_init_neighbor(X, n, m, k, rng, candD, candA);
    bo_learn_traditional(X, candE, candD, candA, n, m, k, NULL, 0);
    uint64_t candL = ~0ull;
    for (int rep = 0; rep < 10; ++rep) {
      bo_init_neighbor(X, n, m, k, rng, candD, candA);
      bo_learn_traditional(X, candE, candD, candA, n, m, k, NULL, 0);
      const uint64_t L = bo_model_codelength(candE, candD, candA, n, m, k);
      if (L < candL) candL = L;
    }
    if (candL < bestL) {
      bestL = candL;
      bestk = k;
      memcpy(E, candE, n * wpr * sizeof(bo_word));
      free(D); free(A);
      D = candD; A = candA;
    } else {
      free(candD); free(candA);
    }
  }
  free(candE);
  return mdl_result(n, m, bestk, bestL, D, A);
}

/* ------------------------------------------------------------------------------------------
 * Bit planes of a grey image: the loop of bitplane_tool (src/bitplane_tool.cpp:24-39) over pixels read as
 * read_pgm_p5_data reads them (src/pnm.cpp:54-78): one byte per pixel when maxval < 256, else two, high byte
 * first. planes: nplanes consecutive rows x cols matrices, plane bi for mask 1 << bi, for every mask < maxval.
 * Returns the number of planes.
 * ------------------------------------------------------------------------------------------ */
uint32_t bo_split_bitplanes(const uint8_t* payload, uint64_t rows, uint64_t cols, uint32_t maxval, bo_word* planes) {
  const uint64_t wpr = bo_wpr(cols);
  uint32_t bi = 0;
  for (uint64_t b = 1; b < maxval; b <<= 1, bi++) {           /* :24 */
    bo_word* A = planes + (uint64_t)bi * rows * wpr;
    memset(A, 0, rows * wpr * sizeof(bo_word));
    for (uint64_t i = 0, li = 0; i < rows; i++)
      for (uint64_t j = 0; j < cols; j++, li++) {
        const uint32_t pix = maxval < 256 ? payload[li] : (((uint32_t)payload[2 * li] << 8) + payload[2 * li + 1]);  /* pnm.cpp:62,71 */
        if (pix & b) A[i * wpr + (j >> 6)] |= (bo_word)1 << (63 - (j & 63));   /* A.set(i,j,gray_img[li] & b), :28 */
      }
  }
  return bi;
}

/* ------------------------------------------------------------------------------------------
 * Role-switched learners: learn_model_alter1 src/bsvd.cpp:1245-1311, learn_model_alter2
 * :1314-1388, learn_model_alter3 :1391-1434, over binary_matrix::transpose_to (src/binmat.cpp:199-208) and the
 * default update_coefficients / update_dictionary plug points. In the transposed calls the roles are
 * E' = Et (m x n), D' = At (p x n), A' = Dt (m x p).
 * ------------------------------------------------------------------------------------------ */
void bo_transpose(const bo_word* M, uint64_t rows, uint64_t cols, bo_word* T) {
  const uint64_t wi = bo_wpr(cols), wo = bo_wpr(rows);
  memset(T, 0, cols * wo * sizeof(bo_word));
  for (uint64_t i = 0; i < rows; ++i)
    for (uint64_t j = 0; j < cols; ++j)
      if ((M[i * wi + (j >> 6)] >> (63 - (j & 63))) & 1u) T[j * wo + (i >> 6)] |= (bo_word)1 << (63 - (i & 63));
}

uint64_t bo_learn_alter(int variant, const bo_word* X, bo_word* E, bo_word* D, bo_word* A, uint64_t n, uint64_t m, uint64_t p) {
  bo_residual(X, A, D, E, n, m, p);                                   /* :1254-1255 / :1323-1324 / :1399-1400 */
  bo_word* Dt = mat_alloc(m, p);
  bo_word* At = mat_alloc(p, n);
  bo_word* Et = mat_alloc(m, n);
  uint64_t iter = 0;
  if (variant == 1) {                                                 /* :1264-1307 */
    uint64_t changed = 1;
    while (changed > 0) {
      iter++;
      uint64_t cc = bo_update_coefficients(E, D, A, n, m, p);
      changed = cc + bo_update_dictionary(E, D, A, n, m, p);
      bo_transpose(A, n, p, At); bo_transpose(D, p, m, Dt); bo_transpose(E, n, m, Et);
      cc = bo_update_coefficients(Et, At, Dt, m, n, p);
      (void)cc;
      changed = bo_update_dictionary(Et, At, Dt, m, n, p);            /* :1297: only this count drives the loop */
      bo_transpose(At, p, n, A); bo_transpose(Dt, m, p, D); bo_transpose(Et, m, n, E);
    }
  } else if (variant == 2) {                                          /* :1331-1383 */
    uint64_t changed = 1, outer_changed = 1;
    while (outer_changed > 0) {
      outer_changed = 0;
      while (changed > 0) {
        iter++;
        const uint64_t cc = bo_update_coefficients(E, D, A, n, m, p);
        changed = cc + bo_update_dictionary(E, D, A, n, m, p);
        outer_changed += changed;
      }
      bo_transpose(A, n, p, At); bo_transpose(D, p, m, Dt); bo_transpose(E, n, m, Et);
      changed = 1;
      iter = 0;                                                       /* :1361 */
      while (changed > 0) {
        iter++;
        const uint64_t cc = bo_update_coefficients(Et, At, Dt, m, n, p);
        changed = cc + bo_update_dictionary(Et, At, Dt, m, n, p);
        outer_changed += changed;
      }
      bo_transpose(At, p, n, A); bo_transpose(Dt, m, p, D); bo_transpose(Et, m, n, E);
    }
  } else {                                                            /* :1407-1429 */
    uint64_t changed = p + 1;
    while (changed > 0) {
      iter++;
      bo_transpose(A, n, p, At); bo_transpose(D, p, m, Dt); bo_transpose(E, n, m, Et);
      changed = bo_update_dictionary(Et, At, Dt, m, n, p);
      bo_transpose(At, p, n, A); bo_transpose(Dt, m, p, D); bo_transpose(Et, m, n, E);
      changed = bo_update_dictionary(E, D, A, n, m, p);               /* :1423: overwrites the transposed count */
    }
  }
  free(Dt); free(At); free(Et);
  return iter;
}

/* ------------------------------------------------------------------------------------------
 * update_dictionary_proximus, src/bsvd.cpp:528-729 (the `#if 0` initialisation block :563-620 is dead code).
 * Per atom, in order: alternate a majority vote for the atom over its users (as the steepest update does) and a
 * majority vote for the atom's coefficient column over the atom's set bits, patching E after each, until neither
 * changes. Counts the atoms whose D row changed (:655, :711 keeps coefficient-only changes out of the count).
 * ------------------------------------------------------------------------------------------ */
uint64_t bo_update_dictionary_proximus(bo_word* E, bo_word* D, bo_word* A, uint64_t n, uint64_t m, uint64_t p) {
  const uint64_t wpr = bo_wpr(m), apr = bo_wpr(p);
  uint64_t changed = 0;
  uint64_t* Dw = (uint64_t*)malloc(sizeof(uint64_t) * (m ? m : 1));
  bo_word* newDk = (bo_word*)malloc(sizeof(bo_word) * (wpr ? wpr : 1));
  for (uint64_t k = 0; k < p; k++) {
    bo_word* Dk = D + k * wpr;
    int kchanged = 0, converged;
    do {
      converged = 1;
      /* ---- the atom: :627-668 */
      uint64_t u = 0;
      memset(Dw, 0, sizeof(uint64_t) * m);
      for (uint64_t i = 0; i < n; i++) {
        if (!mget(A, p, i, k)) continue;
        u++;
        for (uint64_t j = 0; j < m; j++)
          if (mget(E, m, i, j) ^ mget(Dk, m, 0, j)) Dw[j]++;          /* add-back old atom, :637-643 */
      }
      if (u) {
        u /= 2;                                                       /* :649 */
        memcpy(newDk, Dk, sizeof(bo_word) * wpr);
        for (uint64_t j = 0; j < m; j++) mset(newDk, m, 0, j, Dw[j] > u);
        int dd = 0;
        for (uint64_t b = 0; b < wpr; ++b) dd |= (newDk[b] != Dk[b]);
        if (dd) {
          for (uint64_t i = 0; i < n; i++) {                          /* :659-666 */
            if (!mget(A, p, i, k)) continue;
            for (uint64_t b = 0; b < wpr; ++b) E[i * wpr + b] ^= Dk[b] ^ newDk[b];
          }
          memcpy(Dk, newDk, sizeof(bo_word) * wpr);                   /* :655 */
          converged = 0;
          kchanged = 1;
        }
      }
      /* ---- the coefficient column: :673-715. u = bits of the (updated) atom; Aw[i] = sum over them of E[i][j] xor A[i][k] */
      u = 0;
      for (uint64_t j = 0; j < m; j++) u += (uint64_t)mget(Dk, m, 0, j);
      if (u) {
        const uint64_t half = u / 2;
        for (uint64_t i = 0; i < n; i++) {
          const int a = mget(A, p, i, k);
          uint64_t Aw = 0;
          for (uint64_t j = 0; j < m; j++)
            if (mget(Dk, m, 0, j) && (mget(E, m, i, j) ^ a)) Aw++;
          const int na = Aw > half;
          if (na != a) {                                              /* rows are independent: each patches its own E row */
            mset(A, p, i, k, na);
            for (uint64_t b = 0; b < wpr; ++b) E[i * wpr + b] ^= Dk[b];
            converged = 0;
          }
        }
      }
    } while (!converged);
    if (kchanged) changed++;
  }
  free(Dw); free(newDk);
  (void)apr;
  return changed;
}


/* ------------------------------------------------------------------------------------------
 * compress*_test: template matching with enumerative + Golomb costing.
 * src/compress_test.cpp:37-141 (v1) and src/compress4_test.cpp:35-171 (v4).
 * ------------------------------------------------------------------------------------------ */
double bo_enumL(uint64_t n, uint64_t r) { /* compress_test.cpp:37-40; lnchoose as oracle/gsl_shim/gsl/gsl_sf_gamma.h */
  if (r == 0 || r >= n) return 0.0;
  const double ln = lgamma((double)n + 1.0) - lgamma((double)r + 1.0) - lgamma((double)(n - r) + 1.0);
  return ln * 1.442695040888963387004650940070860087872;
}

/* W bits (W <= 64, left aligned in the result) of image row r starting at column j, as get_submatrix sees them
 * (binmat.cpp:267-298): the words of the matrix are one flat array, rows are ceil(cols/64) blocks long with zero pad bits
 * (read_pbm_data clears the matrix, pbm.cpp:31), a read that runs past the last block of a row continues in the next row,
 * and blocks past the end of the array read as zero. */
static bo_word flat_bits(const bo_word* I, uint64_t rows, uint64_t cols, uint64_t r, uint64_t j, uint64_t W) {
  const uint64_t bpr = bo_wpr(cols), nblk = rows * bpr;
  const uint64_t k1 = r * bpr + j / 64, off = j % 64;
  const bo_word s1 = k1 < nblk ? I[k1] : 0, s2 = (k1 + 1) < nblk ? I[k1 + 1] : 0;
  const bo_word v = off ? ((s1 << off) | (s2 >> (64 - off))) : s1;
  return W >= 64 ? v : (v & ~(~(bo_word)0 >> W));  /* get_block's trail mask of the W-column patch, binmat.h:188-190 */
}

static uint64_t patch_dist(const bo_word* I, uint64_t rows, uint64_t cols, const bo_word* P, uint64_t i2, uint64_t j2, uint64_t W) {
  uint64_t d = 0;                               /* dist(P, P2), binmat.cpp:499-512 */
  for (uint64_t di = 0; di < W; ++di) d += (uint64_t)popc64(P[di] ^ flat_bits(I, rows, cols, i2 + di, j2, W));
  return d;
}

static uint64_t ceil_log2_u64(uint64_t li) {    /* ceil(log2(li)) for li >= 1 */
  uint64_t k = 0;
  while (((uint64_t)1 << k) < li) k++;
  return k;
}

static void match_account(bo_match_rec* rec, bo_golomb* gm, bo_golomb* gn, bo_match_totals* tot) {
  if (rec->use_match) {                         /* compress_test.cpp:130-136 */
    bo_golomb_code_sample(gm, NULL, (uint32_t)rec->bestd);
    tot->weight_sum += rec->bestd;
    tot->matches++;
    tot->L += (double)rec->match_len;
  } else {                                      /* :137-140 */
    bo_golomb_code_sample(gn, NULL, (uint32_t)rec->weight);
    tot->L += (double)rec->nomatch_len;
  }
}

void bo_compress_v1(const bo_word* I, uint64_t rows, uint64_t cols, uint64_t W, bo_match_rec* recs, bo_match_totals* tot) {
  const uint64_t Ny = (W - 1 + rows) / W, Nx = (W - 1 + cols) / W, M = W * W;
  bo_golomb gm, gn;
  bo_golomb_init(&gm); bo_golomb_init(&gn);
  memset(tot, 0, sizeof(*tot));
  bo_word P[64];
  uint64_t li = 0;
  for (uint64_t i = 0; i < Ny; ++i)
    for (uint64_t j = 0; j < Nx; ++j, ++li) {
      const uint64_t i0 = i * W, j0 = j * W;
      uint64_t w = 0;
      for (uint64_t di = 0; di < W; ++di) { P[di] = flat_bits(I, rows, cols, i0 + di, j0, W); w += (uint64_t)popc64(P[di]); }
      uint64_t besti = 0, bestj = 0, bestd = M;  /* :78 */
      int perfect = 0;
      int64_t i2 = 0;
      for (; i2 <= (int64_t)i0 - (int64_t)W && !perfect; ++i2)       /* :81-96 rows fully above: every column */
        for (uint64_t j2 = 0; j2 < cols; ++j2) {
          const uint64_t d = patch_dist(I, rows, cols, P, (uint64_t)i2, j2, W);
          if (d < bestd) { bestd = d; besti = (uint64_t)i2; bestj = j2; }
          if (bestd == 0) { perfect = 1; break; }
        }
      for (; i2 <= (int64_t)i0 && !perfect; ++i2)                     /* :97-111 the patch's own band: columns to its left */
        for (int64_t j2 = 0; j2 <= (int64_t)j0 - (int64_t)W; ++j2) {
          const uint64_t d = patch_dist(I, rows, cols, P, (uint64_t)i2, (uint64_t)j2, W);
          if (d < bestd) { bestd = d; besti = (uint64_t)i2; bestj = (uint64_t)j2; }
          if (bestd == 0) { perfect = 1; break; }
        }
      bo_match_rec* rec = recs + li;
      rec->besti = besti; rec->bestj = bestj; rec->bestd = bestd; rec->weight = w;
      rec->nomatch_len = (uint64_t)(1 + bo_enumL(M, w));              /* :126 */
      if (li == 0) {
        /* ceil(log2(0)) = -inf converted to idx_t is undefined; on x86-64 it comes out as 2^63, so the first patch
         * (which has no candidate anyway) never takes the match branch */
        rec->match_len = (uint64_t)1 << 63;
        rec->use_match = 0;
      } else {
        rec->match_len = (uint64_t)(1 + ceil_log2_u64(li) + bo_enumL(M, bestd));  /* :120,127 */
        rec->use_match = rec->nomatch_len > rec->match_len;            /* :129 */
      }
      match_account(rec, &gm, &gn, tot);
    }
  tot->bits_match = (uint64_t)gm.bitcount;
  tot->bits_nomatch = (uint64_t)gn.bitcount;
}

void bo_compress_v4(bo_word* I, uint64_t rows, uint64_t cols, uint64_t W, uint64_t T, uint64_t R, bo_match_rec* recs,
                    bo_match_totals* tot) {
  const uint64_t Ny = (W - 1 + rows) / W, Nx = (W - 1 + cols) / W, M = W * W, bpr = bo_wpr(cols);
  bo_golomb gm, gn;
  bo_golomb_init(&gm); bo_golomb_init(&gn);
  memset(tot, 0, sizeof(*tot));
  bo_word P[64];
  uint64_t li = 0;
  for (uint64_t i = 0; i < Ny; ++i)
    for (uint64_t j = 0; j < Nx; ++j, ++li) {
      const int64_t i0 = (int64_t)(i * W), j0 = (int64_t)(j * W), Wi = (int64_t)W, Ri = (int64_t)R;
      uint64_t w = 0;
      for (uint64_t di = 0; di < W; ++di) { P[di] = flat_bits(I, rows, cols, (uint64_t)i0 + di, (uint64_t)j0, W); w += (uint64_t)popc64(P[di]); }
      uint64_t besti = 0, bestj = 0, bestd = M + 1;                  /* compress4_test.cpp:94 */
      int perfect = 0;
      const int64_t mini = i0 > Ri ? i0 - Ri : 0, mini2 = i0 > Wi ? i0 - Wi : 0;     /* :97-98 */
      const int64_t minj = j0 > Ri ? j0 - Ri : 0;
      const int64_t maxj = (j0 + Ri) > ((int64_t)cols - Wi) ? (int64_t)cols - Wi : j0 + Ri;
      for (int64_t i2 = i0; i2 >= mini2 && !perfect; --i2)            /* :103-119 behind the patch, backwards */
        for (int64_t j2 = j0 - Wi; j2 >= minj; --j2) {
          const uint64_t d = patch_dist(I, rows, cols, P, (uint64_t)i2, (uint64_t)j2, W);
          if (d < bestd) { bestd = d; besti = (uint64_t)i2; bestj = (uint64_t)j2; }
          if (bestd <= T) { perfect = 1; break; }
        }
      for (int64_t i2 = i0 - Wi; i2 >= mini && !perfect; --i2)        /* :120-135 everything above, backwards */
        for (int64_t j2 = maxj; j2 >= minj; --j2) {
          const uint64_t d = patch_dist(I, rows, cols, P, (uint64_t)i2, (uint64_t)j2, W);
          if (d < bestd) { bestd = d; besti = (uint64_t)i2; bestj = (uint64_t)j2; }
          if (bestd <= T) { perfect = 1; break; }
        }
      bo_match_rec* rec = recs + li;
      rec->besti = besti; rec->bestj = bestj; rec->bestd = bestd; rec->weight = w;
      rec->nomatch_len = (uint64_t)(1 + bo_enumL(M, w));              /* :153 */
      /* :154; with bestd > W*W (no candidate: only the first patch) the constant 100000 keeps the undefined
       * ceil(log2(0)) out of the comparison */
      rec->match_len = bestd <= M ? (uint64_t)(1 + (li ? ceil_log2_u64(li) : 0) + bo_enumL(M, bestd)) : 100000;
      rec->use_match = rec->nomatch_len > rec->match_len;             /* :157 */
      match_account(rec, &gm, &gn, tot);
      if (rec->use_match) {                                           /* :164 I.set_submatrix(i0,j0,P3), P3 = P xor P2 */
        for (uint64_t di = 0; di < W && (uint64_t)i0 + di < rows; ++di) {
          const bo_word p3 = P[di] ^ flat_bits(I, rows, cols, besti + di, bestj, W);
          bo_word* blk = I + ((uint64_t)i0 + di) * bpr + (uint64_t)j0 / 64;
          const uint64_t off = (uint64_t)j0 % 64;                      /* W | 64 and W | cols: the patch sits inside one block */
          const bo_word mask = (W >= 64 ? ~(bo_word)0 : ~(~(bo_word)0 >> W)) >> off;
          *blk = (*blk & ~mask) | ((p3 >> off) & mask);
        }
      }
    }
  tot->bits_match = (uint64_t)gm.bitcount;
  tot->bits_nomatch = (uint64_t)gn.bitcount;
}
