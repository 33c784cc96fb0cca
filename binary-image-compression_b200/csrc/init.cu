// initialize_model_neighbor on the device. Reference: src/bsvd.cpp:227-267.
//
// The reference, per pivot row P (a random non-zero row of X), scans all rows j, forms
// Ej = X[j] & P, and if Ej != 0 counts u++ and s[b] += Ej[b]; then D[k][b] = (s[b] >= u/2).
// Two observations make this one pass over X instead of p bit-serial passes, with identical
// results:
//   (1) Ej[b] = 1 already implies Ej != 0, so s[b] = P[b] ? c[b] : 0 where c[b] is the plain
//       column count of X (independent of the pivot);
//   (2) only u = #{j : X[j] & P != 0} depends on the pivot.
// So: one column histogram of X, one "does row j meet pivot k" count per pivot, one threshold.
// The RNG draw (rand48 + rejection of all-zero rows, :241-243) is serial and stays on the host;
// the device only supplies the zero-row bitmap.
#include "bic_internal.cuh"

#include <vector>

// ------------------------------------------------------------------ zero-row bitmap
__global__ void k_row_nonzero(const uint32_t* __restrict__ X, uint64_t n, uint64_t wpr, uint32_t* __restrict__ bitmap) {
  const uint64_t nround = (n + 31) & ~(uint64_t)31;
  for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < nround; r += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t any = 0;
    if (r < n) {
      const uint32_t* row = X + r * wpr;
      for (uint64_t w = 0; w < wpr; ++w) any |= __ldg(row + w);
    }
    const uint32_t b = __ballot_sync(0xffffffffu, any != 0);
    if ((threadIdx.x & 31) == 0) bitmap[r >> 5] = b;  // bit (r & 31), LSB first
  }
}

bic_status bic_k_row_nonzero_bitmap(bic_ctx* c, const bic_mat* X, uint32_t* d_bitmap) {
  if (X->rows == 0) return BIC_OK;
  BIC_PROF(c, KID_ROW_NONZERO);
  k_row_nonzero<<<bic_grid_for(c, X->rows, 256, 8), 256, 0, c->stream>>>(X->d, X->rows, X->wpr, d_bitmap);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

// The draw itself on the device, for callers that must not wait for the host (pipeline.cu): one warp replays rand48 32
// steps at a time. Lane l holds the affine map of l + 1 generator steps (s -> a_l * s + c_l mod 2^48), so a round costs one
// multiply-add, one division (gsl_rng_uniform_int: k = (s >> 16) / scale, k >= n is drawn again, src/bsvd.cpp:241) and one
// bitmap lookup per lane; accepted rows are taken in lane (= draw) order until p pivots exist, and the generator is left
// exactly after the draw that produced the last one. status[0] = 1 if X has no nonzero row (the reference would never
// return), status[1] = draws of the outer loop (accepted + rejected zero rows).
// Row-sharded variant (starts != nullptr): the draw runs over the GLOBAL row index of the concatenated shards; `bitmap` holds
// every rank's zero-row bitmap, bw words each, and starts[r] is the global index of rank r's first row. Every rank replays the
// same draw; pivots[k] receives the LOCAL row index where this rank owns the row and ~0 elsewhere.
__global__ void __launch_bounds__(256) k_draw_pivots(const uint32_t* __restrict__ bitmap, uint64_t n, uint32_t p,
                                                     uint64_t* __restrict__ state, uint64_t* __restrict__ pivots,
                                                     unsigned long long* __restrict__ status, const uint64_t* __restrict__ starts = nullptr,
                                                     uint32_t nranks = 1, uint64_t bw = 0, uint32_t my_rank = 0) {
  const uint64_t nw = starts ? (uint64_t)nranks * bw : (n + 31) >> 5;
  int any = 0;
  for (uint64_t i = threadIdx.x; i < nw; i += blockDim.x) any |= (bitmap[i] != 0);
  any = __syncthreads_or(any);
  if (!any) {
    for (uint32_t k = threadIdx.x; k < p; k += blockDim.x) pivots[k] = 0;
    if (threadIdx.x == 0) { status[0] = 1; status[1] = 0; }
    return;
  }
  if (threadIdx.x >= 32) return;
  const unsigned lane = threadIdx.x;
  const uint64_t M48 = 0xFFFFFFFFFFFFull, A = 0x5DEECE66Dull, C = 0xBull;
  uint64_t a = 1, cc = 0;
  for (unsigned i = 0; i <= lane; ++i) { cc = (cc * A + C) & M48; a = (a * A) & M48; }
  const uint64_t scale = 0xFFFFFFFFull / n;
  uint64_t base = *state, draws = 0;
  uint32_t got = 0;
  while (got < p) {
    const uint64_t s = (base * a + cc) & M48;
    const uint64_t k = (s >> 16) / scale;
    const bool in_range = k < n;                       // else uniform_int draws again: a generator step, not a draw of the loop
    uint64_t bit = k, local = k;                       // where row k's bit sits in `bitmap`, and the row's index on its owner
    uint32_t owner = my_rank;
    if (starts && in_range) {
      owner = 0;
      while (owner + 1 < nranks && k >= starts[owner + 1]) ++owner;
      local = k - starts[owner];
      bit = (uint64_t)owner * bw * 32 + local;
    }
    const bool nz = in_range && ((bitmap[bit >> 5] >> (bit & 31)) & 1u);
    const uint32_t vmask = __ballot_sync(0xffffffffu, nz), rmask = __ballot_sync(0xffffffffu, in_range);
    const uint32_t need = p - got, cnt = (uint32_t)__popc(vmask);
    const uint32_t rank = (uint32_t)__popc(vmask & ((1u << lane) - 1u));
    if (nz && rank < need) pivots[got + rank] = (owner == my_rank) ? local : ~0ull;
    if (cnt >= need) {
      const unsigned L = __fns(vmask, 0, (int)need);   // lane of the draw that gave the last pivot
      base = __shfl_sync(0xffffffffu, s, L);
      draws += (uint64_t)__popc(rmask & (L == 31 ? 0xFFFFFFFFu : ((2u << L) - 1u)));
      got = p;
    } else {
      got += cnt;
      base = __shfl_sync(0xffffffffu, s, 31);
      draws += (uint64_t)__popc(rmask);
    }
  }
  if (lane == 0) { *state = base; status[0] = 0; status[1] = draws; }
}

// bitmap + draw, all queued on the stream: pivots (device, p u64), the generator state (device u64, in/out) and status
// (device, 2 u64) are ready when the stream gets there. n must be below 2^32 (gsl_rng_uniform_int's range).
bic_status bic_k_draw_pivots_device(bic_ctx* c, const bic_mat* X, uint64_t p, uint64_t* d_state, uint64_t* d_pivots,
                                    unsigned long long* d_status) {
  const uint64_t n = X->rows;
  if (n == 0 || n > 0xFFFFFFFFull) return bic_fail(c, BIC_ERR_INVALID, "draw_pivots: row count outside gsl_rng_uniform_int's range");
  const uint64_t nw = div_up_u64(n, 32);
  BIC_TRY(bic_scratch_reserve(c, &c->work[0], nw * 4));
  BIC_TRY(bic_k_row_nonzero_bitmap(c, X, (uint32_t*)c->work[0].p));
  BIC_PROF(c, KID_ROW_NONZERO);
  k_draw_pivots<<<1, 256, 0, c->stream>>>((const uint32_t*)c->work[0].p, n, (uint32_t)p, d_state, d_pivots, d_status);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

extern "C" bic_status bic_draw_pivots(bic_ctx* c, const bic_mat* X, uint64_t p, uint64_t* rng_state,
                                      uint64_t* pivots_out, uint64_t* ndraws_out) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !X || !rng_state || (!pivots_out && p)) return BIC_ERR_INVALID;
  const uint64_t n = X->rows;
  if (p == 0) { if (ndraws_out) *ndraws_out = 0; return BIC_OK; }
  if (n == 0) return bic_fail(c, BIC_ERR_INVALID, "draw_pivots: empty X");
  // the draw runs on the device (k_draw_pivots); the caller's generator state goes there and comes back
  uint64_t* d_state = c->d_scalars + 48;                               // [48] state, [49..50] status
  unsigned long long* d_status = (unsigned long long*)(c->d_scalars + 49);
  BIC_TRY(bic_scratch_reserve(c, &c->work[1], p * 8 + 64));
  uint64_t* d_piv = (uint64_t*)c->work[1].p;
  BIC_CUDA(c, cudaMemcpyAsync(d_state, rng_state, 8, cudaMemcpyHostToDevice, c->stream));
  BIC_TRY(bic_k_draw_pivots_device(c, X, p, d_state, d_piv, d_status));
  BIC_CUDA(c, cudaMemcpyAsync(pivots_out, d_piv, p * 8, cudaMemcpyDeviceToHost, c->stream));
  BIC_CUDA(c, cudaMemcpyAsync(c->h_scalars + 48, d_state, 24, cudaMemcpyDeviceToHost, c->stream));
  BIC_CUDA(c, bic_wait_stream(c));
  if (c->h_scalars[49]) return bic_fail(c, BIC_ERR_INVALID, "draw_pivots: X is all zero (the reference's draw loop never ends)");
  *rng_state = c->h_scalars[48];
  if (ndraws_out) *ndraws_out = c->h_scalars[50];
  return BIC_OK;
}

// ------------------------------------------------------------------ column histogram c[b]
// A warp takes blocks of 32 consecutive rows: lane r loads word w of row r (coalesced across the block's rows),
// the 32x32 bit tile is transposed in registers (shuffle butterfly), and lane b then holds column b of the tile:
// one POPC per 32 rows per column instead of 32 bit extractions. WORDS = words of a row per launch.
template <int WORDS>
__global__ void __launch_bounds__(256) k_col_hist(const uint32_t* __restrict__ X, uint64_t n, uint64_t wpr, uint64_t word0,
                                                  uint64_t nwords, uint32_t* __restrict__ hist /* [wpr*32] */) {
  __shared__ uint32_t s_cnt[WORDS * 32];
  for (int i = threadIdx.x; i < WORDS * 32; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t nblocks = div_up_u64(n, 32);
  uint32_t cnt[WORDS];
#pragma unroll
  for (int w = 0; w < WORDS; ++w) cnt[w] = 0;
  for (uint64_t blk = gw; blk < nblocks; blk += nwarps) {
    const uint64_t r = blk * 32 + lane;
    const uint32_t* row = X + r * wpr + word0;
#pragma unroll
    for (int w = 0; w < WORDS; ++w) {
      const uint32_t x = (r < n && (uint64_t)w < nwords) ? __ldg(row + w) : 0u;
      cnt[w] += __popc(warp_transpose32(x));
    }
  }
#pragma unroll
  for (int w = 0; w < WORDS; ++w)
    if ((uint64_t)w < nwords && cnt[w]) atomicAdd(&s_cnt[w * 32 + lane], cnt[w]);
  __syncthreads();
  for (int i = threadIdx.x; i < WORDS * 32; i += blockDim.x)
    if (s_cnt[i] && (uint64_t)(i >> 5) < nwords) atomicAdd(&hist[word0 * 32 + i], s_cnt[i]);
}

// ------------------------------------------------------------------ u[k] = #{j : X[j] & P_k != 0}
// Thread per row (row words in registers), pivot rows staged in shared memory (broadcast reads).
template <int WORDS>
__global__ void k_pivot_usage(const uint32_t* __restrict__ X, uint64_t n, uint64_t wpr,
                              const uint32_t* __restrict__ P /* [np][wpr] */, uint32_t np, uint32_t* __restrict__ usage) {
  extern __shared__ uint32_t sm[];
  uint32_t* Ps = sm;             // np * WORDS
  uint32_t* us = sm + (size_t)np * WORDS;  // np
  for (uint32_t i = threadIdx.x; i < np * WORDS; i += blockDim.x) {
    const uint32_t k = i / WORDS, w = i - k * WORDS;
    Ps[i] = (w < wpr) ? P[(uint64_t)k * wpr + w] : 0u;
  }
  for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) us[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint64_t nround = (n + 31) & ~(uint64_t)31;
  for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < nround; r += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t x[WORDS];
#pragma unroll
    for (int w = 0; w < WORDS; ++w) x[w] = (r < n && (uint64_t)w < wpr) ? __ldg(X + r * wpr + w) : 0u;
    for (uint32_t k = 0; k < np; ++k) {
      uint32_t any = 0;
#pragma unroll
      for (int w = 0; w < WORDS; ++w) any |= x[w] & Ps[k * WORDS + w];
      const uint32_t b = __ballot_sync(0xffffffffu, any != 0);
      if (lane == 0 && b) atomicAdd(&us[k], (uint32_t)__popc(b));
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < np; i += blockDim.x)
    if (us[i]) atomicAdd(&usage[i], us[i]);
}

// slow path for rows wider than 32 words: a warp per row, lanes stride the words
__global__ void k_pivot_usage_wide(const uint32_t* __restrict__ X, uint64_t n, uint64_t wpr,
                                   const uint32_t* __restrict__ P, uint32_t np, uint32_t* __restrict__ usage) {
  const int lane = threadIdx.x & 31;
  const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t r = gw; r < n; r += nwarps) {
    for (uint32_t k = 0; k < np; ++k) {
      uint32_t any = 0;
      for (uint64_t w = lane; w < wpr; w += 32) any |= __ldg(X + r * wpr + w) & __ldg(P + (uint64_t)k * wpr + w);
      if (__any_sync(0xffffffffu, any != 0) && lane == 0) atomicAdd(&usage[k], 1u);
    }
  }
}

__global__ void k_gather_rows(const uint32_t* __restrict__ X, uint64_t wpr, const uint64_t* __restrict__ pivots,
                              uint32_t np, uint32_t* __restrict__ P) {
  const uint64_t total = (uint64_t)np * wpr;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = i / wpr, w = i - k * wpr;
    const uint64_t r = pivots[k];
    P[i] = (r == ~0ull) ? 0u : X[r * wpr + w];  // ~0: the row lives on another rank
  }
}

// D[k][b] = (P_k[b] ? c[b] : 0) >= u_k / 2     (src/bsvd.cpp:258-262; ">=" and integer u/2)
__global__ void k_init_finalize(const uint32_t* __restrict__ P, const uint32_t* __restrict__ hist,
                                const uint32_t* __restrict__ usage, uint32_t np, uint64_t wpr, uint64_t m,
                                uint32_t* __restrict__ D) {
  const uint64_t total = (uint64_t)np * wpr;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = i / wpr, w = i - k * wpr;
    const uint32_t pw = P[i];
    const uint32_t half = usage[k] >> 1;
    uint32_t out = 0;
    for (int b = 0; b < 32; ++b) {
      const uint64_t j = w * 32 + b;
      if (j >= m) break;
      const uint32_t s = ((pw >> (31 - b)) & 1u) ? hist[j] : 0u;
      if (s >= half) out |= 0x80000000u >> b;
    }
    D[i] = out;
  }
}

template <int WORDS>
static bic_status launch_usage(bic_ctx* c, const bic_mat* X, const uint32_t* P, uint32_t np, uint32_t* usage) {
  const size_t smem = ((size_t)np * WORDS + np) * 4;
  if (smem > 48 * 1024)
    BIC_CUDA(c, cudaFuncSetAttribute(k_pivot_usage<WORDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BIC_PROF(c, KID_PIVOT_USAGE);
  k_pivot_usage<WORDS><<<bic_grid_for(c, X->rows, 256, 4), 256, smem, c->stream>>>(X->d, X->rows, X->wpr, P, np, usage);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

// ---- stages, shared with the row-sharded driver (dist.cu) ------------------------------------------
// scratch (work[1]): pivots (u64 p) | P rows (p*wpr u32) | hist (wpr*32 u32) | usage (p u32)
bic_status bic_k_init_scratch(bic_ctx* c, uint64_t p, uint64_t wpr, InitWork* w) {
  const size_t off_P = (size_t)p * 8;
  const size_t off_h = off_P + (size_t)p * wpr * 4;
  const size_t off_u = off_h + (size_t)wpr * 32 * 4;
  const size_t total = off_u + (size_t)p * 4;
  BIC_TRY(bic_scratch_reserve(c, &c->work[1], total + 16));
  uint8_t* base = (uint8_t*)c->work[1].p;
  w->piv = (uint64_t*)base;
  w->P = (uint32_t*)(base + off_P);
  w->hist = (uint32_t*)(base + off_h);
  w->usage = (uint32_t*)(base + off_u);
  w->p = p; w->wpr = wpr;
  return BIC_OK;
}

// P[k] = X[pivots[k]] (zero row where pivots[k] == ~0, i.e. owned by another rank)
bic_status bic_k_init_gather(bic_ctx* c, const bic_mat* X, const uint64_t* host_pivots, InitWork* w) {
  BIC_CUDA(c, cudaMemcpyAsync(w->piv, host_pivots, (size_t)w->p * 8, cudaMemcpyHostToDevice, c->stream));
  BIC_PROF(c, KID_GATHER_ROWS);
  k_gather_rows<<<bic_grid_for(c, w->p * w->wpr, 256, 4), 256, 0, c->stream>>>(X->d, w->wpr, w->piv, (uint32_t)w->p, w->P);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

// the same with the pivot list already in device memory at w->piv (k_draw_pivots wrote it)
bic_status bic_k_init_gather_dev(bic_ctx* c, const bic_mat* X, InitWork* w) {
  BIC_PROF(c, KID_GATHER_ROWS);
  k_gather_rows<<<bic_grid_for(c, w->p * w->wpr, 256, 4), 256, 0, c->stream>>>(X->d, w->wpr, w->piv, (uint32_t)w->p, w->P);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

bic_status bic_k_init_stats(bic_ctx* c, const bic_mat* X, InitWork* w);
bic_status bic_k_init_finalize(bic_ctx* c, InitWork* w, uint64_t m, bic_mat* D);

// initialize_model_neighbor with nothing waiting for the host: bitmap, draw, gather, statistics and thresholds are queued on
// the stream. d_state: the rand48 state (device u64, in/out); d_status: 2 u64 ([0] != 0: X is all zero, D is garbage).
bic_status bic_k_init_neighbor_async(bic_ctx* c, const bic_mat* X, bic_mat* D, bic_mat* A, uint64_t* d_state,
                                     unsigned long long* d_status) {
  BIC_RANGE("bic:initialize_model_neighbor(async)");
  const uint64_t p = D->rows;
  if (D->cols != X->cols || A->rows != X->rows || A->cols != p)
    return bic_fail(c, BIC_ERR_INVALID, "init: shapes must be X n x m, D p x m, A n x p");
  BIC_TRY(bic_mat_clear(c, A));
  BIC_TRY(bic_mat_clear(c, D));
  if (p == 0 || X->rows == 0 || X->cols == 0) return BIC_OK;
  InitWork w;
  BIC_TRY(bic_k_init_scratch(c, p, X->wpr, &w));
  BIC_TRY(bic_k_draw_pivots_device(c, X, p, d_state, w.piv, d_status));
  BIC_TRY(bic_k_init_gather_dev(c, X, &w));
  BIC_TRY(bic_k_init_stats(c, X, &w));
  return bic_k_init_finalize(c, &w, X->cols, D);
}

// local column histogram of X and, per pivot, the number of local rows that intersect it
bic_status bic_k_init_stats(bic_ctx* c, const bic_mat* X, InitWork* w) {
  const uint64_t wpr = w->wpr, p = w->p;
  BIC_CUDA(c, cudaMemsetAsync(w->hist, 0, (size_t)wpr * 32 * 4 + (size_t)p * 4, c->stream));
  if (X->rows == 0) return BIC_OK;
  for (uint64_t w0 = 0; w0 < wpr; w0 += 32) {  // 32 words of the row per launch
    const uint64_t nw = (wpr - w0 < 32) ? wpr - w0 : 32;
    const int grid = bic_grid_for(c, X->rows, 256, 4);  // a warp per 32-row block, a few blocks per warp
    BIC_PROF(c, KID_COL_HIST);
    if (nw <= 2) k_col_hist<2><<<grid, 256, 0, c->stream>>>(X->d, X->rows, wpr, w0, nw, w->hist);
    else if (nw <= 8) k_col_hist<8><<<grid, 256, 0, c->stream>>>(X->d, X->rows, wpr, w0, nw, w->hist);
    else k_col_hist<32><<<grid, 256, 0, c->stream>>>(X->d, X->rows, wpr, w0, nw, w->hist);
    BIC_LAUNCH_CHECK(c);
  }
  if (wpr <= 32) {  // pivot usage, in chunks of pivots whose rows fit in shared memory
    const int WORDS = wpr <= 1 ? 1 : wpr <= 2 ? 2 : wpr <= 4 ? 4 : wpr <= 8 ? 8 : wpr <= 16 ? 16 : 32;
    const uint64_t max_np = (96 * 1024 / 4) / (WORDS + 1);
    for (uint64_t k0 = 0; k0 < p; k0 += max_np) {
      const uint32_t np = (uint32_t)((p - k0 < max_np) ? p - k0 : max_np);
      const uint32_t* Pk = w->P + k0 * wpr;
      uint32_t* uk = w->usage + k0;
      switch (WORDS) {
        case 1: BIC_TRY(launch_usage<1>(c, X, Pk, np, uk)); break;
        case 2: BIC_TRY(launch_usage<2>(c, X, Pk, np, uk)); break;
        case 4: BIC_TRY(launch_usage<4>(c, X, Pk, np, uk)); break;
        case 8: BIC_TRY(launch_usage<8>(c, X, Pk, np, uk)); break;
        case 16: BIC_TRY(launch_usage<16>(c, X, Pk, np, uk)); break;
        default: BIC_TRY(launch_usage<32>(c, X, Pk, np, uk)); break;
      }
    }
  } else {
    BIC_PROF(c, KID_PIVOT_USAGE);
    k_pivot_usage_wide<<<bic_grid_for(c, X->rows * 32, 256, 8), 256, 0, c->stream>>>(X->d, X->rows, wpr, w->P, (uint32_t)p, w->usage);
    BIC_LAUNCH_CHECK(c);
  }
  return BIC_OK;
}

bic_status bic_k_init_finalize(bic_ctx* c, InitWork* w, uint64_t m, bic_mat* D) {
  BIC_PROF(c, KID_INIT_FINALIZE);
  k_init_finalize<<<bic_grid_for(c, w->p * w->wpr, 256, 4), 256, 0, c->stream>>>(w->P, w->hist, w->usage, (uint32_t)w->p, w->wpr, m, D->d);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

extern "C" bic_status bic_initialize_model_neighbor_pivots(bic_ctx* c, const bic_mat* X, const uint64_t* pivots,
                                                           uint64_t p, bic_mat* D, bic_mat* A) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !X || !D || !A || (!pivots && p)) return BIC_ERR_INVALID;
  if (D->rows != p || D->cols != X->cols || A->rows != X->rows || A->cols != p)
    return bic_fail(c, BIC_ERR_INVALID, "init: shapes must be X n x m, D p x m, A n x p");
  for (uint64_t k = 0; k < p; ++k)
    if (pivots[k] >= X->rows) return bic_fail(c, BIC_ERR_INVALID, "init: pivot out of range");
  BIC_TRY(bic_mat_clear(c, A));  // A.clear(), src/bsvd.cpp:237
  BIC_TRY(bic_mat_clear(c, D));  // D.clear(), :238
  if (p == 0 || X->rows == 0 || X->cols == 0) return BIC_OK;
  InitWork w;
  BIC_TRY(bic_k_init_scratch(c, p, X->wpr, &w));
  BIC_TRY(bic_k_init_gather(c, X, pivots, &w));
  BIC_TRY(bic_k_init_stats(c, X, &w));
  return bic_k_init_finalize(c, &w, X->cols, D);
}

extern "C" bic_status bic_initialize_model_neighbor(bic_ctx* c, const bic_mat* X, bic_mat* D, bic_mat* A,
                                                    uint64_t* rng_state) {
  BIC_RANGE("bic:initialize_model_neighbor");
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !X || !D || !A || !rng_state) return BIC_ERR_INVALID;
  const uint64_t p = D->rows;
  std::vector<uint64_t> piv(p ? p : 1);
  BIC_TRY(bic_draw_pivots(c, X, p, rng_state, piv.data(), nullptr));
  return bic_initialize_model_neighbor_pivots(c, X, piv.data(), p, D, A);
}
