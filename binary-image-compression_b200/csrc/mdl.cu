// MDL model selection over the bsvd fit: the learners that call the hot path tens of times.
// Reference: model_codelength src/bsvd.cpp:1438-1461 over universal_codelength src/coding.cpp:24-32;
// learn_model_mdl_forward_selection :1463-1546, learn_model_mdl_backward_selection :1548-1660,
// learn_model_mdl_full_search :1662-1717; inner learner learn_model_traditional, initialiser
// initialize_model_neighbor (the reference's default plug points).
//
// Everything that touches a matrix runs on the device: the weights the description length needs (|E|, the
// row weights of D, the column weights of A) are integer reductions, the backward selection's "residual if
// atom k were dropped" is one pass for ALL atoms (|E xor A_k' D_k| = |E| + sum over users of k of
// (|E_i xor D_k| - |E_i|)), growing / shrinking D and A are row copies and a column shift. The only floating
// point on the path is the reference's own: double log2 on the host over those integers, truncated into
// integers at every accumulation exactly where the reference's idx_t accumulators truncate.
#include "bic_internal.cuh"

#include <math.h>

#include <vector>

bic_status bic_k_update_dictionary(bic_ctx* c, bic_mat* E, bic_mat* D, const bic_mat* A, unsigned long long* d_changed);

// ------------------------------------------------------------------ kernels
// weight of every row (warp per row)
__global__ void __launch_bounds__(256) k_row_weights(const uint32_t* __restrict__ M, uint64_t rows, uint64_t wpr,
                                                     uint32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t r = gw; r < rows; r += nw) {
    uint32_t c = 0;
    for (uint64_t w = lane; w < wpr; w += 32) c += __popc(__ldg(M + r * wpr + w));
    c = warp_sum_u32(c);
    if (lane == 0) out[r] = c;
  }
}

// weight of every column: 32 rows x 32 columns at a time, transposed in registers so lane j holds column j's 32 bits
__global__ void __launch_bounds__(256) k_col_weights(const uint32_t* __restrict__ M, uint64_t n, uint64_t wpr,
                                                     uint32_t* __restrict__ out) {
  extern __shared__ uint32_t s_cnt[];  // wpr * 32
  const int lane = threadIdx.x & 31;
  for (uint64_t i = threadIdx.x; i < wpr * 32; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t nblk = div_up_u64(n, 32);
  for (uint64_t blk = gw; blk < nblk; blk += nw) {
    const uint64_t row = blk * 32 + lane;
    for (uint64_t w = 0; w < wpr; ++w) {
      const uint32_t v = (row < n) ? __ldg(M + row * wpr + w) : 0u;
      if (!__any_sync(0xffffffffu, v != 0)) continue;
      const uint32_t c = __popc(warp_transpose32(v));
      if (c) atomicAdd(&s_cnt[w * 32 + lane], c);
    }
  }
  __syncthreads();
  for (uint64_t i = threadIdx.x; i < wpr * 32; i += blockDim.x)
    if (s_cnt[i]) atomicAdd(out + i, s_cnt[i]);
}

// delta[k] = sum over the rows that use atom k of (|E_i xor D_k| - |E_i|): the change of |E| if atom k were
// dropped from the model (E xor A_k' D_k, src/bsvd.cpp:1582-1584), for every atom in one pass over the rows
__global__ void __launch_bounds__(256) k_atom_removal_delta(const uint32_t* __restrict__ E, const uint32_t* __restrict__ D,
                                                            const uint32_t* __restrict__ A, uint64_t n, uint64_t wprE,
                                                            uint64_t wprA, uint32_t p, long long* __restrict__ delta) {
  extern __shared__ int s_d[];  // p
  for (uint32_t i = threadIdx.x; i < p; i += blockDim.x) s_d[i] = 0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t row = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; row < n; row += stride) {
    const uint32_t* a = A + row * wprA;
    const uint32_t* e = E + row * wprE;
    int we = -1;
    for (uint64_t t = 0; t < wprA; ++t) {
      uint32_t ab = __ldg(a + t);
      while (ab) {
        const int q = __clz(ab);
        ab &= ~(0x80000000u >> q);
        const uint32_t k = (uint32_t)t * 32 + q;
        if (we < 0) {
          we = 0;
          for (uint64_t w = 0; w < wprE; ++w) we += __popc(__ldg(e + w));
        }
        int wx = 0;
        for (uint64_t w = 0; w < wprE; ++w) wx += __popc(__ldg(e + w) ^ __ldg(D + (uint64_t)k * wprE + w));
        atomicAdd(&s_d[k], wx - we);
      }
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < p; i += blockDim.x)
    if (s_d[i]) atomicAdd((unsigned long long*)delta + i, (unsigned long long)(long long)s_d[i]);
}

// Aout (n x pout) from Ain (n x pin): column `drop` removed when drop < pin (pout = pin - 1), or the columns kept and
// zero columns appended (drop >= pin, pout >= pin)
__global__ void __launch_bounds__(256) k_reshape_cols(const uint32_t* __restrict__ Ain, uint32_t* __restrict__ Aout, uint64_t n,
                                                      uint64_t wprIn, uint64_t wprOut, uint64_t pin, uint64_t drop) {
  const uint64_t total = n * wprOut;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const uint64_t row = i / wprOut, w = i - row * wprOut;
    const uint32_t* a = Ain + row * wprIn;
    const uint32_t cur = (w < wprIn) ? __ldg(a + w) : 0u;
    uint32_t o;
    if (drop >= pin || (w + 1) * 32 <= drop) {
      o = cur;                                        // entirely before the dropped column (or nothing dropped)
    } else {
      const uint32_t nxt = (w + 1 < wprIn) ? __ldg(a + w + 1) : 0u;
      const uint32_t shifted = (cur << 1) | (nxt >> 31);  // column j+1 -> j (MSB first)
      if (w * 32 >= drop) {
        o = shifted;                                  // entirely after it
      } else {
        const uint32_t b = (uint32_t)(drop - w * 32);  // the dropped column sits at bit b of this word
        const uint32_t before = b ? (0xFFFFFFFFu << (32 - b)) : 0u;
        o = (cur & before) | (shifted & ~before);
      }
    }
    Aout[i] = o;
  }
}

// ------------------------------------------------------------------ device-side statistics of a model
struct ModelStats {
  uint64_t wE = 0;
  std::vector<uint32_t> wD, wA;         // row weights of D, column weights of A
  std::vector<long long> removal;       // optional: delta of |E| per dropped atom
};

static bic_status model_stats(bic_ctx* c, const bic_mat* E, const bic_mat* D, const bic_mat* A, bool want_removal, ModelStats* st) {
  const uint64_t p = D ? D->rows : 0, n = E->rows;
  BIC_TRY(bic_mat_weight(c, E, &st->wE));
  st->wD.assign(p, 0);
  st->wA.assign(p, 0);
  st->removal.assign(want_removal ? p : 0, 0);
  if (p == 0) return BIC_OK;
  const uint64_t wprA = A->wpr;
  // work[4]: row weights (p) | column weights (wprA*32) | removal deltas (p x 8 bytes)
  const size_t words = (size_t)p + wprA * 32 + 2 * p + 8;
  BIC_TRY(bic_scratch_reserve(c, &c->work[4], words * 4));
  uint32_t* d_wD = (uint32_t*)c->work[4].p;
  uint32_t* d_wA = d_wD + p;
  long long* d_rm = (long long*)(d_wA + wprA * 32 + ((p + wprA * 32) & 1));
  BIC_CUDA(c, cudaMemsetAsync(d_wD, 0, words * 4, c->stream));
  k_row_weights<<<bic_grid_for(c, p * 32, 256, 4), 256, 0, c->stream>>>(D->d, p, D->wpr, d_wD);
  BIC_LAUNCH_CHECK(c);
  if (n) {
    k_col_weights<<<bic_grid_for(c, n, 256, 4), 256, (size_t)wprA * 32 * 4, c->stream>>>(A->d, n, wprA, d_wA);
    BIC_LAUNCH_CHECK(c);
    if (want_removal) {
      k_atom_removal_delta<<<bic_grid_for(c, n, 256, 8), 256, (size_t)p * 4, c->stream>>>(E->d, D->d, A->d, n, E->wpr, wprA,
                                                                                        (uint32_t)p, d_rm);
      BIC_LAUNCH_CHECK(c);
    }
  }
  BIC_CUDA(c, cudaMemcpyAsync(st->wD.data(), d_wD, p * 4, cudaMemcpyDeviceToHost, c->stream));
  BIC_CUDA(c, cudaMemcpyAsync(st->wA.data(), d_wA, p * 4, cudaMemcpyDeviceToHost, c->stream));
  if (want_removal) BIC_CUDA(c, cudaMemcpyAsync(st->removal.data(), d_rm, p * 8, cudaMemcpyDeviceToHost, c->stream));
  BIC_CUDA(c, bic_wait_stream(c));
  return BIC_OK;
}

// ------------------------------------------------------------------ the reference's description length
// src/coding.cpp:24-32 (the parameters are `unsigned` there: 64-bit counts wrap to 32 bits)
extern "C" double bic_universal_codelength(unsigned n, unsigned r) {
  const double p1 = (double)r / (double)n;
  if ((r > 0) && (r < n)) {
    return double(n) * (-p1 * log2(p1) - (1.0 - p1) * log2(1.0 - p1)) + 0.5 * log2(n);
  } else {
    return 0.5 * log2(n);
  }
}

// src/bsvd.cpp:1438-1461 from the integer statistics; LD, LA are idx_t there, so `+= double` truncates every time
static uint64_t codelength_from_stats(uint64_t n, uint64_t m, uint64_t wE, const std::vector<uint32_t>& wD,
                                      const std::vector<uint32_t>& wA) {
  const uint64_t LE = (uint64_t)bic_universal_codelength((unsigned)(n * m), (unsigned)wE);
  uint64_t LD = 0, LA = 0;
  for (size_t k = 0; k < wD.size(); ++k) {
    LD = (uint64_t)((double)LD + bic_universal_codelength((unsigned)m, wD[k]));
    LA = (uint64_t)((double)LA + bic_universal_codelength((unsigned)n, wA[k]));
  }
  return LE + LD + LA;
}

extern "C" bic_status bic_model_codelength(bic_ctx* c, const bic_mat* E, const bic_mat* D, const bic_mat* A, uint64_t* L) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !E || !L) return BIC_ERR_INVALID;
  if (D && (!A || D->cols != E->cols || A->rows != E->rows || A->cols != D->rows))
    return bic_fail(c, BIC_ERR_INVALID, "model_codelength: shapes must be E n x m, D p x m, A n x p");
  ModelStats st;
  BIC_TRY(model_stats(c, E, D, A, false, &st));
  *L = codelength_from_stats(E->rows, E->cols, st.wE, st.wD, st.wA);
  return BIC_OK;
}

// ------------------------------------------------------------------ growing / shrinking a model
// The learners' working matrices come from the stream-ordered pool (no device synchronisation per create / destroy);
// what is handed back to the caller is copied into ordinary allocations at the end (unpool).
static bic_status reshape_cols(bic_ctx* c, const bic_mat* Ain, uint64_t pout, uint64_t drop, bic_mat** out) {
  BIC_TRY(bic_mat_create_pooled(c, Ain->rows, pout, out));
  if (Ain->rows && pout) {
    k_reshape_cols<<<bic_grid_for(c, Ain->rows * (*out)->wpr, 256, 8), 256, 0, c->stream>>>(Ain->d, (*out)->d, Ain->rows, Ain->wpr,
                                                                                         (*out)->wpr, Ain->cols, drop);
    BIC_LAUNCH_CHECK(c);
  }
  return BIC_OK;
}

static bic_status clone_mat(bic_ctx* c, const bic_mat* M, bic_mat** out) {
  BIC_TRY(bic_mat_create_pooled(c, M->rows, M->cols, out));
  return bic_mat_copy(c, M, *out);
}

static bic_status unpool(bic_ctx* c, bic_mat** M) {
  if (!*M || !(*M)->pooled) return BIC_OK;
  bic_mat* r = nullptr;
  BIC_TRY(bic_mat_create(c, (*M)->rows, (*M)->cols, &r));
  BIC_TRY(bic_mat_copy(c, *M, r));
  bic_mat_destroy(c, *M);
  *M = r;
  return BIC_OK;
}

static void drop_mat(bic_ctx* c, bic_mat** M) {
  if (*M) bic_mat_destroy(c, *M);
  *M = nullptr;
}

static bic_status inner_learn(bic_ctx* c, const bic_mat* X, bic_mat* E, bic_mat* D, bic_mat* A) {
  uint64_t it = 0;
  return bic_learn_model_traditional(c, X, E, D, A, &it, nullptr, 0);   // learn_model_inner, src/bsvd.cpp:21
}

// learn_model_mdl_forward_selection, src/bsvd.cpp:1463-1546
static bic_status mdl_forward(bic_ctx* c, const bic_mat* X, bic_mat* E, bic_mat** Dp, bic_mat** Ap, uint64_t* rng, uint64_t* bestL_out) {
  const uint64_t n = E->rows, m = E->cols;
  uint64_t K = (*Dp)->rows;
  BIC_TRY(inner_learn(c, X, E, *Dp, *Ap));                                 // :1470
  bic_mat *nextAtom = nullptr, *nextCoefs = nullptr, *currD = nullptr, *currA = nullptr, *currE = nullptr;
  BIC_TRY(bic_mat_create_pooled(c, 1, m, &nextAtom));
  BIC_TRY(bic_mat_create_pooled(c, n, 1, &nextCoefs));
  BIC_TRY(clone_mat(c, *Dp, &currD));
  BIC_TRY(clone_mat(c, *Ap, &currA));
  BIC_TRY(clone_mat(c, E, &currE));
  uint64_t bestL = 0;
  BIC_TRY(bic_model_codelength(c, E, *Dp, *Ap, &bestL));                   // :1476
  uint64_t stuck = 0, sumStuck = 0, allStuck = 0;
  bic_status st = BIC_OK;
  do {
    const int dev = allStuck > 0 ? (int)(sumStuck / allStuck) : 0;         // :1486
    // one new atom from the current residual; its coefficient column starts at zero (initialize_model, :1488)
    if ((st = bic_initialize_model_neighbor(c, currE, nextAtom, nextCoefs, rng)) != BIC_OK) break;
    bic_mat *nD = nullptr, *nA = nullptr;
    if ((st = bic_mat_create_pooled(c, K + 1, m, &nD)) != BIC_OK) break;          // :1499-1506
    if ((st = bic_mat_copy_rows(c, currD, 0, K, nD, 0)) != BIC_OK) break;
    if ((st = bic_mat_copy_rows(c, nextAtom, 0, 1, nD, K)) != BIC_OK) break;
    if ((st = reshape_cols(c, currA, K + 1, ~0ull, &nA)) != BIC_OK) break;  // :1508-1516 (the new column is all zero)
    drop_mat(c, &currD);
    drop_mat(c, &currA);
    currD = nD;
    currA = nA;
    if ((st = inner_learn(c, X, currE, currD, currA)) != BIC_OK) break;    // :1518
    uint64_t currL = 0;
    if ((st = bic_model_codelength(c, currE, currD, currA, &currL)) != BIC_OK) break;
    if ((currL + (uint64_t)(int64_t)dev) < bestL) {                        // :1520
      stuck = 0;
      bestL = currL;
      drop_mat(c, Dp);
      drop_mat(c, Ap);
      if ((st = clone_mat(c, currD, Dp)) != BIC_OK) break;
      if ((st = clone_mat(c, currA, Ap)) != BIC_OK) break;
      if ((st = bic_mat_copy(c, currE, E)) != BIC_OK) break;
    } else {
      stuck++;
      allStuck++;
      sumStuck += (currL - bestL);
      if (stuck >= 10) break;                                              // :1532-1535
    }
    K++;
  } while (stuck < 10);
  drop_mat(c, &nextAtom); drop_mat(c, &nextCoefs); drop_mat(c, &currD); drop_mat(c, &currA); drop_mat(c, &currE);
  *bestL_out = bestL;
  return st;
}

// learn_model_mdl_backward_selection, src/bsvd.cpp:1548-1660
static bic_status mdl_backward(bic_ctx* c, const bic_mat* X, bic_mat* E, bic_mat** Dp, bic_mat** Ap, uint64_t* bestL_out) {
  const uint64_t n = E->rows, m = E->cols;
  uint64_t K = (*Dp)->rows;
  BIC_TRY(inner_learn(c, X, E, *Dp, *Ap));                                 // :1555
  uint64_t bestL = 0;
  BIC_TRY(bic_model_codelength(c, E, *Dp, *Ap, &bestL));
  bic_mat *currD = nullptr, *currA = nullptr, *nextD = nullptr, *nextA = nullptr, *nextE = nullptr;
  BIC_TRY(clone_mat(c, *Dp, &currD));
  BIC_TRY(clone_mat(c, *Ap, &currA));
  BIC_TRY(bic_mat_create_pooled(c, n, m, &nextE));
  uint64_t stuck = 0, sumStuck = 0, allStuck = 0;
  bic_status st = BIC_OK;
  for (; K > 0; K--) {
    const int dev = allStuck > 0 ? (int)(sumStuck / allStuck) : 0;         // :1575
    // the atom whose removal leaves the shortest description (:1577-1592). The reference measures the removal
    // against E -- the residual of the best model so far -- with the coefficients and atoms of the current one.
    ModelStats ms;
    if ((st = model_stats(c, E, currD, currA, true, &ms)) != BIC_OK) break;
    uint64_t LD = 0, LA = 0;
    for (uint64_t k = 0; k < K; ++k) {
      LD = (uint64_t)((double)LD + bic_universal_codelength((unsigned)m, ms.wD[k]));
      LA = (uint64_t)((double)LA + bic_universal_codelength((unsigned)n, ms.wA[k]));
    }
    uint64_t nextk = 0;
    uint64_t nextL = ~(1UL << (sizeof(uint64_t) - 1));                     // :1578
    uint64_t w_last = 0;
    for (uint64_t k = 0; k < K; ++k) {
      const uint64_t wk = (uint64_t)((long long)ms.wE + ms.removal[k]);    // |E xor A_k' D_k|
      w_last = wk;
      uint64_t tmpL = (uint64_t)bic_universal_codelength((unsigned)(n * m), (unsigned)wk) + LD + LA;
      tmpL = (uint64_t)((double)tmpL - bic_universal_codelength((unsigned)m, ms.wD[k]));
      tmpL = (uint64_t)((double)tmpL - bic_universal_codelength((unsigned)n, ms.wA[k]));
      if (tmpL < nextL) { nextL = tmpL; nextk = k; }
    }
    drop_mat(c, &nextD);
    drop_mat(c, &nextA);
    if (K > 1) {                                                           // :1598-1616
      if ((st = bic_mat_create_pooled(c, K - 1, m, &nextD)) != BIC_OK) break;
      if ((st = bic_mat_copy_rows(c, currD, 0, nextk, nextD, 0)) != BIC_OK) break;
      if ((st = bic_mat_copy_rows(c, currD, nextk + 1, K - 1 - nextk, nextD, nextk)) != BIC_OK) break;
      if ((st = reshape_cols(c, currA, K - 1, nextk, &nextA)) != BIC_OK) break;
      if ((st = inner_learn(c, X, nextE, nextD, nextA)) != BIC_OK) break;
      if ((st = bic_model_codelength(c, nextE, nextD, nextA, &nextL)) != BIC_OK) break;
    } else {
      // no atom left: the description is the residual alone, E xor A_0' D_0 (:1618)
      nextL = (uint64_t)bic_universal_codelength((unsigned)(n * m), (unsigned)w_last);
    }
    if (nextL + (uint64_t)(int64_t)dev < bestL) {                          // :1621
      if (K == 1) {                                                        // "Resulted in empty model!", :1623-1629
        drop_mat(c, Dp);
        drop_mat(c, Ap);
        st = bic_mat_copy(c, X, E);
        break;
      }
      stuck = 0;
      bestL = nextL;
      drop_mat(c, Dp);
      drop_mat(c, Ap);
      if ((st = clone_mat(c, nextD, Dp)) != BIC_OK) break;
      if ((st = clone_mat(c, nextA, Ap)) != BIC_OK) break;
      if ((st = bic_mat_copy(c, nextE, E)) != BIC_OK) break;
    } else {
      stuck++;
      allStuck++;
      sumStuck += (nextL - bestL);
      if (stuck >= 10) break;
    }
    drop_mat(c, &currD);                                                   // :1649-1655
    drop_mat(c, &currA);
    if (K > 1) {
      if ((st = clone_mat(c, nextD, &currD)) != BIC_OK) break;
      if ((st = clone_mat(c, nextA, &currA)) != BIC_OK) break;
    } else {
      if ((st = bic_mat_create_pooled(c, 0, m, &currD)) != BIC_OK) break;
      if ((st = bic_mat_create_pooled(c, n, 0, &currA)) != BIC_OK) break;
    }
  }
  drop_mat(c, &currD); drop_mat(c, &currA); drop_mat(c, &nextD); drop_mat(c, &nextA); drop_mat(c, &nextE);
  *bestL_out = bestL;
  return st;
}

// learn_model_mdl_full_search, src/bsvd.cpp:1662-1717: sizes 20, 40, ... up to the rows of the D handed in; eleven
// fits per size on one continuing RNG stream; the matrices kept are those of the size's LAST fit, the length
// recorded the minimum over its last ten.
static bic_status mdl_full_search(bic_ctx* c, const bic_mat* X, bic_mat* E, bic_mat** Dp, bic_mat** Ap, uint64_t* rng, uint64_t* bestL_out) {
  const uint64_t n = E->rows, m = E->cols, Kmax = (*Dp)->rows;
  bic_mat* candE = nullptr;
  BIC_TRY(bic_mat_create_pooled(c, n, m, &candE));
  uint64_t bestL = 1UL << 30;
  bic_status st = BIC_OK;
  for (uint64_t k = 20; k <= Kmax && st == BIC_OK; k += 20) {
    bic_mat *candD = nullptr, *candA = nullptr;
    if ((st = bic_mat_create_pooled(c, k, m, &candD)) != BIC_OK) break;
    if ((st = bic_mat_create_pooled(c, n, k, &candA)) != BIC_OK) { drop_mat(c, &candD); break; }
    uint64_t candL = ~0ull;
    for (int rep = 0; rep < 11 && st == BIC_OK; ++rep) {
      if ((st = bic_initialize_model_neighbor(c, X, candD, candA, rng)) != BIC_OK) break;
      if ((st = inner_learn(c, X, candE, candD, candA)) != BIC_OK) break;
      if (rep == 0) continue;                                              // the first fit is not scored (:1672-1673)
      uint64_t L = 0;
      if ((st = bic_model_codelength(c, candE, candD, candA, &L)) != BIC_OK) break;
      if (L < candL) candL = L;
    }
    if (st == BIC_OK && candL < bestL) {
      bestL = candL;
      st = bic_mat_copy(c, candE, E);
      drop_mat(c, Dp);
      drop_mat(c, Ap);
      *Dp = candD;
      *Ap = candA;
    } else {
      drop_mat(c, &candD);
      drop_mat(c, &candA);
    }
  }
  drop_mat(c, &candE);
  *bestL_out = bestL;
  return st;
}

extern "C" bic_status bic_learn_model_mdl(bic_ctx* c, int lm, const bic_mat* X, bic_mat* E, bic_mat** D, bic_mat** A,
                                          uint64_t* rng_state, uint64_t* best_codelength) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !X || !E || !D || !A || !*D || !*A || !best_codelength) return BIC_ERR_INVALID;
  if (X->rows != E->rows || X->cols != E->cols || (*D)->cols != E->cols || (*A)->rows != E->rows || (*A)->cols != (*D)->rows)
    return bic_fail(c, BIC_ERR_INVALID, "learn_model_mdl: shapes must be X, E n x m, D p x m, A n x p");
  bic_status st;
  if (lm == 4) st = rng_state ? mdl_forward(c, X, E, D, A, rng_state, best_codelength) : BIC_ERR_INVALID;
  else if (lm == 5) st = mdl_backward(c, X, E, D, A, best_codelength);
  else if (lm == 6) st = rng_state ? mdl_full_search(c, X, E, D, A, rng_state, best_codelength) : BIC_ERR_INVALID;
  else return bic_fail(c, BIC_ERR_INVALID, "learn_model_mdl: lm must be 4 (forward), 5 (backward) or 6 (full search)");
  const bic_status s1 = unpool(c, D), s2 = unpool(c, A);
  return st != BIC_OK ? st : (s1 != BIC_OK ? s1 : s2);
}
