// update_coefficients (greedy matching pursuit over GF(2)) and the GF(2) residual product.
// Reference: update_coefficients_omp src/bsvd.cpp:1029-1107 (serial twin :399-460);
// mul_AB src/binmat.cpp:516-543 + add :463-478.
//
// Per row i, repeat: w = |E_i|; pick k minimising |E_i xor D_k| (strict <, so the lowest k wins
// ties); if that distance is strictly below w flip A[i,k] and E_i ^= D_k, else stop.
// Rows are independent given D, so this is one thread per row with the row in registers and D
// staged in shared memory as [atom][word] (every lane of a warp reads the same atom word, a
// broadcast). The work is XOR + POPC: it is bound by the popcount pipe, not by HBM.
#include "bic_internal.cuh"

// distance and atom index packed in one key so the argmin with the reference's tie-break is a
// plain min: key = dist << 16 | k  (dist <= 32*32 = 1024 bits, k < 65536)
template <int WORDS>
__device__ __forceinline__ uint32_t best_atom_key(const uint32_t (&e)[WORDS], const uint32_t* __restrict__ Ds, uint32_t p) {
  uint32_t best = 0xFFFFFFFFu;
  for (uint32_t k = 0; k < p; ++k) {
    const uint32_t* dk = Ds + k * WORDS;
    uint32_t d = 0;
#pragma unroll
    for (int w = 0; w < WORDS; ++w) d += __popc(e[w] ^ dk[w]);
    best = min(best, (d << 16) | k);
  }
  return best;
}

template <int WORDS>
__global__ void __launch_bounds__(256) k_update_coefficients(uint32_t* __restrict__ E, const uint32_t* __restrict__ D,
                                                             uint32_t* __restrict__ A, uint64_t n, uint64_t wprE,
                                                             uint32_t p, uint64_t wprA,
                                                             unsigned long long* __restrict__ changed,
                                                             const ProbDev* __restrict__ probs,
                                                             const uint32_t* __restrict__ active,
                                                             const uint32_t* __restrict__ skip,
                                                             unsigned long long* __restrict__ passes) {
  if (skip && *skip) return;  // the learner's loop already ended on the device (queued-ahead iteration)
  if (probs) {  // batched launch: blockIdx.y selects the problem
    if (!active[blockIdx.y]) return;
    const ProbDev pr = probs[blockIdx.y];
    E = pr.E; D = pr.D; A = pr.A; changed = pr.counts;
  }
  extern __shared__ __align__(16) uint32_t Ds[];  // p * WORDS, rows zero padded to WORDS
  __shared__ uint64_t tma_bar;
  // The atom tile: when the rows need no padding (m = 64, 256, 1024 ...) the whole dictionary is one
  // contiguous block and is staged by the TMA engine (cp.async.bulk) while the threads set up;
  // otherwise it is re-strided with ordinary loads.
  const bool bulk = (wprE == WORDS) && ((p * WORDS * 4u) % 16u == 0);
  if (bulk) {
    if (threadIdx.x == 0) tma_stage_begin(&tma_bar);
    __syncthreads();
    if (threadIdx.x == 0) tma_stage_copy(Ds, D, p * WORDS * 4u, &tma_bar);
    tma_stage_wait(&tma_bar);
  } else {
    for (uint32_t i = threadIdx.x; i < p * WORDS; i += blockDim.x) {
      const uint32_t k = i / WORDS, w = i - k * WORDS;
      Ds[i] = (w < wprE) ? D[(uint64_t)k * wprE + w] : 0u;
    }
    __syncthreads();
  }
  // Rows need different numbers of greedy passes (0 .. weight). A lane that finishes its row takes the
  // next row of its warp's range at once instead of idling until the slowest lane of the warp is
  // done: every trip of the loop below is ONE pass (all atoms) for each lane that holds a row, and
  // lanes at different passes of different rows run the same instructions.
  uint32_t nchanged = 0, npasses = 0;   // npasses: greedy passes over the dictionary (the algorithmic popcount work: p * WORDS each)
  const int lane = threadIdx.x & 31;
  const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t chunk = div_up_u64(n, nwarps);
  uint64_t next = gw * chunk;                       // warp-uniform: next unassigned row of this warp
  const uint64_t w_end = (next + chunk < n) ? next + chunk : n;
  bool have = false, row_changed = false;
  uint64_t r = 0;
  uint32_t e[WORDS];
#pragma unroll
  for (int w = 0; w < WORDS; ++w) e[w] = 0;
  for (;;) {
    const uint32_t need = __ballot_sync(0xffffffffu, !have);
    if (need) {
      const uint64_t cand = next + __popc(need & ((1u << lane) - 1u));
      if (!have && cand < w_end) {
        r = cand;
        const uint32_t* erow = E + r * wprE;
#pragma unroll
        for (int w = 0; w < WORDS; ++w) e[w] = ((uint64_t)w < wprE) ? erow[w] : 0u;
        have = true;
        row_changed = false;
      }
      next += __popc(need);
    }
    if (!__any_sync(0xffffffffu, have)) break;
    if (have) {
      uint32_t wt = 0;  // w = Ei.weight(), src/bsvd.cpp:1065
#pragma unroll
      for (int w = 0; w < WORDS; ++w) wt += __popc(e[w]);
      bool again = false;
      if (wt) {         // with weight 0 no distance can be smaller
        ++npasses;
        const uint32_t key = best_atom_key<WORDS>(e, Ds, p);  // :1067-1082
        const uint32_t bestd = key >> 16, bestk = key & 0xFFFFu;
        if (bestd < wt) {  // :1084, strict <
          BIC_DCHECK(bestk < p && r < n);
          A[r * wprA + (bestk >> 5)] ^= 0x80000000u >> (bestk & 31);  // Ai.flip(0,bestk), :1086
          const uint32_t* dk = Ds + bestk * WORDS;
#pragma unroll
          for (int w = 0; w < WORDS; ++w) e[w] ^= dk[w];  // :1087
          row_changed = true;
          again = true;
        }
      }
      if (!again) {      // the row is finished (:1090-1099)
        if (row_changed) {
          nchanged++;
          uint32_t* erow = E + r * wprE;
#pragma unroll
          for (int w = 0; w < WORDS; ++w)
            if ((uint64_t)w < wprE) erow[w] = e[w];
        }
        have = false;
      }
    }
  }
  nchanged = warp_sum_u32(nchanged);
  npasses = warp_sum_u32(npasses);
  if ((threadIdx.x & 31) == 0 && nchanged) atomicAdd(changed, (unsigned long long)nchanged);
  if ((threadIdx.x & 31) == 0 && npasses && passes) atomicAdd(passes, (unsigned long long)npasses);
}

// ------------------------------------------------------------------ large dictionaries: a warp per row, atoms pruned by weight
// With many atoms (p >= 64) most of them cannot win a pass: |E_i xor D_k| >= | |E_i| - |D_k| |. The CTA keeps the dictionary
// in shared memory SORTED BY WEIGHT (with the original indices); a warp takes one row and visits atoms outward from the
// row's own weight, 16 lighter and 16 heavier ones per round (one atom per lane), tightening the bound after every round:
// an atom can only win, or tie with the current best, if | |E_i| - |D_k| | <= T = min(|E_i| - 1, best distance so far).
// The sort makes the survivors two contiguous runs, so a round is skipped for ALL lanes at once -- which a lane-per-row
// kernel cannot do. Exactness: every atom that could beat or tie the final best is evaluated (the bound is inclusive),
// and the key (distance << 16 | original index) keeps the reference's lowest-index tie-break whatever the visiting order.
template <int WORDS>
__global__ void __launch_bounds__(256) k_update_coefficients_sorted(uint32_t* __restrict__ E, const uint32_t* __restrict__ D,
                                                                    uint32_t* __restrict__ A, uint64_t n, uint64_t wprE, uint32_t p,
                                                                    uint64_t wprA, unsigned long long* __restrict__ changed,
                                                                    unsigned long long* __restrict__ next_row, uint32_t grab,
                                                                    const uint32_t* __restrict__ skip) {
  if (skip && *skip) return;
  constexpr int STR = WORDS + 1;                 // row stride of the sorted dictionary: consecutive atoms hit distinct banks
  extern __shared__ __align__(16) uint32_t sm[];
  uint32_t* Ds = sm;                             // p * STR, sorted by (weight, original index)
  uint32_t* ws = Ds + (size_t)p * STR;           // p weights, ascending
  uint16_t* ks = (uint16_t*)(ws + p);            // p original indices
  uint16_t* pos = ks + p;                        // p: sorted position of original atom k
  uint32_t* wraw = (uint32_t*)(pos + p);         // p: weights in original order (scratch for the rank sort); ks + pos = 4p bytes
  for (uint32_t k = threadIdx.x; k < p; k += blockDim.x) {
    uint32_t w = 0;
    for (uint32_t j = 0; j < (uint32_t)wprE; ++j) w += __popc(__ldg(D + (uint64_t)k * wprE + j));
    wraw[k] = w;
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < p; k += blockDim.x) {  // rank sort, stable in the original index
    const uint32_t w = wraw[k];
    uint32_t r = 0;
    for (uint32_t j = 0; j < p; ++j) r += (wraw[j] < w) || (wraw[j] == w && j < k);
    ws[r] = w;
    ks[r] = (uint16_t)k;
    pos[k] = (uint16_t)r;
    for (int j = 0; j < WORDS; ++j) Ds[(size_t)r * STR + j] = ((uint64_t)j < wprE) ? __ldg(D + (uint64_t)k * wprE + j) : 0u;
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  uint32_t nchanged = 0;
  for (;;) {
    // rows are handed out `grab` at a time (blank rows cost almost nothing, text rows many rounds)
    unsigned long long r0 = 0;
    if (lane == 0) r0 = atomicAdd(next_row, (unsigned long long)grab);
    r0 = __shfl_sync(0xffffffffu, r0, 0);
    if (r0 >= n) break;
    const uint64_t r1 = (r0 + grab < n) ? r0 + grab : n;
    for (uint64_t r = r0; r < r1; ++r) {
      uint32_t e[WORDS];
      uint32_t wt = 0;
#pragma unroll
      for (int j = 0; j < WORDS; ++j) {
        e[j] = ((uint64_t)j < wprE) ? __ldg(E + r * wprE + j) : 0u;   // the same address in every lane: a broadcast
        wt += __popc(e[j]);
      }
      bool row_changed = false;
      while (wt) {                               // one greedy pass per trip (src/bsvd.cpp:1063-1099)
        uint32_t T = wt - 1;                     // a winner needs distance < wt
        uint32_t best = 0xFFFFFFFFu;             // distance << 16 | original index
        // first position whose weight is >= wt (all lanes the same search)
        uint32_t lo = 0, hi = p;
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if (ws[mid] < wt) lo = mid + 1; else hi = mid;
        }
        uint32_t left = lo, right = lo;          // atoms [left, right) have been visited
        for (;;) {
          // the nearest unvisited atom on either side decides whether that side still has candidates
          const bool more_l = left > 0 && (wt - ws[left - 1]) <= T;
          const bool more_r = right < p && (ws[right] - wt) <= T;
          if (!more_l && !more_r) break;
          uint32_t idx = 0xFFFFFFFFu;
          if (lane < 16) { if (left > lane) idx = left - 1 - lane; }
          else { if (right + (lane - 16) < p) idx = right + (lane - 16); }
          uint32_t key = 0xFFFFFFFFu;
          if (idx != 0xFFFFFFFFu) {
            const uint32_t wk = ws[idx];
            const uint32_t lb = wk > wt ? wk - wt : wt - wk;
            if (lb <= T) {
              const uint32_t* dk = Ds + (size_t)idx * STR;
              uint32_t d = 0;
#pragma unroll
              for (int j = 0; j < WORDS; ++j) d += __popc(e[j] ^ dk[j]);
              key = (d << 16) | ks[idx];
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
          best = min(best, key);
          if ((best >> 16) < T) T = best >> 16;  // ties with the best so far must still be evaluated: inclusive bound
          left = left > 16 ? left - 16 : 0;
          right = right + 16 < p ? right + 16 : p;
        }
        const uint32_t bestd = best >> 16, bestk = best & 0xFFFFu;
        if (best == 0xFFFFFFFFu || bestd >= wt) break;   // :1084, strict <
        if (lane == 0) A[r * wprA + (bestk >> 5)] ^= 0x80000000u >> (bestk & 31);  // :1086
        const uint32_t* dk = Ds + (size_t)pos[bestk] * STR;
        wt = 0;
#pragma unroll
        for (int j = 0; j < WORDS; ++j) { e[j] ^= dk[j]; wt += __popc(e[j]); }     // :1087
        row_changed = true;
      }
      if (row_changed) {
        nchanged++;
#pragma unroll
        for (int j = 0; j < WORDS; ++j)
          if (lane == (uint32_t)j && (uint64_t)j < wprE) E[r * wprE + j] = e[j];
      }
    }
  }
  if (lane == 0 && nchanged) atomicAdd(changed, (unsigned long long)nchanged);
}

template <int WORDS>
static bic_status launch_coef_sorted(bic_ctx* c, bic_mat* E, const bic_mat* D, bic_mat* A, unsigned long long* d_changed) {
  const uint64_t p = D->rows;
  const size_t smem = (size_t)p * (WORDS + 1) * 4 + p * 12 + 16;
  if (smem > 48 * 1024)
    BIC_CUDA(c, cudaFuncSetAttribute(k_update_coefficients_sorted<WORDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = (int)((200 * 1024) / smem);
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  const int grid = bic_grid_for(c, E->rows * 4, 256, per_sm);   // a warp per row
  // rows per grab: about four grabs per warp, between 8 and 128 (one global atomic each)
  uint64_t grab = E->rows / ((uint64_t)grid * 8 * 4);
  grab = grab < 8 ? 8 : (grab > 128 ? 128 : grab);
  unsigned long long* next_row = (unsigned long long*)(c->d_scalars + 8);
  BIC_CUDA(c, cudaMemsetAsync(next_row, 0, 8, c->stream));
  BIC_PROF(c, KID_UPDATE_COEF);
  k_update_coefficients_sorted<WORDS><<<grid, 256, smem, c->stream>>>(E->d, D->d, A->d, E->rows, E->wpr, (uint32_t)p, A->wpr, d_changed,
                                                                     next_row, (uint32_t)grab, c->loop_skip);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

// Fallback for rows wider than 32 words (m > 1024) or a dictionary that does not fit in shared
// memory: one warp per row, the row staged in shared memory, D read through L1/L2.
__global__ void __launch_bounds__(256) k_update_coefficients_wide(uint32_t* __restrict__ E, const uint32_t* __restrict__ D,
                                                                  uint32_t* __restrict__ A, uint64_t n, uint64_t wprE,
                                                                  uint32_t p, uint64_t wprA,
                                                                  unsigned long long* __restrict__ changed,
                                                                  const uint32_t* __restrict__ skip) {
  if (skip && *skip) return;
  extern __shared__ uint32_t es_all[];  // (blockDim/32) * wprE
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t* es = es_all + (size_t)wib * wprE;
  const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  uint32_t nchanged = 0;
  for (uint64_t r = gw; r < n; r += nwarps) {
    uint32_t* erow = E + r * wprE;
    for (uint64_t w = lane; w < wprE; w += 32) es[w] = erow[w];
    __syncwarp();
    bool row_changed = false;
    for (;;) {
      uint32_t wt = 0;
      for (uint64_t w = lane; w < wprE; w += 32) wt += __popc(es[w]);
      wt = warp_sum_u32(wt);
      if (wt == 0) break;
      unsigned long long best = ~0ull;  // dist << 32 | k
      for (uint32_t k = 0; k < p; ++k) {
        const uint32_t* dk = D + (uint64_t)k * wprE;
        uint32_t d = 0;
        for (uint64_t w = lane; w < wprE; w += 32) d += __popc(es[w] ^ __ldg(dk + w));
        d = warp_sum_u32(d);
        const unsigned long long key = ((unsigned long long)d << 32) | k;
        best = key < best ? key : best;
      }
      const uint32_t bestd = (uint32_t)(best >> 32), bestk = (uint32_t)best;
      if (bestd >= wt) break;
      if (lane == 0) A[r * wprA + (bestk >> 5)] ^= 0x80000000u >> (bestk & 31);
      const uint32_t* dk = D + (uint64_t)bestk * wprE;
      for (uint64_t w = lane; w < wprE; w += 32) es[w] ^= __ldg(dk + w);
      __syncwarp();
      row_changed = true;
    }
    if (row_changed) {
      if (lane == 0) nchanged++;
      for (uint64_t w = lane; w < wprE; w += 32) erow[w] = es[w];
    }
    __syncwarp();
  }
  if (lane == 0 && nchanged) atomicAdd(changed, (unsigned long long)nchanged);
}

template <int WORDS>
static bic_status launch_coef(bic_ctx* c, bic_mat* E, const bic_mat* D, bic_mat* A, unsigned long long* d_changed) {
  const size_t smem = (size_t)D->rows * WORDS * 4;
  if (smem > 48 * 1024)
    BIC_CUDA(c, cudaFuncSetAttribute(k_update_coefficients<WORDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // persistent grid: a multiple of the SM count, as many CTAs per SM as the dictionary allows
  int per_sm = smem ? (int)((200 * 1024) / smem) : 8;
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  const int grid = bic_grid_for(c, E->rows, 256, per_sm);
  BIC_PROF(c, KID_UPDATE_COEF);
  k_update_coefficients<WORDS><<<grid, 256, smem, c->stream>>>(E->d, D->d, A->d, E->rows, E->wpr, (uint32_t)D->rows,
                                                              A->wpr, d_changed, nullptr, nullptr, c->loop_skip, (unsigned long long*)(c->d_scalars + BIC_SCALAR_COEF_PASSES));
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

template <int WORDS>
static bic_status launch_coef_batched(bic_ctx* c, uint64_t n, uint64_t wprE, uint64_t p, uint64_t wprA, const ProbDev* probs,
                                      const uint32_t* active, uint32_t nprob) {
  const size_t smem = (size_t)p * WORDS * 4;
  if (smem > 48 * 1024)
    BIC_CUDA(c, cudaFuncSetAttribute(k_update_coefficients<WORDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = smem ? (int)((200 * 1024) / smem) : 8;
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  uint64_t gx = ((uint64_t)c->sm_count * per_sm + nprob - 1) / nprob;  // the whole batch fills the GPU once
  const uint64_t need = div_up_u64(n, 256);
  if (gx > need) gx = need;
  if (gx < 1) gx = 1;
  BIC_PROF(c, KID_UPDATE_COEF);
  k_update_coefficients<WORDS><<<dim3((unsigned)gx, nprob), 256, smem, c->stream>>>(nullptr, nullptr, nullptr, n, wprE, (uint32_t)p,
                                                                                   wprA, nullptr, probs, active, nullptr, nullptr);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

// all problems of a batch in one launch (same shapes); only rows up to 1024 bits and a dictionary that fits
// in shared memory (the single-problem path has a fallback for the rest)
bic_status bic_k_update_coefficients_batched(bic_ctx* c, uint64_t n, uint64_t m, uint64_t p, const ProbDev* probs,
                                             const uint32_t* active, uint32_t nprob) {
  const uint64_t wpr = div_up_u64(m, 32), wprA = div_up_u64(p, 32);
  const int WORDS = wpr <= 1 ? 1 : wpr <= 2 ? 2 : wpr <= 4 ? 4 : wpr <= 8 ? 8 : wpr <= 16 ? 16 : wpr <= 32 ? 32 : 0;
  if (!WORDS || (size_t)p * WORDS * 4 > 200 * 1024 || p > 65535) return BIC_ERR_UNSUPPORTED;
  switch (WORDS) {
    case 1: return launch_coef_batched<1>(c, n, wpr, p, wprA, probs, active, nprob);
    case 2: return launch_coef_batched<2>(c, n, wpr, p, wprA, probs, active, nprob);
    case 4: return launch_coef_batched<4>(c, n, wpr, p, wprA, probs, active, nprob);
    case 8: return launch_coef_batched<8>(c, n, wpr, p, wprA, probs, active, nprob);
    case 16: return launch_coef_batched<16>(c, n, wpr, p, wprA, probs, active, nprob);
    default: return launch_coef_batched<32>(c, n, wpr, p, wprA, probs, active, nprob);
  }
}

// device-side entry used by the learner too: adds the changed-row count to *d_changed
bic_status bic_k_update_coefficients(bic_ctx* c, bic_mat* E, const bic_mat* D, bic_mat* A, unsigned long long* d_changed) {
  BIC_RANGE("bic:update_coefficients");
  if (E->rows != A->rows || E->cols != D->cols || A->cols != D->rows)
    return bic_fail(c, BIC_ERR_INVALID, "update_coefficients: shapes must be E n x m, D p x m, A n x p");
  if (E->rows == 0 || D->rows == 0) return BIC_OK;
  if (D->rows > 65535) return bic_fail(c, BIC_ERR_UNSUPPORTED, "update_coefficients: more than 65535 atoms");
  const uint64_t wpr = E->wpr;
  const int WORDS = wpr <= 1 ? 1 : wpr <= 2 ? 2 : wpr <= 4 ? 4 : wpr <= 8 ? 8 : wpr <= 16 ? 16 : wpr <= 32 ? 32 : 0;
  // many atoms, rows of 4..32 words: the weight-sorted warp-per-row kernel ("coef_algo" 0 keeps the lane-per-row kernel)
  if (c->coef_algo != 0 && WORDS >= 4 && D->rows >= 64 && (size_t)D->rows * (WORDS + 4) * 4 + 64 <= 200 * 1024) {
    switch (WORDS) {
      case 4: return launch_coef_sorted<4>(c, E, D, A, d_changed);
      case 8: return launch_coef_sorted<8>(c, E, D, A, d_changed);
      case 16: return launch_coef_sorted<16>(c, E, D, A, d_changed);
      default: return launch_coef_sorted<32>(c, E, D, A, d_changed);
    }
  }
  if (WORDS && (size_t)D->rows * WORDS * 4 <= 200 * 1024) {
    switch (WORDS) {
      case 1: return launch_coef<1>(c, E, D, A, d_changed);
      case 2: return launch_coef<2>(c, E, D, A, d_changed);
      case 4: return launch_coef<4>(c, E, D, A, d_changed);
      case 8: return launch_coef<8>(c, E, D, A, d_changed);
      case 16: return launch_coef<16>(c, E, D, A, d_changed);
      default: return launch_coef<32>(c, E, D, A, d_changed);
    }
  }
  const size_t smem = (size_t)8 * wpr * 4;
  if (smem > 200 * 1024) return bic_fail(c, BIC_ERR_UNSUPPORTED, "update_coefficients: rows wider than 200 KB");
  if (smem > 48 * 1024)
    BIC_CUDA(c, cudaFuncSetAttribute(k_update_coefficients_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BIC_PROF(c, KID_UPDATE_COEF);
  k_update_coefficients_wide<<<bic_grid_for(c, E->rows * 32, 256, 4), 256, smem, c->stream>>>(
      E->d, D->d, A->d, E->rows, wpr, (uint32_t)D->rows, A->wpr, d_changed, c->loop_skip);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

extern "C" bic_status bic_update_coefficients(bic_ctx* c, bic_mat* E, const bic_mat* D, bic_mat* A, uint64_t* changed) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !E || !D || !A) return BIC_ERR_INVALID;
  BIC_CUDA(c, cudaMemsetAsync(c->d_scalars, 0, sizeof(uint64_t), c->stream));
  BIC_TRY(bic_k_update_coefficients(c, E, D, A, (unsigned long long*)c->d_scalars));
  BIC_TRY(bic_read_scalars(c, 1));
  if (changed) *changed = c->h_scalars[0];
  return BIC_OK;
}

// ------------------------------------------------------------------ E = A*D xor X
template <int WORDS>
__global__ void __launch_bounds__(256) k_residual(const uint32_t* __restrict__ X, const uint32_t* __restrict__ A,
                                                  const uint32_t* __restrict__ D, uint32_t* __restrict__ E, uint64_t n,
                                                  uint64_t wprE, uint32_t p, uint64_t wprA, bool d_in_smem) {
  extern __shared__ uint32_t Ds[];
  if (d_in_smem) {
    for (uint32_t i = threadIdx.x; i < p * WORDS; i += blockDim.x) {
      const uint32_t k = i / WORDS, w = i - k * WORDS;
      Ds[i] = (w < wprE) ? D[(uint64_t)k * wprE + w] : 0u;
    }
    __syncthreads();
  }
  for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t e[WORDS];
#pragma unroll
    for (int w = 0; w < WORDS; ++w) e[w] = ((uint64_t)w < wprE) ? X[r * wprE + w] : 0u;
    for (uint64_t aw = 0; aw < wprA; ++aw) {
      uint32_t bits = A[r * wprA + aw];
      while (bits) {
        const int pos = __clz(bits);
        bits &= ~(0x80000000u >> pos);
        const uint32_t k = (uint32_t)aw * 32 + pos;
        if (d_in_smem) {
#pragma unroll
          for (int w = 0; w < WORDS; ++w) e[w] ^= Ds[k * WORDS + w];
        } else {
#pragma unroll
          for (int w = 0; w < WORDS; ++w)
            if ((uint64_t)w < wprE) e[w] ^= __ldg(D + (uint64_t)k * wprE + w);
        }
      }
    }
#pragma unroll
    for (int w = 0; w < WORDS; ++w)
      if ((uint64_t)w < wprE) E[r * wprE + w] = e[w];
  }
}

// generic: a thread per word of E
__global__ void k_residual_wide(const uint32_t* __restrict__ X, const uint32_t* __restrict__ A,
                                const uint32_t* __restrict__ D, uint32_t* __restrict__ E, uint64_t n, uint64_t wprE,
                                uint64_t wprA) {
  const uint64_t total = n * wprE;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = i / wprE, w = i - r * wprE;
    uint32_t e = X[i];
    for (uint64_t aw = 0; aw < wprA; ++aw) {
      uint32_t bits = __ldg(A + r * wprA + aw);
      while (bits) {
        const int pos = __clz(bits);
        bits &= ~(0x80000000u >> pos);
        e ^= __ldg(D + (aw * 32 + pos) * wprE + w);
      }
    }
    E[i] = e;
  }
}

template <int WORDS>
static bic_status launch_residual(bic_ctx* c, const bic_mat* X, const bic_mat* A, const bic_mat* D, bic_mat* E) {
  size_t smem = (size_t)D->rows * WORDS * 4;
  const bool in_smem = smem <= 200 * 1024;
  if (!in_smem) smem = 0;
  if (smem > 48 * 1024)
    BIC_CUDA(c, cudaFuncSetAttribute(k_residual<WORDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = smem ? (int)((200 * 1024) / smem) : 8;
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  BIC_PROF(c, KID_RESIDUAL);
  k_residual<WORDS><<<bic_grid_for(c, X->rows, 256, per_sm), 256, smem, c->stream>>>(
      X->d, A->d, D->d, E->d, X->rows, X->wpr, (uint32_t)D->rows, A->wpr, in_smem);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

extern "C" bic_status bic_residual(bic_ctx* c, const bic_mat* X, const bic_mat* A, const bic_mat* D, bic_mat* E) {
  BIC_RANGE("bic:residual");
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !X || !A || !D || !E) return BIC_ERR_INVALID;
  if (X->rows != A->rows || X->cols != D->cols || A->cols != D->rows || E->rows != X->rows || E->cols != X->cols)
    return bic_fail(c, BIC_ERR_INVALID, "residual: shapes must be X,E n x m, D p x m, A n x p");
  if (X->words() == 0) return BIC_OK;
  const uint64_t wpr = X->wpr;
  if (wpr <= 1) return launch_residual<1>(c, X, A, D, E);
  if (wpr <= 2) return launch_residual<2>(c, X, A, D, E);
  if (wpr <= 4) return launch_residual<4>(c, X, A, D, E);
  if (wpr <= 8) return launch_residual<8>(c, X, A, D, E);
  if (wpr <= 16) return launch_residual<16>(c, X, A, D, E);
  if (wpr <= 32) return launch_residual<32>(c, X, A, D, E);
  BIC_PROF(c, KID_RESIDUAL);
  k_residual_wide<<<bic_grid_for(c, X->words(), 256, 8), 256, 0, c->stream>>>(X->d, A->d, D->d, E->d, X->rows, wpr, A->wpr);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}
