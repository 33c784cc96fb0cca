// update_dictionary_proximus, src/bsvd.cpp:528-729: per atom, in order, alternate
//   (a) a majority vote for the atom over its users (the steepest update's vote, :627-668) and
//   (b) a majority vote for the atom's coefficient COLUMN over the atom's set bits (:673-715):
//       Aw[i] = sum_{j in D_k} E[i][j] xor A[i][k], new A[i][k] = Aw[i] > |D_k| / 2,
// patching E after each, until neither changes. (a) is a histogram over the atom's users plus a patch pass, (b) is one
// pass over all rows (rows are independent: popc(E_i & D_k) decides, a flipped coefficient XORs D_k into the row).
// The atoms are sequential and every inner round needs its two "changed?" flags on the host, so this path is a chain
// of small launches; it exists for parity with the reference's `-d 1`, the throughput path is the steepest update.
#include "bic_internal.cuh"

// (a1) column counts of (E_i xor D_k) over the users of atom k, and the number of users (hist[m])
__global__ void __launch_bounds__(256) k_prox_hist(const uint32_t* __restrict__ E, const uint32_t* __restrict__ Dk,
                                                   const uint32_t* __restrict__ A, uint64_t n, uint64_t wprE, uint64_t wprA,
                                                   uint32_t k, uint64_t m, uint32_t* __restrict__ hist) {
  extern __shared__ uint32_t s_h[];  // m + 1
  for (uint64_t i = threadIdx.x; i <= m; i += blockDim.x) s_h[i] = 0;
  __syncthreads();
  const uint32_t kw = k >> 5, kbit = 0x80000000u >> (k & 31);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t row = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; row < n; row += stride) {
    if (!(__ldg(A + row * wprA + kw) & kbit)) continue;
    atomicAdd(&s_h[m], 1u);
    for (uint64_t w = 0; w < wprE; ++w) {
      uint32_t x = E[row * wprE + w] ^ __ldg(Dk + w);  // add-back old atom, :637-638
      while (x) {
        const int b = __clz(x);
        x &= ~(0x80000000u >> b);
        atomicAdd(&s_h[w * 32 + b], 1u);
      }
    }
  }
  __syncthreads();
  for (uint64_t i = threadIdx.x; i <= m; i += blockDim.x)
    if (s_h[i]) atomicAdd(hist + i, s_h[i]);
}

// (a2) every CTA derives the new atom from the counts; CTA 0 stores it and raises the flag; all patch the users' rows
__global__ void __launch_bounds__(256) k_prox_fix_atom(uint32_t* __restrict__ E, uint32_t* __restrict__ Dk, const uint32_t* __restrict__ A,
                                                       uint64_t n, uint64_t wprE, uint64_t wprA, uint32_t k, uint64_t m,
                                                       const uint32_t* __restrict__ hist, uint32_t* __restrict__ flags) {
  extern __shared__ uint32_t s_delta[];  // wprE
  __shared__ int s_any;
  const int lane = threadIdx.x & 31;
  const uint32_t u = __ldcg(hist + m);
  if (u == 0) return;                      // :647
  const uint32_t half = u / 2;             // :649
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  for (uint64_t w = threadIdx.x >> 5; w < wprE; w += blockDim.x >> 5) {
    const uint64_t j = w * 32 + lane;
    const uint32_t bit = (j < m) ? (__ldcg(hist + j) > half) : 0u;  // :651-652
    const uint32_t nd = __brev(__ballot_sync(0xffffffffu, bit));
    if (lane == 0) {
      const uint32_t d = nd ^ Dk[w];
      s_delta[w] = d;
      if (d) s_any = 1;
    }
  }
  __syncthreads();
  if (!s_any) return;                      // dd == 0, :654
  const uint32_t kw = k >> 5, kbit = 0x80000000u >> (k & 31);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t row = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; row < n; row += stride) {
    if (!(__ldg(A + row * wprA + kw) & kbit)) continue;
    for (uint64_t w = 0; w < wprE; ++w)
      if (s_delta[w]) E[row * wprE + w] ^= s_delta[w];   // :659-666
  }
  // the atom itself is rewritten only after every CTA has read the old one: done by the NEXT kernel (k_prox_commit_atom)
  if (blockIdx.x == 0 && threadIdx.x == 0) flags[0] = 1;
}

// (a3) D_k ^= delta, recomputed from the same counts (a kernel boundary after every CTA of (a2) used the old atom)
__global__ void k_prox_commit_atom(uint32_t* __restrict__ Dk, uint64_t wprE, uint64_t m, const uint32_t* __restrict__ hist,
                                   const uint32_t* __restrict__ flags) {
  if (!__ldcg(flags)) return;
  const int lane = threadIdx.x & 31;
  const uint32_t half = __ldcg(hist + m) / 2;
  for (uint64_t w = threadIdx.x >> 5; w < wprE; w += blockDim.x >> 5) {
    const uint64_t j = w * 32 + lane;
    const uint32_t bit = (j < m) ? (__ldcg(hist + j) > half) : 0u;
    const uint32_t nd = __brev(__ballot_sync(0xffffffffu, bit));
    if (lane == 0) Dk[w] = nd;             // :655
  }
}

// (b) the coefficient column of atom k: one pass over all rows
__global__ void __launch_bounds__(256) k_prox_column(uint32_t* __restrict__ E, const uint32_t* __restrict__ Dk, uint32_t* __restrict__ A,
                                                     uint64_t n, uint64_t wprE, uint64_t wprA, uint32_t k, uint32_t* __restrict__ flags) {
  extern __shared__ uint32_t s_d[];        // wprE
  __shared__ uint32_t s_w;
  if (threadIdx.x == 0) s_w = 0;
  __syncthreads();
  uint32_t wsum = 0;
  for (uint64_t w = threadIdx.x; w < wprE; w += blockDim.x) {
    const uint32_t d = __ldcg(Dk + w);
    s_d[w] = d;
    wsum += __popc(d);
  }
  if (wsum) atomicAdd(&s_w, wsum);
  __syncthreads();
  const uint32_t u = s_w;                  // bits of the atom, :680-682
  if (u == 0) return;                      // :693
  const uint32_t half = u / 2;             // :695
  const uint32_t kw = k >> 5, kbit = 0x80000000u >> (k & 31);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  bool any = false;
  for (uint64_t row = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; row < n; row += stride) {
    uint32_t c = 0;
    for (uint64_t w = 0; w < wprE; ++w) c += __popc(E[row * wprE + w] & s_d[w]);
    const uint32_t aw = A[row * wprA + kw];
    const bool a = (aw & kbit) != 0;
    const uint32_t Aw = a ? u - c : c;     // sum over the atom's bits of E[i][j] xor A[i][k], :683-689
    const bool na = Aw > half;             // :698
    if (na != a) {
      A[row * wprA + kw] = aw ^ kbit;      // :702
      for (uint64_t w = 0; w < wprE; ++w)
        if (s_d[w]) E[row * wprE + w] ^= s_d[w];   // :706-713
      any = true;
    }
  }
  if (any) flags[1] = 1;
}

extern "C" bic_status bic_update_dictionary_proximus(bic_ctx* c, bic_mat* E, bic_mat* D, bic_mat* A, uint64_t* changed) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !E || !D || !A) return BIC_ERR_INVALID;
  if (E->rows != A->rows || E->cols != D->cols || A->cols != D->rows)
    return bic_fail(c, BIC_ERR_INVALID, "update_dictionary_proximus: shapes must be E n x m, D p x m, A n x p");
  const uint64_t n = E->rows, m = E->cols, p = D->rows, wprE = E->wpr, wprA = A->wpr;
  uint64_t nchanged = 0;
  if (n && m && p) {
    if ((m + 1) * 4 > 200 * 1024) return bic_fail(c, BIC_ERR_UNSUPPORTED, "update_dictionary_proximus: rows wider than 50K bits");
    // work[5]: hist (m + 1) | flags (2)
    BIC_TRY(bic_scratch_reserve(c, &c->work[5], (size_t)(m + 1 + 2 + 8) * 4));
    uint32_t* hist = (uint32_t*)c->work[5].p;
    uint32_t* flags = hist + m + 1;
    const size_t hs_bytes = (size_t)(m + 1) * 4;
    if (hs_bytes > 48 * 1024) BIC_CUDA(c, cudaFuncSetAttribute(k_prox_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hs_bytes));
    const int grid = bic_grid_for(c, n, 256, 4);
    uint32_t* h_flags = (uint32_t*)(c->h_scalars + 48);
    for (uint64_t k = 0; k < p; ++k) {
      uint32_t* Dk = D->d + k * wprE;
      bool kchanged = false;
      for (;;) {
        BIC_CUDA(c, cudaMemsetAsync(hist, 0, (size_t)(m + 1 + 2) * 4, c->stream));
        BIC_PROF(c, KID_PROXIMUS);
        k_prox_hist<<<grid, 256, hs_bytes, c->stream>>>(E->d, Dk, A->d, n, wprE, wprA, (uint32_t)k, m, hist);
        BIC_LAUNCH_CHECK(c);
        BIC_PROF(c, KID_PROXIMUS);
        k_prox_fix_atom<<<grid, 256, (size_t)wprE * 4, c->stream>>>(E->d, Dk, A->d, n, wprE, wprA, (uint32_t)k, m, hist, flags);
        BIC_LAUNCH_CHECK(c);
        BIC_PROF(c, KID_PROXIMUS);
        k_prox_commit_atom<<<1, 256, 0, c->stream>>>(Dk, wprE, m, hist, flags);
        BIC_LAUNCH_CHECK(c);
        BIC_PROF(c, KID_PROXIMUS);
        k_prox_column<<<grid, 256, (size_t)wprE * 4, c->stream>>>(E->d, Dk, A->d, n, wprE, wprA, (uint32_t)k, flags);
        BIC_LAUNCH_CHECK(c);
        BIC_CUDA(c, cudaMemcpyAsync(h_flags, flags, 8, cudaMemcpyDeviceToHost, c->stream));
        BIC_CUDA(c, bic_wait_stream(c));
        if (h_flags[0]) kchanged = true;         // only a changed ATOM counts (:657, :711)
        if (!h_flags[0] && !h_flags[1]) break;   // converged, :716
      }
      if (kchanged) nchanged++;
    }
  }
  if (changed) *changed = nchanged;
  return BIC_OK;
}
