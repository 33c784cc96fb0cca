// update_dictionary_steepest on the device. Reference: src/bsvd.cpp:463-527.
//
// For atom k = 0..p-1, IN ORDER: over the rows that use it (A[i,k] = 1) take the per-bit
// majority of E_i xor D_k ("strictly more than usage/2" -> 1); if the atom changed, patch those
// rows' residuals, E_i ^= D_k xor newD_k, before the next atom is looked at (Gauss-Seidel: a row
// using atoms k and l > k carries atom k's change into atom l's vote).
//
// Device plan
//   1. A (n x p) is bit-transposed once into AT (p x n): atom k's users are then one contiguous
//      n-bit row, read coalesced. A does not change during the update.
//   2. One persistent cooperative kernel walks the atoms. Per atom: (a) every warp scans a slice
//      of AT[k], and for each user row adds the row's bits into per-lane counters (lane l owns
//      bit l of every word; the row word is a warp-broadcast load); counters go to a global
//      histogram with integer atomics, which are order independent, so the result is exact;
//      (b) grid barrier; (c) every CTA derives newD_k from the histogram (same result in all
//      CTAs); (d) if it differs from D_k, all CTAs patch their slice of user rows and a second
//      grid barrier orders those writes before the next atom's reads.
//   weights[j] = sum_i (E_i[j] xor D_k[j]) = D_k[j] ? usage - cE[j] : cE[j], cE = column count of
//   the users' E rows, so only cE and usage are accumulated.
#include "bic_internal.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

// ------------------------------------------------------------------ A -> AT (bit transpose)
// CTA = 1024 threads = 32 warps; warp i takes rows [r0+32i, r0+32i+32) of one 32-column slab and
// transposes the 32x32 bit tile with 32 ballots; the 32 tiles of a CTA give each of the slab's 32
// atoms 32 consecutive words (128 B) of AT, written coalesced through shared memory.
__global__ void __launch_bounds__(1024) k_transpose_bits(const uint32_t* __restrict__ A, uint64_t n, uint64_t wprA,
                                                         uint32_t* __restrict__ AT, uint64_t wprN, uint64_t p,
                                                         const ProbDev* __restrict__ probs, const uint32_t* __restrict__ active) {
  if (probs) {  // batched launch: blockIdx.z selects the problem
    if (!active[blockIdx.z]) return;
    A = probs[blockIdx.z].A;
    AT = probs[blockIdx.z].AT;
  }
  __shared__ uint32_t tile[32][33];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint64_t cw = blockIdx.y;                    // column word of A (32 atoms)
  const uint64_t nblk = div_up_u64(n, 1024);
  for (uint64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const uint64_t r = blk * 1024 + (uint64_t)wib * 32 + lane;
    const uint32_t v = (r < n) ? A[r * wprA + cw] : 0u;
    const uint32_t mine = warp_transpose32(v);  // lane b: atom b of the slab, rows of this warp
    tile[lane][wib] = mine;  // atom `lane` of the slab, row-word `wib` of the block
    __syncthreads();
    const uint64_t atom = cw * 32 + wib;
    const uint64_t word = blk * 32 + lane;
    if (atom < p && word < wprN) AT[atom * wprN + word] = tile[wib][lane];
    __syncthreads();
  }
}

// ------------------------------------------------------------------ the per-atom walk
struct DictParams {
  uint32_t* E;          // n x wprE
  uint32_t* D;          // p x wprE
  const uint32_t* AT;   // p x wprN
  uint32_t* hist;       // 3 x (hist_stride): [0, m) column counts, [m] usage
  unsigned long long* changed;
  uint64_t n, wprE, wprN, m, hist_stride;
  uint32_t p;
};

template <int WORDS>
__global__ void __launch_bounds__(256) k_update_dictionary(DictParams P) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ uint32_t sm[];
  uint32_t* s_hist = sm;                 // WORDS*32 + 1
  uint32_t* s_delta = sm + WORDS * 32 + 1;  // WORDS
  __shared__ int s_any;
  const int lane = threadIdx.x & 31;
  const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t nchunks = div_up_u64(P.wprN, 32);  // 32 AT words (1024 rows) per warp step

  for (uint32_t k = 0; k < P.p; ++k) {
    uint32_t* hist = P.hist + (uint64_t)(k % 3) * P.hist_stride;
    const uint32_t* at = P.AT + (uint64_t)k * P.wprN;
    // ---- (a) column counts of the users' rows
    for (int i = threadIdx.x; i < WORDS * 32 + 1; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    uint32_t cnt[WORDS];
#pragma unroll
    for (int w = 0; w < WORDS; ++w) cnt[w] = 0;
    uint32_t ucnt = 0;
    for (uint64_t ch = gw; ch < nchunks; ch += nwarps) {
      const uint64_t wi = ch * 32 + lane;
      const uint32_t bits = (wi < P.wprN) ? __ldg(at + wi) : 0u;
      ucnt += __popc(bits);
      uint32_t live = __ballot_sync(0xffffffffu, bits != 0);
      while (live) {
        const int src = __ffs(live) - 1;
        live &= live - 1;
        uint32_t b = __shfl_sync(0xffffffffu, bits, src);
        const uint64_t row0 = (ch * 32 + src) * 32;
        while (b) {
          const int pos = __clz(b);
          b &= ~(0x80000000u >> pos);
          const uint32_t* erow = P.E + (row0 + pos) * P.wprE;
#pragma unroll
          for (int w = 0; w < WORDS; ++w)
            if ((uint64_t)w < P.wprE) cnt[w] += (__ldcg(erow + w) >> (31 - lane)) & 1u;
        }
      }
    }
#pragma unroll
    for (int w = 0; w < WORDS; ++w)
      if (cnt[w]) atomicAdd(&s_hist[w * 32 + lane], cnt[w]);
    ucnt = warp_sum_u32(ucnt);
    if (lane == 0 && ucnt) atomicAdd(&s_hist[WORDS * 32], ucnt);
    __syncthreads();
    for (int i = threadIdx.x; i < WORDS * 32; i += blockDim.x)
      if (s_hist[i] && (uint64_t)i < P.m) atomicAdd(&hist[i], s_hist[i]);
    if (threadIdx.x == 0 && s_hist[WORDS * 32]) atomicAdd(&hist[P.m], s_hist[WORDS * 32]);
    // ---- (b)
    grid.sync();
    // ---- (c) newD_k, identical in every CTA
    if (blockIdx.x == 0) {  // recycle the buffer atom k+2 will use (last read for atom k-1)
      uint32_t* hz = P.hist + (uint64_t)((k + 2) % 3) * P.hist_stride;
      for (uint64_t i = threadIdx.x; i <= P.m; i += blockDim.x) hz[i] = 0;
    }
    const uint32_t usage = __ldcg(hist + P.m);
    if (usage == 0) continue;  // src/bsvd.cpp:499-500 (uniform over the grid)
    const uint32_t half = usage >> 1;  // :502
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    // warp w of the CTA builds words w, w+8, ... : lane l decides bit l
    for (uint64_t w = threadIdx.x >> 5; w < P.wprE; w += blockDim.x >> 5) {
      const uint64_t j = w * 32 + lane;
      const uint32_t dk = P.D[(uint64_t)k * P.wprE + w];
      uint32_t bit = 0;
      if (j < P.m) {
        const uint32_t ce = __ldcg(hist + j);
        const uint32_t dbit = (dk >> (31 - lane)) & 1u;
        const uint32_t weight = dbit ? usage - ce : ce;  // sum of (E_i xor D_k)[j] over users
        bit = weight > half;                             // strict >, :504-506
      }
      const uint32_t nd = __brev(__ballot_sync(0xffffffffu, bit));
      if (lane == 0) {
        s_delta[w] = nd ^ dk;
        if (nd != dk) s_any = 1;
      }
    }
    __syncthreads();
    const int any = s_any;
    if (!any) continue;  // dist(newDk,Dk) == 0, :507
    // ---- (d) patch the users' residual rows: E_i ^= Dk ^ newDk  (:512-520)
    for (uint64_t ch = gw; ch < nchunks; ch += nwarps) {
      const uint64_t wi = ch * 32 + lane;
      const uint32_t bits = (wi < P.wprN) ? __ldg(at + wi) : 0u;
      uint32_t live = __ballot_sync(0xffffffffu, bits != 0);
      while (live) {
        const int src = __ffs(live) - 1;
        live &= live - 1;
        uint32_t b = __shfl_sync(0xffffffffu, bits, src);
        const uint64_t row0 = (ch * 32 + src) * 32;
        while (b) {
          const int pos = __clz(b);
          b &= ~(0x80000000u >> pos);
          uint32_t* erow = P.E + (row0 + pos) * P.wprE;
          for (uint64_t w = lane; w < P.wprE; w += 32) {
            const uint32_t dl = s_delta[w];
            if (dl) erow[w] = __ldcg(erow + w) ^ dl;
          }
        }
      }
    }
    grid.sync();
    // D_k is replaced only now: until the barrier other CTAs may still be deriving the same
    // delta from the old D_k. s_delta stays valid until the next atom's step (c), which is
    // behind the next barrier.
    if (blockIdx.x == 0) {
      for (uint64_t w = threadIdx.x; w < P.wprE; w += blockDim.x) P.D[(uint64_t)k * P.wprE + w] ^= s_delta[w];  // :510
      if (threadIdx.x == 0) atomicAdd(P.changed, 1ull);  // :509
    }
  }
}

template <int WORDS>
static bic_status launch_dict(bic_ctx* c, DictParams& P) {
  const size_t smem = (size_t)(WORDS * 32 + 1 + WORDS) * 4;
  int per_sm = 0;
  BIC_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_update_dictionary<WORDS>, 256, smem));
  if (per_sm < 1) return bic_fail(c, BIC_ERR_CUDA, "update_dictionary: kernel does not fit on an SM");
  if (per_sm > 4) per_sm = 4;
  // enough warps for the scan, never more CTAs than can be co-resident
  uint64_t want = div_up_u64(div_up_u64(P.wprN, 32), 8);
  uint64_t cap = (uint64_t)c->sm_count * per_sm;
  int grid = (int)(want < cap ? (want ? want : 1) : cap);
  void* args[] = {&P};
  BIC_PROF(c, KID_UPDATE_DICT);
  BIC_CUDA(c, cudaLaunchCooperativeKernel((void*)k_update_dictionary<WORDS>, dim3(grid), dim3(256), args, smem, c->stream));
  c->launches++;
  if (c->prof_on) bic_prof_end(c);
  return BIC_OK;
}

bic_status bic_k_transpose_A(bic_ctx* c, const bic_mat* A, uint32_t* AT, uint64_t wprN) {
  const uint64_t n = A->rows;
  const uint64_t nblk = div_up_u64(n, 1024);
  const uint64_t gx = nblk < (uint64_t)c->sm_count * 2 ? nblk : (uint64_t)c->sm_count * 2;
  dim3 grid((unsigned)gx, (unsigned)A->wpr);
  BIC_PROF(c, KID_TRANSPOSE_BITS);
  k_transpose_bits<<<grid, 1024, 0, c->stream>>>(A->d, n, A->wpr, AT, wprN, A->cols, nullptr, nullptr);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

bic_status bic_k_transpose_A_batched(bic_ctx* c, uint64_t n, uint64_t p, const ProbDev* probs, const uint32_t* active,
                                     uint32_t nprob) {
  const uint64_t wprA = div_up_u64(p, 32), wprN = div_up_u64(n, 32);
  const uint64_t nblk = div_up_u64(n, 1024);
  uint64_t gx = div_up_u64((uint64_t)c->sm_count * 2, wprA * nprob);
  if (gx > nblk) gx = nblk;
  if (gx < 1) gx = 1;
  BIC_PROF(c, KID_TRANSPOSE_BITS);
  k_transpose_bits<<<dim3((unsigned)gx, (unsigned)wprA, nprob), 1024, 0, c->stream>>>(nullptr, n, wprA, nullptr, wprN, p, probs, active);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

bic_status bic_k_update_dictionary_v2(bic_ctx* c, bic_mat* E, bic_mat* D, const bic_mat* A, unsigned long long* d_changed);
bic_status bic_k_update_dictionary_v3(bic_ctx* c, bic_mat* E, bic_mat* D, const bic_mat* A, unsigned long long* d_changed);
bool bic_dict_chain_eligible(bic_ctx* c, uint64_t n, uint64_t p, uint64_t wprE);

bic_status bic_k_update_dictionary(bic_ctx* c, bic_mat* E, bic_mat* D, const bic_mat* A, unsigned long long* d_changed) {
  BIC_RANGE("bic:update_dictionary_steepest");
  if (E->rows != A->rows || E->cols != D->cols || A->cols != D->rows)
    return bic_fail(c, BIC_ERR_INVALID, "update_dictionary: shapes must be E n x m, D p x m, A n x p");
  if (c->dict_algo == 2 && bic_dict_chain_eligible(c, E->rows, D->rows, E->wpr)) return bic_k_update_dictionary_v3(c, E, D, A, d_changed);
  if (c->dict_algo != 0) return bic_k_update_dictionary_v2(c, E, D, A, d_changed);
  const uint64_t n = E->rows, p = D->rows, m = E->cols;
  if (n == 0 || p == 0 || m == 0) return BIC_OK;
  if (E->wpr > 128) return bic_fail(c, BIC_ERR_UNSUPPORTED, "update_dictionary: rows wider than 4096 bits");
  const uint64_t wprN = div_up_u64(n, 32);
  const uint64_t hist_stride = (m + 1 + 31) & ~(uint64_t)31;
  // work[2]: AT (p * wprN u32) ; work[3]: hist (3 * hist_stride u32)
  BIC_TRY(bic_scratch_reserve(c, &c->work[2], (size_t)p * wprN * 4));
  BIC_TRY(bic_scratch_reserve(c, &c->work[3], (size_t)3 * hist_stride * 4));
  uint32_t* AT = (uint32_t*)c->work[2].p;
  uint32_t* hist = (uint32_t*)c->work[3].p;
  BIC_CUDA(c, cudaMemsetAsync(hist, 0, (size_t)3 * hist_stride * 4, c->stream));
  BIC_TRY(bic_k_transpose_A(c, A, AT, wprN));
  DictParams P;
  P.E = E->d; P.D = D->d; P.AT = AT; P.hist = hist; P.changed = d_changed;
  P.n = n; P.wprE = E->wpr; P.wprN = wprN; P.m = m; P.hist_stride = hist_stride; P.p = (uint32_t)p;
  const uint64_t wpr = E->wpr;
  if (wpr <= 2) return launch_dict<2>(c, P);
  if (wpr <= 8) return launch_dict<8>(c, P);
  if (wpr <= 32) return launch_dict<32>(c, P);
  return launch_dict<128>(c, P);
}

extern "C" bic_status bic_update_dictionary_steepest(bic_ctx* c, bic_mat* E, bic_mat* D, const bic_mat* A, uint64_t* changed) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !E || !D || !A) return BIC_ERR_INVALID;
  BIC_CUDA(c, cudaMemsetAsync(c->d_scalars, 0, sizeof(uint64_t), c->stream));
  BIC_TRY(bic_k_update_dictionary(c, E, D, A, (unsigned long long*)c->d_scalars));
  BIC_TRY(bic_read_scalars(c, 1));
  if (changed) *changed = c->h_scalars[0];
  return BIC_OK;
}
