// update_dictionary_steepest, second formulation ("histogram first, resolve in order").
// Reference: src/bsvd.cpp:463-527. Same results as dict.cu, far fewer dependent steps.
//
// dict.cu walks the atoms with one grid barrier (and one latency-bound gather) per atom. But the
// vote of atom l only differs from what the iteration-start residual gives when an EARLIER atom
// k < l changed AND some row uses both. So:
//   pass 1 (one ordinary launch, all atoms at once): H[l][j] = sum over users i of atom l of E_i[j],
//           usage[l] = number of users -- read from the bit-transposed coefficient matrix AT.
//   pass 2 (one cooperative launch): for k = 0..p-1 in order, every CTA derives newD_k from H[k]
//           (weights[j] = D_k[j] ? usage - H[k][j] : H[k][j]; bit = weights[j] > usage/2). If the atom
//           does not change nothing else happens -- no barrier. If it changes by delta = D_k ^ newD_k,
//           the grid patches E_i ^= delta for the users i of k and, for every later atom l > k that
//           row i also uses and every bit j of delta, corrects H[l][j] by +1 (E_i[j] was 0) or -1
//           (it was 1); then ONE grid barrier. The invariant "H[l] is the column count of atom l's
//           users in the CURRENT E" therefore holds whenever atom l is resolved: exactly the
//           Gauss-Seidel order of the reference, with barriers only for atoms that change.
// Integer atomics are order independent, so the result is deterministic and bit exact.
#include "bic_internal.cuh"

#include <mutex>


// ------------------------------------------------------------------ pass 1
// H = A^T * E as integer counts: H[k][j] = sum_i A[i][k] * E[i][j] = popc(column k of A AND column j
// of E). A warp takes a block of 32 consecutive rows: lane r loads A word kw and JW words of E of
// row r (coalesced), both 32x32 bit tiles are transposed in registers (shuffle butterfly), so lane
// l then holds, for atom kw*32+l, the 32 rows' usage bits, and for bit column l of every E word the
// 32 rows' residual bits. Atom words are broadcast with shuffles and every lane accumulates
// popc(users_k & column_j) for its 32 x JW (k, j) pairs: AND + POPC on the XU pipe, no divergence,
// no gather. Tiles (kw, JW-word group of E) are spread over blockIdx.y for larger p and m.
template <int JW>
__global__ void __launch_bounds__(256) k_dict_hist_popc(const uint32_t* __restrict__ E, const uint32_t* __restrict__ A,
                                                        uint32_t* __restrict__ H, uint32_t* __restrict__ U, uint64_t n,
                                                        uint64_t wprE, uint64_t wprA, uint64_t hs, uint32_t ntile_j,
                                                        const ProbDev* __restrict__ probs, const uint32_t* __restrict__ active,
                                                        uint32_t* __restrict__ listA, uint32_t* __restrict__ listE,
                                                        uint32_t* __restrict__ list_count, uint32_t* __restrict__ hcount,
                                                        const uint32_t* __restrict__ skip) {
  if (skip && *skip) return;  // the learner's loop already ended on the device (queued-ahead iteration)
  if (probs) {  // batched launch: blockIdx.z selects the problem
    if (!active[blockIdx.z]) return;
    const ProbDev pr = probs[blockIdx.z];
    E = pr.E; A = pr.A; H = pr.H; U = pr.U;
  }
  __shared__ uint32_t s_acc[32 * JW * 32];
  __shared__ uint32_t s_u[32];
  __shared__ uint32_t s_h[32];
  uint32_t hcnt = 0;
  const int lane = threadIdx.x & 31;
  const uint32_t kw = blockIdx.y / ntile_j, jt = blockIdx.y - kw * ntile_j;
  const uint64_t jw0 = (uint64_t)jt * JW;
  for (int i = threadIdx.x; i < 32 * JW * 32; i += blockDim.x) s_acc[i] = 0;
  if (threadIdx.x < 32) { s_u[threadIdx.x] = 0; s_h[threadIdx.x] = 0; }
  __syncthreads();
  uint32_t acc[32][JW];
#pragma unroll
  for (int k = 0; k < 32; ++k)
#pragma unroll
    for (int w = 0; w < JW; ++w) acc[k][w] = 0;
  uint32_t ucnt = 0;
  const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t nblocks = div_up_u64(n, 32);
  // software pipeline: the next block's words are requested before the current block is processed
  uint64_t blk = gw;
  uint32_t a_n = 0, e_n[JW];
#pragma unroll
  for (int w = 0; w < JW; ++w) e_n[w] = 0;
  if (blk < nblocks) {
    const uint64_t row = blk * 32 + lane;
    if (row < n) {
      a_n = __ldg(A + row * wprA + kw);
#pragma unroll
      for (int w = 0; w < JW; ++w) e_n[w] = (jw0 + w < wprE) ? __ldg(E + row * wprE + jw0 + w) : 0u;
    }
  }
  while (blk < nblocks) {
    const uint32_t a = a_n;
    uint32_t e[JW];
#pragma unroll
    for (int w = 0; w < JW; ++w) e[w] = e_n[w];
    blk += nwarps;
    a_n = 0;
#pragma unroll
    for (int w = 0; w < JW; ++w) e_n[w] = 0;
    if (blk < nblocks) {
      const uint64_t row = blk * 32 + lane;
      if (row < n) {
        a_n = __ldg(A + row * wprA + kw);
#pragma unroll
        for (int w = 0; w < JW; ++w) e_n[w] = (jw0 + w < wprE) ? __ldg(E + row * wprE + jw0 + w) : 0u;
      }
    }
    if (listA) {
      // on the side (single-tile shapes only: this warp holds the whole rows): rows using two or more atoms
      // are appended to the list the cluster chain of dict3.cu walks
      const bool multi = __popc(a) >= 2;
      const uint32_t bal = __ballot_sync(0xffffffffu, multi);
      if (bal) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(list_count, (uint32_t)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (multi) {
          const uint64_t pos = base + __popc(bal & ((1u << lane) - 1));
          listA[pos] = a;
#pragma unroll
          for (int w = 0; w < JW; ++w)
            if ((uint64_t)w < wprE) listE[pos * wprE + w] = e[w];
        }
        // bucket sizes of the chain: per atom, the rows that use it and a later atom (a & (a - 1) drops
        // the row's last atom, the lowest-valued set bit in MSB-first order)
        hcnt += __popc(warp_transpose32(a & (a - 1)));
      }
    }
    if (!__any_sync(0xffffffffu, a != 0)) continue;  // none of these 32 rows uses these 32 atoms
    uint32_t et[JW];
#pragma unroll
    for (int w = 0; w < JW; ++w) et[w] = warp_transpose32(e[w]);
    const uint32_t at = warp_transpose32(a);
    ucnt += __popc(at);
    // With a large dictionary a block of 32 rows uses only a handful of these 32 atoms (rows are sparse codes): then
    // only the atoms present are visited and their counts go straight to the CTA's shared-memory counters (lane-distinct
    // banks); otherwise all 32 atoms are accumulated in registers.
    uint32_t present = __ballot_sync(0xffffffffu, at != 0);
    if (__popc(present) <= 12) {
      while (present) {
        const int k = __ffs(present) - 1;
        present &= present - 1;
        const uint32_t ak = __shfl_sync(0xffffffffu, at, k);
#pragma unroll
        for (int w = 0; w < JW; ++w) {
          const uint32_t cnt = __popc(ak & et[w]);
          if (cnt) atomicAdd(&s_acc[(k * JW + w) * 32 + lane], cnt);
        }
      }
      continue;
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const uint32_t ak = __shfl_sync(0xffffffffu, at, k);
#pragma unroll
      for (int w = 0; w < JW; ++w) acc[k][w] += __popc(ak & et[w]);
    }
  }
  // CTA reduction in shared memory (lane-distinct banks), then one global atomic per nonzero count
#pragma unroll
  for (int k = 0; k < 32; ++k)
#pragma unroll
    for (int w = 0; w < JW; ++w)
      if (acc[k][w]) atomicAdd(&s_acc[(k * JW + w) * 32 + lane], acc[k][w]);
  if (jt == 0 && ucnt) atomicAdd(&s_u[lane], ucnt);
  if (hcnt) atomicAdd(&s_h[lane], hcnt);
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * JW * 32; i += blockDim.x) {
    const uint32_t v = s_acc[i];
    if (v) {
      const int l = i & 31, kwv = i >> 5, w = kwv % JW, k = kwv / JW;
      atomicAdd(&H[((uint64_t)kw * 32 + k) * hs + (jw0 + w) * 32 + l], v);
    }
  }
  if (jt == 0 && threadIdx.x < 32 && s_u[threadIdx.x]) atomicAdd(&U[kw * 32 + threadIdx.x], s_u[threadIdx.x]);
  if (hcount && threadIdx.x < 32 && s_h[threadIdx.x]) atomicAdd(&hcount[threadIdx.x], s_h[threadIdx.x]);
}

// ------------------------------------------------------------------ pass 2
// One ordinary launch resolves atoms in order from a device-side cursor up to and including the
// FIRST atom that changes, applies that atom's corrections, and stores the next cursor. Launches
// of one stream run in order, so the next launch sees the corrected histograms: the grid barrier
// of a cooperative kernel is replaced by the launch boundary (cooperative launches from several
// streams -- one per page in flight -- serialise against everything else on the GPU).
// A launch whose cursor already reached p returns at once, so the host may queue a few launches
// ahead without knowing how many atoms will change.
struct ResolveParams {
  uint32_t* E;
  const uint32_t* D;     // the dictionary as it was when the update started (never written here)
  uint32_t* Dnew;        // receives the changed atoms
  const uint32_t* A;
  const uint32_t* AT;
  uint32_t* H;
  uint32_t* Hc;          // where corrections for later atoms are accumulated (H itself, or a per-rank delta buffer)
  const uint32_t* U;
  unsigned long long* changed;
  uint32_t* cursor;      // [2]: this launch reads [parity], writes [parity ^ 1]
  uint32_t* first;       // [2]: scan/fix variant: first changing atom found by the scan of this parity
  uint64_t n, wprE, wprA, wprN, m, hs;
  uint32_t p, parity, win;  // win: histogram rows cached in shared memory per refill
  uint32_t dmax;            // k_dict_fix: bits of the delta whose corrections a CTA sums in shared memory (p * dmax counters)
  uint32_t corr_smem;       // 1: k_dict_resolve_step sums the corrections per CTA in shared memory (p*hs words after the window)
  const ProbDev* probs;     // batched launch: blockIdx.y selects the problem (else null)
  const uint32_t* active;
  XPeers x;                 // row-sharded fit with peer-memory exchange: corrections go to EVERY rank's H (else nranks <= 1)
};

// batched launch: swap in the problem's pointers; false = this problem already converged
__device__ __forceinline__ bool resolve_select_problem(ResolveParams& P) {
  if (!P.probs) return true;
  if (!P.active[blockIdx.y]) return false;
  const ProbDev pr = P.probs[blockIdx.y];
  P.E = pr.E; P.D = pr.D; P.Dnew = pr.Dnew; P.A = pr.A; P.AT = pr.AT; P.H = pr.H; P.Hc = pr.H; P.U = pr.U;
  P.changed = pr.counts + 1; P.cursor = pr.cursor; P.first = pr.first;
  return true;
}

// Atom k changes by s_delta = D_k ^ newD_k (shared memory, same in every CTA): patch its users'
// residual rows, correct the histograms of the later atoms those rows use, publish newD_k and the cursor.
// one correction: to the local buffer, or -- row-sharded fit with peer windows -- straight into every rank's histograms
// (remote atomics over NVLink; integer sums, so the order in which the ranks' contributions land does not matter)
__device__ __forceinline__ void hc_add(const ResolveParams& P, uint64_t idx, uint32_t v) {
  if (P.x.nranks > 1) {
    for (uint32_t r = 0; r < P.x.nranks; ++r) atomicAdd(P.x.win[r] + P.x.h_off + idx, v);
  } else {
    atomicAdd(P.Hc + idx, v);
  }
}

// Barrier over the ranks at the end of a kernel that pushed into the peers' windows: every thread fences its remote atomics,
// the LAST CTA of the local grid tells every peer "rank r reached barrier e" and waits until every peer has said the same.
// When the kernel ends, every rank's contributions to this rank's window have landed. One spinning thread per GPU; the wait is
// bounded (about half a minute) and traps instead of hanging if a peer never arrives.
__device__ __forceinline__ void xgpu_barrier(const XPeers& x) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t* me = x.win[x.rank];
    const uint32_t ncta = gridDim.x * gridDim.y * gridDim.z;
    if (atomicAdd(me + 64, 1u) == ncta - 1) {
      atomicExch(me + 64, 0u);
      __threadfence_system();
      for (uint32_t r = 0; r < x.nranks; ++r)
        if (r != x.rank) *(volatile uint32_t*)(x.win[r] + x.rank) = x.epoch;
      for (uint32_t r = 0; r < x.nranks; ++r) {
        if (r == x.rank) continue;
        uint32_t spins = 0;
        while ((int32_t)(*(volatile uint32_t*)(me + r) - x.epoch) < 0) {
          __nanosleep(200);
          if (++spins > (1u << 27)) __trap();  // ~30 s: a peer that never arrives ends in an error, not a hang
        }
      }
      __threadfence_system();
    }
  }
}

// s_cc / s_rank / s_pos (k_dict_fix, large dictionaries): the histogram does not fit shared memory, but only the bits of
// the delta are ever corrected. They are numbered 0..nd-1 (s_rank[w] = bits before word w, s_pos[r] = bit position) and the
// first P.dmax of them get per-CTA counters s_cc[l * dmax + r]; one global atomic per touched counter and CTA at the end.
__device__ __forceinline__ void dict_apply_change(const ResolveParams& P, uint32_t k, const uint32_t* s_delta, uint32_t* s_corr,
                                                  uint32_t* s_cc = nullptr, const uint32_t* s_rank = nullptr,
                                                  const uint32_t* s_pos = nullptr) {
  const int lane = threadIdx.x & 31;
  const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t* at = P.AT + (uint64_t)k * P.wprN;
  if (P.wprE > 8) {
    // wide rows: a WARP per user row, lanes over the row's words (coalesced), so a 1024-bit row is 32
    // lanes x 1 word instead of one lane x 32 words. AT words are spread over warps one by one.
    for (uint64_t wi = gw; wi < P.wprN; wi += nwarps) {
      uint32_t bits = __ldg(at + wi);
      while (bits) {  // warp-uniform
        const int pos = __clz(bits);
        bits &= ~(0x80000000u >> pos);
        const uint64_t i = wi * 32 + pos;
        uint32_t* erow = P.E + i * P.wprE;
        const uint32_t* arow = P.A + i * P.wprA;
        for (uint64_t w0 = 0; w0 < P.wprE; w0 += 32) {
          const uint64_t w = w0 + lane;
          const uint32_t dl0 = (w < P.wprE) ? s_delta[w] : 0u;
          const uint32_t e = (dl0 != 0) ? erow[w] : 0u;
          if (__any_sync(0xffffffffu, dl0 != 0)) {
            for (uint64_t aw = k >> 5; aw < P.wprA; ++aw) {   // later atoms used by this row (warp-uniform)
              uint32_t ab = __ldg(arow + aw);
              if (aw == (k >> 5)) ab &= (0x7FFFFFFFu >> (k & 31));
              while (ab) {
                const int ap = __clz(ab);
                ab &= ~(0x80000000u >> ap);
                uint32_t* hl = P.Hc + (aw * 32 + ap) * P.hs + w * 32;
                uint32_t dl = dl0;
                while (dl) {
                  const int bp = __clz(dl);
                  dl &= ~(0x80000000u >> bp);
                  const uint32_t v = ((e >> (31 - bp)) & 1u) ? 0xFFFFFFFFu : 1u;
                  const uint32_t r = s_cc ? s_rank[w] + __popc(dl0 & ~(0xFFFFFFFFu >> bp)) : 0xFFFFFFFFu;
                  if (r < P.dmax) atomicAdd(&s_cc[(aw * 32 + ap) * P.dmax + r], v);
                  else hc_add(P, (uint64_t)(hl - P.Hc) + bp, v);
                }
              }
            }
          }
          if (dl0) erow[w] = e ^ dl0;                         // E_i ^= Dk ^ newDk, :512-520
        }
      }
    }
  } else {
    // narrow rows: a lane per user row. The users of a group of 32 AT words (1024 rows) are first listed
    // in a per-warp shared-memory queue so the lanes get an even share (two or three rows each): the pass
    // is a chain of dependent L2 round trips per row, so its duration is the LONGEST lane's chain.
    __shared__ uint16_t s_q[8][1024];  // row index within the group
    uint16_t* q = s_q[threadIdx.x >> 5];
    const uint64_t ngroups = div_up_u64(P.wprN, 32);
    for (uint64_t g = gw; g < ngroups; g += nwarps) {
      const uint64_t wi = g * 32 + lane;
      uint32_t bits = (wi < P.wprN) ? __ldg(at + wi) : 0u;
      const uint32_t cnt = __popc(bits);
      uint32_t incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
      uint32_t off = incl - cnt;
      while (bits) {
        const int pos = __clz(bits);
        bits &= ~(0x80000000u >> pos);
        q[off++] = (uint16_t)(lane * 32 + pos);
      }
      __syncwarp();
      for (uint32_t t = lane; t < total; t += 32) {
        const uint64_t i = g * 1024 + q[t];
        uint32_t* erow = P.E + i * P.wprE;
        const uint32_t* arow = P.A + i * P.wprA;
        for (uint64_t aw = k >> 5; aw < P.wprA; ++aw) {   // later atoms used by this row
          uint32_t ab = __ldg(arow + aw);
          if (aw == (k >> 5)) ab &= (0x7FFFFFFFu >> (k & 31));  // strictly after k
          while (ab) {
            const int ap = __clz(ab);
            ab &= ~(0x80000000u >> ap);
            // corrections are summed per CTA in shared memory when the histogram fits there (one global
            // atomic per touched counter per CTA instead of one per row and bit), else go straight to L2
            uint32_t* hl = (s_corr ? s_corr : P.Hc) + (aw * 32 + ap) * P.hs;
            for (uint64_t w = 0; w < P.wprE; ++w) {
              const uint32_t dl0 = s_delta[w];
              uint32_t dl = dl0;
              if (!dl) continue;
              const uint32_t e = erow[w];
              while (dl) {
                const int bp = __clz(dl);
                dl &= ~(0x80000000u >> bp);
                const uint32_t v = ((e >> (31 - bp)) & 1u) ? 0xFFFFFFFFu : 1u;
                const uint32_t r = s_cc ? s_rank[w] + __popc(dl0 & ~(0xFFFFFFFFu >> bp)) : 0xFFFFFFFFu;
                if (r < P.dmax) atomicAdd(&s_cc[(aw * 32 + ap) * P.dmax + r], v);
                else if (s_corr) atomicAdd(hl + w * 32 + bp, v);
                else hc_add(P, (aw * 32 + ap) * P.hs + w * 32 + bp, v);
              }
            }
          }
        }
        for (uint64_t w = 0; w < P.wprE; ++w) {           // E_i ^= Dk ^ newDk, :512-520
          const uint32_t dl = s_delta[w];
          if (dl) erow[w] ^= dl;
        }
      }
      __syncwarp();
    }
    if (s_corr) {
      __syncthreads();
      const uint64_t lo = (uint64_t)(k + 1) * P.hs, hi = (uint64_t)P.p * P.hs;
      for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const uint32_t v = s_corr[i];
        if (v) hc_add(P, i, v);
      }
    }
  }
  if (s_cc) {
    __syncthreads();
    for (uint32_t i = (k + 1) * P.dmax + threadIdx.x; i < P.p * P.dmax; i += blockDim.x) {
      const uint32_t v = s_cc[i];
      if (v) hc_add(P, (uint64_t)(i / P.dmax) * P.hs + s_pos[i % P.dmax], v);
    }
  }
  if (blockIdx.x == 0) {
    for (uint64_t w = threadIdx.x; w < P.wprE; w += blockDim.x)
      P.Dnew[(uint64_t)k * P.wprE + w] = __ldg(P.D + (uint64_t)k * P.wprE + w) ^ s_delta[w];  // :510
    if (threadIdx.x == 0) {
      atomicAdd(P.changed, 1ull);                       // :509
      P.cursor[P.parity ^ 1] = k + 1;
    }
  }
  if (P.x.nranks > 1) xgpu_barrier(P.x);                // every rank's corrections for atom k are in this rank's H
}

__global__ void __launch_bounds__(256) k_dict_resolve_step(ResolveParams P) {
  if (!resolve_select_problem(P)) return;
  extern __shared__ uint32_t s_mem[];
  uint32_t* s_delta = s_mem;              // wprE
  uint32_t* s_U = s_delta + P.wprE;       // win
  uint32_t* s_H = s_U + P.win;            // win * hs
  __shared__ int s_any;
  const int lane = threadIdx.x & 31;
  const uint32_t start = __ldcg(P.cursor + P.parity);
  if (start >= P.p) {
    if (blockIdx.x == 0 && threadIdx.x == 0) P.cursor[P.parity ^ 1] = P.p;
    return;
  }
  uint32_t k = start;
  uint32_t win0 = start, win1 = start;    // cached rows [win0, win1)
  int any = 0;
  for (; k < P.p; ++k) {
    if (k >= win1) {                      // refill: rows <= the first changing atom are stable
      __syncthreads();
      win0 = k;
      win1 = (k + P.win < P.p) ? k + P.win : P.p;
      for (uint64_t i = threadIdx.x; i < (uint64_t)(win1 - win0) * P.hs; i += blockDim.x)
        s_H[i] = __ldcg(P.H + (uint64_t)win0 * P.hs + i);
      for (uint32_t i = threadIdx.x; i < win1 - win0; i += blockDim.x) s_U[i] = __ldcg(P.U + win0 + i);
      __syncthreads();
    }
    const uint32_t usage = s_U[k - win0];
    if (usage == 0) continue;             // src/bsvd.cpp:499-500
    const uint32_t half = usage >> 1;     // :502
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    const uint32_t* hk = s_H + (uint64_t)(k - win0) * P.hs;
    for (uint64_t w = threadIdx.x >> 5; w < P.wprE; w += blockDim.x >> 5) {
      const uint64_t j = w * 32 + lane;
      const uint32_t dk = __ldg(P.D + (uint64_t)k * P.wprE + w);
      uint32_t bit = 0;
      if (j < P.m) {
        const uint32_t ce = hk[j];
        const uint32_t weight = ((dk >> (31 - lane)) & 1u) ? usage - ce : ce;  // sum of (E_i ^ D_k)[j] over users
        bit = weight > half;              // strict >, :504-506
      }
      const uint32_t nd = __brev(__ballot_sync(0xffffffffu, bit));
      if (lane == 0) {
        s_delta[w] = nd ^ dk;
        if (nd != dk) s_any = 1;
      }
    }
    __syncthreads();
    any = s_any;
    __syncthreads();                      // s_any / s_delta are rewritten at the next atom
    if (any) break;                       // dist(newDk, Dk) > 0, :507 (same decision in every CTA)
  }
  if (!any) {
    if (blockIdx.x == 0 && threadIdx.x == 0) P.cursor[P.parity ^ 1] = P.p;
    return;
  }
  uint32_t* s_corr = nullptr;
  if (P.corr_smem) {  // zero the CTA's correction counters for the atoms after k
    s_corr = s_H + (uint64_t)P.win * P.hs;
    for (uint64_t i = (uint64_t)(k + 1) * P.hs + threadIdx.x; i < (uint64_t)P.p * P.hs; i += blockDim.x) s_corr[i] = 0;
    __syncthreads();
  }
  dict_apply_change(P, k, s_delta, s_corr);
}

// ------------------------------------------------------------------ pass 2 for large dictionaries
// With many atoms the serial walk above (every CTA re-derives every atom) is the bottleneck. Instead:
//   k_dict_scan : all atoms at or after the cursor are tested IN PARALLEL against the current H
//                 ("would this atom change?"); atomicMin keeps the first one.
//   k_dict_fix  : that atom is the next to change in the reference's order (every earlier one is
//                 unchanged under the same H); the grid applies its corrections and advances the cursor.
// Atoms after the first changing one are re-tested by the next scan, after the corrections landed.
__global__ void __launch_bounds__(256) k_dict_scan(ResolveParams P) {
  if (!resolve_select_problem(P)) return;
  const int lane = threadIdx.x & 31;
  const uint32_t start = __ldcg(P.cursor + P.parity);
  if (start >= P.p) return;
  const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t k = start + gw; k < P.p; k += nwarps) {
    if (k >= __ldcg(P.first + P.parity)) break;  // an earlier atom already changes
    const uint32_t usage = __ldg(P.U + k);
    if (usage == 0) continue;
    const uint32_t half = usage >> 1;
    const uint32_t* hk = P.H + k * P.hs;
    bool any = false;
    for (uint64_t w = 0; w < P.wprE; ++w) {
      const uint64_t j = w * 32 + lane;
      const uint32_t dk = __ldg(P.D + k * P.wprE + w);
      uint32_t bit = 0;
      if (j < P.m) {
        const uint32_t ce = __ldcg(hk + j);
        const uint32_t weight = ((dk >> (31 - lane)) & 1u) ? usage - ce : ce;
        bit = weight > half;
      }
      const uint32_t nd = __brev(__ballot_sync(0xffffffffu, bit));
      any |= (nd != dk);
    }
    if (any && lane == 0) atomicMin(P.first + P.parity, (uint32_t)k);
  }
}

__global__ void __launch_bounds__(256) k_dict_fix(ResolveParams P) {
  if (!resolve_select_problem(P)) return;
  extern __shared__ uint32_t s_delta[];  // wprE
  const int lane = threadIdx.x & 31;
  const uint32_t start = __ldcg(P.cursor + P.parity);
  const uint32_t k = __ldcg(P.first + P.parity);
  if (blockIdx.x == 0 && threadIdx.x == 0) P.first[P.parity ^ 1] = P.p;  // arm the next scan
  if (start >= P.p || k >= P.p) {
    if (blockIdx.x == 0 && threadIdx.x == 0) P.cursor[P.parity ^ 1] = P.p;
    return;
  }
  const uint32_t usage = __ldg(P.U + k);
  const uint32_t half = usage >> 1;
  const uint32_t* hk = P.H + (uint64_t)k * P.hs;
  for (uint64_t w = threadIdx.x >> 5; w < P.wprE; w += blockDim.x >> 5) {
    const uint64_t j = w * 32 + lane;
    const uint32_t dk = __ldg(P.D + (uint64_t)k * P.wprE + w);
    uint32_t bit = 0;
    if (j < P.m) {
      const uint32_t ce = __ldcg(hk + j);
      const uint32_t weight = ((dk >> (31 - lane)) & 1u) ? usage - ce : ce;
      bit = weight > half;
    }
    const uint32_t nd = __brev(__ballot_sync(0xffffffffu, bit));
    if (lane == 0) s_delta[w] = nd ^ dk;
  }
  __syncthreads();
  uint32_t* s_rank = s_delta + P.wprE;      // wprE
  uint32_t* s_pos = s_rank + P.wprE;        // dmax
  uint32_t* s_cc = s_pos + P.dmax;          // p * dmax
  if (P.dmax) {
    if (threadIdx.x == 0) {
      uint32_t r = 0;
      for (uint64_t w = 0; w < P.wprE; ++w) {
        s_rank[w] = r;
        uint32_t dl = s_delta[w];
        while (dl) {
          const int bp = __clz(dl);
          dl &= ~(0x80000000u >> bp);
          if (r < P.dmax) s_pos[r] = (uint32_t)w * 32 + bp;
          ++r;
        }
      }
    }
    for (uint32_t i = (k + 1) * P.dmax + threadIdx.x; i < P.p * P.dmax; i += blockDim.x) s_cc[i] = 0;
    __syncthreads();
    dict_apply_change(P, k, s_delta, nullptr, s_cc, s_rank, s_pos);
  } else {
    dict_apply_change(P, k, s_delta, nullptr);
  }
}

// k_dict_fix's shared memory: delta, rank per word, positions, and p * dmax correction counters (at most ~40 KB)
static void fix_set_dmax(ResolveParams& P) {
  uint64_t d = 10240 / (P.p ? P.p : 1);
  P.dmax = (uint32_t)(d > 64 ? 64 : d);
}
static size_t fix_smem_bytes(const ResolveParams& P) {
  return (size_t)(2 * P.wprE + P.dmax + (uint64_t)P.p * P.dmax) * 4;
}

static bic_status launch_hist(bic_ctx* c, const bic_mat* E, const bic_mat* A, uint32_t* H, uint32_t* U, uint64_t hs,
                              uint32_t* listA = nullptr, uint32_t* listE = nullptr, uint32_t* count = nullptr, uint32_t* hcount = nullptr) {
  const uint64_t n = E->rows;
  const bool two = E->wpr >= 2;
  const uint32_t ntile_j = (uint32_t)(two ? div_up_u64(E->wpr, 2) : 1);
  const uint32_t ntiles = (uint32_t)A->wpr * ntile_j;
  // persistent in x: about two CTAs per SM overall, at least one block of rows per warp
  uint64_t gx = div_up_u64((uint64_t)c->sm_count * 2, ntiles);
  const uint64_t need = div_up_u64(div_up_u64(n, 32), 8);
  if (gx > need) gx = need;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, ntiles);
  BIC_PROF(c, KID_DICT_HIST);
  if (two) k_dict_hist_popc<2><<<grid, 256, 0, c->stream>>>(E->d, A->d, H, U, n, E->wpr, A->wpr, hs, ntile_j, nullptr, nullptr, listA, listE, count, hcount, c->loop_skip);
  else k_dict_hist_popc<1><<<grid, 256, 0, c->stream>>>(E->d, A->d, H, U, n, E->wpr, A->wpr, hs, ntile_j, nullptr, nullptr, listA, listE, count, hcount, c->loop_skip);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

// histogram pass for dict3.cu; the multi-atom row list is built on the side when a warp sees whole rows
bic_status bic_k_dict_hist_compact(bic_ctx* c, const bic_mat* E, const bic_mat* A, uint32_t* H, uint32_t* U, uint64_t hs,
                                   uint32_t* listA, uint32_t* listE, uint32_t* count, uint32_t* hcount, bool* fused) {
  *fused = (A->wpr == 1 && E->wpr <= 2);
  if (*fused) return launch_hist(c, E, A, H, U, hs, listA, listE, count, hcount);
  return launch_hist(c, E, A, H, U, hs);
}

bic_status bic_k_transpose_A(bic_ctx* c, const bic_mat* A, uint32_t* AT, uint64_t wprN);

// window (<= 32 KB) + correction counters (<= 16 KB) + the static user queue (16 KB) can pass the 48 KB default
static bic_status resolve_step_smem_optin(bic_ctx* c) {
  static bool done[64] = {false};
  static std::mutex mu;  // contexts of several host threads come through here at once
  std::lock_guard<std::mutex> lk(mu);
  if (c->device < 64 && done[c->device]) return BIC_OK;
  BIC_CUDA(c, cudaFuncSetAttribute(k_dict_resolve_step, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  BIC_CUDA(c, cudaFuncSetAttribute(k_dict_fix, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));  // + 16 KB static
  if (c->device < 64) done[c->device] = true;
  return BIC_OK;
}

// The update as three reusable stages, so the row-sharded driver (dist.cu) can put its collectives
// between them: (1) prepare: AT, local H/U; (2) resolve steps; (3) commit Dnew -> D.
bic_status bic_k_dict_prepare(bic_ctx* c, bic_mat* E, bic_mat* D, const bic_mat* A, DictWork* w, uint32_t* hbase) {
  if (E->rows != A->rows || E->cols != D->cols || A->cols != D->rows)
    return bic_fail(c, BIC_ERR_INVALID, "update_dictionary: shapes must be E n x m, D p x m, A n x p");
  const uint64_t n = E->rows, p = D->rows, wpr = E->wpr;
  w->n = n; w->p = p; w->wpr = wpr; w->hs = wpr * 32; w->wprN = div_up_u64(n, 32);
  // work[2]: AT (p * wprN u32) ; work[3]: H (p*hs) | U (p) | extra (64) | Hd (p*hs) ; work[0]: Dnew (p*wpr) | cursor
  BIC_TRY(bic_scratch_reserve(c, &c->work[2], (size_t)p * w->wprN * 4 + 16));
  BIC_TRY(bic_scratch_reserve(c, &c->work[3], (size_t)(2 * p * w->hs + p + 64) * 4));
  BIC_TRY(bic_scratch_reserve(c, &c->work[0], (size_t)p * wpr * 4 + 64));
  w->AT = (uint32_t*)c->work[2].p;
  w->x = XPeers();
  if (hbase) {  // the caller's buffer (a peer window) holds [H | U | extra]; the delta buffer stays in the scratch
    w->H = hbase;
    w->U = w->H + p * w->hs;
    w->extra = w->U + p;
    w->Hd = (uint32_t*)c->work[3].p;
  } else {
    w->H = (uint32_t*)c->work[3].p;
    w->U = w->H + p * w->hs;
    w->extra = w->U + p;
    w->Hd = w->extra + 64;
  }
  w->Dnew = (uint32_t*)c->work[0].p;
  w->cursor = w->Dnew + p * wpr;
  w->first = w->cursor + 2;
  w->launched = 0;
  // a dictionary whose histogram fits one shared-memory window is walked serially inside one launch;
  // larger ones use the parallel scan + fix pair
  w->use_scan = (p * (w->hs + 1) * 4 > 32 * 1024);
  if (hbase) {
    BIC_CUDA(c, cudaMemsetAsync(w->H, 0, (size_t)(p * w->hs + p + 64) * 4, c->stream));
  } else {
    BIC_CUDA(c, cudaMemsetAsync(w->H, 0, (size_t)(2 * p * w->hs + p + 64) * 4, c->stream));
  }
  BIC_CUDA(c, cudaMemcpyAsync(w->Dnew, D->d, (size_t)p * wpr * 4, cudaMemcpyDeviceToDevice, c->stream));
  {
    const uint32_t init[4] = {0u, 0u, (uint32_t)p, (uint32_t)p};  // cursor[2], first[2]
    BIC_CUDA(c, cudaMemcpyAsync(w->cursor, init, 16, cudaMemcpyHostToDevice, c->stream));
  }
  if (n) {
    BIC_TRY(bic_k_transpose_A(c, A, w->AT, w->wprN));
    BIC_TRY(launch_hist(c, E, A, w->H, w->U, w->hs));
  }
  return BIC_OK;
}

// one resolve launch; corrections go to `Hc` (w->H in place, or w->Hd for the sharded driver)
bic_status bic_k_dict_step(bic_ctx* c, bic_mat* E, const bic_mat* D, const bic_mat* A, DictWork* w, uint32_t* Hc,
                           unsigned long long* d_changed) {
  ResolveParams P;
  P.E = E->d; P.D = D->d; P.Dnew = w->Dnew; P.A = A->d; P.AT = w->AT; P.H = w->H; P.Hc = Hc; P.U = w->U;
  P.changed = d_changed; P.cursor = w->cursor; P.first = w->first; P.probs = nullptr; P.active = nullptr;
  P.x = w->x;
  if (P.x.nranks > 1) P.x.epoch += w->launched + 1;  // one barrier number per launch (a launch that changes no atom skips its barrier)
  P.n = w->n; P.wprE = w->wpr; P.wprA = A->wpr; P.wprN = w->wprN; P.m = E->cols; P.hs = w->hs; P.p = (uint32_t)w->p;
  uint64_t win = (32 * 1024 / 4) / (w->hs + 1);
  if (win < 1) win = 1;
  if (win > w->p) win = w->p;
  P.win = (uint32_t)win;
  fix_set_dmax(P);
  P.corr_smem = (!w->use_scan && w->wpr < 8 && w->p * w->hs <= 4096) ? 1u : 0u;
  const size_t smem = (size_t)(w->wpr + win * (w->hs + 1) + (P.corr_smem ? w->p * w->hs : 0)) * 4;
  const int grid = (w->wpr > 8) ? bic_grid_for(c, (w->wprN ? w->wprN : 1) * 32, 256, 8)
                                 : bic_grid_for(c, div_up_u64(w->wprN ? w->wprN : 1, 32) * 32, 256, 4);
  P.parity = w->launched & 1;
  if (w->use_scan) {
    BIC_TRY(resolve_step_smem_optin(c));
    const int sgrid = bic_grid_for(c, w->p * 32, 256, 4);
    BIC_PROF(c, KID_DICT_SCAN);
    k_dict_scan<<<sgrid, 256, 0, c->stream>>>(P);
    BIC_LAUNCH_CHECK(c);
    BIC_PROF(c, KID_DICT_RESOLVE);
    k_dict_fix<<<grid, 256, fix_smem_bytes(P), c->stream>>>(P);
    BIC_LAUNCH_CHECK(c);
  } else {
    BIC_TRY(resolve_step_smem_optin(c));
    BIC_PROF(c, KID_DICT_RESOLVE);
    k_dict_resolve_step<<<grid, 256, smem, c->stream>>>(P);
    BIC_LAUNCH_CHECK(c);
  }
  w->launched++;
  return BIC_OK;
}

// ---- batched variants (batch.cu): the same kernels over a problem table
bic_status bic_k_dict_hist_batched(bic_ctx* c, uint64_t n, uint64_t m, uint64_t p, const ProbDev* probs, const uint32_t* active,
                                   uint32_t nprob) {
  const uint64_t wprE = div_up_u64(m, 32), wprA = div_up_u64(p, 32), hs = wprE * 32;
  const bool two = wprE >= 2;
  const uint32_t ntile_j = (uint32_t)(two ? div_up_u64(wprE, 2) : 1);
  const uint32_t ntiles = (uint32_t)wprA * ntile_j;
  uint64_t gx = div_up_u64((uint64_t)c->sm_count * 2, (uint64_t)ntiles * nprob);
  const uint64_t need = div_up_u64(div_up_u64(n, 32), 8);
  if (gx > need) gx = need;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, ntiles, nprob);
  BIC_PROF(c, KID_DICT_HIST);
  if (two) k_dict_hist_popc<2><<<grid, 256, 0, c->stream>>>(nullptr, nullptr, nullptr, nullptr, n, wprE, wprA, hs, ntile_j, probs, active, nullptr, nullptr, nullptr, nullptr, nullptr);
  else k_dict_hist_popc<1><<<grid, 256, 0, c->stream>>>(nullptr, nullptr, nullptr, nullptr, n, wprE, wprA, hs, ntile_j, probs, active, nullptr, nullptr, nullptr, nullptr, nullptr);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

bic_status bic_k_dict_step_batched(bic_ctx* c, uint64_t n, uint64_t m, uint64_t p, const ProbDev* probs, const uint32_t* active,
                                   uint32_t nprob, uint32_t launched) {
  const uint64_t wpr = div_up_u64(m, 32), hs = wpr * 32, wprN = div_up_u64(n, 32);
  ResolveParams P;
  memset(&P, 0, sizeof(P));
  P.n = n; P.wprE = wpr; P.wprA = div_up_u64(p, 32); P.wprN = wprN; P.m = m; P.hs = hs; P.p = (uint32_t)p;
  P.probs = probs; P.active = active;
  uint64_t win = (32 * 1024 / 4) / (hs + 1);
  if (win < 1) win = 1;
  if (win > p) win = p;
  P.win = (uint32_t)win;
  fix_set_dmax(P);
  P.corr_smem = (p * (hs + 1) * 4 <= 32 * 1024 && wpr < 8 && p * hs <= 4096) ? 1u : 0u;
  const size_t smem = (size_t)(wpr + win * (hs + 1) + (P.corr_smem ? p * hs : 0)) * 4;
  const int gx = (wpr > 8) ? bic_grid_for(c, (wprN ? wprN : 1) * 32, 256, 8)
                            : bic_grid_for(c, div_up_u64(wprN ? wprN : 1, 32) * 32, 256, 4);
  P.parity = launched & 1;
  if (p * (hs + 1) * 4 > 32 * 1024) {
    BIC_TRY(resolve_step_smem_optin(c));
    const int sgrid = bic_grid_for(c, p * 32, 256, 4);
    BIC_PROF(c, KID_DICT_SCAN);
    k_dict_scan<<<dim3(sgrid, nprob), 256, 0, c->stream>>>(P);
    BIC_LAUNCH_CHECK(c);
    BIC_PROF(c, KID_DICT_RESOLVE);
    k_dict_fix<<<dim3(gx, nprob), 256, fix_smem_bytes(P), c->stream>>>(P);
    BIC_LAUNCH_CHECK(c);
  } else {
    BIC_TRY(resolve_step_smem_optin(c));
    BIC_PROF(c, KID_DICT_RESOLVE);
    k_dict_resolve_step<<<dim3(gx, nprob), 256, smem, c->stream>>>(P);
    BIC_LAUNCH_CHECK(c);
  }
  return BIC_OK;
}

// cursor after the launches queued so far (synchronises the stream)
bic_status bic_k_dict_cursor(bic_ctx* c, DictWork* w, uint32_t* cursor_out) {
  uint32_t* h_cursor = (uint32_t*)(c->h_scalars + 32);
  BIC_CUDA(c, cudaMemcpyAsync(h_cursor, w->cursor + (w->launched & 1), 4, cudaMemcpyDeviceToHost, c->stream));
  BIC_CUDA(c, bic_wait_stream(c));
  *cursor_out = *h_cursor;
  return BIC_OK;
}

bic_status bic_k_dict_commit(bic_ctx* c, bic_mat* D, DictWork* w) {
  BIC_CUDA(c, cudaMemcpyAsync(D->d, w->Dnew, (size_t)w->p * w->wpr * 4, cudaMemcpyDeviceToDevice, c->stream));
  return BIC_OK;
}

bic_status bic_k_update_dictionary_v2(bic_ctx* c, bic_mat* E, bic_mat* D, const bic_mat* A, unsigned long long* d_changed) {
  if (E->rows == 0 || D->rows == 0 || E->cols == 0) return BIC_OK;
  DictWork w;
  BIC_TRY(bic_k_dict_prepare(c, E, D, A, &w, nullptr));
  // Queue launches ahead; each returns at once when the cursor is already at p. The cursor is read
  // back and more launches follow only if atoms are still pending.
  uint32_t batch = 4, cursor = 0;
  for (;;) {
    for (uint32_t i = 0; i < batch && w.launched < w.p; ++i) BIC_TRY(bic_k_dict_step(c, E, D, A, &w, w.H, d_changed));
    BIC_TRY(bic_k_dict_cursor(c, &w, &cursor));
    if (cursor >= w.p || w.launched >= w.p) break;
    batch = (batch * 2 < 64) ? batch * 2 : 64;
  }
  return bic_k_dict_commit(c, D, &w);
}
