// update_dictionary_steepest, third formulation: the in-order atom chain inside ONE thread-block cluster.
// Reference: src/bsvd.cpp:463-527. Same results as dict.cu / dict2.cu.
//
// dict2.cu resolves the atoms in order with one launch per atom that CHANGES (the launch boundary is
// its grid barrier) and a gather over that atom's users through the transposed coefficient matrix:
// a chain of ~15 us links, 10-32 of them in the first iterations of a fit. Two observations cut the link:
//   * A row that uses a single atom never feeds a later atom's vote. Only rows with >= 2 atoms can
//     carry a change of atom k into the histogram of an atom l > k. Those rows are compacted once
//     per update (by the histogram pass, which holds every row in registers anyway) into a list.
//   * The residual rows need not be patched while the chain runs: the residual of row i as atom k
//     sees it is E_i ^ XOR{delta_k' : k' < k changed, k' in S_i}, recomputed on the fly from the
//     (tiny) table of deltas. The list is read-only; E is patched once, after the chain, by a
//     streaming pass (k_dict_apply).
//   * The list is bucketed by atom once per update (k_dict_bucket_count / k_dict_bucket_fill): bucket k
//     holds the list rows that use atom k AND a later atom, i.e. exactly the rows a change of atom k
//     has to look at, so a change costs its own rows only (no scan of the list per change).
// So the chain needs no global writes and no grid-wide barrier: one cluster of CTAs walks the atoms,
// each CTA scanning its share of the list, correction counts summed per CTA in shared memory and
// exchanged through distributed shared memory, with one hardware cluster barrier per CHANGED atom.
// Every CTA keeps its own copy of the histograms H (p x m counters), so the decisions are taken
// redundantly and identically in every CTA (integer sums, order independent => deterministic).
#include "bic_internal.cuh"

#include <cooperative_groups.h>

#include <mutex>
namespace cg = cooperative_groups;

static const int CHAIN_THREADS = 1024;

struct ChainParams {
  uint32_t* D;              // p x wprE, patched in place at the end (atoms are decided from the old D, :501-506)
  const uint32_t* H;        // p x hs histograms of the iteration-start residual
  const uint32_t* U;        // p usage counts
  const uint32_t* listA;    // rows using >= 2 atoms: coefficient words ...
  const uint32_t* listE;    // ... and residual words
  const uint32_t* count;    // list length
  uint32_t* delta;          // out: p x wprE, D_k ^ newD_k (0 for unchanged atoms)
  uint32_t* chmask;         // out: wprA words, bit k set iff atom k changed; then [wprA] = number of changed atoms
  unsigned long long* changed;
  uint32_t p, wprE, wprA, hs;
  const uint32_t* hcount;   // p: bucket sizes
  const uint32_t* bucket;   // list indices grouped by atom (offsets = exclusive scan of hcount)
  uint32_t bucket_cap;      // capacity of `bucket`; when the buckets do not fit the chain scans the list instead
  uint64_t m;
  const uint32_t* skip;     // non-null and nonzero: the learner's loop already ended on the device, do nothing
  const uint32_t* gcount;   // p: bucket sizes summed over all ranks (== hcount on one GPU): decides whether a change needs an exchange
  ChainDist x;              // nranks > 1: rows are sharded over several GPUs, corrections are exchanged over peer memory
};

// vote of atom k from the histograms in shared memory (src/bsvd.cpp:499-507); one warp. Returns whether
// the atom changes; optionally stores delta = D_k ^ newD_k.
__device__ __forceinline__ bool chain_decide(const uint32_t* sH, const uint32_t* sU, const uint32_t* sD, uint32_t k,
                                             const ChainParams& P, uint32_t* out_delta) {
  const int lane = threadIdx.x & 31;
  const uint32_t usage = sU[k];
  if (usage == 0) {                       // :499-500
    if (out_delta && lane == 0) for (uint32_t w = 0; w < P.wprE; ++w) out_delta[w] = 0;
    return false;
  }
  const uint32_t half = usage >> 1;       // :502
  bool any = false;
  for (uint32_t w = 0; w < P.wprE; ++w) {
    const uint32_t j = w * 32 + lane;
    const uint32_t dk = sD[k * P.wprE + w];
    uint32_t bit = 0;
    if (j < P.m) {
      const uint32_t ce = sH[k * P.hs + j];
      const uint32_t weight = ((dk >> (31 - lane)) & 1u) ? usage - ce : ce;  // sum of (E_i ^ D_k)[j] over users
      bit = weight > half;                // strict >, :504-506
    }
    const uint32_t nd = __brev(__ballot_sync(0xffffffffu, bit));
    any |= (nd != dk);
    if (out_delta && lane == 0) out_delta[w] = nd ^ dk;
  }
  return any;
}

// A batch of up to 32 list rows that use the changed atom k AND a later atom (lane r holds list index
// `idx`, or valid = false): the corrections sum_i [l in S_i] * (1 - 2 * cur_i[j]) for every later atom l
// and every bit j of delta_k, as a bit-matrix product. The 32 rows' later-atom bits and current residual
// bits are transposed in registers (lane l then holds atom l's user bits u_l, lane j the residual column
// c_j), and corr[l][j] += popc(u_l) - 2 * popc(u_l & c_j): one shared-memory atomic per (atom, bit) per
// 32 rows instead of one per row, atom and bit. corr rows are hsC = hs + 1 words apart (lane l -> bank l + j).
__device__ __forceinline__ void chain_batch(const ChainParams& P, uint32_t k, uint32_t idx, bool valid,
                                            const uint32_t* sDelta, const uint32_t* sCh, uint32_t* corr, uint32_t hsC) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t kw = k >> 5;
  BIC_DCHECK(!valid || idx < __ldcg(P.count));                       // a bucket / queue entry is an index into the list
  const uint32_t* a = P.listA + (uint64_t)idx * P.wprA;
  const uint32_t* e = P.listE + (uint64_t)idx * P.wprE;
  for (uint32_t w = 0; w < P.wprE; ++w) {
    const uint32_t dl0 = sDelta[k * P.wprE + w];
    if (!dl0) continue;                     // warp-uniform
    uint32_t cur = 0;
    if (valid) {
      cur = __ldg(e + w);
      for (uint32_t t = 0; t <= kw; ++t) {  // the row as atom k sees it: earlier changed atoms already applied
        uint32_t eb = __ldg(a + t) & sCh[t];
        while (eb) {
          const int q = __clz(eb);
          eb &= ~(0x80000000u >> q);
          cur ^= sDelta[(t * 32 + q) * P.wprE + w];
        }
      }
    }
    const uint32_t xt = warp_transpose32(cur & dl0);  // lane j: bit j of the 32 rows
    for (uint32_t t = kw; t < P.wprA; ++t) {
      uint32_t ab = valid ? __ldg(a + t) : 0u;
      if (t == kw) ab &= (0x7FFFFFFFu >> (k & 31));   // strictly after k
      if (!__any_sync(0xffffffffu, ab != 0)) continue;
      const uint32_t ul = warp_transpose32(ab);       // lane l: which of the 32 rows use atom t*32+l
      const uint32_t pu = __popc(ul);
      BIC_DCHECK(!ul || t * 32 + lane < P.p);                            // a user bit beyond the last atom would be a stray pad bit of A
      uint32_t* hl = corr + (t * 32 + lane) * hsC + w * 32;
      uint32_t dl = dl0;
      while (dl) {                                    // warp-uniform
        const int bp = __clz(dl);
        dl &= ~(0x80000000u >> bp);
        const uint32_t cj = __shfl_sync(0xffffffffu, xt, bp);
        if (ul) atomicAdd(hl + bp, pu - 2u * __popc(ul & cj));
      }
    }
  }
}

// exclusive prefix sum of cnt[0..p) into off[0..p], by one warp (p <= 1024)
__device__ __forceinline__ void warp_offsets(const uint32_t* cnt, uint32_t* off, uint32_t p) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t per = (p + 31) / 32;
  uint32_t s = 0;
  for (uint32_t i = lane * per; i < (lane + 1) * per && i < p; ++i) s += cnt[i];
  uint32_t inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  uint32_t run = inc - s;
  for (uint32_t i = lane * per; i < (lane + 1) * per && i < p; ++i) { off[i] = run; run += cnt[i]; }
  if (lane == 31) off[p] = inc;
}

// hit mask of one list row for coefficient word t: its atoms that are followed by a later atom of the row
__device__ __forceinline__ uint32_t hit_word(const uint32_t* __restrict__ a, uint32_t t, uint32_t wprA) {
  const uint32_t v = __ldg(a + t);
  bool later = false;
  for (uint32_t u = t + 1; u < wprA; ++u) later |= __ldg(a + u) != 0;
  return later ? v : (v & (v - 1));  // MSB first: the lowest-valued set bit is the row's last atom
}

// bucket sizes: hcount[k] = list rows using atom k and a later atom
__global__ void __launch_bounds__(256) k_dict_bucket_count(const uint32_t* __restrict__ listA, const uint32_t* __restrict__ count,
                                                           uint32_t* __restrict__ hcount, uint32_t wprA, uint32_t p,
                                                           const uint32_t* __restrict__ skip) {
  if (skip && *skip) return;
  extern __shared__ uint32_t sCnt[];       // wprA * 32
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < wprA * 32; i += blockDim.x) sCnt[i] = 0;
  __syncthreads();
  const uint32_t L = __ldcg(count);
  const uint32_t nblk = (L + 31) / 32;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t blk = gw; blk < nblk; blk += nw) {
    const uint32_t idx = blk * 32 + lane;
    for (uint32_t t = 0; t < wprA; ++t) {
      const uint32_t hm = (idx < L) ? hit_word(listA + (uint64_t)idx * wprA, t, wprA) : 0u;
      if (!__any_sync(0xffffffffu, hm != 0)) continue;
      const uint32_t c = __popc(warp_transpose32(hm));
      if (c) atomicAdd(&sCnt[t * 32 + lane], c);
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < p; i += blockDim.x)
    if (sCnt[i]) atomicAdd(hcount + i, sCnt[i]);
}

// bucket contents. A CTA takes chunks of the list; per chunk it counts its rows per atom in shared memory, reserves
// its range of every bucket with one global atomic per atom, and writes the row indices.
static const int FILL_CHUNK = 1024;
__global__ void __launch_bounds__(256) k_dict_bucket_fill(const uint32_t* __restrict__ listA, const uint32_t* __restrict__ count,
                                                          const uint32_t* __restrict__ hcount, uint32_t* __restrict__ cursor,
                                                          uint32_t* __restrict__ bucket, uint32_t bucket_cap, uint32_t wprA, uint32_t p,
                                                          const uint32_t* __restrict__ skip) {
  if (skip && *skip) return;
  extern __shared__ uint32_t s_mem[];
  uint32_t* sOff = s_mem;                  // p + 1
  uint32_t* sCnt = sOff + p + 1;           // wprA * 32
  uint32_t* sBase = sCnt + wprA * 32;      // wprA * 32
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < p; i += blockDim.x) sCnt[i] = __ldcg(hcount + i);
  __syncthreads();
  if (warp == 0) warp_offsets(sCnt, sOff, p);
  __syncthreads();
  if (sOff[p] > bucket_cap) return;        // the chain falls back to scanning the list
  const uint32_t L = __ldcg(count);
  const uint32_t nchunk = (L + FILL_CHUNK - 1) / FILL_CHUNK;
  for (uint32_t ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
    for (uint32_t i = threadIdx.x; i < wprA * 32; i += blockDim.x) sCnt[i] = 0;
    __syncthreads();
    for (int pass = 0; pass < 2; ++pass) {
      for (uint32_t blk = warp; blk < FILL_CHUNK / 32; blk += 8) {
        const uint32_t idx0 = ch * FILL_CHUNK + blk * 32;
        if (idx0 >= L) break;
        const uint32_t idx = idx0 + lane;
        for (uint32_t t = 0; t < wprA; ++t) {
          const uint32_t hm = (idx < L) ? hit_word(listA + (uint64_t)idx * wprA, t, wprA) : 0u;
          if (!__any_sync(0xffffffffu, hm != 0)) continue;
          uint32_t rows = warp_transpose32(hm);   // lane k: which of the 32 rows go to bucket t*32+k
          const uint32_t c = __popc(rows);
          if (!c) continue;
          const uint32_t k = t * 32 + lane;
          const uint32_t o0 = atomicAdd(&sCnt[k], c);
          if (pass == 1) {
            uint32_t o = sOff[k] + sBase[k] + o0;
            while (rows) {
              const int r = __clz(rows);
              rows &= ~(0x80000000u >> r);
              BIC_DCHECK(o < bucket_cap && o < sOff[k + 1]);               // stays inside bucket k's range
              bucket[o++] = idx0 + r;
            }
          }
        }
      }
      __syncthreads();
      if (pass == 0) {
        for (uint32_t k = threadIdx.x; k < p; k += blockDim.x) {
          const uint32_t c = sCnt[k];
          sBase[k] = c ? atomicAdd(cursor + k, c) : 0u;
          sCnt[k] = 0;
        }
        __syncthreads();
      }
    }
  }
}

__global__ void __launch_bounds__(CHAIN_THREADS, 1) k_dict_chain(ChainParams P) {
  if (P.skip && *P.skip) return;           // every CTA of the cluster takes the same branch
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank(), csize = cluster.num_blocks();
  extern __shared__ __align__(16) uint32_t s_mem[];
  const uint32_t nH = P.p * P.hs;
  const uint32_t hsC = P.hs + 1, nC = P.p * hsC;
  uint32_t* sH = s_mem;                    // p*hs
  uint32_t* sCorr = sH + nH;               // p*(hs+1): this CTA's correction counts of the current change
  uint32_t* sDelta = sCorr + nC;           // p*wprE
  uint32_t* sD = sDelta + P.p * P.wprE;    // p*wprE
  uint32_t* sU = sD + P.p * P.wprE;        // p
  uint32_t* sCh = sU + P.p;                // wprA
  uint32_t* sCnt = sCh + P.wprA;           // p: bucket sizes
  uint32_t* sOff = sCnt + P.p;             // p + 1: bucket offsets
  // cluster-wide sums of the correction counts: counter i lives in CTA i % csize at [i / csize]; three buffers in rotation
  const uint32_t slice = (nH + csize - 1) / csize;
  uint32_t* sAcc = sOff + P.p + 1;         // 3 * slice
  __shared__ uint32_t sFirst;
  __shared__ uint32_t sQ[CHAIN_THREADS / 32][64];  // scan fallback, per warp: list rows waiting for a full batch
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = CHAIN_THREADS / 32;
  for (uint32_t i = tid; i < nH; i += CHAIN_THREADS) sH[i] = __ldcg(P.H + i);
  for (uint32_t i = tid; i < P.p * P.wprE; i += CHAIN_THREADS) { sDelta[i] = 0; sD[i] = __ldcg(P.D + i); }
  for (uint32_t i = tid; i < P.p; i += CHAIN_THREADS) { sU[i] = __ldcg(P.U + i); sCnt[i] = __ldcg(P.hcount + i); }
  const bool multi = P.x.nranks > 1;
  uint32_t* const my_win = multi ? P.x.win[P.x.rank] : nullptr;
  const uint32_t epoch0 = multi ? __ldcg(my_win + XWIN_EPOCH2) : 0u;   // exchanges done so far (the same number on every rank)
  for (uint32_t i = tid; i < P.wprA; i += CHAIN_THREADS) sCh[i] = 0;
  for (uint32_t i = tid; i < 3 * slice; i += CHAIN_THREADS) sAcc[i] = 0;
  const uint32_t L = __ldcg(P.count);
  __syncthreads();
  if (warp == 0) warp_offsets(sCnt, sOff, P.p);
  cluster.sync();                          // every CTA of the cluster runs before any remote shared-memory access
  const bool bucketed = sOff[P.p] <= P.bucket_cap;
  uint32_t cursor = 0, nchanged = 0, nx = 0;  // nx: changes that needed an exchange (the sum buffers rotate on it)
  for (;;) {
    // first atom at or after the cursor that changes under the current histograms: it is the next one
    // to change in the reference's order (every earlier one is unchanged under the same H)
    if (tid == 0) sFirst = P.p;
    __syncthreads();
    for (uint32_t k = cursor + warp; k < P.p; k += nwarps) {
      if (k >= *(volatile uint32_t*)&sFirst) break;
      if (chain_decide(sH, sU, sD, k, P, nullptr)) {
        if (lane == 0) atomicMin(&sFirst, k);
        break;
      }
    }
    __syncthreads();
    const uint32_t k = sFirst;
    if (k >= P.p) break;
    if (warp == 0) chain_decide(sH, sU, sD, k, P, sDelta + k * P.wprE);
    if (__ldcg(P.gcount + k) == 0) {
      // no row (on any rank) uses atom k together with a later atom: the change reaches no other histogram (same decision in
      // every CTA and on every rank, so nobody waits at a barrier)
      __syncthreads();                       // sDelta[k] is complete before anyone moves on
      if (tid == 0) sCh[k >> 5] |= 0x80000000u >> (k & 31);
      nchanged++;
      cursor = k + 1;
      continue;
    }
    for (uint32_t i = (k + 1) * hsC + tid; i < nC; i += CHAIN_THREADS) sCorr[i] = 0;
    // Buffer rotation of the cluster-wide sums: exchange number c (nx) uses buffer c % 3. The buffer of change c + 1 is
    // cleared here: it was last read after the barrier of change c - 2, and every CTA finished those reads before
    // it arrived at the barrier of change c - 1, which we have passed; nobody adds to it before the barrier of
    // change c, which we have not reached.
    uint32_t* acc = sAcc + (nx % 3) * slice;
    {
      uint32_t* nxt = sAcc + ((nx + 1) % 3) * slice;
      for (uint32_t i = tid; i < slice; i += CHAIN_THREADS) nxt[i] = 0;
    }
    __syncthreads();
    const uint32_t kw = k >> 5, kbit = 0x80000000u >> (k & 31);
    if (bucketed) {
      // ---- this CTA's share of bucket k, 32 rows per warp and batch
      const uint32_t cnt = sCnt[k], off = sOff[k];
      const uint32_t nb = (cnt + 31) >> 5;
      for (uint32_t b = warp * csize + rank; b < nb; b += csize * nwarps) {
        const uint32_t e = b * 32 + lane;
        const bool valid = e < cnt;
        BIC_DCHECK(!valid || off + e < P.bucket_cap);
        chain_batch(P, k, valid ? __ldg(P.bucket + off + e) : 0u, valid, sDelta, sCh, sCorr, hsC);
      }
    } else {
      // ---- no room for the buckets: every warp scans its share of the list, queues the rows that use k and a
      // later atom and hands them to chain_batch 32 at a time
      uint32_t* q = sQ[warp];
      uint32_t qn = 0;                       // warp-uniform
      const uint32_t lt = (1u << lane) - 1;
      for (uint32_t g0 = rank * CHAIN_THREADS + warp * 32; g0 < L; g0 += csize * CHAIN_THREADS) {
        const uint32_t g = g0 + lane;
        bool hit = false;
        if (g < L) hit = (hit_word(P.listA + (uint64_t)g * P.wprA, kw, P.wprA) & kbit) != 0;
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        if (!bal) continue;
        if (hit) q[qn + __popc(bal & lt)] = g;
        qn += __popc(bal);
        __syncwarp();
        if (qn >= 32) {
          chain_batch(P, k, q[lane], true, sDelta, sCh, sCorr, hsC);
          const uint32_t rest = qn - 32;     // < 32
          const uint32_t mv = (lane < rest) ? q[32 + lane] : 0u;
          __syncwarp();
          if (lane < rest) q[lane] = mv;
          qn = rest;
          __syncwarp();
        }
      }
      if (qn) chain_batch(P, k, (lane < qn) ? q[lane] : 0u, lane < qn, sDelta, sCh, sCorr, hsC);
    }
    __syncthreads();
    // ---- cluster-wide sum: every CTA adds its nonzero counts to the counter's owner (one remote shared-memory
    // reduction per counter, fire and forget), barrier, then every CTA reads the sums it needs from the owners
    for (uint32_t i = (k + 1) * P.hs + tid; i < nH; i += CHAIN_THREADS) {
      const uint32_t v = sCorr[(i / P.hs) * hsC + (i % P.hs)];
      if (v) atomicAdd(cluster.map_shared_rank(acc, i % csize) + i / csize, v);
    }
    cluster.sync();
    if (!multi) {
      for (uint32_t i = (k + 1) * P.hs + tid; i < nH; i += CHAIN_THREADS) {
        const uint32_t j = i % P.hs;
        if ((sDelta[k * P.wprE + (j >> 5)] >> (31 - (j & 31))) & 1u) sH[i] += cluster.map_shared_rank(acc, i % csize)[i / csize];
      }
    } else {
      // Several GPUs: the sums above are this rank's rows only. CTA 0 writes them into every peer's window (a contiguous vector
      // per source rank, two areas alternating with the exchange number), publishes "rank r has pushed exchange e" and waits
      // for the same from every peer; the other CTAs wait at the hardware cluster barrier. Then everybody adds the peers' vectors.
      // A rank that runs ahead can be at most one exchange further (it needs this rank's flag to get past the next one), and
      // that one goes to the other area.
      const uint32_t e = epoch0 + nx + 1;
      const uint64_t area = P.x.xoff + (uint64_t)(e & 1u) * P.x.nranks * nH;
      BIC_DCHECK(P.x.rank < P.x.nranks && P.x.xoff >= XWIN_DATA);
      for (uint32_t i = (k + 1) * P.hs + tid; i < nH; i += CHAIN_THREADS) {
        const uint32_t j = i % P.hs;
        const uint32_t tot = cluster.map_shared_rank(acc, i % csize)[i / csize];
        if ((sDelta[k * P.wprE + (j >> 5)] >> (31 - (j & 31))) & 1u) sH[i] += tot;
        if (rank == 0)
          for (uint32_t r = 0; r < P.x.nranks; ++r)
            if (r != P.x.rank) P.x.win[r][area + (uint64_t)P.x.rank * nH + i] = tot;
      }
      if (rank == 0) {
        __threadfence_system();
        __syncthreads();
        if (tid == 0) {
          for (uint32_t r = 0; r < P.x.nranks; ++r)
            if (r != P.x.rank) *(volatile uint32_t*)(P.x.win[r] + XWIN_FLAGS2 + P.x.rank) = e;
          for (uint32_t r = 0; r < P.x.nranks; ++r) {
            if (r == P.x.rank) continue;
            uint32_t spins = 0;
            while ((int32_t)(*(volatile uint32_t*)(my_win + XWIN_FLAGS2 + r) - e) < 0) {
              __nanosleep(100);
              if (++spins > (1u << 27)) __trap();  // a peer that never arrives ends in an error, not a hang
            }
          }
          __threadfence_system();
        }
        __syncthreads();
      }
      cluster.sync();
      for (uint32_t i = (k + 1) * P.hs + tid; i < nH; i += CHAIN_THREADS) {
        const uint32_t j = i % P.hs;
        if (!((sDelta[k * P.wprE + (j >> 5)] >> (31 - (j & 31))) & 1u)) continue;
        uint32_t add = 0;
        for (uint32_t r = 0; r < P.x.nranks; ++r)
          if (r != P.x.rank) add += __ldcg(my_win + area + (uint64_t)r * nH + i);
        sH[i] += add;
      }
    }
    if (tid == 0) sCh[kw] |= kbit;
    nchanged++;
    nx++;
    cursor = k + 1;
    // the __syncthreads at the top of the loop orders these writes before the next decisions
  }
  cluster.sync();                          // no CTA leaves while another may still access its shared memory
  if (multi && rank == 0 && tid == 0) my_win[XWIN_EPOCH2] = epoch0 + nx;
  if (rank == 0) {
    for (uint32_t i = tid; i < P.p * P.wprE; i += CHAIN_THREADS) {
      const uint32_t d = sDelta[i];
      P.delta[i] = d;
      if (d) P.D[i] = sD[i] ^ d;           // :510
    }
    for (uint32_t i = tid; i < P.wprA; i += CHAIN_THREADS) P.chmask[i] = sCh[i];
    if (tid == 0) {
      P.chmask[P.wprA] = nchanged;
      if (nchanged) atomicAdd(P.changed, (unsigned long long)nchanged);  // :509
    }
  }
}

// E_i ^= XOR{delta_k : k changed, k in S_i} for every row (src/bsvd.cpp:512-520 for all changed atoms at once)
__global__ void __launch_bounds__(256) k_dict_apply(uint32_t* __restrict__ E, const uint32_t* __restrict__ A,
                                                    const uint32_t* __restrict__ delta, const uint32_t* __restrict__ chmask,
                                                    uint64_t n, uint32_t wprE, uint32_t wprA, uint32_t p,
                                                    const uint32_t* __restrict__ skip) {
  if (skip && *skip) return;
  if (__ldcg(chmask + wprA) == 0) return;  // no atom changed
  extern __shared__ uint32_t s_mem[];
  uint32_t* sDelta = s_mem;                // p*wprE
  uint32_t* sCh = sDelta + p * wprE;       // wprA
  for (uint32_t i = threadIdx.x; i < p * wprE; i += blockDim.x) sDelta[i] = __ldcg(delta + i);
  for (uint32_t i = threadIdx.x; i < wprA; i += blockDim.x) sCh[i] = __ldcg(chmask + i);
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t row = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; row < n; row += stride) {
    const uint32_t* a = A + row * wprA;
    if (wprA == 1 && wprE == 2) {          // 8x8 patches, <= 32 atoms: the bench shape
      uint32_t ab = __ldg(a) & sCh[0];
      if (!ab) continue;
      uint32_t x0 = 0, x1 = 0;
      while (ab) {
        const int q = __clz(ab);
        ab &= ~(0x80000000u >> q);
        x0 ^= sDelta[q * 2];
        x1 ^= sDelta[q * 2 + 1];
      }
      if (x0 | x1) {
        uint2* er = (uint2*)(E + row * 2);
        uint2 v = *er;
        v.x ^= x0; v.y ^= x1;
        *er = v;
      }
      continue;
    }
    bool any = false;
    for (uint32_t t = 0; t < wprA; ++t) any |= (__ldg(a + t) & sCh[t]) != 0;
    if (!any) continue;
    for (uint32_t w = 0; w < wprE; ++w) {
      uint32_t x = 0;
      for (uint32_t t = 0; t < wprA; ++t) {
        uint32_t ab = __ldg(a + t) & sCh[t];
        while (ab) {
          const int q = __clz(ab);
          ab &= ~(0x80000000u >> q);
          x ^= sDelta[(t * 32 + q) * wprE + w];
        }
      }
      if (x) E[row * wprE + w] ^= x;
    }
  }
}

// rows using >= 2 atoms -> list (order irrelevant: only integer sums are taken over it). Used when the
// histogram pass cannot do it on the side (more than one tile per row).
__global__ void __launch_bounds__(256) k_dict_compact(const uint32_t* __restrict__ E, const uint32_t* __restrict__ A,
                                                      uint32_t* __restrict__ listA, uint32_t* __restrict__ listE,
                                                      uint32_t* __restrict__ count, uint64_t n, uint32_t wprE, uint32_t wprA,
                                                      const uint32_t* __restrict__ skip) {
  if (skip && *skip) return;
  const int lane = threadIdx.x & 31;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t nround = div_up_u64(n, stride);
  for (uint64_t it = 0; it < nround; ++it) {
    const uint64_t row = it * stride + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t c = 0;
    if (row < n) for (uint32_t t = 0; t < wprA; ++t) c += __popc(__ldg(A + row * wprA + t));
    const bool multi = c >= 2;
    const uint32_t bal = __ballot_sync(0xffffffffu, multi);
    if (!bal) continue;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(count, (uint32_t)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (multi) {
      const uint64_t pos = base + __popc(bal & ((1u << lane) - 1));
      for (uint32_t t = 0; t < wprA; ++t) listA[pos * wprA + t] = __ldg(A + row * wprA + t);
      for (uint32_t w = 0; w < wprE; ++w) listE[pos * wprE + w] = __ldg(E + row * wprE + w);
    }
  }
}

bic_status bic_k_dict_hist_compact(bic_ctx* c, const bic_mat* E, const bic_mat* A, uint32_t* H, uint32_t* U, uint64_t hs,
                                   uint32_t* listA, uint32_t* listE, uint32_t* count, uint32_t* hcount, bool* fused);

// cluster size to launch with: the context's choice, but 16 (a non-portable size) only where the device can
// actually co-schedule such a cluster with the kernel's shared-memory footprint
static int bic_chain_cluster_size(bic_ctx* c) {
  if (c->chain_cluster != 16) return c->chain_cluster;
  static int ok16[64] = {0};  // 0 unknown, 1 yes, -1 no
  static std::mutex mu;       // contexts of several host threads come through here at once
  std::lock_guard<std::mutex> lk(mu);
  const int d = c->device < 64 ? c->device : 63;
  if (ok16[d] == 0) {
    cudaFuncSetAttribute(k_dict_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
    cudaFuncSetAttribute(k_dict_chain, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(16, 1, 1);
    cfg.blockDim = dim3(CHAIN_THREADS, 1, 1);
    cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 16; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nclusters = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, k_dict_chain, &cfg);
    if (e != cudaSuccess) cudaGetLastError();
    ok16[d] = (e == cudaSuccess && nclusters >= 1) ? 1 : -1;
  }
  return ok16[d] == 1 ? 16 : 8;
}

// Is this shape handled here? Every CTA holds H, its correction counts and its slice of the cluster-wide sums.
bool bic_dict_chain_eligible(bic_ctx* c, uint64_t n, uint64_t p, uint64_t wprE) {
  const uint64_t hs = wprE * 32, csize = (uint64_t)bic_chain_cluster_size(c);
  const uint64_t smem = (p * hs + p * (hs + 1) + 3 * div_up_u64(p * hs, csize) + 2 * p * wprE + 3 * p + div_up_u64(p, 32) + 8) * 4 + 8 * 1024 + 64;
  return n < (1ull << 30) &&  // list indices and bucket offsets are 32-bit (a row sits in at most p - 1 buckets; the capacity is 2n)
         smem + 1024 <= c->smem_optin && smem <= 200 * 1024;
}

__global__ void k_copy_u32(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint32_t n, const uint32_t* __restrict__ skip) {
  if (skip && *skip) return;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// hook == nullptr: one GPU. Otherwise the rows are one rank's shard: hook->reduce sums [H | U | bucket sizes | extra] over the
// ranks between the histogram pass and the chain, and the chain exchanges the corrections of every atom that changes through
// the peer windows in hook->x.
bic_status bic_k_update_dictionary_v3x(bic_ctx* c, bic_mat* E, bic_mat* D, const bic_mat* A, unsigned long long* d_changed, V3Hook* hook) {
  BIC_RANGE("bic:update_dictionary_steepest(cluster chain)");
  const uint64_t n = E->rows, p = D->rows, wprE = E->wpr, wprA = A->wpr, hs = wprE * 32;
  if (p == 0 || E->cols == 0) return BIC_OK;
  if (n == 0 && !hook) return BIC_OK;
  // global part (summed over the ranks when sharded): H (p*hs) | U (p) | gcount (wprA*32) | extra (4)
  // local part: count (4) | hcount (wprA*32) | cursor (wprA*32) | delta (p*wprE) | chmask (wprA + 1)
  // work[2]: listA (n*wprA + 4) | listE (n*wprE) | bucket (bucket_cap)
  const size_t glob_words = (size_t)(p * hs + p + wprA * 32 + 4);
  const size_t zero_local = (size_t)(4 + 2 * wprA * 32);
  const size_t local_words = zero_local + (size_t)(p * wprE + wprA + 1);
  uint32_t* G = nullptr;
  if (hook && hook->window_words) {
    // sharded with peer windows: the global part lives in this rank's window (so a collective of the driver's choice can work
    // in place), followed by the chain's exchange area
    const size_t xwords = (size_t)2 * hook->x.nranks * p * hs;
    uint32_t* base = nullptr;
    uint64_t base_off = 0;
    const size_t got = hook->window_words(hook->user, glob_words + 64 + xwords, &base, &base_off);
    if (got < glob_words + 64 + xwords || !base) return bic_fail(c, BIC_ERR_NOMEM, "update_dictionary: peer window too small");
    G = base;
    hook->x.xoff = base_off + ((glob_words + 63) & ~(size_t)63);   // (hook->x.win is a table in DEVICE memory: not for the host to read)
  }
  BIC_TRY(bic_scratch_reserve(c, &c->work[3], (glob_words + local_words) * 4 + 64));
  if (!G) G = (uint32_t*)c->work[3].p;
  uint32_t* Lc = (uint32_t*)c->work[3].p + glob_words;
  const size_t la_words = ((size_t)n * wprA + 7) & ~(size_t)3;  // listE stays 16-byte aligned
  // a row with s atoms sits in s - 1 buckets; 2 entries per row covers sparse codes, denser ones fall back to the scan
  uint64_t bucket_cap = c->chain_bucket_cap >= 0 ? (uint64_t)c->chain_bucket_cap : 2 * n;
  if (bucket_cap > 0xFFFFFFF0ull) bucket_cap = 0xFFFFFFF0ull;
  BIC_TRY(bic_scratch_reserve(c, &c->work[2], (la_words + (size_t)n * wprE + bucket_cap) * 4 + 64));
  uint32_t* H = G;
  uint32_t* U = H + p * hs;
  uint32_t* gcount = U + p;
  uint32_t* extra = gcount + wprA * 32;
  uint32_t* count = Lc;
  uint32_t* hcount = count + 4;
  uint32_t* cursor = hcount + wprA * 32;
  uint32_t* delta = cursor + wprA * 32;
  uint32_t* chmask = delta + p * wprE;
  uint32_t* listA = (uint32_t*)c->work[2].p;
  uint32_t* listE = listA + la_words;
  uint32_t* bucket = listE + (size_t)n * wprE;
  BIC_CUDA(c, cudaMemsetAsync(G, 0, glob_words * 4, c->stream));
  BIC_CUDA(c, cudaMemsetAsync(Lc, 0, zero_local * 4, c->stream));
  bool fused = false;
  if (n) {
    BIC_TRY(bic_k_dict_hist_compact(c, E, A, H, U, hs, listA, listE, count, hcount, &fused));
    if (!fused) {
      const int grid = bic_grid_for(c, n, 256, 8);
      BIC_PROF(c, KID_DICT_COMPACT);
      k_dict_compact<<<grid, 256, 0, c->stream>>>(E->d, A->d, listA, listE, count, n, (uint32_t)wprE, (uint32_t)wprA, c->loop_skip);
      BIC_LAUNCH_CHECK(c);
    }
    if (!fused) {  // the fused histogram pass counted the buckets as well
      const int grid = bic_grid_for(c, n, 256, 2);
      BIC_PROF(c, KID_DICT_BUCKET);
      k_dict_bucket_count<<<grid, 256, (size_t)wprA * 32 * 4, c->stream>>>(listA, count, hcount, (uint32_t)wprA, (uint32_t)p, c->loop_skip);
      BIC_LAUNCH_CHECK(c);
    }
  }
  const uint32_t* gc = hcount;
  if (hook) {
    k_copy_u32<<<1, 256, 0, c->stream>>>(gcount, hcount, (uint32_t)(wprA * 32), c->loop_skip);
    BIC_LAUNCH_CHECK(c);
    BIC_TRY(hook->reduce(hook->user, c, G, glob_words, extra));
    gc = gcount;
  }
  if (n) {
    const int grid2 = bic_grid_for(c, div_up_u64(n, FILL_CHUNK) * 256, 256, 4);
    BIC_PROF(c, KID_DICT_BUCKET);
    k_dict_bucket_fill<<<grid2, 256, (size_t)(p + 1 + 2 * wprA * 32) * 4, c->stream>>>(listA, count, hcount, cursor, bucket,
                                                                                     (uint32_t)bucket_cap, (uint32_t)wprA, (uint32_t)p, c->loop_skip);
    BIC_LAUNCH_CHECK(c);
  }
  ChainParams P;
  P.D = D->d; P.H = H; P.U = U; P.listA = listA; P.listE = listE; P.count = count; P.delta = delta; P.chmask = chmask;
  P.hcount = hcount; P.bucket = bucket; P.bucket_cap = (uint32_t)bucket_cap;
  P.changed = d_changed; P.p = (uint32_t)p; P.wprE = (uint32_t)wprE; P.wprA = (uint32_t)wprA; P.hs = (uint32_t)hs; P.m = E->cols;
  P.skip = c->loop_skip;
  P.gcount = gc;
  if (hook) P.x = hook->x;
  const unsigned csize = (unsigned)bic_chain_cluster_size(c);
  const size_t smem = (size_t)(p * hs + p * (hs + 1) + 2 * p * wprE + 3 * p + 1 + wprA + 3 * div_up_u64(p * hs, csize)) * 4;
  {
    static size_t optin_done[64] = {0};
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (c->device >= 64 || optin_done[c->device] < smem) {
      BIC_CUDA(c, cudaFuncSetAttribute(k_dict_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
      BIC_CUDA(c, cudaFuncSetAttribute(k_dict_chain, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));  // chain_cluster = 16
      if (c->device < 64) optin_done[c->device] = 200 * 1024;
    }
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(csize, 1, 1);
  cfg.blockDim = dim3(CHAIN_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BIC_PROF(c, KID_DICT_CHAIN);
  BIC_CUDA(c, cudaLaunchKernelEx(&cfg, k_dict_chain, P));
  BIC_LAUNCH_CHECK(c);
  if (n) {
    const int grid = bic_grid_for(c, n, 256, 8);
    BIC_PROF(c, KID_DICT_APPLY);
    k_dict_apply<<<grid, 256, (size_t)(p * wprE + wprA) * 4, c->stream>>>(E->d, A->d, delta, chmask, n, (uint32_t)wprE,
                                                                        (uint32_t)wprA, (uint32_t)p, c->loop_skip);
    BIC_LAUNCH_CHECK(c);
  }
  return BIC_OK;
}

bic_status bic_k_update_dictionary_v3(bic_ctx* c, bic_mat* E, bic_mat* D, const bic_mat* A, unsigned long long* d_changed) {
  return bic_k_update_dictionary_v3x(c, E, D, A, d_changed, nullptr);
}
