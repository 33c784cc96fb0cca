// learn_model_traditional over a BATCH of independent, equally shaped fits (e.g. the 16 bitplanes of a
// grey image, or the pages of a batch that each get their own dictionary): every kernel of the
// iteration is launched once for the whole batch (problem index in blockIdx.y/.z, pointers from a
// device-side problem table), so the GPU sees a few large launches per iteration instead of a chain of
// small ones per plane. Each problem follows exactly the reference's loop (src/bsvd.cpp:1215-1244):
// it stops taking part as soon as one of its iterations changes nothing.
#include "bic_internal.cuh"

#include <vector>

bic_status bic_k_update_coefficients_batched(bic_ctx* c, uint64_t n, uint64_t m, uint64_t p, const ProbDev* probs,
                                             const uint32_t* active, uint32_t nprob);
bic_status bic_k_transpose_A_batched(bic_ctx* c, uint64_t n, uint64_t p, const ProbDev* probs, const uint32_t* active,
                                     uint32_t nprob);
bic_status bic_k_dict_hist_batched(bic_ctx* c, uint64_t n, uint64_t m, uint64_t p, const ProbDev* probs, const uint32_t* active,
                                   uint32_t nprob);
bic_status bic_k_dict_step_batched(bic_ctx* c, uint64_t n, uint64_t m, uint64_t p, const ProbDev* probs, const uint32_t* active,
                                   uint32_t nprob, uint32_t launched);

// per iteration: Dnew <- D, cursors <- 0, first <- p, counts <- 0 for every active problem
__global__ void k_batch_begin_iteration(const ProbDev* __restrict__ probs, const uint32_t* __restrict__ active, uint64_t dwords,
                                        uint32_t p) {
  if (!active[blockIdx.y]) return;
  const ProbDev pr = probs[blockIdx.y];
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < dwords; i += (uint64_t)gridDim.x * blockDim.x)
    pr.Dnew[i] = pr.D[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    pr.cursor[0] = 0; pr.cursor[1] = 0; pr.first[0] = p; pr.first[1] = p;
    pr.counts[0] = 0; pr.counts[1] = 0;
  }
}

// D <- Dnew after the resolve (src/bsvd.cpp:510 for every atom that changed)
__global__ void k_batch_commit(const ProbDev* __restrict__ probs, const uint32_t* __restrict__ active, uint64_t dwords) {
  if (!active[blockIdx.y]) return;
  const ProbDev pr = probs[blockIdx.y];
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < dwords; i += (uint64_t)gridDim.x * blockDim.x)
    pr.D[i] = pr.Dnew[i];
}

extern "C" bic_status bic_learn_model_traditional_batched(bic_ctx* c, uint32_t nprob, const bic_mat* const* X, bic_mat* const* E,
                                                          bic_mat* const* D, bic_mat* const* A, uint64_t* iterations) {
  if (!c || !X || !E || !D || !A || nprob == 0) return BIC_ERR_INVALID;
  cudaSetDevice(c->device);
  const uint64_t n = X[0]->rows, m = X[0]->cols, p = D[0]->rows;
  for (uint32_t b = 0; b < nprob; ++b) {
    if (!X[b] || !E[b] || !D[b] || !A[b]) return BIC_ERR_INVALID;
    if (X[b]->rows != n || X[b]->cols != m || E[b]->rows != n || E[b]->cols != m || D[b]->rows != p || D[b]->cols != m ||
        A[b]->rows != n || A[b]->cols != p)
      return bic_fail(c, BIC_ERR_INVALID, "batched learner: all problems must have the same shapes");
  }
  for (uint32_t b = 0; b < nprob; ++b) BIC_TRY(bic_residual(c, X[b], A[b], D[b], E[b]));  // src/bsvd.cpp:1219-1220
  if (iterations) for (uint32_t b = 0; b < nprob; ++b) iterations[b] = 0;
  if (n == 0 || p == 0 || m == 0) {  // nothing can change: one empty iteration each, like the reference
    if (iterations) for (uint32_t b = 0; b < nprob; ++b) iterations[b] = 1;
    return BIC_OK;
  }
  const uint64_t wpr = div_up_u64(m, 32), hs = wpr * 32, wprN = div_up_u64(n, 32);
  // scratch per problem (u32 words): AT | H | U | Dnew | cursor[2] first[2] | counts (2 u64) -- 16-byte aligned pieces
  auto al = [](uint64_t w) { return (w + 3) & ~(uint64_t)3; };
  const uint64_t o_AT = 0, o_H = o_AT + al(p * wprN), o_U = o_H + al(p * hs), o_Dn = o_U + al(p), o_cur = o_Dn + al(p * wpr),
                 o_cnt = o_cur + 4, per = o_cnt + 4;
  const size_t table_bytes = ((size_t)nprob * sizeof(ProbDev) + 255) & ~(size_t)255;
  const size_t active_bytes = ((size_t)nprob * 4 + 255) & ~(size_t)255;
  BIC_TRY(bic_scratch_reserve(c, &c->work[2], table_bytes + active_bytes + (size_t)nprob * per * 4 + 256));
  uint8_t* base = (uint8_t*)c->work[2].p;
  ProbDev* d_probs = (ProbDev*)base;
  uint32_t* d_active = (uint32_t*)(base + table_bytes);
  uint32_t* d_pool = (uint32_t*)(base + table_bytes + active_bytes);
  std::vector<ProbDev> h_probs(nprob);
  for (uint32_t b = 0; b < nprob; ++b) {
    uint32_t* q = d_pool + (size_t)b * per;
    ProbDev& pr = h_probs[b];
    pr.E = E[b]->d; pr.D = D[b]->d; pr.A = A[b]->d;
    pr.AT = q + o_AT; pr.H = q + o_H; pr.U = q + o_U; pr.Dnew = q + o_Dn; pr.cursor = q + o_cur; pr.first = q + o_cur + 2;
    pr.counts = (unsigned long long*)(q + o_cnt);
  }
  BIC_CUDA(c, cudaMemcpyAsync(d_probs, h_probs.data(), nprob * sizeof(ProbDev), cudaMemcpyHostToDevice, c->stream));
  std::vector<uint32_t> active(nprob, 1);
  std::vector<uint32_t> cur(nprob * 4);
  std::vector<unsigned long long> cnt(nprob * 2);
  uint32_t nactive = nprob;
  // the only batched kernels with shape limits are the coefficient ones; fall back per problem otherwise
  while (nactive) {
    BIC_CUDA(c, cudaMemcpyAsync(d_active, active.data(), nprob * 4, cudaMemcpyHostToDevice, c->stream));
    k_batch_begin_iteration<<<dim3(8, nprob), 256, 0, c->stream>>>(d_probs, d_active, p * wpr, (uint32_t)p);
    BIC_LAUNCH_CHECK(c);
    // H, U <- 0 (contiguous per problem)
    for (uint32_t b = 0; b < nprob; ++b)
      if (active[b]) BIC_CUDA(c, cudaMemsetAsync(h_probs[b].H, 0, (size_t)(o_Dn - o_H) * 4, c->stream));
    bic_status st = bic_k_update_coefficients_batched(c, n, m, p, d_probs, d_active, nprob);  // src/bsvd.cpp:1229
    if (st == BIC_ERR_UNSUPPORTED) return bic_fail(c, st, "batched learner: rows wider than 1024 bits or dictionary too large for shared memory");
    BIC_TRY(st);
    BIC_TRY(bic_k_transpose_A_batched(c, n, p, d_probs, d_active, nprob));                    // src/bsvd.cpp:1235 ...
    BIC_TRY(bic_k_dict_hist_batched(c, n, m, p, d_probs, d_active, nprob));
    uint32_t launched = 0, batch = 8;
    for (;;) {
      for (uint32_t i = 0; i < batch && launched < p; ++i, ++launched)
        BIC_TRY(bic_k_dict_step_batched(c, n, m, p, d_probs, d_active, nprob, launched));
      // cursors of all problems: one strided copy per problem would be nprob copies; they are 16 B apart
      // inside each problem's scratch, so gather them with a 2-D copy
      BIC_CUDA(c, cudaMemcpy2DAsync(cur.data(), 16, d_pool + o_cur, per * 4, 16, nprob, cudaMemcpyDeviceToHost, c->stream));
      BIC_CUDA(c, bic_wait_stream(c));
      bool done = true;
      for (uint32_t b = 0; b < nprob && done; ++b)
        if (active[b] && cur[b * 4 + (launched & 1)] < p) done = false;
      if (done || launched >= p) break;
      batch = (batch * 2 < 64) ? batch * 2 : 64;
    }
    k_batch_commit<<<dim3(8, nprob), 256, 0, c->stream>>>(d_probs, d_active, p * wpr);
    BIC_LAUNCH_CHECK(c);
    BIC_CUDA(c, cudaMemcpy2DAsync(cnt.data(), 16, d_pool + o_cnt, per * 4, 16, nprob, cudaMemcpyDeviceToHost, c->stream));
    BIC_CUDA(c, bic_wait_stream(c));
    for (uint32_t b = 0; b < nprob; ++b) {
      if (!active[b]) continue;
      if (iterations) iterations[b]++;
      if (cnt[b * 2] + cnt[b * 2 + 1] == 0) { active[b] = 0; nactive--; }  // while (changed > 0), src/bsvd.cpp:1227
      else if (cnt[b * 2 + 1] == 0) {
        // no atom changed: the next iteration of this problem provably changes nothing (see
        // bic_learn_model_traditional); it is counted without being run
        if (iterations) iterations[b]++;
        active[b] = 0;
        nactive--;
      }
    }
  }
  return BIC_OK;
}
