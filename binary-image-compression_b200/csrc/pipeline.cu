// The encoder as a pipeline that ONE host thread drives: a pool of slots (a CUDA stream + a workspace each), a queue of rasters,
// and per slot a small state machine that only ever enqueues work and polls an event -- it never waits for the device.
//
// What bic_encode_raster does per raster (src/bsvd_test.cpp:56-125 with the three PBM dumps replaced by Golomb streams) has three
// points where the host needs a number from the device: the pivot draw of initialize_model_neighbor (src/bsvd.cpp:239-243), the
// loop test of learn_model_traditional (:1227) and the stream sizes for the container. Here
//   * the draw runs on the device (init.cu: k_draw_pivots, a warp replaying rand48 with jump-ahead),
//   * the learner's iterations are queued in small batches; a device flag (k_loop_end) turns the iterations queued past the end of
//     the loop into no-ops (every kernel of an iteration returns at once when it is set), so the result is exactly the reference's
//     loop -- same iteration count, same trace -- and the host only looks at the flag when a batch's event has fired,
//   * the Golomb encoder writes into a pre-sized buffer with the bit count staying on the device (coding.cu:
//     bic_k_golomb_encode_async); the sizes reach the host with one small copy, then the container's pieces are copied out.
// So a raster costs the host three event polls and ~100 enqueues, and any number of rasters are in flight on one thread.
// Shapes the cluster-chain dictionary update does not take (large dictionaries) run the synchronous path inside the slot.
#include "bic_internal.cuh"

#include <sched.h>

#include <deque>
#include <memory>
#include <vector>

bic_status bic_k_update_coefficients(bic_ctx* c, bic_mat* E, const bic_mat* D, bic_mat* A, unsigned long long* d_changed);
bic_status bic_k_update_dictionary_v3(bic_ctx* c, bic_mat* E, bic_mat* D, const bic_mat* A, unsigned long long* d_changed);
bool bic_dict_chain_eligible(bic_ctx* c, uint64_t n, uint64_t p, uint64_t wprE);
bic_status bic_k_init_neighbor_async(bic_ctx* c, const bic_mat* X, bic_mat* D, bic_mat* A, uint64_t* d_state, unsigned long long* d_status);
bic_status bic_k_golomb_encode_async(bic_ctx* c, const bic_mat* M, uint32_t chunk_samples, bic_stream* out, unsigned long long* d_info);
bic_status bic_golomb_async_finish(bic_stream* out, const uint64_t* host_info);
bic_status bic_k_golomb_encode_multi(bic_ctx* c, const bic_mat* const* mats, int nmat, uint32_t chunk_samples, bic_stream* const* outs,
                                     unsigned long long* d_info);

struct bic_comm;
bic_status bic_k_dist_prepare(bic_ctx* c, bic_comm* m, uint64_t n_local, uint64_t p, uint64_t wprE, uint64_t wprA);
bool bic_k_dist_fused(const bic_comm* m);
bic_status bic_k_dist_init_async(bic_ctx* c, bic_comm* m, const bic_mat* X, bic_mat* D, bic_mat* A, uint64_t* d_state, unsigned long long* d_status);
bic_status bic_k_dist_iteration_dict(bic_ctx* c, bic_comm* m, bic_mat* E, bic_mat* D, const bic_mat* A, unsigned long long* d_counts);
bic_status bic_k_dist_golomb_async(bic_ctx* c, bic_comm* m, const bic_mat* const* mats, int nmat, uint32_t chunk_samples,
                                   bic_stream* const* outs, unsigned long long* d_info, unsigned long long* d_shard);
int bic_comm_rank(const bic_comm* m);
int bic_comm_size(const bic_comm* m);

#define LOOP_TRACE 64
struct LoopState {
  uint32_t done, iter;            // done: the loop has ended; iterations queued after that do nothing
  unsigned long long iters;       // learn_model_traditional's return value (src/bsvd.cpp:1243)
  unsigned long long trace[2 * LOOP_TRACE];
};

__global__ void k_loop_reset(LoopState* st, unsigned long long* counts) {
  st->done = 0; st->iter = 0; st->iters = 0;
  counts[0] = counts[1] = 0;
}

// end of one iteration of src/bsvd.cpp:1227-1242: changed = changed_coefs + changed_atoms decides whether the loop goes on.
// When no atom changed, D and E are exactly what the coefficient update left and every row ended that update with a pass that
// found no improving atom, so the reference's next iteration changes nothing and ends the loop: it is counted without being run
// (the argument of bic_learn_model_traditional, encoder.cu).
__global__ void k_loop_end(LoopState* st, unsigned long long* counts) {
  if (st->done) return;
  const uint32_t it = ++st->iter;
  const unsigned long long cc = counts[0], ca = counts[1];
  counts[0] = counts[1] = 0;
  if (it <= LOOP_TRACE) { st->trace[2 * (it - 1)] = cc; st->trace[2 * (it - 1) + 1] = ca; }
  if (cc + ca == 0) { st->done = 1; st->iters = it; }
  else if (ca == 0) {
    st->done = 1; st->iters = it + 1;
    if (it + 1 <= LOOP_TRACE) { st->trace[2 * it] = 0; st->trace[2 * it + 1] = 0; }
  }
}

namespace {

enum Stage { ST_IDLE = 0, ST_LEARN, ST_CODE, ST_COPY };

struct Job {
  uint64_t id = 0;
  const uint8_t* payload = nullptr;   // host P4 payload, or
  const bic_mat* raster = nullptr;    // a raster already on the device
  uint64_t rows = 0, cols = 0, W = 0, K = 0;
  unsigned long seed = 0;
  uint8_t* out = nullptr;
  uint64_t cap = 0;
  bic_encode_info* info = nullptr;
  cudaEvent_t after = nullptr;        // recorded at submit time on the producer's stream (resident rasters)
  bic_status status = BIC_OK;
  bool done = false;
  std::string err;
};

// device words of a slot's status block
enum { SW_RNG = 0, SW_ALLZERO = 1, SW_DRAWS = 2, SW_INFO = 8, SW_SHARD = 32, SW_WORDS = 48 };

struct HostMirror {
  uint64_t status[SW_WORDS];
  LoopState loop;
  uint64_t rng_in;
};

struct Slot {
  bic_ctx* c = nullptr;
  uint64_t rows = 0, cols = 0, W = 0, K = 0;
  bic_mat *raster = nullptr, *X = nullptr, *E = nullptr, *D = nullptr, *A = nullptr;
  bic_stream* st[3] = {nullptr, nullptr, nullptr};
  uint64_t* d_status = nullptr;
  LoopState* d_loop = nullptr;
  HostMirror* h = nullptr;            // pinned
  cudaEvent_t ev = nullptr;
  Job* job = nullptr;
  Stage stage = ST_IDLE;
  bool async_ok = false;
  bic_comm* comm = nullptr;           // sharded mode: this slot's communicator (the same slot index on every rank)
  std::deque<Job*> queue;             // sharded mode: jobs are dealt to slots by their sequence number, identically on every rank
};

}  // namespace

struct bic_pipeline {
  int device = 0;
  std::vector<Slot> slots;
  std::deque<Job*> pending;
  std::deque<std::unique_ptr<Job>> jobs;   // every job ever submitted and not yet forgotten
  uint64_t first_id = 1, next_id = 1;
  uint64_t in_flight = 0;
  int first_batch = 2, next_batch = 2;
  uint64_t polls = 0, batches = 0, sync_fallbacks = 0, recodes = 0;
  bool sharded = false;               // every job is one rank's row shard of a matrix that all ranks fit together
  uint64_t seq = 0;                   // jobs submitted so far (sharded mode: job q runs on slot q % nslots)
  std::string err;
};

static void slot_release(Slot& s) {
  bic_mat** ms[] = {&s.raster, &s.X, &s.E, &s.D, &s.A};
  for (auto pm : ms) if (*pm) { bic_mat_destroy(s.c, *pm); *pm = nullptr; }
}

static bic_status slot_prepare(Slot& s, uint64_t rows, uint64_t cols, uint64_t W, uint64_t K, bool need_raster) {
  bic_ctx* c = s.c;
  if (s.X && s.rows == rows && s.cols == cols && s.W == W && s.K == K) {
    if (need_raster && !s.raster) BIC_TRY(bic_mat_create(c, rows, cols, &s.raster));
    return BIC_OK;
  }
  slot_release(s);
  const uint64_t Ny = (W - 1 + rows) / W, Nx = (W - 1 + cols) / W, n = Nx * Ny, m = W * W;
  if (need_raster) BIC_TRY(bic_mat_create(c, rows, cols, &s.raster));
  BIC_TRY(bic_mat_create(c, n, m, &s.X));
  BIC_TRY(bic_mat_create(c, n, m, &s.E));
  BIC_TRY(bic_mat_create(c, K, m, &s.D));
  BIC_TRY(bic_mat_create(c, n, K, &s.A));
  for (auto& st : s.st) if (!st) BIC_TRY(bic_stream_create(c, &st));
  s.rows = rows; s.cols = cols; s.W = W; s.K = K;
  s.async_ok = c->dict_algo == 2 && c->dict_update == 0 && n > 0 && n <= 0xFFFFFFFFull && K > 0 && K <= 65535 &&
               bic_dict_chain_eligible(c, n, K, s.E->wpr);
  if (s.comm) {
    // first raster of this shape on this slot: shard sizes and the peer window are exchanged now (blocking, collective -- every
    // rank gets here for the same slot in the same order, see bic_pipeline_attach_comms)
    if (!s.async_ok) return bic_fail(c, BIC_ERR_UNSUPPORTED, "sharded pipeline: the shape must be one the cluster-chain update takes");
    BIC_TRY(bic_k_dist_prepare(c, s.comm, n, K, s.E->wpr, s.A->wpr));
    if (!bic_k_dist_fused(s.comm)) return bic_fail(c, BIC_ERR_UNSUPPORTED, "sharded pipeline: the ranks cannot map each other's memory");
  }
  return BIC_OK;
}

static void job_finish(bic_pipeline* P, Slot& s, bic_status st, const char* msg = nullptr) {
  Job* j = s.job;
  j->status = st;
  if (st != BIC_OK) j->err = msg ? msg : s.c->err;
  j->done = true;
  if (j->after) { cudaEventDestroy(j->after); j->after = nullptr; }
  s.job = nullptr;
  s.stage = ST_IDLE;
  P->in_flight--;
}

// queue `niter` iterations of src/bsvd.cpp:1227-1242 (and, first, E = A*D xor X, :1219-1220)
static bic_status learn_enqueue(Slot& s, bool first, int niter) {
  bic_ctx* c = s.c;
  unsigned long long* d_cc = (unsigned long long*)c->d_scalars;
  if (first) {
    BIC_TRY(bic_residual(c, s.X, s.A, s.D, s.E));
    k_loop_reset<<<1, 1, 0, c->stream>>>(s.d_loop, d_cc);
    BIC_LAUNCH_CHECK(c);
  }
  c->loop_skip = &s.d_loop->done;
  bic_status st = BIC_OK;
  for (int i = 0; i < niter && st == BIC_OK; ++i) {
    st = bic_k_update_coefficients(c, s.E, s.D, s.A, d_cc);                    // :1229
    if (st == BIC_OK) {                                                        // :1235
      if (s.comm) st = bic_k_dist_iteration_dict(c, s.comm, s.E, s.D, s.A, d_cc);   // statistics summed over the ranks, d_cc[0] made global
      else st = bic_k_update_dictionary_v3(c, s.E, s.D, s.A, d_cc + 1);
    }
    if (st == BIC_OK) {
      k_loop_end<<<1, 1, 0, c->stream>>>(s.d_loop, d_cc);
      c->launches++;
      if (cudaGetLastError() != cudaSuccess) st = BIC_ERR_CUDA;
    }
  }
  c->loop_skip = nullptr;
  return st;
}

static bic_status mirror_and_mark(Slot& s, bool with_trace) {
  bic_ctx* c = s.c;
  BIC_CUDA(c, cudaMemcpyAsync(s.h->status, s.d_status, sizeof(uint64_t) * SW_WORDS, cudaMemcpyDeviceToHost, c->stream));
  BIC_CUDA(c, cudaMemcpyAsync(&s.h->loop, s.d_loop, with_trace ? sizeof(LoopState) : 16, cudaMemcpyDeviceToHost, c->stream));
  BIC_CUDA(c, cudaEventRecord(s.ev, c->stream));
  return BIC_OK;
}

static bic_status encode_sync(Slot& s, Job* j, const bic_mat* rast);

static void job_start(bic_pipeline* P, Slot& s, Job* j) {
  BIC_RANGE("bic:pipeline:job_start");
  s.job = j;
  P->in_flight++;
  bic_ctx* c = s.c;
  bic_status st = slot_prepare(s, j->rows, j->cols, j->W, j->K, j->payload != nullptr);
  if (st != BIC_OK) return job_finish(P, s, st);
  const bic_mat* rast = j->raster;
  if (j->payload) {
    st = bic_mat_upload_pbm(c, s.raster, j->payload);
    if (st != BIC_OK) return job_finish(P, s, st);
    rast = s.raster;
  } else if (j->after) {
    if (cudaStreamWaitEvent(c->stream, j->after, 0) != cudaSuccess) return job_finish(P, s, BIC_ERR_CUDA, "cudaStreamWaitEvent");
  }
  st = bic_extract_patches(c, rast, j->W, s.X);                                  // src/bsvd_test.cpp:80-99
  if (st != BIC_OK) return job_finish(P, s, st);
  if (!s.async_ok) {
    if (s.comm) return job_finish(P, s, BIC_ERR_UNSUPPORTED, "sharded pipeline: shape outside the asynchronous path");
    P->sync_fallbacks++;
    st = encode_sync(s, j, rast);
    return job_finish(P, s, st);
  }
  bic_rand48_seed(&s.h->rng_in, j->seed);                                        // -r / random_seed, src/bsvd.cpp:12
  if (cudaMemcpyAsync(s.d_status + SW_RNG, &s.h->rng_in, 8, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
    return job_finish(P, s, BIC_ERR_CUDA, "cudaMemcpyAsync");
  if (s.comm) st = bic_k_dist_init_async(c, s.comm, s.X, s.D, s.A, s.d_status + SW_RNG, (unsigned long long*)(s.d_status + SW_ALLZERO));
  else st = bic_k_init_neighbor_async(c, s.X, s.D, s.A, s.d_status + SW_RNG, (unsigned long long*)(s.d_status + SW_ALLZERO));  // :114
  if (st == BIC_OK) st = learn_enqueue(s, true, P->first_batch);                 // :119
  if (st == BIC_OK) st = mirror_and_mark(s, false);
  if (st != BIC_OK) return job_finish(P, s, st);
  P->batches++;
  s.stage = ST_LEARN;
}

// shapes outside the asynchronous path: the synchronous calls, blocking this thread (large dictionaries; rare in a pipeline)
static bic_status encode_sync(Slot& s, Job* j, const bic_mat* rast) {
  bic_ctx* c = s.c;
  uint64_t rng;
  bic_rand48_seed(&rng, j->seed);
  BIC_TRY(bic_initialize_model_neighbor(c, s.X, s.D, s.A, &rng));
  uint64_t iters = 0;
  BIC_TRY(bic_learn_model_traditional(c, s.X, s.E, s.D, s.A, &iters, nullptr, 0));
  const bic_mat* mats[3] = {s.D, s.A, s.E};
  for (int i = 0; i < 3; ++i) BIC_TRY(bic_golomb_encode(c, mats[i], 256, s.st[i]));
  bic_stream_info si[3];
  for (int i = 0; i < 3; ++i) si[i] = s.st[i]->info;
  const uint64_t need = bic_container_bytes(si);
  if (j->info) {
    bic_encode_info* info = j->info;
    memset(info, 0, sizeof(*info));
    info->rows = rast->rows; info->cols = rast->cols; info->W = j->W; info->K = j->K; info->n = s.X->rows; info->m = s.X->cols;
    info->iterations = iters;
    info->bits_D = si[0].bitcount; info->bits_A = si[1].bitcount; info->bits_E = si[2].bitcount;
    info->weight_D = si[0].nsamples - 1; info->weight_A = si[1].nsamples - 1; info->weight_E = si[2].nsamples - 1;
    info->container_bytes = need;
  }
  if (!j->out) return BIC_OK;
  if (j->cap < need) return BIC_ERR_CAPACITY;
  if (((uintptr_t)j->out & 7) != 0) return bic_fail(c, BIC_ERR_INVALID, "encode: out must be 8-byte aligned");
  bic_container_header(j->out, rast->rows, rast->cols, j->W, j->K, s.X->rows, s.X->cols, iters, (uint64_t)j->seed, si);
  uint64_t off = (BIC_HDR_FIELDS + 3 * BIC_STREAM_FIELDS) * 8;
  for (int i = 0; i < 3; ++i) {
    const uint64_t nb = div_up_u64(si[i].bitcount, 8), nbp = div_up_u64(nb, 8) * 8;
    memset(j->out + off + nb, 0, nbp - nb);
    BIC_TRY(bic_stream_download(c, s.st[i], j->out + off, nb, (uint64_t*)(j->out + off + nbp), si[i].nchunks));
    off += nbp + si[i].nchunks * 16;
  }
  return BIC_OK;
}

// A sharded job's output ("shard container"), little-endian u64 fields:
//   [0] magic "BICSHRD\0" [1] version [2] rank [3] ranks [4] rows (this rank's band) [5] cols [6] W [7] K [8] n (local patches) [9] m
//   [10] iterations [11] seed, [12 + 7 i ...) per stream (D, A, E) coder, chunk_samples, rows, cols, bits of the local buffer,
//   local samples, local chunk-index entries; [33 + 6 j ...) for A and E the bic_shard_info fields (global bit count, global samples,
//   code bit offset, local code bits, first chunk, local chunks); then per stream the bytes padded to 8 and the chunk index.
//   The global stream of A (of E) is the word-wise OR of the ranks' buffers placed at 32-bit word (code bit offset >> 5).
static const uint64_t BIC_SHARD_MAGIC = 0x0044524853434942ull;  // "BICSHRD\0"
static const uint64_t SHARD_HDR_U64 = 48;

static void slot_finish_sharded(bic_pipeline* P, Slot& s) {
  bic_ctx* c = s.c;
  Job* j = s.job;
  bic_status st = bic_golomb_async_finish(s.st[0], s.h->status + SW_INFO);
  if (st != BIC_OK) return job_finish(P, s, st, "sharded pipeline: the dictionary's code did not fit its buffer");
  for (int i = 1; i < 3; ++i) {
    const uint64_t* info = s.h->status + SW_INFO + 8 * i;
    if (info[4]) return job_finish(P, s, BIC_ERR_CAPACITY, "sharded pipeline: a shard's code did not fit its pre-sized buffer (raise gol_presize_pct)");
    s.st[i]->info.bitcount = info[0];
    s.st[i]->info.nsamples = info[1];
    s.st[i]->info.nchunks = info[3];
  }
  const uint64_t* shA = s.h->status + SW_SHARD;
  const uint64_t* shE = s.h->status + SW_SHARD + 6;
  bic_stream_info si[3];
  for (int i = 0; i < 3; ++i) si[i] = s.st[i]->info;
  uint64_t need = SHARD_HDR_U64 * 8;
  for (int i = 0; i < 3; ++i) need += div_up_u64(div_up_u64(si[i].bitcount, 8), 8) * 8 + si[i].nchunks * 16;
  const bic_mat* rast = j->payload ? s.raster : j->raster;
  if (j->info) {
    bic_encode_info* info = j->info;
    memset(info, 0, sizeof(*info));
    info->rows = rast->rows; info->cols = rast->cols; info->W = j->W; info->K = j->K; info->n = s.X->rows; info->m = s.X->cols;
    info->iterations = s.h->loop.iters;
    info->bits_D = si[0].bitcount; info->bits_A = shA[0]; info->bits_E = shE[0];          // A, E: of the ONE global stream
    info->weight_D = si[0].nsamples - 1; info->weight_A = shA[1] - 1; info->weight_E = shE[1] - 1;
    info->container_bytes = need;
  }
  if (!j->out) return job_finish(P, s, BIC_OK);
  if (j->cap < need) return job_finish(P, s, BIC_ERR_CAPACITY, "encode: container larger than the caller's buffer");
  if (((uintptr_t)j->out & 7) != 0) return job_finish(P, s, BIC_ERR_INVALID, "encode: out must be 8-byte aligned");
  uint64_t* h = (uint64_t*)j->out;
  memset(h, 0, SHARD_HDR_U64 * 8);
  h[0] = BIC_SHARD_MAGIC; h[1] = 1; h[2] = (uint64_t)bic_comm_rank(s.comm); h[3] = (uint64_t)bic_comm_size(s.comm);
  h[4] = rast->rows; h[5] = rast->cols; h[6] = j->W; h[7] = j->K; h[8] = s.X->rows; h[9] = s.X->cols; h[10] = s.h->loop.iters; h[11] = (uint64_t)j->seed;
  for (int i = 0; i < 3; ++i) {
    uint64_t* f = h + 12 + 7 * i;
    f[0] = si[i].coder; f[1] = si[i].chunk_samples; f[2] = si[i].rows; f[3] = si[i].cols; f[4] = si[i].bitcount; f[5] = si[i].nsamples; f[6] = si[i].nchunks;
  }
  memcpy(h + 33, shA, 6 * 8);
  memcpy(h + 39, shE, 6 * 8);
  uint64_t off = SHARD_HDR_U64 * 8;
  for (int i = 0; i < 3; ++i) {
    const uint64_t nb = div_up_u64(si[i].bitcount, 8), nbp = div_up_u64(nb, 8) * 8;
    cudaError_t e = cudaSuccess;
    if (nbp > nb) memset(j->out + off + nb, 0, nbp - nb);   // (the copy below never touches the pad bytes)
    if (nb) e = cudaMemcpyAsync(j->out + off, s.st[i]->d_bytes, nb, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && si[i].nchunks)
      e = cudaMemcpyAsync(j->out + off + nbp, s.st[i]->d_index, si[i].nchunks * 16, cudaMemcpyDeviceToHost, c->stream);
    if (e != cudaSuccess) return job_finish(P, s, BIC_ERR_CUDA, cudaGetErrorString(e));
    off += nbp + si[i].nchunks * 16;
  }
  if (cudaEventRecord(s.ev, c->stream) != cudaSuccess) return job_finish(P, s, BIC_ERR_CUDA, "cudaEventRecord");
  s.stage = ST_COPY;
}

static void slot_advance(bic_pipeline* P, Slot& s) {
  BIC_RANGE("bic:pipeline:advance");
  bic_ctx* c = s.c;
  Job* j = s.job;
  if (s.stage == ST_LEARN) {
    if (s.h->status[SW_ALLZERO]) return job_finish(P, s, BIC_ERR_INVALID, "init: X is all zero (the reference's draw loop never ends)");
    bic_status st = BIC_OK;
    if (!s.h->loop.done) {                       // the loop goes on: another batch
      st = learn_enqueue(s, false, P->next_batch);
      if (st == BIC_OK) st = mirror_and_mark(s, false);
      if (st != BIC_OK) return job_finish(P, s, st);
      P->batches++;
      return;
    }
    const bic_mat* mats[3] = {s.D, s.A, s.E};
    if (s.comm) {
      // D is replicated: every rank codes the same stream; A and E are row shards of ONE global stream each
      st = bic_k_golomb_encode_multi(c, mats, 1, 256, s.st, (unsigned long long*)(s.d_status + SW_INFO));
      if (st == BIC_OK)
        st = bic_k_dist_golomb_async(c, s.comm, mats + 1, 2, 256, s.st + 1, (unsigned long long*)(s.d_status + SW_INFO + 8),
                                     (unsigned long long*)(s.d_status + SW_SHARD));
    } else if (c->gol_algo == 2) {               // D, A and E in one set of launches (coding2.cu)
      st = bic_k_golomb_encode_multi(c, mats, 3, 256, s.st, (unsigned long long*)(s.d_status + SW_INFO));
    } else {
      for (int i = 0; i < 3 && st == BIC_OK; ++i)
        st = bic_k_golomb_encode_async(c, mats[i], 256, s.st[i], (unsigned long long*)(s.d_status + SW_INFO + 8 * i));
    }
    if (st == BIC_OK) st = mirror_and_mark(s, true);
    if (st != BIC_OK) return job_finish(P, s, st);
    s.stage = ST_CODE;
    return;
  }
  if (s.stage == ST_CODE && s.comm) return slot_finish_sharded(P, s);
  if (s.stage == ST_CODE) {
    const bic_mat* mats[3] = {s.D, s.A, s.E};
    bic_stream_info si[3];
    for (int i = 0; i < 3; ++i) {
      bic_status st = bic_golomb_async_finish(s.st[i], s.h->status + SW_INFO + 8 * i);
      if (st == BIC_ERR_CAPACITY) {              // the code outgrew the pre-sized buffer: once more, exactly sized (blocking; rare)
        P->recodes++;
        st = bic_golomb_encode(c, mats[i], 256, s.st[i]);
      }
      if (st != BIC_OK) return job_finish(P, s, st);
      si[i] = s.st[i]->info;
    }
    const uint64_t need = bic_container_bytes(si);
    const bic_mat* rast = j->payload ? s.raster : j->raster;
    if (j->info) {
      bic_encode_info* info = j->info;
      memset(info, 0, sizeof(*info));
      info->rows = rast->rows; info->cols = rast->cols; info->W = j->W; info->K = j->K; info->n = s.X->rows; info->m = s.X->cols;
      info->iterations = s.h->loop.iters;
      info->bits_D = si[0].bitcount; info->bits_A = si[1].bitcount; info->bits_E = si[2].bitcount;
      info->weight_D = si[0].nsamples - 1; info->weight_A = si[1].nsamples - 1; info->weight_E = si[2].nsamples - 1;
      info->container_bytes = need;
    }
    if (!j->out) return job_finish(P, s, BIC_OK);
    if (j->cap < need) return job_finish(P, s, BIC_ERR_CAPACITY, "encode: container larger than the caller's buffer");
    if (((uintptr_t)j->out & 7) != 0) return job_finish(P, s, BIC_ERR_INVALID, "encode: out must be 8-byte aligned");
    uint64_t off = (BIC_HDR_FIELDS + 3 * BIC_STREAM_FIELDS) * 8;
    for (int i = 0; i < 3; ++i) {
      const uint64_t nb = div_up_u64(si[i].bitcount, 8), nbp = div_up_u64(nb, 8) * 8;
      cudaError_t e = cudaSuccess;
      if (nb) e = cudaMemcpyAsync(j->out + off, s.st[i]->d_bytes, nb, cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess && si[i].nchunks)
        e = cudaMemcpyAsync(j->out + off + nbp, s.st[i]->d_index, si[i].nchunks * 16, cudaMemcpyDeviceToHost, c->stream);
      if (e != cudaSuccess) return job_finish(P, s, BIC_ERR_CUDA, cudaGetErrorString(e));
      off += nbp + si[i].nchunks * 16;
    }
    if (cudaEventRecord(s.ev, c->stream) != cudaSuccess) return job_finish(P, s, BIC_ERR_CUDA, "cudaEventRecord");
    s.stage = ST_COPY;
    return;
  }
  if (s.stage == ST_COPY && s.comm) return job_finish(P, s, BIC_OK);   // the header was written before the copies were queued
  if (s.stage == ST_COPY) {
    bic_stream_info si[3];
    for (int i = 0; i < 3; ++i) si[i] = s.st[i]->info;
    const bic_mat* rast = j->payload ? s.raster : j->raster;
    bic_container_header(j->out, rast->rows, rast->cols, j->W, j->K, s.X->rows, s.X->cols, s.h->loop.iters, (uint64_t)j->seed, si);
    uint64_t off = (BIC_HDR_FIELDS + 3 * BIC_STREAM_FIELDS) * 8;
    for (int i = 0; i < 3; ++i) {                // the pad bytes between a code and its index (host memory, after the copy landed)
      const uint64_t nb = div_up_u64(si[i].bitcount, 8), nbp = div_up_u64(nb, 8) * 8;
      memset(j->out + off + nb, 0, nbp - nb);
      off += nbp + si[i].nchunks * 16;
    }
    return job_finish(P, s, BIC_OK);
  }
}

extern "C" bic_status bic_pipeline_create(int device, int nslots, bic_pipeline** out) {
  if (!out || nslots < 1 || nslots > 256) return BIC_ERR_INVALID;
  *out = nullptr;
  std::unique_ptr<bic_pipeline> P(new (std::nothrow) bic_pipeline());
  if (!P) return BIC_ERR_NOMEM;
  P->device = device;
  P->slots.resize(nslots);
  for (auto& s : P->slots) {
    bic_status st = bic_ctx_create(device, &s.c);
    if (st != BIC_OK) { bic_pipeline_destroy(P.release()); return st; }
    bool ok = cudaMalloc((void**)&s.d_status, sizeof(uint64_t) * SW_WORDS) == cudaSuccess &&
              cudaMalloc((void**)&s.d_loop, sizeof(LoopState)) == cudaSuccess &&
              cudaMallocHost((void**)&s.h, sizeof(HostMirror)) == cudaSuccess &&
              cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming) == cudaSuccess;
    if (ok) ok = cudaMemset(s.d_status, 0, sizeof(uint64_t) * SW_WORDS) == cudaSuccess && cudaMemset(s.d_loop, 0, sizeof(LoopState)) == cudaSuccess;
    if (!ok) { cudaGetLastError(); bic_pipeline_destroy(P.release()); return BIC_ERR_NOMEM; }
    memset(s.h, 0, sizeof(HostMirror));
  }
  *out = P.release();
  return BIC_OK;
}

extern "C" bic_status bic_pipeline_destroy(bic_pipeline* P) {
  if (!P) return BIC_ERR_INVALID;
  cudaSetDevice(P->device);
  for (auto& s : P->slots) {
    if (!s.c) continue;
    cudaStreamSynchronize(s.c->stream);
    slot_release(s);
    for (auto& st : s.st) if (st) { bic_stream_destroy(s.c, st); st = nullptr; }
    if (s.d_status) cudaFree(s.d_status);
    if (s.d_loop) cudaFree(s.d_loop);
    if (s.h) cudaFreeHost(s.h);
    if (s.ev) cudaEventDestroy(s.ev);
    bic_ctx_destroy(s.c);
  }
  for (auto& j : P->jobs) if (j && j->after) cudaEventDestroy(j->after);
  delete P;
  return BIC_OK;
}

extern "C" bic_status bic_pipeline_set_option(bic_pipeline* P, const char* name, int64_t value) {
  if (!P || !name) return BIC_ERR_INVALID;
  if (!strcmp(name, "first_batch")) { if (value < 1 || value > 64) return BIC_ERR_INVALID; P->first_batch = (int)value; return BIC_OK; }
  if (!strcmp(name, "next_batch")) { if (value < 1 || value > 64) return BIC_ERR_INVALID; P->next_batch = (int)value; return BIC_OK; }
  bic_status st = BIC_OK;
  for (auto& s : P->slots) { st = bic_ctx_set_option(s.c, name, value); if (st != BIC_OK) break; }   // e.g. chain_cluster, coef_algo
  return st;
}

static bic_status submit(bic_pipeline* P, Job* j, uint64_t* id_out) {
  j->id = P->next_id++;
  if (id_out) *id_out = j->id;
  P->jobs.emplace_back(j);
  if (P->sharded) P->slots[P->seq % P->slots.size()].queue.push_back(j);   // the same slot (= communicator) on every rank
  else P->pending.push_back(j);
  P->seq++;
  return BIC_OK;
}

extern "C" bic_status bic_pipeline_submit(bic_pipeline* P, const uint8_t* pbm_payload, uint64_t rows, uint64_t cols, uint64_t W, uint64_t K,
                                          unsigned long seed, uint8_t* out, uint64_t cap_bytes, bic_encode_info* info, uint64_t* job) {
  if (!P || !pbm_payload || W == 0 || K == 0 || rows == 0 || cols == 0) return BIC_ERR_INVALID;
  Job* j = new (std::nothrow) Job();
  if (!j) return BIC_ERR_NOMEM;
  j->payload = pbm_payload; j->rows = rows; j->cols = cols; j->W = W; j->K = K; j->seed = seed; j->out = out; j->cap = cap_bytes; j->info = info;
  return submit(P, j, job);
}

extern "C" bic_status bic_pipeline_submit_resident(bic_pipeline* P, const bic_mat* raster, bic_ctx* producer, uint64_t W, uint64_t K,
                                                   unsigned long seed, uint8_t* out, uint64_t cap_bytes, bic_encode_info* info, uint64_t* job) {
  if (!P || !raster || W == 0 || K == 0 || raster->rows == 0 || raster->cols == 0) return BIC_ERR_INVALID;
  cudaSetDevice(P->device);
  Job* j = new (std::nothrow) Job();
  if (!j) return BIC_ERR_NOMEM;
  j->raster = raster; j->rows = raster->rows; j->cols = raster->cols; j->W = W; j->K = K; j->seed = seed; j->out = out; j->cap = cap_bytes; j->info = info;
  if (producer) {   // the raster is being produced on that context's stream: whatever it has queued so far comes first
    if (cudaEventCreateWithFlags(&j->after, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(j->after, producer->stream) != cudaSuccess) {
      cudaGetLastError();
      if (j->after) cudaEventDestroy(j->after);
      delete j;
      return BIC_ERR_CUDA;
    }
  }
  return submit(P, j, job);
}

extern "C" bic_status bic_pipeline_poll(bic_pipeline* P, uint64_t* in_flight) {
  if (!P) return BIC_ERR_INVALID;
  cudaSetDevice(P->device);
  P->polls++;
  for (auto& s : P->slots) {
    if (s.stage != ST_IDLE) {
      const cudaError_t e = cudaEventQuery(s.ev);
      if (e == cudaErrorNotReady) continue;
      if (e != cudaSuccess) { job_finish(P, s, BIC_ERR_CUDA, cudaGetErrorString(e)); continue; }
      slot_advance(P, s);
    }
    std::deque<Job*>& q = P->sharded ? s.queue : P->pending;
    while (s.stage == ST_IDLE && !q.empty()) {   // a job that finishes inside job_start (error, synchronous shape) frees the slot again
      Job* j = q.front();
      q.pop_front();
      job_start(P, s, j);
    }
  }
  uint64_t waiting = P->pending.size();
  for (auto& s : P->slots) waiting += s.queue.size();
  if (in_flight) *in_flight = P->in_flight + waiting;
  return BIC_OK;
}

extern "C" bic_status bic_pipeline_wait(bic_pipeline* P, uint64_t job) {
  if (!P) return BIC_ERR_INVALID;
  for (;;) {
    uint64_t left = 0;
    BIC_TRY(bic_pipeline_poll(P, &left));
    if (job == 0) { if (left == 0) return BIC_OK; }
    else {
      if (job < P->first_id || job >= P->next_id) return BIC_ERR_INVALID;
      if (P->jobs[job - P->first_id]->done) return BIC_OK;
    }
    sched_yield();
  }
}

extern "C" bic_status bic_pipeline_job_status(bic_pipeline* P, uint64_t job, int* done, const char** error) {
  if (!P || job < P->first_id || job >= P->next_id) return BIC_ERR_INVALID;
  const Job* j = P->jobs[job - P->first_id].get();
  if (done) *done = j->done ? 1 : 0;
  if (error) *error = j->err.c_str();
  return j->done ? j->status : BIC_OK;
}

// drop the records of finished jobs at the front of the queue (ids stay valid for the others)
extern "C" bic_status bic_pipeline_forget_finished(bic_pipeline* P) {
  if (!P) return BIC_ERR_INVALID;
  while (!P->jobs.empty() && P->jobs.front()->done) { P->jobs.pop_front(); P->first_id++; }
  return BIC_OK;
}

extern "C" bic_status bic_pipeline_stats(bic_pipeline* P, uint64_t* launches, uint64_t* polls, uint64_t* batches, uint64_t* sync_fallbacks,
                                         uint64_t* recodes) {
  if (!P) return BIC_ERR_INVALID;
  uint64_t l = 0;
  for (auto& s : P->slots) l += s.c->launches;
  if (launches) *launches = l;
  if (polls) *polls = P->polls;
  if (batches) *batches = P->batches;
  if (sync_fallbacks) *sync_fallbacks = P->sync_fallbacks;
  if (recodes) *recodes = P->recodes;
  return BIC_OK;
}

// stream ordering against a context outside the pool (timers, producers): every slot waits for what `signal` has queued so far /
// `waiter` waits for what every slot has queued so far
extern "C" bic_status bic_pipeline_wait_ctx(bic_pipeline* P, bic_ctx* signal) {
  if (!P || !signal) return BIC_ERR_INVALID;
  for (auto& s : P->slots) BIC_TRY(bic_ctx_wait_ctx(s.c, signal));
  return BIC_OK;
}
extern "C" bic_status bic_ctx_wait_pipeline(bic_ctx* waiter, bic_pipeline* P) {
  if (!P || !waiter) return BIC_ERR_INVALID;
  for (auto& s : P->slots) BIC_TRY(bic_ctx_wait_ctx(waiter, s.c));
  return BIC_OK;
}


// ---- sharded mode: every job is this rank's row shard (a band of the raster) of ONE matrix that all ranks fit together with one
// dictionary (csrc/dist.cu). Slot i of every rank shares communicator i: create them with bic_comm_create on
// bic_pipeline_slot_ctx(p, i) (ids exchanged by the caller's plumbing), attach them, and from then on submit the SAME jobs in the
// SAME order on every rank -- job q runs on slot q % nslots everywhere. A job's output is a shard container (see above).
extern "C" bic_ctx* bic_pipeline_slot_ctx(bic_pipeline* P, int slot) {
  if (!P || slot < 0 || slot >= (int)P->slots.size()) return nullptr;
  return P->slots[slot].c;
}
extern "C" bic_status bic_pipeline_attach_comms(bic_pipeline* P, bic_comm* const* comms, int n) {
  if (!P || !comms || n != (int)P->slots.size()) return BIC_ERR_INVALID;
  if (P->in_flight || !P->pending.empty()) return BIC_ERR_INVALID;
  for (int i = 0; i < n; ++i) {
    if (!comms[i]) return BIC_ERR_INVALID;
    P->slots[i].comm = comms[i];
    P->slots[i].c->defer_free = true;      // no cudaFree (a device-wide wait) while peers may be waiting for this rank
  }
  P->sharded = true;
  P->seq = 0;
  return BIC_OK;
}

// ---- the shard containers of ONE sharded job (one per rank) -> the ordinary container of the whole raster.
// Pure host code: byte surgery on buffers the caller gathered from the ranks (MPI, NCCL, files -- not our business). The result
// is the container bic_encode_raster produces for the concatenated bands, byte for byte: D's stream is replicated (rank 0's is
// taken, the others are checked against it), the global A / E stream is the OR of the ranks' buffers at their 32-bit word
// offsets (a shard's buffer starts with code_bit_offset mod 32 zero bits), the chunk indexes are already in global terms and
// concatenate in rank order. Every field is bounded before it sizes anything: the inputs may come from anywhere.
namespace {
struct ShardView {
  const uint64_t* h = nullptr;
  uint64_t bytes = 0;
  uint64_t off[3] = {0, 0, 0}, nb[3] = {0, 0, 0}, nbp[3] = {0, 0, 0}, nidx[3] = {0, 0, 0};
};
bool shard_view(const uint8_t* p, uint64_t bytes, ShardView* v) {
  if (!p || bytes < SHARD_HDR_U64 * 8 || ((uintptr_t)p & 7) != 0) return false;
  const uint64_t* h = (const uint64_t*)p;
  if (h[0] != BIC_SHARD_MAGIC || h[1] != 1) return false;
  v->h = h;
  v->bytes = bytes;
  uint64_t off = SHARD_HDR_U64 * 8;
  for (int i = 0; i < 3; ++i) {
    const uint64_t* f = h + 12 + 7 * i;
    const uint64_t room = bytes - off;
    if (f[4] > room * 8 || f[6] > room / 16) return false;                 // bits of the local buffer, chunk-index entries
    const uint64_t nb = div_up_u64(f[4], 8), nbp = div_up_u64(nb, 8) * 8;
    if (nbp > room || f[6] * 16 > room - nbp) return false;
    v->off[i] = off; v->nb[i] = nb; v->nbp[i] = nbp; v->nidx[i] = f[6];
    off += nbp + f[6] * 16;
  }
  return true;
}
}  // namespace

extern "C" bic_status bic_merge_shard_containers(const uint8_t* const* shards, const uint64_t* shard_bytes, int nshards, uint8_t* out,
                                                 uint64_t cap_bytes, uint64_t* bytes) {
  if (!shards || !shard_bytes || nshards < 1 || nshards > 64) return BIC_ERR_INVALID;
  try {
    std::vector<ShardView> v((size_t)nshards);          // by rank
    for (int i = 0; i < nshards; ++i) {
      ShardView t;
      if (!shard_view(shards[i], shard_bytes[i], &t)) return BIC_ERR_CORRUPT;
      const uint64_t r = t.h[2];
      if (t.h[3] != (uint64_t)nshards || r >= (uint64_t)nshards || v[r].h) return BIC_ERR_CORRUPT;   // every rank exactly once
      v[r] = t;
    }
    const uint64_t* h0 = v[0].h;
    const uint64_t cols = h0[5], W = h0[6], K = h0[7], m = h0[9], iters = h0[10], seed = h0[11];
    if (W == 0 || W > 1024 || K == 0 || K > 65535 || cols == 0 || cols > (1ull << 40) || m != W * W) return BIC_ERR_CORRUPT;
    uint64_t rows = 0, n = 0;
    for (int r = 0; r < nshards; ++r) {
      const uint64_t* h = v[r].h;
      if (h[5] != cols || h[6] != W || h[7] != K || h[9] != m || h[10] != iters || h[11] != seed) return BIC_ERR_CORRUPT;
      if (h[4] == 0 || h[4] > (1ull << 40) || h[8] > (1ull << 40)) return BIC_ERR_CORRUPT;
      if (r + 1 < nshards && h[4] % W != 0) return BIC_ERR_INVALID;        // only the last band may end inside a patch row
      if (h[8] != div_up_u64(h[4], W) * div_up_u64(cols, W)) return BIC_ERR_CORRUPT;
      rows += h[4];
      n += h[8];
    }
    // ---- the three streams of the whole: D from rank 0, A and E from the shard fields
    bic_stream_info si[3];
    memset(si, 0, sizeof(si));
    const uint64_t* fD = h0 + 12;
    si[0].coder = (uint32_t)fD[0]; si[0].chunk_samples = (uint32_t)fD[1]; si[0].rows = fD[2]; si[0].cols = fD[3];
    si[0].bitcount = fD[4]; si[0].nsamples = fD[5]; si[0].nchunks = fD[6];
    if (fD[2] != K || fD[3] != m) return BIC_ERR_CORRUPT;
    for (int r = 1; r < nshards; ++r) {                                     // the replicas of D must agree
      if (memcmp(v[r].h + 12, fD, 7 * 8) != 0 || v[r].nb[0] != v[0].nb[0]) return BIC_ERR_CORRUPT;
      if (memcmp((const uint8_t*)v[r].h + v[r].off[0], (const uint8_t*)h0 + v[0].off[0], v[0].nbp[0] + v[0].nidx[0] * 16) != 0) return BIC_ERR_CORRUPT;
    }
    for (int s = 1; s < 3; ++s) {
      const uint64_t mcols = s == 1 ? K : m;
      const uint64_t* f0 = h0 + 12 + 7 * s;
      const uint64_t* g0 = h0 + 33 + 6 * (s - 1);       // global bit count, global samples, code bit offset, local code bits, first chunk, local chunks
      const uint64_t gbits = g0[0], gsamples = g0[1], chunk = f0[1];
      if (chunk == 0 || (chunk & (chunk - 1)) || gbits > (1ull << 45) || gsamples == 0 || gsamples > n * mcols + 1) return BIC_ERR_CORRUPT;
      uint64_t code = 0, chunks = 0, mrows = 0;
      for (int r = 0; r < nshards; ++r) {
        const uint64_t* f = v[r].h + 12 + 7 * s;
        const uint64_t* g = v[r].h + 33 + 6 * (s - 1);
        if (f[0] != f0[0] || f[1] != chunk || f[3] != mcols || f[2] != v[r].h[8]) return BIC_ERR_CORRUPT;
        if (g[0] != gbits || g[1] != gsamples) return BIC_ERR_CORRUPT;
        if (g[2] != code || g[4] != chunks || g[5] != f[6]) return BIC_ERR_CORRUPT;       // shards follow each other without a gap
        if (g[3] > gbits - code || f[4] != (g[2] & 31) + g[3]) return BIC_ERR_CORRUPT;
        code += g[3];
        chunks += g[5];
        mrows += f[2];
      }
      if (code != gbits || chunks != div_up_u64(gsamples, chunk) || mrows != n) return BIC_ERR_CORRUPT;
      si[s].coder = (uint32_t)f0[0]; si[s].chunk_samples = (uint32_t)chunk; si[s].rows = n; si[s].cols = mcols;
      si[s].bitcount = gbits; si[s].nsamples = gsamples; si[s].nchunks = chunks;
    }
    const uint64_t need = bic_container_bytes(si);
    if (bytes) *bytes = need;
    if (!out) return BIC_OK;
    if (cap_bytes < need) return BIC_ERR_CAPACITY;
    if (((uintptr_t)out & 7) != 0) return BIC_ERR_INVALID;
    memset(out, 0, need);
    bic_container_header(out, rows, cols, W, K, n, m, iters, seed, si);
    uint64_t off = (BIC_HDR_FIELDS + 3 * BIC_STREAM_FIELDS) * 8;
    for (int s = 0; s < 3; ++s) {
      const uint64_t nb = div_up_u64(si[s].bitcount, 8), nbp = div_up_u64(nb, 8) * 8;
      uint8_t* dst = out + off;
      uint64_t* idx = (uint64_t*)(out + off + nbp);
      if (s == 0) {
        memcpy(dst, (const uint8_t*)h0 + v[0].off[0], v[0].nb[0]);
        memcpy(idx, (const uint8_t*)h0 + v[0].off[0] + v[0].nbp[0], v[0].nidx[0] * 16);
      } else {
        uint64_t ci = 0;
        for (int r = 0; r < nshards; ++r) {
          const uint64_t* g = v[r].h + 33 + 6 * (s - 1);
          const uint8_t* src = (const uint8_t*)v[r].h + v[r].off[s];
          const uint64_t b0 = (g[2] >> 5) * 4;                              // the shard's buffer starts at this byte of the whole
          for (uint64_t i = 0; i < v[r].nb[s] && b0 + i < nb; ++i) dst[b0 + i] |= src[i];
          memcpy(idx + 2 * ci, src + v[r].nbp[s], v[r].nidx[s] * 16);
          ci += v[r].nidx[s];
        }
      }
      off += nbp + si[s].nchunks * 16;
    }
    return BIC_OK;
  } catch (const std::bad_alloc&) {
    return BIC_ERR_NOMEM;
  } catch (...) {
    return BIC_ERR_CORRUPT;
  }
}
