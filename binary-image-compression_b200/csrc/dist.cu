// Row-sharded fit over several GPUs: every rank holds a contiguous block of patch rows
// (its X, E, A), the dictionary D is replicated, and the integer statistics are combined with NCCL
// over NVLink. Integer sums are order independent, so the result is bit-identical to the single-GPU
// (and to the reference's serial) fit of the concatenated rows.
//
//   init   : allgather of the per-rank zero-row bitmaps (the rand48 accept/reject needs "is global row i
//            zero"), every rank replays the same draw; pivot rows are contributed by their owners through
//            one allreduce; one allreduce of [column histogram | per-pivot intersect counts].
//   coef   : local (rows are independent given D).
//   dict   : one allreduce of [H | U | changed-rows] per iteration, then the in-order resolve runs
//            replicated on every rank; only an atom that CHANGES costs another (small) allreduce of the
//            histogram corrections its users produced (dict2.cu).
//
// NCCL is loaded at run time (the copy torch already loaded if there is one, else the system
// libnccl.so.2); nccl.h is only used for its types.
#include "bic_internal.cuh"

#include <dlfcn.h>
#include <stdlib.h>
#include <nccl.h>

#include <mutex>
#include <vector>

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
  bool ok = false;
};

static NcclApi* nccl_load(NcclApi& api);
static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] { nccl_load(api); });
  return api.ok ? &api : nullptr;
}
static NcclApi* nccl_load(NcclApi& api) {
  // 1. a libnccl.so.2 some other component of the process (e.g. torch) already mapped; 2. the one
  // BIC_NCCL_LIB names; 3. the system one. Two different NCCL builds in one process do not mix.
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) { const char* e = getenv("BIC_NCCL_LIB"); if (e && *e) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL); }
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return nullptr;
#define LOAD(name) *(void**)(&api.name) = dlsym(h, "nccl" #name); if (!api.name) return nullptr
  LOAD(GetUniqueId); LOAD(CommInitRank); LOAD(CommDestroy); LOAD(AllReduce); LOAD(AllGather); LOAD(GetErrorString);
#undef LOAD
  api.ok = true;
  return &api;
}

struct bic_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  std::vector<uint64_t> nrows;    // rows per rank (set by the first collective that needs them)
  uint64_t row0 = 0, nglobal = 0;
  uint64_t* d_starts = nullptr;          // device copy of the shard boundaries (nranks + 1 global row indices), for the device-side draw
  uint64_t collectives = 0;
  // peer windows (cudaIpc) for the dictionary update's per-atom exchange: see XPeers in bic_internal.cuh
  int fused = -1;                        // -1 not tried yet, 0 unavailable (NCCL path), 1 windows mapped
  uint32_t* win = nullptr;               // this rank's window
  size_t win_words = 0;
  std::vector<uint32_t*> peer_win;       // every rank's window as mapped here (own = win)
  uint32_t** d_peer_win = nullptr;       // the same table in device memory
  uint32_t epoch = 0;                    // barriers used so far
  const uint32_t* last_extra = nullptr;  // where the last dictionary update left the summed changed-rows count (three 22-bit limbs)
};

#define BIC_NCCL(ctx, expr)                                                                   \
  do {                                                                                        \
    ncclResult_t _r = (expr);                                                                 \
    if (_r != ncclSuccess) {                                                                  \
      (ctx)->err = std::string(#expr) + ": " + nccl_api()->GetErrorString(_r);                \
      return BIC_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

extern "C" bic_status bic_comm_unique_id(uint8_t id[128]) {
  NcclApi* n = nccl_api();
  if (!n || !id) return BIC_ERR_UNSUPPORTED;
  ncclUniqueId u;
  if (n->GetUniqueId(&u) != ncclSuccess) return BIC_ERR_CUDA;
  memcpy(id, u.internal, 128);
  return BIC_OK;
}

extern "C" bic_status bic_comm_create(bic_ctx* c, int rank, int nranks, const uint8_t id[128], bic_comm** out) {
  if (!c || !out || !id || nranks < 1 || rank < 0 || rank >= nranks) return BIC_ERR_INVALID;
  cudaSetDevice(c->device);
  NcclApi* n = nccl_api();
  if (!n) return bic_fail(c, BIC_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
  bic_comm* m = new bic_comm();
  m->rank = rank;
  m->nranks = nranks;
  ncclUniqueId u;
  memcpy(u.internal, id, 128);
  ncclResult_t r = n->CommInitRank(&m->comm, nranks, u, rank);
  if (r != ncclSuccess) { c->err = std::string("ncclCommInitRank: ") + n->GetErrorString(r); delete m; return BIC_ERR_CUDA; }
  *out = m;
  return BIC_OK;
}

static void window_release(bic_comm* m) {
  for (int r = 0; r < (int)m->peer_win.size(); ++r)
    if (r != m->rank && m->peer_win[r]) cudaIpcCloseMemHandle(m->peer_win[r]);
  m->peer_win.clear();
  if (m->win) cudaFree(m->win);
  if (m->d_peer_win) cudaFree(m->d_peer_win);
  m->win = nullptr;
  m->d_peer_win = nullptr;
  m->win_words = 0;
}

extern "C" bic_status bic_comm_destroy(bic_ctx* c, bic_comm* m) {
  if (!c || !m) return BIC_ERR_INVALID;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  window_release(m);
  if (m->d_starts) cudaFree(m->d_starts);
  if (m->comm) nccl_api()->CommDestroy(m->comm);
  delete m;
  return BIC_OK;
}

extern "C" uint64_t bic_comm_collective_count(const bic_comm* m) { return m ? m->collectives : 0; }

static bic_status allreduce_u32(bic_ctx* c, bic_comm* m, uint32_t* buf, size_t count) {
  if (m->nranks == 1 || count == 0) return BIC_OK;
  BIC_NCCL(c, nccl_api()->AllReduce(buf, buf, count, ncclUint32, ncclSum, m->comm, c->stream));
  m->collectives++;
  return BIC_OK;
}

// small host-visible allgather of `count` u64 per rank (count <= 4): all[r*count + i]
bic_status bic_comm_allgather_u64(bic_ctx* c, bic_comm* m, const uint64_t* mine, int count, uint64_t* all) {
  if (count < 1 || count > 4 || m->nranks > 8) return BIC_ERR_INVALID;
  uint64_t* d = c->d_scalars + 16;  // [16, 16 + 8*4)
  BIC_CUDA(c, cudaMemcpyAsync(d + (size_t)m->rank * count, mine, 8 * count, cudaMemcpyHostToDevice, c->stream));
  if (m->nranks > 1) {
    BIC_NCCL(c, nccl_api()->AllGather(d + (size_t)m->rank * count, d, count, ncclUint64, m->comm, c->stream));
    m->collectives++;
  }
  BIC_CUDA(c, cudaMemcpyAsync(all, d, 8 * (size_t)count * m->nranks, cudaMemcpyDeviceToHost, c->stream));
  BIC_CUDA(c, bic_wait_stream(c));
  return BIC_OK;
}
int bic_comm_rank(const bic_comm* m) { return m->rank; }
int bic_comm_size(const bic_comm* m) { return m->nranks; }

// rows per rank -> offsets (one tiny allgather, repeated when the local row count changes)
static bic_status share_rows(bic_ctx* c, bic_comm* m, uint64_t n_local) {
  if ((int)m->nrows.size() == m->nranks && m->nrows[m->rank] == n_local) return BIC_OK;
  m->nrows.assign(m->nranks, 0);
  uint64_t* d = c->d_scalars + 40;  // [40, 40+nranks)
  if (m->nranks > 16) return bic_fail(c, BIC_ERR_UNSUPPORTED, "more than 16 ranks");
  BIC_CUDA(c, cudaMemcpyAsync(d + m->rank, &n_local, 8, cudaMemcpyHostToDevice, c->stream));
  if (m->nranks > 1) {
    BIC_NCCL(c, nccl_api()->AllGather(d + m->rank, d, 1, ncclUint64, m->comm, c->stream));
    m->collectives++;
  }
  BIC_CUDA(c, cudaMemcpyAsync(m->nrows.data(), d, 8 * m->nranks, cudaMemcpyDeviceToHost, c->stream));
  BIC_CUDA(c, bic_wait_stream(c));
  m->row0 = 0;
  m->nglobal = 0;
  for (int r = 0; r < m->nranks; ++r) { if (r < m->rank) m->row0 += m->nrows[r]; m->nglobal += m->nrows[r]; }
  std::vector<uint64_t> starts(m->nranks + 1, 0);
  for (int r = 0; r < m->nranks; ++r) starts[r + 1] = starts[r] + m->nrows[r];
  if (!m->d_starts) BIC_CUDA(c, cudaMalloc((void**)&m->d_starts, 8 * 65));
  BIC_CUDA(c, cudaMemcpyAsync(m->d_starts, starts.data(), 8 * (m->nranks + 1), cudaMemcpyHostToDevice, c->stream));
  BIC_CUDA(c, bic_wait_stream(c));
  return BIC_OK;
}

// the sharded calls of the pipeline (pipeline.cu) never wait for the host, so the shard boundaries must be known beforehand
bic_status bic_comm_share_rows(bic_ctx* c, bic_comm* m, uint64_t n_local) { return share_rows(c, m, n_local); }
bool bic_comm_rows_known(const bic_comm* m, uint64_t n_local) {
  return (int)m->nrows.size() == m->nranks && m->nrows[m->rank] == n_local && m->d_starts != nullptr;
}

bic_status bic_k_init_scratch(bic_ctx* c, uint64_t p, uint64_t wpr, InitWork* w);
bic_status bic_k_init_gather(bic_ctx* c, const bic_mat* X, const uint64_t* host_pivots, InitWork* w);
bic_status bic_k_init_stats(bic_ctx* c, const bic_mat* X, InitWork* w);
bic_status bic_k_init_finalize(bic_ctx* c, InitWork* w, uint64_t m, bic_mat* D);
bic_status bic_k_update_coefficients(bic_ctx* c, bic_mat* E, const bic_mat* D, bic_mat* A, unsigned long long* d_changed);
bic_status bic_k_dict_prepare(bic_ctx* c, bic_mat* E, bic_mat* D, const bic_mat* A, DictWork* w, uint32_t* hbase);
bic_status bic_k_dict_step(bic_ctx* c, bic_mat* E, const bic_mat* D, const bic_mat* A, DictWork* w, uint32_t* Hc,
                           unsigned long long* d_changed);
bic_status bic_k_dict_cursor(bic_ctx* c, DictWork* w, uint32_t* cursor_out);
bic_status bic_k_dict_commit(bic_ctx* c, bic_mat* D, DictWork* w);

// ------------------------------------------------------------------ initialize_model_neighbor, sharded
// src/bsvd.cpp:227-267 over the concatenation of all ranks' rows.
extern "C" bic_status bic_dist_initialize_model_neighbor(bic_ctx* c, bic_comm* m, const bic_mat* X, bic_mat* D, bic_mat* A,
                                                         uint64_t* rng_state) {
  if (!c || !m || !X || !D || !A || !rng_state) return BIC_ERR_INVALID;
  cudaSetDevice(c->device);
  const uint64_t p = D->rows;
  if (D->cols != X->cols || A->rows != X->rows || A->cols != p)
    return bic_fail(c, BIC_ERR_INVALID, "init: shapes must be X n x m, D p x m, A n x p");
  BIC_TRY(share_rows(c, m, X->rows));
  BIC_TRY(bic_mat_clear(c, A));
  BIC_TRY(bic_mat_clear(c, D));
  if (p == 0 || m->nglobal == 0 || X->cols == 0) return BIC_OK;
  // zero-row bitmaps of all ranks, padded to the largest shard
  uint64_t maxn = 0;
  for (uint64_t v : m->nrows) maxn = v > maxn ? v : maxn;
  const uint64_t bw = div_up_u64(maxn, 32);
  BIC_TRY(bic_scratch_reserve(c, &c->work[0], (size_t)bw * 4 * m->nranks + 16));
  uint32_t* d_bm = (uint32_t*)c->work[0].p;
  BIC_CUDA(c, cudaMemsetAsync(d_bm, 0, (size_t)bw * 4 * m->nranks, c->stream));
  BIC_TRY(bic_k_row_nonzero_bitmap(c, X, d_bm + (size_t)m->rank * bw));
  if (m->nranks > 1) {
    BIC_NCCL(c, nccl_api()->AllGather(d_bm + (size_t)m->rank * bw, d_bm, bw, ncclUint32, m->comm, c->stream));
    m->collectives++;
  }
  std::vector<uint32_t> bm((size_t)bw * m->nranks);
  BIC_CUDA(c, cudaMemcpyAsync(bm.data(), d_bm, bm.size() * 4, cudaMemcpyDeviceToHost, c->stream));
  BIC_CUDA(c, bic_wait_stream(c));
  bool any = false;
  for (size_t i = 0; i < bm.size() && !any; ++i) any = bm[i] != 0;
  if (!any) return bic_fail(c, BIC_ERR_INVALID, "init: X is all zero (the reference's draw loop never ends)");
  // every rank replays the same draw over the global row index (src/bsvd.cpp:239-243)
  std::vector<uint64_t> local_piv(p);
  std::vector<uint64_t> start(m->nranks + 1, 0);
  for (int r = 0; r < m->nranks; ++r) start[r + 1] = start[r] + m->nrows[r];
  for (uint64_t k = 0; k < p;) {
    const uint64_t g = bic_rand48_uniform_int(rng_state, m->nglobal);
    int owner = 0;
    while (g >= start[owner + 1]) owner++;
    const uint64_t li = g - start[owner];
    if (!((bm[(size_t)owner * bw + (li >> 5)] >> (li & 31)) & 1u)) continue;
    local_piv[k++] = (owner == m->rank) ? li : ~0ull;
  }
  InitWork w;
  BIC_TRY(bic_k_init_scratch(c, p, X->wpr, &w));
  BIC_TRY(bic_k_init_gather(c, X, local_piv.data(), &w));
  BIC_TRY(allreduce_u32(c, m, w.P, (size_t)p * X->wpr));          // each pivot row has exactly one owner
  BIC_TRY(bic_k_init_stats(c, X, &w));
  BIC_TRY(allreduce_u32(c, m, w.hist, (size_t)X->wpr * 32 + p));  // hist and usage are adjacent
  return bic_k_init_finalize(c, &w, X->cols, D);
}

bic_status bic_k_init_gather_dev(bic_ctx* c, const bic_mat* X, InitWork* w);
__global__ void k_draw_pivots(const uint32_t* bitmap, uint64_t n, uint32_t p, uint64_t* state, uint64_t* pivots, unsigned long long* status,
                              const uint64_t* starts, uint32_t nranks, uint64_t bw, uint32_t my_rank);

// The same with nothing waiting for the host (pipeline.cu): the bitmaps are gathered on the device, the draw runs there over the
// global row index (init.cu: k_draw_pivots, sharded variant -- every rank replays the same generator), pivot rows and statistics
// are combined with NCCL calls queued on the stream. d_state: the rand48 state (device u64, in/out, the same on every rank);
// d_status[0] != 0 afterwards: the whole matrix is zero. bic_comm_share_rows must have been called for this shard size.
bic_status bic_k_dist_init_async(bic_ctx* c, bic_comm* m, const bic_mat* X, bic_mat* D, bic_mat* A, uint64_t* d_state,
                                 unsigned long long* d_status) {
  const uint64_t p = D->rows;
  if (D->cols != X->cols || A->rows != X->rows || A->cols != p)
    return bic_fail(c, BIC_ERR_INVALID, "init: shapes must be X n x m, D p x m, A n x p");
  if (!bic_comm_rows_known(m, X->rows)) return bic_fail(c, BIC_ERR_INVALID, "init: shard sizes not exchanged (bic_comm_share_rows)");
  if (m->nglobal == 0 || m->nglobal > 0xFFFFFFFFull) return bic_fail(c, BIC_ERR_INVALID, "init: row count outside gsl_rng_uniform_int's range");
  BIC_TRY(bic_mat_clear(c, A));
  BIC_TRY(bic_mat_clear(c, D));
  if (p == 0 || X->cols == 0) return BIC_OK;
  uint64_t maxn = 0;
  for (uint64_t v : m->nrows) maxn = v > maxn ? v : maxn;
  const uint64_t bw = div_up_u64(maxn, 32);
  BIC_TRY(bic_scratch_reserve(c, &c->work[0], (size_t)bw * 4 * m->nranks + 16));
  uint32_t* d_bm = (uint32_t*)c->work[0].p;
  BIC_CUDA(c, cudaMemsetAsync(d_bm, 0, (size_t)bw * 4 * m->nranks, c->stream));
  BIC_TRY(bic_k_row_nonzero_bitmap(c, X, d_bm + (size_t)m->rank * bw));
  if (m->nranks > 1) {
    BIC_NCCL(c, nccl_api()->AllGather(d_bm + (size_t)m->rank * bw, d_bm, bw, ncclUint32, m->comm, c->stream));
    m->collectives++;
  }
  InitWork w;
  BIC_TRY(bic_k_init_scratch(c, p, X->wpr, &w));
  k_draw_pivots<<<1, 256, 0, c->stream>>>(d_bm, m->nglobal, (uint32_t)p, d_state, w.piv, d_status, m->d_starts, (uint32_t)m->nranks, bw,
                                          (uint32_t)m->rank);
  BIC_LAUNCH_CHECK(c);
  BIC_TRY(bic_k_init_gather_dev(c, X, &w));
  BIC_TRY(allreduce_u32(c, m, w.P, (size_t)p * X->wpr));          // each pivot row has exactly one owner
  BIC_TRY(bic_k_init_stats(c, X, &w));
  BIC_TRY(allreduce_u32(c, m, w.hist, (size_t)X->wpr * 32 + p));  // hist and usage are adjacent
  return bic_k_init_finalize(c, &w, X->cols, D);
}

// ---- sharded Golomb coding with nothing waiting for the host: coding2.cu runs the passes, the two exchanges are NCCL all-gathers
// queued on the same stream
bic_status bic_k_golomb_encode_multi_sharded(bic_ctx* c, const bic_mat* const* mats, int nmat, uint32_t chunk_samples, bic_stream* const* outs,
                                             unsigned long long* d_info, unsigned long long* d_shard, uint32_t nranks, uint32_t rank,
                                             bic_status (*allgather)(void* user, bic_ctx* c, const unsigned long long* d_src, int count,
                                                                     unsigned long long* d_dst),
                                             void* user);
static bic_status allgather_u64_dev(void* user, bic_ctx* c, const unsigned long long* d_src, int count, unsigned long long* d_dst) {
  bic_comm* m = (bic_comm*)user;
  if (m->nranks == 1) {
    BIC_CUDA(c, cudaMemcpyAsync(d_dst, d_src, 8 * (size_t)count, cudaMemcpyDeviceToDevice, c->stream));
    return BIC_OK;
  }
  BIC_NCCL(c, nccl_api()->AllGather(d_src, d_dst, (size_t)count, ncclUint64, m->comm, c->stream));
  m->collectives++;
  return BIC_OK;
}
bic_status bic_k_dist_golomb_async(bic_ctx* c, bic_comm* m, const bic_mat* const* mats, int nmat, uint32_t chunk_samples,
                                   bic_stream* const* outs, unsigned long long* d_info, unsigned long long* d_shard) {
  return bic_k_golomb_encode_multi_sharded(c, mats, nmat, chunk_samples, outs, d_info, d_shard, (uint32_t)m->nranks, (uint32_t)m->rank,
                                           allgather_u64_dev, m);
}

// ---- the dictionary update of a sharded learner iteration, queued (no host wait): dist_update_dictionary's chain path + the
// global changed-rows count joined back into d_counts[0]
bic_status bic_k_dist_iteration_dict(bic_ctx* c, bic_comm* m, bic_mat* E, bic_mat* D, const bic_mat* A, unsigned long long* d_counts);

// H += Hd; Hd = 0// H += Hd; Hd = 0// H += Hd; Hd = 0
__global__ void k_apply_corrections(uint32_t* __restrict__ H, uint32_t* __restrict__ Hd, uint64_t nwords) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t d = Hd[i];
    if (d) { H[i] += d; Hd[i] = 0; }
  }
}

// ------------------------------------------------------------------ peer windows
// Every rank allocates a window of the same size, the cudaIpc handles go round with one NCCL allgather, and every rank
// maps the others' windows. Collective: every rank must call it with the same size at the same point. BIC_DIST_FUSED=0
// (or any failure to map a peer) keeps the NCCL-only path.
static bic_status window_reserve(bic_ctx* c, bic_comm* m, size_t data_words) {
  if (m->fused == 0 || m->nranks == 1) return BIC_OK;
  if (m->fused == -1) {
    const char* e = getenv("BIC_DIST_FUSED");
    if ((e && e[0] == '0') || m->nranks > 64) { m->fused = 0; return BIC_OK; }
  }
  const size_t need = XWIN_DATA + data_words;
  if (m->win && m->win_words >= need) return BIC_OK;
  BIC_CUDA(c, bic_wait_stream(c));
  window_release(m);
  const size_t words = need + (need >> 2);
  int ok = 1;
  if (cudaMalloc((void**)&m->win, words * 4) != cudaSuccess) { cudaGetLastError(); m->win = nullptr; ok = 0; }
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) {
    cudaMemset(m->win, 0, words * 4);
    if (cudaIpcGetMemHandle(&mine, m->win) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  }
  // handles (64 bytes each) + an "ok" byte per rank through NCCL; staged in a small device buffer
  const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;
  // the staging record lives in the context's scalar area: no allocation here that could fail on one rank only -- a rank that
  // returned early would leave the others hanging in the allgather below
  if (rec * m->nranks > (BIC_SCALARS - 128) * 8) return bic_fail(c, BIC_ERR_UNSUPPORTED, "too many ranks for the peer-window exchange");
  uint8_t* d_rec = (uint8_t*)(c->d_scalars + 128);
  std::vector<uint8_t> h_rec(rec * m->nranks, 0);
  memcpy(h_rec.data() + rec * m->rank, &mine, sizeof(mine));
  h_rec[rec * m->rank + sizeof(mine)] = (uint8_t)ok;
  BIC_CUDA(c, cudaMemcpyAsync(d_rec + rec * m->rank, h_rec.data() + rec * m->rank, rec, cudaMemcpyHostToDevice, c->stream));
  BIC_NCCL(c, nccl_api()->AllGather(d_rec + rec * m->rank, d_rec, rec, ncclUint8, m->comm, c->stream));
  m->collectives++;
  BIC_CUDA(c, cudaMemcpyAsync(h_rec.data(), d_rec, rec * m->nranks, cudaMemcpyDeviceToHost, c->stream));
  BIC_CUDA(c, bic_wait_stream(c));
  for (int r = 0; r < m->nranks; ++r) ok &= h_rec[rec * r + sizeof(mine)];
  m->peer_win.assign(m->nranks, nullptr);
  if (ok) {
    for (int r = 0; r < m->nranks && ok; ++r) {
      if (r == m->rank) { m->peer_win[r] = m->win; continue; }
      cudaIpcMemHandle_t h;
      memcpy(&h, h_rec.data() + rec * r, sizeof(h));
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
      m->peer_win[r] = (uint32_t*)ptr;
    }
  }
  // every rank must take the same path: agree on the outcome
  uint64_t mine_ok = (uint64_t)ok, all_ok[16] = {0};
  if (m->nranks <= 8) {
    BIC_TRY(bic_comm_allgather_u64(c, m, &mine_ok, 1, all_ok));
    for (int r = 0; r < m->nranks; ++r) ok &= (int)all_ok[r];
  } else {
    ok = 0;
  }
  if (!ok) {
    window_release(m);
    m->fused = 0;
    return BIC_OK;
  }
  BIC_CUDA(c, cudaMalloc((void**)&m->d_peer_win, sizeof(uint32_t*) * m->nranks));
  BIC_CUDA(c, cudaMemcpy(m->d_peer_win, m->peer_win.data(), sizeof(uint32_t*) * m->nranks, cudaMemcpyHostToDevice));
  m->win_words = words;
  m->fused = 1;
  m->epoch = 0;
  return BIC_OK;
}

// barrier over the ranks as a kernel of its own (after a collective whose result the peers are about to add to)
__global__ void k_xgpu_barrier(XPeers x) {
  __threadfence_system();
  if (threadIdx.x == 0) {
    uint32_t* me = x.win[x.rank];
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

__global__ void k_split_u64_limbs(const unsigned long long* __restrict__ in, uint32_t* __restrict__ limbs) {
  const unsigned long long v = *in;
  limbs[0] = (uint32_t)(v & 0x3FFFFFu);
  limbs[1] = (uint32_t)((v >> 22) & 0x3FFFFFu);
  limbs[2] = (uint32_t)(v >> 44);   // < 2^20
}

bic_status bic_k_update_dictionary_v3x(bic_ctx* c, bic_mat* E, bic_mat* D, const bic_mat* A, unsigned long long* d_changed, V3Hook* hook);
bool bic_dict_chain_eligible(bic_ctx* c, uint64_t n, uint64_t p, uint64_t wprE);

// the sharded driver's side of dict3.cu's hook
struct ChainHookUser {
  bic_comm* m;
  unsigned long long* d_counts;
};
static bic_status chain_reduce(void* user, bic_ctx* c, uint32_t* buf, size_t words, uint32_t* extra) {
  ChainHookUser* u = (ChainHookUser*)user;
  // extra[0..2]: the 64-bit changed-rows count of the preceding coefficient update as three 22-bit limbs (see below)
  k_split_u64_limbs<<<1, 1, 0, c->stream>>>(u->d_counts, extra);
  BIC_LAUNCH_CHECK(c);
  u->m->last_extra = extra;
  return allreduce_u32(c, u->m, buf, words);
}
static size_t chain_window(void* user, size_t need_words, uint32_t** base, uint64_t* base_off) {
  ChainHookUser* u = (ChainHookUser*)user;
  (void)need_words;
  *base = u->m->win ? u->m->win + XWIN_DATA : nullptr;
  *base_off = XWIN_DATA;
  return u->m->win_words > XWIN_DATA ? u->m->win_words - XWIN_DATA : 0;
}

// ------------------------------------------------------------------ update_dictionary_steepest, sharded
// d_counts[0] (changed rows of the preceding coefficient update, local) is summed over ranks in the same
// allreduce as H and U; d_counts[1] receives the changed atoms (identical on every rank).
static bic_status dist_update_dictionary(bic_ctx* c, bic_comm* m, bic_mat* E, bic_mat* D, const bic_mat* A,
                                         unsigned long long* d_counts) {
  BIC_RANGE("bic:dist:update_dictionary");
  if (D->rows == 0 || E->cols == 0) return BIC_OK;
  // Where the histograms fit shared memory: the cluster chain (dict3.cu), one launch per update; the corrections of an atom
  // that changes go from rank to rank inside the chain kernel (peer windows). The row count that decides eligibility is the
  // largest shard's, so every rank takes the same path.
  {
    uint64_t maxn = E->rows;
    for (uint64_t v : m->nrows) maxn = v > maxn ? v : maxn;
    if (c->dict_algo == 2 && bic_dict_chain_eligible(c, maxn, D->rows, E->wpr)) {
      const size_t hwords = (size_t)D->rows * E->wpr * 32;
      const size_t glob = hwords + D->rows + A->wpr * 32 + 4;
      if (m->nranks > 1) BIC_TRY(window_reserve(c, m, glob + 128 + 2 * (size_t)m->nranks * hwords));
      if (m->nranks == 1 || m->fused == 1) {
        ChainHookUser u{m, d_counts};
        V3Hook hook;
        hook.reduce = chain_reduce;
        hook.user = &u;
        if (m->nranks > 1) {
          hook.window_words = chain_window;
          hook.x.win = m->d_peer_win;
          hook.x.nranks = (uint32_t)m->nranks;
          hook.x.rank = (uint32_t)m->rank;
        }
        return bic_k_update_dictionary_v3x(c, E, D, A, d_counts + 1, &hook);
      }
    }
  }
  DictWork w;
  // With peer windows the histograms live in the window: the fix kernel of an atom that changes adds its corrections
  // straight into EVERY rank's H over NVLink and ends with a barrier over the ranks (dict2.cu: hc_add, xgpu_barrier) --
  // compute and exchange are one kernel, there is no per-atom collective call. Without (BIC_DIST_FUSED=0, no peer
  // access) the corrections go to a delta buffer that is allreduced with NCCL after every step.
  const uint64_t hwords = D->rows * E->wpr * 32;
  BIC_TRY(window_reserve(c, m, hwords + D->rows + 64));
  const bool fused = (m->fused == 1);
  BIC_TRY(bic_k_dict_prepare(c, E, D, A, &w, fused ? m->win + XWIN_DATA : nullptr));
  // [H | U | extra]: extra[0..2] carries the 64-bit changed-rows count as three 22-bit limbs, so the u32 sums over the
  // ranks cannot overflow (<= 1024 ranks) and no carry between the limbs is lost
  k_split_u64_limbs<<<1, 1, 0, c->stream>>>(d_counts, w.extra);
  BIC_LAUNCH_CHECK(c);
  m->last_extra = w.extra;
  BIC_TRY(allreduce_u32(c, m, w.H, (size_t)w.p * w.hs + w.p + 3));
  if (fused) {
    w.x.win = m->d_peer_win;
    w.x.nranks = (uint32_t)m->nranks;
    w.x.rank = (uint32_t)m->rank;
    w.x.h_off = XWIN_DATA;
    // no rank may add to a peer's H before that peer's allreduce has delivered it
    XPeers b = w.x;
    b.epoch = ++m->epoch;
    k_xgpu_barrier<<<1, 32, 0, c->stream>>>(b);
    BIC_LAUNCH_CHECK(c);
    w.x.epoch = m->epoch;  // launch i of this update uses barrier number epoch + i + 1
  }
  uint32_t cursor = 0;
  uint32_t batch = 1;  // most updates after the first iteration change no atom: one launch, one look at the cursor
  for (;;) {
    for (uint32_t i = 0; i < batch && w.launched < w.p; ++i) {
      if (fused) {
        BIC_TRY(bic_k_dict_step(c, E, D, A, &w, w.H, d_counts + 1));
      } else {
        BIC_TRY(bic_k_dict_step(c, E, D, A, &w, w.Hd, d_counts + 1));
        // an atom may have changed: combine the corrections its users produced on every rank
        BIC_TRY(allreduce_u32(c, m, w.Hd, hwords));
        k_apply_corrections<<<bic_grid_for(c, hwords, 256, 2), 256, 0, c->stream>>>(w.H, w.Hd, hwords);
        BIC_LAUNCH_CHECK(c);
      }
    }
    BIC_TRY(bic_k_dict_cursor(c, &w, &cursor));
    if (cursor >= w.p || w.launched >= w.p) break;
    batch = (batch * 2 < 32) ? batch * 2 : 32;
  }
  if (fused) m->epoch += w.launched + 1;
  return bic_k_dict_commit(c, D, &w);
}

__global__ void k_join_u64_limbs(const uint32_t* __restrict__ limbs, unsigned long long* __restrict__ out) {
  // each limb is a sum over the ranks of a 22-bit piece: the pieces overlap after the sum, plain 64-bit adds recombine them
  *out = (unsigned long long)limbs[0] + ((unsigned long long)limbs[1] << 22) + ((unsigned long long)limbs[2] << 44);
}

extern "C" bic_status bic_dist_update_dictionary_steepest(bic_ctx* c, bic_comm* m, bic_mat* E, bic_mat* D, const bic_mat* A,
                                                          uint64_t* changed) {
  if (!c || !m || !E || !D || !A) return BIC_ERR_INVALID;
  cudaSetDevice(c->device);
  BIC_CUDA(c, cudaMemsetAsync(c->d_scalars, 0, 16, c->stream));
  BIC_TRY(dist_update_dictionary(c, m, E, D, A, (unsigned long long*)c->d_scalars));
  BIC_TRY(bic_read_scalars(c, 2));
  if (changed) *changed = c->h_scalars[1];
  return BIC_OK;
}

// ------------------------------------------------------------------ learn_model_traditional, sharded
extern "C" bic_status bic_dist_learn_model_traditional(bic_ctx* c, bic_comm* m, const bic_mat* X, bic_mat* E, bic_mat* D,
                                                       bic_mat* A, uint64_t* iterations, uint64_t* trace, uint64_t trace_cap) {
  if (!c || !m || !X || !E || !D || !A) return BIC_ERR_INVALID;
  cudaSetDevice(c->device);
  BIC_TRY(bic_residual(c, X, A, D, E));
  uint64_t changed = 1, iter = 0;
  unsigned long long* d_cc = (unsigned long long*)c->d_scalars;
  while (changed > 0) {
    iter++;
    BIC_CUDA(c, cudaMemsetAsync(c->d_scalars, 0, 16, c->stream));
    BIC_TRY(bic_k_update_coefficients(c, E, D, A, d_cc));
    // global changed-rows: rides in the dictionary update's first allreduce (extra[0..1]); read it back
    // from there after the update
    BIC_TRY(dist_update_dictionary(c, m, E, D, A, d_cc));
    // extra[] sits right after H and U, wherever the update put them (scratch or peer window)
    {
      const uint64_t p = D->rows;
      const uint32_t* extra = m->last_extra;
      if (p && E->cols) {
        k_join_u64_limbs<<<1, 1, 0, c->stream>>>(extra, d_cc);
        BIC_LAUNCH_CHECK(c);
      } else if (m->nranks > 1) {
        BIC_NCCL(c, nccl_api()->AllReduce(d_cc, d_cc, 1, ncclUint64, ncclSum, m->comm, c->stream));
        m->collectives++;
      }
    }
    BIC_TRY(bic_read_scalars(c, 2));
    changed = c->h_scalars[0] + c->h_scalars[1];
    if (trace && iter <= trace_cap) { trace[2 * (iter - 1)] = c->h_scalars[0]; trace[2 * (iter - 1) + 1] = c->h_scalars[1]; }
    if (changed > 0 && c->h_scalars[1] == 0) {  // the next iteration provably changes nothing (see bic_learn_model_traditional)
      iter++;
      if (trace && iter <= trace_cap) { trace[2 * (iter - 1)] = 0; trace[2 * (iter - 1) + 1] = 0; }
      break;
    }
  }
  if (iterations) *iterations = iter;
  return BIC_OK;
}


bic_status bic_k_dist_iteration_dict(bic_ctx* c, bic_comm* m, bic_mat* E, bic_mat* D, const bic_mat* A, unsigned long long* d_counts) {
  uint64_t maxn = E->rows;
  for (uint64_t v : m->nrows) maxn = v > maxn ? v : maxn;
  if (!(c->dict_algo == 2 && bic_dict_chain_eligible(c, maxn, D->rows, E->wpr)) || (m->nranks > 1 && m->fused != 1))
    return bic_fail(c, BIC_ERR_UNSUPPORTED, "sharded pipeline: the dictionary update must take the cluster-chain path with peer windows");
  BIC_TRY(dist_update_dictionary(c, m, E, D, A, d_counts));
  k_join_u64_limbs<<<1, 1, 0, c->stream>>>(m->last_extra, d_counts);   // the changed-rows count summed over the ranks
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

// first use of a shape on a communicator: shard sizes and the peer window (both block and are collective: every rank calls this
// for its slots in the same order, before any job of that shape is queued)
bic_status bic_k_dist_prepare(bic_ctx* c, bic_comm* m, uint64_t n_local, uint64_t p, uint64_t wprE, uint64_t wprA) {
  BIC_TRY(share_rows(c, m, n_local));
  if (m->nranks > 1) {
    const size_t hwords = (size_t)p * wprE * 32;
    BIC_TRY(window_reserve(c, m, hwords + p + wprA * 32 + 4 + 128 + 2 * (size_t)m->nranks * hwords));
  }
  return BIC_OK;
}
bool bic_k_dist_fused(const bic_comm* m) { return m->nranks == 1 || m->fused == 1; }
