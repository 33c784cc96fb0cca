// Internal declarations shared by the translation units of libbic_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "bic_b200.h"

// NVTX ranges per phase of the path (tracing): visible in Nsight Systems / ncu --nvtx, free when no tool is attached
// (nvtx3 is header-only and resolves its injection library lazily).
#include <nvtx3/nvToolsExt.h>
struct bic_nvtx_range {
  explicit bic_nvtx_range(const char* name) { nvtxRangePushA(name); }
  ~bic_nvtx_range() { nvtxRangePop(); }
};
#define BIC_RANGE(name) bic_nvtx_range _bic_range_(name)

// ---------------------------------------------------------------------------------------------
// Device layout of a bit matrix: uint32 words, MSB first (bit j of a row is bit 31-(j&31) of
// word j>>5), stride = ceil(cols/32) words, pad bits zero. The allocation is rounded up to a
// multiple of 256 B and zero filled so 128-bit loads may run past the last row harmlessly.
// ---------------------------------------------------------------------------------------------
struct bic_mat {
  uint64_t rows = 0, cols = 0;
  uint64_t wpr = 0;         // 32-bit words per row
  uint32_t* d = nullptr;    // device words
  size_t alloc_bytes = 0;
  bool owns = true;
  bool pooled = false;      // allocated with cudaMallocAsync on its context's stream (library-internal temporaries)
  uint64_t words() const { return rows * wpr; }
};

struct bic_scratch {
  void* p = nullptr;
  size_t bytes = 0;
};

struct bic_stream {
  bic_stream_info info{};
  uint8_t* d_bytes = nullptr;   // device byte stream
  size_t cap_bytes = 0;
  uint64_t* d_index = nullptr;  // 2 * nchunks
  size_t cap_index = 0;         // in uint64 entries
};

// kernels of the library, for the per-kernel device timers (bic_prof_*)
enum bic_kernel_id {
  KID_WORDS64_TO_DEV = 0, KID_DEV_TO_WORDS64, KID_PBM_TO_DEV, KID_DEV_TO_PBM, KID_WEIGHT, KID_XOR,
  KID_EXTRACT, KID_ASSEMBLE, KID_ROW_NONZERO, KID_GATHER_ROWS, KID_COL_HIST, KID_PIVOT_USAGE, KID_INIT_FINALIZE,
  KID_UPDATE_COEF, KID_RESIDUAL, KID_TRANSPOSE_BITS, KID_UPDATE_DICT, KID_DICT_HIST, KID_DICT_RESOLVE, KID_DICT_SCAN,
  KID_COMPACT_ROWS, KID_EXPAND_ROWS, KID_GOL_TILE_COUNTS, KID_GOL_SCAN_A, KID_GOL_LENGTHS, KID_GOL_SCAN_B,
  KID_GOL_SCATTER, KID_GOL_DECODE, KID_EG_FIRST, KID_EG_FILL, KID_EG_ENCODE, KID_EG_DECODE, KID_DICT_CHAIN, KID_DICT_APPLY, KID_DICT_COMPACT, KID_DICT_BUCKET, KID_BITPLANES, KID_PROXIMUS, KID_MATCH, KID_MATCH_DECIDE,
  KID_COUNT
};

#define BIC_SCALARS 512
#define BIC_SCALAR_COEF_PASSES 100   // d_scalars slot: greedy passes (row x pass) of the lane-per-row coefficient kernel since the counter was last read

struct bic_prof_rec { int kid; cudaEvent_t e0, e1; };

struct bic_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = true;
  int sm_count = 0;
  size_t smem_optin = 0;
  uint64_t launches = 0;
  std::string err;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // pinned host scalars for results read back after a kernel
  uint64_t* h_scalars = nullptr;  // BIC_SCALARS entries
  uint64_t* d_scalars = nullptr;  // BIC_SCALARS entries; [0, 64) per-call results, [128, 512) staging of small collectives
  // grow-only scratch areas
  bic_scratch staging;    // host-layout staging for uploads/downloads
  bic_scratch work[6];    // per-algorithm work buffers
  // optional per-launch device timers
  int wait_mode = 0;       // how host threads wait for the stream: 0 cudaStreamSynchronize, 1 poll + sched_yield, 2 blocking event
  cudaEvent_t wait_ev = nullptr, wait_ev_blocking = nullptr;
  int gol_algo = 2;        // 2: wide-tile encoder with fused scans (coding2.cu), 1: the first formulation (coding.cu)
  int gol_list = 1;        // sparse tiles coded from a list of their ones (coding2.cu): 0 never, 1 for streams long enough for wide tiles, 2 always
  int gol_scan = 1;        // tile scans of coding2.cu: 0 in the passes' last CTA, 1 as their own launch for long streams (>= 8192 tiles), 2 always
  int gol_presize_pct = 125;  // the asynchronous encoders size the code buffer to this percentage of the input bits (+ 4 KB)
  int gol_onepass = 0;     // 1: single-pass Golomb encoder (decoupled look-back) when the buffer is pre-sized
  int coef_algo = 1;  // 1: dictionaries of >= 64 atoms use the weight-sorted warp-per-row coefficient kernel; 0: always lane per row
  int dict_update = 0;  // which dictionary update the learners call: 0 update_dictionary_steepest, 1 update_dictionary_proximus (the reference's -d 1)
  int dict_algo = 2;  // 0: per-atom walk (dict.cu), 1: histogram first, resolve in order (dict2.cu), 2: cluster chain (dict3.cu) where the shape allows, else 1
  long long chain_bucket_cap = -1;  // entries of dict3.cu's per-atom buckets; -1 = 2 per row (0 forces the list-scan fallback)
  int chain_cluster = 16;  // CTAs in the cluster of dict3.cu's chain kernel (1, 2, 4, 8 or 16)
  // device flag of a learner loop that runs ahead of the host (pipeline.cu): non-null while iterations are queued without
  // waiting for the previous one's counts; the iteration's kernels return at once when *loop_skip != 0
  const uint32_t* loop_skip = nullptr;
  bool prof_on = false;
  // A context that is one slot of a SHARDED pipeline must never call cudaFree while jobs are in flight: cudaFree waits for every
  // stream of the device, including other slots' kernels that are waiting for a peer GPU -- whose host may be waiting the same
  // way for ours. Such contexts park outgrown blocks here until they are destroyed.
  bool defer_free = false;
  std::vector<void*> graveyard;
  std::vector<bic_prof_rec> prof_recs;
  std::vector<cudaEvent_t> prof_free;
  double prof_ms[KID_COUNT] = {0};
  uint64_t prof_n[KID_COUNT] = {0};
};

void bic_prof_begin(bic_ctx* c, int kid);
void bic_prof_end(bic_ctx* c);
#define BIC_PROF(ctx, kid) do { if ((ctx)->prof_on) bic_prof_begin((ctx), (kid)); } while (0)

#define BIC_CUDA(ctx, expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                   \
      return BIC_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define BIC_TRY(expr)                 \
  do {                                \
    bic_status _s = (expr);           \
    if (_s != BIC_OK) return _s;      \
  } while (0)

#define BIC_LAUNCH_CHECK(ctx)                                   \
  do {                                                          \
    (ctx)->launches++;                                          \
    if ((ctx)->prof_on) bic_prof_end(ctx);                      \
    BIC_CUDA(ctx, cudaGetLastError());                          \
  } while (0)

static inline bic_status bic_fail(bic_ctx* ctx, bic_status s, const char* msg) {
  if (ctx) ctx->err = msg;
  return s;
}

// wait until everything queued on the context's stream is done (honours wait_mode)
cudaError_t bic_wait_stream(bic_ctx* ctx);
// grow-only scratch; contents are undefined after growth
static inline void bic_free_device(bic_ctx* c, void* p) {
  if (!p) return;
  if (c && c->defer_free) c->graveyard.push_back(p);
  else cudaFree(p);
}
bic_status bic_scratch_reserve(bic_ctx* ctx, bic_scratch* s, size_t bytes);
// D2H of the first n scalar slots after the stream drained
bic_status bic_read_scalars(bic_ctx* ctx, int n);
bic_status bic_zero_scalars(bic_ctx* ctx);

// device-side bounds checks of the debug build (make -C csrc debug): a violated check prints its place and traps, which the
// host sees as a failed launch -- the memcheck of a pool where compute-sanitizer is not available
#if defined(BIC_DEBUG_CHECKS) && defined(__CUDA_ARCH__)
#define BIC_DCHECK(cond)                                                                          \
  do {                                                                                            \
    if (!(cond)) {                                                                                \
      printf("BIC_DCHECK failed: %s at %s:%d (block %d thread %d)\n", #cond, __FILE__, __LINE__, \
             (int)blockIdx.x, (int)threadIdx.x);                                                  \
      __trap();                                                                                   \
    }                                                                                             \
  } while (0)
#else
#define BIC_DCHECK(cond) do { } while (0)
#endif

#ifdef __CUDACC__
#define BIC_HD __host__ __device__
#else
#define BIC_HD
#endif
BIC_HD static inline uint64_t div_up_u64(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// rows * ceil(cols / 32) * 4 bytes and rows * cols bits must fit comfortably in 64 bits (and in what a device can hold:
// 2^44 bytes = 16 TB is far beyond any HBM, so anything larger is a corrupt header or a caller bug, not a request)
static inline bool bic_shape_ok(uint64_t rows, uint64_t cols) {
  if (cols > (1ull << 40) || rows > (1ull << 40)) return false;
  const uint64_t wpr = div_up_u64(cols, 32);
  if (rows && wpr > (1ull << 42) / rows) return false;
  return true;
}

// persistent-grid sizing: a multiple of the SM count
static inline int bic_grid_for(const bic_ctx* ctx, uint64_t work_items, int per_block, int blocks_per_sm) {
  uint64_t need = div_up_u64(work_items ? work_items : 1, (uint64_t)per_block);
  uint64_t cap = (uint64_t)ctx->sm_count * (uint64_t)blocks_per_sm;
  return (int)(need < cap ? need : cap);
}

// One problem of a batch of equally shaped fits (batch.cu): kernels launched with a problem table take
// their pointers from probs[blockIdx.y or .z] and return at once for problems that already converged.
struct ProbDev {
  uint32_t *E, *D, *A, *AT, *H, *U, *Dnew, *cursor, *first;
  unsigned long long* counts;  // [0] changed rows, [1] changed atoms of the current iteration
};

// Peers of a row-sharded fit, for kernels that exchange over NVLink peer memory (dist.cu): every rank maps every other
// rank's window with cudaIpc; all pointers are valid on THIS device. Window layout in u32 words: [0, 64) flags (flags[r] =
// the last barrier rank r has reached), [64] arrival counter of the local grid, [XWIN_DATA, ...) data.
#define XWIN_DATA 128
struct XPeers {
  uint32_t* const* win = nullptr;  // [nranks] window of every rank (own included)
  uint32_t nranks = 0, rank = 0;
  uint32_t epoch = 0;              // number of this launch's barrier: strictly increasing, identical on every rank
  uint64_t h_off = 0;              // word offset of the histograms H inside a window
};

// Row-sharded cluster chain (dict3.cu + dist.cu): the corrections of an atom that changes are exchanged between the ranks from
// inside the chain kernel. Window words used (header, < XWIN_DATA): [XWIN_EPOCH2] this rank's exchange counter (local),
// [XWIN_FLAGS2 + r] the last exchange rank r has pushed into this window. The exchange area holds 2 (parity) x nranks vectors of
// p * hs counters at word offset xoff of every window.
#define XWIN_EPOCH2 95
#define XWIN_FLAGS2 96
struct ChainDist {
  uint32_t* const* win = nullptr;  // [nranks] every rank's window as mapped on this device
  uint32_t nranks = 1, rank = 0;
  uint64_t xoff = 0;
};
// what the sharded driver plugs into the update: `reduce` sums buf[0 .. words) over the ranks (it is called once, after the
// local histograms, usage counts and bucket sizes are complete and `extra` -- 4 words at the end of buf -- may be filled in)
struct V3Hook {
  bic_status (*reduce)(void* user, bic_ctx* c, uint32_t* buf, size_t words, uint32_t* extra) = nullptr;
  void* user = nullptr;
  ChainDist x;
  // this rank's window: *base = the first usable word (a device pointer), *base_off = its word offset inside the window;
  // returns the words available from there
  size_t (*window_words)(void* user, size_t need_words, uint32_t** base, uint64_t* base_off) = nullptr;
};

// scratch layout of one dictionary update (dict2.cu), shared with the row-sharded driver (dist.cu)
struct DictWork {
  uint64_t n, p, wpr, hs, wprN;
  uint32_t *AT, *H, *U, *extra, *Hd, *Dnew, *cursor, *first;
  uint32_t launched;
  bool use_scan;
  XPeers x;                        // nranks > 1: corrections are pushed into every rank's H by the fix kernel itself
};

// scratch layout of one neighbour initialisation (init.cu), shared with dist.cu
struct InitWork {
  uint64_t p, wpr;
  uint64_t* piv;
  uint32_t *P, *hist, *usage;
};

// ---- container (encoder.cu writes it synchronously, pipeline.cu asynchronously): little-endian u64 fields
//   [0] magic "BICB200\0"  [1] version  [2] rows [3] cols [4] W [5] K [6] n [7] m [8] iterations [9] seed
//   then per stream (D, A, E) 7 fields: coder, chunk_samples, rows, cols, bitcount, nsamples, nchunks
//   then per stream: code bytes padded to 8, chunk index (nchunks * 2 u64)
static const uint64_t BIC_MAGIC = 0x0030303242434942ull;  // "BICB200\0"
static const uint64_t BIC_HDR_FIELDS = 10, BIC_STREAM_FIELDS = 7;
static inline uint64_t bic_container_bytes(const bic_stream_info si[3]) {
  uint64_t need = (BIC_HDR_FIELDS + 3 * BIC_STREAM_FIELDS) * 8;
  for (int i = 0; i < 3; ++i) need += div_up_u64(div_up_u64(si[i].bitcount, 8), 8) * 8 + si[i].nchunks * 16;
  return need;
}
static inline void bic_container_header(uint8_t* out, uint64_t rows, uint64_t cols, uint64_t W, uint64_t K, uint64_t n, uint64_t m,
                                        uint64_t iters, uint64_t seed, const bic_stream_info si[3]) {
  uint64_t* h = (uint64_t*)out;
  h[0] = BIC_MAGIC; h[1] = 1; h[2] = rows; h[3] = cols; h[4] = W; h[5] = K; h[6] = n; h[7] = m; h[8] = iters; h[9] = seed;
  for (int i = 0; i < 3; ++i) {
    uint64_t* f = h + BIC_HDR_FIELDS + i * BIC_STREAM_FIELDS;
    f[0] = si[i].coder; f[1] = si[i].chunk_samples; f[2] = si[i].rows; f[3] = si[i].cols; f[4] = si[i].bitcount; f[5] = si[i].nsamples;
    f[6] = si[i].nchunks;
  }
}

// a matrix from the stream-ordered pool: create and destroy cost no device synchronisation. Only for temporaries that are
// used and destroyed on ONE context (the one passed here).
bic_status bic_mat_create_pooled(bic_ctx* ctx, uint64_t rows, uint64_t cols, bic_mat** out);

// ---- cross-TU entry points (implemented next to their kernels) -------------------------------
bic_status bic_k_row_nonzero_bitmap(bic_ctx* ctx, const bic_mat* X, uint32_t* d_bitmap);

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// 32x32 bit-matrix transpose across a warp. In: lane r holds row r (bit c of the row at bit
// 31-c of the word, MSB first). Out: lane c holds column c (row r at bit 31-r). Five butterfly
// stages, each one shuffle + a masked merge.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x) {
  const unsigned lane = threadIdx.x & 31;
#define BIC_T32_STAGE(S, HI)                                                         \
  {                                                                                   \
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, S);                            \
    x = (lane & S) ? (((y << S) & (HI)) | (x & ~(HI))) : ((x & (HI)) | ((y >> S) & ~(HI))); \
  }
  BIC_T32_STAGE(16, 0xFFFF0000u)
  BIC_T32_STAGE(8, 0xFF00FF00u)
  BIC_T32_STAGE(4, 0xF0F0F0F0u)
  BIC_T32_STAGE(2, 0xCCCCCCCCu)
  BIC_T32_STAGE(1, 0xAAAAAAAAu)
#undef BIC_T32_STAGE
  return x;
}
// ---- TMA bulk copy (cp.async.bulk, 1-D) of a contiguous global block into shared memory, completion
// signalled on an mbarrier. Used to stage the dictionary (the atom tile every row of a CTA is matched
// against). bytes must be a multiple of 16, both addresses 16-byte aligned.
__device__ __forceinline__ void tma_stage_begin(uint64_t* bar) {  // one thread, before the copies
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_stage_copy(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done = 0;
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  while (done < bytes) {  // pieces of at most 32 KB
    const uint32_t n = (bytes - done < 32768u) ? bytes - done : 32768u;
    const uint32_t d = (uint32_t)__cvta_generic_to_shared((char*)smem_dst + done);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"((const char*)gmem_src + done), "r"(n), "r"(b) : "memory");
    done += n;
  }
}
__device__ __forceinline__ void tma_stage_wait(uint64_t* bar) {  // every thread that reads the staged data
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(b) : "memory");
    if (spin > (1u << 26)) __trap();  // never spin forever on a lost copy
  }
}

// mask of the valid bits of the last word of a row with `cols` columns
__host__ __device__ __forceinline__ uint32_t tail_mask32(uint64_t cols) {
  const unsigned r = (unsigned)(cols & 31);
  return r ? (0xFFFFFFFFu << (32 - r)) : 0xFFFFFFFFu;
}
#endif
