// Context, device bit matrices, host<->device layout conversion and the small binmat
// reductions (weight / dist / xor). Reference interfaces: src/binmat.h:29-232.
#include "bic_internal.cuh"

#include <mutex>

#include <new>

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
static bic_status ctx_init(bic_ctx* c, int device, void* stream, bool own) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return BIC_ERR_NO_DEVICE;
  if (device < 0 || device >= ndev) return BIC_ERR_INVALID;
  c->device = device;
  BIC_CUDA(c, cudaSetDevice(device));
  cudaDeviceProp prop;
  BIC_CUDA(c, cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  if (own) {
    BIC_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  } else {
    c->stream = (cudaStream_t)stream;
  }
  c->owns_stream = own;
  BIC_CUDA(c, cudaEventCreate(&c->ev0));
  BIC_CUDA(c, cudaEventCreate(&c->ev1));
  BIC_CUDA(c, cudaEventCreateWithFlags(&c->wait_ev, cudaEventDisableTiming));
  BIC_CUDA(c, cudaEventCreateWithFlags(&c->wait_ev_blocking, cudaEventDisableTiming | cudaEventBlockingSync));
  BIC_CUDA(c, cudaMallocHost(&c->h_scalars, BIC_SCALARS * sizeof(uint64_t)));
  BIC_CUDA(c, cudaMalloc(&c->d_scalars, BIC_SCALARS * sizeof(uint64_t)));
  BIC_CUDA(c, cudaMemsetAsync(c->d_scalars, 0, BIC_SCALARS * sizeof(uint64_t), c->stream));
  return BIC_OK;
}

extern "C" bic_status bic_ctx_create(int device, bic_ctx** out) {
  if (!out) return BIC_ERR_INVALID;
  *out = nullptr;
  bic_ctx* c = new (std::nothrow) bic_ctx();
  if (!c) return BIC_ERR_NOMEM;
  bic_status s = ctx_init(c, device, nullptr, true);
  if (s != BIC_OK) { delete c; return s; }
  *out = c;
  return BIC_OK;
}

extern "C" bic_status bic_ctx_create_on_stream(int device, void* cuda_stream, bic_ctx** out) {
  if (!out) return BIC_ERR_INVALID;
  *out = nullptr;
  bic_ctx* c = new (std::nothrow) bic_ctx();
  if (!c) return BIC_ERR_NOMEM;
  bic_status s = ctx_init(c, device, cuda_stream, false);
  if (s != BIC_OK) { delete c; return s; }
  *out = c;
  return BIC_OK;
}

extern "C" void bic_internal_drop_workspace(bic_ctx* c);
static void prof_collect_for_destroy(bic_ctx* c) {
  cudaStreamSynchronize(c->stream);
  for (auto& r : c->prof_recs) { if (r.e0) cudaEventDestroy(r.e0); if (r.e1) cudaEventDestroy(r.e1); }
  c->prof_recs.clear();
  for (auto e : c->prof_free) cudaEventDestroy(e);
  c->prof_free.clear();
}

extern "C" bic_status bic_ctx_destroy(bic_ctx* c) {
  if (!c) return BIC_ERR_INVALID;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  bic_internal_drop_workspace(c);
  prof_collect_for_destroy(c);
  if (c->staging.p) cudaFree(c->staging.p);
  for (auto& w : c->work) if (w.p) cudaFree(w.p);
  for (void* p : c->graveyard) cudaFree(p);
  c->graveyard.clear();
  if (c->h_scalars) cudaFreeHost(c->h_scalars);
  if (c->d_scalars) cudaFree(c->d_scalars);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->wait_ev) cudaEventDestroy(c->wait_ev);
  if (c->wait_ev_blocking) cudaEventDestroy(c->wait_ev_blocking);
  if (c->owns_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return BIC_OK;
}

extern "C" bic_status bic_ctx_sync(bic_ctx* c) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c) return BIC_ERR_INVALID;
  BIC_CUDA(c, bic_wait_stream(c));
  return BIC_OK;
}

extern "C" const char* bic_ctx_last_error(bic_ctx* c) { return c ? c->err.c_str() : "null context"; }

extern "C" const char* bic_status_string(bic_status s) {
  switch (s) {
    case BIC_OK: return "ok";
    case BIC_ERR_INVALID: return "invalid argument";
    case BIC_ERR_CUDA: return "CUDA error";
    case BIC_ERR_NOMEM: return "out of memory";
    case BIC_ERR_CAPACITY: return "buffer too small";
    case BIC_ERR_NO_DEVICE: return "no CUDA device (libbic_b200 has no CPU fallback)";
    case BIC_ERR_CORRUPT: return "corrupt stream";
    case BIC_ERR_UNSUPPORTED: return "unsupported";
    default: return "unknown";
  }
}

// counters the kernels keep on the device for the measurement harness; reading one waits for the stream and resets it
extern "C" bic_status bic_ctx_read_counter(bic_ctx* c, const char* name, uint64_t* value) {
  if (!c || !name || !value) return BIC_ERR_INVALID;
  cudaSetDevice(c->device);
  int slot = -1;
  if (!strcmp(name, "coef_passes")) slot = BIC_SCALAR_COEF_PASSES;
  if (slot < 0) return bic_fail(c, BIC_ERR_INVALID, "unknown counter");
  BIC_CUDA(c, cudaMemcpyAsync(c->h_scalars + slot, c->d_scalars + slot, 8, cudaMemcpyDeviceToHost, c->stream));
  BIC_CUDA(c, cudaMemsetAsync(c->d_scalars + slot, 0, 8, c->stream));
  BIC_CUDA(c, bic_wait_stream(c));
  *value = c->h_scalars[slot];
  return BIC_OK;
}

extern "C" void* bic_ctx_cuda_stream(bic_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" int bic_ctx_sm_count(bic_ctx* c) { return c ? c->sm_count : 0; }
extern "C" uint64_t bic_ctx_launch_count(bic_ctx* c) { return c ? c->launches : 0; }

extern "C" bic_status bic_timer_start(bic_ctx* c) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c) return BIC_ERR_INVALID;
  BIC_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  return BIC_OK;
}
extern "C" bic_status bic_timer_stop(bic_ctx* c, float* ms) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !ms) return BIC_ERR_INVALID;
  BIC_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  BIC_CUDA(c, cudaEventSynchronize(c->ev1));
  BIC_CUDA(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return BIC_OK;
}

extern "C" bic_status bic_host_alloc(size_t bytes, void** out) {
  if (!out) return BIC_ERR_INVALID;
  return cudaMallocHost(out, bytes ? bytes : 1) == cudaSuccess ? BIC_OK : BIC_ERR_NOMEM;
}
extern "C" bic_status bic_host_free(void* p) {
  return cudaFreeHost(p) == cudaSuccess ? BIC_OK : BIC_ERR_CUDA;
}

#include <sched.h>
// Many contexts per process (one per page in flight) and several processes per box (one per GPU) mean
// many host threads waiting at once; spinning inside cudaStreamSynchronize then oversubscribes the cores.
cudaError_t bic_wait_stream(bic_ctx* c) {
  if (c->wait_mode == 1) {
    cudaError_t e = cudaEventRecord(c->wait_ev, c->stream);
    if (e != cudaSuccess) return e;
    for (;;) {
      e = cudaEventQuery(c->wait_ev);
      if (e != cudaErrorNotReady) return e;
      sched_yield();
    }
  }
  if (c->wait_mode == 2) {
    cudaError_t e = cudaEventRecord(c->wait_ev_blocking, c->stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(c->wait_ev_blocking);
  }
  return cudaStreamSynchronize(c->stream);
}

bic_status bic_scratch_reserve(bic_ctx* c, bic_scratch* s, size_t bytes) {
  if (bytes <= s->bytes) return BIC_OK;
  // the old block may still be in use by queued kernels
  BIC_CUDA(c, bic_wait_stream(c));
  if (s->p) { bic_free_device(c, s->p); s->p = nullptr; s->bytes = 0; }
  size_t want = (bytes + (bytes >> 2) + 255) & ~(size_t)255;
  cudaError_t e = cudaMalloc(&s->p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = (bytes + 255) & ~(size_t)255;
    e = cudaMalloc(&s->p, want);
    if (e != cudaSuccess) { cudaGetLastError(); c->err = "scratch allocation failed"; return BIC_ERR_NOMEM; }
  }
  s->bytes = want;
  return BIC_OK;
}

bic_status bic_read_scalars(bic_ctx* c, int n) {
  BIC_CUDA(c, cudaMemcpyAsync(c->h_scalars, c->d_scalars, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, c->stream));
  BIC_CUDA(c, bic_wait_stream(c));
  return BIC_OK;
}
bic_status bic_zero_scalars(bic_ctx* c) {
  BIC_CUDA(c, cudaMemsetAsync(c->d_scalars, 0, 64 * sizeof(uint64_t), c->stream));
  return BIC_OK;
}

// ---------------------------------------------------------------------------------------------
// matrices
// ---------------------------------------------------------------------------------------------
extern "C" bic_status bic_mat_create(bic_ctx* c, uint64_t rows, uint64_t cols, bic_mat** out) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !out) return BIC_ERR_INVALID;
  *out = nullptr;
  bic_mat* m = new (std::nothrow) bic_mat();
  if (!m) return BIC_ERR_NOMEM;
  m->rows = rows;
  m->cols = cols;
  m->wpr = div_up_u64(cols, 32);
  if (!bic_shape_ok(rows, cols)) { delete m; c->err = "matrix shape overflows"; return BIC_ERR_INVALID; }
  size_t bytes = (size_t)(m->rows * m->wpr) * 4;
  m->alloc_bytes = ((bytes + 255) & ~(size_t)255) + 256;
  cudaSetDevice(c->device);
  if (cudaMalloc(&m->d, m->alloc_bytes) != cudaSuccess) {
    cudaGetLastError();
    delete m;
    c->err = "cudaMalloc failed for matrix";
    return BIC_ERR_NOMEM;
  }
  cudaError_t e = cudaMemsetAsync(m->d, 0, m->alloc_bytes, c->stream);
  if (e != cudaSuccess) { cudaFree(m->d); delete m; c->err = cudaGetErrorString(e); return BIC_ERR_CUDA; }
  *out = m;
  return BIC_OK;
}

bic_status bic_mat_create_pooled(bic_ctx* c, uint64_t rows, uint64_t cols, bic_mat** out) {
  *out = nullptr;
  static bool threshold_set[64] = {false};
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (c->device < 64 && !threshold_set[c->device]) {  // keep freed blocks in the pool instead of returning them to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    threshold_set[c->device] = true;
  }
  bic_mat* m = new (std::nothrow) bic_mat();
  if (!m) return BIC_ERR_NOMEM;
  m->rows = rows;
  m->cols = cols;
  m->wpr = div_up_u64(cols, 32);
  m->pooled = true;
  if (!bic_shape_ok(rows, cols)) { delete m; c->err = "matrix shape overflows"; return BIC_ERR_INVALID; }
  const size_t bytes = (size_t)(m->rows * m->wpr) * 4;
  m->alloc_bytes = ((bytes + 255) & ~(size_t)255) + 256;
  if (cudaMallocAsync((void**)&m->d, m->alloc_bytes, c->stream) != cudaSuccess) {
    cudaGetLastError();
    delete m;
    c->err = "cudaMallocAsync failed for matrix";
    return BIC_ERR_NOMEM;
  }
  cudaError_t e = cudaMemsetAsync(m->d, 0, m->alloc_bytes, c->stream);
  if (e != cudaSuccess) { cudaFreeAsync(m->d, c->stream); delete m; c->err = cudaGetErrorString(e); return BIC_ERR_CUDA; }
  *out = m;
  return BIC_OK;
}

extern "C" bic_status bic_mat_destroy(bic_ctx* c, bic_mat* m) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !m) return BIC_ERR_INVALID;
  if (m->pooled) {  // stream ordered: everything queued so far on this stream may still use it
    if (m->d) BIC_CUDA(c, cudaFreeAsync(m->d, c->stream));
    delete m;
    return BIC_OK;
  }
  BIC_CUDA(c, bic_wait_stream(c));
  if (m->owns && m->d) cudaFree(m->d);
  delete m;
  return BIC_OK;
}

extern "C" uint64_t bic_mat_rows(const bic_mat* m) { return m ? m->rows : 0; }
extern "C" uint64_t bic_mat_cols(const bic_mat* m) { return m ? m->cols : 0; }
extern "C" void* bic_mat_device_ptr(const bic_mat* m) { return m ? (void*)m->d : nullptr; }
extern "C" uint64_t bic_mat_stride_words32(const bic_mat* m) { return m ? m->wpr : 0; }

// host u64 words (MSB first) -> device u32 words: word q of a row is the high (q even) or low
// (q odd) half of host word q/2; the tail word is masked so stale pad bits never enter.
__global__ void k_words64_to_dev(const uint64_t* __restrict__ src, uint32_t* __restrict__ dst,
                                 uint64_t rows, uint64_t wpr32, uint64_t wpr64, uint32_t tmask) {
  const uint64_t total = rows * wpr32;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = i / wpr32, q = i - r * wpr32;
    const uint64_t w = src[r * wpr64 + (q >> 1)];
    uint32_t v = (q & 1) ? (uint32_t)w : (uint32_t)(w >> 32);
    if (q == wpr32 - 1) v &= tmask;
    dst[i] = v;
  }
}

__global__ void k_dev_to_words64(const uint32_t* __restrict__ src, uint64_t* __restrict__ dst,
                                 uint64_t rows, uint64_t wpr32, uint64_t wpr64) {
  const uint64_t total = rows * wpr64;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = i / wpr64, j = i - r * wpr64;
    const uint64_t hi = src[r * wpr32 + 2 * j];
    const uint64_t lo = (2 * j + 1 < wpr32) ? src[r * wpr32 + 2 * j + 1] : 0u;
    dst[i] = (hi << 32) | lo;
  }
}

// P4 payload (rows of ceil(cols/8) bytes) <-> device words
__global__ void k_pbm_to_dev(const uint8_t* __restrict__ src, uint32_t* __restrict__ dst,
                             uint64_t rows, uint64_t wpr32, uint64_t bpr, uint32_t tmask) {
  const uint64_t total = rows * wpr32;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = i / wpr32, q = i - r * wpr32;
    const uint8_t* p = src + r * bpr + q * 4;
    const uint64_t left = bpr - q * 4;
    uint32_t v = (uint32_t)p[0] << 24;
    if (left > 1) v |= (uint32_t)p[1] << 16;
    if (left > 2) v |= (uint32_t)p[2] << 8;
    if (left > 3) v |= (uint32_t)p[3];
    if (q == wpr32 - 1) v &= tmask;
    dst[i] = v;
  }
}

__global__ void k_dev_to_pbm(const uint32_t* __restrict__ src, uint8_t* __restrict__ dst,
                             uint64_t rows, uint64_t wpr32, uint64_t bpr) {
  const uint64_t total = rows * bpr;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = i / bpr, b = i - r * bpr;
    const uint32_t w = src[r * wpr32 + (b >> 2)];
    dst[i] = (uint8_t)(w >> (24 - 8 * (b & 3)));
  }
}

static int copy_grid(const bic_ctx* c, uint64_t items) { return bic_grid_for(c, items, 256, 16); }

extern "C" bic_status bic_mat_upload_words64(bic_ctx* c, bic_mat* m, const uint64_t* host) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !m || (!host && m->rows * m->cols)) return BIC_ERR_INVALID;
  if (m->rows * m->cols == 0) return BIC_OK;
  const uint64_t wpr64 = div_up_u64(m->cols, 64);
  const size_t bytes = (size_t)(m->rows * wpr64) * 8;
  BIC_TRY(bic_scratch_reserve(c, &c->staging, bytes));
  BIC_CUDA(c, cudaMemcpyAsync(c->staging.p, host, bytes, cudaMemcpyHostToDevice, c->stream));
  BIC_PROF(c, KID_WORDS64_TO_DEV);
  k_words64_to_dev<<<copy_grid(c, m->words()), 256, 0, c->stream>>>((const uint64_t*)c->staging.p, m->d, m->rows,
                                                                   m->wpr, wpr64, tail_mask32(m->cols));
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

extern "C" bic_status bic_mat_download_words64(bic_ctx* c, const bic_mat* m, uint64_t* host) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !m || (!host && m->rows * m->cols)) return BIC_ERR_INVALID;
  if (m->rows * m->cols == 0) return BIC_OK;
  const uint64_t wpr64 = div_up_u64(m->cols, 64);
  const size_t bytes = (size_t)(m->rows * wpr64) * 8;
  BIC_TRY(bic_scratch_reserve(c, &c->staging, bytes));
  BIC_PROF(c, KID_DEV_TO_WORDS64);
  k_dev_to_words64<<<copy_grid(c, m->rows * wpr64), 256, 0, c->stream>>>(m->d, (uint64_t*)c->staging.p, m->rows,
                                                                         m->wpr, wpr64);
  BIC_LAUNCH_CHECK(c);
  BIC_CUDA(c, cudaMemcpyAsync(host, c->staging.p, bytes, cudaMemcpyDeviceToHost, c->stream));
  BIC_CUDA(c, bic_wait_stream(c));
  return BIC_OK;
}

extern "C" bic_status bic_mat_upload_pbm(bic_ctx* c, bic_mat* m, const uint8_t* payload) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !m || (!payload && m->rows * m->cols)) return BIC_ERR_INVALID;
  if (m->rows * m->cols == 0) return BIC_OK;
  const uint64_t bpr = div_up_u64(m->cols, 8);
  const size_t bytes = (size_t)(m->rows * bpr);
  BIC_TRY(bic_scratch_reserve(c, &c->staging, bytes + 16));
  BIC_CUDA(c, cudaMemcpyAsync(c->staging.p, payload, bytes, cudaMemcpyHostToDevice, c->stream));
  BIC_PROF(c, KID_PBM_TO_DEV);
  k_pbm_to_dev<<<copy_grid(c, m->words()), 256, 0, c->stream>>>((const uint8_t*)c->staging.p, m->d, m->rows, m->wpr,
                                                               bpr, tail_mask32(m->cols));
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

extern "C" bic_status bic_mat_download_pbm(bic_ctx* c, const bic_mat* m, uint8_t* payload) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !m || (!payload && m->rows * m->cols)) return BIC_ERR_INVALID;
  if (m->rows * m->cols == 0) return BIC_OK;
  const uint64_t bpr = div_up_u64(m->cols, 8);
  const size_t bytes = (size_t)(m->rows * bpr);
  BIC_TRY(bic_scratch_reserve(c, &c->staging, bytes + 16));
  BIC_PROF(c, KID_DEV_TO_PBM);
  k_dev_to_pbm<<<copy_grid(c, bytes), 256, 0, c->stream>>>(m->d, (uint8_t*)c->staging.p, m->rows, m->wpr, bpr);
  BIC_LAUNCH_CHECK(c);
  BIC_CUDA(c, cudaMemcpyAsync(payload, c->staging.p, bytes, cudaMemcpyDeviceToHost, c->stream));
  BIC_CUDA(c, bic_wait_stream(c));
  return BIC_OK;
}

extern "C" bic_status bic_mat_clear(bic_ctx* c, bic_mat* m) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !m) return BIC_ERR_INVALID;
  BIC_CUDA(c, cudaMemsetAsync(m->d, 0, m->alloc_bytes, c->stream));
  return BIC_OK;
}

extern "C" bic_status bic_mat_copy(bic_ctx* c, const bic_mat* src, bic_mat* dst) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !src || !dst || src->rows != dst->rows || src->cols != dst->cols) return BIC_ERR_INVALID;
  BIC_CUDA(c, cudaMemcpyAsync(dst->d, src->d, (size_t)src->words() * 4, cudaMemcpyDeviceToDevice, c->stream));
  return BIC_OK;
}

// rows [src_row0, src_row0 + nrows) of src -> rows [dst_row0, ...) of dst (same number of columns): the row-wise half of
// set_submatrix / copy_submatrix_to (src/binmat.cpp:267-298, 373-414) that stacking patch matrices needs
extern "C" bic_status bic_mat_copy_rows(bic_ctx* c, const bic_mat* src, uint64_t src_row0, uint64_t nrows, bic_mat* dst,
                                        uint64_t dst_row0) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !src || !dst || src->cols != dst->cols || src_row0 + nrows > src->rows || dst_row0 + nrows > dst->rows)
    return BIC_ERR_INVALID;
  if (nrows == 0) return BIC_OK;
  BIC_CUDA(c, cudaMemcpyAsync(dst->d + dst_row0 * dst->wpr, src->d + src_row0 * src->wpr, (size_t)nrows * src->wpr * 4,
                              cudaMemcpyDeviceToDevice, c->stream));
  return BIC_OK;
}

// popcount reductions: weight(A) and dist(A,B) = weight(A xor B)
__global__ void k_weight(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint64_t nwords,
                         unsigned long long* out) {
  unsigned long long acc = 0;
  const uint64_t nv = nwords >> 2;
  const uint4* a4 = reinterpret_cast<const uint4*>(a);
  const uint4* b4 = reinterpret_cast<const uint4*>(b);
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nv; i += (uint64_t)gridDim.x * blockDim.x) {
    uint4 v = a4[i];
    if (b) { const uint4 u = b4[i]; v.x ^= u.x; v.y ^= u.y; v.z ^= u.z; v.w ^= u.w; }
    acc += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (nwords & 3)) {
    const uint64_t i = (nv << 2) + threadIdx.x;
    acc += __popc(b ? (a[i] ^ b[i]) : a[i]);
  }
  acc = warp_sum_u64(acc);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

static bic_status weight_impl(bic_ctx* c, const bic_mat* a, const bic_mat* b, uint64_t* w) {
  BIC_CUDA(c, cudaMemsetAsync(c->d_scalars, 0, sizeof(uint64_t), c->stream));
  if (a->words()) {
    BIC_PROF(c, KID_WEIGHT);
    k_weight<<<bic_grid_for(c, a->words() / 4 + 1, 256, 8), 256, 0, c->stream>>>(
        a->d, b ? b->d : nullptr, a->words(), (unsigned long long*)c->d_scalars);
    BIC_LAUNCH_CHECK(c);
  }
  BIC_TRY(bic_read_scalars(c, 1));
  *w = c->h_scalars[0];
  return BIC_OK;
}

extern "C" bic_status bic_mat_weight(bic_ctx* c, const bic_mat* m, uint64_t* w) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !m || !w) return BIC_ERR_INVALID;
  return weight_impl(c, m, nullptr, w);
}

extern "C" bic_status bic_mat_dist(bic_ctx* c, const bic_mat* a, const bic_mat* b, uint64_t* d) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !a || !b || !d || a->rows != b->rows || a->cols != b->cols) return BIC_ERR_INVALID;
  return weight_impl(c, a, b, d);
}

__global__ void k_xor(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t* __restrict__ o, uint64_t nv) {
  const uint4* a4 = reinterpret_cast<const uint4*>(a);
  const uint4* b4 = reinterpret_cast<const uint4*>(b);
  uint4* o4 = reinterpret_cast<uint4*>(o);
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nv; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 x = a4[i], y = b4[i];
    o4[i] = make_uint4(x.x ^ y.x, x.y ^ y.y, x.z ^ y.z, x.w ^ y.w);
  }
}

extern "C" bic_status bic_mat_xor(bic_ctx* c, const bic_mat* a, const bic_mat* b, bic_mat* o) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !a || !b || !o || a->rows != b->rows || a->cols != b->cols || a->rows != o->rows || a->cols != o->cols)
    return BIC_ERR_INVALID;
  // allocations are padded to 16 B multiples with zero words, so whole uint4s are safe
  const uint64_t nv = div_up_u64(a->words(), 4);
  if (nv) {
    BIC_PROF(c, KID_XOR);
    k_xor<<<bic_grid_for(c, nv, 256, 8), 256, 0, c->stream>>>(a->d, b->d, o->d, nv);
    BIC_LAUNCH_CHECK(c);
  }
  return BIC_OK;
}

// ---------------------------------------------------------------------------------------------
// GSL rand48 / gsl_rng_uniform_int as used at src/bsvd.cpp:8-15, :241 (published GSL algorithm:
// x <- 0x5DEECE66D x + 0xB mod 2^48; seed s -> (0x330E | (s & 0xFFFFFFFF) << 16); output = top
// 32 bits; uniform_int: scale = 0xFFFFFFFF / n, redraw while get()/scale >= n).
// ---------------------------------------------------------------------------------------------
extern "C" void bic_rand48_seed(uint64_t* state, unsigned long seed) {
  if (seed == 0) *state = ((uint64_t)0x1234 << 32) | ((uint64_t)0xABCD << 16) | 0x330E;
  else *state = (((uint64_t)seed & 0xFFFFFFFFull) << 16) | 0x330E;
}

extern "C" uint64_t bic_rand48_uniform_int(uint64_t* state, uint64_t n) {
  const uint64_t range = 0xFFFFFFFFull;
  if (n == 0 || n > range) return 0;
  const uint64_t scale = range / n;
  uint64_t k;
  do {
    *state = (*state * 0x5DEECE66Dull + 0xBull) & 0xFFFFFFFFFFFFull;
    k = (*state >> 16) / scale;
  } while (k >= n);
  return k;
}

// ---------------------------------------------------------------------------------------------
// per-launch device timers: an event pair around every kernel of this context's stream
// ---------------------------------------------------------------------------------------------
static const char* const kKernelNames[KID_COUNT] = {
  "k_words64_to_dev", "k_dev_to_words64", "k_pbm_to_dev", "k_dev_to_pbm", "k_weight", "k_xor",
  "k_extract", "k_assemble", "k_row_nonzero", "k_gather_rows", "k_col_hist", "k_pivot_usage", "k_init_finalize",
  "k_update_coefficients", "k_residual", "k_transpose_bits", "k_update_dictionary", "k_dict_hist_popc", "k_dict_resolve", "k_dict_scan",
  "k_compact_rows", "k_expand_rows", "k_gol_tile_counts", "k_gol_scan_tiles_a", "k_gol_walk<0>", "k_gol_scan_tiles_b",
  "k_gol_walk<1>", "k_gol_decode", "k_first_one/zero", "k_fill_ones", "k_eg_encode", "k_eg_decode", "k_dict_chain", "k_dict_apply",
  "k_dict_compact", "k_dict_bucket", "k_bitplanes", "k_proximus", "k_match", "k_match_decide"};

static cudaEvent_t prof_event(bic_ctx* c) {
  if (!c->prof_free.empty()) { cudaEvent_t e = c->prof_free.back(); c->prof_free.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

void bic_prof_begin(bic_ctx* c, int kid) {
  bic_prof_rec r;
  r.kid = kid;
  r.e0 = prof_event(c);
  r.e1 = nullptr;
  cudaEventRecord(r.e0, c->stream);
  c->prof_recs.push_back(r);
}

void bic_prof_end(bic_ctx* c) {
  if (c->prof_recs.empty() || c->prof_recs.back().e1) return;
  bic_prof_rec& r = c->prof_recs.back();
  r.e1 = prof_event(c);
  cudaEventRecord(r.e1, c->stream);
}

static void prof_collect(bic_ctx* c) {
  cudaStreamSynchronize(c->stream);
  for (auto& r : c->prof_recs) {
    if (r.e1) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { c->prof_ms[r.kid] += ms; c->prof_n[r.kid]++; }
      c->prof_free.push_back(r.e1);
    }
    c->prof_free.push_back(r.e0);
  }
  c->prof_recs.clear();
}

extern "C" bic_status bic_prof_enable(bic_ctx* c, int on) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c) return BIC_ERR_INVALID;
  if (!on && c->prof_on) prof_collect(c);
  c->prof_on = on != 0;
  return BIC_OK;
}

extern "C" bic_status bic_prof_reset(bic_ctx* c) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c) return BIC_ERR_INVALID;
  prof_collect(c);
  for (int i = 0; i < KID_COUNT; ++i) { c->prof_ms[i] = 0; c->prof_n[i] = 0; }
  return BIC_OK;
}

extern "C" int bic_prof_kernel_count(void) { return KID_COUNT; }

extern "C" bic_status bic_prof_get(bic_ctx* c, int kid, const char** name, uint64_t* launches, double* total_ms) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || kid < 0 || kid >= KID_COUNT) return BIC_ERR_INVALID;
  prof_collect(c);
  if (name) *name = kKernelNames[kid];
  if (launches) *launches = c->prof_n[kid];
  if (total_ms) *total_ms = c->prof_ms[kid];
  return BIC_OK;
}

extern "C" bic_status bic_ctx_set_option(bic_ctx* c, const char* name, int64_t value) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !name) return BIC_ERR_INVALID;
  if (!strcmp(name, "wait_mode")) { if (value < 0 || value > 2) return BIC_ERR_INVALID; c->wait_mode = (int)value; return BIC_OK; }
  if (!strcmp(name, "gol_algo")) { if (value < 1 || value > 2) return BIC_ERR_INVALID; c->gol_algo = (int)value; return BIC_OK; }
  if (!strcmp(name, "gol_scan")) { if (value < 0 || value > 2) return BIC_ERR_INVALID; c->gol_scan = (int)value; return BIC_OK; }
  if (!strcmp(name, "gol_list")) { if (value < 0 || value > 2) return BIC_ERR_INVALID; c->gol_list = (int)value; return BIC_OK; }
  if (!strcmp(name, "gol_presize_pct")) { if (value < 1 || value > 1000) return BIC_ERR_INVALID; c->gol_presize_pct = (int)value; return BIC_OK; }
  if (!strcmp(name, "gol_onepass")) { c->gol_onepass = value != 0; return BIC_OK; }
  if (!strcmp(name, "dict_algo")) { if (value < 0 || value > 2) return BIC_ERR_INVALID; c->dict_algo = (int)value; return BIC_OK; }
  if (!strcmp(name, "dict_update")) { if (value < 0 || value > 1) return BIC_ERR_INVALID; c->dict_update = (int)value; return BIC_OK; }
  if (!strcmp(name, "coef_algo")) { if (value < 0 || value > 1) return BIC_ERR_INVALID; c->coef_algo = (int)value; return BIC_OK; }
  if (!strcmp(name, "chain_bucket_cap")) { c->chain_bucket_cap = value; return BIC_OK; }
  if (!strcmp(name, "chain_cluster")) {
    if (value != 1 && value != 2 && value != 4 && value != 8 && value != 16) return BIC_ERR_INVALID;
    c->chain_cluster = (int)value;
    return BIC_OK;
  }
  return bic_fail(c, BIC_ERR_INVALID, "unknown option");
}

// Order `waiter`'s stream after everything queued so far on `signal`'s stream (cross-context
// dependency for callers that drive several contexts at once).
extern "C" bic_status bic_ctx_wait_ctx(bic_ctx* waiter, bic_ctx* signal) {
  if (!waiter || !signal) return BIC_ERR_INVALID;
  cudaSetDevice(signal->device);
  cudaEvent_t e = nullptr;
  BIC_CUDA(waiter, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  BIC_CUDA(waiter, cudaEventRecord(e, signal->stream));
  BIC_CUDA(waiter, cudaStreamWaitEvent(waiter->stream, e, 0));
  BIC_CUDA(waiter, cudaEventDestroy(e));
  return BIC_OK;
}
