// learn_model_traditional and the whole-page encoder / decoder.
// Reference: learn_model_traditional src/bsvd.cpp:1215-1244; driver src/bsvd_test.cpp:56-155.
#include "bic_internal.cuh"

#include <new>
#include <vector>

bic_status bic_k_update_coefficients(bic_ctx* c, bic_mat* E, const bic_mat* D, bic_mat* A, unsigned long long* d_changed);
bic_status bic_k_update_dictionary(bic_ctx* c, bic_mat* E, bic_mat* D, const bic_mat* A, unsigned long long* d_changed);

extern "C" bic_status bic_learn_model_traditional(bic_ctx* c, const bic_mat* X, bic_mat* E, bic_mat* D, bic_mat* A,
                                                  uint64_t* iterations, uint64_t* trace, uint64_t trace_cap) {
  BIC_RANGE("bic:learn_model_traditional");
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !X || !E || !D || !A) return BIC_ERR_INVALID;
  BIC_TRY(bic_residual(c, X, A, D, E));  // mul(A,false,D,false,E); add(E,X,E)  src/bsvd.cpp:1219-1220
  uint64_t changed = 1, iter = 0;
  if (c->dict_update == 1) {
    // update_dictionary points at update_dictionary_proximus (-d 1): the reference's loop over its plug points, :1227-1243
    while (changed > 0) {
      iter++;
      uint64_t cc = 0, ca = 0;
      BIC_TRY(bic_update_coefficients(c, E, D, A, &cc));
      BIC_TRY(bic_update_dictionary_proximus(c, E, D, A, &ca));
      changed = cc + ca;
      if (trace && iter <= trace_cap) { trace[2 * (iter - 1)] = cc; trace[2 * (iter - 1) + 1] = ca; }
    }
    if (iterations) *iterations = iter;
    return BIC_OK;
  }
  unsigned long long* d_cc = (unsigned long long*)c->d_scalars;
  while (changed > 0) {  // :1227
    iter++;
    BIC_CUDA(c, cudaMemsetAsync(c->d_scalars, 0, 2 * sizeof(uint64_t), c->stream));
    BIC_TRY(bic_k_update_coefficients(c, E, D, A, d_cc));      // :1229
    BIC_TRY(bic_k_update_dictionary(c, E, D, A, d_cc + 1));    // :1235
    BIC_TRY(bic_read_scalars(c, 2));                           // the loop condition needs the counts
    changed = c->h_scalars[0] + c->h_scalars[1];
    if (trace && iter <= trace_cap) { trace[2 * (iter - 1)] = c->h_scalars[0]; trace[2 * (iter - 1) + 1] = c->h_scalars[1]; }
    if (changed > 0 && c->h_scalars[1] == 0) {
      // No atom changed, so D and E are exactly what the coefficient update left behind, and that update
      // ended every row with a pass that found no improving atom. The reference's next iteration therefore
      // changes no row, sees the same (E, A, D) in its dictionary update as this one did, changes no atom,
      // and ends the loop: it is counted (and traced as 0, 0) without being run.
      iter++;
      if (trace && iter <= trace_cap) { trace[2 * (iter - 1)] = 0; trace[2 * (iter - 1) + 1] = 0; }
      break;
    }
  }
  if (iterations) *iterations = iter;
  return BIC_OK;
}

// container layout: bic_internal.cuh
static const uint64_t HDR_FIELDS = BIC_HDR_FIELDS, STREAM_FIELDS = BIC_STREAM_FIELDS;

struct EncWorkspace {
  uint64_t rows = 0, cols = 0, W = 0, K = 0;
  bic_mat *raster = nullptr, *X = nullptr, *E = nullptr, *D = nullptr, *A = nullptr;
  bic_stream* st[3] = {nullptr, nullptr, nullptr};
};

static EncWorkspace* ws_of(bic_ctx* c);

// one workspace per context, kept in a side table so bic_ctx stays a plain struct
#include <map>
#include <mutex>
static std::map<bic_ctx*, EncWorkspace> g_ws;
static std::mutex g_ws_mu;
static EncWorkspace* ws_of(bic_ctx* c) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  return &g_ws[c];
}

static void ws_release(bic_ctx* c, EncWorkspace* w) {
  bic_mat** ms[] = {&w->raster, &w->X, &w->E, &w->D, &w->A};
  for (auto pm : ms) if (*pm) { bic_mat_destroy(c, *pm); *pm = nullptr; }
}

extern "C" void bic_internal_drop_workspace(bic_ctx* c) {
  EncWorkspace* w = ws_of(c);
  ws_release(c, w);
  for (auto& s : w->st) if (s) { bic_stream_destroy(c, s); s = nullptr; }
  std::lock_guard<std::mutex> lk(g_ws_mu);
  g_ws.erase(c);
}

static bic_status ws_prepare(bic_ctx* c, EncWorkspace* w, uint64_t rows, uint64_t cols, uint64_t W, uint64_t K) {
  if (w->raster && w->rows == rows && w->cols == cols && w->W == W && w->K == K) return BIC_OK;
  ws_release(c, w);
  const uint64_t Ny = (W - 1 + rows) / W, Nx = (W - 1 + cols) / W, n = Nx * Ny, m = W * W;
  BIC_TRY(bic_mat_create(c, rows, cols, &w->raster));
  BIC_TRY(bic_mat_create(c, n, m, &w->X));
  BIC_TRY(bic_mat_create(c, n, m, &w->E));
  BIC_TRY(bic_mat_create(c, K, m, &w->D));
  BIC_TRY(bic_mat_create(c, n, K, &w->A));
  for (auto& s : w->st) if (!s) BIC_TRY(bic_stream_create(c, &s));
  w->rows = rows; w->cols = cols; w->W = W; w->K = K;
  return BIC_OK;
}

static bic_status encode_from_raster(bic_ctx* c, EncWorkspace* w, const bic_mat* raster, uint64_t W, uint64_t K, unsigned long seed,
                                     uint8_t* out, uint64_t cap_bytes, bic_encode_info* info);

extern "C" bic_status bic_encode_raster(bic_ctx* c, const uint8_t* pbm_payload, uint64_t rows, uint64_t cols, uint64_t W,
                                        uint64_t K, unsigned long seed, uint8_t* out, uint64_t cap_bytes,
                                        bic_encode_info* info) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !pbm_payload || W == 0 || K == 0 || rows == 0 || cols == 0) return BIC_ERR_INVALID;
  if (out && ((uintptr_t)out & 7) != 0) return bic_fail(c, BIC_ERR_INVALID, "encode_raster: out must be 8-byte aligned");   // before the fit, not after
  EncWorkspace* w = ws_of(c);
  BIC_TRY(ws_prepare(c, w, rows, cols, W, K));
  BIC_TRY(bic_mat_upload_pbm(c, w->raster, pbm_payload));
  return encode_from_raster(c, w, w->raster, W, K, seed, out, cap_bytes, info);
}

// the same encoder for a raster that already sits in device memory (e.g. one of bic_split_bitplanes' planes)
extern "C" bic_status bic_encode_raster_resident(bic_ctx* c, const bic_mat* raster, uint64_t W, uint64_t K, unsigned long seed,
                                                 uint8_t* out, uint64_t cap_bytes, bic_encode_info* info) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !raster || W == 0 || K == 0 || raster->rows == 0 || raster->cols == 0) return BIC_ERR_INVALID;
  if (out && ((uintptr_t)out & 7) != 0) return bic_fail(c, BIC_ERR_INVALID, "encode_raster: out must be 8-byte aligned");
  EncWorkspace* w = ws_of(c);
  BIC_TRY(ws_prepare(c, w, raster->rows, raster->cols, W, K));
  return encode_from_raster(c, w, raster, W, K, seed, out, cap_bytes, info);
}

static bic_status encode_from_raster(bic_ctx* c, EncWorkspace* w, const bic_mat* raster, uint64_t W, uint64_t K, unsigned long seed,
                                     uint8_t* out, uint64_t cap_bytes, bic_encode_info* info) {
  const uint64_t rows = raster->rows, cols = raster->cols;
  BIC_TRY(bic_extract_patches(c, raster, W, w->X));          // src/bsvd_test.cpp:80-99
  uint64_t rng;
  bic_rand48_seed(&rng, seed);                               // -r / random_seed, src/bsvd.cpp:12
  BIC_TRY(bic_initialize_model_neighbor(c, w->X, w->D, w->A, &rng));  // :114
  uint64_t iters = 0;
  BIC_TRY(bic_learn_model_traditional(c, w->X, w->E, w->D, w->A, &iters, nullptr, 0));  // :119
  const bic_mat* mats[3] = {w->D, w->A, w->E};
  for (int i = 0; i < 3; ++i) BIC_TRY(bic_golomb_encode(c, mats[i], 256, w->st[i]));
  // lay out the container
  uint64_t need = (HDR_FIELDS + 3 * STREAM_FIELDS) * 8;
  for (int i = 0; i < 3; ++i) need += div_up_u64(div_up_u64(w->st[i]->info.bitcount, 8), 8) * 8 + w->st[i]->info.nchunks * 16;
  if (info) {
    memset(info, 0, sizeof(*info));
    info->rows = rows; info->cols = cols; info->W = W; info->K = K;
    info->n = w->X->rows; info->m = w->X->cols; info->iterations = iters;
    info->bits_D = w->st[0]->info.bitcount; info->bits_A = w->st[1]->info.bitcount; info->bits_E = w->st[2]->info.bitcount;
    info->weight_D = w->st[0]->info.nsamples - 1; info->weight_A = w->st[1]->info.nsamples - 1;
    info->weight_E = w->st[2]->info.nsamples - 1;  // samples = ones + 1
    info->container_bytes = need;
  }
  if (!out) return BIC_OK;
  if (cap_bytes < need) return BIC_ERR_CAPACITY;
  if (((uintptr_t)out & 7) != 0) return bic_fail(c, BIC_ERR_INVALID, "encode_raster: out must be 8-byte aligned");
  uint64_t* h = (uint64_t*)out;
  h[0] = BIC_MAGIC; h[1] = 1; h[2] = rows; h[3] = cols; h[4] = W; h[5] = K; h[6] = w->X->rows; h[7] = w->X->cols;
  h[8] = iters; h[9] = (uint64_t)seed;
  uint64_t off = (HDR_FIELDS + 3 * STREAM_FIELDS) * 8;
  for (int i = 0; i < 3; ++i) {
    const bic_stream_info& si = w->st[i]->info;
    uint64_t* f = h + HDR_FIELDS + i * STREAM_FIELDS;
    f[0] = si.coder; f[1] = si.chunk_samples; f[2] = si.rows; f[3] = si.cols; f[4] = si.bitcount; f[5] = si.nsamples; f[6] = si.nchunks;
    const uint64_t nb = div_up_u64(si.bitcount, 8), nbp = div_up_u64(nb, 8) * 8;
    memset(out + off + nb, 0, nbp - nb);
    BIC_TRY(bic_stream_download(c, w->st[i], out + off, nb, (uint64_t*)(out + off + nbp), si.nchunks));
    off += nbp + si.nchunks * 16;
  }
  return BIC_OK;
}

static bic_status decode_raster_checked(bic_ctx* c, const uint8_t* cont, uint64_t nbytes, uint8_t* pbm_payload,
                                        uint64_t cap_bytes, uint64_t* rows_out, uint64_t* cols_out) {
  const uint64_t hdr = (HDR_FIELDS + 3 * STREAM_FIELDS) * 8;
  if (nbytes < hdr) return bic_fail(c, BIC_ERR_CORRUPT, "container: truncated header");
  std::vector<uint64_t> h((HDR_FIELDS + 3 * STREAM_FIELDS));
  memcpy(h.data(), cont, hdr);
  if (h[0] != BIC_MAGIC || h[1] != 1) return bic_fail(c, BIC_ERR_CORRUPT, "container: bad magic/version");
  const uint64_t rows = h[2], cols = h[3], W = h[4], K = h[5], n = h[6], m = h[7];
  // Every field is bounded BEFORE it enters any arithmetic: the file controls them all. A raster cannot have more pixels
  // than 64 x the container's bits can possibly describe... there is no such bound for a compressor, so the limits are
  // absolute: sides < 2^31, W <= 2^12, 1 <= K <= 65535 (the coefficient kernels' key packs the atom index in 16 bits).
  const uint64_t SIDE_MAX = 1ull << 31, W_MAX = 1ull << 12;
  if (W == 0 || W > W_MAX || rows == 0 || cols == 0 || rows >= SIDE_MAX || cols >= SIDE_MAX || K == 0 || K > 65535)
    return bic_fail(c, BIC_ERR_CORRUPT, "container: field out of range");
  if (m != W * W || n != ((W - 1 + rows) / W) * ((W - 1 + cols) / W))   // both products < 2^62
    return bic_fail(c, BIC_ERR_CORRUPT, "container: inconsistent shape");
  if (!bic_shape_ok(n, m) || !bic_shape_ok(n, K) || !bic_shape_ok(rows, cols))
    return bic_fail(c, BIC_ERR_CORRUPT, "container: shape too large");
  if (rows_out) *rows_out = rows;
  if (cols_out) *cols_out = cols;
  const uint64_t payload = rows * div_up_u64(cols, 8);
  if (!pbm_payload) return BIC_OK;
  if (cap_bytes < payload) return BIC_ERR_CAPACITY;
  // validate all three stream headers against the bytes that are really there before touching the device
  const uint64_t shp[3][2] = {{K, m}, {n, K}, {n, m}};
  bic_stream_info sis[3];
  uint64_t offs[3];
  uint64_t off = hdr;
  for (int i = 0; i < 3; ++i) {
    const uint64_t* f = h.data() + HDR_FIELDS + i * STREAM_FIELDS;
    bic_stream_info& si = sis[i];
    if (f[0] != BIC_CODER_GOLOMB && f[0] != BIC_CODER_EG) return bic_fail(c, BIC_ERR_CORRUPT, "container: unknown coder");
    si.coder = (uint32_t)f[0]; si.rows = f[2]; si.cols = f[3];
    si.bitcount = f[4]; si.nsamples = f[5]; si.nchunks = f[6];
    if (si.rows != shp[i][0] || si.cols != shp[i][1]) return bic_fail(c, BIC_ERR_CORRUPT, "container: stream shape mismatch");
    const uint64_t left = nbytes - off, N = si.rows * si.cols;
    if (si.bitcount > 8 * left) return bic_fail(c, BIC_ERR_CORRUPT, "container: truncated stream");   // left < 2^61
    const uint64_t nb = div_up_u64(si.bitcount, 8), nbp = div_up_u64(nb, 8) * 8;
    if (nbp > left || si.nchunks > (left - nbp) / 16) return bic_fail(c, BIC_ERR_CORRUPT, "container: truncated stream");
    if (si.coder == BIC_CODER_GOLOMB) {
      if (f[1] == 0 || f[1] > (1ull << 30) || (f[1] & (f[1] - 1))) return bic_fail(c, BIC_ERR_CORRUPT, "container: bad chunk size");
      si.chunk_samples = (uint32_t)f[1];
      // popcount + 1 samples, each at least one code bit
      if (si.nsamples == 0 || si.nsamples > N + 1 || si.nsamples > si.bitcount ||
          si.nchunks != div_up_u64(si.nsamples, si.chunk_samples))
        return bic_fail(c, BIC_ERR_CORRUPT, "container: inconsistent stream header");
    } else {
      si.chunk_samples = 0;
      if (si.nchunks != 0 || si.nsamples != 0 || (si.bitcount != N + si.rows && si.bitcount != N + si.rows + 1))
        return bic_fail(c, BIC_ERR_CORRUPT, "container: inconsistent stream header");
    }
    offs[i] = off;
    off += nbp + si.nchunks * 16;
  }
  EncWorkspace* w = ws_of(c);
  BIC_TRY(ws_prepare(c, w, rows, cols, W, K));
  bic_mat* mats[3] = {w->D, w->A, w->E};
  for (int i = 0; i < 3; ++i) {
    const bic_stream_info& si = sis[i];
    const uint64_t nb = div_up_u64(si.bitcount, 8), nbp = div_up_u64(nb, 8) * 8;
    std::vector<uint64_t> idx(si.nchunks * 2 + 1);
    memcpy(idx.data(), cont + offs[i] + nbp, si.nchunks * 16);
    // chunk-index entries are offsets into the code and into the matrix: bound them here, the kernel re-checks
    for (uint64_t j = 0; j < si.nchunks; ++j)
      if (idx[2 * j] >= si.bitcount || idx[2 * j + 1] > si.rows * si.cols)
        return bic_fail(c, BIC_ERR_CORRUPT, "container: chunk index out of range");
    BIC_TRY(bic_stream_upload(c, w->st[i], &si, cont + offs[i], idx.data()));
    BIC_CUDA(c, bic_wait_stream(c));  // idx is a local
    if (si.coder == BIC_CODER_GOLOMB) BIC_TRY(bic_golomb_decode(c, w->st[i], mats[i]));
    else BIC_TRY(bic_eg_decode(c, w->st[i], mats[i]));
  }
  // X = A*D xor E  (the residual identity read backwards), then patches -> raster
  BIC_TRY(bic_residual(c, w->E, w->A, w->D, w->X));
  BIC_TRY(bic_assemble_patches(c, w->X, W, w->raster));
  BIC_TRY(bic_mat_download_pbm(c, w->raster, pbm_payload));
  return BIC_OK;
}

extern "C" bic_status bic_decode_raster(bic_ctx* c, const uint8_t* cont, uint64_t nbytes, uint8_t* pbm_payload,
                                        uint64_t cap_bytes, uint64_t* rows_out, uint64_t* cols_out) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !cont) return BIC_ERR_INVALID;
  try {  // nothing may unwind through the C boundary
    return decode_raster_checked(c, cont, nbytes, pbm_payload, cap_bytes, rows_out, cols_out);
  } catch (const std::bad_alloc&) {
    return bic_fail(c, BIC_ERR_NOMEM, "decode_raster: host allocation failed");
  } catch (...) {
    return bic_fail(c, BIC_ERR_CORRUPT, "decode_raster: container rejected");
  }
}
