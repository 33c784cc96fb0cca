// Sliding-window template matching of the compress*_test experiments: for every W x W patch of a binary
// raster, the earlier position of the image whose W x W window is closest in Hamming distance, then the reference's
// enumerative + Golomb costing of "code the patch as a difference to that window" against "code the patch itself".
//
//   v1  src/compress_test.cpp:73-141    every patch searches ALL earlier positions of the unmodified image; the first smallest
//                                       distance in scan order wins (the reference stops at distance 0, which picks the same one).
//                                       Patches are independent: one launch for all of them.
//   v4  src/compress4_test.cpp:89-171   window of radius R behind / above the patch, scanned backwards, stop at the first
//                                       distance <= T; a matched patch is REPLACED by its residual in the image, so later patches
//                                       search the coded image: the patches form one serial chain (two small launches each).
//
// A candidate window is W rows of W bits taken at an arbitrary column: two adjacent 32-bit words and a funnel shift per row,
// XOR with the patch row, POPC. A thread takes one candidate; neighbouring threads take neighbouring columns, so the words are
// shared through L1. The argmin with the reference's tie-break is a plain min over keys (distance << 40 | scan index); "stop at
// the first distance <= T" is a second min over (scan index << 16 | distance) restricted to those candidates.
// The code lengths are the reference's double arithmetic over integers: enumL (lgamma, src/compress_test.cpp:37-40) is
// tabulated on the host for every weight, the device only adds and truncates; the two GolombCoder bit counts are the serial
// counter over the per-patch samples (src/GolombCoder.cpp:13-34), run on the host over the downloaded records.
#include "bic_internal.cuh"

#include <math.h>

#include <vector>

struct MatchParams {
  const uint32_t* I;
  uint64_t rows, cols, wpr, S64;   // S64: bits per row of the reference's flat block array, 64 * ceil(cols / 64)
  uint32_t W, mode;                // mode 1: v1, 4: v4
  uint64_t T, R;
  uint64_t Nx;
};

// 32 bits of row r starting at column j; columns past the stored words and rows past the image read as zero
__device__ __forceinline__ uint32_t row_bits32(const MatchParams& P, uint64_t r, uint64_t j) {
  if (r >= P.rows) return 0u;
  const uint64_t wi = j >> 5;
  const unsigned off = (unsigned)(j & 31);
  const uint32_t* row = P.I + r * P.wpr;
  const uint32_t hi = wi < P.wpr ? __ldg(row + wi) : 0u;
  const uint32_t lo = (off && wi + 1 < P.wpr) ? __ldg(row + wi + 1) : 0u;
  return __funnelshift_l(lo, hi, off);
}

// the W bits get_submatrix delivers for (row r, column j), left aligned (src/binmat.cpp:267-298: the matrix is one flat array
// of 64-bit blocks, so a window that runs past the last block of a row continues in the next row)
__device__ __forceinline__ uint32_t window_bits(const MatchParams& P, uint64_t r, uint64_t j) {
  uint32_t v = row_bits32(P, r, j);
  if (j + P.W > P.S64) {
    const unsigned k = (unsigned)(P.S64 - j);       // 1 .. W-1 bits are left in this row
    v = (v & ~(0xFFFFFFFFu >> k)) | (row_bits32(P, r + 1, 0) >> k);
  }
  return v & (P.W >= 32 ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> P.W));
}

// the two candidate rectangles of a patch and their scan order
struct Rects {
  long long ilo[2], ihi[2], jlo[2], jhi[2];
  unsigned long long n[2];
  bool desc;
};
__device__ __forceinline__ Rects patch_rects(const MatchParams& P, long long i0, long long j0) {
  Rects q;
  const long long W = P.W;
  if (P.mode == 1) {               // compress_test.cpp:81-111, ascending
    q.desc = false;
    q.ilo[0] = 0; q.ihi[0] = i0 - W; q.jlo[0] = 0; q.jhi[0] = (long long)P.cols - 1;
    q.ilo[1] = i0 - W + 1 > 0 ? i0 - W + 1 : 0; q.ihi[1] = i0; q.jlo[1] = 0; q.jhi[1] = j0 - W;
  } else {                         // compress4_test.cpp:97-135, descending
    q.desc = true;
    const long long R = (long long)P.R;
    const long long mini = i0 > R ? i0 - R : 0, mini2 = i0 > W ? i0 - W : 0, minj = j0 > R ? j0 - R : 0;
    const long long maxj = (j0 + R) > ((long long)P.cols - W) ? (long long)P.cols - W : j0 + R;
    q.ilo[0] = mini2; q.ihi[0] = i0; q.jlo[0] = minj; q.jhi[0] = j0 - W;
    q.ilo[1] = mini; q.ihi[1] = i0 - W; q.jlo[1] = minj; q.jhi[1] = maxj;
  }
  for (int s = 0; s < 2; ++s)
    q.n[s] = (q.ihi[s] >= q.ilo[s] && q.jhi[s] >= q.jlo[s]) ? (unsigned long long)(q.ihi[s] - q.ilo[s] + 1) * (unsigned long long)(q.jhi[s] - q.jlo[s] + 1) : 0ull;
  return q;
}
__device__ __forceinline__ void scan_to_pos(const Rects& q, unsigned long long idx, long long* i2, long long* j2) {
  const int s = idx < q.n[0] ? 0 : 1;
  const unsigned long long c = s ? idx - q.n[0] : idx;
  const unsigned long long wj = (unsigned long long)(q.jhi[s] - q.jlo[s] + 1);
  const long long a = (long long)(c / wj), b = (long long)(c % wj);
  *i2 = q.desc ? q.ihi[s] - a : q.ilo[s] + a;
  *j2 = q.desc ? q.jhi[s] - b : q.jlo[s] + b;
}

// best1[patch] = min (distance << 40 | scan index); best2[patch] = min (scan index << 16 | distance) over distance <= T
__global__ void __launch_bounds__(256) k_match(MatchParams P, uint64_t li0, unsigned long long* __restrict__ best1,
                                               unsigned long long* __restrict__ best2) {
  __shared__ uint32_t sP[32];
  __shared__ unsigned long long s1[8], s2[8];
  const uint64_t li = li0 + blockIdx.y;
  const long long i0 = (long long)((li / P.Nx) * P.W), j0 = (long long)((li % P.Nx) * P.W);
  if (threadIdx.x < P.W) sP[threadIdx.x] = window_bits(P, (uint64_t)i0 + threadIdx.x, (uint64_t)j0);
  __syncthreads();
  const Rects q = patch_rects(P, i0, j0);
  const unsigned long long total = q.n[0] + q.n[1];
  unsigned long long k1 = ~0ull, k2 = ~0ull;
  for (unsigned long long c = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; c < total; c += (unsigned long long)gridDim.x * blockDim.x) {
    long long i2, j2;
    scan_to_pos(q, c, &i2, &j2);
    uint32_t d = 0;
    for (uint32_t di = 0; di < P.W; ++di) d += __popc(sP[di] ^ window_bits(P, (uint64_t)i2 + di, (uint64_t)j2));
    const unsigned long long a = ((unsigned long long)d << 40) | c;
    k1 = a < k1 ? a : k1;
    if (P.mode == 4 && d <= P.T) { const unsigned long long b = (c << 16) | d; k2 = b < k2 ? b : k2; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long y1 = __shfl_xor_sync(0xffffffffu, k1, o), y2 = __shfl_xor_sync(0xffffffffu, k2, o);
    k1 = y1 < k1 ? y1 : k1;
    k2 = y2 < k2 ? y2 : k2;
  }
  if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = k1; s2[threadIdx.x >> 5] = k2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { k1 = s1[w] < k1 ? s1[w] : k1; k2 = s2[w] < k2 ? s2[w] : k2; }
    if (k1 != ~0ull) atomicMin(best1 + li, k1);
    if (k2 != ~0ull) atomicMin(best2 + li, k2);
  }
}

// per patch: the record (compress_test.cpp:113-140 / compress4_test.cpp:136-170) and, for v4, the in-place replacement of a
// matched patch by its residual (:164). One block per patch, W threads do the rows.
__global__ void __launch_bounds__(32) k_match_decide(MatchParams P, uint64_t li0, const unsigned long long* __restrict__ best1,
                                                     const unsigned long long* __restrict__ best2, const double* __restrict__ enumL,
                                                     bic_match_rec* __restrict__ recs, uint32_t* __restrict__ Iw) {
  const uint64_t li = li0 + blockIdx.x;
  const long long i0 = (long long)((li / P.Nx) * P.W), j0 = (long long)((li % P.Nx) * P.W);
  const uint64_t M = (uint64_t)P.W * P.W;
  const Rects q = patch_rects(P, i0, j0);
  const unsigned long long b1 = best1[li], b2 = best2[li];
  uint64_t besti = 0, bestj = 0, bestd = P.mode == 1 ? M : M + 1;
  bool found = false;
  unsigned long long idx = 0;
  if (P.mode == 4 && b2 != ~0ull) { idx = b2 >> 16; bestd = b2 & 0xFFFFu; found = true; }       // the first distance <= T ends the search
  else if (b1 != ~0ull && (b1 >> 40) < bestd) { idx = b1 & ((1ull << 40) - 1); bestd = b1 >> 40; found = true; }  // strict <, :86 / :107
  if (found) {
    long long i2, j2;
    scan_to_pos(q, idx, &i2, &j2);
    besti = (uint64_t)i2; bestj = (uint64_t)j2;
  }
  const unsigned lane = threadIdx.x;
  const uint32_t prow = lane < P.W ? window_bits(P, (uint64_t)i0 + lane, (uint64_t)j0) : 0u;
  const uint32_t p2row = lane < P.W ? window_bits(P, besti + lane, bestj) : 0u;   // get_submatrix(besti, ..., bestj, ...), :116 / :143
  uint32_t w = __popc(prow);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
  const unsigned long long nomatch_len = (unsigned long long)(1 + enumL[w]);                       // idx_t = 1 + double
  unsigned long long match_len;
  bool use_match;
  unsigned long long idx_len = 0;
  while ((1ull << idx_len) < li) idx_len++;                                                        // ceil(log2(li)), li >= 1
  if (P.mode == 1) {
    if (li == 0) { match_len = 1ull << 63; use_match = false; }   // ceil(log2(0)) is undefined in the reference; see oracle/bic_oracle.c
    else { match_len = (unsigned long long)((double)(1 + idx_len) + enumL[bestd]); use_match = nomatch_len > match_len; }
  } else {
    match_len = bestd <= M ? (unsigned long long)((double)(1 + idx_len) + enumL[bestd]) : 100000ull;
    use_match = nomatch_len > match_len;
  }
  if (lane == 0) {
    bic_match_rec r;
    r.besti = besti; r.bestj = bestj; r.bestd = bestd; r.weight = w; r.match_len = match_len; r.nomatch_len = nomatch_len;
    r.use_match = use_match ? 1 : 0;
    recs[li] = r;
  }
  __syncwarp();  // every lane has read its rows of P and P2 before any lane rewrites a word of the image
  if (P.mode == 4 && use_match && lane < P.W && (uint64_t)i0 + lane < P.rows) {
    // the patch sits inside one 32-bit word (W | 32 and W | cols are required for v4) and no other thread writes this row's word
    const unsigned off = (unsigned)(j0 & 31);
    const uint32_t mask = (P.W >= 32 ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> P.W)) >> off;
    uint32_t* word = Iw + ((uint64_t)i0 + lane) * P.wpr + ((uint64_t)j0 >> 5);
    *word = (*word & ~mask) | (((prow ^ p2row) >> off) & mask);
  }
}

extern "C" double bic_enumL(uint64_t n, uint64_t r) {  // src/compress_test.cpp:37-40; gsl_sf_lnchoose through lgamma
  if (r == 0 || r >= n) return 0.0;
  const double ln = lgamma((double)n + 1.0) - lgamma((double)r + 1.0) - lgamma((double)(n - r) + 1.0);
  return ln * 1.442695040888963387004650940070860087872;
}

namespace {
struct GolombCounter {  // GolombCoder as the reference ships it: a bit COUNTER (src/GolombCoder.cpp:13-34, src/Golomb.h:12-29)
  uint32_t acc = 0, samples = 0, k = 1;
  uint64_t bitcount = 0;
  void code(uint32_t x) {
    bitcount += (uint64_t)k + (x >> k) + 1;
    samples++;
    acc += x;
    uint32_t kk = 0;
    while (kk < 31 && (uint32_t)(samples << kk) < acc) kk++;
    k = kk;
  }
};
}  // namespace

static bic_status match_run(bic_ctx* c, bic_mat* raster, uint32_t mode, uint64_t W, uint64_t T, uint64_t R, bic_match_rec* recs,
                            bic_match_totals* tot) {
  BIC_RANGE("bic:match_patches");
  if (!c || !raster || !recs || !tot) return BIC_ERR_INVALID;
  cudaSetDevice(c->device);
  const uint64_t rows = raster->rows, cols = raster->cols;
  if (W == 0 || W > 32 || rows == 0 || cols == 0 || cols < W) return bic_fail(c, BIC_ERR_UNSUPPORTED, "match: 1 <= W <= 32 <= cols");
  if (mode == 4 && ((32 % W) != 0 || (cols % W) != 0))
    return bic_fail(c, BIC_ERR_UNSUPPORTED, "match v4: W must divide 32 and the number of columns (the shapes where the reference's set_submatrix stays in its row)");
  if (T > 0xFFFF || rows * cols >= (1ull << 39)) return bic_fail(c, BIC_ERR_UNSUPPORTED, "match: raster too large for the packed keys");
  const uint64_t Ny = (W - 1 + rows) / W, Nx = (W - 1 + cols) / W, n = Nx * Ny, M = W * W;
  MatchParams P;
  P.I = raster->d; P.rows = rows; P.cols = cols; P.wpr = raster->wpr; P.S64 = 64 * div_up_u64(cols, 64);
  P.W = (uint32_t)W; P.mode = mode; P.T = T; P.R = R; P.Nx = Nx;
  // work[0]: best1 (n u64) | best2 (n u64) | enumL table (M + 2 doubles) | records (n)
  const size_t bytes = n * 16 + (M + 2) * 8 + n * sizeof(bic_match_rec);
  BIC_TRY(bic_scratch_reserve(c, &c->work[0], bytes + 64));
  unsigned long long* best1 = (unsigned long long*)c->work[0].p;
  unsigned long long* best2 = best1 + n;
  double* d_enum = (double*)(best2 + n);
  bic_match_rec* d_recs = (bic_match_rec*)(d_enum + M + 2);
  std::vector<double> tab(M + 2);
  for (uint64_t r = 0; r < M + 2; ++r) tab[r] = bic_enumL(M, r);
  BIC_CUDA(c, cudaMemsetAsync(best1, 0xFF, n * 16, c->stream));
  BIC_CUDA(c, cudaMemcpyAsync(d_enum, tab.data(), (M + 2) * 8, cudaMemcpyHostToDevice, c->stream));
  BIC_CUDA(c, bic_wait_stream(c));  // tab is a local
  if (mode == 1) {
    // all patches at once: blockIdx.y = patch, blockIdx.x strides over its candidates
    const uint64_t max_total = rows * cols;
    uint64_t gx = div_up_u64(max_total, 256 * 8);
    const uint64_t cap = ((uint64_t)c->sm_count * 16 + n - 1) / n;
    gx = gx > cap ? cap : gx;
    gx = gx < 1 ? 1 : gx;
    for (uint64_t l0 = 0; l0 < n; l0 += 65535) {
      const unsigned ny = (unsigned)((n - l0) < 65535 ? (n - l0) : 65535);
      BIC_PROF(c, KID_MATCH);
      k_match<<<dim3((unsigned)gx, ny), 256, 0, c->stream>>>(P, l0, best1, best2);
      BIC_LAUNCH_CHECK(c);
      BIC_PROF(c, KID_MATCH_DECIDE);
      k_match_decide<<<ny, 32, 0, c->stream>>>(P, l0, best1, best2, d_enum, d_recs, raster->d);
      BIC_LAUNCH_CHECK(c);
    }
  } else {
    // the serial chain: patch li searches the image as patches 0 .. li-1 left it
    for (uint64_t li = 0; li < n; ++li) {
      const long long i0 = (long long)((li / Nx) * W), j0 = (long long)((li % Nx) * W), Wl = (long long)W, Rl = (long long)R;
      const long long mini = i0 > Rl ? i0 - Rl : 0, mini2 = i0 > Wl ? i0 - Wl : 0, minj = j0 > Rl ? j0 - Rl : 0;
      const long long maxj = (j0 + Rl) > ((long long)cols - Wl) ? (long long)cols - Wl : j0 + Rl;
      uint64_t total = 0;
      if (j0 - Wl >= minj) total += (uint64_t)(i0 - mini2 + 1) * (uint64_t)(j0 - Wl - minj + 1);
      if (i0 - Wl >= mini && maxj >= minj) total += (uint64_t)(i0 - Wl - mini + 1) * (uint64_t)(maxj - minj + 1);
      if (total) {
        uint64_t gx = div_up_u64(total, 256);
        const uint64_t cap = (uint64_t)c->sm_count * 8;
        gx = gx > cap ? cap : gx;
        BIC_PROF(c, KID_MATCH);
        k_match<<<dim3((unsigned)gx, 1), 256, 0, c->stream>>>(P, li, best1, best2);
        BIC_LAUNCH_CHECK(c);
      }
      BIC_PROF(c, KID_MATCH_DECIDE);
      k_match_decide<<<1, 32, 0, c->stream>>>(P, li, best1, best2, d_enum, d_recs, raster->d);
      BIC_LAUNCH_CHECK(c);
    }
  }
  BIC_CUDA(c, cudaMemcpyAsync(recs, d_recs, n * sizeof(bic_match_rec), cudaMemcpyDeviceToHost, c->stream));
  BIC_CUDA(c, bic_wait_stream(c));
  // the accounting of the drivers' tail (compress_test.cpp:129-140): two Golomb counters over the samples in patch order
  GolombCounter gm, gn;
  memset(tot, 0, sizeof(*tot));
  for (uint64_t li = 0; li < n; ++li) {
    const bic_match_rec& r = recs[li];
    if (r.use_match) { gm.code((uint32_t)r.bestd); tot->weight_sum += r.bestd; tot->matches++; tot->L += (double)r.match_len; }
    else { gn.code((uint32_t)r.weight); tot->L += (double)r.nomatch_len; }
  }
  tot->bits_match = gm.bitcount;
  tot->bits_nomatch = gn.bitcount;
  return BIC_OK;
}

extern "C" bic_status bic_match_patches_v1(bic_ctx* c, const bic_mat* raster, uint64_t W, bic_match_rec* recs, bic_match_totals* tot) {
  return match_run(c, const_cast<bic_mat*>(raster), 1, W, 0, 0, recs, tot);  // mode 1 never writes the raster
}

extern "C" bic_status bic_match_patches_v4(bic_ctx* c, bic_mat* raster, uint64_t W, uint64_t T, uint64_t R, bic_match_rec* recs,
                                           bic_match_totals* tot) {
  return match_run(c, raster, 4, W, T, R, recs, tot);
}
