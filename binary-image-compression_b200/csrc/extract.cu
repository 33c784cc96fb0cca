// Patch extraction (raster -> patch matrix X) and its inverse.
// Reference: src/bsvd_test.cpp:80-99 using copy_submatrix_to (src/binmat.cpp:267-298),
// copy_vectorized_to (:306-320), set_row (:362-371); inverse src/bsvd_test.cpp:128-139.
#include "bic_internal.cuh"

// Read `len` (1..32) raster bits of row r starting at column c, right-aligned in the result.
// Pixels outside the raster read as zero. `bpr64_bits` = 64*ceil(cols/64): the reference reads
// the raster through linear word indices (binmat.cpp:275-291), so a column at or past the
// padded row end belongs to the next raster row; only tiles with W not dividing 64 get there.
__device__ __forceinline__ uint32_t raster_bits(const uint32_t* __restrict__ I, uint64_t rows, uint64_t wpr,
                                                uint64_t bpr64_bits, uint64_t r, uint64_t c, unsigned len) {
  if (c >= bpr64_bits) { c -= bpr64_bits; r += 1; }
  if (r >= rows) return 0u;
  const uint64_t wi = c >> 5;
  const unsigned off = (unsigned)(c & 31);
  const uint32_t* row = I + r * wpr;
  const uint32_t hi = (wi < wpr) ? __ldg(row + wi) : 0u;
  const uint32_t lo = (off + len > 32 && wi + 1 < wpr) ? __ldg(row + wi + 1) : 0u;
  const uint32_t v = __funnelshift_l(lo, hi, off);  // bits c.. at the top
  return v >> (32 - len);
}

// One thread per 32-bit word of X. Word q of patch (i,j) holds vectorised bits [32q, 32q+32):
// bit b of the vectorisation is pixel (b / W, b % W) of the tile, so the word is a few
// contiguous segments of consecutive tile rows.
__global__ void k_extract(const uint32_t* __restrict__ I, uint32_t* __restrict__ X, uint64_t rows, uint64_t cols,
                          uint64_t wprI, uint64_t W, uint64_t Nx, uint64_t n, uint64_t wprX, uint64_t m) {
  const uint64_t total = n * wprX;
  const uint64_t bpr64_bits = ((cols + 63) >> 6) << 6;
  for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t li = idx / wprX, q = idx - li * wprX;
    const uint64_t ti = li / Nx, tj = li - ti * Nx;
    uint64_t b = q * 32;
    const uint64_t bend = (b + 32 < m) ? b + 32 : m;
    uint32_t out = 0;
    unsigned filled = 0;
    while (b < bend) {
      const uint64_t pr = b / W, pc = b - pr * W;
      uint64_t len = W - pc;
      if (len > bend - b) len = bend - b;
      const uint64_t c = tj * W + pc;
      // a segment must not straddle the padded row end (the two halves live in different rows)
      if (c < bpr64_bits && c + len > bpr64_bits) len = bpr64_bits - c;
      const uint32_t v = raster_bits(I, rows, wprI, bpr64_bits, ti * W + pr, c, (unsigned)len);
      out |= v << (32 - filled - (unsigned)len);
      filled += (unsigned)len;
      b += len;
    }
    X[idx] = out;
  }
}

// Fast path for W = 2^LW <= 32 (8, 16, 32: every config BASELINE.json names): a word of X is 32/W whole tile
// rows, each W aligned bits of one raster word, so there is no division, no funnel shift and no edge wrap
// (W divides 64). blockIdx.y = tile row; threads run over (tile column, word) pairs.
template <int LW>
__global__ void __launch_bounds__(256) k_extract_pow2(const uint32_t* __restrict__ I, uint32_t* __restrict__ X, uint32_t rows,
                                                      uint32_t cols, uint32_t wprI, uint32_t Nx) {
  constexpr uint32_t W = 1u << LW;
  constexpr uint32_t SEG = 32u / W;          // tile rows per word of X
  constexpr uint32_t WPX = (W * W) / 32u;    // words per patch (W >= 8) -- W*W is a multiple of 32 for W >= 8
  const uint32_t ti = blockIdx.y;
  const uint32_t per_row = Nx * WPX;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < per_row; i += gridDim.x * blockDim.x) {
    const uint32_t tj = i / WPX, q = i - tj * WPX;   // WPX is a power of two: shifts
    const uint32_t c = tj << LW;                     // first raster column of the tile
    uint32_t out = 0;
    if (c < cols) {
      const uint32_t wi = c >> 5, sh = 32u - W - (c & 31u);
#pragma unroll
      for (uint32_t s = 0; s < SEG; ++s) {
        const uint32_t r = (ti << LW) + q * SEG + s;
        uint32_t seg = 0;
        if (r < rows) seg = (W == 32) ? __ldg(I + (uint64_t)r * wprI + wi) : ((__ldg(I + (uint64_t)r * wprI + wi) >> sh) & ((1u << (W & 31)) - 1u));
        out |= seg << (32u - W * (s + 1u));
      }
    }
    X[((uint64_t)ti * Nx + tj) * WPX + q] = out;
  }
}

// Inverse: one thread per 32-bit word of the raster; gathers its bits from the patches.
__global__ void k_assemble(const uint32_t* __restrict__ X, uint32_t* __restrict__ I, uint64_t rows, uint64_t cols,
                           uint64_t wprI, uint64_t W, uint64_t Nx, uint64_t wprX) {
  const uint64_t total = rows * wprI;
  for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = idx / wprI, q = idx - r * wprI;
    const uint64_t ti = r / W, pr = r - ti * W;
    uint64_t c = q * 32;
    const uint64_t cend = (c + 32 < cols) ? c + 32 : cols;
    uint32_t out = 0;
    unsigned filled = 0;
    while (c < cend) {
      const uint64_t tj = c / W, pc = c - tj * W;
      uint64_t len = W - pc;
      if (len > cend - c) len = cend - c;
      const uint64_t b = pr * W + pc;  // first vectorised bit
      const uint32_t* xr = X + (ti * Nx + tj) * wprX;
      const uint64_t wi = b >> 5;
      const unsigned off = (unsigned)(b & 31);
      const uint32_t hi = __ldg(xr + wi);
      const uint32_t lo = (off + len > 32 && wi + 1 < wprX) ? __ldg(xr + wi + 1) : 0u;
      const uint32_t v = __funnelshift_l(lo, hi, off) >> (32 - (unsigned)len);
      out |= v << (32 - filled - (unsigned)len);
      filled += (unsigned)len;
      c += len;
    }
    I[idx] = out;
  }
}

extern "C" bic_status bic_extract_patches(bic_ctx* c, const bic_mat* raster, uint64_t W, bic_mat* X) {
  BIC_RANGE("bic:extract_patches");
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !raster || !X || W == 0) return BIC_ERR_INVALID;
  const uint64_t Ny = (W - 1 + raster->rows) / W, Nx = (W - 1 + raster->cols) / W;  // bsvd_test.cpp:82-83
  if (X->rows != Nx * Ny || X->cols != W * W) return bic_fail(c, BIC_ERR_INVALID, "extract: X must be Nx*Ny x W*W");
  if (X->rows == 0) return BIC_OK;
  if ((W == 8 || W == 16 || W == 32) && raster->rows < (1ull << 31) && raster->cols < (1ull << 31) && Ny < 65536) {
    const uint32_t per_row = (uint32_t)(Nx * (W * W / 32));
    dim3 grid((unsigned)((per_row + 255) / 256 < 64 ? (per_row + 255) / 256 : 64), (unsigned)Ny);
    BIC_PROF(c, KID_EXTRACT);
    if (W == 8) k_extract_pow2<3><<<grid, 256, 0, c->stream>>>(raster->d, X->d, (uint32_t)raster->rows, (uint32_t)raster->cols, (uint32_t)raster->wpr, (uint32_t)Nx);
    else if (W == 16) k_extract_pow2<4><<<grid, 256, 0, c->stream>>>(raster->d, X->d, (uint32_t)raster->rows, (uint32_t)raster->cols, (uint32_t)raster->wpr, (uint32_t)Nx);
    else k_extract_pow2<5><<<grid, 256, 0, c->stream>>>(raster->d, X->d, (uint32_t)raster->rows, (uint32_t)raster->cols, (uint32_t)raster->wpr, (uint32_t)Nx);
    BIC_LAUNCH_CHECK(c);
    return BIC_OK;
  }
  BIC_PROF(c, KID_EXTRACT);
  k_extract<<<bic_grid_for(c, X->words(), 256, 16), 256, 0, c->stream>>>(raster->d, X->d, raster->rows, raster->cols,
                                                                        raster->wpr, W, Nx, X->rows, X->wpr, X->cols);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

extern "C" bic_status bic_assemble_patches(bic_ctx* c, const bic_mat* X, uint64_t W, bic_mat* raster) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !raster || !X || W == 0) return BIC_ERR_INVALID;
  const uint64_t Ny = (W - 1 + raster->rows) / W, Nx = (W - 1 + raster->cols) / W;
  if (X->rows != Nx * Ny || X->cols != W * W) return bic_fail(c, BIC_ERR_INVALID, "assemble: X must be Nx*Ny x W*W");
  if (raster->words() == 0) return BIC_OK;
  BIC_PROF(c, KID_ASSEMBLE);
  k_assemble<<<bic_grid_for(c, raster->words(), 256, 16), 256, 0, c->stream>>>(X->d, raster->d, raster->rows,
                                                                              raster->cols, raster->wpr, W, Nx, X->wpr);
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}
