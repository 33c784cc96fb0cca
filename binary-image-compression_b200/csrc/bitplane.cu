// Bit planes of a grey image: the loop of the reference's bitplane_tool (src/bitplane_tool.cpp:24-39) as a kernel.
// Plane bi (mask b = 1 << bi, for every b < maxval) holds A(i, j) = gray(i, j) & b; the input is the P5 payload as
// read_pgm_p5_data reads it (src/pnm.cpp:54-78): one byte per pixel when maxval < 256, else two, high byte first.
// BASELINE.json configs[1] starts here: one 16-bit 8192 x 8192 PGM -> 16 binary rasters, each then fitted and coded.
//
// 16-bit fast path: a lane loads 8 pixels (16 bytes, coalesced 512 B per warp), splits them into the 8 x 8 bit
// matrices of their high and low bytes (PRMT), transposes both with three masked-swap steps each, so every byte of
// the result is "bit b of 8 consecutive pixels"; four neighbouring lanes exchange those bytes (3 shuffles + a 4 x 4
// byte transpose with PRMT) so each lane ends with four complete 32-pixel words of four planes, stored as full
// 32-byte sectors per plane. 2 bytes in + 2 bytes out per pixel: HBM bound.
#include "bic_internal.cuh"

struct PlanePtrs { uint32_t* p[16]; };

// 8 x 8 bit transpose of a 64-bit word in MSB-first coordinates: byte r (from the top) is row r, bit c of it column c
__device__ __forceinline__ unsigned long long transpose8x8(unsigned long long x) {
  unsigned long long t;
  t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull;  x = x ^ t ^ (t << 7);
  t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull; x = x ^ t ^ (t << 14);
  t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull; x = x ^ t ^ (t << 28);
  return x;
}

__global__ void __launch_bounds__(256) k_bitplanes16(const uint8_t* __restrict__ pgm, uint64_t rows, uint64_t cols, uint64_t wpr,
                                                     uint32_t nplanes, PlanePtrs P) {
  __shared__ uint32_t* s_plane[16];       // the lane's planes depend on its position in the group: indexed at run time
  if (threadIdx.x < 16) s_plane[threadIdx.x] = P.p[threadIdx.x];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, q = lane & 3, g = lane >> 2;
  // A warp keeps ONE column tile and walks down the rows (the host sizes the grid so that the warps in use are a multiple
  // of the tiles per row): no division in the loop, and a warp's stores to a plane are a fixed-stride stream.
  const uint32_t tiles_per_row = (uint32_t)div_up_u64(cols, 256);
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const uint32_t row_step = nw / tiles_per_row;          // >= 1 (host)
  if (gw >= row_step * tiles_per_row) return;            // spare warps
  const uint32_t tile = gw % tiles_per_row, row0 = gw / tiles_per_row;
  const uint64_t px = (uint64_t)tile * 256 + lane * 8;   // this lane's 8 pixels of every row (cols % 8 == 0: all or nothing)
  const bool in_row = px < cols;
  const uint8_t* src = pgm + ((uint64_t)row0 * cols + px) * 2;
  const uint64_t src_step = (uint64_t)row_step * cols * 2;
  // software pipeline: the next row's 16 bytes are requested before the current ones are transposed
  uint4 v_next = make_uint4(0, 0, 0, 0);
  if (in_row && row0 < rows) v_next = __ldg(reinterpret_cast<const uint4*>(src));
  for (uint64_t row = row0; row < rows; row += row_step) {
    const uint4 v = v_next;
    v_next = make_uint4(0, 0, 0, 0);
    src += src_step;
    if (in_row && row + row_step < rows) v_next = __ldg(reinterpret_cast<const uint4*>(src));
    // bytes in memory: hi0 lo0 hi1 lo1 ... ; gather the 8 high bytes and the 8 low bytes, pixel 0 on top
    const unsigned long long H = ((unsigned long long)__byte_perm(v.x, v.y, 0x0246) << 32) | __byte_perm(v.z, v.w, 0x0246);
    const unsigned long long L = ((unsigned long long)__byte_perm(v.x, v.y, 0x1357) << 32) | __byte_perm(v.z, v.w, 0x1357);
    const unsigned long long TH = transpose8x8(H), TL = transpose8x8(L);
    // W[c]: four plane bytes, planes 15-4c .. 12-4c from the top byte down, each with this lane's 8 pixels
    const uint32_t W0 = (uint32_t)(TH >> 32), W1 = (uint32_t)TH, W2 = (uint32_t)(TL >> 32), W3 = (uint32_t)TL;
    // 4 x 4 transpose of the W words across each group of four lanes (lane q ends with X[s] = W[q] of lane s): two
    // butterfly stages, each swapping half of the words with a partner lane
    uint32_t X[4];
    {
      // stage 1, partner lane ^ 2: lanes 0,1 keep W0,W1 and take the partner's W0,W1; lanes 2,3 keep W2,W3
      const bool up = (q & 2) != 0;
      const uint32_t s0 = __shfl_xor_sync(0xffffffffu, up ? W0 : W2, 2);
      const uint32_t s1 = __shfl_xor_sync(0xffffffffu, up ? W1 : W3, 2);
      // after stage 1: A = word c0 of (lane with bit1 = 0), B = same word of (lane with bit1 = 1), for c0 in {0,1} (or {2,3})
      const uint32_t a0 = up ? s0 : W0, b0 = up ? W2 : s0;   // words (0 or 2): from lanes q&1, (q&1)|2
      const uint32_t a1 = up ? s1 : W1, b1 = up ? W3 : s1;   // words (1 or 3)
      // stage 2, partner lane ^ 1: even lanes keep the lower word of each pair, odd lanes the upper
      const bool odd = (q & 1) != 0;
      const uint32_t r0 = __shfl_xor_sync(0xffffffffu, odd ? a0 : a1, 1);
      const uint32_t r1 = __shfl_xor_sync(0xffffffffu, odd ? b0 : b1, 1);
      // lane q now holds word q of lanes {0,1} (from a*) and {2,3} (from b*)
      X[0] = odd ? r0 : a0;
      X[1] = odd ? a1 : r0;
      X[2] = odd ? r1 : b0;
      X[3] = odd ? b1 : r1;
    }
    // 4 x 4 byte transpose: word i = byte (3 - i) of X0..X3, X0 (the first 8 pixels) on top
    const uint32_t a = __byte_perm(X[0], X[1], 0x3715), b = __byte_perm(X[0], X[1], 0x2604);
    const uint32_t c2 = __byte_perm(X[2], X[3], 0x3715), d = __byte_perm(X[2], X[3], 0x2604);
    uint32_t o[4];
    o[0] = __byte_perm(c2, a, 0x7632);
    o[1] = __byte_perm(d, b, 0x7632);
    o[2] = __byte_perm(c2, a, 0x5410);
    o[3] = __byte_perm(d, b, 0x5410);
    const uint64_t word = (uint64_t)tile * 8 + g;   // the group's 32 pixels
    if (word < wpr) {  // row-invariant
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t plane = 15 - (4 * q + i);
        if (plane < nplanes) s_plane[plane][row * wpr + word] = o[i];
      }
    }
  }
}

// any width, one or two bytes per pixel: a lane per pixel, a vote per plane
__global__ void __launch_bounds__(256) k_bitplanes_generic(const uint8_t* __restrict__ pgm, uint64_t rows, uint64_t cols, uint64_t wpr,
                                                           uint32_t nplanes, int two_bytes, PlanePtrs P) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t nwords = rows * wpr;
  const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t t = gw; t < nwords; t += nw) {
    const uint64_t row = t / wpr, w = t - row * wpr;
    const uint64_t j = w * 32 + lane;
    uint32_t pix = 0;
    if (j < cols) {
      const uint64_t i = row * cols + j;
      pix = two_bytes ? (((uint32_t)__ldg(pgm + 2 * i) << 8) + __ldg(pgm + 2 * i + 1)) : __ldg(pgm + i);  // src/pnm.cpp:62,71
    }
    for (uint32_t b = 0; b < nplanes; ++b) {
      const uint32_t word = __brev(__ballot_sync(0xffffffffu, (pix >> b) & 1u));
      if (lane == 0) P.p[b][t] = word;
    }
  }
}

// number of planes bitplane_tool writes: one per b = 1, 2, 4, ... with b < maxval (src/bitplane_tool.cpp:24)
extern "C" uint32_t bic_bitplane_count(uint32_t maxval) {
  uint32_t n = 0;
  for (uint64_t b = 1; b < maxval; b <<= 1) n++;
  return n;
}

bic_status bic_k_split_bitplanes_dev(bic_ctx* c, const uint8_t* d_payload, uint64_t rows, uint64_t cols, uint32_t maxval,
                                     bic_mat* const* planes, uint32_t nplanes) {
  if (nplanes == 0 || rows == 0 || cols == 0) return BIC_OK;
  PlanePtrs P;
  for (uint32_t b = 0; b < 16; ++b) P.p[b] = b < nplanes ? planes[b]->d : nullptr;
  const uint64_t wpr = planes[0]->wpr;
  const int two = maxval >= 256;
  BIC_PROF(c, KID_BITPLANES);
  const uint64_t ntiles = rows * div_up_u64(cols, 256);
  if (two && cols % 8 == 0 && ((uintptr_t)d_payload & 15) == 0 && ntiles < (1ull << 31)) {
    // warps in use: a multiple of the tiles per row (each warp keeps one column tile), about 64 per SM
    const uint64_t tpr = div_up_u64(cols, 256);
    uint64_t warps = (uint64_t)c->sm_count * 64;
    if (warps > ntiles) warps = ntiles;
    warps = warps < tpr ? tpr : (warps / tpr) * tpr;
    const int grid = (int)div_up_u64(warps, 8);
    k_bitplanes16<<<grid, 256, 0, c->stream>>>(d_payload, rows, cols, wpr, nplanes, P);
  } else {
    k_bitplanes_generic<<<bic_grid_for(c, rows * wpr * 32, 256, 8), 256, 0, c->stream>>>(d_payload, rows, cols, wpr, nplanes, two, P);
  }
  BIC_LAUNCH_CHECK(c);
  return BIC_OK;
}

extern "C" bic_status bic_split_bitplanes(bic_ctx* c, const uint8_t* p5_payload, uint64_t rows, uint64_t cols, uint32_t maxval,
                                          bic_mat* const* planes, uint32_t nplanes) {
  BIC_RANGE("bic:split_bitplanes");
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !p5_payload || !planes || maxval == 0 || maxval > 65535) return BIC_ERR_INVALID;
  if (nplanes != bic_bitplane_count(maxval) || nplanes > 16) return bic_fail(c, BIC_ERR_INVALID, "split_bitplanes: one matrix per b = 1, 2, 4, ... < maxval");
  for (uint32_t b = 0; b < nplanes; ++b)
    if (!planes[b] || planes[b]->rows != rows || planes[b]->cols != cols) return bic_fail(c, BIC_ERR_INVALID, "split_bitplanes: every plane must be rows x cols");
  const size_t bytes = (size_t)rows * cols * (maxval >= 256 ? 2 : 1);
  BIC_TRY(bic_scratch_reserve(c, &c->staging, bytes));
  BIC_CUDA(c, cudaMemcpyAsync(c->staging.p, p5_payload, bytes, cudaMemcpyHostToDevice, c->stream));
  return bic_k_split_bitplanes_dev(c, (const uint8_t*)c->staging.p, rows, cols, maxval, planes, nplanes);
}
