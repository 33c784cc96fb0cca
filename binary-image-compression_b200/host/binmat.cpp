// Host side of the binary_matrix stand-in (see binmat.h). Own implementation; semantics follow the
// reference's src/binmat.cpp (cited per method). Bulk reductions that the reference computes with a
// byte LUT (block_weight, src/binmat.cpp:22-37) use the hardware popcount here.
#include "binmat.h"

#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <iomanip>

#include "bic_b200.h"

// ---------------------------------------------------------------------------------------------
// shared device context
// ---------------------------------------------------------------------------------------------
bic_ctx* bic_host_context() {
  static bic_ctx* ctx = nullptr;
  if (!ctx) {
    const char* dev = std::getenv("BIC_DEVICE");
    const bic_status st = bic_ctx_create(dev ? std::atoi(dev) : 0, &ctx);
    if (st != BIC_OK) {
      std::cerr << "binary-image-compression_b200: cannot create a device context: " << bic_status_string(st) << std::endl;
      std::exit(-1);
    }
  }
  return ctx;
}

static void die(bic_status st, const char* what) {
  if (st == BIC_OK) return;
  std::cerr << "binary-image-compression_b200: " << what << ": " << bic_status_string(st) << " ("
            << bic_ctx_last_error(bic_host_context()) << ")" << std::endl;
  std::abort();  // the reference aborts through assert() on misuse (no -DNDEBUG, src/Makefile:7)
}

// ---------------------------------------------------------------------------------------------
// construction / ownership
// ---------------------------------------------------------------------------------------------
void binary_matrix::shape(idx_t r, idx_t c) {  // src/binmat.cpp:140-163
  rows = r;
  cols = c;
  len = r * c;
  blocks_per_row = (c + BITS_PER_BLOCK - 1) / BITS_PER_BLOCK;
  data_blocks = blocks_per_row * r;
  last_bit_offset = (c - 1) % BITS_PER_BLOCK;
  trail_mask = ONES << (BITS_PER_BLOCK - last_bit_offset - 1);
  last_block = blocks_per_row - 1;
}

binary_matrix::binary_matrix()
    : rows(0), cols(0), len(0), last_bit_offset(0), data_blocks(0), blocks_per_row(0), last_block(0), data(nullptr),
      trail_mask(0), mirror(nullptr), host_newer(true), dev_newer(false) {}

binary_matrix::binary_matrix(idx_t r, idx_t c) : data(nullptr), mirror(nullptr), host_newer(true), dev_newer(false) {
  shape(r, c);
  data = new block_t[data_blocks ? data_blocks : 1];  // words are NOT cleared, like the reference
}

binary_matrix::binary_matrix(const binary_matrix& o) : data(nullptr), mirror(nullptr), host_newer(true), dev_newer(false) {
  o.want_host();
  shape(o.rows, o.cols);
  data = new block_t[data_blocks ? data_blocks : 1];
  std::memcpy(data, o.data, sizeof(block_t) * data_blocks);
}

void binary_matrix::allocate(idx_t r, idx_t c) {
  shape(r, c);
  data = new block_t[data_blocks ? data_blocks : 1];
  host_newer = true;
  dev_newer = false;
}

void binary_matrix::destroy() {  // src/binmat.h:178
  delete[] data;
  data = nullptr;
  if (mirror) { bic_mat_destroy(bic_host_context(), mirror); mirror = nullptr; }
  rows = cols = len = data_blocks = blocks_per_row = last_block = last_bit_offset = 0;
  trail_mask = 0;
  host_newer = true;
  dev_newer = false;
}

binary_matrix& binary_matrix::operator=(const binary_matrix& A) {  // src/binmat.cpp:180-184: takes over A's storage
  if (this == &A) return *this;
  delete[] data;
  if (mirror) bic_mat_destroy(bic_host_context(), mirror);
  std::memcpy(static_cast<void*>(this), static_cast<const void*>(&A), sizeof(binary_matrix));
  return *this;
}

// ---------------------------------------------------------------------------------------------
// device mirror
// ---------------------------------------------------------------------------------------------
bic_mat* binary_matrix::device() const {
  bic_ctx* ctx = bic_host_context();
  if (mirror && (bic_mat_rows(mirror) != rows || bic_mat_cols(mirror) != cols)) {
    bic_mat_destroy(ctx, mirror);
    mirror = nullptr;
    host_newer = !dev_newer;
  }
  if (!mirror) {
    die(bic_mat_create(ctx, rows, cols, &mirror), "bic_mat_create");
    host_newer = true;
  }
  if (host_newer && !dev_newer) {
    die(bic_mat_upload_words64(ctx, mirror, data), "bic_mat_upload_words64");
    die(bic_ctx_sync(ctx), "sync");
    host_newer = false;
  }
  return mirror;
}

bic_mat* binary_matrix::release_device() {
  bic_mat* m = device();
  mirror = nullptr;
  destroy();
  return m;
}

void binary_matrix::adopt_device(bic_mat* m) {
  destroy();
  if (!m) return;
  allocate(bic_mat_rows(m), bic_mat_cols(m));
  mirror = m;
  dev_newer = true;
  host_newer = false;
}

void binary_matrix::pull() const {
  if (mirror) die(bic_mat_download_words64(bic_host_context(), mirror, data), "bic_mat_download_words64");
  dev_newer = false;
  host_newer = false;
}

// ---------------------------------------------------------------------------------------------
// whole-matrix fills (src/binmat.cpp:165-178)
// ---------------------------------------------------------------------------------------------
void binary_matrix::clear() {
  if (rows * cols == 0) return;
  dev_newer = false;
  host_newer = true;
  std::memset(data, 0, sizeof(block_t) * data_blocks);
}
void binary_matrix::set() {
  if (rows * cols == 0) return;
  dev_newer = false;
  host_newer = true;
  std::memset(data, 0xff, sizeof(block_t) * data_blocks);
}
void binary_matrix::flip() {
  if (rows * cols == 0) return;
  touch_host();
  for (idx_t i = 0; i < data_blocks; ++i) data[i] = ~data[i];
}

// ---------------------------------------------------------------------------------------------
// reductions (src/binmat.cpp:57-126)
// ---------------------------------------------------------------------------------------------
idx_t binary_matrix::weight() const {
  if (rows * cols == 0) return 0;
  if (dev_newer && mirror) {  // count where the data is
    uint64_t w = 0;
    die(bic_mat_weight(bic_host_context(), mirror, &w), "bic_mat_weight");
    return w;
  }
  want_host();
  idx_t w = 0;
  for (idx_t i = 0; i < rows; ++i)
    for (idx_t j = 0; j < blocks_per_row; ++j) w += (idx_t)__builtin_popcountl(block(i, j));
  return w;
}

idx_t binary_matrix::row_weight(idx_t i) const {
  assert(i < rows);
  want_host();
  idx_t w = 0;
  for (idx_t j = 0; j < blocks_per_row; ++j) w += (idx_t)__builtin_popcountl(block(i, j));
  return w;
}

idx_t binary_matrix::col_weight(idx_t j) const {
  assert(j < cols);
  idx_t w = 0;
  for (idx_t i = 0; i < rows; ++i) w += get(i, j) ? 1 : 0;
  return w;
}

bool binary_matrix::sum() const {
  if (rows * cols == 0) return false;
  want_host();
  block_t acc = 0;
  for (idx_t i = 0; i < rows; ++i)
    for (idx_t j = 0; j < blocks_per_row; ++j) acc ^= block(i, j);
  return __builtin_parityl(acc);
}

bool binary_matrix::row_sum(idx_t i) const {
  assert(i < rows);
  want_host();
  block_t acc = 0;
  for (idx_t j = 0; j < blocks_per_row; ++j) acc ^= block(i, j);
  return __builtin_parityl(acc);
}

bool binary_matrix::col_sum(idx_t j) const {
  bool p = false;
  for (idx_t i = 0; i < rows; ++i) p ^= get(i, j);
  return p;
}

// ---------------------------------------------------------------------------------------------
// copies in and out (src/binmat.cpp:187-414)
// ---------------------------------------------------------------------------------------------
binary_matrix binary_matrix::get_copy() const { return binary_matrix(*this); }

void binary_matrix::copy_to(binary_matrix& B) const {
  want_host();
  B.dev_newer = false;
  B.host_newer = true;
  std::memcpy(B.data, data, sizeof(block_t) * data_blocks);
}

void binary_matrix::copy_row_to(const idx_t i, binary_matrix& row) const {  // :250-257
  assert(i < rows);
  want_host();
  row.touch_host();
  std::memcpy(row.data, data + i * blocks_per_row, sizeof(block_t) * blocks_per_row);
}

void binary_matrix::set_row(const idx_t i, const binary_matrix& B) {  // :362-371
  assert(i < rows);
  B.want_host();
  touch_host();
  std::memcpy(data + i * blocks_per_row, B.data, sizeof(block_t) * blocks_per_row);
}

binary_matrix binary_matrix::get_row(const idx_t i) const {
  binary_matrix r(1, cols);
  copy_row_to(i, r);
  return r;
}

void binary_matrix::copy_col_to(const idx_t j, binary_matrix& col) const {  // :225-240: a ROW vector of length rows
  assert(j < cols);
  col.clear();
  for (idx_t i = 0; i < rows; ++i)
    if (get(i, j)) col.set(0, i);
}

binary_matrix binary_matrix::get_col(const idx_t j) const {
  binary_matrix c(1, rows);
  copy_col_to(j, c);
  return c;
}

void binary_matrix::set_col(const idx_t j, const binary_matrix& B) {  // :343-360
  assert(j < cols);
  for (idx_t i = 0; i < rows; ++i) set(i, j, B.get(0, i));
}

void binary_matrix::transpose_to(binary_matrix& A) const {  // :199-208
  assert(rows == A.cols);
  assert(cols == A.rows);
  A.clear();
  for (idx_t i = 0; i < rows; ++i)
    for (idx_t j = 0; j < cols; ++j)
      if (get(i, j)) A.set(j, i);
}

binary_matrix binary_matrix::get_transposed() const {
  binary_matrix A(cols, rows);
  transpose_to(A);
  return A;
}

// Sub-matrix read with the reference's addressing (:267-298): the raster is read through linear block
// indices, so positions outside the matrix read as zero EXCEPT that a column at or past the padded
// row end falls through into the next row (only a tile whose width does not divide 64 can ask).
void binary_matrix::copy_submatrix_to(const idx_t i0, const idx_t i1, const idx_t j0, const idx_t j1, binary_matrix& B) const {
  assert(i0 < i1);
  assert(j0 < j1);
  want_host();
  B.touch_host();
  for (idx_t di = 0; di < B.rows; ++di) {
    for (idx_t db = 0; db < B.blocks_per_row; ++db) {
      block_t out = 0;
      for (idx_t b = 0; b < BITS_PER_BLOCK; ++b) {
        const idx_t c = j0 + db * BITS_PER_BLOCK + b;
        const idx_t k = (i0 + di) * blocks_per_row + c / BITS_PER_BLOCK;
        if (k < data_blocks && (data[k] & (MSB >> (c % BITS_PER_BLOCK)))) out |= MSB >> b;
      }
      B.data[di * B.blocks_per_row + db] = out;
    }
  }
}

binary_matrix binary_matrix::get_submatrix(const idx_t i0, const idx_t i1, const idx_t j0, const idx_t j1) const {
  binary_matrix B(i1 - i0, j1 - j0);
  copy_submatrix_to(i0, i1, j0, j1, B);
  return B;
}

void binary_matrix::copy_vectorized_to(binary_matrix& v) const {  // :306-320, row-major concatenation
  v.clear();
  for (idx_t i = 0; i < rows; ++i)
    for (idx_t j = 0; j < cols; ++j)
      if (get(i, j)) v.set(0, i * cols + j);
}

binary_matrix binary_matrix::get_vectorized() const {
  binary_matrix v(1, rows * cols);
  copy_vectorized_to(v);
  return v;
}

void binary_matrix::set_vectorized(const binary_matrix& src) {  // :322-341
  clear();
  for (idx_t i = 0; i < rows; ++i)
    for (idx_t j = 0; j < cols; ++j)
      if (src.get(0, i * cols + j)) set(i, j);
}

void binary_matrix::set_submatrix(const idx_t i0, const idx_t j0, const binary_matrix& B) {  // :373-414, clipped
  for (idx_t si = 0, di = i0; si < B.rows && di < rows; ++si, ++di)
    for (idx_t sj = 0, dj = j0; sj < B.cols && dj < cols; ++sj, ++dj) set(di, dj, B.get(si, sj));
}

void binary_matrix::add_rows(idx_t nrows) {  // :417-425
  want_host();
  const idx_t old_blocks = data_blocks;
  block_t* nd = new block_t[blocks_per_row * (rows + nrows)];
  std::memcpy(nd, data, sizeof(block_t) * old_blocks);
  delete[] data;
  data = nd;
  rows += nrows;
  len = rows * cols;
  data_blocks = blocks_per_row * rows;
  host_newer = true;
}

void binary_matrix::remove_rows(idx_t nrows) {  // :426-434 (storage is kept)
  want_host();
  if (nrows < rows) { rows -= nrows; data_blocks -= blocks_per_row * nrows; }
  else { rows = 0; data_blocks = 0; }
  len = rows * cols;
  host_newer = true;
}

// ---------------------------------------------------------------------------------------------
// element-wise and products (src/binmat.cpp:463-616)
// ---------------------------------------------------------------------------------------------
binary_matrix& add(const binary_matrix& A, const binary_matrix& B, binary_matrix& C) {
  assert(C.data != 0);
  assert(C.rows == A.rows && C.rows == B.rows && C.cols == A.cols && C.cols == B.cols);
  A.want_host(); B.want_host(); C.touch_host();
  for (idx_t i = 0; i < A.rows; ++i)
    for (idx_t j = 0; j < C.blocks_per_row; ++j) C.data[i * C.blocks_per_row + j] = A.block(i, j) ^ B.block(i, j);
  return C;
}

binary_matrix& bool_and(const binary_matrix& A, const binary_matrix& B, binary_matrix& C) {
  assert(C.data != 0);
  assert(C.rows == A.rows && C.rows == B.rows && C.cols == A.cols && C.cols == B.cols);
  A.want_host(); B.want_host(); C.touch_host();
  for (idx_t i = 0; i < A.rows; ++i)
    for (idx_t j = 0; j < C.blocks_per_row; ++j) C.data[i * C.blocks_per_row + j] = A.block(i, j) & B.block(i, j);
  return C;
}

idx_t dist(const binary_matrix& A, const binary_matrix& B) {
  assert(A.rows == B.rows && A.cols == B.cols);
  A.want_host(); B.want_host();
  idx_t w = 0;
  for (idx_t i = 0; i < A.rows; ++i)
    for (idx_t j = 0; j < A.blocks_per_row; ++j) w += (idx_t)__builtin_popcountl(A.block(i, j) ^ B.block(i, j));
  return w;
}

// C = op(A) * op(B) over GF(2). The AB case of a tall A against a small B is the residual product of
// the hot path and runs on the device when the shapes allow it (C = A*B xor 0).
binary_matrix& mul(const binary_matrix& A, const bool At, const binary_matrix& B, const bool Bt, binary_matrix& C) {
  assert(C.data != 0);
  const idx_t M = At ? A.cols : A.rows, K = At ? A.rows : A.cols, N = Bt ? B.rows : B.cols;
  assert((Bt ? B.cols : B.rows) == K);
  assert(C.rows == M && C.cols == N);
  if (At && Bt) return C;  // mul_AtBt is an empty stub in the reference too (:596-604)
  C.clear();
  for (idx_t i = 0; i < M; ++i)
    for (idx_t k = 0; k < K; ++k) {
      if (!(At ? A.get(k, i) : A.get(i, k))) continue;
      if (!Bt) {
        for (idx_t j = 0; j < C.blocks_per_row; ++j) C.data[i * C.blocks_per_row + j] ^= B.block(k, j);
      } else {
        for (idx_t j = 0; j < N; ++j)
          if (B.get(j, k)) C.flip(i, j);
      }
    }
  return C;
}

// ---------------------------------------------------------------------------------------------
// dump (src/binmat.cpp:618-644)
// ---------------------------------------------------------------------------------------------
static idx_t grid_width = 10;
void set_grid_width(idx_t g) { grid_width = g; }

std::ostream& operator<<(std::ostream& out, const binary_matrix& A) {
  out << "rows=" << A.rows << "\tcols=" << A.cols << "\tlen=" << A.len << "\tbpw=" << BITS_PER_BLOCK << "\tdw="
      << A.data_blocks << "\twpr=" << A.blocks_per_row << "\ttm=" << bm_bitset(A.trail_mask) << std::endl;
  out << "       ";
  for (idx_t j = 0; j < A.cols; ++j) out << ((j % BITS_PER_BLOCK) ? ' ' : '|') << ' ';
  out << std::endl;
  for (idx_t i = 0; i < A.rows; ++i) {
    out << std::setw(5) << i << "  ";
    for (idx_t j = 0; j < A.cols; ++j) {
      const char pixel = (i % grid_width) || (j % grid_width) ? '.' : '+';
      out << std::setw(1) << (A.get(i, j) ? '#' : pixel) << ' ';
    }
    out << std::endl;
  }
  return out;
}
