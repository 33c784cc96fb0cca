// Source-compatible stand-in for the reference's binmat.h (src/binmat.h:1-236): same type names,
// macros, public methods and free functions, so code written against the reference compiles
// unchanged. New implementation: the words live on the host in the reference's layout (row-major
// 64-bit blocks, bit j of a row at MSB >> (j % 64)) AND, lazily, in a device mirror owned through
// the C ABI (include/bic_b200.h). Host accessors pull the mirror when the device copy is newer;
// the CUDA plug-ins (bsvd.h) push it when the host copy is newer.
//
// Kept on purpose (callers rely on them): no destructor + explicit destroy() (binmat.h:47,178);
// operator= releases the left side and takes over the right side's storage (binmat.cpp:180-184);
// the copy constructor is deep (:128-138); constructors do not clear the words.
// Not kept (bugs the reference's own comments flag, never hit by the hot path): col_weight's row
// stride for more than one block per row (:87), get_transposed's shape (:210-214), the
// out-of-range shift in copy_vectorized_to for W % 64 == 0 (:316).
#ifndef BIC_HOST_BINMAT_H
#define BIC_HOST_BINMAT_H

#include <bitset>
#include <cstring>
#include <iostream>

typedef unsigned long idx_t;
typedef unsigned long block_t;

#define BITS_PER_BLOCK (sizeof(block_t) * 8)
#define ONES (~block_t(0))
#define ZEROES (block_t(0))
#define LSB block_t(1)
#define MSB (LSB << (BITS_PER_BLOCK - 1))
#define IMSB (ONES >> 1)
#define ILSB (ONES << 1)
#define XOR(a, b) ((!(a) && (b)) || ((a) && !(b)))

typedef std::bitset<BITS_PER_BLOCK> bm_bitset;

struct bic_mat;  // device mirror (C ABI)

class binary_matrix {
 public:
  binary_matrix(idx_t _rows, idx_t _cols);
  binary_matrix();
  binary_matrix(const binary_matrix& other);  // deep copy
  ~binary_matrix() {}                         // storage is released explicitly, see destroy()

  void allocate(idx_t _rows, idx_t _cols);
  void destroy();

  binary_matrix get_vectorized() const;
  binary_matrix get_col(const idx_t j) const;
  binary_matrix get_row(const idx_t i) const;
  binary_matrix get_submatrix(const idx_t i0, const idx_t i1, const idx_t j0, const idx_t j1) const;
  binary_matrix get_copy() const;
  binary_matrix get_transposed() const;

  void copy_vectorized_to(binary_matrix& B) const;
  void copy_col_to(const idx_t j, binary_matrix& B) const;
  void copy_row_to(const idx_t i, binary_matrix& B) const;
  void copy_submatrix_to(const idx_t i0, const idx_t i1, const idx_t j0, const idx_t j1, binary_matrix& B) const;
  void copy_to(binary_matrix& B) const;
  void transpose_to(binary_matrix& B) const;

  void set_vectorized(const binary_matrix& src);
  void set_col(const idx_t j, const binary_matrix& src);
  void set_row(const idx_t i, const binary_matrix& src);
  void set_submatrix(const idx_t i0, const idx_t j0, const binary_matrix& src);

  void add_rows(idx_t nrows);
  void remove_rows(idx_t nrows);

  inline idx_t get_rows() const { return rows; }
  inline idx_t get_cols() const { return cols; }
  inline idx_t get_len() const { return len; }

  void clear();
  void set();
  void flip();

  inline bool get(const idx_t i, const idx_t j) const {
    want_host();
    return (data[i * blocks_per_row + (j / BITS_PER_BLOCK)] & (MSB >> (j % BITS_PER_BLOCK))) != 0;
  }
  inline void set(const idx_t i, const idx_t j, const bool v) {
    if (v) set(i, j); else clear(i, j);
  }
  inline void set(const idx_t i, const idx_t j) {
    touch_host();
    data[i * blocks_per_row + (j / BITS_PER_BLOCK)] |= (MSB >> (j % BITS_PER_BLOCK));
  }
  inline void flip(const idx_t i, const idx_t j) {
    touch_host();
    data[i * blocks_per_row + (j / BITS_PER_BLOCK)] ^= (MSB >> (j % BITS_PER_BLOCK));
  }
  inline void clear(const idx_t i, const idx_t j) {
    touch_host();
    data[i * blocks_per_row + (j / BITS_PER_BLOCK)] &= ~(MSB >> (j % BITS_PER_BLOCK));
  }

  idx_t weight() const;
  idx_t row_weight(idx_t i) const;
  idx_t col_weight(idx_t j) const;
  bool sum() const;
  bool row_sum(idx_t i) const;
  bool col_sum(idx_t j) const;

  friend std::ostream& operator<<(std::ostream& out, const binary_matrix& A);
  friend binary_matrix& add(const binary_matrix& A, const binary_matrix& B, binary_matrix& C);
#define bool_xor add
  friend binary_matrix& bool_and(const binary_matrix& A, const binary_matrix& B, binary_matrix& C);
  friend binary_matrix& mul(const binary_matrix& A, const bool At, const binary_matrix& B, const bool Bt, binary_matrix& C);
  friend idx_t dist(const binary_matrix& A, const binary_matrix& B);

  binary_matrix& operator=(const binary_matrix& A);

  // ---- B200 extension (not in the reference): the device mirror ------------------------------
  /** device handle with the current contents (uploads if the host copy is newer) */
  bic_mat* device() const;
  /** the device copy was written by a kernel: the host words are stale until next read */
  void device_written() const { dev_newer = true; host_newer = false; }
  /** give up the device mirror (made current first): the caller owns the handle; this matrix is destroy()ed */
  bic_mat* release_device();
  /** destroy() + allocate(rows, cols of m) with m as the (newer) device copy; m = nullptr leaves the 0 x 0 matrix */
  void adopt_device(bic_mat* m);
  /** raw host words in the reference layout (pulls the mirror first) */
  const block_t* host_words() const { want_host(); return data; }

 private:
  inline block_t block(const idx_t i, const idx_t j) const {
    return (j < last_block) ? data[i * blocks_per_row + j] : (data[i * blocks_per_row + j] & trail_mask);
  }
  void shape(idx_t r, idx_t c);
  void want_host() const { if (dev_newer) pull(); }
  void touch_host() { if (dev_newer) pull(); host_newer = true; }
  void pull() const;

  idx_t rows, cols, len;
  idx_t last_bit_offset, data_blocks, blocks_per_row, last_block;
  block_t* data;
  block_t trail_mask;
  mutable bic_mat* mirror;
  mutable bool host_newer, dev_newer;
};

void set_grid_width(idx_t g);

// the process-wide device context the shim's matrices and plug-ins share (created on first use;
// aborts with a message if there is no CUDA device: there is no CPU fallback)
struct bic_ctx;
bic_ctx* bic_host_context();

#endif
