// P4 I/O for the shim (behaviour of src/pbm.cpp:4-77, src/util.cpp:83-89; own code, whole-byte moves
// instead of the reference's bit-by-bit set/get).
#include "pbm.h"

#include <vector>

ErrorCode read_pbm_header(FILE* f, idx_t& rows, idx_t& cols) {
  if (fgetc(f) != 'P' || fgetc(f) != '4') return PBM_INVALID_HEADER;
  int w = 0, h = 0;
  if (fscanf(f, " %d", &w) != 1 || w == 0) return PBM_INVALID_HEADER;
  if (fscanf(f, " %d ", &h) != 1 || h == 0) return PBM_INVALID_HEADER;  // consumes the single whitespace after the height
  rows = (idx_t)h;
  cols = (idx_t)w;
  return PBM_OK;
}

ErrorCode read_pbm_data(FILE* f, binary_matrix& A) {
  A.clear();
  const idx_t bpr = (A.get_cols() + 7) / 8;
  std::vector<unsigned char> line(bpr ? bpr : 1);
  for (idx_t i = 0; i < A.get_rows(); ++i) {
    // like the reference (src/pbm.cpp:37-47) a short read keeps the pixels that did arrive: its
    // header parser (" %d ", :18) also swallows leading data bytes that happen to be white space
    const size_t got = fread(line.data(), 1, bpr, f);
    for (idx_t j = 0; j < A.get_cols() && (j >> 3) < got; ++j)
      if (line[j >> 3] & (0x80u >> (j & 7))) A.set(i, j);
    if (got != bpr) return PBM_INVALID_DATA;
  }
  return PBM_OK;
}

ErrorCode write_pbm(binary_matrix& A, FILE* f) {
  fprintf(f, "P4\n%lu %lu\n", A.get_cols(), A.get_rows());
  const idx_t bpr = (A.get_cols() + 7) / 8;
  std::vector<unsigned char> line(bpr ? bpr : 1);
  for (idx_t i = 0; i < A.get_rows(); ++i) {
    std::fill(line.begin(), line.end(), 0);
    for (idx_t j = 0; j < A.get_cols(); ++j)
      if (A.get(i, j)) line[j >> 3] |= (unsigned char)(0x80u >> (j & 7));
    if (fwrite(line.data(), 1, bpr, f) != bpr) return PBM_WRITE_ERROR;
  }
  return PBM_OK;
}

int write_pbm(binary_matrix& A, const char* fname) {
  FILE* f = fopen(fname, "w");
  if (!f) return -2;
  write_pbm(A, f);
  fclose(f);
  return 0;
}
