// Stand-in for src/pbm.h:1-22 (P4 reader/writer, host side; file I/O stays on the host).
#ifndef BIC_HOST_PBM_H
#define BIC_HOST_PBM_H
#include <cstdio>
#include "binmat.h"
typedef enum error_code {
  PBM_OK = 0, PBM_READ_ERROR = 1, PBM_FILE_NOT_FOUND = 2, PBM_INVALID_HEADER = 3, PBM_INVALID_DATA = 4,
  PBM_WRITE_ERROR = 5, PBM_INVALID_FORMAT = 6
} ErrorCode;
ErrorCode read_pbm_header(FILE* fimg, idx_t& rows, idx_t& cols);
ErrorCode read_pbm_data(FILE* fimg, binary_matrix& A);
ErrorCode write_pbm(binary_matrix& A, FILE* fimg);
int write_pbm(binary_matrix& A, const char* fname);  // src/util.h:8
#endif
