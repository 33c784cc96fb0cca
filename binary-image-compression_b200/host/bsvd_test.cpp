// Driver with the reference's command line (src/bsvd_test.cpp:23-50) on top of the shim: read a PBM,
// build the sample matrix (image mode: patches; matrix mode: rows), initialise, learn, write
// dictionary.pbm / coefficients.pbm / residual.pbm (+ the image-shaped residual in image mode, as the
// reference does, :126-145), print |E| recomputed from A*D xor X (:153-155). Additionally (-g 1) the
// three matrices are Golomb coded on the device and the bit counts printed. The mosaics
// (render_mosaic, host visualisation) are not produced.
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "GolombCoder.h"
#include "bsvd.h"
#include "pbm.h"

int mi_algo = 0, cu_algo = 0, du_algo = 0, lm_algo = 0, lmi_algo = 0, golomb = 0;
idx_t W = 16, K = 512;
bool image_mode = false;
const char* iname = "data/test.pbm";

static void parse_args(int argc, char** argv) {
  for (int i = 1; i < argc; ++i) {
    if (argv[i][0] == '-') {
      if (i == argc - 1) { std::cerr << "Missing argument for " << argv[i] << std::endl; exit(-1); }
      const char* val = argv[i + 1];
      switch (argv[i][1]) {
        case 'i': mi_algo = atoi(val); break;
        case 'c': cu_algo = atoi(val); break;
        case 'd': du_algo = atoi(val); break;
        case 'l': lm_algo = atoi(val); break;
        case 'L': lmi_algo = atoi(val); break;
        case 'w': W = (idx_t)atoi(val); break;
        case 'k': K = (idx_t)atoi(val); break;
        case 'r': random_seed = atol(val); break;
        case 'I': image_mode = atoi(val) > 0; break;
        case 'm': case 'M': break;  // mosaics: accepted, not produced
        case 'g': golomb = atoi(val); break;
        default: std::cerr << "Invalid option " << argv[i] << std::endl; exit(-1);
      }
      i++;
    } else {
      iname = argv[i];
    }
  }
}

int main(int argc, char** argv) {
  parse_args(argc, argv);
  learn_model_setup(mi_algo, cu_algo, du_algo, lm_algo, lmi_algo);
  FILE* fimg = fopen(iname, "r");
  if (!fimg) return -1;
  idx_t rows, cols;
  const int res = read_pbm_header(fimg, rows, cols);
  std::cout << "rows=" << rows << " cols=" << cols << std::endl;
  if (res != PBM_OK) { std::cerr << "Error " << res << " reading image." << std::endl; std::exit(1); }
  binary_matrix I(rows, cols);
  read_pbm_data(fimg, I);
  fclose(fimg);

  idx_t M, N;
  binary_matrix X;
  if (image_mode) {
    std::cout << "==== DATA TREATED AS IMAGE, VECTORS ARE PATCHES =====\n" << std::endl;
    const idx_t Ny = (W - 1 + rows) / W, Nx = (W - 1 + cols) / W;
    M = W * W;
    N = Nx * Ny;
    std::cout << "Nx=" << Nx << " Ny=" << Ny << std::endl;
    X.allocate(N, M);
    extract_patches(I, W, X);  // the loop of src/bsvd_test.cpp:89-96 as one kernel
  } else {
    std::cout << "==== DATA TREATED AS MATRIX, VECTORS ARE ROWS =====\n" << std::endl;
    X = I.get_copy();
    M = I.get_cols();
    N = I.get_rows();
  }
  binary_matrix D(K, M), A(N, K);
  std::cout << "M=" << M << " N=" << N << " K=" << K << std::endl;
  initialize_model(X, D, A);
  binary_matrix E(N, M);
  const idx_t iters = learn_model(X, E, D, A);
  std::cout << "iterations=" << iters << std::endl;
  write_pbm(D, "dictionary.pbm");
  write_pbm(A, "coefficients.pbm");
  write_pbm(E, "residual.pbm");
  if (image_mode) {
    binary_matrix R(rows, cols);
    assemble_patches(E, W, R);  // src/bsvd_test.cpp:128-139
    write_pbm(R, "residual.pbm");
    R.destroy();
  }
  if (golomb) {
    std::vector<uint8_t> bytes;
    std::cout << "golomb bits D=" << golomb_encode(D, bytes);
    std::cout << " A=" << golomb_encode(A, bytes);
    std::cout << " E=" << golomb_encode(E, bytes) << std::endl;
  }
  binary_matrix E2(N, M);
  residual(X, A, D, E2);  // mul(A,false,D,false,E); add(E,X,E)
  std::cout << "|E|" << E2.weight() << std::endl;
  E2.destroy(); A.destroy(); I.destroy(); E.destroy(); D.destroy(); X.destroy();
  return 0;
}
