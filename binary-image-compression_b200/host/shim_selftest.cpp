// Self-checking twin of the reference's print-and-eyeball programs (src/binmat_test.cpp,
// src/patch_test.cpp) for the shim: exits non-zero on the first mismatch. Needs a GPU.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "GolombCoder.h"
#include "bsvd.h"
#include "eg.h"

#define CHECK(c) do { if (!(c)) { std::cerr << "FAILED line " << __LINE__ << ": " #c << std::endl; return 1; } } while (0)

static unsigned rnd(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

int main() {
  // --- binmat_test.cpp:29-50 known answers: BC, B*Bt, Ct*C
  binary_matrix B(3, 2), C(2, 3), D(3, 3);
  B.clear(); C.clear(); D.clear();
  B.set(0, 0, 1); B.set(1, 0, 1); B.set(0, 1, 1); B.set(2, 1, 1);
  C.set(0, 1, 1); C.set(1, 0, 1); C.set(1, 2, 1);
  const char* BC[3] = {"###", ".#.", "#.#"};
  const char* BBt[3] = {".##", "##.", "#.#"};
  const char* CtC[3] = {"#.#", ".#.", "#.#"};
  mul(B, false, C, false, D);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) CHECK(D.get(i, j) == (BC[i][j] == '#'));
  mul(B, false, B, true, D);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) CHECK(D.get(i, j) == (BBt[i][j] == '#'));
  mul(C, true, C, false, D);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) CHECK(D.get(i, j) == (CtC[i][j] == '#'));
  CHECK(D.weight() == 5 && D.row_weight(0) == 2 && D.col_weight(1) == 1);
  binary_matrix V = D.get_vectorized();
  CHECK(V.get_cols() == 9 && V.get(0, 0) && !V.get(0, 1) && V.get(0, 4));

  // --- patch_test.cpp: extract -> vectorise -> un-vectorise -> reassemble, host path vs kernels
  unsigned s = 7;
  const idx_t rows = 150, cols = 131;
  binary_matrix I(rows, cols);
  I.clear();
  for (idx_t i = 0; i < rows; ++i) for (idx_t j = 0; j < cols; ++j) if (rnd(s) % 5 == 0) I.set(i, j);
  for (idx_t W : {8ul, 12ul, 16ul}) {
    const idx_t Ny = (W - 1 + rows) / W, Nx = (W - 1 + cols) / W;
    binary_matrix Xh(Nx * Ny, W * W), Xd(Nx * Ny, W * W), P(W, W), Vv(1, W * W), I2(rows, cols);
    idx_t li = 0;
    for (idx_t i = 0; i < Ny; i++)
      for (idx_t j = 0; j < Nx; j++, li++) {  // src/bsvd_test.cpp:92-98 on the host
        I.copy_submatrix_to(i * W, (i + 1) * W, j * W, (j + 1) * W, P);
        P.copy_vectorized_to(Vv);
        Xh.set_row(li, Vv);
        binary_matrix P2(W, W);
        P2.set_vectorized(Vv);
        CHECK(dist(P, P2) == 0);  // "Difference after vect!" (patch_test.cpp:48-50)
        P2.destroy();
      }
    extract_patches(I, W, Xd);
    CHECK(dist(Xh, Xd) == 0);
    assemble_patches(Xd, W, I2);
    CHECK(dist(I, I2) == 0);
    Xh.destroy(); Xd.destroy(); P.destroy(); Vv.destroy(); I2.destroy();
  }

  // --- GolombCoder: SURVEY 8c trace, and serial writer == device encoder
  {
    const unsigned xs[10] = {0, 0, 5, 100, 3, 0, 0, 0, 1000, 2};
    const long bits[10] = {2, 1, 6, 52, 6, 6, 6, 5, 67, 8};
    GolombCoder gc;
    for (int t = 0; t < 10; ++t) { const long b0 = gc.bitcount; gc.codeSample(xs[t]); CHECK(gc.bitcount - b0 == bits[t]); }
    EGCoder ec;
    const int lens[7] = {5, 0, 3, 7, 0, 0, 12};
    const bool eols[7] = {false, false, true, false, false, true, false};
    const unsigned long eb[7] = {7, 1, 4, 8, 1, 1, 13};
    for (int t = 0; t < 7; ++t) { const unsigned long b0 = ec.bitcount; ec.codeRun(lens[t], eols[t]); CHECK(ec.bitcount - b0 == eb[t]); }
  }
  {
    BinaryFileWriter w;
    GolombCoder gc(&w);
    unsigned run = 0;
    for (idx_t i = 0; i < rows; ++i) for (idx_t j = 0; j < cols; ++j) { if (I.get(i, j)) { gc.codeSample(run); run = 0; } else run++; }
    gc.codeSample(run);
    std::vector<uint8_t> bytes;
    std::vector<uint64_t> idx;
    unsigned long ns = 0;
    const unsigned long bc = golomb_encode(I, bytes, &idx, &ns);
    CHECK(bc == (unsigned long)gc.bitcount && bc == w.bits());
    CHECK(bytes == w.bytes());
    CHECK(ns == I.weight() + 1);
    binary_matrix I3(rows, cols);
    golomb_decode(bytes, bc, ns, idx, I3);
    CHECK(dist(I, I3) == 0);
    // serial decoder on the device stream
    BinaryFileReader r(bytes.data(), bc);
    GolombDecoder gd(&r);
    idx_t pos = 0;
    binary_matrix I4(rows, cols);
    I4.clear();
    for (unsigned long t = 0; t < ns; ++t) { pos += gd.decodeSample(); if (pos < rows * cols) { I4.set(pos / cols, pos % cols); pos++; } }
    CHECK(dist(I, I4) == 0);
    I3.destroy(); I4.destroy();
  }

  // --- plug points: the catalog runs the fit; fixed point; E == A*D xor X; ownership quirks
  learn_model_setup(0, 0, 0, 0, 0);
  {
    const idx_t W = 8, K = 16, Ny = (W - 1 + rows) / W, Nx = (W - 1 + cols) / W;
    binary_matrix X(Nx * Ny, W * W), Dd(K, W * W), A(Nx * Ny, K), E(Nx * Ny, W * W), E2(Nx * Ny, W * W);
    extract_patches(I, W, X);
    initialize_model(X, Dd, A);
    CHECK(A.weight() == 0);
    const idx_t iters = learn_model(X, E, Dd, A);
    CHECK(iters >= 1);
    CHECK(update_coefficients(E, Dd, A) == 0 && update_dictionary(E, Dd, A) == 0);
    mul(A, false, Dd, false, E2);  // host product of the shim ...
    add(E2, X, E2);
    CHECK(dist(E, E2) == 0);       // ... agrees with the device residual
    residual(X, A, Dd, E2);
    CHECK(dist(E, E2) == 0);
    binary_matrix T;
    T = E.get_copy();              // operator= takes over the temporary's storage (binmat.cpp:180-184)
    CHECK(dist(T, E) == 0);
    T.destroy();
    X.destroy(); Dd.destroy(); A.destroy(); E.destroy(); E2.destroy();
  }
  // --- MDL learners through the catalog (src/bsvd.cpp:1463-1660): D and A are resized to the selected model
  for (int lm : {4, 5}) {
    learn_model_setup(0, 0, 0, lm, 0);
    const idx_t W = 8, K = 6, Ny = (W - 1 + rows) / W, Nx = (W - 1 + cols) / W;
    binary_matrix X(Nx * Ny, W * W), Dd(K, W * W), A(Nx * Ny, K), E(Nx * Ny, W * W), E2(Nx * Ny, W * W);
    extract_patches(I, W, X);
    initialize_model(X, Dd, A);
    const idx_t bestL = learn_model(X, E, Dd, A);
    CHECK(Dd.get_rows() == A.get_cols());
    CHECK(Dd.get_rows() == 0 ? A.get_rows() == 0 : A.get_rows() == X.get_rows());  // the empty model is 0 x 0 (src/bsvd.cpp:1623-1629)
    CHECK(lm == 4 ? Dd.get_rows() >= K : Dd.get_rows() <= K);
    if (Dd.get_rows() > 0) {
      CHECK(bestL == model_codelength(E, Dd, A));
      CHECK(update_coefficients(E, Dd, A) == 0 && update_dictionary(E, Dd, A) == 0);
      residual(X, A, Dd, E2);
      CHECK(dist(E, E2) == 0);
      // the description length from host-side weights (the reference's own loop, src/bsvd.cpp:1449-1456)
      idx_t LE = universal_codelength(E.get_rows() * E.get_cols(), E.weight()), LD = 0, LA = 0;
      for (idx_t k = 0; k < Dd.get_rows(); k++) {
        LD += universal_codelength(Dd.get_cols(), Dd.row_weight(k));
        LA += universal_codelength(A.get_rows(), A.col_weight(k));
      }
      CHECK(bestL == LE + LD + LA);
    }
    X.destroy(); Dd.destroy(); A.destroy(); E.destroy(); E2.destroy();
  }
  learn_model_setup(0, 0, 0, 0, 0);
  std::cout << "shim selftest ok" << std::endl;
  return 0;
}
