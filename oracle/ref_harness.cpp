/*
 * TEST INFRASTRUCTURE -- not part of the product.
 *
 * C-callable harness around the UNMODIFIED reference objects (compiled from where they lie
 * under /root/reference/src by oracle/Makefile into oracle/_ref/libbic_ref.so). It lets the
 * Python tests and bench.py's `--impl reference` / cpu_baseline legs drive the reference's
 * own functions on plain buffers in the reference's word layout (row-major uint64 words,
 * bit j of a row at MSB >> (j % 64), rows padded to whole words: src/binmat.h:114-116,
 * src/binmat.cpp:140-149).
 *
 * Nothing in here restates the algorithm: every wrapper calls straight into the reference
 * (the file:line of each callee is given next to the wrapper). The only non-reference code
 * is buffer marshalling and wall-clock timers (the reference has none).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <bitset>
#include <cmath>
#include <cassert>
#include <chrono>
#include <algorithm>
#include <omp.h>

/* marshalling needs the raw word pointer, which the reference keeps private (binmat.h:183-231).
 * Standard headers are included above so this only affects the reference's own classes. */
#define private public
#define protected public
#include "binmat.h"
#include "bsvd.h"
#include "GolombCoder.h"
#include "eg.h"
#include "coding.h"
#undef private
#undef protected
#include "gsl/gsl_rng.h"

typedef unsigned long u64;

/* defined in bsvd.cpp:1438 but not declared in bsvd.h */
idx_t model_codelength(const binary_matrix& E, const binary_matrix& D, const binary_matrix& A);

static void load(binary_matrix& M, const u64* w) {
  if (M.data_blocks) std::memcpy(M.data, w, sizeof(u64) * M.data_blocks);
}
static void store(const binary_matrix& M, u64* w) {
  if (M.data_blocks) std::memcpy(w, M.data, sizeof(u64) * M.data_blocks);
}

static bool g_setup_done = false;
static void ensure_setup() {
  if (g_setup_done) return;
  /* bsvd.cpp:79-96 -- assigns the five global plug points. 0,0,0,0,0 = neighbor init,
   * OpenMP coefficient update, SERIAL steepest dictionary update, traditional learner:
   * the only deterministic combination. Its "Using ..." chatter goes to a sink. */
  std::streambuf* old = std::cout.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf());
  learn_model_setup(0, 0, 0, 0, 0);
  std::cout.rdbuf(old);
  g_setup_done = true;
}

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

extern "C" {

int ref_max_threads() { return omp_get_max_threads(); }
void ref_set_threads(int t) { omp_set_num_threads(t); }

/* Re-seed the reference's function-static rand48 (bsvd.cpp:8-15) so successive runs in one
 * process behave like fresh processes started with `-r seed` (bsvd_test.cpp:39). */
void ref_reseed(long seed) {
  random_seed = seed;
  gsl_shim_reseed_all((unsigned long)seed);
}

/* Patch extraction exactly as the driver does it: bsvd_test.cpp:80-99, using
 * copy_submatrix_to (binmat.cpp:267-298), copy_vectorized_to (:306-320), set_row (:362-371). */
void ref_extract_patches(const u64* I_words, u64 rows, u64 cols, u64 W, u64* X_words) {
  binary_matrix I(rows, cols);
  load(I, I_words);
  const idx_t Ny = (W - 1 + rows) / W;
  const idx_t Nx = (W - 1 + cols) / W;
  binary_matrix X;
  X.allocate(Nx * Ny, W * W);
  idx_t li = 0;
  binary_matrix P(W, W), V(1, W * W);
  for (idx_t i = 0; i < Ny; i++) {
    for (idx_t j = 0; j < Nx; j++, li++) {
      I.copy_submatrix_to(i * W, (i + 1) * W, j * W, (j + 1) * W, P);
      P.copy_vectorized_to(V);
      X.set_row(li, V);
    }
  }
  store(X, X_words);
  P.destroy(); V.destroy(); X.destroy(); I.destroy();
}

/* initialize_model_neighbor: bsvd.cpp:227-267. Returns the number of RNG draws made; the
 * draws (accepted and rejected, in order) are written to `draws` (capacity `cap`). */
u64 ref_init_neighbor(const u64* X_words, u64 n, u64 m, u64 p, u64* D_words, u64* A_words,
                      long seed, int reseed, u64* draws, u64 cap) {
  ensure_setup();
  binary_matrix X(n, m), D(p, m), A(n, p);
  load(X, X_words);
  /* make sure the static generator exists before re-seeding it */
  if (reseed) { random_seed = seed; ref_reseed(seed); }
  gsl_shim_set_log(draws, draws ? cap : 0);
  initialize_model_neighbor(X, D, A);
  const u64 ndraws = gsl_shim_log_len;
  gsl_shim_set_log(0, 0);
  store(D, D_words);
  store(A, A_words);
  X.destroy(); D.destroy(); A.destroy();
  return ndraws;
}

/* update_coefficients_omp: bsvd.cpp:1029-1107 (E, A updated in place). */
u64 ref_update_coefficients(u64* E_words, const u64* D_words, u64* A_words, u64 n, u64 m, u64 p) {
  binary_matrix E(n, m), D(p, m), A(n, p);
  load(E, E_words); load(D, D_words); load(A, A_words);
  const u64 changed = update_coefficients_omp(E, D, A);
  store(E, E_words); store(A, A_words);
  E.destroy(); D.destroy(); A.destroy();
  return changed;
}

/* update_coefficients_basic: bsvd.cpp:399-460 (serial twin; prints "cu/basic"). */
u64 ref_update_coefficients_basic(u64* E_words, const u64* D_words, u64* A_words, u64 n, u64 m, u64 p) {
  binary_matrix E(n, m), D(p, m), A(n, p);
  load(E, E_words); load(D, D_words); load(A, A_words);
  std::streambuf* old = std::cout.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf());
  const u64 changed = update_coefficients_basic(E, D, A);
  std::cout.rdbuf(old);
  store(E, E_words); store(A, A_words);
  E.destroy(); D.destroy(); A.destroy();
  return changed;
}

/* update_dictionary_steepest: bsvd.cpp:463-527 (E, D updated in place). */
u64 ref_update_dictionary(u64* E_words, u64* D_words, const u64* A_words, u64 n, u64 m, u64 p) {
  binary_matrix E(n, m), D(p, m), A(n, p);
  load(E, E_words); load(D, D_words); load(A, A_words);
  const u64 changed = update_dictionary_steepest(E, D, A);
  store(E, E_words); store(D, D_words);
  E.destroy(); D.destroy(); A.destroy();
  return changed;
}

/* E = A*D xor X : mul (binmat.cpp:606-616 -> mul_AB :516-543) then add (:463-478),
 * as learn_model_traditional (bsvd.cpp:1219-1220) and bsvd_test.cpp:153-154 do. */
void ref_residual(const u64* X_words, const u64* A_words, const u64* D_words, u64* E_words,
                  u64 n, u64 m, u64 p) {
  binary_matrix X(n, m), D(p, m), A(n, p), E(n, m);
  load(X, X_words); load(D, D_words); load(A, A_words);
  mul(A, false, D, false, E);
  add(E, X, E);
  store(E, E_words);
  E.destroy(); D.destroy(); A.destroy(); X.destroy();
}

/* learn_model_traditional: bsvd.cpp:1215-1244. D, A in/out; E out. Returns iterations. */
u64 ref_learn_traditional(const u64* X_words, u64* E_words, u64* D_words, u64* A_words,
                          u64 n, u64 m, u64 p) {
  ensure_setup();
  binary_matrix X(n, m), D(p, m), A(n, p), E(n, m);
  load(X, X_words); load(D, D_words); load(A, A_words);
  const u64 iters = learn_model_traditional(X, E, D, A);
  store(E, E_words); store(D, D_words); store(A, A_words);
  E.destroy(); D.destroy(); A.destroy(); X.destroy();
  return iters;
}

/* binary_matrix::weight: binmat.cpp:57-67 */
u64 ref_weight(const u64* words, u64 rows, u64 cols) {
  binary_matrix M(rows, cols);
  load(M, words);
  const u64 w = M.weight();
  M.destroy();
  return w;
}

/* GolombCoder::codeSample: GolombCoder.cpp:29-34 (+ binaryEncode :13-27). Records the k used
 * for each sample and the bits it added. Returns the final bitcount. */
long ref_golomb(const unsigned* samples, u64 n, unsigned* k_used, long* bits_added) {
  GolombCoder gc;
  for (u64 t = 0; t < n; ++t) {
    const long before = gc.bitcount;
    if (k_used) k_used[t] = gc.k;
    gc.codeSample(samples[t]);
    if (bits_added) bits_added[t] = gc.bitcount - before;
  }
  return gc.bitcount;
}

/* The reference's GolombCoder (GolombCoder.cpp:29-34) fed with the zero-run lengths of M read
 * row-major through binary_matrix::get (binmat.h:114-116); a virtual one closes the last run.
 * This is the serial coding baseline (single core: the state machine cannot be threaded). */
long ref_golomb_matrix(const u64* words, u64 rows, u64 cols) {
  binary_matrix M(rows, cols);
  load(M, words);
  GolombCoder gc;
  unsigned run = 0;
  for (u64 i = 0; i < rows; ++i)
    for (u64 j = 0; j < cols; ++j) {
      if (M.get(i, j)) { gc.codeSample(run); run = 0; }
      else run++;
    }
  gc.codeSample(run);
  M.destroy();
  return gc.bitcount;
}

/* EGCoder::codeRun: eg.cpp:20-37. Returns the final bitcount. */
u64 ref_eg(const int* lens, const unsigned char* eols, u64 n, u64* bits_added) {
  EGCoder ec;
  for (u64 t = 0; t < n; ++t) {
    const u64 before = ec.bitcount;
    ec.codeRun(lens[t], eols[t] != 0);
    if (bits_added) bits_added[t] = ec.bitcount - before;
  }
  return ec.bitcount;
}

/* Whole fit as bsvd_test.cpp:56-155 runs it in image mode (-I 1), minus file I/O, with
 * wall-clock timers around each phase (the reference has none).
 * times[0]=extract, [1]=init, [2]=coefficient updates (sum), [3]=dictionary updates (sum),
 * [4]=initial mul+add, [5]=total. D/A/E outputs optional (may be null). */
u64 ref_fit_timed(const u64* I_words, u64 rows, u64 cols, u64 W, u64 K, long seed,
                  double* times, u64* D_words, u64* A_words, u64* E_words) {
  ensure_setup();
  const double t_begin = now_s();
  binary_matrix I(rows, cols);
  load(I, I_words);
  const idx_t Ny = (W - 1 + rows) / W;
  const idx_t Nx = (W - 1 + cols) / W;
  const idx_t M = W * W, N = Nx * Ny;
  double t0 = now_s();
  binary_matrix X;
  X.allocate(N, M);
  {
    idx_t li = 0;
    binary_matrix P(W, W), V(1, W * W);
    for (idx_t i = 0; i < Ny; i++) {
      for (idx_t j = 0; j < Nx; j++, li++) {
        I.copy_submatrix_to(i * W, (i + 1) * W, j * W, (j + 1) * W, P);
        P.copy_vectorized_to(V);
        X.set_row(li, V);
      }
    }
    P.destroy(); V.destroy();
  }
  times[0] = now_s() - t0;
  binary_matrix D(K, M), A(N, K), E(N, M);
  ref_reseed(seed);
  t0 = now_s();
  initialize_model(X, D, A);
  times[1] = now_s() - t0;
  /* learn_model_traditional (bsvd.cpp:1215-1244) unrolled only to put timers around the two
   * plug points; the calls and their order are the reference's. */
  t0 = now_s();
  mul(A, false, D, false, E);
  add(E, X, E);
  times[4] = now_s() - t0;
  times[2] = times[3] = 0.0;
  idx_t changed = 1, iter = 0;
  while (changed > 0) {
    iter++;
    t0 = now_s();
    idx_t cc = update_coefficients(E, D, A);
    times[2] += now_s() - t0;
    t0 = now_s();
    changed = cc + update_dictionary(E, D, A);
    times[3] += now_s() - t0;
  }
  times[5] = now_s() - t_begin;
  if (D_words) store(D, D_words);
  if (A_words) store(A, A_words);
  if (E_words) store(E, E_words);
  I.destroy(); X.destroy(); D.destroy(); A.destroy(); E.destroy();
  return iter;
}

/* universal_codelength: coding.cpp:24-32 */
double ref_universal_codelength(unsigned n, unsigned r) { return universal_codelength(n, r); }

/* model_codelength: bsvd.cpp:1438-1461 */
u64 ref_model_codelength(const u64* E_words, const u64* D_words, const u64* A_words, u64 n, u64 m, u64 p) {
  binary_matrix E(n, m), D(p, m), A(n, p);
  load(E, E_words); load(D, D_words); load(A, A_words);
  const u64 L = model_codelength(E, D, A);
  E.destroy(); D.destroy(); A.destroy();
  return L;
}

/* The MDL learners: learn_model_mdl_forward_selection bsvd.cpp:1463-1546 (lm = 4), _backward_selection :1548-1660
 * (lm = 5), _full_search :1662-1717 (lm = 6), with the reference's default plug points (ensure_setup). They resize
 * D and A, so the result stays in a handle until the caller has sized its buffers. */
struct ref_mdl_result { binary_matrix D, A; u64 bestL; };

void* ref_learn_mdl(int lm, const u64* X_words, u64* E_words, const u64* D_words, const u64* A_words,
                    u64 n, u64 m, u64 p, long seed, int reseed) {
  ensure_setup();
  binary_matrix X(n, m), E(n, m);
  load(X, X_words);
  if (E_words) load(E, E_words);
  ref_mdl_result* r = new ref_mdl_result;
  r->D.allocate(p, m);
  r->A.allocate(n, p);
  if (D_words) load(r->D, D_words); else r->D.clear();
  if (A_words) load(r->A, A_words); else r->A.clear();
  if (reseed) ref_reseed(seed);
  std::streambuf* old = std::cout.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf());
  if (lm == 4) r->bestL = learn_model_mdl_forward_selection(X, E, r->D, r->A);
  else if (lm == 5) r->bestL = learn_model_mdl_backward_selection(X, E, r->D, r->A);
  else r->bestL = learn_model_mdl_full_search(X, E, r->D, r->A);
  std::cout.rdbuf(old);
  store(E, E_words);
  X.destroy(); E.destroy();
  return r;
}
void ref_mdl_result_info(void* h, u64* p, u64* bestL) {
  ref_mdl_result* r = (ref_mdl_result*)h;
  *p = r->D.get_rows();
  *bestL = r->bestL;
}
void ref_mdl_result_copy(void* h, u64* D_words, u64* A_words) {
  ref_mdl_result* r = (ref_mdl_result*)h;
  if (r->D.get_rows()) { store(r->D, D_words); store(r->A, A_words); }
}
void ref_mdl_result_free(void* h) {
  ref_mdl_result* r = (ref_mdl_result*)h;
  if (r->D.data) r->D.destroy();
  if (r->A.data) r->A.destroy();
  delete r;
}

/* learn_model_alter1 / 2 / 3: bsvd.cpp:1245-1311, :1314-1388, :1391-1434 (default plug points). D, A in/out; E out. */
u64 ref_learn_alter(int variant, const u64* X_words, u64* E_words, u64* D_words, u64* A_words, u64 n, u64 m, u64 p) {
  ensure_setup();
  binary_matrix X(n, m), D(p, m), A(n, p), E(n, m);
  load(X, X_words); load(D, D_words); load(A, A_words);
  E.clear();
  u64 iters;
  if (variant == 1) iters = learn_model_alter1(X, E, D, A);
  else if (variant == 2) iters = learn_model_alter2(X, E, D, A);
  else iters = learn_model_alter3(X, E, D, A);
  store(E, E_words); store(D, D_words); store(A, A_words);
  E.destroy(); D.destroy(); A.destroy(); X.destroy();
  return iters;
}

/* update_dictionary_proximus: bsvd.cpp:528-729 (E, D, A updated in place). */
u64 ref_update_dictionary_proximus(u64* E_words, u64* D_words, u64* A_words, u64 n, u64 m, u64 p) {
  binary_matrix E(n, m), D(p, m), A(n, p);
  load(E, E_words); load(D, D_words); load(A, A_words);
  const u64 changed = update_dictionary_proximus(E, D, A);
  store(E, E_words); store(D, D_words); store(A, A_words);
  E.destroy(); D.destroy(); A.destroy();
  return changed;
}

} /* extern "C" */
