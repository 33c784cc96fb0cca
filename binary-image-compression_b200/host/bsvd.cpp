// The reference's plug-in catalog (src/bsvd.cpp:17-96) with B200 entries. Own code over the C ABI.
#include "bsvd.h"

#include <cstdlib>
#include <iostream>

#include "bic_b200.h"

static void ck(bic_status st, const char* what) {
  if (st == BIC_OK) return;
  std::cerr << "binary-image-compression_b200: " << what << ": " << bic_status_string(st) << " ("
            << bic_ctx_last_error(bic_host_context()) << ")" << std::endl;
  std::abort();
}

// The reference keeps one function-static rand48 seeded from random_seed on first use and advanced
// by every initialiser call (src/bsvd.cpp:8-15). Same lifetime here.
long random_seed = 34503498;  // src/bsvd.cpp:23
static uint64_t* get_rng() {
  static uint64_t state;
  static bool seeded = false;
  if (!seeded) { bic_rand48_seed(&state, (unsigned long)random_seed); seeded = true; }
  return &state;
}

void initialize_model_neighbor(const binary_matrix& E, binary_matrix& D, binary_matrix& A) {
  ck(bic_initialize_model_neighbor(bic_host_context(), E.device(), D.device(), A.device(), get_rng()), "initialize_model_neighbor");
  D.device_written();
  A.device_written();
}

idx_t update_coefficients_omp(binary_matrix& E, const binary_matrix& D, binary_matrix& A) {
  uint64_t changed = 0;
  ck(bic_update_coefficients(bic_host_context(), E.device(), D.device(), A.device(), &changed), "update_coefficients");
  E.device_written();
  A.device_written();
  return changed;
}
idx_t update_coefficients_basic(binary_matrix& E, const binary_matrix& D, binary_matrix& A) {
  return update_coefficients_omp(E, D, A);
}

idx_t update_dictionary_steepest(binary_matrix& E, binary_matrix& D, binary_matrix& A) {
  uint64_t changed = 0;
  ck(bic_update_dictionary_steepest(bic_host_context(), E.device(), D.device(), A.device(), &changed), "update_dictionary");
  E.device_written();
  D.device_written();
  return changed;
}
idx_t update_dictionary_steepest_omp(binary_matrix& E, binary_matrix& D, binary_matrix& A) {
  return update_dictionary_steepest(E, D, A);
}
idx_t update_dictionary_proximus(binary_matrix& E, binary_matrix& D, binary_matrix& A) {
  uint64_t changed = 0;
  ck(bic_update_dictionary_proximus(bic_host_context(), E.device(), D.device(), A.device(), &changed), "update_dictionary_proximus");
  E.device_written();
  D.device_written();
  A.device_written();
  return changed;
}

void residual(const binary_matrix& X, const binary_matrix& A, const binary_matrix& D, binary_matrix& E) {
  ck(bic_residual(bic_host_context(), X.device(), A.device(), D.device(), E.device()), "residual");
  E.device_written();
}

void extract_patches(const binary_matrix& I, idx_t W, binary_matrix& X) {
  ck(bic_extract_patches(bic_host_context(), I.device(), W, X.device()), "extract_patches");
  X.device_written();
}
void assemble_patches(const binary_matrix& X, idx_t W, binary_matrix& I) {
  ck(bic_assemble_patches(bic_host_context(), X.device(), W, I.device()), "assemble_patches");
  I.device_written();
}

mi_algorithm_t initialize_model = initialize_model_neighbor;
cu_algorithm_t update_coefficients = update_coefficients_omp;
du_algorithm_t update_dictionary = update_dictionary_steepest;  // the reference's default is a null pointer (src/bsvd.cpp:19)
ml_algorithm_t learn_model = learn_model_traditional;
ml_algorithm_t learn_model_inner = learn_model_traditional;

idx_t learn_model_traditional(binary_matrix& X, binary_matrix& E, binary_matrix& D, binary_matrix& A) {
  if (update_coefficients == update_coefficients_omp || update_coefficients == update_coefficients_basic) {
    if (update_dictionary == update_dictionary_steepest || update_dictionary == update_dictionary_steepest_omp) {
      // both plug points are the device kernels: run the whole loop next to the data
      uint64_t iters = 0;
      ck(bic_learn_model_traditional(bic_host_context(), X.device(), E.device(), D.device(), A.device(), &iters, nullptr, 0),
         "learn_model_traditional");
      E.device_written(); D.device_written(); A.device_written();
      return iters;
    }
  }
  // a caller plugged its own update in: the reference's loop over the pointers (src/bsvd.cpp:1219-1243)
  residual(X, A, D, E);
  idx_t changed = 1, iter = 0;
  while (changed > 0) {
    iter++;
    const idx_t changed_coefs = update_coefficients(E, D, A);
    changed = changed_coefs + update_dictionary(E, D, A);
  }
  return iter;
}

// ---- role-switched learners (src/bsvd.cpp:1245-1434): loops, transposes and updates all run in the library
static idx_t learn_alter(int variant, binary_matrix& X, binary_matrix& E, binary_matrix& D, binary_matrix& A, const char* what) {
  const bool cu_ok = update_coefficients == update_coefficients_omp || update_coefficients == update_coefficients_basic;
  const bool du_ok = update_dictionary == update_dictionary_steepest || update_dictionary == update_dictionary_steepest_omp;
  if (!cu_ok || !du_ok) {
    std::cerr << what << ": the B200 build runs it with its own coefficient and dictionary updates only" << std::endl;
    std::exit(-1);
  }
  uint64_t iters = 0;
  ck(bic_learn_model_alter(bic_host_context(), variant, X.device(), E.device(), D.device(), A.device(), &iters), what);
  E.device_written(); D.device_written(); A.device_written();
  return iters;
}
idx_t learn_model_alter1(binary_matrix& X, binary_matrix& E, binary_matrix& D, binary_matrix& A) { return learn_alter(1, X, E, D, A, "learn_model_alter1"); }
idx_t learn_model_alter2(binary_matrix& X, binary_matrix& E, binary_matrix& D, binary_matrix& A) { return learn_alter(2, X, E, D, A, "learn_model_alter2"); }
idx_t learn_model_alter3(binary_matrix& X, binary_matrix& E, binary_matrix& D, binary_matrix& A) { return learn_alter(3, X, E, D, A, "learn_model_alter3"); }

// ---- MDL model selection (src/bsvd.cpp:1438-1717): the loops run in the library next to the data; D and A change size
static idx_t learn_mdl(int lm, binary_matrix& X, binary_matrix& E, binary_matrix& D, binary_matrix& A, const char* what) {
  if (initialize_model != initialize_model_neighbor || learn_model_inner != learn_model_traditional) {
    std::cerr << what << ": the B200 build runs it with the neighbor initialisation and the traditional inner learner only" << std::endl;
    std::exit(-1);
  }
  // the inner learner goes through the global update pointers (src/bsvd.cpp:1229,1235): -d 1 runs PROXIMUS inside the search
  const bool cu_ok = update_coefficients == update_coefficients_omp || update_coefficients == update_coefficients_basic;
  const bool du_steepest = update_dictionary == update_dictionary_steepest || update_dictionary == update_dictionary_steepest_omp;
  const bool du_proximus = update_dictionary == update_dictionary_proximus;
  if (!cu_ok || !(du_steepest || du_proximus)) {
    std::cerr << what << ": the B200 build runs it with its own coefficient and dictionary updates only" << std::endl;
    std::exit(-1);
  }
  bic_mat* x = X.device();
  bic_mat* e = E.device();
  bic_mat* d = D.release_device();
  bic_mat* a = A.release_device();
  uint64_t bestL = 0;
  ck(bic_ctx_set_option(bic_host_context(), "dict_update", du_proximus ? 1 : 0), what);
  const bic_status st = bic_learn_model_mdl(bic_host_context(), lm, x, e, &d, &a, get_rng(), &bestL);
  bic_ctx_set_option(bic_host_context(), "dict_update", 0);
  ck(st, what);
  E.device_written();
  D.adopt_device(d);
  A.adopt_device(a);
  return bestL;
}
idx_t learn_model_mdl_forward_selection(binary_matrix& X, binary_matrix& E, binary_matrix& D, binary_matrix& A) {
  return learn_mdl(4, X, E, D, A, "learn_model_mdl_forward_selection");
}
idx_t learn_model_mdl_backward_selection(binary_matrix& X, binary_matrix& E, binary_matrix& D, binary_matrix& A) {
  return learn_mdl(5, X, E, D, A, "learn_model_mdl_backward_selection");
}
idx_t learn_model_mdl_full_search(binary_matrix& X, binary_matrix& E, binary_matrix& D, binary_matrix& A) {
  return learn_mdl(6, X, E, D, A, "learn_model_mdl_full_search");
}
idx_t model_codelength(const binary_matrix& E, const binary_matrix& D, const binary_matrix& A) {
  uint64_t L = 0;
  const bool empty = D.get_rows() == 0;
  ck(bic_model_codelength(bic_host_context(), E.device(), empty ? nullptr : D.device(), empty ? nullptr : A.device(), &L), "model_codelength");
  return L;
}
double universal_codelength(const unsigned n, const unsigned r) { return bic_universal_codelength(n, r); }

// ---- catalog: same positions and names as src/bsvd.cpp:25-77; null = not provided by this build
static mi_algorithm_t mi_catalog[] = {initialize_model_neighbor, 0, 0, 0, 0, 0};
const char* mi_algorithm_names[] = {"Neighbor initialization", "Partition initialization", "Random centroids initialization",
                                    "Random centroids (in mod-2 algebra) initialization", "Graph growing initialization", 0};
static cu_algorithm_t cu_catalog[] = {update_coefficients_omp, update_coefficients_basic, 0, 0};
const char* cu_algorithm_names[] = {"OpenMP basic coefficients update", "Basic coefficients update",
                                    "Fast coefficients update (broken!)", 0};
static du_algorithm_t du_catalog[] = {update_dictionary_steepest, update_dictionary_proximus, update_dictionary_steepest_omp, 0, 0};
const char* du_algorithm_names[] = {"Steepest descent (a la MOD)  dictionary update", "Proximus-like dictionary update",
                                    "Steepest descent (a la MOD)  dictionary update (OMP)",
                                    "Proximus-like dictionary update (OMP)", 0};
static ml_algorithm_t lm_catalog[] = {learn_model_traditional, learn_model_alter1, learn_model_alter2, learn_model_alter3,
                                      learn_model_mdl_forward_selection,
                                      learn_model_mdl_backward_selection, learn_model_mdl_full_search, 0};
const char* lm_algorithm_names[] = {"Model learning by traditional alternate descent",
                                    "Role-switching learning 1: at each iteration, the role of A and D are switched",
                                    "Role-switched learning 2: after convergence, the role of A and D are switched and traditional model is applied again",
                                    "Role switched learning 3: like RS1 but only update_dictionary is applied (for use with Proximus",
                                    "MDL/forward selection", "MDO/backward selection", "MDL/full search"};

template <typename T>
static T pick(T* catalog, int idx, const char* what, const char* const* names) {
  if (!catalog[idx]) {
    std::cerr << what << " '" << names[idx] << "' is not provided by the B200 build (only the reference's deterministic "
              << "configuration -i 0 -c 0|1 -d 0|1|2 -l 0..6 -L 0 is)" << std::endl;
    std::exit(-1);
  }
  return catalog[idx];
}

void learn_model_setup(int mi_algo, int cu_algo, int du_algo, int lm_algo, int lmi_algo) {  // src/bsvd.cpp:79-96
  if (mi_algo < 0 || mi_algo > 4) { std::cerr << "Invalid model initialization algorithm (0-" << 4 << ')' << std::endl; exit(-1); }
  if (cu_algo < 0 || cu_algo > 2) { std::cerr << "Invalid coefficients update algorithm (0-" << 2 << ')' << std::endl; exit(-1); }
  if (du_algo < 0 || du_algo > 3) { std::cerr << "Invalid dictionary update algorithm (0-" << 3 << ')' << std::endl; exit(-1); }
  if (lm_algo < 0 || lm_algo > 6) { std::cerr << "Invalid model learning algorithm (0-" << 6 << ')' << std::endl; exit(-1); }
  if (lmi_algo < 0 || lmi_algo > 3) { std::cerr << "Invalid inner model learning algorithm (0-" << 3 << ')' << std::endl; exit(-1); }
  initialize_model = pick(mi_catalog, mi_algo, "initialisation", mi_algorithm_names);
  std::cout << "Using " << mi_algorithm_names[mi_algo] << std::endl;
  update_coefficients = pick(cu_catalog, cu_algo, "coefficients update", cu_algorithm_names);
  std::cout << "Using " << cu_algorithm_names[cu_algo] << std::endl;
  update_dictionary = pick(du_catalog, du_algo, "dictionary update", du_algorithm_names);
  std::cout << "Using " << du_algorithm_names[du_algo] << std::endl;
  learn_model = pick(lm_catalog, lm_algo, "learner", lm_algorithm_names);
  std::cout << "Using " << lm_algorithm_names[lm_algo] << " for outer learning loop." << std::endl;
  learn_model_inner = pick(lm_catalog, lmi_algo, "inner learner", lm_algorithm_names);
  std::cout << "Using " << lm_algorithm_names[lmi_algo] << " for inner learning." << std::endl;
}
