// Stand-in for the reference's bsvd.h (src/bsvd.h:1-136): the same five global plug points, the same
// typedefs, selector and seed variable. The catalog entries behind them run on the B200 through the
// C ABI (include/bic_b200.h). Only the reference's deterministic oracle configuration is provided
// (SURVEY 8c): the other catalog slots exit(-1) from learn_model_setup, like an out-of-range index.
#ifndef BIC_HOST_BSVD_H
#define BIC_HOST_BSVD_H

#include "binmat.h"

// src/bsvd.h:6-8 / src/bsvd.cpp:227-267
void initialize_model_neighbor(const binary_matrix& E, binary_matrix& D, binary_matrix& A);
// src/bsvd.h:39 / src/bsvd.cpp:463-527 (atoms strictly in order)
idx_t update_dictionary_steepest(binary_matrix& E, binary_matrix& D, binary_matrix& A);
// src/bsvd.h:41 / src/bsvd.cpp:528-729: per atom, the atom's vote alternates with a vote for its coefficient column
idx_t update_dictionary_proximus(binary_matrix& E, binary_matrix& D, binary_matrix& A);
// src/bsvd.h:40: the reference's OpenMP variant races (src/bsvd.cpp:770-787); here it is the same
// deterministic kernel as update_dictionary_steepest
idx_t update_dictionary_steepest_omp(binary_matrix& E, binary_matrix& D, binary_matrix& A);
// src/bsvd.h:57-59 / src/bsvd.cpp:399-460, :1029-1107 (identical semantics, one kernel)
idx_t update_coefficients_basic(binary_matrix& E, const binary_matrix& D, binary_matrix& A);
idx_t update_coefficients_omp(binary_matrix& E, const binary_matrix& D, binary_matrix& A);
// src/bsvd.h:71-74 / src/bsvd.cpp:1215-1244
idx_t learn_model_traditional(binary_matrix& X, binary_matrix& E, binary_matrix& D, binary_matrix& A);

// src/bsvd.h:76-98 / src/bsvd.cpp:1245-1434: the role-switched learners (the fit's updates alternate with the same updates
// on the transposed problem)
idx_t learn_model_alter1(binary_matrix& X, binary_matrix& E, binary_matrix& D, binary_matrix& A);
idx_t learn_model_alter2(binary_matrix& X, binary_matrix& E, binary_matrix& D, binary_matrix& A);
idx_t learn_model_alter3(binary_matrix& X, binary_matrix& E, binary_matrix& D, binary_matrix& A);
// src/bsvd.h:100-118 / src/bsvd.cpp:1463-1717: MDL model selection around the fit. They resize D and A (destroy + allocate)
// to the selected number of atoms and return the best description length; an empty model leaves D and A 0 x 0.
idx_t learn_model_mdl_forward_selection(binary_matrix& X, binary_matrix& E, binary_matrix& D, binary_matrix& A);
idx_t learn_model_mdl_backward_selection(binary_matrix& X, binary_matrix& E, binary_matrix& D, binary_matrix& A);
idx_t learn_model_mdl_full_search(binary_matrix& X, binary_matrix& E, binary_matrix& D, binary_matrix& A);
// src/bsvd.cpp:1438-1461 and src/coding.h / coding.cpp:24-32
idx_t model_codelength(const binary_matrix& E, const binary_matrix& D, const binary_matrix& A);
double universal_codelength(const unsigned n, const unsigned r);

typedef void (*mi_algorithm_t)(const binary_matrix& E, binary_matrix& D, binary_matrix& A);
typedef idx_t (*cu_algorithm_t)(binary_matrix& E, const binary_matrix& D, binary_matrix& A);
typedef idx_t (*du_algorithm_t)(binary_matrix& E, binary_matrix& D, binary_matrix& A);
typedef idx_t (*ml_algorithm_t)(binary_matrix& X, binary_matrix& E, binary_matrix& D, binary_matrix& A);

extern mi_algorithm_t initialize_model;
extern cu_algorithm_t update_coefficients;
extern du_algorithm_t update_dictionary;
extern ml_algorithm_t learn_model;
extern ml_algorithm_t learn_model_inner;

extern const char* mi_algorithm_names[];
extern const char* cu_algorithm_names[];
extern const char* du_algorithm_names[];
extern const char* lm_algorithm_names[];

extern long random_seed;

void learn_model_setup(int mi_algo, int cu_algo, int du_algo, int lm_algo, int lmi_algo);

// ---- B200 extensions -----------------------------------------------------------------------------
// patch extraction loop of the driver (src/bsvd_test.cpp:80-99) and its inverse (:128-139) as kernels
void extract_patches(const binary_matrix& I, idx_t W, binary_matrix& X);
void assemble_patches(const binary_matrix& X, idx_t W, binary_matrix& I);
// E = A*D xor X on the device (mul + add, src/bsvd.cpp:1219-1220)
void residual(const binary_matrix& X, const binary_matrix& A, const binary_matrix& D, binary_matrix& E);

#endif
