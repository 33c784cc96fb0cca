// Role-switched learners: learn_model_alter1 src/bsvd.cpp:1245-1311, learn_model_alter2 :1314-1388,
// learn_model_alter3 :1391-1434. They alternate the fit's two updates with the same two updates applied to the
// TRANSPOSED problem, where the roles are E' = Et (m x n), D' = At (p x n: p "atoms" as wide as the number of patches),
// A' = Dt (m x p). Everything stays on the device: binary_matrix::transpose_to (src/binmat.cpp:199-208) is the 32x32
// bit-tile transpose kernel of dict.cu, the transposed updates are the same entry points at their unusual shape
// (a few very wide rows: the warp-per-row coefficient kernel and the large-dictionary path of dict2.cu).
#include "bic_internal.cuh"

bic_status bic_k_update_coefficients(bic_ctx* c, bic_mat* E, const bic_mat* D, bic_mat* A, unsigned long long* d_changed);
bic_status bic_k_update_dictionary(bic_ctx* c, bic_mat* E, bic_mat* D, const bic_mat* A, unsigned long long* d_changed);
bic_status bic_k_transpose_A(bic_ctx* c, const bic_mat* A, uint32_t* AT, uint64_t wprN);

extern "C" bic_status bic_mat_transpose(bic_ctx* c, const bic_mat* src, bic_mat* dst) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !src || !dst || src->rows != dst->cols || src->cols != dst->rows) return BIC_ERR_INVALID;
  if (src->rows == 0 || src->cols == 0) return BIC_OK;
  if (src->wpr > 65535) return bic_fail(c, BIC_ERR_UNSUPPORTED, "transpose: more than 2M columns");
  return bic_k_transpose_A(c, src, dst->d, dst->wpr);
}

// one call of each plug point with its count read back (the learners' loop conditions need them)
static bic_status coef(bic_ctx* c, bic_mat* E, bic_mat* D, bic_mat* A, uint64_t* changed) {
  return bic_update_coefficients(c, E, D, A, changed);
}
static bic_status dict(bic_ctx* c, bic_mat* E, bic_mat* D, bic_mat* A, uint64_t* changed) {
  return bic_update_dictionary_steepest(c, E, D, A, changed);
}

extern "C" bic_status bic_learn_model_alter(bic_ctx* c, int variant, const bic_mat* X, bic_mat* E, bic_mat* D, bic_mat* A,
                                            uint64_t* iterations) {
  if (c) cudaSetDevice(c->device);  // the calling thread may be new to this device
  if (!c || !X || !E || !D || !A || variant < 1 || variant > 3) return BIC_ERR_INVALID;
  if (X->rows != E->rows || X->cols != E->cols || D->cols != E->cols || A->rows != E->rows || A->cols != D->rows)
    return bic_fail(c, BIC_ERR_INVALID, "learn_model_alter: shapes must be X, E n x m, D p x m, A n x p");
  const uint64_t n = E->rows, m = E->cols, p = D->rows;
  BIC_TRY(bic_residual(c, X, A, D, E));                               // :1254-1255 / :1323-1324 / :1399-1400
  bic_mat *Dt = nullptr, *At = nullptr, *Et = nullptr;
  BIC_TRY(bic_mat_create(c, m, p, &Dt));
  BIC_TRY(bic_mat_create(c, p, n, &At));
  BIC_TRY(bic_mat_create(c, m, n, &Et));
  bic_status st = BIC_OK;
  uint64_t iter = 0, cc = 0, ca = 0;
#define ALT(expr) if ((st = (expr)) != BIC_OK) break
#define TO_T()   ALT(bic_mat_transpose(c, A, At)); ALT(bic_mat_transpose(c, D, Dt)); ALT(bic_mat_transpose(c, E, Et))
#define FROM_T() ALT(bic_mat_transpose(c, At, A)); ALT(bic_mat_transpose(c, Dt, D)); ALT(bic_mat_transpose(c, Et, E))
  if (variant == 1) {                                                 // :1264-1307
    uint64_t changed = 1;
    while (changed > 0) {
      iter++;
      ALT(coef(c, E, D, A, &cc));
      ALT(dict(c, E, D, A, &ca));
      TO_T();
      ALT(coef(c, Et, At, Dt, &cc));
      ALT(dict(c, Et, At, Dt, &ca));
      changed = ca;                                                   // :1297: only the transposed atom count drives the loop
      FROM_T();
    }
  } else if (variant == 2) {                                          // :1331-1383
    uint64_t changed = 1, outer_changed = 1;
    while (outer_changed > 0 && st == BIC_OK) {
      outer_changed = 0;
      while (changed > 0) {
        iter++;
        ALT(coef(c, E, D, A, &cc));
        ALT(dict(c, E, D, A, &ca));
        changed = cc + ca;
        outer_changed += changed;
      }
      if (st != BIC_OK) break;
      TO_T();
      changed = 1;
      iter = 0;                                                       // :1361
      while (changed > 0) {
        iter++;
        ALT(coef(c, Et, At, Dt, &cc));
        ALT(dict(c, Et, At, Dt, &ca));
        changed = cc + ca;
        outer_changed += changed;
      }
      if (st != BIC_OK) break;
      FROM_T();
    }
  } else {                                                            // :1407-1429
    uint64_t changed = p + 1;
    while (changed > 0) {
      iter++;
      TO_T();
      ALT(dict(c, Et, At, Dt, &ca));
      FROM_T();
      ALT(dict(c, E, D, A, &ca));
      changed = ca;                                                   // :1423 overwrites the transposed count
    }
  }
#undef ALT
#undef TO_T
#undef FROM_T
  bic_mat_destroy(c, Dt);
  bic_mat_destroy(c, At);
  bic_mat_destroy(c, Et);
  if (iterations) *iterations = iter;
  return st;
}
