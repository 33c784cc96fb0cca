/* TEST INFRASTRUCTURE (oracle build recipe). Stand-in for <gsl/gsl_sf_gamma.h>.
 * Only gsl_sf_lnchoose is used (reference: src/coding.cpp:21, compress*_test.cpp); it is
 * OFF the hot path (float cost model of the MDL learners). ln C(n,m) via lgamma. */
#ifndef BIC_GSL_SHIM_SF_GAMMA_H
#define BIC_GSL_SHIM_SF_GAMMA_H
#include <math.h>
#ifdef __cplusplus
extern "C" {
#endif
static inline double gsl_sf_lnchoose(unsigned int n, unsigned int m) {
  if (m > n) return 0.0;
  if (m == n || m == 0) return 0.0;
  return lgamma((double)n + 1.0) - lgamma((double)m + 1.0) - lgamma((double)(n - m) + 1.0);
}
#ifdef __cplusplus
}
#endif
#endif
