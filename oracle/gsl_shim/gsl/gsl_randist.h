/* TEST INFRASTRUCTURE (oracle build recipe). Stand-in for <gsl/gsl_randist.h>; see gsl_rng.h. */
#ifndef BIC_GSL_SHIM_RANDIST_H
#define BIC_GSL_SHIM_RANDIST_H
#include "gsl_rng.h"
#ifdef __cplusplus
extern "C" {
#endif
/* GSL randist/bernoulli.c: u = gsl_rng_uniform(r); return u < p.  (reference: src/bsvd.cpp:393) */
static inline unsigned int gsl_ran_bernoulli(gsl_rng* r, double p) {
  return gsl_rng_uniform(r) < p ? 1u : 0u;
}
#ifdef __cplusplus
}
#endif
#endif
