/*
 * TEST INFRASTRUCTURE (oracle build recipe) -- not part of the product.
 *
 * Header-only stand-in for the handful of GSL symbols the reference uses
 * (GSL is an un-vendored, un-pinned dependency of /root/reference: src/Makefile:17,20
 * links -lgsl -lgslcblas; it is not installed in this image).
 *
 * Call sites in the reference that this serves:
 *   src/bsvd.cpp:8-15   gsl_rng_alloc(gsl_rng_rand48), gsl_rng_set(rng, random_seed)
 *   src/bsvd.cpp:241    gsl_rng_uniform_int(rng, n)     (also :117,:151,:312,:354)
 *   src/bsvd.cpp:393    gsl_ran_bernoulli(rng, 0.5)     (unreachable helper)
 *
 * Algorithm restated from GSL's published rng/rand48.c and rng/rng.c:
 *   rand48: 48-bit LCG  x <- 0x5DEECE66D * x + 0xB  (mod 2^48), kept as three 16-bit limbs;
 *           seed s != 0 -> limbs (0x330E, s & 0xFFFF, (s >> 16) & 0xFFFF),
 *           seed s == 0 -> limbs (0x330E, 0xABCD, 0x1234); output = top 32 bits; range [0, 2^32-1].
 *   gsl_rng_uniform_int(r, n): scale = range / n; do k = get() / scale; while (k >= n).
 * PARITY NOTE: real GSL cannot be run offline here, so the pivot sequence is pinned against
 * glibc's srand48/mrand48 (same LCG) in tests/test_oracle_cpu.py, not against libgsl itself.
 *
 * Extras (not GSL): a registry so the harness can re-seed the reference's function-static
 * generator between runs, and a draw log so tests can recover the pivots the reference drew.
 */
#ifndef BIC_GSL_SHIM_RNG_H
#define BIC_GSL_SHIM_RNG_H

#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int id; } gsl_rng_type;

typedef struct gsl_rng_s {
  unsigned short x0, x1, x2;
  struct gsl_rng_s* next_registered;
} gsl_rng;

static const gsl_rng_type gsl_shim_rand48_type = {48};
static const gsl_rng_type* const gsl_rng_rand48 = &gsl_shim_rand48_type;

/* shared across translation units of one link (weak, so header-only is fine) */
__attribute__((weak)) gsl_rng* gsl_shim_registry = 0;
__attribute__((weak)) unsigned long* gsl_shim_log = 0;
__attribute__((weak)) unsigned long gsl_shim_log_len = 0;
__attribute__((weak)) unsigned long gsl_shim_log_cap = 0;

static inline void gsl_rng_set(gsl_rng* r, unsigned long s) {
  if (s == 0) {
    r->x0 = 0x330E; r->x1 = 0xABCD; r->x2 = 0x1234;
  } else {
    r->x0 = 0x330E;
    r->x1 = (unsigned short)(s & 0xFFFF);
    r->x2 = (unsigned short)((s >> 16) & 0xFFFF);
  }
}

static inline gsl_rng* gsl_rng_alloc(const gsl_rng_type* t) {
  (void)t;
  gsl_rng* r = (gsl_rng*)malloc(sizeof(gsl_rng));
  gsl_rng_set(r, 0);
  r->next_registered = gsl_shim_registry;
  gsl_shim_registry = r;
  return r;
}

static inline void gsl_rng_free(gsl_rng* r) { (void)r; /* registry keeps it */ }

static inline unsigned long gsl_rng_get(gsl_rng* r) {
  const unsigned long a0 = 0xE66D, a1 = 0xDEEC, a2 = 0x0005, c0 = 0x000B;
  const unsigned long x0 = r->x0, x1 = r->x1, x2 = r->x2;
  unsigned long a;
  a = a0 * x0 + c0;
  r->x0 = (unsigned short)(a & 0xFFFF);
  a >>= 16;
  a += a0 * x1 + a1 * x0;
  r->x1 = (unsigned short)(a & 0xFFFF);
  a >>= 16;
  a += a0 * x2 + a1 * x1 + a2 * x0;
  r->x2 = (unsigned short)(a & 0xFFFF);
  return ((unsigned long)r->x2 << 16) + (unsigned long)r->x1;
}

static inline double gsl_rng_uniform(gsl_rng* r) {
  return (double)gsl_rng_get(r) / 4294967296.0;
}

static inline unsigned long gsl_rng_uniform_int(gsl_rng* r, unsigned long n) {
  const unsigned long range = 0xFFFFFFFFUL;
  if (n > range || n == 0) return 0; /* GSL raises GSL_EINVAL and returns 0 */
  const unsigned long scale = range / n;
  unsigned long k;
  do {
    k = gsl_rng_get(r) / scale;
  } while (k >= n);
  if (gsl_shim_log && gsl_shim_log_len < gsl_shim_log_cap) gsl_shim_log[gsl_shim_log_len++] = k;
  return k;
}

/* --- harness-only helpers (not GSL API) --- */
static inline void gsl_shim_reseed_all(unsigned long s) {
  for (gsl_rng* r = gsl_shim_registry; r; r = r->next_registered) gsl_rng_set(r, s);
}
static inline void gsl_shim_set_log(unsigned long* buf, unsigned long cap) {
  gsl_shim_log = buf; gsl_shim_log_cap = cap; gsl_shim_log_len = 0;
}

#ifdef __cplusplus
}
#endif
#endif
