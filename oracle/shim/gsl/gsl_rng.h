/* oracle/shim/gsl/gsl_rng.h -- TEST INFRASTRUCTURE ONLY.
 * Minimal stand-in for GSL's RNG interface so that the unmodified reference
 * sources under /root/reference/Pf compile in an image without GSL.  None of
 * this is on the likelihood hot path. */
#ifndef ORACLE_SHIM_GSL_RNG_H
#define ORACLE_SHIM_GSL_RNG_H
#include <stddef.h>
typedef struct {
    const char *name;
    unsigned long max, min;
    size_t size;
} gsl_rng_type;
typedef struct {
    const gsl_rng_type *type;
    void *state;
} gsl_rng;
extern const gsl_rng_type *gsl_rng_default;
const gsl_rng_type *gsl_rng_env_setup(void);
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_free(gsl_rng *r);
void gsl_rng_set(const gsl_rng *r, unsigned long seed);
unsigned long gsl_rng_get(const gsl_rng *r);
double gsl_rng_uniform(const gsl_rng *r);
unsigned long gsl_rng_uniform_int(const gsl_rng *r, unsigned long n);
size_t gsl_rng_size(const gsl_rng *r);
void *gsl_rng_state(const gsl_rng *r);
#endif
