/* oracle/shim -- TEST INFRASTRUCTURE ONLY (see gsl_rng.h). */
#ifndef ORACLE_SHIM_GSL_STATS_H
#define ORACLE_SHIM_GSL_STATS_H
#include <stddef.h>
double gsl_stats_mean(const double data[], size_t stride, size_t n);
double gsl_stats_variance_m(const double data[], size_t stride, size_t n, double mean);
#endif
