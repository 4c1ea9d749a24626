/* oracle/shim -- TEST INFRASTRUCTURE ONLY (see gsl_rng.h). */
#ifndef ORACLE_SHIM_GSL_SF_GAMMA_H
#define ORACLE_SHIM_GSL_SF_GAMMA_H
double gsl_sf_lngamma(double x);
double gsl_sf_gamma(double x);
double gsl_sf_beta(double a, double b);
#endif
