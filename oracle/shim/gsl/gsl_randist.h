/* oracle/shim -- TEST INFRASTRUCTURE ONLY (see gsl_rng.h). */
#ifndef ORACLE_SHIM_GSL_RANDIST_H
#define ORACLE_SHIM_GSL_RANDIST_H
#include <gsl/gsl_rng.h>
double gsl_ran_chisq_pdf(double x, double nu);
double gsl_ran_gamma(const gsl_rng *r, double a, double b);
double gsl_ran_gamma_pdf(double x, double a, double b);
double gsl_ran_exponential_pdf(double x, double mu);
double gsl_ran_lognormal_pdf(double x, double zeta, double sigma);
void gsl_ran_dirichlet(const gsl_rng *r, size_t K, const double alpha[], double theta[]);
double gsl_ran_dirichlet_pdf(size_t K, const double alpha[], const double theta[]);
double gsl_ran_dirichlet_lnpdf(size_t K, const double alpha[], const double theta[]);
#endif
