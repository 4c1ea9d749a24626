/* oracle/shim/nlopt.h -- TEST INFRASTRUCTURE ONLY.
 * nlopt (BOBYQA) is absent from this image.  The stub lets Pf/p4_treeOpt.c
 * compile; nlopt_optimize() reports failure, so the BOBYQA optimisers are
 * unavailable in the oracle while the Brent/Powell ones (self-contained in
 * Pf/brent.c) keep working.  The optimisers are callers of the hot path, not
 * part of it. */
#ifndef ORACLE_SHIM_NLOPT_H
#define ORACLE_SHIM_NLOPT_H
typedef enum { NLOPT_LN_BOBYQA = 34 } nlopt_algorithm;
typedef enum {
    NLOPT_FAILURE = -1, NLOPT_INVALID_ARGS = -2, NLOPT_OUT_OF_MEMORY = -3,
    NLOPT_ROUNDOFF_LIMITED = -4, NLOPT_FORCED_STOP = -5, NLOPT_SUCCESS = 1
} nlopt_result;
typedef struct nlopt_opt_s *nlopt_opt;
typedef double (*nlopt_func)(unsigned n, const double *x, double *grad, void *data);
nlopt_opt nlopt_create(nlopt_algorithm a, unsigned n);
void nlopt_destroy(nlopt_opt o);
nlopt_result nlopt_set_lower_bounds(nlopt_opt o, const double *lb);
nlopt_result nlopt_set_upper_bounds(nlopt_opt o, const double *ub);
nlopt_result nlopt_set_max_objective(nlopt_opt o, nlopt_func f, void *data);
nlopt_result nlopt_set_ftol_abs(nlopt_opt o, double tol);
nlopt_result nlopt_optimize(nlopt_opt o, double *x, double *opt_f);
#endif
