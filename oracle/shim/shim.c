/* oracle/shim/shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Functional stand-ins for the handful of GSL and nlopt entry points that the
 * reference's Pf sources link against.  They exist so
 * that the UNMODIFIED reference sources under /root/reference/Pf can be built
 * into oracle/_ref/pf.so in an image that has neither GSL nor nlopt.  Nothing
 * here is on the likelihood hot path: the reference uses GSL only for random
 * numbers / special functions in proposals and nlopt only for BOBYQA.
 *
 * The RNG is MT19937 (what GSL's default generator is), written from the
 * published algorithm (Matsumoto & Nishimura 1998).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <gsl/gsl_rng.h>
#include <gsl/gsl_randist.h>
#include <gsl/gsl_sf_gamma.h>
#include <gsl/gsl_statistics_double.h>
#include <nlopt.h>

#define MT_N 624
#define MT_M 397
typedef struct {
    unsigned long mt[MT_N];
    int mti;
} mt_state_t;

static const gsl_rng_type mt_type = {"mt19937", 0xffffffffUL, 0, sizeof(mt_state_t)};
const gsl_rng_type *gsl_rng_default = &mt_type;

const gsl_rng_type *gsl_rng_env_setup(void) { return gsl_rng_default; }

static void mt_set(mt_state_t *s, unsigned long seed)
{
    int i;
    if (seed == 0) seed = 4357; /* GSL's convention for seed 0 */
    s->mt[0] = seed & 0xffffffffUL;
    for (i = 1; i < MT_N; i++) {
        s->mt[i] = (1812433253UL * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (unsigned long)i);
        s->mt[i] &= 0xffffffffUL;
    }
    s->mti = MT_N;
}

static unsigned long mt_get(mt_state_t *s)
{
    unsigned long y;
    static const unsigned long mag01[2] = {0x0UL, 0x9908b0dfUL};
    if (s->mti >= MT_N) {
        int kk;
        for (kk = 0; kk < MT_N - MT_M; kk++) {
            y = (s->mt[kk] & 0x80000000UL) | (s->mt[kk + 1] & 0x7fffffffUL);
            s->mt[kk] = s->mt[kk + MT_M] ^ (y >> 1) ^ mag01[y & 1UL];
        }
        for (; kk < MT_N - 1; kk++) {
            y = (s->mt[kk] & 0x80000000UL) | (s->mt[kk + 1] & 0x7fffffffUL);
            s->mt[kk] = s->mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 1UL];
        }
        y = (s->mt[MT_N - 1] & 0x80000000UL) | (s->mt[0] & 0x7fffffffUL);
        s->mt[MT_N - 1] = s->mt[MT_M - 1] ^ (y >> 1) ^ mag01[y & 1UL];
        s->mti = 0;
    }
    y = s->mt[s->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680UL;
    y ^= (y << 15) & 0xefc60000UL;
    y ^= (y >> 18);
    return y & 0xffffffffUL;
}

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T)
{
    gsl_rng *r = (gsl_rng *)malloc(sizeof(gsl_rng));
    if (!r) return NULL;
    r->type = T;
    r->state = calloc(1, T->size);
    if (!r->state) { free(r); return NULL; }
    mt_set((mt_state_t *)r->state, 0);
    return r;
}
void gsl_rng_free(gsl_rng *r) { if (r) { free(r->state); free(r); } }
void gsl_rng_set(const gsl_rng *r, unsigned long seed) { mt_set((mt_state_t *)r->state, seed); }
unsigned long gsl_rng_get(const gsl_rng *r) { return mt_get((mt_state_t *)r->state); }
double gsl_rng_uniform(const gsl_rng *r) { return mt_get((mt_state_t *)r->state) / 4294967296.0; }
unsigned long gsl_rng_uniform_int(const gsl_rng *r, unsigned long n)
{
    unsigned long range = 0xffffffffUL, scale, k;
    if (n == 0 || n > range) return 0;
    scale = range / n;
    do { k = mt_get((mt_state_t *)r->state) / scale; } while (k >= n);
    return k;
}
size_t gsl_rng_size(const gsl_rng *r) { return r->type->size; }
void *gsl_rng_state(const gsl_rng *r) { return r->state; }

/* --- distributions ------------------------------------------------------ */
static double uniform_pos(const gsl_rng *r)
{
    double x;
    do { x = gsl_rng_uniform(r); } while (x == 0.0);
    return x;
}
static double gaussian(const gsl_rng *r)
{
    double x, y, r2;
    do {
        x = -1 + 2 * uniform_pos(r);
        y = -1 + 2 * uniform_pos(r);
        r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0);
    return y * sqrt(-2.0 * log(r2) / r2);
}
double gsl_ran_gamma(const gsl_rng *r, double a, double b)
{
    /* Marsaglia & Tsang (2000) */
    if (a < 1) {
        double u = uniform_pos(r);
        return gsl_ran_gamma(r, 1.0 + a, b) * pow(u, 1.0 / a);
    }
    {
        double x, v, u;
        double d = a - 1.0 / 3.0;
        double c = (1.0 / 3.0) / sqrt(d);
        while (1) {
            do { x = gaussian(r); v = 1.0 + c * x; } while (v <= 0);
            v = v * v * v;
            u = uniform_pos(r);
            if (u < 1 - 0.0331 * x * x * x * x) break;
            if (log(u) < 0.5 * x * x + d * (1 - v + log(v))) break;
        }
        return b * d * v;
    }
}
double gsl_ran_gamma_pdf(double x, double a, double b)
{
    if (x < 0) return 0;
    if (x == 0) return (a == 1) ? 1 / b : 0;
    if (a == 1) return exp(-x / b) / b;
    return exp((a - 1) * log(x / b) - x / b - lgamma(a)) / b;
}
double gsl_ran_chisq_pdf(double x, double nu)
{
    if (x < 0) return 0;
    if (nu == 2.0) return exp(-x / 2.0) / 2.0;
    return exp((nu / 2 - 1) * log(x / 2) - x / 2 - lgamma(nu / 2)) / 2;
}
double gsl_ran_exponential_pdf(double x, double mu) { return x < 0 ? 0 : exp(-x / mu) / mu; }
double gsl_ran_lognormal_pdf(double x, double zeta, double sigma)
{
    if (x <= 0) return 0;
    {
        double u = (log(x) - zeta) / sigma;
        return 1 / (x * fabs(sigma) * sqrt(2 * M_PI)) * exp(-(u * u) / 2);
    }
}
void gsl_ran_dirichlet(const gsl_rng *r, size_t K, const double alpha[], double theta[])
{
    size_t i;
    double norm = 0.0;
    for (i = 0; i < K; i++) theta[i] = gsl_ran_gamma(r, alpha[i], 1.0);
    for (i = 0; i < K; i++) norm += theta[i];
    for (i = 0; i < K; i++) theta[i] /= norm;
}
double gsl_ran_dirichlet_lnpdf(size_t K, const double alpha[], const double theta[])
{
    size_t i;
    double log_p = 0.0, sum_alpha = 0.0;
    for (i = 0; i < K; i++) log_p += (alpha[i] - 1.0) * log(theta[i]);
    for (i = 0; i < K; i++) sum_alpha += alpha[i];
    log_p += lgamma(sum_alpha);
    for (i = 0; i < K; i++) log_p -= lgamma(alpha[i]);
    return log_p;
}
double gsl_ran_dirichlet_pdf(size_t K, const double alpha[], const double theta[])
{
    return exp(gsl_ran_dirichlet_lnpdf(K, alpha, theta));
}
double gsl_sf_lngamma(double x) { return lgamma(x); }
double gsl_sf_gamma(double x) { return tgamma(x); }
double gsl_sf_beta(double a, double b) { return exp(lgamma(a) + lgamma(b) - lgamma(a + b)); }
double gsl_stats_mean(const double data[], size_t stride, size_t n)
{
    long double mean = 0;
    size_t i;
    for (i = 0; i < n; i++) mean += (data[i * stride] - mean) / (i + 1);
    return (double)mean;
}
double gsl_stats_variance_m(const double data[], size_t stride, size_t n, double mean)
{
    long double variance = 0;
    size_t i;
    for (i = 0; i < n; i++) {
        const long double delta = (data[i * stride] - mean);
        variance += (delta * delta - variance) / (i + 1);
    }
    return (double)(variance * ((double)n / (double)(n - 1)));
}

/* --- nlopt: BOBYQA unavailable in the oracle build ---------------------- */
struct nlopt_opt_s { unsigned n; };
nlopt_opt nlopt_create(nlopt_algorithm a, unsigned n)
{
    nlopt_opt o = (nlopt_opt)malloc(sizeof(struct nlopt_opt_s));
    (void)a;
    if (o) o->n = n;
    return o;
}
void nlopt_destroy(nlopt_opt o) { free(o); }
nlopt_result nlopt_set_lower_bounds(nlopt_opt o, const double *lb) { (void)o; (void)lb; return NLOPT_SUCCESS; }
nlopt_result nlopt_set_upper_bounds(nlopt_opt o, const double *ub) { (void)o; (void)ub; return NLOPT_SUCCESS; }
nlopt_result nlopt_set_max_objective(nlopt_opt o, nlopt_func f, void *d) { (void)o; (void)f; (void)d; return NLOPT_SUCCESS; }
nlopt_result nlopt_set_ftol_abs(nlopt_opt o, double tol) { (void)o; (void)tol; return NLOPT_SUCCESS; }
nlopt_result nlopt_optimize(nlopt_opt o, double *x, double *opt_f)
{
    (void)o; (void)x; (void)opt_f;
    return NLOPT_FAILURE;
}
