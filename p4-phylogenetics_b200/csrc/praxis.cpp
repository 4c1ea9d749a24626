// praxis.cpp -- derivative-free minimisers behind the optimiser entry points of the C ABI.
//
//   Praxis        R. P. Brent's principal-axis method ("Algorithms for Minimization without Derivatives",
//                 Prentice-Hall 1973, chapter 7: procedures praxis / min / flin / quad / minfit / sort), the
//                 method the reference's Brent-Powell optimisers run (Pf/brent.c, called from
//                 Pf/p4_treeOpt.c:947-1180 p4_allBrentPowellOptimize and :1182-1330 p4_newtAndBrentPowellOpt).
//                 Restated here from the published algorithm with the reference's settings -- scbd = 1,
//                 illc = false, ktm = 1, the start vector's q0 zeroed, libc random() for the random steps of
//                 the ill-conditioned branch -- so that the same objective is walked the same way.
//   boundedPowell Powell's conjugate-direction method with every line search confined to the box, each line
//                 search Brent's golden-section / parabolic minimiser.  It stands where the reference calls
//                 nlopt's BOBYQA (Pf/p4_treeOpt.c:617-753): nlopt is a third-party library this engine does
//                 not link; the contract kept is the box and the optimum, not BOBYQA's trajectory.
//
// Both take the objective as a callable on a parameter vector; they know nothing about trees.  They are host
// code: every evaluation of the objective is a p4_setPrams + p4_treeLogLike on the GPU (csrc/opt.cpp).
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <vector>

#include "optim.h"

namespace p4b {

namespace {
inline double uniform01() { return (double)random() / ((double)RAND_MAX + 1.0); }   // Pf/util.c:38-42
inline double hyp(double a, double b)     // sqrt(a*a + b*b) the way minfit forms it: scaled by the larger magnitude
{
    if (fabs(a) < fabs(b)) return fabs(b) * sqrt(1.0 + (a / b) * (a / b));
    if (a != 0.0) return fabs(a) * sqrt(1.0 + (b / a) * (b / a));
    return 0.0;
}
}  // namespace

// ------------------------------------------------------------------------------------------------------------
// Praxis
// ------------------------------------------------------------------------------------------------------------
struct Praxis::State {
    int n = 0;
    Objective f;
    double *x = nullptr;             // the caller's vector: the current point
    double fx = 0.0;                 // f at x
    std::vector<double> V;           // directions, column j = direction j: V[i * n + j]
    std::vector<double> d;           // second-derivative estimates along the directions
    std::vector<double> q0, q1, xnew, y, z;
    double qd0 = 0.0, qd1 = 0.0, qf1 = 0.0;
    double toler = 0.0, htol = 0.0, ldt = 0.0, dmin = 0.0;
    double eps2, m2, m4, vsmall, large, vlarge;
    long nl = 0;                     // line searches so far
    double &v(int i, int j) { return V[(size_t)i * n + j]; }
};

Praxis::Praxis(int n) : S(new State())
{
    S->n = n;
    S->V.assign((size_t)n * n, 0.0);
    S->d.assign(n, 0.0);
    S->q0.assign(n, 0.0);
    S->q1.assign(n, 0.0);
    S->xnew.assign(n, 0.0);
    S->y.assign(n, 0.0);
    S->z.assign(n, 0.0);
}
Praxis::~Praxis() { delete S; }

// f at distance lambda from x along direction j (j >= 0), or along the parabola through the last three
// points of the quadratic-extrapolation step (j < 0).  [Brent: flin]
static double along(Praxis::State &s, int j, double lambda)
{
    const int n = s.n;
    if (j >= 0) {
        for (int i = 0; i < n; i++) s.xnew[i] = s.x[i] + lambda * s.v(i, j);
    } else {
        const double qa = lambda * (lambda - s.qd1) / (s.qd0 * (s.qd0 + s.qd1));
        const double qb = (lambda + s.qd0) * (s.qd1 - lambda) / (s.qd0 * s.qd1);
        const double qc = lambda * (lambda + s.qd0) / (s.qd1 * (s.qd0 + s.qd1));
        for (int i = 0; i < n; i++) s.xnew[i] = qa * s.q0[i] + qb * s.x[i] + qc * s.q1[i];
    }
    return s.f(s.xnew.data());
}

// One-dimensional search along direction j (or the parabola, j < 0): on return x1 is the step taken and d2 the
// new second-derivative estimate; x moves only for j >= 0.  fk: f1 = f(x1) is already known.  [Brent: min]
static void lineSearch(Praxis::State &s, int j, int nits, double &d2io, double &x1io, double f1, bool fk)
{
    const int n = s.n;
    double d2 = d2io, x1 = x1io;
    const double sf1 = f1, sx1 = x1;
    int k = 0;
    double xm = 0.0, f0 = s.fx, fm = s.fx;
    bool needD2 = d2 < DBL_EPSILON;           // then f''(0) has to be estimated first
    // step size
    double len = 0.0;
    for (int i = 0; i < n; i++) len += s.x[i] * s.x[i];
    len = sqrt(len);
    double t2 = s.m4 * sqrt(fabs(s.fx) / (needD2 ? s.dmin : d2) + len * s.ldt) + s.m2 * s.ldt;
    len = s.m4 * len + s.toler;
    if (needD2 && t2 > len) t2 = len;
    if (t2 < s.eps2) t2 = s.eps2;
    if (t2 > 0.01 * s.htol) t2 = 0.01 * s.htol;
    if (fk && f1 <= fm) { xm = x1; fm = f1; }
    if (!fk || fabs(x1) < t2) {
        x1 = x1 >= 0.0 ? t2 : -t2;
        f1 = along(s, j, x1);
    }
    if (f1 <= fm) { xm = x1; fm = f1; }
    double x2 = 0.0, f2 = 0.0;
    bool done = false;
    while (!done) {
        if (needD2) {      // a third point gives the second derivative
            x2 = f0 < f1 ? -x1 : 2.0 * x1;
            f2 = along(s, j, x2);
            if (f2 <= fm) { xm = x2; fm = f2; }
            d2 = (x2 * (f1 - f0) - x1 * (f2 - f0)) / (x1 * x2 * (x1 - x2));
        }
        const double d1 = (f1 - f0) / x1 - x1 * d2;      // first derivative at 0
        needD2 = true;
        if (d2 <= s.eps2) x2 = d1 < 0.0 ? s.htol : -s.htol;     // predicted minimum
        else x2 = -0.5 * d1 / d2;
        if (fabs(x2) > s.htol) x2 = x2 > 0.0 ? s.htol : -s.htol;
        for (;;) {         // f at the predicted minimum; halve the step while it is no improvement
            f2 = along(s, j, x2);
            done = true;
            if (k < nits && f2 > f0) {
                done = false;
                k++;
                if (f0 < f1 && x1 * x2 > 0.0) break;     // try the other side: back to the outer loop
                x2 *= 0.5;
            }
            if (done) break;
        }
    }
    s.nl++;
    if (f2 > fm) x2 = xm;
    else fm = f2;
    if (fabs(x2 * (x2 - x1)) > s.eps2) d2 = (x2 * (f1 - f0) - x1 * (fm - f0)) / (x1 * x2 * (x1 - x2));
    else if (k > 0) d2 = 0.0;
    if (d2 < s.eps2) d2 = s.eps2;
    x1 = x2;
    s.fx = fm;
    if (sf1 < s.fx) { s.fx = sf1; x1 = sx1; }
    if (j >= 0)
        for (int i = 0; i < n; i++) s.x[i] += x1 * s.v(i, j);
    x1io = x1;
    d2io = d2;
}

// Quadratic extrapolation through the last three iterates, in case the search is following a curved valley.
// [Brent: quad]
static void quadStep(Praxis::State &s)
{
    const int n = s.n;
    double t = s.fx;
    s.fx = s.qf1;
    s.qf1 = t;
    s.qd1 = 0.0;
    for (int i = 0; i < n; i++) {
        t = s.x[i];
        s.x[i] = s.q1[i];
        s.q1[i] = t;
        s.qd1 += (s.q1[i] - s.x[i]) * (s.q1[i] - s.x[i]);
    }
    s.qd1 = sqrt(s.qd1);
    double qa, qb, qc;
    if (s.qd0 > 0.0 && s.qd1 > 0.0 && s.nl >= 3L * n * n) {
        double zero = 0.0, lambda = s.qd1;
        lineSearch(s, -1, 2, zero, lambda, s.qf1, true);
        qa = lambda * (lambda - s.qd1) / (s.qd0 * (s.qd0 + s.qd1));
        qb = (lambda + s.qd0) * (s.qd1 - lambda) / (s.qd0 * s.qd1);
        qc = lambda * (lambda + s.qd0) / (s.qd1 * (s.qd0 + s.qd1));
    } else {
        s.fx = s.qf1;
        qa = qb = 0.0;
        qc = 1.0;
    }
    s.qd0 = s.qd1;
    for (int i = 0; i < n; i++) {
        t = s.q0[i];
        s.q0[i] = s.x[i];
        s.x[i] = qa * t + qb * s.x[i] + qc * s.q1[i];
    }
}

// Singular values q and right singular vectors (returned in ab) of the n x n matrix ab: Golub & Reinsch's
// Householder bidiagonalisation + implicit-shift QR, as Brent uses it.  e is work space.  [Brent: minfit]
static void minfit(int n, double eps, double tol, std::vector<double> &A, std::vector<double> &q, std::vector<double> &e)
{
    auto ab = [&](int i, int j) -> double & { return A[(size_t)i * n + j]; };
    int l = 0;
    double g = 0.0, x = 0.0;
    for (int i = 0; i < n; i++) {              // Householder reduction to bidiagonal form
        e[i] = g;
        l = i + 1;
        double s = 0.0;
        for (int j = i; j < n; j++) s += ab(j, i) * ab(j, i);
        if (s < tol) g = 0.0;
        else {
            const double f = ab(i, i);
            g = f < 0.0 ? sqrt(s) : -sqrt(s);
            const double h = f * g - s;
            ab(i, i) = f - g;
            for (int j = l; j < n; j++) {
                double ff = 0.0;
                for (int k = i; k < n; k++) ff += ab(k, i) * ab(k, j);
                ff /= h;
                for (int k = i; k < n; k++) ab(k, j) += ff * ab(k, i);
            }
        }
        q[i] = g;
        s = 0.0;
        for (int j = l; j < n; j++) s += ab(i, j) * ab(i, j);
        if (s < tol) g = 0.0;
        else {
            const double f = ab(i, i + 1);
            g = f < 0.0 ? sqrt(s) : -sqrt(s);
            const double h = f * g - s;
            ab(i, i + 1) = f - g;
            for (int j = l; j < n; j++) e[j] = ab(i, j) / h;
            for (int j = l; j < n; j++) {
                double ss = 0.0;
                for (int k = l; k < n; k++) ss += ab(j, k) * ab(i, k);
                for (int k = l; k < n; k++) ab(j, k) += ss * e[k];
            }
        }
        const double yy = fabs(q[i]) + fabs(e[i]);
        if (yy > x) x = yy;
    }
    for (int i = n - 1; i >= 0; i--) {         // accumulation of the right-hand transformations
        if (g != 0.0) {
            const double h = ab(i, i + 1) * g;
            for (int j = l; j < n; j++) ab(j, i) = ab(i, j) / h;
            for (int j = l; j < n; j++) {
                double s = 0.0;
                for (int k = l; k < n; k++) s += ab(i, k) * ab(k, j);
                for (int k = l; k < n; k++) ab(k, j) += s * ab(k, i);
            }
        }
        for (int j = l; j < n; j++) ab(i, j) = ab(j, i) = 0.0;
        ab(i, i) = 1.0;
        g = e[i];
        l = i;
    }
    eps *= x;                                   // diagonalisation of the bidiagonal form
    for (int k = n - 1; k >= 0; k--) {
        for (int kt = 1;; kt++) {
            if (kt > 30) e[k] = 0.0;
            bool cancel = false;
            for (int l2 = k; l2 >= 0; l2--) {   // test for splitting
                l = l2;
                if (fabs(e[l]) <= eps) break;
                if (fabs(q[l - 1]) <= eps) { cancel = true; break; }
            }
            if (cancel) {                       // cancellation of e[l]
                double c = 0.0, s = 1.0;
                for (int i = l; i <= k; i++) {
                    const double f = s * e[i];
                    e[i] *= c;
                    if (fabs(f) <= eps) break;
                    double gg = q[i];
                    double h = hyp(f, gg);
                    q[i] = h;
                    if (h == 0.0) gg = h = 1.0;
                    c = gg / h;
                    s = -f / h;
                }
            }
            double z2 = q[k];
            if (l == k) {                       // converged
                if (z2 < 0.0) {
                    q[k] = -z2;
                    for (int j = 0; j < n; j++) ab(j, k) = -ab(j, k);
                }
                break;
            }
            double xx = q[l], y2 = q[k - 1], gg = e[k - 1], h = e[k];       // shift from the bottom 2 x 2 minor
            double f = ((y2 - z2) * (y2 + z2) + (gg - h) * (gg + h)) / (2.0 * h * y2);
            gg = sqrt(f * f + 1.0);
            double s = f < 0.0 ? f - gg : f + gg;
            f = ((xx - z2) * (xx + z2) + h * (y2 / s - h)) / xx;
            double c = 1.0;
            s = 1.0;
            for (int i = l + 1; i <= k; i++) {  // next QR transformation
                gg = e[i];
                y2 = q[i];
                h = s * gg;
                gg *= c;
                z2 = hyp(f, h);
                e[i - 1] = z2;
                if (z2 == 0.0) z2 = f = 1.0;
                c = f / z2;
                s = h / z2;
                f = xx * c + gg * s;
                gg = -xx * s + gg * c;
                h = y2 * s;
                y2 *= c;
                for (int j = 0; j < n; j++) {
                    const double a0 = ab(j, i - 1), a1 = ab(j, i);
                    ab(j, i - 1) = a0 * c + a1 * s;
                    ab(j, i) = -a0 * s + a1 * c;
                }
                z2 = hyp(f, h);
                q[i - 1] = z2;
                if (z2 == 0.0) z2 = f = 1.0;
                c = f / z2;
                s = h / z2;
                f = c * gg + s * y2;
                xx = -s * gg + c * y2;
            }
            e[l] = 0.0;
            e[k] = f;
            q[k] = xx;
        }
    }
}

// Directions by decreasing second derivative.  [Brent: sort]
static void sortDirections(Praxis::State &s)
{
    const int n = s.n;
    for (int i = 0; i < n - 1; i++) {
        int k = i;
        double big = s.d[i];
        for (int j = i + 1; j < n; j++)
            if (s.d[j] > big) { k = j; big = s.d[j]; }
        if (k > i) {
            s.d[k] = s.d[i];
            s.d[i] = big;
            for (int j = 0; j < n; j++) std::swap(s.v(j, i), s.v(j, k));
        }
    }
}

double Praxis::minimize(double tol, double h, double *x, Objective f)
{
    State &s = *S;
    const int n = s.n;
    s.f = f;
    s.x = x;
    s.eps2 = DBL_EPSILON * DBL_EPSILON;
    s.m2 = sqrt(DBL_EPSILON);
    s.m4 = sqrt(s.m2);
    s.vsmall = s.eps2 * s.eps2;
    s.large = 1.0 / s.eps2;
    s.vlarge = 1.0 / s.vsmall;
    const int ktm = 1;               // ktm + 1 iterations without improvement end the search
    bool illc = false;
    const double ldfac = illc ? 0.1 : 0.01;
    int kt = 0;
    s.nl = 0;
    s.qf1 = s.fx = s.f(s.x);
    double t2 = s.eps2 + fabs(tol);
    s.toler = t2;
    s.dmin = s.eps2;
    s.htol = h < 100.0 * s.toler ? 100.0 * s.toler : h;
    s.ldt = s.htol;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) s.v(i, j) = i == j ? 1.0 : 0.0;
    s.d[0] = s.qd0 = 0.0;
    for (int i = 0; i < n; i++) { s.q0[i] = 0.0; s.q1[i] = s.x[i]; }

    for (;;) {
        double sf = s.d[0], step = 0.0;
        s.d[0] = 0.0;
        lineSearch(s, 0, 2, s.d[0], step, s.fx, false);            // along the first direction
        if (step < 0.0)
            for (int i = 0; i < n; i++) s.v(i, 0) = -s.v(i, 0);
        if (sf <= 0.9 * s.d[0] || 0.9 * sf >= s.d[0])
            for (int i = 1; i < n; i++) s.d[i] = 0.0;
        for (int k = 1; k < n; k++) {
            for (int i = 0; i < n; i++) s.y[i] = s.x[i];
            sf = s.fx;
            illc = illc || kt > 0;
            int kl;
            double df;
            for (;;) {
                kl = k;
                df = 0.0;
                if (illc) {            // a random step to get off a resolution valley
                    for (int i = 0; i < n; i++) {
                        const double r = s.z[i] = (0.1 * s.ldt + t2 * pow(10.0, kt)) * (uniform01() - 0.5);
                        for (int j = 0; j < n; j++) s.x[j] += r * s.v(j, i);
                    }
                    s.fx = s.f(s.x);
                }
                for (int k2 = k; k2 < n; k2++) {                   // the non-conjugate directions
                    const double sl = s.fx;
                    double st = 0.0;
                    lineSearch(s, k2, 2, s.d[k2], st, s.fx, false);
                    double gain;
                    if (illc) { const double sz = st + s.z[k2]; gain = s.d[k2] * sz * sz; }
                    else gain = sl - s.fx;
                    if (df < gain) { df = gain; kl = k2; }
                }
                if (!illc && df < fabs(100.0 * DBL_EPSILON * s.fx)) illc = true;     // no success: once more, with random steps
                else break;
            }
            for (int k2 = 0; k2 < k; k2++) {                       // the conjugate directions
                double st = 0.0;
                lineSearch(s, k2, 2, s.d[k2], st, s.fx, false);
            }
            const double f1 = s.fx;
            s.fx = sf;
            double lds = 0.0;
            for (int i = 0; i < n; i++) {
                double sl = s.x[i];
                s.x[i] = s.y[i];
                sl -= s.y[i];
                s.y[i] = sl;
                lds += sl * sl;
            }
            lds = sqrt(lds);
            if (lds > s.eps2) {
                for (int i = kl - 1; i >= k; i--) {                // direction kl makes room for the new one
                    for (int j = 0; j < n; j++) s.v(j, i + 1) = s.v(j, i);
                    s.d[i + 1] = s.d[i];
                }
                s.d[k] = 0.0;
                for (int i = 0; i < n; i++) s.v(i, k) = s.y[i] / lds;
                lineSearch(s, k, 4, s.d[k], lds, f1, true);        // ... and the search goes along it
                if (lds <= 0.0) {
                    lds = -lds;
                    for (int i = 0; i < n; i++) s.v(i, k) = -s.v(i, k);
                }
            }
            s.ldt *= ldfac;
            if (s.ldt < lds) s.ldt = lds;
            t2 = 0.0;
            for (int i = 0; i < n; i++) t2 += s.x[i] * s.x[i];
            t2 = s.m2 * sqrt(t2) + s.toler;
            kt = s.ldt > 0.5 * t2 ? 0 : kt + 1;                    // step shorter than half the tolerance?
            if (kt > ktm) return s.fx;
        }
        if (n == 1) {      // a single direction: the loop above is empty, the stopping test lives here
            t2 = s.m2 * fabs(s.x[0]) + s.toler;
            s.ldt *= ldfac;
            kt = s.ldt > 0.5 * t2 ? 0 : kt + 1;
            if (kt > ktm) return s.fx;
        }
        quadStep(s);
        // V = U D^(-1/2), then its singular value decomposition: the principal axes of the approximating
        // quadratic form, without squaring the condition number
        double dn = 0.0;
        for (int i = 0; i < n; i++) {
            s.d[i] = 1.0 / sqrt(s.d[i]);
            if (s.d[i] > dn) dn = s.d[i];
        }
        for (int j = 0; j < n; j++) {
            const double sc = s.d[j] / dn;
            for (int i = 0; i < n; i++) s.v(i, j) *= sc;
        }
        for (int i = 1; i < n; i++)
            for (int j = 0; j < i; j++) std::swap(s.v(i, j), s.v(j, i));
        minfit(n, DBL_EPSILON, s.vsmall, s.V, s.d, s.y);
        for (int i = 0; i < n; i++) {
            const double sv = dn * s.d[i];
            if (sv > s.large) s.d[i] = s.vsmall;
            else if (sv < s.eps2) s.d[i] = s.vlarge;
            else s.d[i] = 1.0 / (sv * sv);
        }
        sortDirections(s);
        s.dmin = s.d[n - 1] < s.eps2 ? s.eps2 : s.d[n - 1];
        illc = s.m2 * s.d[0] > s.dmin;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Bounded Powell
// ------------------------------------------------------------------------------------------------------------
namespace {
// Brent's minimiser of g on [a, b] (golden section + successive parabolic interpolation), started from the
// known point (x0, g0) inside the interval.
template <class G>
double brentBounded(G g, double a, double b, double x0, double g0, double xtol, int maxIter, double *gBest)
{
    const double golden = 0.3819660112501051;
    double x = x0, w = x0, v = x0, fx = g0, fw = g0, fv = g0, dd = 0.0, e = 0.0;
    for (int it = 0; it < maxIter; it++) {
        const double xm = 0.5 * (a + b), tol1 = xtol * fabs(x) + 1e-11, tol2 = 2.0 * tol1;
        if (fabs(x - xm) <= tol2 - 0.5 * (b - a)) break;
        bool gold = true;
        if (fabs(e) > tol1) {
            double r = (x - w) * (fx - fv), q = (x - v) * (fx - fw), p = (x - v) * q - (x - w) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) p = -p;
            q = fabs(q);
            const double eOld = e;
            e = dd;
            if (!(fabs(p) >= fabs(0.5 * q * eOld) || p <= q * (a - x) || p >= q * (b - x))) {
                dd = p / q;
                const double u = x + dd;
                if (u - a < tol2 || b - u < tol2) dd = xm >= x ? tol1 : -tol1;
                gold = false;
            }
        }
        if (gold) {
            e = x >= xm ? a - x : b - x;
            dd = golden * e;
        }
        const double u = fabs(dd) >= tol1 ? x + dd : x + (dd >= 0 ? tol1 : -tol1);
        const double fu = g(u);
        if (fu <= fx) {
            if (u >= x) a = x; else b = x;
            v = w; fv = fw; w = x; fw = fx; x = u; fx = fu;
        } else {
            if (u < x) a = u; else b = u;
            if (fu <= fw || w == x) { v = w; fv = fw; w = u; fw = fu; }
            else if (fu <= fv || v == x || v == w) { v = u; fv = fu; }
        }
    }
    *gBest = fx;
    return x;
}
}  // namespace

double boundedPowell(int n, double *x, const double *lo, const double *hi, Objective f, double xtol, double ftol, long maxEvals, long *nEvals)
{
    std::vector<double> dir((size_t)n * n, 0.0), x0(n), xt(n), dnew(n);
    for (int i = 0; i < n; i++) {
        dir[(size_t)i * n + i] = 1.0;
        if (x[i] < lo[i]) x[i] = lo[i];
        if (x[i] > hi[i]) x[i] = hi[i];
    }
    long evals = 0;
    auto F = [&](const double *p) { evals++; for (int i = 0; i < n; i++) xt[i] = p[i]; return f(xt.data()); };
    double fx = F(x);
    // minimise along direction d from x inside the box; moves x, returns the new f
    std::vector<double> probe(n);
    auto lineMin = [&](const double *d, double fStart) {
        double tLo = -HUGE_VAL, tHi = HUGE_VAL;       // the segment of the line inside the box
        for (int i = 0; i < n; i++) {
            if (d[i] == 0.0) continue;
            double t1 = (lo[i] - x[i]) / d[i], t2 = (hi[i] - x[i]) / d[i];
            if (t1 > t2) std::swap(t1, t2);
            if (t1 > tLo) tLo = t1;
            if (t2 < tHi) tHi = t2;
        }
        if (!(tHi - tLo > 0.0) || tLo == -HUGE_VAL) return fStart;
        if (tLo > 0.0) tLo = 0.0;
        if (tHi < 0.0) tHi = 0.0;
        auto g = [&](double t) {
            for (int i = 0; i < n; i++) {
                double v = x[i] + t * d[i];
                probe[i] = v < lo[i] ? lo[i] : (v > hi[i] ? hi[i] : v);
            }
            return F(probe.data());
        };
        double gBest = fStart;
        const double t = brentBounded(g, tLo, tHi, 0.0, fStart, xtol, 60, &gBest);
        if (gBest < fStart) {
            for (int i = 0; i < n; i++) {
                double v = x[i] + t * d[i];
                x[i] = v < lo[i] ? lo[i] : (v > hi[i] ? hi[i] : v);
            }
            return gBest;
        }
        return fStart;
    };
    bool pristine = true;      // the directions are the coordinate axes
    for (int iter = 0; iter < 200 && evals < maxEvals; iter++) {
        const double fIterStart = fx;
        for (int i = 0; i < n; i++) x0[i] = x[i];
        int big = 0;
        double bigDrop = 0.0;
        for (int k = 0; k < n && evals < maxEvals; k++) {
            const double before = fx;
            fx = lineMin(&dir[(size_t)k * n], fx);
            if (before - fx > bigDrop) { bigDrop = before - fx; big = k; }
        }
        if (2.0 * (fIterStart - fx) <= ftol * (fabs(fIterStart) + fabs(fx)) + 1e-20) {
            // no progress: a direction set that has become (nearly) dependent -- it happens on a face of the box -- can
            // stall short of the minimum, so the coordinate directions get one more round before the search ends
            if (pristine) break;
            for (int i = 0; i < n; i++)
                for (int j = 0; j < n; j++) dir[(size_t)i * n + j] = i == j ? 1.0 : 0.0;
            pristine = true;
            continue;
        }
        // the direction of the iteration's total move replaces the direction of largest decrease when Powell's
        // criterion says the extrapolated point promises more
        double norm = 0.0;
        for (int i = 0; i < n; i++) {
            dnew[i] = x[i] - x0[i];
            norm += dnew[i] * dnew[i];
            double v = 2.0 * x[i] - x0[i];
            probe[i] = v < lo[i] ? lo[i] : (v > hi[i] ? hi[i] : v);
        }
        if (norm == 0.0) break;
        const double fE = F(probe.data());
        if (fE < fIterStart) {
            const double t = 2.0 * (fIterStart - 2.0 * fx + fE) * (fIterStart - fx - bigDrop) * (fIterStart - fx - bigDrop) -
                             bigDrop * (fIterStart - fE) * (fIterStart - fE);
            if (t < 0.0) {
                fx = lineMin(dnew.data(), fx);
                for (int i = 0; i < n; i++) {
                    dir[(size_t)big * n + i] = dir[(size_t)(n - 1) * n + i];
                    dir[(size_t)(n - 1) * n + i] = dnew[i];
                }
                pristine = false;
            }
        }
    }
    if (nEvals) *nEvals += evals;
    return fx;
}

}  // namespace p4b
