// opt.cpp -- the optimiser's seam onto the likelihood path.
//
// The reference's optimisers (Pf/p4_treeOpt.c) pack every free parameter and, optionally, every
// branch length into one vector (p4_windUpParameters :17-120), and their objective unpacks it,
// rebuilds the model and evaluates the tree (p4_logLikeForNLOpt :579-615:
// p4_unWindParameters -> p4_setPrams -> p4_treeLogLike).  This file provides the same packing,
// unpacking and objective on the B200 engine, so that any bounded optimiser can drive the GPU
// with the reference's own parameterisation, and the reference's four optimiser entry points on top of it
// (p4_allBrentPowellOptimize :947-1180, p4_newtAndBrentPowellOpt :1182-1330, p4_allBOBYQAOptimize :617-753,
// p4_newtAndBOBYQAOpt :755-945): the same schedules of calls, with Brent's praxis restated in csrc/praxis.cpp
// and a bounded Powell method where the reference calls nlopt's BOBYQA (a third-party library this engine does
// not link).  Every objective evaluation is p4_setPrams + p4_treeLogLike on the GPU.
//
// Unpacking follows p4_unWindParameters (:186-565) with one simplification: where the reference
// nudges an out-of-bounds rate by a random amount (RATE_MIN * ranDoubleUpToOne(), :356, :373) the
// value is clamped, so the objective is a deterministic function of the vector.  Like the reference
// (:568-571) a vector that hit a limit is wound up again from the clamped model, in place.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/p4b200.h"
#include "engine.h"
#include "optim.h"

namespace p4b {

static inline double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

int countParameters(Tree *t, int doBrLens)
{
    int n = 0;
    Model *m = t->model;
    for (int p = 0; p < m->nParts; p++) {
        ModelPart *mp = m->parts[p];
        for (auto &c : mp->comps)
            if (c.isFree) n += mp->dim - 1;
        for (auto &r : mp->rMatrices)
            if (r.isFree) n += (r.spec == 5) ? 1 : (mp->dim * (mp->dim - 1) / 2 - 1);
        for (Gdasrv *g : mp->gdasrvs)
            if (g && g->isFree) n += 1;
        if (mp->pInvarFree) n += 1;
    }
    if (m->relRatesAreFree) n += m->nParts - 1;
    if (doBrLens)
        for (Node *nd : t->nodes)
            if (nd && nd != t->root) n++;
    return n;
}

int windUpParameters(Tree *t, int doBrLens, double *x, double *lb, double *ub)
{
    Model *m = t->model;
    int pos = 0;
    auto put = [&](double v, double lo, double hi) {
        x[pos] = v;
        if (lb) lb[pos] = lo;
        if (ub) ub[pos] = hi;
        pos++;
    };
    for (int p = 0; p < m->nParts; p++) {
        ModelPart *mp = m->parts[p];
        const int dim = mp->dim;
        for (auto &c : mp->comps)
            if (c.isFree)
                for (int i = 0; i < dim - 1; i++) put(c.val[i], m->PIVEC_MIN[0], m->PIVEC_MAX[0]);
        for (auto &r : mp->rMatrices)
            if (r.isFree) {
                if (r.spec == 5) put(r.kappa, m->KAPPA_MIN[0], m->KAPPA_MAX[0]);
                else
                    for (int i = 0; i < dim - 2; i++)
                        for (int j = i + 1; j < dim; j++) put(r.bigR[i * dim + j], m->RATE_MIN[0], m->RATE_MAX[0]);
            }
        for (Gdasrv *g : mp->gdasrvs)
            if (g && g->isFree) put(g->val[0], m->GAMMA_SHAPE_MIN[0], m->GAMMA_SHAPE_MAX[0]);
        if (mp->pInvarFree) put(mp->pInvar, m->PINVAR_MIN[0], m->PINVAR_MAX[0]);
    }
    if (m->relRatesAreFree)
        for (int p = 0; p < m->nParts - 1; p++) put(m->parts[p]->relRate, m->RELRATE_MIN[0], m->RELRATE_MAX[0]);
    if (doBrLens)
        for (Node *nd : t->nodes)
            if (nd && nd != t->root) put(nd->brLen, m->BRLEN_MIN[0], m->BRLEN_MAX[0]);
    return pos;
}

static bool g_hitLimit = false;   // set by unWindParameters when a value had to be clamped
static inline double clampHit(double v, double lo, double hi)
{
    if (v < lo) { g_hitLimit = true; return lo; }
    if (v > hi) { g_hitLimit = true; return hi; }
    return v;
}

int unWindParameters(Tree *t, int doBrLens, const double *x)
{
    Model *m = t->model;
    int pos = 0;
    g_hitLimit = false;
    for (int p = 0; p < m->nParts; p++) {
        ModelPart *mp = m->parts[p];
        const int dim = mp->dim;
        const double pmin = m->PIVEC_MIN[0];
        for (auto &c : mp->comps)
            if (c.isFree) {   // :212-247: shift by PIVEC_MIN, last value by difference, rescale
                double sum = 0.0;
                for (int i = 0; i < dim - 1; i++) {
                    double par = x[pos++];
                    if (par <= pmin) { g_hitLimit = true; par = 0.0; } else par = par - pmin;
                    c.val[i] = par;
                    sum += par;
                }
                const double reducedUpper = 1.0 - dim * pmin;
                const double diff = reducedUpper - sum;
                if (diff < 0.0) c.val[dim - 1] = 0.0;
                else { c.val[dim - 1] = diff; sum += diff; }
                const double factor = reducedUpper / sum;
                for (int i = 0; i < dim; i++) c.val[i] = c.val[i] * factor + pmin;
            }
        for (auto &r : mp->rMatrices)
            if (r.isFree) {
                if (r.spec == 5) {
                    r.kappa = clampHit(x[pos++], m->KAPPA_MIN[0], m->KAPPA_MAX[0]);
                    setKappaBigR(r);
                } else if (m->rMatrixNormalizeTo1 && m->rMatrixNormalizeTo1[0]) {   // :343-404
                    double sum = 0.0;
                    for (int i = 0; i < dim - 2; i++)
                        for (int j = i + 1; j < dim; j++) {
                            const double par = clampHit(x[pos++], m->RATE_MIN[0], 0.999);
                            r.bigR[i * dim + j] = par;
                            sum += par;
                        }
                    double last = sum < 1.0 ? clampHit(1.0 - sum, m->RATE_MIN[0], 0.999) : m->RATE_MIN[0] * sum;
                    sum += last;
                    r.bigR[(dim - 2) * dim + (dim - 1)] = last;
                    for (int i = 0; i < dim - 1; i++)
                        for (int j = i + 1; j < dim; j++) {
                            if (sum != 1.0) r.bigR[i * dim + j] /= sum;
                            r.bigR[j * dim + i] = r.bigR[i * dim + j];
                        }
                } else {
                    for (int i = 0; i < dim - 2; i++)
                        for (int j = i + 1; j < dim; j++) {
                            const double par = clampHit(x[pos++], m->RATE_MIN[0], m->RATE_MAX[0]);
                            r.bigR[i * dim + j] = r.bigR[j * dim + i] = par;
                        }
                }
            }
        for (Gdasrv *g : mp->gdasrvs)
            if (g && g->isFree) g->val[0] = clampHit(x[pos++], m->GAMMA_SHAPE_MIN[0], m->GAMMA_SHAPE_MAX[0]);
        if (mp->pInvarFree) mp->pInvar = clampHit(x[pos++], m->PINVAR_MIN[0], m->PINVAR_MAX[0]);
    }
    if (m->relRatesAreFree) {   // :521-545: the last part's rate keeps the site-weighted mean at 1
        for (int p = 0; p < m->nParts - 1; p++) m->parts[p]->relRate = clampHit(x[pos++], m->RELRATE_MIN[0], m->RELRATE_MAX[0]);
        long totLen = 0;
        for (int p = 0; p < m->nParts; p++) totLen += t->data->parts[p]->nChar;
        double sum = 0.0;
        for (int p = 0; p < m->nParts - 1; p++) sum = sum + (m->parts[p]->relRate * t->data->parts[p]->nChar);
        m->parts[m->nParts - 1]->relRate = (((double)totLen) - sum) / t->data->parts[m->nParts - 1]->nChar;
    }
    if (doBrLens)
        for (Node *nd : t->nodes)
            if (nd && nd != t->root) nd->brLen = clampHit(x[pos++], m->BRLEN_MIN[0], m->BRLEN_MAX[0]);
    return pos;
}

// ---------------------------------------------------------------------------
// Branch lengths, one at a time, through the dirty path.
//
// The reference's newtAnd* optimisers (Pf/p4_treeOpt.c:755-945, Pf/p4_treeNewt.c) go around the tree
// optimising one branch length at a time (Newton-Raphson on partials kept for both directions of every
// branch) and hand the model parameters to Brent-Powell or BOBYQA.  Here each one-branch objective is
// what Chain.proposeSp evaluates after a branch-length proposal: the branch's P(t), the conditional
// likelihoods from its parent to the root, the part log-likelihoods -- one P(t) launch and one
// step-list launch per evaluation (tree.cu, queued calls), maximised by Brent's bounded method.
// ---------------------------------------------------------------------------
static double evalBranch(Tree *t, Node *n, double len, long *nEvals)
{
    n->brLen = len;
    if (nodeCalculateBigPDecks(n)) return NAN;
    for (Node *q = n->parent; q; q = q->parent)
        for (int p = 0; p < t->nParts; p++)
            if (nodeSetCL(q, p)) return NAN;
    double sum = 0.0;
    for (int p = 0; p < t->nParts; p++) {
        const double v = treePartLogLike(t, nullptr, p, 0);
        if (v != v) return NAN;
        sum += v;
    }
    if (nEvals) (*nEvals)++;
    return sum;
}

// Brent's bounded minimiser (golden section + successive parabolic interpolation) of f on [a, b],
// written for maximising the log-likelihood in u = log(branch length).
template <class F>
static double brentMax(F f, double a, double b, double x0, double f0, double tolU, int maxIter, double *fBest, bool *failed)
{
    const double golden = 0.3819660112501051;
    double x = x0, w = x0, v = x0;
    double fx = -f0, fw = fx, fv = fx;     // minimise -lnL
    double d = 0.0, e = 0.0;
    for (int it = 0; it < maxIter; it++) {
        const double xm = 0.5 * (a + b);
        const double tol1 = tolU * fabs(x) + 1e-10, tol2 = 2.0 * tol1;
        if (fabs(x - xm) <= tol2 - 0.5 * (b - a)) break;
        bool useGolden = true;
        if (fabs(e) > tol1) {
            double r = (x - w) * (fx - fv), q = (x - v) * (fx - fw), p = (x - v) * q - (x - w) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) p = -p;
            q = fabs(q);
            const double eOld = e;
            e = d;
            if (!(fabs(p) >= fabs(0.5 * q * eOld) || p <= q * (a - x) || p >= q * (b - x))) {
                d = p / q;
                const double u = x + d;
                if (u - a < tol2 || b - u < tol2) d = xm >= x ? tol1 : -tol1;
                useGolden = false;
            }
        }
        if (useGolden) {
            e = (x >= xm) ? a - x : b - x;
            d = golden * e;
        }
        const double u = fabs(d) >= tol1 ? x + d : x + (d >= 0 ? tol1 : -tol1);
        const double fuRaw = f(u);
        if (fuRaw != fuRaw) { *failed = true; break; }
        const double fu = -fuRaw;
        if (fu <= fx) {
            if (u >= x) a = x; else b = x;
            v = w; fv = fw; w = x; fw = fx; x = u; fx = fu;
        } else {
            if (u < x) a = u; else b = u;
            if (fu <= fw || w == x) { v = w; fv = fw; w = u; fw = fu; }
            else if (fu <= fv || v == x || v == w) { v = u; fv = fu; }
        }
    }
    *fBest = -fx;
    return x;
}

// One or more passes over all branches.  Returns the final log-likelihood (NAN on error).
double optimizeBrLens(Tree *t, int maxPasses, double tol, long *nEvals)
{
    Model *m = t->model;
    const double lo = m->BRLEN_MIN[0], hi = m->BRLEN_MAX[0];
    if (treeSetPrams(t, -1)) return NAN;
    double cur = treeLogLike(t, 0);
    if (cur != cur) return NAN;
    if (nEvals) (*nEvals)++;
    std::vector<Node *> order;       // leaves first, root-most branches last: post-order
    for (int j = 0; j < t->nNodes; j++) {
        const int i = t->postOrder[j];
        if (i == P4B_NO_ORDER) continue;
        Node *n = t->nodes[i];
        if (n && n != t->root) order.push_back(n);
    }
    for (int pass = 0; pass < maxPasses; pass++) {
        const double atStart = cur;
        for (Node *n : order) {
            const double t0 = n->brLen < lo ? lo : (n->brLen > hi ? hi : n->brLen);
            // search a decade either side of the current length (the next pass moves the window)
            const double a = log(t0 / 10.0 > lo ? t0 / 10.0 : lo), b = log(t0 * 10.0 < hi ? t0 * 10.0 : hi);
            double fBest = cur;
            bool failed = false;
            const double u = brentMax([&](double uu) { return evalBranch(t, n, exp(uu), nEvals); }, a, b, log(t0), cur, 1e-4, 40, &fBest, &failed);
            if (failed) return NAN;
            // leave the tree in the best state found (the last point evaluated need not be it)
            const double best = fBest >= cur ? exp(u) : t0;
            const double v = evalBranch(t, n, best, nEvals);
            if (v != v) return NAN;
            cur = v;
        }
        if (cur - atStart < tol) break;
    }
    t->logLike = cur;
    return cur;
}


// ---------------------------------------------------------------------------
// The reference's optimiser entry points
// ---------------------------------------------------------------------------
namespace {
struct OptRun {
    Tree *t;
    int doBrLens;
    long evals = 0;
    bool failed = false;
    // minus the log-likelihood at x: p4_minusLogLikeForBrent (Pf/p4_treeOpt.c:947-985)
    double minusLogLike(double *x)
    {
        if (failed) return 1.0e99;
        unWindParameters(t, doBrLens, x);
        if (g_hitLimit) windUpParameters(t, doBrLens, x, nullptr, nullptr);
        if (treeSetPrams(t, -1)) { failed = true; return 1.0e99; }
        const double v = treeLogLike(t, 0);
        evals++;
        if (v != v) { failed = true; return 1.0e99; }
        return -v;
    }
    // make the model and the tree those of x (the last point evaluated need not be the best one)
    int settle(double *x)
    {
        unWindParameters(t, doBrLens, x);
        return treeSetPrams(t, -1);
    }
};

// evaluations that never read a CL back run in lnL-only mode (Tree::storeCL = 0); restored on the way out
struct LnLOnly {
    Tree *t;
    int saved;
    explicit LnLOnly(Tree *tr) : t(tr), saved(tr->storeCL) { t->storeCL = 0; }
    ~LnLOnly() { t->storeCL = saved; }
};

inline bool isBad(double v) { return v != v; }
const double kNewtFull[4][2] = {{1.0, 10.0}, {1.0e-1, 1.0}, {1.0e-2, 0.1}, {1.0e-5, 1.0e-7}};
}  // namespace

// p4_allBrentPowellOptimize (Pf/p4_treeOpt.c:996-1180): praxis(1e-4, 0.1) over all parameters and branch lengths
// until a call gains less than 1e-6, then the same again with h = 0.05.  Returns the number of evaluations, -1 on error.
long allBrentPowellOptimize(Tree *t)
{
    const int n = countParameters(t, 1);
    if (n <= 0) return 0;
    OptRun run{t, 1};
    std::vector<double> x(n);
    windUpParameters(t, 1, x.data(), nullptr, nullptr);
    Praxis px(n);
    double previous = treeLogLike(t, 0);
    if (previous != previous) return -1;
    run.evals = 1;
    Objective f = [&](double *p) { return run.minusLogLike(p); };
    {
        LnLOnly guard(t);
        double diff = previous;
        while (fabs(diff) > 1.0e-6) {
            const double lnL = -px.minimize(1.0e-4, 0.1, x.data(), f);
            if (run.failed || run.settle(x.data())) return -1;
            diff = lnL - previous;
            previous = lnL;
        }
    }
    if (isBad(treeLogLike(t, 0))) return -1;     // (the reference evaluates the tree here, :1113)
    windUpParameters(t, 1, x.data(), nullptr, nullptr);
    {
        LnLOnly guard(t);
        double diff = 1.0;
        while (fabs(diff) > 1.0e-6) {
            const double lnL = -px.minimize(1.0e-4, 0.05, x.data(), f);
            if (run.failed || run.settle(x.data())) return -1;
            diff = previous - lnL;
            previous = lnL;
        }
    }
    const double last = treeLogLike(t, 0);
    if (last != last) return -1;
    return run.evals + 3;
}

// p4_allBOBYQAOptimize (Pf/p4_treeOpt.c:617-753): two bounded derivative-free maximisations of the objective,
// the second from the re-wound result of the first with the tighter stopping rule (ftol_abs 1e-6, then 1e-8).
long allBoundedOptimize(Tree *t, int doBrLens)
{
    const int n = countParameters(t, doBrLens);
    if (n <= 0) return 0;
    OptRun run{t, doBrLens};
    std::vector<double> x(n), lo(n), hi(n);
    windUpParameters(t, doBrLens, x.data(), lo.data(), hi.data());
    Objective f = [&](double *p) { return run.minusLogLike(p); };
    long evals = 0;
    for (int pass = 0; pass < 2; pass++) {
        {
            LnLOnly guard(t);
            const double fAbs = pass == 0 ? 1.0e-6 : 1.0e-8;
            double scale = fabs(treeLogLike(t, 0));
            if (!(scale > 1.0)) scale = 1.0;
            boundedPowell(n, x.data(), lo.data(), hi.data(), f, 1.0e-7, fAbs / scale, 400L * n + 4000, &evals);
            if (run.failed || run.settle(x.data())) return -1;
        }
        const double v = treeLogLike(t, 0);
        if (v != v) return -1;
        windUpParameters(t, doBrLens, x.data(), lo.data(), hi.data());
    }
    return evals + run.evals;
}

static int newtSchedule(Tree *t, int from, int to)
{
    for (int i = from; i < to; i++) {
        const double v = treeNewtAround(t, kNewtFull[i][0], kNewtFull[i][1]);
        if (v != v) return 1;
    }
    return 0;
}

// p4_newtAndBrentPowellOpt (Pf/p4_treeOpt.c:1182-1330) and p4_newtAndBOBYQAOpt (:755-945): branch lengths by
// Newton-Raphson (p4_newtAround), the free model parameters by praxis (bounded = 0) or the bounded method
// (bounded = 1), alternating until a round gains less than 1e-6 or the pass limit is reached.
long newtAndModelOpt(Tree *t, int bounded)
{
    if (treeNewtSetup(t)) return -1;
    const int n = countParameters(t, 0);
    if (n == 0) return newtSchedule(t, 0, 4) ? -1 : 0;          // :1214-1226
    OptRun run{t, 0};
    std::vector<double> x(n), lo(n), hi(n);
    windUpParameters(t, 0, x.data(), lo.data(), hi.data());
    Objective f = [&](double *p) { return run.minusLogLike(p); };
    long evals = 0;
    if (n == 1 && !bounded) {
        // p4_newtAnd1DBrent (:1395-1436): Newton rounds, then the one parameter by Brent's one-dimensional
        // minimiser, alternating.  The reference brackets the minimum first and then runs LocalMin; here the
        // parameter's own bounds are the bracket.
        if (newtSchedule(t, 0, 3)) return -1;
        double before = 0.0;
        for (int iter = 0; iter <= 21; iter++) {
            if (iter > 0 && isBad(treeNewtAround(t, 1.0e-5, 1.0e-7))) return -1;
            windUpParameters(t, 0, x.data(), lo.data(), hi.data());
            double lnL;
            {
                LnLOnly guard(t);
                lnL = -boundedPowell(1, x.data(), lo.data(), hi.data(), f, 1.0e-7, 1.0e-14, 400, &evals);
                if (run.failed || run.settle(x.data())) return -1;
            }
            if (iter > 0 && fabs(lnL - before) < 1.0e-6) break;
            before = lnL;
        }
        return isBad(treeLogLike(t, 0)) ? -1 : evals + run.evals;
    }
    if (isBad(treeNewtAround(t, 1.0, 10.0))) return -1;       // :1242-1244 / :800-802
    {
        const double a = treeNewtAround(t, 1.0e-1, 1.0), b = treeNewtAround(t, 1.0e-5, 1.0e-7);
        if (a != a || b != b) return -1;
    }
    Praxis px(n);
    double previous = treeLogLike(t, 0);
    if (previous != previous) return -1;
    const int limit = bounded ? 50 : ((t->passLimit && t->passLimit[0] > 0) ? t->passLimit[0] : 50);
    for (int pass = 0;; ) {
        if (isBad(treeNewtAround(t, 1.0e-5, 1.0e-7))) return -1;
        double lnL;
        {
            LnLOnly guard(t);
            if (bounded) {
                double scale = fabs(previous) > 1.0 ? fabs(previous) : 1.0;
                lnL = -boundedPowell(n, x.data(), lo.data(), hi.data(), f, 1.0e-7, 1.0e-8 / scale, 400L * n + 4000, &evals);
            } else {
                lnL = -px.minimize(1.0e-4, 1.0, x.data(), f);
            }
            if (run.failed || run.settle(x.data())) return -1;
        }
        const double diff = lnL - previous;
        previous = lnL;
        if (fabs(diff) < 1.0e-6) break;
        if (++pass > limit) break;       // the reference prints "Pass limit exceeded without convergence" and stops too
    }
    const double last = treeLogLike(t, 0);
    if (last != last) return -1;
    return evals + run.evals;
}

}  // namespace p4b

using namespace p4b;

extern "C" {

int p4b_countParameters(p4b_tree t, int doBrLens)
{
    if (!t) { setError("p4b_countParameters: NULL handle"); return -1; }
    return countParameters((Tree *)t, doBrLens);
}
int p4b_windUpParameters(p4b_tree t, int doBrLens, double *x, double *lb, double *ub)
{
    if (!t || !x) { setError("p4b_windUpParameters: NULL argument"); return -1; }
    return windUpParameters((Tree *)t, doBrLens, x, lb, ub);
}
int p4b_unWindParameters(p4b_tree t, int doBrLens, const double *x)
{
    if (!t || !x) { setError("p4b_unWindParameters: NULL argument"); return -1; }
    return unWindParameters((Tree *)t, doBrLens, x);
}
double p4b_logLikeForParameters(p4b_tree t, int doBrLens, const double *x)
{
    if (!t || !x) { setError("p4b_logLikeForParameters: NULL argument"); return NAN; }
    Tree *T = (Tree *)t;
    unWindParameters(T, doBrLens, x);
    if (treeSetPrams(T, -1)) return NAN;
    return treeLogLike(T, 0);
}
double p4b_optimizeBrLens(p4b_tree t, int maxPasses, double tol, long *nEvals)
{
    if (!t) { setError("p4b_optimizeBrLens: NULL handle"); return NAN; }
    if (nEvals) *nEvals = 0;
    return optimizeBrLens((Tree *)t, maxPasses < 1 ? 1 : maxPasses, tol, nEvals);
}
long p4b_allBrentPowellOptimize(p4b_tree t)
{
    if (!t) { setError("p4_allBrentPowellOptimize: NULL handle"); return -1; }
    return allBrentPowellOptimize((Tree *)t);
}
long p4b_allBOBYQAOptimize(p4b_tree t, int doBrLens)
{
    if (!t) { setError("p4_allBOBYQAOptimize: NULL handle"); return -1; }
    return allBoundedOptimize((Tree *)t, doBrLens);
}
long p4b_newtAndBrentPowellOpt(p4b_tree t)
{
    if (!t) { setError("p4_newtAndBrentPowellOpt: NULL handle"); return -1; }
    return newtAndModelOpt((Tree *)t, 0);
}
long p4b_newtAndBOBYQAOpt(p4b_tree t)
{
    if (!t) { setError("p4_newtAndBOBYQAOpt: NULL handle"); return -1; }
    return newtAndModelOpt((Tree *)t, 1);
}
// The minimisers themselves, on a caller's objective (tests drive them with analytic functions on the CPU).
double p4b_praxisMinimize(int n, double *x, double tol, double h, double (*fn)(const double *, void *), void *ctx)
{
    if (n < 1 || !x || !fn) { setError("p4b_praxisMinimize: bad argument"); return NAN; }
    Praxis px(n);
    return px.minimize(tol, h, x, [&](double *p) { return fn(p, ctx); });
}
double p4b_boundedMinimize(int n, double *x, const double *lo, const double *hi, double xtol, double ftol, long maxEvals,
                           double (*fn)(const double *, void *), void *ctx, long *nEvals)
{
    if (n < 1 || !x || !lo || !hi || !fn) { setError("p4b_boundedMinimize: bad argument"); return NAN; }
    long e = 0;
    const double v = boundedPowell(n, x, lo, hi, [&](double *p) { return fn(p, ctx); }, xtol, ftol, maxEvals, &e);
    if (nEvals) *nEvals = e;
    return v;
}
int p4b_newtSetup(p4b_tree t)
{
    if (!t) { setError("p4b_newtSetup: NULL handle"); return 1; }
    return treeNewtSetup((Tree *)t);
}
double p4b_newtAround(p4b_tree t, double epsilon, double likeDelta)
{
    if (!t) { setError("p4b_newtAround: NULL handle"); return NAN; }
    return treeNewtAround((Tree *)t, epsilon, likeDelta);
}
int p4b_newtDerivs(p4b_node n, double out3[3])
{
    if (!n || !out3) { setError("p4b_newtDerivs: NULL argument"); return 1; }
    return nodeNewtDerivs((Node *)n, out3);
}
int p4b_getNodeCL2(p4b_node n, int pNum, double *out)
{
    if (!n || !out) { setError("p4b_getNodeCL2: NULL argument"); return 1; }
    return nodeGetCL2((Node *)n, pNum, out);
}
long long p4b_newtIterations(p4b_tree t) { return t ? treeNewtIterations((Tree *)t) : 0; }
int p4b_treePassLimit(p4b_tree t)
{
    Tree *T = (Tree *)t;
    return (T && T->passLimit) ? T->passLimit[0] : 50;
}
int p4b_treeNNodes(p4b_tree t) { return t ? ((Tree *)t)->nNodes : -1; }
int p4b_treeNLeaves(p4b_tree t) { return t ? ((Tree *)t)->nLeaves : -1; }
int p4b_treeNParts(p4b_tree t) { return t ? ((Tree *)t)->nParts : -1; }
int p4b_treePartDim(p4b_tree t, int pNum)
{
    Tree *T = (Tree *)t;
    return (T && T->model && pNum >= 0 && pNum < T->nParts) ? T->model->parts[pNum]->dim : -1;
}
int p4b_getBrLens(p4b_tree t, double *out)
{
    if (!t || !out) { setError("p4b_getBrLens: NULL argument"); return 1; }
    Tree *T = (Tree *)t;
    for (int i = 0; i < T->nNodes; i++) out[i] = (T->nodes[i] && T->nodes[i] != T->root) ? T->nodes[i]->brLen : -1.0;
    return 0;
}

}  // extern "C"
