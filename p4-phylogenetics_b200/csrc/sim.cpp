// sim.cpp -- consumers of the P decks beside the likelihood: expected tip
// compositions, and simulation of data down the tree (p4_simulate) on the caller's random stream.
//
// p4_expectedComposition / p4_expectedCompositionCounts (Pf/p4_treeSim.c:859-1045) push the root's
// composition down the tree through every branch's transition matrices (p4_calculateExpectedComp,
// Pf/p4_node.c:1038-1250): the composition a tree-heterogeneous model expects at each tip, which
// Tree.compoTestUsingSimulations and the model-fit tests of p4 compare with the observed one
// (p4/tree.py:8192, 8544, 9071-9180).  The P decks are the ones resident on the device; the
// recursion itself is dim-sized per node and runs on the host in the reference's order of operations.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <utility>
#include <vector>

#include "../../include/p4b200.h"
#include "engine.h"

namespace p4b {

// expected[node][cat][state] for part p, nodes in preOrder; returns 0 on success.
static int expectedComp(Tree *t, int p, std::vector<double> &e)
{
    ModelPart *mp = t->model->parts[p];
    const int dim = mp->dim, nCat = mp->nCat;
    e.assign((size_t)t->nNodes * nCat * dim, 0.0);
    if (!t->root) { setError("p4_expectedComposition: the tree has no root"); return 1; }
    const int rc = t->root->compNums[p];
    if (rc < 0 || rc >= mp->nComps || !mp->comps[rc].val) { setError("root uses comp %d which does not exist", rc); return 1; }
    std::vector<double> P((size_t)nCat * dim * dim), invar(dim);
    for (int j = 0; j < t->nNodes; j++) {
        const int i = t->preOrder[j];
        if (i == P4B_NO_ORDER) continue;
        if (i < 0 || i >= (int)t->nodes.size() || !t->nodes[i]) { setError("preOrder[%d] = %d is not a node", j, i); return 1; }
        Node *n = t->nodes[i];
        double *mine = &e[(size_t)i * nCat * dim];
        if (n == t->root) {                                  // Pf/p4_node.c:1063-1073
            for (int c = 0; c < nCat; c++)
                for (int s = 0; s < dim; s++) mine[c * dim + s] = mp->comps[rc].val[s];
            continue;
        }
        if (!n->parent) { setError("node %d has no parent", n->nodeNum); return 1; }
        const double *par = &e[(size_t)n->parent->nodeNum * nCat * dim];
        const double *rootE = &e[(size_t)t->root->nodeNum * nCat * dim];
        if (nodeGetBigP(n, p, P.data())) return 1;
        const bool inv = mp->pInvar > 0.0;                   // :1107-1119; pInvar is -1 until set
        for (int s = 0; s < dim; s++) invar[s] = inv ? mp->pInvar * rootE[s] : 0.0;
        for (int c = 0; c < nCat; c++) {
            double *out = mine + c * dim;
            // picker[i][j] = P[i][j] * (parent[i] - invar[i]); expected[j] = sum_i picker[i][j]   (:1121-1164)
            for (int a = 0; a < dim; a++) {
                const double w = par[c * dim + a] - invar[a];
                for (int b = 0; b < dim; b++) out[b] += P[((size_t)c * dim + a) * dim + b] * w;
            }
            double factor = 0.0;                             // :1175-1185
            for (int s = 0; s < dim; s++) factor += out[s];
            for (int s = 0; s < dim; s++) out[s] /= factor;
            if (inv) {                                       // :1187-1206
                const double f = 1.0 - mp->pInvar;
                for (int s = 0; s < dim; s++) out[s] *= f;
                for (int s = 0; s < dim; s++) out[s] += invar[s];
            }
        }
    }
    return 0;
}

// out[seqNum][state]: mean over the rate categories at every leaf (Pf/p4_treeSim.c:986-1041); with
// counts != 0 multiplied by the number of non-gap, non-'?' sites of that sequence (:907-944).
int treeExpectedComposition(Tree *t, int p, int counts, double *out)
{
    if (!t->dev) { setError("tree has no device state"); return 1; }
    if (p < 0 || p >= t->nParts) { setError("p4_expectedComposition: bad part %d", p); return 1; }
    std::vector<double> e;
    if (expectedComp(t, p, e)) return 1;
    ModelPart *mp = t->model->parts[p];
    Part *dp = t->data->parts[p];
    const int dim = mp->dim, nCat = mp->nCat;
    for (Node *n : t->nodes) {
        if (!n || !n->isLeaf) continue;
        if (n->seqNum < 0 || n->seqNum >= dp->nTax) { setError("leaf node %d has seqNum %d", n->nodeNum, n->seqNum); return 1; }
        double scale = 1.0;
        if (counts) {
            int nGaps = 0;
            const int *seq = &dp->sequences[(size_t)n->seqNum * dp->nChar];
            for (int k = 0; k < dp->nChar; k++)
                if (seq[k] == P4B_GAP_CODE || seq[k] == P4B_QMARK_CODE) nGaps++;
            scale = (double)(dp->nChar - nGaps);
        }
        for (int s = 0; s < dim; s++) {
            double v = 0.0;
            for (int c = 0; c < nCat; c++) v += e[((size_t)n->nodeNum * nCat + c) * dim + s];
            v /= (double)nCat;
            out[(size_t)n->seqNum * dim + s] = counts ? scale * v : v;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------
// Random stream.  p4 hands its simulations a GSL generator (var.gsl_rng = pf.gsl_rng_get(), the library default
// mt19937; Pf/pfmodule.c:674-749) and the reference consumes one gsl_rng_uniform per decision in a fixed order
// (Pf/p4_treeSim.c:235-340).  Drawing the same stream here makes a simulation a pure function of (tree, model,
// seed) that can be compared with the reference's site by site.  MT19937 (Matsumoto & Nishimura 1998) with GSL's
// seeding: seed 0 means 4357; uniform = 32-bit output / 2^32.
// ---------------------------------------------------------------------------
struct Rng {
    uint32_t mt[624];
    int mti = 625;
    void set(unsigned long seed)
    {
        if (seed == 0) seed = 4357;
        mt[0] = (uint32_t)(seed & 0xffffffffUL);
        for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        mti = 624;
    }
    uint32_t get()
    {
        if (mti >= 624) reload();
        uint32_t y = mt[mti++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    double uniform() { return get() / 4294967296.0; }
    void reload()
    {
        if (mti == 625) set(0);
        static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
        int kk = 0;
        for (; kk < 624 - 397; kk++) {
            const uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1u];
        }
        for (; kk < 623; kk++) {
            const uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u];
        }
        const uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ mag01[y & 1u];
        mti = 0;
    }
    // The next n uniforms of the stream, a state block at a time: the tempering of a block is a straight loop over
    // the state words (no call, no branch per draw), which is what the simulation's 4e8 draws spend their time in.
    void fillUniform(double *dst, size_t n)
    {
        size_t done = 0;
        while (done < n) {
            if (mti >= 624) reload();
            size_t take = (size_t)(624 - mti);
            if (take > n - done) take = n - done;
            const uint32_t *src = mt + mti;
            double *out = dst + done;
            for (size_t i = 0; i < take; i++) {
                uint32_t y = src[i];
                y ^= (y >> 11);
                y ^= (y << 7) & 0x9d2c5680u;
                y ^= (y << 15) & 0xefc60000u;
                y ^= (y >> 18);
                out[i] = y / 4294967296.0;
            }
            mti += (int)take;
            done += take;
        }
    }
    // gsl_rng_uniform_int for a generator with min 0 and max 2^32 - 1: equal-width bins, draws beyond the last bin rejected
    unsigned long uniformInt(unsigned long n)
    {
        const unsigned long range = 0xffffffffUL;
        if (n == 0 || n > range) return 0;
        const unsigned long scale = range / n;
        unsigned long k;
        do { k = get() / scale; } while (k >= n);
        return k;
    }
};

// p4_drawAncState (Pf/p4_treeSim.c:591-845): one draw of (root state, rate category | invariant) for one site from
// the posterior the root's conditional likelihoods imply.  The reference draws from the C library's random()
// (seeded through pf.reseedCRandomizer -> srandom), not from the GSL stream; so does this.
// draw = {chStNum, catNum, isInvar, invarChNum} as p4_drawAncStateP returns them (:847-857).
// The draw itself, given the root's CL as [cat*dim + state][pattern] with row stride `stride`.
static int drawFromRootCL(const double *cl, size_t stride, Part *dp, int p, int seqPos, int dim, int nCat, double pInvar, int pInvarFree,
                          const double *pi, int draw[4])
{
    if (seqPos < 0 || seqPos >= dp->nChar) { setError("p4_drawAncState: bad site %d", seqPos); return 1; }
    const int patNum = dp->sequencePositionPatternIndex[seqPos];
    const size_t ps = stride;
    // running total of the site's likelihood, term by term in the reference's order: first every (category, state) of
    // the variable-site mixture, then -- pInvar set -- the invariant share of every state the site could be constant for
    const double fr = (1.0 - pInvar) / (double)nCat;     // freqsTimesOneMinusPInvar, Pf/p4_tree.c:992-996
    std::vector<double> upTo((size_t)nCat * dim);
    std::vector<std::pair<int, double>> invUpTo;
    double total = 0.0;
    for (int c = 0; c < nCat; c++)
        for (int s = 0; s < dim; s++) {
            double term = pi[s] * cl[((size_t)c * dim + s) * ps + patNum];
            term *= fr;
            total += term;
            upTo[(size_t)c * dim + s] = total;
        }
    if (pInvar != 0.0) {
        if (dp->globalInvarSitesArray.empty()) { setError("pInvar is set but pf.setGlobalInvarSitesVec was not called on part %d", p); return 1; }
        for (int s = 0; s < dim; s++)
            if (dp->globalInvarSitesArray[(size_t)s * dp->nChar + patNum]) {
                total += pi[s] * pInvar;
                invUpTo.emplace_back(s, total);
            }
    }
    double u = (double)random() / ((double)(RAND_MAX) + 1.0);
    u *= total;
    for (size_t i = 0; i < upTo.size(); i++)
        if (u < upTo[i]) {
            draw[0] = (int)(i % dim);
            draw[1] = (int)(i / dim);
            draw[2] = 0;
            draw[3] = -1;
            return 0;
        }
    if (pInvarFree)                                      // :780-792: only a FREE pInvar's share is looked at
        for (const auto &sv : invUpTo)
            if (u < sv.second) {
                draw[0] = -1;
                draw[1] = -1;
                draw[2] = 1;
                draw[3] = sv.first;
                return 0;
            }
    setError("Something is wrong with the ancestral state picker. gotIt is zero.");   // the reference's message (:794-797)
    return 1;
}

int treeDrawAncState(Tree *t, int p, int seqPos, int draw[4])
{
    int ps = 0;
    const double *cl = treeRootCLHost(t, p, &ps);
    if (!cl) return 1;
    ModelPart *mp = t->model->parts[p];
    const int rc = t->root->compNums[p];
    if (rc < 0 || rc >= mp->nComps || !mp->comps[rc].val) { setError("root uses comp %d which does not exist", rc); return 1; }
    return drawFromRootCL(cl, (size_t)ps, t->data->parts[p], p, seqPos, mp->dim, mp->nCat, mp->pInvar, mp->pInvarFree, mp->comps[rc].val, draw);
}

// p4_simulate(t, refTree, g), Pf/p4_treeSim.c:14-420.  Host: the per-site draws that need no tree (rate category,
// root state, invariant or not -- :235-300, in the reference's order over parts) and the stream of uniforms of the
// mutation phase, which the reference consumes node by node in preOrder and, within a node, site by site over the
// VARIABLE sites only (:330-360).  Device (tree.cu): picker decks (running row sums of the P decks,
// Pf/p4_node.c p4_calculatePickerDecks) and one thread per site walking the nodes.
int treeSimulate(Tree *t, Tree *refTree, Rng *g)
{
    if (!t->dev) { setError("tree has no device state"); return 1; }
    if (!t->root) { setError("p4_simulate: the tree has no root"); return 1; }
    Data *d = t->data;
    const int nParts = d->nParts;
    std::vector<std::vector<uint8_t>> cats(nParts), rootSt(nParts), inv(nParts);
    if (refTree) {
        // :200-232: root state, rate category and invariant-or-not of every site drawn from the posterior at the
        // root of refTree (p4_drawAncState, the C library's random()), before this tree's data are touched
        if (refTree->data == d) { setError("p4_simulate: refTree must have its own data"); return 1; }
        if (refTree->data->nParts != nParts) { setError("p4_simulate: refTree has %d parts, the tree %d", refTree->data->nParts, nParts); return 1; }
        for (int p = 0; p < nParts; p++) {
            Part *dp = d->parts[p], *rp = refTree->data->parts[p];
            if (rp->nChar != dp->nChar) { setError("p4_simulate: part %d of refTree has %d sites, the tree's %d", p, rp->nChar, dp->nChar); return 1; }
            cats[p].assign(dp->nChar, 0);
            rootSt[p].assign(dp->nChar, 0);
            inv[p].assign(dp->nChar, 0);
            for (int k = 0; k < rp->nChar; k++) {
                int draw[4];
                if (treeDrawAncState(refTree, p, k, draw)) return 1;
                if (draw[2]) { rootSt[p][k] = (uint8_t)draw[3]; inv[p][k] = 1; }
                else { cats[p][k] = (uint8_t)draw[1]; rootSt[p][k] = (uint8_t)draw[0]; }
            }
        }
    }
    for (int p = 0; p < nParts; p++) d->parts[p]->nPatterns = 0;                       // :60-62
    if (refTree) {
        for (int p = 0; p < nParts; p++) {
            Part *dp = d->parts[p];
            dp->globalInvarSitesVec.assign(dp->nChar, 0);
            for (int k = 0; k < dp->nChar; k++) dp->globalInvarSitesVec[k] = inv[p][k];
        }
    } else {
        for (int p = 0; p < nParts; p++) {                                                 // rate categories, :235-246
            Part *dp = d->parts[p];
            ModelPart *mp = t->model->parts[p];
            if (mp->dim > 255 || mp->nCat > 255) { setError("p4_simulate: dim or nCat too large"); return 1; }
            cats[p].assign(dp->nChar, 0);
            if (mp->nCat > 1)
                for (int i = 0; i < dp->nChar; i++) cats[p][i] = (uint8_t)(int)floor(((double)mp->nCat) * g->uniform());
        }
        for (int p = 0; p < nParts; p++) {                                                 // root states, :256-277
            Part *dp = d->parts[p];
            ModelPart *mp = t->model->parts[p];
            const int rc = t->root->compNums[p];
            if (rc < 0 || rc >= mp->nComps || !mp->comps[rc].val) { setError("root uses comp %d which does not exist", rc); return 1; }
            const double *pi = mp->comps[rc].val;
            std::vector<double> picker(mp->dim);
            picker[0] = pi[0];
            for (int j = 1; j < mp->dim - 1; j++) picker[j] = picker[j - 1] + pi[j];
            picker[mp->dim - 1] = 1.0;
            rootSt[p].assign(dp->nChar, 0);
            for (int j = 0; j < dp->nChar; j++) {
                const double u = g->uniform();
                for (int k = 0; k < mp->dim; k++)
                    if (u < picker[k]) { rootSt[p][j] = (uint8_t)k; break; }
            }
        }
        for (int p = 0; p < nParts; p++) {                                                 // invariant sites, :284-300
            Part *dp = d->parts[p];
            ModelPart *mp = t->model->parts[p];
            inv[p].assign(dp->nChar, 0);
            dp->globalInvarSitesVec.assign(dp->nChar, 0);
            if (mp->pInvar > 0.0)
                for (int i = 0; i < dp->nChar; i++) {
                    const double u = g->uniform();
                    inv[p][i] = u < mp->pInvar ? 1 : 0;
                    dp->globalInvarSitesVec[i] = inv[p][i];
                }
        }
    }
    if (treeCalculateAllBigPDecks(t)) return 1;                                        // :315-321
    for (int p = 0; p < nParts; p++) {                                                 // mutation, :330-360
        Part *dp = d->parts[p];
        std::vector<int> rank(dp->nChar, 0);
        int nVar = 0;
        for (int k = 0; k < dp->nChar; k++) {
            rank[k] = nVar;
            if (!inv[p][k]) nVar++;
        }
        auto fill = [&](double *dst, size_t n) { g->fillUniform(dst, n); };
        if (treeSimulateDevice(t, p, cats[p].data(), rootSt[p].data(), inv[p].data(), rank.data(), nVar, fill)) return 1;
        dp->version++;
    }
    return 0;
}

}  // namespace p4b

using namespace p4b;

extern "C" {

void *p4b_rngNew(void) { return new Rng(); }
void p4b_rngFree(void *g) { delete (Rng *)g; }
void p4b_rngSet(void *g, unsigned long seed)
{
    if (g) ((Rng *)g)->set(seed);
}
unsigned long p4b_rngGet(void *g) { return g ? ((Rng *)g)->get() : 0ul; }
double p4b_rngUniform(void *g) { return g ? ((Rng *)g)->uniform() : 0.0; }
void p4b_rngFillUniform(void *g, double *out, long n)
{
    if (g && out && n > 0) ((Rng *)g)->fillUniform(out, (size_t)n);
}
// ---- the random draws and special functions p4's MCMC proposals take from GSL through pf (Pf/pfmodule.c:724-1057) ----
// GSL is a third-party dependency of the reference (absent here).  The draws below are the published algorithms GSL
// documents for them -- gamma by Marsaglia & Tsang (2000) on a polar-method normal deviate, Dirichlet as normalised
// gammas -- on this file's mt19937 stream; densities and lnGamma through libm.  They are host code beside the path:
// Chain's proposals call them between likelihood evaluations (p4/chain.py:2274-2565).
static double rngUniformPos(Rng *r)
{
    double x;
    do { x = r->uniform(); } while (x == 0.0);
    return x;
}
static double rngGaussian(Rng *r)
{
    double x, y, r2;
    do {
        x = -1 + 2 * rngUniformPos(r);
        y = -1 + 2 * rngUniformPos(r);
        r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0);
    return y * sqrt(-2.0 * log(r2) / r2);
}
static double rngGamma(Rng *r, double a, double b)
{
    if (a < 1) {
        const double u = rngUniformPos(r);
        return rngGamma(r, 1.0 + a, b) * pow(u, 1.0 / a);
    }
    const double d = a - 1.0 / 3.0, c = (1.0 / 3.0) / sqrt(d);
    double x, v, u;
    for (;;) {
        do { x = rngGaussian(r); v = 1.0 + c * x; } while (v <= 0);
        v = v * v * v;
        u = rngUniformPos(r);
        if (u < 1 - 0.0331 * x * x * x * x) break;
        if (log(u) < 0.5 * x * x + d * (1 - v + log(v))) break;
    }
    return b * d * v;
}
long p4b_rngSize(void *g) { (void)g; return (long)sizeof(Rng); }                       /* pf.gsl_rng_size :724 */
void p4b_rngGetState(void *g, void *buf) { if (g && buf) memcpy(buf, g, sizeof(Rng)); }   /* pf.gsl_rng_getstate :752 */
void p4b_rngSetState(void *g, const void *buf) { if (g && buf) memcpy(g, buf, sizeof(Rng)); }   /* pf.gsl_rng_setstate :779 */
double p4b_ranGamma(void *g, double a, double b) { return g ? rngGamma((Rng *)g, a, b) : NAN; }   /* pf.gsl_ran_gamma :827 */
void p4b_ranDirichlet(void *g, int k, const double *alpha, double *theta)              /* pf.gsl_ran_dirichlet :940 */
{
    if (!g || !alpha || !theta) return;
    double norm = 0.0;
    for (int i = 0; i < k; i++) theta[i] = rngGamma((Rng *)g, alpha[i], 1.0);
    for (int i = 0; i < k; i++) norm += theta[i];
    for (int i = 0; i < k; i++) theta[i] /= norm;
}
double p4b_ranDirichletLnPdf(int k, const double *alpha, const double *theta)          /* pf.gsl_ran_dirichlet_lnpdf :989 */
{
    double logP = 0.0, sumAlpha = 0.0;
    for (int i = 0; i < k; i++) logP += (alpha[i] - 1.0) * log(theta[i]);
    for (int i = 0; i < k; i++) sumAlpha += alpha[i];
    logP += lgamma(sumAlpha);
    for (int i = 0; i < k; i++) logP -= lgamma(alpha[i]);
    return logP;
}
double p4b_sfLnGamma(double x) { return lgamma(x); }                                   /* pf.gsl_sf_lngamma :887 */
double p4b_ranGammaPdf(double x, double a, double b)                                   /* pf.gsl_ran_gamma_pdf :848 */
{
    if (x < 0) return 0;
    if (x == 0) return (a == 1) ? 1 / b : 0;
    if (a == 1) return exp(-x / b) / b;
    return exp((a - 1) * log(x / b) - x / b - lgamma(a)) / b;
}
void p4b_meanVariance(const double *seq, int n, double *mean, double *variance)        /* pf.gsl_meanVariance :1025 */
{
    long double m = 0;
    for (int i = 0; i < n; i++) m += (seq[i] - m) / (i + 1);
    long double v = 0;
    for (int i = 0; i < n; i++) {
        const long double delta = seq[i] - (double)m;
        v += (delta * delta - v) / (i + 1);
    }
    *mean = (double)m;
    *variance = (double)(v * ((double)n / (double)(n - 1)));
}

int p4b_simulate(p4b_tree t, p4b_tree refTree, void *rng)
{
    if (!t || !rng) { setError("p4b_simulate: NULL argument"); return 1; }
    return treeSimulate((Tree *)t, (Tree *)refTree, (Rng *)rng);
}
/* pf.bootstrapData(referenceData, toFillData, gsl_rng) Pf/pfmodule.c:90 -> bootstrapData Pf/data.c:107-139: every site of every
 * part of toFill becomes a site of the reference part drawn with gsl_rng_uniform_int on the caller's stream, then makePatterns. */
int p4b_bootstrapData(p4b_data reference, p4b_data toFill, void *rng)
{
    Data *ref = (Data *)reference, *out = (Data *)toFill;
    Rng *g = (Rng *)rng;
    if (!ref || !out || !g) { setError("bootstrapData: NULL argument"); return 1; }
    if (ref->nParts < out->nParts) { setError("bootstrapData: the reference has %d parts, the data to fill %d", ref->nParts, out->nParts); return 1; }
    for (int p = 0; p < out->nParts; p++) {
        Part *rp = ref->parts[p], *fp = out->parts[p];
        if (!rp || !fp) { setError("bootstrapData: part %d missing", p); return 1; }
        if (rp->nTax != fp->nTax || rp->nChar < fp->nChar) { setError("bootstrapData: part %d of the two data objects differ in shape", p); return 1; }
        for (int pos = 0; pos < fp->nChar; pos++) {
            const int ran = (int)g->uniformInt((unsigned long)fp->nChar);
            for (int t = 0; t < fp->nTax; t++) fp->sequences[(size_t)t * fp->nChar + pos] = rp->sequences[(size_t)t * rp->nChar + ran];
        }
        if (makePatterns(fp)) return 1;
    }
    return 0;
}
/* Test hook: the draw of p4b_drawAncState for a root CL handed in by the caller ([cat][state][pattern], nPatterns columns),
 * so that the host logic can be checked against the reference without a device. */
int p4b_drawAncStateFromCL(p4b_part part, int seqPos, int nCat, double pInvar, int pInvarFree, const double *pi, const double *rootCL, int *draw4)
{
    Part *dp = (Part *)part;
    if (!dp || !pi || !rootCL || !draw4) { setError("p4b_drawAncStateFromCL: NULL argument"); return 1; }
    return drawFromRootCL(rootCL, (size_t)dp->nPatterns, dp, 0, seqPos, dp->dim, nCat, pInvar, pInvarFree, pi, draw4);
}
int p4b_drawAncState(p4b_tree t, int pNum, int seqPos, int *draw4)
{
    if (!t || !draw4) { setError("p4b_drawAncState: NULL argument"); return 1; }
    return treeDrawAncState((Tree *)t, pNum, seqPos, draw4);
}
void p4b_reseedCRandomizer(int seed) { srandom((unsigned)seed); }
int p4b_expectedComposition(p4b_tree t, int pNum, double *out)
{
    if (!t || !out) { setError("p4b_expectedComposition: NULL argument"); return 1; }
    return treeExpectedComposition((Tree *)t, pNum, 0, out);
}
int p4b_expectedCompositionCounts(p4b_tree t, int pNum, double *out)
{
    if (!t || !out) { setError("p4b_expectedCompositionCounts: NULL argument"); return 1; }
    return treeExpectedComposition((Tree *)t, pNum, 1, out);
}

}  // extern "C"
