// opt.cpp -- the optimiser's seam onto the likelihood path.
//
// The reference's optimisers (Pf/p4_treeOpt.c) pack every free parameter and, optionally, every
// branch length into one vector (p4_windUpParameters :17-120), and their objective unpacks it,
// rebuilds the model and evaluates the tree (p4_logLikeForNLOpt :579-615:
// p4_unWindParameters -> p4_setPrams -> p4_treeLogLike).  This file provides the same packing,
// unpacking and objective on the B200 engine, so that any bounded optimiser can drive the GPU
// with the reference's own parameterisation.  The optimiser itself is not here: nlopt's BOBYQA
// is a third-party library the reference links (absent from this image); pf.py drives this
// objective with a bounded derivative-free method instead.
//
// Unpacking follows p4_unWindParameters (:186-565) with one simplification: where the reference
// nudges an out-of-bounds rate by a random amount (RATE_MIN * ranDoubleUpToOne(), :356, :373) the
// value is clamped, so the objective is a deterministic function of the vector.
#include <cmath>
#include <cstring>

#include "../../include/p4b200.h"
#include "engine.h"

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

int unWindParameters(Tree *t, int doBrLens, const double *x)
{
    Model *m = t->model;
    int pos = 0;
    for (int p = 0; p < m->nParts; p++) {
        ModelPart *mp = m->parts[p];
        const int dim = mp->dim;
        const double pmin = m->PIVEC_MIN[0];
        for (auto &c : mp->comps)
            if (c.isFree) {   // :212-247: shift by PIVEC_MIN, last value by difference, rescale
                double sum = 0.0;
                for (int i = 0; i < dim - 1; i++) {
                    double par = x[pos++];
                    par = par <= pmin ? 0.0 : par - pmin;
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
                    r.kappa = clampd(x[pos++], m->KAPPA_MIN[0], m->KAPPA_MAX[0]);
                    setKappaBigR(r);
                } else if (m->rMatrixNormalizeTo1 && m->rMatrixNormalizeTo1[0]) {   // :343-404
                    double sum = 0.0;
                    for (int i = 0; i < dim - 2; i++)
                        for (int j = i + 1; j < dim; j++) {
                            const double par = clampd(x[pos++], m->RATE_MIN[0], 0.999);
                            r.bigR[i * dim + j] = par;
                            sum += par;
                        }
                    double last = sum < 1.0 ? clampd(1.0 - sum, m->RATE_MIN[0], 0.999) : m->RATE_MIN[0] * sum;
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
                            const double par = clampd(x[pos++], m->RATE_MIN[0], m->RATE_MAX[0]);
                            r.bigR[i * dim + j] = r.bigR[j * dim + i] = par;
                        }
                }
            }
        for (Gdasrv *g : mp->gdasrvs)
            if (g && g->isFree) g->val[0] = clampd(x[pos++], m->GAMMA_SHAPE_MIN[0], m->GAMMA_SHAPE_MAX[0]);
        if (mp->pInvarFree) mp->pInvar = clampd(x[pos++], m->PINVAR_MIN[0], m->PINVAR_MAX[0]);
    }
    if (m->relRatesAreFree) {   // :521-545: the last part's rate keeps the site-weighted mean at 1
        for (int p = 0; p < m->nParts - 1; p++) m->parts[p]->relRate = clampd(x[pos++], m->RELRATE_MIN[0], m->RELRATE_MAX[0]);
        long totLen = 0;
        for (int p = 0; p < m->nParts; p++) totLen += t->data->parts[p]->nChar;
        double sum = 0.0;
        for (int p = 0; p < m->nParts - 1; p++) sum = sum + (m->parts[p]->relRate * t->data->parts[p]->nChar);
        m->parts[m->nParts - 1]->relRate = (((double)totLen) - sum) / t->data->parts[m->nParts - 1]->nChar;
    }
    if (doBrLens)
        for (Node *nd : t->nodes)
            if (nd && nd != t->root) nd->brLen = clampd(x[pos++], m->BRLEN_MIN[0], m->BRLEN_MAX[0]);
    return pos;
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
int p4b_treeNNodes(p4b_tree t) { return t ? ((Tree *)t)->nNodes : -1; }
int p4b_getBrLens(p4b_tree t, double *out)
{
    if (!t || !out) { setError("p4b_getBrLens: NULL argument"); return 1; }
    Tree *T = (Tree *)t;
    for (int i = 0; i < T->nNodes; i++) out[i] = (T->nodes[i] && T->nodes[i] != T->root) ? T->nodes[i]->brLen : -1.0;
    return 0;
}

}  // extern "C"
