// sim.cpp -- consumers of the P decks beside the likelihood.
//
// p4_expectedComposition / p4_expectedCompositionCounts (Pf/p4_treeSim.c:859-1045) push the root's
// composition down the tree through every branch's transition matrices (p4_calculateExpectedComp,
// Pf/p4_node.c:1038-1250): the composition a tree-heterogeneous model expects at each tip, which
// Tree.compoTestUsingSimulations and the model-fit tests of p4 compare with the observed one
// (p4/tree.py:8192, 8544, 9071-9180).  The P decks are the ones resident on the device; the
// recursion itself is dim-sized per node and runs on the host in the reference's order of operations.
#include <cmath>
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

}  // namespace p4b

using namespace p4b;

extern "C" {

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
