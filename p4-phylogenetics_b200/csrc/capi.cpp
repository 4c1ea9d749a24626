// capi.cpp -- the extern "C" entry points declared in include/p4b200.h.
// Each function is the engine's counterpart of one pf.* wrapper of the
// reference (Pf/pfmodule.c); see the header for the line citations.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../include/p4b200.h"
#include "engine.h"

using namespace p4b;

#define CHECK_PTR(p, what, ret)                         \
    do {                                                \
        if (!(p)) {                                     \
            setError("%s: NULL handle", what);          \
            return ret;                                 \
        }                                               \
    } while (0)

extern "C" {

const char *p4b_version(void) { return "p4b200 0.1 (sm_100a)"; }
const char *p4b_lastError(void) { return lastError(); }
int p4b_deviceCount(void) { return deviceCount(); }
int p4b_setDevice(int device) { return setDevice(device); }
int p4b_setShard(int rank, int world) { return setShard(rank, world); }
int p4b_shardRangeFor(int nPatterns, int rank, int world, int *lo, int *hi)
{
    if (world < 1 || rank < 0 || rank >= world || nPatterns < 0 || !lo || !hi) { setError("p4b_shardRangeFor: bad arguments"); return 1; }
    *lo = (int)(((long long)nPatterns * rank) / world);
    *hi = (int)(((long long)nPatterns * (rank + 1)) / world);
    return 0;
}
int p4b_commGetUniqueId(char id128[128]) { return commGetUniqueId(id128); }
int p4b_commInitRank(const char id128[128], int rank, int world) { return commInitRank(id128, rank, world); }
int p4b_commDestroy(void) { engineMailShutdown(); return commDestroy(); }
int p4b_peerReduceState(void) { return peerReduceState(); }
long long p4b_kernelLaunchCount(void) { return kernelLaunchCount(); }
void p4b_setFusedTreeKernel(int on) { setFusedEnabled(on); }
int p4b_setFusedVariant(int v) { return setFusedVariant(v); }
const char *p4b_lastCLKernelName(void) { return lastCLKernelName(); }
void p4b_setFusedTreeKernel20(int on) { setFusedAAEnabled(on); }
void p4b_setDeferredNodeCalls(int on) { setDeferEnabled(on); }
void p4b_setSharedCondLikes(int on) { setShareEnabled(on); }
void p4b_setMemoize(int on) { setMemoizeEnabled(on); }
void p4b_setTensorCoreKernel(int on) { setDmmaEnabled(on); }
void p4b_setScalers(int on) { setScalersEnabled(on); }

// ---- data ------------------------------------------------------------------
p4b_data p4b_newData(int nTax, int nParts)
{
    if (nParts <= 0) { setError("newData: nParts=%d", nParts); return nullptr; }
    Data *d = new Data();
    d->nTax = nTax;
    d->nParts = nParts;
    d->parts.assign(nParts, nullptr);
    return d;
}
void p4b_freeData(p4b_data d) { delete (Data *)d; }   // parts are freed by their owners, like the reference
int p4b_pokePartInData(p4b_part p, p4b_data d, int i)
{
    CHECK_PTR(p, "pokePartInData", 1);
    CHECK_PTR(d, "pokePartInData", 1);
    Data *D = (Data *)d;
    if (i < 0 || i >= D->nParts) { setError("pokePartInData: index %d out of range", i); return 1; }
    D->parts[i] = (Part *)p;
    return 0;
}
p4b_part p4b_newPart(int nTax, int nChar, const char *equateSymbols, int nEquates, const char *symbols, int dim)
{
    return newPart(nTax, nChar, equateSymbols, nEquates, symbols, dim);
}
void p4b_freePart(p4b_part p) { freePart((Part *)p); }
int p4b_pokeEquatesTable(p4b_part p, const char *table) { CHECK_PTR(p, "pokeEquatesTable", 1); return pokeEquatesTable((Part *)p, table); }
int p4b_pokeSequences(p4b_part p, const char *s) { CHECK_PTR(p, "pokeSequences", 1); return pokeSequences((Part *)p, s); }
int p4b_makePatterns(p4b_part p) { CHECK_PTR(p, "makePatterns", 1); return makePatterns((Part *)p); }
int p4b_setGlobalInvarSitesVec(p4b_part p) { CHECK_PTR(p, "setGlobalInvarSitesVec", 1); return setGlobalInvarSitesVec((Part *)p); }
int p4b_partPatternCount(p4b_part p) { CHECK_PTR(p, "partPatternCount", -1); return ((Part *)p)->nPatterns; }
int p4b_getUnconstrainedLogLike(p4b_part p, double *out)
{
    CHECK_PTR(p, "getUnconstrainedLogLike", 1);
    CHECK_PTR(out, "getUnconstrainedLogLike", 1);
    Part *P = (Part *)p;
    if (!P->nPatterns) { setError("part.c: unconstrainedLogLike: no patterns.  Needs both sequences and patterns."); return 1; }
    for (int v : P->sequences)
        if (v < 0) { setError("part.c: unconstrainedLogLike: bad character.  Can't do this calculation if there are any gaps, unknowns, ambiguities, or equates"); return 1; }
    double dsum = 0.0;
    for (int i = 0; i < P->nPatterns; i++) dsum = dsum + (P->patternCounts[i] * log((double)(P->patternCounts[i])));
    dsum = dsum - (P->nChar * log((double)(P->nChar)));
    *out = dsum;
    return 0;
}
int p4b_getSiteLikes(p4b_part p, double *out, int nOut)
{
    CHECK_PTR(p, "getSiteLikes", -1);
    Part *P = (Part *)p;
    if (P->siteLikes.empty()) { setError("getSiteLikes: no site likelihoods have been computed for this part"); return -1; }
    const int n = nOut < P->nChar ? nOut : P->nChar;
    memcpy(out, P->siteLikes.data(), sizeof(double) * n);
    return P->nChar;
}
const int *p4b_partSequences(p4b_part p) { return p ? ((Part *)p)->sequences.data() : nullptr; }
const int *p4b_partPatterns(p4b_part p) { return p ? ((Part *)p)->patterns.data() : nullptr; }
const int *p4b_partPatternCounts(p4b_part p) { return p ? ((Part *)p)->patternCounts.data() : nullptr; }
const int *p4b_partSequencePositionPatternIndex(p4b_part p) { return p ? ((Part *)p)->sequencePositionPatternIndex.data() : nullptr; }
const int *p4b_partGlobalInvarSitesVec(p4b_part p)
{
    return (p && !((Part *)p)->globalInvarSitesVec.empty()) ? ((Part *)p)->globalInvarSitesVec.data() : nullptr;
}
const int *p4b_partGlobalInvarSitesArray(p4b_part p)
{
    return (p && !((Part *)p)->globalInvarSitesArray.empty()) ? ((Part *)p)->globalInvarSitesArray.data() : nullptr;
}
const int *p4b_partEquates(p4b_part p) { return (p && ((Part *)p)->nEquates) ? ((Part *)p)->equates.data() : nullptr; }
int p4b_partNChar(p4b_part p) { return p ? ((Part *)p)->nChar : -1; }
int p4b_partNTax(p4b_part p) { return p ? ((Part *)p)->nTax : -1; }
int p4b_partDim(p4b_part p) { return p ? ((Part *)p)->dim : -1; }

// ---- model -----------------------------------------------------------------
p4b_model p4b_newModel(int nParts, int doRelRates, int relRatesAreFree, int nFreePrams, int isHet, int *rMatrixNormalizeTo1,
                       double *PINVAR_MIN, double *PINVAR_MAX, double *KAPPA_MIN, double *KAPPA_MAX, double *GAMMA_SHAPE_MIN,
                       double *GAMMA_SHAPE_MAX, double *PIVEC_MIN, double *PIVEC_MAX, double *RATE_MIN, double *RATE_MAX,
                       double *RELRATE_MIN, double *RELRATE_MAX, double *BRLEN_MIN, double *BRLEN_MAX)
{
    if (nParts <= 0 || !PIVEC_MIN) { setError("p4_newModel: bad arguments"); return nullptr; }
    Model *m = new Model();
    m->nParts = nParts;
    m->doRelRates = doRelRates;
    m->relRatesAreFree = relRatesAreFree;
    m->nFreePrams = nFreePrams;
    m->isHet = isHet;
    m->rMatrixNormalizeTo1 = rMatrixNormalizeTo1;
    m->PINVAR_MIN = PINVAR_MIN; m->PINVAR_MAX = PINVAR_MAX;
    m->KAPPA_MIN = KAPPA_MIN; m->KAPPA_MAX = KAPPA_MAX;
    m->GAMMA_SHAPE_MIN = GAMMA_SHAPE_MIN; m->GAMMA_SHAPE_MAX = GAMMA_SHAPE_MAX;
    m->PIVEC_MIN = PIVEC_MIN; m->PIVEC_MAX = PIVEC_MAX;
    m->RATE_MIN = RATE_MIN; m->RATE_MAX = RATE_MAX;
    m->RELRATE_MIN = RELRATE_MIN; m->RELRATE_MAX = RELRATE_MAX;
    m->BRLEN_MIN = BRLEN_MIN; m->BRLEN_MAX = BRLEN_MAX;
    m->parts.assign(nParts, nullptr);
    return m;
}

void p4b_freeModel(p4b_model m)
{
    Model *M = (Model *)m;
    if (!M) return;
    for (ModelPart *mp : M->parts) {
        if (!mp) continue;
        for (Gdasrv *g : mp->gdasrvs) delete g;
        delete mp;
    }
    delete M;
}

static ModelPart *modelPart(p4b_model m, int pNum, const char *what)
{
    Model *M = (Model *)m;
    if (!M) { setError("%s: NULL model", what); return nullptr; }
    if (pNum < 0 || pNum >= M->nParts || !M->parts[pNum]) { setError("%s: model has no part %d", what, pNum); return nullptr; }
    return M->parts[pNum];
}

int p4b_newModelPart(p4b_model m, int pNum, int dim, int nComps, int nRMatrices, int nGdasrvs, int nCat, int pInvarFree, int *bQETneedsReset)
{
    Model *M = (Model *)m;
    CHECK_PTR(M, "p4_newModelPart", 1);
    if (pNum < 0 || pNum >= M->nParts) { setError("p4_newModelPart: part %d out of range", pNum); return 1; }
    if (dim <= 0 || dim > 64 || nComps <= 0 || nRMatrices <= 0 || nGdasrvs < 0 || nCat <= 0 || !bQETneedsReset) {
        setError("p4_newModelPart: bad arguments (dim=%d nComps=%d nRMatrices=%d nGdasrvs=%d nCat=%d)", dim, nComps, nRMatrices, nGdasrvs, nCat);
        return 1;
    }
    ModelPart *mp = new ModelPart();
    mp->dim = dim;
    mp->nComps = nComps;
    mp->nRMatrices = nRMatrices;
    mp->nGdasrvs = nGdasrvs;
    mp->nCat = nCat;
    mp->pInvarFree = pInvarFree;
    mp->comps.resize(nComps);
    mp->rMatrices.resize(nRMatrices);
    mp->gdasrvs.assign(nGdasrvs, nullptr);
    mp->compSet.assign(nComps, 0);
    mp->rMatrixSet.assign(nRMatrices, 0);
    mp->bqe.resize((size_t)nComps * nRMatrices);
    mp->bQETneedsReset = bQETneedsReset;
    delete M->parts[pNum];
    M->parts[pNum] = mp;
    return 0;
}

int p4b_newComp(p4b_model m, int pNum, int mNum, int isFree, double *val)
{
    ModelPart *mp = modelPart(m, pNum, "p4_newComp");
    if (!mp) return 1;
    if (mNum < 0 || mNum >= mp->nComps || !val) { setError("p4_newComp: bad comp %d", mNum); return 1; }
    mp->comps[mNum].isFree = isFree;
    mp->comps[mNum].val = val;
    mp->compSet[mNum] = 1;
    return 0;
}

int p4b_newRMatrix(p4b_model m, int pNum, int mNum, int isFree, int spec)
{
    ModelPart *mp = modelPart(m, pNum, "p4_newRMatrix");
    if (!mp) return 1;
    if (mNum < 0 || mNum >= mp->nRMatrices) { setError("p4_newRMatrix: bad rMatrix %d", mNum); return 1; }
    RMatrix &r = mp->rMatrices[mNum];
    r.isFree = isFree;
    r.spec = spec;
    r.hasKappa = false;
    r.bigR.assign((size_t)mp->dim * mp->dim, 1.0);   // 'ones' / 'specified': all ones until poked (Pf/p4_model.c:427-433)
    if (spec == 5) {                                  // RMATRIX_2P: kappa starts at 2 (Pf/p4_model.c:418-425)
        if (mp->dim != 4) { setError("p4_newRMatrix: the 2-parameter matrix needs dim 4"); return 1; }
        r.hasKappa = true;
        r.kappa = 2.0;
        r.bigR.assign(16, 0.0);   // the reference leaves it unset until p4_setKappa / p4_setPrams
    } else if (spec > 100) {                          // empirical protein tables (Pf/defines.h:10-29)
        if (mp->dim != 20) { setError("p4_newRMatrix: protein spec %d needs dim 20", spec); return 1; }
        if (proteinBigR(spec, r.bigR.data())) { setError("p4_newRMatrix: unknown rMatrix spec %d", spec); return 1; }
    }
    mp->rMatrixSet[mNum] = 1;
    return 0;
}

p4b_gdasrv p4b_newGdasrv(p4b_model m, int pNum, int mNum, int nCat, int isFree, double *val, double *freqs, double *rates)
{
    ModelPart *mp = modelPart(m, pNum, "p4_newGdasrv");
    if (!mp) return nullptr;
    if (mNum < 0 || mNum >= mp->nGdasrvs || !val || !freqs || !rates) { setError("p4_newGdasrv: bad arguments"); return nullptr; }
    Gdasrv *g = new Gdasrv();
    g->isFree = isFree;
    g->nCat = nCat;
    g->val = val;
    g->freqs = freqs;
    g->rates = rates;
    delete mp->gdasrvs[mNum];
    mp->gdasrvs[mNum] = g;
    return g;
}

int p4b_gdasrvCalcRates(p4b_gdasrv g)
{
    Gdasrv *G = (Gdasrv *)g;
    CHECK_PTR(G, "gdasrvCalcRates", 1);
    return discreteGamma(G->freqs, G->rates, G->val[0], G->val[0], G->nCat, 0);
}
int p4b_gdasrvCalcRates_np(int nCat, double alpha, double *freqs, double *rates)
{
    if (nCat < 2 || !freqs || !rates) { setError("gdasrvCalcRates_np: bad arguments"); return 1; }
    return discreteGamma(freqs, rates, alpha, alpha, nCat, 0);
}

int p4b_setRMatrixBigR(p4b_model m, int pNum, int rNum, int i, int j, double val)
{
    ModelPart *mp = modelPart(m, pNum, "p4_setRMatrixBigR");
    if (!mp) return 1;
    if (rNum < 0 || rNum >= mp->nRMatrices || !mp->rMatrixSet[rNum] || i < 0 || j < 0 || i >= mp->dim || j >= mp->dim) {
        setError("p4_setRMatrixBigR: bad index rMatrix %d [%d][%d]", rNum, i, j);
        return 1;
    }
    mp->rMatrices[rNum].bigR[i * mp->dim + j] = val;
    mp->rMatrices[rNum].bigR[j * mp->dim + i] = val;
    return 0;
}

int p4b_setKappa(p4b_model m, int pNum, int rNum, double val)
{
    ModelPart *mp = modelPart(m, pNum, "p4_setKappa");
    if (!mp) return 1;
    if (rNum < 0 || rNum >= mp->nRMatrices || !mp->rMatrices[rNum].hasKappa) { setError("p4_setKappa: rMatrix %d is not a 2-parameter matrix", rNum); return 1; }
    mp->rMatrices[rNum].kappa = val;
    setKappaBigR(mp->rMatrices[rNum]);
    return 0;
}

int p4b_setPInvarVal(p4b_model m, int pNum, double val)
{
    ModelPart *mp = modelPart(m, pNum, "p4_setPInvarVal");
    if (!mp) return 1;
    mp->pInvar = val;
    return 0;
}
int p4b_setRelRateVal(p4b_model m, int pNum, double val)
{
    ModelPart *mp = modelPart(m, pNum, "p4_setRelRateVal");
    if (!mp) return 1;
    mp->relRate = val;
    return 0;
}
int p4b_resetBQET(p4b_model m, int pNum, int compNum, int rMatrixNum) { return resetBQET((Model *)m, pNum, compNum, rMatrixNum); }
double p4b_getRelRate(p4b_model m, int pNum)
{
    ModelPart *mp = modelPart(m, pNum, "p4_getRelRate");
    return mp ? mp->relRate : NAN;
}
static Eig *getEigOf(p4b_model m, int pNum, int c, int r, const char *what)
{
    ModelPart *mp = modelPart(m, pNum, what);
    if (!mp) return nullptr;
    if (c < 0 || c >= mp->nComps || r < 0 || r >= mp->nRMatrices) { setError("%s: bad comp %d / rMatrix %d", what, c, r); return nullptr; }
    Eig *e = &mp->bqe[(size_t)c * mp->nRMatrices + r];
    if (!e->allocated) { setError("%s: comp %d rMatrix %d has no Q yet", what, c, r); return nullptr; }
    return e;
}
int p4b_getBigQ(p4b_model m, int pNum, int compNum, int rMatrixNum, double *out)
{
    Eig *e = getEigOf(m, pNum, compNum, rMatrixNum, "getBigQ");
    if (!e || !out) return 1;
    memcpy(out, e->Q.data(), e->Q.size() * sizeof(double));
    return 0;
}
int p4b_getBigR(int spec, double *out400)
{
    if (!out400 || proteinBigR(spec, out400)) { setError("getBigR: unknown protein rMatrix spec %d", spec); return 1; }
    return 0;
}
int p4b_getModelBigR(p4b_model m, int pNum, int rMatrixNum, double *out)
{
    ModelPart *mp = modelPart(m, pNum, "getModelBigR");
    if (!mp || !out) return 1;
    if (rMatrixNum < 0 || rMatrixNum >= mp->nRMatrices || !mp->rMatrixSet[rMatrixNum]) { setError("getModelBigR: bad rMatrix %d", rMatrixNum); return 1; }
    memcpy(out, mp->rMatrices[rMatrixNum].bigR.data(), sizeof(double) * mp->dim * mp->dim);
    return 0;
}
int p4b_getEig(p4b_model m, int pNum, int compNum, int rMatrixNum, double *eigvecs, double *inverseEigvecs, double *eigvals)
{
    Eig *e = getEigOf(m, pNum, compNum, rMatrixNum, "getEig");
    if (!e) return 1;
    if (eigvecs) memcpy(eigvecs, e->V.data(), e->V.size() * sizeof(double));
    if (inverseEigvecs) memcpy(inverseEigvecs, e->Vinv.data(), e->Vinv.size() * sizeof(double));
    if (eigvals) memcpy(eigvals, e->lam.data(), e->lam.size() * sizeof(double));
    return 0;
}

// ---- tree ------------------------------------------------------------------
p4b_tree p4b_newTree(int nNodes, int nLeaves, int *preOrder, int *postOrder, int *passLimit, double *partLikes, p4b_data d, p4b_model m)
{
    if (nNodes <= 0 || !preOrder || !postOrder || !partLikes || !d || !m) { setError("p4_newTree: bad arguments"); return nullptr; }
    Tree *t = new Tree();
    t->nNodes = nNodes;
    t->nLeaves = nLeaves;
    t->preOrder = preOrder;
    t->postOrder = postOrder;
    t->passLimit = passLimit;
    t->partLikes = partLikes;
    t->data = (Data *)d;
    t->model = (Model *)m;
    t->nParts = t->data->nParts;
    t->nodes.assign(nNodes, nullptr);
    for (int p = 0; p < t->nParts; p++)
        if (!t->model->parts[p] || !t->data->parts[p]) { setError("p4_newTree: part %d of the model or data is missing", p); delete t; return nullptr; }
    if (treeDeviceCreate(t)) {
        treeDeviceDestroy(t);
        delete t;
        return nullptr;
    }
    return t;
}

void p4b_freeTree(p4b_tree t)
{
    Tree *T = (Tree *)t;
    if (!T) return;
    treeDeviceDestroy(T);
    // nodes are freed one by one through p4_freeNode, before the tree (p4/tree.py:9202-9253); one freed
    // later must not reach back into this tree
    for (Node *n : T->nodes)
        if (n) n->tree = nullptr;
    delete T;
}

p4b_node p4b_newNode(int nodeNum, p4b_tree t, int seqNum, int isLeaf, int inTree)
{
    Tree *T = (Tree *)t;
    CHECK_PTR(T, "p4_newNode", nullptr);
    if (nodeNum < 0 || nodeNum >= T->nNodes) { setError("p4_newNode: nodeNum %d out of range", nodeNum); return nullptr; }
    Node *n = new Node();
    n->nodeNum = nodeNum;
    n->tree = T;
    n->seqNum = seqNum;
    n->isLeaf = isLeaf;
    n->inTree = inTree;
    n->compNums.assign(T->nParts, 0);
    n->rMatrixNums.assign(T->nParts, 0);
    n->gdasrvNums.assign(T->nParts, 0);
    n->clNeedsUpdating = (!isLeaf && inTree) ? 1 : 0;   // Pf/p4_node.c:114-122
    if (Node *old = T->nodes[nodeNum]) {   // a replaced node gives its CL slots back before the new one asks for its own
        if (treeHasPending(T)) treeFlushPending(T);
        nodeDeviceRelease(old);
        T->nodes[nodeNum] = nullptr;
        delete old;
    }
    if (nodeDeviceCreate(n)) { nodeDeviceRelease(n); delete n; return nullptr; }
    T->nodes[nodeNum] = n;
    T->topoStamp++;
    return n;
}

void p4b_freeNode(p4b_node n)
{
    Node *N = (Node *)n;
    if (!N) return;
    if (N->tree && treeHasPending(N->tree)) treeFlushPending(N->tree);   // queued CL calls may name this node
    nodeDeviceRelease(N);
    if (N->tree) N->tree->topoStamp++;
    if (N->tree && N->nodeNum >= 0 && N->nodeNum < (int)N->tree->nodes.size() && N->tree->nodes[N->nodeNum] == N) N->tree->nodes[N->nodeNum] = nullptr;
    delete N;
}

int p4b_setNodeRelation(p4b_node n, int relation, int relNum)
{
    Node *N = (Node *)n;
    CHECK_PTR(N, "p4_setNodeRelation", 1);
    // queued node-level CL calls were issued against the current topology: run them before it changes
    if (treeHasPending(N->tree) && treeFlushPending(N->tree)) return 1;
    Node *rel = nullptr;
    if (relNum >= 0) {
        if (relNum >= N->tree->nNodes || !N->tree->nodes[relNum]) { setError("p4_setNodeRelation: node %d does not exist", relNum); return 1; }
        rel = N->tree->nodes[relNum];
    }
    Node **slot = relation == 0 ? &N->parent : (relation == 1 ? &N->leftChild : (relation == 2 ? &N->sibling : nullptr));
    if (!slot) { setError("Error in p4_setNodeRelation: \"relation\" is out of range"); return 1; }
    if (*slot != rel) { *slot = rel; N->tree->topoStamp++; }     // p4 re-sends the whole topology with every Tree.setCStuff: only a CHANGE invalidates cached launch plans
    return 0;
}
int p4b_setTreeRoot(p4b_tree t, p4b_node n)
{
    CHECK_PTR(t, "p4_setTreeRoot", 1);
    Tree *T = (Tree *)t;
    if (T->root != (Node *)n) { T->root = (Node *)n; T->topoStamp++; }
    return 0;
}
int p4b_setBrLen(p4b_node n, double brLen)
{
    CHECK_PTR(n, "p4_setBrLen", 1);
    ((Node *)n)->brLen = brLen;
    return 0;
}
// Tree.setCStuff in one call: what 3*nNodes p4_setNodeRelation, p4_setTreeRoot and nNodes-1 p4_setBrLen
// calls say (p4/tree.py:9338-9355), as arrays indexed by node number.
int p4b_setTreeCStuff(p4b_tree t, int nNodes, const int *parent, const int *leftChild, const int *sibling, const double *brLen, int rootNum)
{
    Tree *T = (Tree *)t;
    CHECK_PTR(T, "p4b_setTreeCStuff", 1);
    if (nNodes != T->nNodes || !parent || !leftChild || !sibling || !brLen) { setError("p4b_setTreeCStuff: bad arguments"); return 1; }
    if (rootNum < 0 || rootNum >= nNodes || !T->nodes[rootNum]) { setError("p4b_setTreeCStuff: bad root %d", rootNum); return 1; }
    if (treeHasPending(T) && treeFlushPending(T)) return 1;
    auto at = [&](int i, Node **out) -> int {
        if (i < 0) { *out = nullptr; return 0; }
        if (i >= nNodes || !T->nodes[i]) { setError("p4b_setTreeCStuff: node %d does not exist", i); return 1; }
        *out = T->nodes[i];
        return 0;
    };
    for (int i = 0; i < nNodes; i++) {
        Node *n = T->nodes[i];
        if (!n) continue;
        Node *par = nullptr, *lc = nullptr, *sib = nullptr;
        if (at(parent[i], &par) || at(leftChild[i], &lc) || at(sibling[i], &sib)) return 1;
        if (par != n->parent || lc != n->leftChild || sib != n->sibling) { n->parent = par; n->leftChild = lc; n->sibling = sib; T->topoStamp++; }
        if (i != rootNum) n->brLen = brLen[i];
    }
    if (T->root != T->nodes[rootNum]) { T->root = T->nodes[rootNum]; T->topoStamp++; }
    return 0;
}
static int setNum(p4b_node n, int pNum, int val, int which)
{
    Node *N = (Node *)n;
    CHECK_PTR(N, "p4_set*Num", 1);
    if (pNum < 0 || pNum >= N->tree->nParts) { setError("p4_set*Num: bad part %d", pNum); return 1; }
    (which == 0 ? N->compNums : which == 1 ? N->rMatrixNums : N->gdasrvNums)[pNum] = val;
    return 0;
}
int p4b_setCompNum(p4b_node n, int pNum, int val) { return setNum(n, pNum, val, 0); }
int p4b_setRMatrixNum(p4b_node n, int pNum, int val) { return setNum(n, pNum, val, 1); }
int p4b_setGdasrvNum(p4b_node n, int pNum, int val) { return setNum(n, pNum, val, 2); }
double p4b_getTreeLen(p4b_tree t)
{
    Tree *T = (Tree *)t;
    CHECK_PTR(T, "p4_getTreeLen", NAN);
    double leng = 0.0;
    for (int i = 0; i < T->nNodes; i++) {
        const int j = T->postOrder[i];
        if (j == P4B_NO_ORDER) continue;
        Node *n = T->nodes[j];
        if (n && n != T->root) leng += n->brLen;
    }
    return leng;
}

// ---- hot path ----------------------------------------------------------------
int p4b_setPrams(p4b_tree t, int pNum) { CHECK_PTR(t, "p4_setPrams", 1); return treeSetPrams((Tree *)t, pNum); }
int p4b_calculateBigPDecks(p4b_node n) { CHECK_PTR(n, "p4_calculateBigPDecks", 1); return nodeCalculateBigPDecks((Node *)n); }
int p4b_calculateAllBigPDecksAllParts(p4b_tree t) { CHECK_PTR(t, "p4_calculateAllBigPDecksAllParts", 1); return treeCalculateAllBigPDecks((Tree *)t); }
int p4b_setConditionalLikelihoodsOfInternalNodePart(p4b_node n, int pNum)
{
    CHECK_PTR(n, "p4_setConditionalLikelihoodsOfInternalNodePart", 1);
    return nodeSetCL((Node *)n, pNum);
}
double p4b_partLogLike(p4b_tree t, p4b_part p, int pNum, int getSiteLikes)
{
    CHECK_PTR(t, "p4_partLogLike", NAN);
    return treePartLogLike((Tree *)t, (Part *)p, pNum, getSiteLikes);
}
double p4b_treeLogLike(p4b_tree t, int getSiteLikes)
{
    CHECK_PTR(t, "p4_treeLogLike", NAN);
    return treeLogLike((Tree *)t, getSiteLikes);
}

int p4b_partLogLikeBegin(p4b_tree t, int pNum)
{
    CHECK_PTR(t, "p4b_partLogLikeBegin", 1);
    return treePartLogLikeBegin((Tree *)t, pNum);
}
int p4b_treesPartLogLike(int nTrees, const p4b_tree *trees, int pNum, double *out)
{
    if (nTrees < 0 || (nTrees > 0 && (!trees || !out))) { setError("p4b_treesPartLogLike: bad arguments"); return 1; }
    return treesPartLogLike((Tree **)trees, nTrees, pNum, out);
}

// ---- state transfer -----------------------------------------------------------
int p4b_copyCondLikes(p4b_tree a, p4b_tree b, int doAll)
{
    CHECK_PTR(a, "p4_copyCondLikes", 1);
    CHECK_PTR(b, "p4_copyCondLikes", 1);
    return treeCopyCondLikes((Tree *)a, (Tree *)b, doAll);
}
int p4b_copyBigPDecks(p4b_tree a, p4b_tree b, int doAll)
{
    CHECK_PTR(a, "p4_copyBigPDecks", 1);
    CHECK_PTR(b, "p4_copyBigPDecks", 1);
    return treeCopyBigPDecks((Tree *)a, (Tree *)b, doAll);
}

int p4b_copyModelPrams(p4b_tree ta, p4b_tree tb)   // Pf/p4_treeCopyVerify.c:63-185
{
    Tree *A = (Tree *)ta, *B = (Tree *)tb;
    CHECK_PTR(A, "p4_copyModelPrams", 1);
    CHECK_PTR(B, "p4_copyModelPrams", 1);
    if (A->model->nParts != B->model->nParts) { setError("p4_copyModelPrams: the models differ in part count"); return 1; }
    for (int p = 0; p < A->model->nParts; p++) {
        ModelPart *a = A->model->parts[p], *b = B->model->parts[p];
        const int dim = a->dim;
        if (b->dim != dim || b->nComps != a->nComps || b->nRMatrices != a->nRMatrices || b->nGdasrvs != a->nGdasrvs || b->nCat != a->nCat) {
            setError("p4_copyModelPrams: part %d differs in shape between the two models", p);
            return 1;
        }
        for (int i = 0; i < a->nComps; i++)
            if (a->comps[i].isFree)
                for (int s = 0; s < dim; s++) b->comps[i].val[s] = a->comps[i].val[s];
        for (int i = 0; i < a->nRMatrices; i++)
            if (a->rMatrices[i].isFree) {
                // the reference copies the upper/lower triangles except the last pair (i < dim-2)
                for (int r = 0; r < dim - 2; r++)
                    for (int c = r + 1; c < dim; c++) {
                        b->rMatrices[i].bigR[r * dim + c] = a->rMatrices[i].bigR[r * dim + c];
                        b->rMatrices[i].bigR[c * dim + r] = a->rMatrices[i].bigR[c * dim + r];
                    }
                if (a->rMatrices[i].spec == 5) b->rMatrices[i].kappa = a->rMatrices[i].kappa;
            }
        for (size_t i = 0; i < a->bqe.size(); i++) {
            Eig &ea = a->bqe[i], &eb = b->bqe[i];
            if (ea.allocated && eb.allocated && (ea.content != eb.content || ea.content == 0)) {
                // same solve on both sides (content id) -> nothing to move, nothing to re-upload
                eb.Q = ea.Q;
                eb.V = ea.V;
                eb.Vinv = ea.Vinv;
                eb.lam = ea.lam;
                eb.inPi = ea.inPi;
                eb.inR = ea.inR;
                eb.content = ea.content;
                eb.version++;
            }
        }
        for (int i = 0; i < a->nGdasrvs; i++)
            if (a->gdasrvs[i] && b->gdasrvs[i] && a->gdasrvs[i]->isFree) {
                b->gdasrvs[i]->val[0] = a->gdasrvs[i]->val[0];
                for (int c = 0; c < a->nCat; c++) b->gdasrvs[i]->rates[c] = a->gdasrvs[i]->rates[c];
            }
        if (a->pInvarFree) b->pInvar = a->pInvar;
        if (A->model->doRelRates && A->model->relRatesAreFree) b->relRate = a->relRate;
    }
    return 0;
}

static int verifyHost(Tree *A, Tree *B)   // Pf/p4_treeCopyVerify.c:270-611
{
    const double eps = 1.e-15;
    int bad = 0;
    for (int j = 0; j < A->nNodes; j++) {
        const int i = A->preOrder[j];
        if (i == P4B_NO_ORDER) continue;
        if (A->nodes[i] && A->nodes[i]->clNeedsUpdating) { printf("Verify: aTree node %i clNeedsUpdating is set.  Bad.\n", i); bad = 1; }
        if (B->nodes[i] && B->nodes[i]->clNeedsUpdating) { printf("Verify: bTree node %i clNeedsUpdating is set.  Bad.\n", i); bad = 1; }
    }
    // model parameters
    int diff = 0;
    for (int p = 0; p < A->model->nParts && !diff; p++) {
        ModelPart *a = A->model->parts[p], *b = B->model->parts[p];
        const int dim = a->dim;
        for (int i = 0; i < a->nComps && !diff; i++)
            if (a->comps[i].isFree)
                for (int s = 0; s < dim; s++)
                    if (fabs(a->comps[i].val[s] - b->comps[i].val[s]) > eps) diff = 1;
        for (int i = 0; i < a->nRMatrices && !diff; i++)
            if (a->rMatrices[i].isFree) {
                for (int r = 0; r < dim - 2; r++)
                    for (int c = r + 1; c < dim; c++)
                        if (fabs(a->rMatrices[i].bigR[r * dim + c] - b->rMatrices[i].bigR[r * dim + c]) > eps ||
                            fabs(a->rMatrices[i].bigR[c * dim + r] - b->rMatrices[i].bigR[c * dim + r]) > eps)
                            diff = 1;
                if (a->rMatrices[i].spec == 5 && fabs(a->rMatrices[i].kappa - b->rMatrices[i].kappa) > eps) diff = 1;
            }
        for (size_t i = 0; i < a->bqe.size() && !diff; i++) {
            const Eig &ea = a->bqe[i], &eb = b->bqe[i];
            if (!(ea.allocated && eb.allocated)) continue;
            for (int k = 0; k < dim * dim; k++)
                if (fabs(ea.Q[k] - eb.Q[k]) > eps || fabs(ea.V[k] - eb.V[k]) > eps || fabs(ea.Vinv[k] - eb.Vinv[k]) > eps) diff = 1;
            for (int k = 0; k < dim; k++)
                if (fabs(ea.lam[k] - eb.lam[k]) > eps) diff = 1;
        }
        for (int i = 0; i < a->nGdasrvs && !diff; i++)
            if (a->gdasrvs[i] && a->gdasrvs[i]->isFree) {
                if (fabs(a->gdasrvs[i]->val[0] - b->gdasrvs[i]->val[0]) > eps) diff = 1;
                for (int c = 0; c < a->nCat; c++)
                    if (fabs(a->gdasrvs[i]->rates[c] - b->gdasrvs[i]->rates[c]) > eps) diff = 1;
            }
        if (a->pInvarFree && fabs(a->pInvar - b->pInvar) > eps) diff = 1;
        if (A->model->doRelRates && A->model->relRatesAreFree && fabs(b->relRate - a->relRate) > eps) diff = 1;
    }
    if (diff) { printf("Verify: model prams are different.  Bad.\n"); bad = 1; }
    // node relations
    diff = 0;
    auto num = [](Node *n) { return n ? n->nodeNum : -1; };
    for (int j = 0; j < A->nNodes; j++) {
        const int i = A->preOrder[j];
        if (i == P4B_NO_ORDER) continue;
        Node *a = A->nodes[i], *b = B->nodes[i];
        if (!a || !b) { diff = 1; continue; }
        if (num(a->parent) != num(b->parent) || num(a->leftChild) != num(b->leftChild) || num(a->sibling) != num(b->sibling)) diff = 1;
    }
    if (num(A->root) != num(B->root)) diff = 1;
    if (diff) { printf("Verify: nodes relations are different.  Bad.\n"); bad = 1; }
    // branch lengths, orders
    diff = 0;
    for (int j = 0; j < A->nNodes; j++) {
        const int i = A->preOrder[j];
        if (i == P4B_NO_ORDER) continue;
        if (A->nodes[i] && B->nodes[i] && A->nodes[i] != A->root && fabs(B->nodes[i]->brLen - A->nodes[i]->brLen) > eps) diff = 1;
    }
    for (int i = 0; i < A->nNodes; i++)
        if (B->postOrder[i] != A->postOrder[i] || B->preOrder[i] != A->preOrder[i]) diff = 1;
    if (diff) { printf("Verify: node brLen, or clNeedsUpdating different.  Bad.\n"); bad = 1; }
    // model usage
    diff = 0;
    for (int j = 0; j < A->nNodes; j++) {
        const int i = A->preOrder[j];
        if (i == P4B_NO_ORDER) continue;
        Node *a = A->nodes[i], *b = B->nodes[i];
        if (!a || !b || a == A->root) continue;
        for (int p = 0; p < A->nParts; p++) {
            if (a->compNums[p] != b->compNums[p] || a->rMatrixNums[p] != b->rMatrixNums[p]) diff = 1;
            else if (A->model->parts[p]->nCat > 1 && a->gdasrvNums[p] != b->gdasrvNums[p]) diff = 1;
        }
    }
    if (diff) { printf("Verify: model arrangements are different.  Bad.\n"); bad = 1; }
    return bad;
}

int p4b_verifyIdentityOfTwoTrees(p4b_tree a, p4b_tree b)
{
    Tree *A = (Tree *)a, *B = (Tree *)b;
    CHECK_PTR(A, "p4_verifyIdentityOfTwoTrees", -1);
    CHECK_PTR(B, "p4_verifyIdentityOfTwoTrees", -1);
    if (A->nNodes != B->nNodes || A->nParts != B->nParts || A->model->nParts != B->model->nParts) return 1;
    int bad = verifyHost(A, B);
    const int dev = treeVerifyDevice(A, B);
    if (dev < 0) return -1;
    return (bad || dev) ? 1 : 0;
}

// ---- inspection ---------------------------------------------------------------
int p4b_treeShardRange(p4b_tree t, int pNum, int *lo, int *hi) { CHECK_PTR(t, "p4b_treeShardRange", 1); return treeShardRangeOf((Tree *)t, pNum, lo, hi); }
int p4b_getNodeCL(p4b_node n, int pNum, double *out) { CHECK_PTR(n, "p4b_getNodeCL", 1); return nodeGetCL((Node *)n, pNum, out); }
int p4b_getNodeBigP(p4b_node n, int pNum, double *out) { CHECK_PTR(n, "p4b_getNodeBigP", 1); return nodeGetBigP((Node *)n, pNum, out); }
int p4b_setNodeBigP(p4b_node n, int pNum, const double *in) { CHECK_PTR(n, "p4b_setNodeBigP", 1); return nodeSetBigP((Node *)n, pNum, in); }
int p4b_setTreeStoresCL(p4b_tree t, int on)
{
    CHECK_PTR(t, "p4b_setTreeStoresCL", 1);
    ((Tree *)t)->storeCL = on ? 1 : 0;
    return 0;
}
int p4b_treeSync(p4b_tree t) { return treeSync((Tree *)t); }
int p4b_treeTimerBegin(p4b_tree t) { CHECK_PTR(t, "p4b_treeTimerBegin", 1); return treeTimerBegin((Tree *)t); }
double p4b_treeTimerEnd(p4b_tree t) { CHECK_PTR(t, "p4b_treeTimerEnd", -1.0); return treeTimerEnd((Tree *)t); }
int p4b_treeLastCLTiming(p4b_tree t, double *ms, int *nLaunches) { CHECK_PTR(t, "p4b_treeLastCLTiming", 1); return treeLastCLTiming((Tree *)t, ms, nLaunches); }
long long p4b_treeDeviceBytes(p4b_tree t) { return t ? treeDeviceBytes((Tree *)t) : 0; }
int p4b_flushL2(p4b_tree t) { return treeFlushL2((Tree *)t); }

}  // extern "C"
