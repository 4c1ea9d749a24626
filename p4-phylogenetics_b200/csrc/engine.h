// engine.h -- host-side object model of the B200 likelihood engine.
//
// The objects mirror what the reference keeps behind its `pf` boundary
// (Pf/pftypes.h) so that every pf.* call has somewhere to land, but they are
// organised for the device: per-node conditional likelihoods (CL) live in one
// HBM arena per (tree, part), transition matrices in one deck per tree, tip
// states as one byte per (taxon, pattern) shared by every tree that uses the
// data, and eigensystems are cached on the host and mirrored to the device
// only when they change.
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace p4b {

// ---- error channel -------------------------------------------------------
void setError(const char *fmt, ...);
const char *lastError();

// ---- data ----------------------------------------------------------------
// Device mirror of one part's pattern shard (owned by Part, shared by trees).
struct PartDevice {
    int lo = 0, hi = 0;        // global pattern range resident here
    int ps = 0;                // padded pattern stride (multiple of 32)
    uint8_t *tips = nullptr;   // [nTax][ps] tip code index (see tipIndex())
    int *counts = nullptr;     // [ps] pattern counts (0 in the padding)
    uint64_t *invarMask = nullptr;  // [ps] bit s set <=> globalInvarSitesArray[s][pat]
    uint64_t *equateMask = nullptr; // [nRealEquates] bit s set <=> state s allowed
    uint64_t dataVersion = 0;  // Part::version this mirror was built from
    int device = -1;
    int rank = 0, world = 1;   // the shard (p4b_setShard) lo/hi were cut for
};

struct Part {
    int dim = 0, nTax = 0, nChar = 0, nEquates = 0, nPatterns = 0;
    std::string symbols, equateSymbols;
    std::vector<int> sequences;   // [nTax][nChar]   (Pf/pftypes.h:41)
    std::vector<int> patterns;    // [nTax][nChar], columns < nPatterns valid (:36)
    std::vector<int> patternCounts;                // [nChar] (:37)
    std::vector<int> sequencePositionPatternIndex; // [nChar] (:38)
    std::vector<int> equates;     // [nEquates][dim] (:43)
    std::vector<int> globalInvarSitesVec;   // [nChar], empty until set (:45)
    std::vector<int> globalInvarSitesArray; // [dim][nChar] (:46)
    std::vector<double> siteLikes;          // [nChar], empty until asked for (:47)
    std::vector<int> taxList;               // [nTax] which sequences partComposition looks at (:49); empty until poked
    uint64_t version = 1;         // bumped whenever patterns / invar arrays change
    // Leaf lookup tables have one column per tip code index:
    //   0..dim-1 the states, dim = "matches everything" (gap, '?', N-like
    //   equates), dim+1+j = the j-th equate that is not N-like.
    std::vector<int> realEquateOfEquate; // [nEquates] -> j or -1 if N-like
    int nRealEquates = 0;
    // ... of which only those that OCCUR in the patterns get a column (recounted whenever the device mirror is
    // rebuilt): with the eleven IUPAC ambiguity codes defined but two of them present a DNA table is 7 wide, not 15
    std::vector<int> usedEquateOfEquate; // [nEquates] -> column j or -1 (N-like, or absent from the data)
    int nUsedEquates = -1;               // -1: not counted yet (then every real equate has a column)
    int equateColumn(int e) const { return nUsedEquates < 0 ? realEquateOfEquate[e] : usedEquateOfEquate[e]; }
    int nEquateColumns() const { return nUsedEquates < 0 ? nRealEquates : nUsedEquates; }
    int tableWidth() const { return dim + 1 + nEquateColumns(); }
    PartDevice dev;
};

struct Data {
    int nTax = 0, nParts = 0;
    std::vector<Part *> parts;
};

// data.cpp
Part *newPart(int nTax, int nChar, const char *equateSymbols, int nEquates, const char *symbols, int dim);
void freePart(Part *p);
int pokeEquatesTable(Part *p, const char *table);
int pokeSequences(Part *p, const char *s);
int makePatterns(Part *p);
int setGlobalInvarSitesVec(Part *p);

// ---- model ---------------------------------------------------------------
struct Eig {                      // cf. p4_bigQAndEigStruct + eigStruct, Pf/pftypes.h:56-72, 219-224
    bool allocated = false;       // reference: aQE->bigQ != NULL
    std::vector<double> Q, V, Vinv, lam;   // dim*dim, dim*dim, dim*dim, dim
    std::vector<double> inPi, inR;         // the (pi, R) this eigensystem was solved for; empty = none
    uint64_t version = 0;         // bumped on every recompute
    uint64_t content = 0;         // id of the solve that produced it (equal ids = identical numbers)
};
struct Comp { int isFree = 0; double *val = nullptr; };
struct RMatrix {
    int isFree = 0, spec = 0;
    std::vector<double> bigR;     // dim*dim
    bool hasKappa = false;
    double kappa = 2.0;
};
struct Gdasrv { int isFree = 0, nCat = 0; double *val = nullptr, *freqs = nullptr, *rates = nullptr; };

struct ModelPart {
    int dim = 0, nComps = 0, nRMatrices = 0, nGdasrvs = 0, nCat = 1, pInvarFree = 0;
    std::vector<Comp> comps;
    std::vector<RMatrix> rMatrices;
    std::vector<Gdasrv *> gdasrvs;
    std::vector<char> compSet, rMatrixSet;
    double pInvar = -1.0, relRate = -1.0;   // Pf/p4_model.c:150, 156
    std::vector<Eig> bqe;                   // [nComps*nRMatrices]
    int *bQETneedsReset = nullptr;          // borrowed numpy int32 [nComps*nRMatrices]
};

struct Model {
    int nParts = 0, doRelRates = 0, relRatesAreFree = 0, nFreePrams = 0, isHet = 0;
    int *rMatrixNormalizeTo1 = nullptr;
    double *PINVAR_MIN = nullptr, *PINVAR_MAX = nullptr, *KAPPA_MIN = nullptr, *KAPPA_MAX = nullptr,
           *GAMMA_SHAPE_MIN = nullptr, *GAMMA_SHAPE_MAX = nullptr, *PIVEC_MIN = nullptr, *PIVEC_MAX = nullptr,
           *RATE_MIN = nullptr, *RATE_MAX = nullptr, *RELRATE_MIN = nullptr, *RELRATE_MAX = nullptr,
           *BRLEN_MIN = nullptr, *BRLEN_MAX = nullptr;
    std::vector<ModelPart *> parts;
};

// model.cpp
int discreteGamma(double *freqK, double *rK, double alfa, double beta, int K, int median);
int resetBQET(Model *m, int pNum, int cNum, int rNum);   // Q + eigensystem of one (comp,rMatrix)
int proteinBigR(int spec, double *out400);
void setKappaBigR(RMatrix &r);

// ---- tree ----------------------------------------------------------------
struct Tree;
struct Node {
    int nodeNum = 0;
    Tree *tree = nullptr;
    Node *parent = nullptr, *leftChild = nullptr, *sibling = nullptr;
    int seqNum = -1, isLeaf = 0, inTree = 1;
    double brLen = -1.0;                      // Pf/p4_node.c:35
    std::vector<int> compNums, rMatrixNums, gdasrvNums;
    int clNeedsUpdating = 0;
    std::vector<int> clSlot;                  // per part: CL arena slot, -1 if none
    std::vector<char> clSel;                  // per part: 0 the slot is in the tree's own arena, 1 in its twin's
    std::vector<uint64_t> clStamp;            // per part: id of the computation that filled the CL
    std::vector<uint64_t> pStamp;             // per part: id of the computation that filled the P deck
    std::vector<char> clResident;             // per part: the CL in the arena is the current one (see Tree::storeCL)
    // what the node's current P deck / CL was computed FROM (p4b_setMemoize): a node-level call whose inputs are
    // exactly these again would reproduce the same numbers and is skipped
    std::vector<std::vector<double>> pKey;    // per part: eigensystem content id, effective branch length per category
    std::vector<std::vector<uint64_t>> clKey; // per part: children in order: node number, CL stamp (or sequence), P stamp
};

struct TreeDevice;   // tree.cu
struct Tree {
    int nNodes = 0, nLeaves = 0, nParts = 0;
    std::vector<Node *> nodes;
    Node *root = nullptr;
    Data *data = nullptr;
    Model *model = nullptr;
    int *preOrder = nullptr, *postOrder = nullptr, *passLimit = nullptr;  // borrowed numpy int32
    double *partLikes = nullptr;                                           // borrowed numpy float64
    double logLike = 0.0;
    // 1 (default, the reference's behaviour): p4_treeLogLike leaves every node's CL in memory.
    // 0: a whole-tree evaluation writes only the CLs it must re-read itself; the others are
    // recomputed (one storing pass) the first time anything asks for them.
    int storeCL = 1;
    uint64_t topoStamp = 1;       // bumped whenever a node relation, the root or a node itself changes (plans of whole-tree launches are cached per topology)
    TreeDevice *dev = nullptr;
};

// tree.cu -- device side
int deviceCount();
int setDevice(int device);
int setShard(int rank, int world);
void shardRange(int nPatterns, int *lo, int *hi);
long long kernelLaunchCount();

int treeDeviceCreate(Tree *t);
void treeDeviceDestroy(Tree *t);
int nodeDeviceCreate(Node *n);
void nodeDeviceRelease(Node *n);
void partDeviceFree(Part *p);

int treeSetPrams(Tree *t, int pNum);
int nodeCalculateBigPDecks(Node *n);
int treeCalculateAllBigPDecks(Tree *t);
int nodeSetCL(Node *n, int pNum);
double treePartLogLike(Tree *t, Part *p, int pNum, int getSiteLikes);
double treeLogLike(Tree *t, int getSiteLikes);
int treeCopyCondLikes(Tree *a, Tree *b, int doAll);
int treeCopyBigPDecks(Tree *a, Tree *b, int doAll);
int treeVerifyDevice(Tree *a, Tree *b);   // 0 same, 1 different, <0 error
int nodeGetCL(Node *n, int pNum, double *out);
int nodeGetBigP(Node *n, int pNum, double *out);
int nodeSetBigP(Node *n, int pNum, const double *in);
int treeSync(Tree *t);
int treeTimerBegin(Tree *t);
double treeTimerEnd(Tree *t);
int treeLastCLTiming(Tree *t, double *ms, int *nLaunches);
long long treeDeviceBytes(Tree *t);
int treeFlushL2(Tree *t);
int treeShardRangeOf(Tree *t, int p, int *lo, int *hi);
int engineInitPublic();
void engineMailShutdown();
int peerReduceState();
void setFusedEnabled(int on);
int setFusedVariant(int v);
const char *lastCLKernelName();
void setDmmaEnabled(int on);
void setScalersEnabled(int on);
int treeEnsureResident(Tree *t, int p);
int treeFlushPending(Tree *t);
bool treeHasPending(Tree *t);
void setDeferEnabled(int on);
void setShareEnabled(int on);
void setFusedAAEnabled(int on);
void setMemoizeEnabled(int on);
int treesPartLogLike(Tree **trees, int n, int p, double *out);
int treePartLogLikeBegin(Tree *t, int p);
// simulate: device part (tree.cu).  `fill(dst, n)` must write the next n uniforms of the caller's stream.
int treeSimulateDevice(Tree *t, int p, const uint8_t *cats, const uint8_t *rootStates, const uint8_t *invar, const int *rank,
                       int nVar, const std::function<void(double *, size_t)> &fill);
const double *treeRootCLHost(Tree *t, int p, int *psOut);
int treeNewtSetup(Tree *t);
double treeNewtAround(Tree *t, double epsilon, double likeDelta);
int nodeNewtDerivs(Node *n, double out[3]);
int nodeGetCL2(Node *n, int pNum, double *out);
long long treeNewtIterations(Tree *t);

// comm.cpp -- NCCL, loaded at run time
int commGetUniqueId(char id128[128]);
int commInitRank(const char id128[128], int rank, int world);
int commDestroy();
bool commActive();
int commWorld();
int commOpenPeerMailboxes(void *mine, int rank, void **peers, void *cudaStream);
void commClosePeerMailboxes(int rank, void **peers);
int commAllReduceSum(double *devBuf, int count, void *cudaStream);

}  // namespace p4b
