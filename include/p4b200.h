/* p4b200.h -- C ABI of the B200 Felsenstein-pruning likelihood engine.
 *
 * This is the drop-in boundary for ONE path of pgfoster/p4-phylogenetics: the
 * likelihood engine behind the CPython extension `p4.pf` (Pf/pfmodule.c).  Every
 * entry point below replaces one `pf.*` wrapper of the reference and keeps its
 * argument order and meaning; the citation after each prototype is the
 * reference wrapper it stands in for.  Handles are opaque pointers that the
 * Python side stores as integers, exactly like the reference's
 * Py_BuildValue("l", ptr) convention (Pf/pfmodule.c:1405, 1455, 1530).
 *
 * Plain C types only: no CUDA, PyTorch or C++ types cross this boundary.
 *
 * Borrowed buffers.  Like the reference (Pf/p4_model.c:339, 504-506;
 * Pf/p4_tree.c:46-49) the engine KEEPS the raw pointers of these caller-owned
 * arrays and re-reads them on every compute call, because p4's Python code
 * mutates them in place without calling a setter:
 *   comp val[dim]; gdasrv val[1], freqs[nCat], rates[nCat];
 *   bQETneedsReset[nComps*nRMatrices] (int32); preOrder/postOrder[nNodes]
 *   (int32); partLikes[nParts]; the model limit arrays.
 * The caller must keep them alive until the owning object is freed.
 *
 * Errors.  The reference prints a message and calls exit(1) on internal errors
 * (e.g. Pf/p4_tree.c:436, 495).  A library must not end the host process, so
 * functions that can fail return int (0 = ok, nonzero = fatal) or a NULL
 * handle, and the message is available from p4b_lastError().  The Python `pf`
 * mirror turns that into an exception whose default outcome is process exit
 * status 1.  The numerical sentinel is unchanged: a non-positive site
 * likelihood makes the part log-likelihood -1.0e99 (Pf/p4_tree.c:1182).
 *
 * There is NO CPU fallback for the compute entry points: without a usable
 * CUDA device they fail with an error.
 */
#ifndef P4B200_H
#define P4B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef void *p4b_data;
typedef void *p4b_part;
typedef void *p4b_model;
typedef void *p4b_gdasrv;
typedef void *p4b_tree;
typedef void *p4b_node;

#define P4B_NO_ORDER (-10000)        /* Pf/defines.h:66, p4/var.py:109 */
#define P4B_GAP_CODE (-1)            /* Pf/defines.h:33 */
#define P4B_QMARK_CODE (-2)          /* Pf/defines.h:34 */
#define P4B_EQUATES_BASE (-64)       /* Pf/defines.h:36 */
#define P4B_BAD_LIKE (-1.0e99)       /* Pf/p4_tree.c:1182 */

/* ---- engine ------------------------------------------------------------- */
const char *p4b_version(void);
const char *p4b_lastError(void);
/* Number of usable CUDA devices (0 on a CPU-only host; never an error). */
int p4b_deviceCount(void);
/* Bind this process to one CUDA device.  Must precede the first p4b_newTree. */
int p4b_setDevice(int device);
/* Pattern sharding for one-process-per-GPU runs: this
 * process keeps patterns [lo,hi) of every part on its device, with
 * lo = floor(nPatterns*rank/world).  Default rank 0 of 1.  Call before
 * p4b_newTree. */
int p4b_setShard(int rank, int world);
/* The rule p4b_setShard applies: patterns [lo,hi) of nPatterns belong to `rank` of `world`. */
int p4b_shardRangeFor(int nPatterns, int rank, int world, int *lo, int *hi);
/* NCCL communicator over the shard ranks.  p4b_commGetUniqueId fills a
 * 128-byte id on rank 0; the host program ships it to the other ranks by any
 * means and every rank calls p4b_commInitRank.  With a communicator present
 * p4b_partLogLike / p4b_treeLogLike return the all-reduced (global) value. */
int p4b_commGetUniqueId(char id128[128]);
int p4b_commInitRank(const char id128[128], int rank, int world);
int p4b_commDestroy(void);
/* How the shard sums of a log-likelihood are combined when a communicator is present: 1 = inside the kernel that folds them,
 * through peer mailboxes over NVLink (every rank's mailbox mapped on every other rank with CUDA IPC; the default where peer
 * access exists), -1 = NCCL all-reduce after the kernel (fallback, or P4B_PEER_REDUCE=0), 0 = not decided yet (no sharded
 * evaluation has run). */
int p4b_peerReduceState(void);
/* p4b_treeLogLike evaluates 4-state parts with ONE whole-tree kernel launch by
 * default; 0 selects the one-launch-per-node kernels instead (same results;
 * kept for comparison and profiling). */
void p4b_setFusedTreeKernel(int on);
/* Launch shape of the 4-state whole-tree kernel.  -1 (default): chosen from the shard size -- 128 threads x 3
 * CTAs/SM for three waves or more, 64 x 5 or 32 x 7 for smaller shards (what a 4- or 8-GPU run of BASELINE
 * configs[1] uses).  0, 1, 2 force one of those three; 3..8 are comparison shapes.  Results are bit-identical
 * across shapes; the setter exists so that the parity tests can run every shape at every shard size. */
int p4b_setFusedVariant(int v);
/* Name and launch shape of the CL kernel launched last (bench.py reports it beside its roofline). */
const char *p4b_lastCLKernelName(void);
/* 20-state parts with 4 rate categories have a whole-tree kernel of their own (FP64 tensor cores, the
 * running CL stays in the accumulator registers from one node to the next); 0 keeps the one-launch-
 * per-node kernels for them. */
void p4b_setFusedTreeKernel20(int on);
/* Node-level calls (p4b_calculateBigPDecks, p4b_setConditionalLikelihoodsOfInternalNodePart) only
 * QUEUE work by default (1): the reference's callers issue them in dependency order along the dirty
 * path of a proposal (p4/chain.py:668-688), and the engine runs a tree's whole queue as one P(t)
 * launch plus one step-list CL launch when the result is needed -- at p4b_partLogLike (which then
 * fuses the root reduction into the same launch), or before anything else reads or changes the state
 * the queue depends on (copy, verify, inspection, a topology or parameter change).  Results are
 * those of immediate execution.  0 launches every call at once (for comparison). */
void p4b_setDeferredNodeCalls(int on);
/* p4_copyCondLikes between the cur and prop tree of a chain moves no data by default (1): the CL
 * arenas of the two trees pair up, a copied node simply references the source node's buffer, and a
 * node whose buffer is shared is given a fresh slot when it is next computed (copy on write; the pair
 * holds 2 x nInternal slots, exactly as two private arenas do).  The observable state is that of a
 * real copy.  A tree can pair with one other tree; copies to or from any further tree, and all copies
 * with 0 here, are device-to-device memcpys. */
void p4b_setSharedCondLikes(int on);
/* Node-level calls whose inputs are exactly those of the result already in memory are skipped by default
 * (1).  p4_calculateBigPDecks / p4_setPrams: a node's P(t) is a pure function of its eigensystem and its
 * effective branch length per category; p4_setConditionalLikelihoodsOfInternalNodePart: a node's CL is a pure
 * function of its children (in order), their CLs and their P decks.  Every computation carries an id, so
 * "same inputs" is an exact comparison of ids, not of numbers.  The reference's callers recompute a whole part
 * after changing one composition or one node's model assignment (p4/chain.py:305-380, 560-608); with this the
 * engine recomputes only what the change reaches: the dirty path.  p4b_treeLogLike always recomputes everything,
 * as the reference does.  0 makes every call do its full work. */
void p4b_setMemoize(int on);
/* 20-state parts use the FP64 tensor-core (mma.sync m8n8k4) CL kernel by default;
 * 0 selects the FMA kernel instead (same results to rounding; for comparison). */
void p4b_setTensorCoreKernel(int on);
/* Per-pattern log-scalers against underflow (default 0 = off, the reference's behaviour: a site
 * likelihood that underflows makes the part -1.0e99, Pf/p4_tree.c:1182).  With 1, trees created
 * AFTERWARDS rescale a pattern's CL by 2^256 whenever its largest entry falls below 2^-256 and
 * carry the exponent to the root (+4 bytes per pattern and node).  Likelihoods of patterns that
 * never rescale are bit-identical either way; trees with hundreds of taxa get a finite lnL instead
 * of the sentinel.  p4b_getNodeCL then returns the rescaled values. */
void p4b_setScalers(int on);
/* Count of engine kernel launches since process start (bench.py gpu_launches). */
long long p4b_kernelLaunchCount(void);

/* ---- data parts --------------------------------------------- Pf/part.c -- */
p4b_data p4b_newData(int nTax, int nParts);                              /* pf.newData, Pf/pfmodule.c:29 */
void p4b_freeData(p4b_data d);                                           /* pf.freeData :43 */
int p4b_pokePartInData(p4b_part p, p4b_data d, int i);                   /* pf.pokePartInData :73 */
p4b_part p4b_newPart(int nTax, int nChar, const char *equateSymbols, int nEquates,
                     const char *symbols, int dim);                      /* pf.newPart :112 */
void p4b_freePart(p4b_part p);                                           /* pf.freePart :134 */
int p4b_pokeEquatesTable(p4b_part p, const char *table);                 /* pf.pokeEquatesTable :151, Pf/part.c:279 */
int p4b_pokeSequences(p4b_part p, const char *allSequences);             /* pf.pokeSequences :169, Pf/part.c:127 */
int p4b_makePatterns(p4b_part p);                                        /* pf.makePatterns :186, Pf/part.c:317 */
int p4b_setGlobalInvarSitesVec(p4b_part p);                              /* pf.setGlobalInvarSitesVec :310, Pf/part.c:716 */
int p4b_partPatternCount(p4b_part p);                                    /* pf.partPatternCount :247 */
/* pf.getUnconstrainedLogLike(part) :432 -> unconstrainedLogLike Pf/part.c:682-714: sum over patterns of
 * count*log(count) - nChar*log(nChar) (the multinomial ceiling Tree.modelFitTests / Data compare lnL with,
 * p4/data.py:223-258, p4/tree.py:8182).  Needs patterns; any gap, '?' or ambiguity is an error.
 * Returns non-zero and leaves *out untouched on error. */
int p4b_getUnconstrainedLogLike(p4b_part p, double *out);
/* The part's own statistics and views (host; csrc/partstats.cpp), each equal to the reference's:
 * pf.singleSequenceBaseCounts :219 -> Pf/part.c:556; pf.symbolSequences :233 -> :604 (out: nTax*nChar + 1 chars);
 * pf.partSequenceSitesCount :363 -> :1068; pf.pokePartTaxListAtIndex :326; pf.partComposition :350 -> :850-1066;
 * pf.partMeanNCharsPerSite :261 -> :1419; pf.partSimpleConstantSitesCount :275 -> :1459; pf.partBigXSquared :296 -> :1490. */
int p4b_singleSequenceBaseCounts(p4b_part p, int seqNum, int *outDim);
int p4b_symbolSequences(p4b_part p, char *outNTaxTimesNCharPlus1);
int p4b_partSequenceSitesCount(p4b_part p, int seqNum);
int p4b_pokePartTaxListAtIndex(p4b_part p, int val, int index);
int p4b_partComposition(p4b_part p, double *outDim);
double p4b_partMeanNCharsPerSite(p4b_part p);
int p4b_partSimpleConstantSitesCount(p4b_part p);
double p4b_partBigXSquared(p4b_part p);
/* pf.getSiteLikes :378 -- copies part->siteLikes (nChar doubles) filled by the
 * last p4b_partLogLike(..., getSiteLikes=1).  Returns nChar, or -1 if none. */
int p4b_getSiteLikes(p4b_part p, double *out, int nOut);
/* Read-only views of the host arrays of struct partStruct (Pf/pftypes.h:31-53),
 * for bit-exact parity checks.  Row-major, same shapes as the reference:
 * sequences/patterns [nTax][nChar] (patterns valid for columns < nPatterns),
 * patternCounts / sequencePositionPatternIndex / globalInvarSitesVec [nChar],
 * globalInvarSitesArray [dim][nChar], equates [nEquates][dim]. */
const int *p4b_partSequences(p4b_part p);
const int *p4b_partPatterns(p4b_part p);
const int *p4b_partPatternCounts(p4b_part p);
const int *p4b_partSequencePositionPatternIndex(p4b_part p);
const int *p4b_partGlobalInvarSitesVec(p4b_part p);
const int *p4b_partGlobalInvarSitesArray(p4b_part p);
const int *p4b_partEquates(p4b_part p);
int p4b_partNChar(p4b_part p);
int p4b_partNTax(p4b_part p);
int p4b_partDim(p4b_part p);

/* ---- model ------------------------------------------------ Pf/p4_model.c -- */
p4b_model p4b_newModel(int nParts, int doRelRates, int relRatesAreFree, int nFreePrams, int isHet,
                       int *rMatrixNormalizeTo1,
                       double *PINVAR_MIN, double *PINVAR_MAX, double *KAPPA_MIN, double *KAPPA_MAX,
                       double *GAMMA_SHAPE_MIN, double *GAMMA_SHAPE_MAX, double *PIVEC_MIN, double *PIVEC_MAX,
                       double *RATE_MIN, double *RATE_MAX, double *RELRATE_MIN, double *RELRATE_MAX,
                       double *BRLEN_MIN, double *BRLEN_MAX);            /* pf.p4_newModel :1484 */
void p4b_freeModel(p4b_model m);                                         /* pf.p4_freeModel :1554 */
int p4b_newModelPart(p4b_model m, int pNum, int dim, int nComps, int nRMatrices, int nGdasrvs,
                     int nCat, int pInvarFree, int *bQETneedsReset);     /* pf.p4_newModelPart :1590 */
int p4b_newComp(p4b_model m, int pNum, int mNum, int isFree, double *val);          /* pf.p4_newComp :1634 */
int p4b_newRMatrix(p4b_model m, int pNum, int mNum, int isFree, int spec);          /* pf.p4_newRMatrix :1653 */
p4b_gdasrv p4b_newGdasrv(p4b_model m, int pNum, int mNum, int nCat, int isFree,
                         double *val, double *freqs, double *rates);                /* pf.p4_newGdasrv :1672 */
int p4b_gdasrvCalcRates(p4b_gdasrv g);                                   /* pf.gdasrvCalcRates :1813 */
int p4b_gdasrvCalcRates_np(int nCat, double alpha, double *freqs, double *rates);   /* pf.gdasrvCalcRates_np :1836 */
int p4b_setRMatrixBigR(p4b_model m, int pNum, int rNum, int i, int j, double val);  /* pf.p4_setRMatrixBigR :1720 */
int p4b_setKappa(p4b_model m, int pNum, int rNum, double val);           /* pf.p4_setKappa :1741 */
int p4b_setPInvarVal(p4b_model m, int pNum, double val);                 /* pf.p4_setPInvarVal :1871 */
int p4b_setRelRateVal(p4b_model m, int pNum, double val);                /* pf.p4_setRelRateVal :1887 */
int p4b_resetBQET(p4b_model m, int pNum, int compNum, int rMatrixNum);   /* pf.p4_resetBQET :1613, Pf/p4_model.c:513 */
double p4b_getRelRate(p4b_model m, int pNum);                            /* pf.p4_getRelRate :2316 */
/* pf.getBigQ :1238 -- copy the cached, normalised Q of (comp,rMatrix): dim*dim doubles, row-major [from][to]. */
int p4b_getBigQ(p4b_model m, int pNum, int compNum, int rMatrixNum, double *out);
/* pf.getBigR :1271 -- the 20x20 empirical protein exchangeability table for an
 * RMATRIX_* spec code (Pf/defines.h:9-29), row-major. */
int p4b_getBigR(int spec, double *out400);
/* The current bigR of one rMatrix of the model: dim*dim doubles (inspection). */
int p4b_getModelBigR(p4b_model m, int pNum, int rMatrixNum, double *out);

/* ---- tree and nodes -------------------------------- Pf/p4_tree.c, p4_node.c -- */
p4b_tree p4b_newTree(int nNodes, int nLeaves, int *preOrder, int *postOrder,
                     int *newtAndBrentPowellOptPassLimit, double *partLikes,
                     p4b_data d, p4b_model m);                           /* pf.p4_newTree :1383 */
void p4b_freeTree(p4b_tree t);                                           /* pf.p4_freeTree :1410 */
p4b_node p4b_newNode(int nodeNum, p4b_tree t, int seqNum, int isLeaf, int inTree);  /* pf.p4_newNode :1443 */
void p4b_freeNode(p4b_node n);                                           /* pf.p4_freeNode :1460 */
/* relation: 0 parent, 1 leftChild, 2 sibling; relNum = nodeNum of the relative or -1 for none. */
int p4b_setNodeRelation(p4b_node n, int relation, int relNum);           /* pf.p4_setNodeRelation :1907 */
/* Addition: everything Tree.setCStuff (p4/tree.py:9338-9355) sends per node -- parent, leftChild,
 * sibling (node numbers, -1 = none), branch length -- and the root, as arrays indexed by node number:
 * one call instead of 4*nNodes.  Optional; the per-node calls above remain. */
int p4b_setTreeCStuff(p4b_tree t, int nNodes, const int *parent, const int *leftChild, const int *sibling,
                      const double *brLen, int rootNum);
int p4b_setTreeRoot(p4b_tree t, p4b_node n);                             /* pf.p4_setTreeRoot :1944 */
int p4b_setBrLen(p4b_node n, double brLen);                              /* pf.p4_setBrLen :1960 */
int p4b_setCompNum(p4b_node n, int pNum, int val);                       /* pf.p4_setCompNum :2013 */
int p4b_setRMatrixNum(p4b_node n, int pNum, int val);                    /* pf.p4_setRMatrixNum :2029 */
int p4b_setGdasrvNum(p4b_node n, int pNum, int val);                     /* pf.p4_setGdasrvNum :2045 */
double p4b_getTreeLen(p4b_tree t);                                       /* pf.p4_getTreeLen :1978 */

/* ---- the hot path ------------------------------------------------------- */
/* pf.p4_setPrams(tree, pNum|-1) :2086 -> p4_setPramsPart, Pf/p4_tree.c:218-541.
 * Host: gamma rates of free gdasrvs, 2-parameter R, comp checks, Q + eigensystem
 * for every (comp,rMatrix) used by a non-root node.  Device: the batched P(t)
 * kernel for every non-root node of the part(s).  Asynchronous. */
int p4b_setPrams(p4b_tree t, int pNum);
/* pf.p4_calculateBigPDecks(node) :2165 -> Pf/p4_node.c:286-346 (all parts). */
int p4b_calculateBigPDecks(p4b_node n);
/* pf.p4_calculateAllBigPDecksAllParts(tree) :2180. */
int p4b_calculateAllBigPDecksAllParts(p4b_tree t);
/* pf.p4_setConditionalLikelihoodsOfInternalNodePart(node, pNum) :2195
 * -> Pf/p4_node.c:636-857.  Enqueues the CL kernel for one node; asynchronous. */
int p4b_setConditionalLikelihoodsOfInternalNodePart(p4b_node n, int pNum);
/* pf.p4_partLogLike(tree, part, pNum, getSiteLikes) :2147 -> Pf/p4_tree.c:924-1197.
 * Synchronises; writes partLikes[pNum]; returns the part lnL (or -1.0e99). */
double p4b_partLogLike(p4b_tree t, p4b_part p, int pNum, int getSiteLikes);
/* pf.p4_treeLogLike(tree, getSiteLikes) :2132 -> Pf/p4_tree.c:868-922.
 * Recomputes the CL of EVERY internal node in postOrder, then all parts. */
double p4b_treeLogLike(p4b_tree t, int getSiteLikes);

/* ---- the optimisers' seam --------------- Pf/p4_treeOpt.c -- */
/* The reference's optimisers pack the free model parameters (and optionally all branch lengths)
 * into one vector and evaluate  p4_unWindParameters -> p4_setPrams -> p4_treeLogLike  per trial
 * (Pf/p4_treeOpt.c:17-120, 186-565, 579-615).  Same packing order, bounds and unpacking here. */
int p4b_countParameters(p4b_tree t, int doBrLens);
int p4b_windUpParameters(p4b_tree t, int doBrLens, double *x, double *lowerBounds, double *upperBounds);
int p4b_unWindParameters(p4b_tree t, int doBrLens, const double *x);
double p4b_logLikeForParameters(p4b_tree t, int doBrLens, const double *x);   /* p4_logLikeForNLOpt :579 */
/* Branch lengths one at a time through the dirty path (the role of Newton-Raphson in the reference's
 * p4_newtAndBrentPowellOpt / p4_newtAndBOBYQAOpt, Pf/p4_treeOpt.c:755-945): each branch is maximised by
 * Brent's bounded method, every evaluation being one P(t) launch + one step-list launch from the branch's
 * parent to the root.  maxPasses passes over all branches or until a pass gains less than tol.  Returns
 * the final log-likelihood (NaN on error); *nEvals receives the number of likelihood evaluations. */
double p4b_optimizeBrLens(p4b_tree t, int maxPasses, double tol, long *nEvals);
/* The reference's four optimiser entry points, native (csrc/opt.cpp, csrc/praxis.cpp): the reference's schedules of
 * calls with every objective evaluation (p4_setPrams + p4_treeLogLike) on the GPU.  Each returns the number of
 * objective evaluations, -1 on error. */
long p4b_allBrentPowellOptimize(p4b_tree t);               /* pf.p4_allBrentPowellOptimize, Pf/pfmodule.c:2247 -> Pf/p4_treeOpt.c:996-1180: Brent's praxis over model parameters and branch lengths */
long p4b_allBOBYQAOptimize(p4b_tree t, int doBrLens);      /* pf.p4_allBOBYQAOptimize :2212 -> Pf/p4_treeOpt.c:617-753: the box of p4_windUpParameters; a bounded Powell method stands where the reference calls nlopt */
long p4b_newtAndBrentPowellOpt(p4b_tree t);                /* pf.p4_newtAndBrentPowellOpt :2262 -> Pf/p4_treeOpt.c:1182-1330 (needs no p4_newtSetup call: it is made) */
long p4b_newtAndBOBYQAOpt(p4b_tree t);                     /* pf.p4_newtAndBOBYQAOpt :2230 -> Pf/p4_treeOpt.c:755-945 */
/* The two minimisers on a caller's objective fn(x, ctx) (host code; used by the entry points above and by the CPU tests):
 * Brent's principal-axis method as Pf/brent.c runs it (tol, h as there), and Powell's method confined to the box [lo, hi]. */
double p4b_praxisMinimize(int n, double *x, double tol, double h, double (*fn)(const double *, void *), void *ctx);
double p4b_boundedMinimize(int n, double *x, const double *lo, const double *hi, double xtol, double ftol, long maxEvals,
                           double (*fn)(const double *, void *), void *ctx, long *nEvals);
/* ---- Newton-Raphson on the branch lengths --- Pf/p4_treeNewt.c -- */
/* pf.p4_newtSetup(tree) Pf/pfmodule.c:2298 -> p4_newtSetup Pf/p4_treeNewt.c:11-75: allocates cl2 (per node,
 * the conditional likelihoods of everything on the far side of the node's branch) and the work space of the
 * first- and second-derivative P decks.  Idempotent.  Fails on a tree whose root is a leaf (the reference's
 * p4_setCL2Up exits there, Pf/p4_node.c:905-908) and on trees created with scalers. */
int p4b_newtSetup(p4b_tree t);
/* p4_newtAround(tree, epsilon, likeDelta) Pf/p4_treeNewt.c:78-205, called by p4_newtAndBrentPowellOpt /
 * p4_newtAndBOBYQAOpt (Pf/p4_treeOpt.c:755-945, 1182-1330): rounds over all branches in postOrder; per branch
 * p4_newtNode (:210-600) iterates v <- v - lnL'/lnL'' (guards: lnL'' >= 0 -> v/5, BRLEN_MIN, 5 x the old length,
 * BRLEN_MAX, 20 iterations, |lnL'| < epsilon) on derivatives computed from cl2 and the node's CL in one pass;
 * a round ends with p4_treeLogLike and the rounds stop when it moves by less than likeDelta (20 at most).
 * P decks must be current (p4_setPrams) as in the reference.  Returns the final log-likelihood (NaN on error). */
double p4b_newtAround(p4b_tree t, double epsilon, double likeDelta);
/* Inspection: {lnL, d lnL/dv, d2 lnL/dv2} in the length v of the node's branch at its current value
 * (the quantities p4_newtNode forms, Pf/p4_treeNewt.c:508-517), cl2 recomputed from the root's child down;
 * the node's cl2 [nCat*dim][nPatterns of this shard]; the number of derivative evaluations so far. */
int p4b_newtDerivs(p4b_node n, double out3[3]);
int p4b_getNodeCL2(p4b_node n, int pNum, double *out);
long long p4b_newtIterations(p4b_tree t);
int p4b_treePassLimit(p4b_tree t);   /* var.newtAndBrentPowellOptPassLimit as given to p4_newTree */
int p4b_treeNNodes(p4b_tree t);
int p4b_treeNLeaves(p4b_tree t);
int p4b_treeNParts(p4b_tree t);
int p4b_treePartDim(p4b_tree t, int pNum);
int p4b_getBrLens(p4b_tree t, double *outNNodes);                            /* pf.p4_getBrLens :2279; root slot = -1 */

/* pf.p4_partLogLike for nTrees trees in one go -- the prop trees of Metropolis-coupled chains after
 * each has had its proposal's node-level calls issued (p4/mcmc.py:2830-2835 runs the chains of a
 * generation one after the other; they are independent until the swap).  When all trees share the
 * data part and every tree's queued calls end at its root, all dirty paths and root reductions run as
 * ONE kernel launch (grid.y = tree), one fold, one all-reduce, one device->host copy.  Otherwise it
 * is a loop over p4b_partLogLike.  out[i] and each tree's partLikes[pNum] receive the values. */
int p4b_treesPartLogLike(int nTrees, const p4b_tree *trees, int pNum, double *out);
/* Start the evaluation p4b_partLogLike(t, pNum) would do -- queued P(t) jobs, queued CL calls, root
 * reduction -- and return without waiting.  The value is collected by the next p4b_partLogLike(t, pNum) or
 * p4b_treesPartLogLike including t (one copy and one synchronisation for all trees).  The host can prepare
 * the next chain's proposal while the GPU evaluates this one. */
int p4b_partLogLikeBegin(p4b_tree t, int pNum);

/* ---- consumers of the P decks beside the likelihood --- Pf/p4_treeSim.c -- */
/* pf.p4_expectedComposition(tree) Pf/pfmodule.c:2404 -> Pf/p4_treeSim.c:951-1045, one part per call here:
 * the root's composition pushed down every branch's P decks (p4_calculateExpectedComp, Pf/p4_node.c:1038-1250,
 * pInvar share included), averaged over the rate categories; out[seqNum][state] for every leaf.
 * pf.p4_expectedCompositionCounts(tree, partNum) :2384 -> :859-949: the same times the number of
 * non-gap, non-'?' sites of the sequence.  P decks must be current (p4_setPrams). */
/* pf.gsl_rng_get() :674, pf.gsl_rng_free :691, pf.gsl_rng_set(g, seed) :708, pf.gsl_rng_uniform(g) :739: the random
 * stream p4 hands to its simulations (var.gsl_rng).  GSL's default generator mt19937 with GSL's seeding
 * (seed 0 = 4357; uniform = 32-bit output / 2^32), so that a simulation is the same function of the seed. */
void *p4b_rngNew(void);
void p4b_rngFree(void *rng);
void p4b_rngSet(void *rng, unsigned long seed);
unsigned long p4b_rngGet(void *rng);
double p4b_rngUniform(void *rng);
void p4b_rngFillUniform(void *rng, double *out, long n);   /* the next n uniforms of the stream, in order */
/* The random draws and special functions Chain's proposals take from GSL through pf between likelihood evaluations
 * (p4/chain.py:2274-2565).  GSL is a third-party dependency of the reference; these are the algorithms GSL documents
 * (gamma: Marsaglia & Tsang 2000; Dirichlet: normalised gammas) on the mt19937 stream above, densities through libm. */
long p4b_rngSize(void *rng);                                                     /* pf.gsl_rng_size, Pf/pfmodule.c:724 */
void p4b_rngGetState(void *rng, void *buf);                                      /* pf.gsl_rng_getstate :752 (Mcmc checkpoints) */
void p4b_rngSetState(void *rng, const void *buf);                                /* pf.gsl_rng_setstate :779 */
double p4b_ranGamma(void *rng, double a, double b);                              /* pf.gsl_ran_gamma :827 */
void p4b_ranDirichlet(void *rng, int k, const double *alpha, double *theta);     /* pf.gsl_ran_dirichlet :940 */
double p4b_ranDirichletLnPdf(int k, const double *alpha, const double *theta);   /* pf.gsl_ran_dirichlet_lnpdf :989 */
double p4b_sfLnGamma(double x);                                                  /* pf.gsl_sf_lngamma :887 */
double p4b_ranGammaPdf(double x, double a, double b);                            /* pf.gsl_ran_gamma_pdf :848 */
void p4b_meanVariance(const double *seq, int n, double *mean, double *variance); /* pf.gsl_meanVariance :1025 */
/* pf.p4_simulate(tree, refTree|0, gsl_rng) :2333 -> p4_simulate Pf/p4_treeSim.c:14-420:
 * new sequences for every leaf, drawn down the tree from the root's composition through every branch's P decks
 * (rate category and invariant-or-not per site, pInvar), consuming the stream in the reference's order -- the
 * same seed gives the reference's sequences.  Writes part->sequences and globalInvarSitesVec, sets nPatterns
 * to 0: the caller re-compresses with pf.makePatterns / pf.setGlobalInvarSitesVec (p4/tree.py:9617-9626); the
 * tree lays its device state out again at its next use. */
int p4b_simulate(p4b_tree t, p4b_tree refTree, void *rng);
/* With a refTree (same tree and model on its own data, likelihood calculated): root state, rate category and
 * invariant-or-not of every site are drawn from the posterior at refTree's root instead (Pf/p4_treeSim.c:200-232).
 * pf.p4_drawAncState(tree, partNum, seqPos, draw) :2353 -> p4_drawAncStateP Pf/p4_treeSim.c:591-857: that draw for
 * one site, draw = {chStNum, catNum, isInvar, invarChNum}; like the reference it draws from the C library's random(),
 * which pf.reseedCRandomizer(seed) :472 seeds (srandom).  The root's conditional likelihoods must be current. */
int p4b_drawAncState(p4b_tree t, int pNum, int seqPos, int *draw4);
/* pf.bootstrapData(referenceData, toFillData, gsl_rng) :90 -> bootstrapData Pf/data.c:107-139: resample the sites of every part
 * with replacement on the caller's stream (gsl_rng_uniform_int), then pf.makePatterns on the filled parts. */
int p4b_bootstrapData(p4b_data reference, p4b_data toFill, void *rng);
void p4b_reseedCRandomizer(int seed);
/* Test hook: p4b_drawAncState's draw for a root CL given by the caller ([cat][state][pattern], nPatterns columns). */
int p4b_drawAncStateFromCL(p4b_part part, int seqPos, int nCat, double pInvar, int pInvarFree, const double *pi, const double *rootCL, int *draw4);
int p4b_expectedComposition(p4b_tree t, int pNum, double *outNTaxTimesDim);
int p4b_expectedCompositionCounts(p4b_tree t, int pNum, double *outNTaxTimesDim);

/* ---- cur/prop state transfer ------------------------ Pf/p4_treeCopyVerify.c -- */
int p4b_copyCondLikes(p4b_tree a, p4b_tree b, int doAll);                /* pf.p4_copyCondLikes :2445, Pf/p4_treeCopyVerify.c:7 */
int p4b_copyBigPDecks(p4b_tree a, p4b_tree b, int doAll);                /* pf.p4_copyBigPDecks :2462, :33 */
int p4b_copyModelPrams(p4b_tree a, p4b_tree b);                          /* pf.p4_copyModelPrams :2479, :63 */
/* 0 = same, 1 = different (epsilon 1e-15, Pf/p4_node.c:1406). */
int p4b_verifyIdentityOfTwoTrees(p4b_tree a, p4b_tree b);                /* pf.p4_verifyIdentityOfTwoTrees :2431, :187 */

/* ---- inspection (parity tests, profiling) ------------------------------- */
/* Number of patterns of part pNum resident on this process's device and the
 * global index of the first one (the shard [lo,hi) of p4b_setShard). */
int p4b_treeShardRange(p4b_tree t, int pNum, int *lo, int *hi);
/* Download one node's CL of part pNum in the reference layout
 * cl[cat][state][pattern] (Pf/pftypes.h:133), shard patterns only:
 * nCat*dim*(hi-lo) doubles.  Synchronises. */
int p4b_getNodeCL(p4b_node n, int pNum, double *out);
/* Download one node's transition matrices bigPDecks[cat][from][to]:
 * nCat*dim*dim doubles.  Synchronises. */
int p4b_getNodeBigP(p4b_node n, int pNum, double *out);
/* Test hook: overwrite one node's transition matrices with nCat*dim*dim given
 * doubles (the leaf lookup table is rebuilt from them).  Lets a parity test hold
 * P identical to the reference's and compare the CL kernels alone. */
int p4b_setNodeBigP(p4b_node n, int pNum, const double *in);
/* Eigen-decomposition cached for (comp,rMatrix): eigvecs, inverseEigvecs
 * (dim*dim each, row-major) and eigvals (dim).  Any may be NULL. */
int p4b_getEig(p4b_model m, int pNum, int compNum, int rMatrixNum,
               double *eigvecs, double *inverseEigvecs, double *eigvals);
/* lnL-only evaluations.  By default (1) p4b_treeLogLike leaves every node's CL in memory, as the
 * reference does.  With 0, a whole-tree evaluation of a 4-state part writes only the CLs it must
 * re-read itself (about a third of them); the rest are recomputed by one storing pass the first time
 * anything needs them (a node-level update, p4b_partLogLike, p4b_getNodeCL, copy, verify).  Results
 * are identical; it suits callers that evaluate the whole tree repeatedly (optimisers). */
int p4b_setTreeStoresCL(p4b_tree t, int on);
/* Wait for every kernel enqueued on the tree's stream. */
int p4b_treeSync(p4b_tree t);
/* Device-side timing on the tree's stream (CUDA events): call Begin, enqueue
 * work through the entry points above, call End; returns milliseconds. */
int p4b_treeTimerBegin(p4b_tree t);
double p4b_treeTimerEnd(p4b_tree t);
/* Timing of the last p4b_treeLogLike on the tree: total ms of the CL kernels
 * and their count (CUDA events around the CL section on the tree's stream). */
int p4b_treeLastCLTiming(p4b_tree t, double *ms, int *nLaunches);
/* Bytes of device memory held by the tree (CL arena + P decks). */
long long p4b_treeDeviceBytes(p4b_tree t);
/* Write a buffer larger than L2 (bench.py: L2 flush between timed steps). */
int p4b_flushL2(p4b_tree t);

#ifdef __cplusplus
}
#endif
#endif /* P4B200_H */
