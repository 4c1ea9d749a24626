// tree.cu -- device state of trees and the launch logic of the hot path.
//
// Reference call sites this file answers (see include/p4b200.h):
//   p4_setPramsPart            Pf/p4_tree.c:218-541
//   p4_calculateBigPDecksPart  Pf/p4_node.c:296-346
//   p4_setConditionalLikelihoodsOfInternalNodePart  Pf/p4_node.c:636-857
//   p4_treeLogLike / p4_partLogLike                 Pf/p4_tree.c:868-1027
//   p4_copyCondLikes / p4_copyBigPDecks / verify    Pf/p4_treeCopyVerify.c
//
// One CUDA stream serves the whole engine: the reference is single-threaded
// and its callers (Tree.calcLogLike, Chain.proposeSp) issue calls in a valid
// dependency order, so stream order is exactly the order the reference would
// have executed them in.  Compute calls only enqueue; p4b_partLogLike and
// p4b_treeLogLike are the synchronisation points.
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <functional>
#include <cstddef>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <memory>
#include <unordered_map>
#include <unordered_set>

#include "../../include/p4b200.h"
#include "engine.h"
#include "kernels.cuh"
#include "tree_dna.cuh"
#include "tree_dmma.cuh"
#include "tree_aa.cuh"
#include "newt.cuh"

namespace p4b {

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            setError("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__,   \
                     cudaGetErrorString(e_));                                                  \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

// ---------------------------------------------------------------------------
// Engine-wide state
// ---------------------------------------------------------------------------
struct Engine {
    bool ready = false;
    int device = 0;
    int rank = 0, world = 1;
    cudaStream_t stream = nullptr;
    // staging ring: small host->device parameter blocks (P jobs, eigensystems)
    char *hStage = nullptr, *dStage = nullptr;
    size_t stageCap = 0, stageHead = 0;
    bool stageDirtySinceSync = false;
    long long launches = 0;
    uint64_t stamp = 0;
    double *flushBuf = nullptr;
    size_t flushN = 0;
    int numSMs = 148;
    // P(t) jobs are collected engine-wide and launched together right before anything reads a P deck
    // (a CL launch, a copy, an inspection): the jobs of every chain of a generation share one launch.
    std::vector<PJob> pJobs;
    std::vector<double> pT;
    // results of a batched evaluation of several trees: [2*kMaxBatchTrees] device + pinned host
    double *dBatch = nullptr, *hBatch = nullptr;
    unsigned *batchTickets = nullptr;          // [kMaxBatchTrees] last-CTA tickets of a batched launch
    // step lists of the second-generation whole-tree kernel: one device buffer; uploads go through rotating pinned
    // slabs; a list identical to the one already on the device (a repeated full-tree evaluation) is not sent again
    Step2 *stepDev = nullptr;
    size_t stepDevCap = 0;
    Step2 *stepSlab[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t stepSlabCap[4] = {0, 0, 0, 0};
    cudaEvent_t stepEv[4] = {nullptr, nullptr, nullptr, nullptr};
    int stepSlabNext = 0;
    std::vector<Step2> stepShadow;             // what stepDev holds (once the stream reaches the last upload)
    // peer mailboxes of the in-kernel all-reduce (kernels.cuh mail_allreduce); mailState 0 untried, 1 on, -1 unavailable
    int mailState = 0;
    double *mailMine = nullptr;
    void *mailPeers[kMailMaxWorld] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    unsigned long long mailSeq = 0;
};
static Engine G;
static char g_lastKernel[128] = "";   // p4b_lastCLKernelName: the CL kernel launched last (bench.py names it in its roofline)
const char *lastCLKernelName() { return g_lastKernel; }
static int flushPJobs();
void nodeDeviceRelease(Node *n);
static bool g_useScalers = false;
static bool g_memoize = true;    // p4b_setMemoize
void setScalersEnabled(int on) { g_useScalers = on != 0; }

int deviceCount()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int setDevice(int device)
{
    if (G.ready && device != G.device) { setError("p4b_setDevice: engine already bound to device %d", G.device); return 1; }
    G.device = device;
    return 0;
}

int setShard(int rank, int world)
{
    if (world < 1 || rank < 0 || rank >= world) { setError("p4b_setShard: bad rank %d of %d", rank, world); return 1; }
    G.rank = rank;
    G.world = world;
    return 0;
}

void shardRange(int nPatterns, int *lo, int *hi)
{
    *lo = (int)(((long long)nPatterns * G.rank) / G.world);
    *hi = (int)(((long long)nPatterns * (G.rank + 1)) / G.world);
}

long long kernelLaunchCount() { return G.launches; }

// The exchange step of a sharded evaluation.  With a communicator present the ranks' partial sums are combined inside
// the kernel that folds them (peer mailboxes over NVLink); NCCL's all-reduce remains the fallback when the mailboxes
// cannot be mapped (no peer access, P4B_PEER_REDUCE=0).  Returns the kernel argument for the NEXT reducing launch.
static bool mailOn()
{
    if (!commActive()) return false;
    if (G.mailState == 0) {
        G.mailState = -1;
        const char *e = getenv("P4B_PEER_REDUCE");
        const int world = commWorld();
        // every rank goes through the hand-shake (it contains collectives) unless the switch is off everywhere
        if (!(e && atoi(e) == 0) && world <= kMailMaxWorld) {
            const size_t bytes = (size_t)kMailSlots * kMailTrees * kMailMaxWorld * 4 * sizeof(double);
            bool ok = cudaMalloc(&G.mailMine, bytes) == cudaSuccess && cudaMemset(G.mailMine, 0, bytes) == cudaSuccess;
            if (!ok) { cudaGetLastError(); if (G.mailMine) { cudaFree(G.mailMine); G.mailMine = nullptr; } }
            void *dummy = nullptr;
            if (!ok) cudaMalloc(&dummy, 256);      // the hand-shake needs SOME allocation to name; its outcome will be "failed"
            if (commOpenPeerMailboxes(ok ? (void *)G.mailMine : dummy, G.rank, G.mailPeers, (void *)G.stream) == 0 && ok) G.mailState = 1;
            if (dummy) cudaFree(dummy);
        }
    }
    return G.mailState == 1;
}
int peerReduceState() { return G.mailState; }
void engineMailShutdown()     // p4b_commDestroy: the mailboxes go with the communicator
{
    if (G.mailState == 1) {
        if (G.stream) cudaStreamSynchronize(G.stream);
        commClosePeerMailboxes(G.rank, G.mailPeers);
    }
    if (G.mailMine) { cudaFree(G.mailMine); G.mailMine = nullptr; }
    G.mailState = 0;
    G.mailSeq = 0;
}
static MailArgs nextMail()
{
    MailArgs m;
    memset(&m, 0, sizeof(m));
    m.world = 1;
    if (mailOn()) {
        m.mine = G.mailMine;
        m.world = commWorld();
        m.rank = G.rank;
        for (int r = 0; r < m.world; r++) m.peer[r] = (double *)G.mailPeers[r];
        m.seq = ++G.mailSeq;
    }
    return m;
}
// what callers do after a reducing launch: NCCL only when the kernel did not exchange the sums itself
static int allReduceIfNeeded(double *devBuf, int count)
{
    if (!commActive() || mailOn()) return 0;
    return commAllReduceSum(devBuf, count, (void *)G.stream);
}

static int engineInit()
{
    if (G.ready) return 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        setError("no usable CUDA device (%s): the likelihood engine has no CPU fallback",
                 e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return 1;
    }
    if (G.device >= n) { setError("device %d requested but only %d present", G.device, n); return 1; }
    CUDA_TRY(cudaSetDevice(G.device));
    CUDA_TRY(cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking));
    G.stageCap = 32u << 20;
    CUDA_TRY(cudaMallocHost(&G.hStage, G.stageCap));
    CUDA_TRY(cudaMalloc(&G.dStage, G.stageCap));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, G.device));
    G.numSMs = prop.multiProcessorCount;
    CUDA_TRY(cudaMalloc(&G.dBatch, 2 * sizeof(double) * kMaxBatchTrees));
    CUDA_TRY(cudaMallocHost(&G.hBatch, 2 * sizeof(double) * kMaxBatchTrees));
    CUDA_TRY(cudaMalloc(&G.batchTickets, sizeof(unsigned) * kMaxBatchTrees));
    CUDA_TRY(cudaMemset(G.batchTickets, 0, sizeof(unsigned) * kMaxBatchTrees));
    G.ready = true;
    return 0;
}

// Copy a small parameter block to the device through the pinned ring; returns
// the device address it will occupy once the stream reaches the copy.
static int stage(const void *src, size_t bytes, void **devPtr)
{
    const size_t need = (bytes + 255) & ~(size_t)255;
    if (need > G.stageCap) { setError("staging block of %zu bytes exceeds the ring", bytes); return 1; }
    if (G.stageHead + need > G.stageCap) {
        // wrap: everything staged before must have been consumed
        if (G.stageDirtySinceSync) CUDA_TRY(cudaStreamSynchronize(G.stream));
        G.stageDirtySinceSync = false;
        G.stageHead = 0;
    }
    memcpy(G.hStage + G.stageHead, src, bytes);
    CUDA_TRY(cudaMemcpyAsync(G.dStage + G.stageHead, G.hStage + G.stageHead, bytes, cudaMemcpyHostToDevice, G.stream));
    *devPtr = G.dStage + G.stageHead;
    G.stageHead += need;
    G.stageDirtySinceSync = true;
    return 0;
}

static int streamSync()
{
    CUDA_TRY(cudaStreamSynchronize(G.stream));
    G.stageDirtySinceSync = false;
    G.stageHead = 0;
    return 0;
}

// ---------------------------------------------------------------------------
// Part mirror: tip code indices, counts, constant-site masks
// ---------------------------------------------------------------------------
void partDeviceFree(Part *p)
{
    PartDevice &d = p->dev;
    if (d.tips) cudaFree(d.tips);
    if (d.counts) cudaFree(d.counts);
    if (d.invarMask) cudaFree(d.invarMask);
    if (d.equateMask) cudaFree(d.equateMask);
    d = PartDevice();
}

static int partDeviceEnsure(Part *p)
{
    PartDevice &d = p->dev;
    if (d.tips && d.dataVersion == p->version && d.device == G.device && d.rank == G.rank && d.world == G.world) return 0;
    if (d.tips && (d.rank != G.rank || d.world != G.world)) p->version++;   // the shard moved: trees laid out on the old range are stale too
    partDeviceFree(p);
    if (p->nPatterns <= 0) { setError("part has no patterns (pf.makePatterns not called?)"); return 1; }
    shardRange(p->nPatterns, &d.lo, &d.hi);
    {   // which ambiguity codes occur at all (over ALL patterns, so that every rank lays its tables out alike)
        std::vector<char> used(p->nEquates > 0 ? p->nEquates : 1, 0);
        for (int t = 0; t < p->nTax; t++) {
            const int *row = p->patterns.data() + (size_t)t * p->nChar;
            for (int k = 0; k < p->nPatterns; k++) {
                const int c = row[k];
                if (c < 0 && c != P4B_GAP_CODE && c != P4B_QMARK_CODE && c != -3) {
                    const int e = c - P4B_EQUATES_BASE;
                    if (e >= 0 && e < p->nEquates) used[e] = 1;
                }
            }
        }
        p->usedEquateOfEquate.assign(p->nEquates, -1);
        p->nUsedEquates = 0;
        for (int e = 0; e < p->nEquates; e++)
            if (p->realEquateOfEquate[e] >= 0 && used[e]) p->usedEquateOfEquate[e] = p->nUsedEquates++;
    }
    const int n = d.hi - d.lo;
    d.ps = ((n > 0 ? n : 1) + 255) & ~255;      // a multiple of every whole-tree kernel's CTA tile: no thread is ever out of range
    const size_t ps = (size_t)d.ps;
    const int dim = p->dim;
    std::vector<uint8_t> tips((size_t)p->nTax * ps, (uint8_t)dim);   // padding behaves like a gap
    for (int t = 0; t < p->nTax; t++) {
        const int *row = p->patterns.data() + (size_t)t * p->nChar + d.lo;
        uint8_t *out = tips.data() + (size_t)t * ps;
        for (int k = 0; k < n; k++) {
            const int c = row[k];
            int w;
            if (c >= 0) w = c;
            else if (c == P4B_GAP_CODE || c == P4B_QMARK_CODE || c == -3) w = dim;
            else {
                const int e = c - P4B_EQUATES_BASE;
                if (e < 0 || e >= p->nEquates) { setError("bad character code %d in patterns", c); return 1; }
                const int j = p->equateColumn(e);
                w = j < 0 ? dim : dim + 1 + j;
            }
            out[k] = (uint8_t)w;
        }
    }
    CUDA_TRY(cudaMalloc(&d.tips, tips.size() + 1024));      // + slack: the whole-tree kernel copies a CTA's 64..256 tip codes of a row at once, the last CTA past ps
    CUDA_TRY(cudaMemcpy(d.tips, tips.data(), tips.size(), cudaMemcpyHostToDevice));
    std::vector<int> counts(ps, 0);
    for (int k = 0; k < n; k++) counts[k] = p->patternCounts[d.lo + k];
    CUDA_TRY(cudaMalloc(&d.counts, ps * sizeof(int)));
    CUDA_TRY(cudaMemcpy(d.counts, counts.data(), ps * sizeof(int), cudaMemcpyHostToDevice));
    if (!p->globalInvarSitesArray.empty()) {
        std::vector<uint64_t> im(ps, 0);
        for (int s = 0; s < dim; s++) {
            const int *row = p->globalInvarSitesArray.data() + (size_t)s * p->nChar + d.lo;
            for (int k = 0; k < n; k++)
                if (row[k]) im[k] |= 1ull << s;
        }
        CUDA_TRY(cudaMalloc(&d.invarMask, ps * sizeof(uint64_t)));
        CUDA_TRY(cudaMemcpy(d.invarMask, im.data(), ps * sizeof(uint64_t), cudaMemcpyHostToDevice));
    }
    std::vector<uint64_t> em(p->nRealEquates > 0 ? p->nRealEquates : 1, 0);
    for (int e = 0; e < p->nEquates; e++) {
        const int j = p->equateColumn(e);
        if (j < 0) continue;
        for (int s = 0; s < dim; s++)
            if (p->equates[(size_t)e * dim + s]) em[j] |= 1ull << s;
    }
    CUDA_TRY(cudaMalloc(&d.equateMask, em.size() * sizeof(uint64_t)));
    CUDA_TRY(cudaMemcpy(d.equateMask, em.data(), em.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
    d.dataVersion = p->version;
    d.device = G.device;
    d.rank = G.rank;
    d.world = G.world;
    return 0;
}

// ---------------------------------------------------------------------------
// Tree device state
// ---------------------------------------------------------------------------
// One allocation of CL buffers ("slots") for one (tree, part).  The cur and prop tree of a chain hold
// the same numbers for every node a proposal did not touch, and p4 re-synchronises them after every
// generation by copying all CLs (p4_copyCondLikes, p4/chain.py:1529).  Here the two trees' arenas form a
// pair: a "copy" makes the destination node REFERENCE the source node's buffer (reference count per
// slot), and a node whose buffer is shared gets a fresh slot the next time it is computed.  No CL is
// ever moved between twins.
struct Arena {
    double *base = nullptr;
    size_t slotDoubles = 0;              // K*ps doubles of CL, then (scalers on) ps int32 exponents, padded to 256 B
    size_t clDoubles = 0;                // K*ps
    bool scalers = false;
    int ps = 0, nSlots = 0;
    std::vector<int> refs;               // per slot: nodes (of either twin) whose CL lives there
    std::vector<int> freeSlots;
    Arena *partner = nullptr;            // the one other arena this one shares buffers with
    ~Arena()
    {
        if (base) cudaFree(base);
    }
};

struct PartLayout {
    int dim = 0, nCat = 1, W = 0, ps = 0, nPat = 0;
    size_t clNodeDoubles = 0;    // nCat*dim*ps
    size_t pOff = 0, pDoubles = 0;       // within a node's P deck
    size_t tblOff = 0, tblDoubles = 0;   // within a node's leaf tables
    size_t auxOff = 0, auxDoubles = 0;   // within a node's operand decks of the tensor-core whole-tree kernels (0: none)
    int auxDP = 0;                       // 0: the 20-state kernel's deck layout; else the padded state count of cl_tree_dmma_kernel
    size_t eigOff = 0, eigStride = 0;    // within the eig mirror; stride per (comp,rMatrix)
    size_t eqOff = 0;                    // within the equate mask mirror
    int nPairs = 0;                      // nComps*nRMatrices
    std::shared_ptr<Arena> own, twin;    // CL buffers: this tree's arena, and the arena of the tree it shares buffers with
    bool scalers = false;
    Arena *arena(int sel) const { return sel ? twin.get() : own.get(); }
    std::vector<uint64_t> eigUploaded;   // version of each (comp,rMatrix) mirrored on the device
};

// The step list of a whole-tree launch depends on the topology, on which CL buffer every node uses and on what was
// asked; planning it (ordering, buffer allocation, operand offsets) costs tens of microseconds on the host, which at
// 8 GPUs is several percent of an evaluation.  A repeated evaluation of an unchanged tree reuses the plan.
struct PlanCache {
    bool valid = false;
    uint64_t topo = 0, slot = 0;
    bool withLike = false, storeAll = true;
    std::vector<Node *> order;          // the job it was made for
    std::vector<Node *> nodes;          // nodes that have steps, in step order (one entry per node)
    std::vector<char> resident;         // their clResident after the launch
    std::vector<Step2> steps;
    unsigned pf0 = 0xffffffffu;
};

struct TreeDevice {
    std::vector<PartLayout> parts;
    std::vector<uint64_t> dataVersion;   // per part: Part::version this device state was laid out for
    std::vector<std::vector<double>> rootCLHost;   // per part: host copy of the root's CL (p4_drawAncState), [K][ps]
    std::vector<uint64_t> rootCLStamp;             // ... and the computation id it is a copy of
    // Node-level CL calls on parts the whole-tree kernel serves are not launched one by one: they queue
    // here (per part, in call order) and run as ONE step-list launch when something needs their result --
    // normally p4_partLogLike, which then also gets the root reduction fused in.
    std::vector<std::vector<Node *>> pending;
    std::vector<char> likeBegun;   // per part: p4b_partLogLikeBegin has launched the evaluation; result not fetched yet
    std::vector<cudaEvent_t> evLike;   // per part: recorded after the begun evaluation's device->host copy
    bool scalers = false;
    size_t pNodeDoubles = 0, tblNodeDoubles = 0, auxNodeDoubles = 0;
    double *P = nullptr, *tbl = nullptr, *eig = nullptr, *aux = nullptr;
    uint64_t *eqMasks = nullptr;
    double *result = nullptr;     // [2*nParts] device
    double *hResult = nullptr;    // pinned
    double *partials = nullptr;   // [2*maxBlocks]
    unsigned *tickets = nullptr;  // [nParts] last-CTA tickets of the fused root reduction
    uint64_t slotStamp = 1;       // bumped whenever a node's CL buffer assignment (or the arena pairing) changes
    std::vector<struct PlanCache> plan;   // per part: the last whole-tree launch plan of the 4-state kernel (reused while nothing it depends on changed)
    int maxLikeBlocks = 0;
    double *patLikes = nullptr;
    int patLikesCap = 0;
    int *flag = nullptr;          // verify
    cudaEvent_t evA = nullptr, evB = nullptr, evCLa = nullptr, evCLb = nullptr;
    bool clTimed = false;
    int lastCLLaunches = 0;
    long long bytes = 0;
    struct NewtState *newt = nullptr;   // work arrays of the Newton-Raphson branch-length step (p4b_newtSetup)
};
static void newtStateFree(TreeDevice *d);

int treeDeviceCreate(Tree *t)
{
    if (engineInit()) return 1;
    if (!t->data || !t->model) { setError("p4_newTree: data and model are required"); return 1; }
    if (t->data->nParts != t->model->nParts) { setError("p4_newTree: data has %d parts, model %d", t->data->nParts, t->model->nParts); return 1; }
    TreeDevice *d = new TreeDevice();
    t->dev = d;
    d->scalers = g_useScalers;
    d->parts.resize(t->nParts);
    d->pending.resize(t->nParts);
    d->plan.resize(t->nParts);
    d->likeBegun.assign(t->nParts, 0);
    d->evLike.assign(t->nParts, nullptr);
    size_t eigTotal = 0, eqTotal = 0;
    const int nInternalSlots = t->nNodes - t->nLeaves + 1;   // +1: a root that is a leaf (Pf/p4_node.c:608-626)
    for (int p = 0; p < t->nParts; p++) {
        Part *dp = t->data->parts[p];
        ModelPart *mp = t->model->parts[p];
        if (!dp || !mp) { setError("p4_newTree: part %d missing in data or model", p); return 1; }
        if (dp->dim != mp->dim) { setError("p4_newTree: part %d dim mismatch data %d model %d", p, dp->dim, mp->dim); return 1; }
        if (partDeviceEnsure(dp)) return 1;
        d->dataVersion.push_back(dp->version);
        PartLayout &L = d->parts[p];
        L.dim = mp->dim;
        L.nCat = mp->nCat;
        L.W = dp->tableWidth();
        L.ps = dp->dev.ps;
        L.nPat = dp->dev.hi - dp->dev.lo;
        L.clNodeDoubles = (size_t)L.nCat * L.dim * L.ps;
        L.pOff = d->pNodeDoubles;
        L.pDoubles = (size_t)L.nCat * L.dim * L.dim;
        d->pNodeDoubles += (L.pDoubles + 1) & ~(size_t)1;      // every part's deck starts 16-byte aligned (bulk copies)
        L.tblOff = d->tblNodeDoubles;
        L.tblDoubles = (size_t)L.nCat * L.dim * L.W;
        d->tblNodeDoubles += (L.tblDoubles + 1) & ~(size_t)1;
        if (L.dim == 20 && L.nCat <= 16) {   // P^T in fragment order + transposed leaf table (kernels.cuh, pmatrix_kernel)
            L.auxOff = d->auxNodeDoubles;
            L.auxDoubles = (size_t)L.nCat * kAAFrag + (size_t)L.nCat * kAATblStates * L.W;
            d->auxNodeDoubles += L.auxDoubles;
        } else if (L.dim > 20 && L.dim <= 64 && !g_useScalers) {   // 21..64 states: the generic tensor-core kernel's decks (tree_dmma.cuh)
            L.auxDP = dmmaPaddedDim(L.dim);
            L.auxOff = d->auxNodeDoubles;
            L.auxDoubles = dmmaAuxDoubles(L.auxDP, L.nCat, L.W);
            d->auxNodeDoubles += L.auxDoubles;
        }
        L.nPairs = mp->nComps * mp->nRMatrices;
        L.eigStride = (size_t)2 * L.dim * L.dim + L.dim;
        L.eigOff = eigTotal;
        eigTotal += L.eigStride * L.nPairs;
        L.eigUploaded.assign(L.nPairs, 0);
        L.eqOff = eqTotal;
        eqTotal += dp->nRealEquates > 0 ? dp->nRealEquates : 1;
        L.own = std::make_shared<Arena>();
        Arena &A = *L.own;
        A.nSlots = nInternalSlots;
        A.ps = L.ps;
        A.clDoubles = L.clNodeDoubles;
        A.scalers = g_useScalers;
        A.slotDoubles = L.clNodeDoubles + (g_useScalers ? (((size_t)L.ps / 2 + 31) & ~(size_t)31) : 0);
        A.refs.assign(A.nSlots, 0);
        for (int s = A.nSlots - 1; s >= 0; s--) A.freeSlots.push_back(s);   // slot 0 is handed out first
        const size_t arenaBytes = A.slotDoubles * sizeof(double) * (size_t)A.nSlots;
        CUDA_TRY(cudaMalloc(&A.base, arenaBytes));
        CUDA_TRY(cudaMemsetAsync(A.base, 0, arenaBytes, G.stream));
        d->bytes += (long long)arenaBytes;
        L.scalers = g_useScalers;
        const int blocks = (L.ps + 255) / 256;
        if (blocks > d->maxLikeBlocks) d->maxLikeBlocks = blocks;
    }
    const size_t pBytes = d->pNodeDoubles * sizeof(double) * (size_t)t->nNodes;
    const size_t tblBytes = d->tblNodeDoubles * sizeof(double) * (size_t)t->nNodes;
    CUDA_TRY(cudaMalloc(&d->P, pBytes));
    CUDA_TRY(cudaMemsetAsync(d->P, 0, pBytes, G.stream));
    CUDA_TRY(cudaMalloc(&d->tbl, tblBytes));
    CUDA_TRY(cudaMemsetAsync(d->tbl, 0, tblBytes, G.stream));
    if (d->auxNodeDoubles) {
        const size_t auxBytes = d->auxNodeDoubles * sizeof(double) * (size_t)t->nNodes;
        CUDA_TRY(cudaMalloc(&d->aux, auxBytes));
        CUDA_TRY(cudaMemsetAsync(d->aux, 0, auxBytes, G.stream));
        d->bytes += (long long)auxBytes;
    }
    CUDA_TRY(cudaMalloc(&d->eig, eigTotal * sizeof(double)));
    CUDA_TRY(cudaMalloc(&d->eqMasks, eqTotal * sizeof(uint64_t)));
    for (int p = 0; p < t->nParts; p++) {
        Part *dp = t->data->parts[p];
        const int n = dp->nRealEquates > 0 ? dp->nRealEquates : 1;
        CUDA_TRY(cudaMemcpyAsync(d->eqMasks + d->parts[p].eqOff, dp->dev.equateMask, n * sizeof(uint64_t),
                                 cudaMemcpyDeviceToDevice, G.stream));
    }
    d->bytes += (long long)(pBytes + tblBytes + eigTotal * sizeof(double));
    CUDA_TRY(cudaMalloc(&d->result, 2 * sizeof(double) * t->nParts));
    CUDA_TRY(cudaMallocHost(&d->hResult, 2 * sizeof(double) * t->nParts));
    CUDA_TRY(cudaMalloc(&d->partials, 2 * sizeof(double) * (size_t)d->maxLikeBlocks * 8 * t->nParts));
    CUDA_TRY(cudaMalloc(&d->flag, sizeof(int)));
    CUDA_TRY(cudaMalloc(&d->tickets, sizeof(unsigned) * t->nParts));
    CUDA_TRY(cudaMemsetAsync(d->tickets, 0, sizeof(unsigned) * t->nParts, G.stream));
    CUDA_TRY(cudaEventCreate(&d->evA));
    CUDA_TRY(cudaEventCreate(&d->evB));
    CUDA_TRY(cudaEventCreate(&d->evCLa));
    CUDA_TRY(cudaEventCreate(&d->evCLb));
    return 0;
}

void treeDeviceDestroy(Tree *t)
{
    TreeDevice *d = t->dev;
    if (!d) return;
    flushPJobs();   // queued jobs may write into this tree's decks
    if (G.stream) cudaStreamSynchronize(G.stream);
    for (Node *n : t->nodes)
        if (n) nodeDeviceRelease(n);
    for (auto &L : d->parts) {
        // the arena's memory lives on while a twin still references buffers in it
        L.own.reset();
        L.twin.reset();
    }
    newtStateFree(d);
    if (d->P) cudaFree(d->P);
    if (d->tbl) cudaFree(d->tbl);
    if (d->aux) cudaFree(d->aux);
    if (d->eig) cudaFree(d->eig);
    if (d->eqMasks) cudaFree(d->eqMasks);
    if (d->result) cudaFree(d->result);
    if (d->hResult) cudaFreeHost(d->hResult);
    if (d->partials) cudaFree(d->partials);
    if (d->patLikes) cudaFree(d->patLikes);
    if (d->flag) cudaFree(d->flag);
    if (d->tickets) cudaFree(d->tickets);
    if (d->evA) cudaEventDestroy(d->evA);
    if (d->evB) cudaEventDestroy(d->evB);
    if (d->evCLa) cudaEventDestroy(d->evCLa);
    if (d->evCLb) cudaEventDestroy(d->evCLb);
    for (cudaEvent_t e : d->evLike)
        if (e) cudaEventDestroy(e);
    delete d;
    t->dev = nullptr;
}

static void slotRelease(Arena *A, int slot)
{
    if (!A || slot < 0) return;
    if (--A->refs[slot] == 0) A->freeSlots.push_back(slot);
}

// A free slot for a node of tree t: from the tree's own arena, else from its twin's.
static int slotAcquire(Tree *t, int p, int *sel, int *slot, int nodeNum)
{
    PartLayout &L = t->dev->parts[p];
    for (int s = 0; s < 2; s++) {
        Arena *A = L.arena(s);
        if (A && !A->freeSlots.empty()) {
            *sel = s;
            *slot = A->freeSlots.back();
            A->freeSlots.pop_back();
            A->refs[*slot] = 1;
            return 0;
        }
    }
    setError("node %d: no CL slot left (nNodes/nLeaves given to p4_newTree were wrong?)", nodeNum);
    return 1;
}

static int nodeEnsureCLSlot(Node *n, int p)
{
    if (n->clSlot[p] >= 0) return 0;
    int sel = 0, slot = -1;
    if (slotAcquire(n->tree, p, &sel, &slot, n->nodeNum)) return 1;
    n->clSel[p] = (char)sel;
    n->clSlot[p] = slot;
    n->tree->dev->slotStamp++;
    return 0;
}

// Before a node's CL is (re)computed: if its buffer is shared with the twin tree, move the node to a
// fresh slot; the twin keeps the old numbers.
static int nodeMakeWritable(Node *n, int p)
{
    if (nodeEnsureCLSlot(n, p)) return 1;
    PartLayout &L = n->tree->dev->parts[p];
    Arena *A = L.arena(n->clSel[p]);
    if (A->refs[n->clSlot[p]] <= 1) return 0;
    int sel = 0, slot = -1;
    if (slotAcquire(n->tree, p, &sel, &slot, n->nodeNum)) return 1;
    A->refs[n->clSlot[p]]--;
    n->clSel[p] = (char)sel;
    n->clSlot[p] = slot;
    n->tree->dev->slotStamp++;
    return 0;
}

void nodeDeviceRelease(Node *n)
{
    Tree *t = n->tree;
    if (!t || !t->dev) return;
    for (int p = 0; p < t->nParts && p < (int)n->clSlot.size(); p++) {
        if (n->clSlot[p] < 0) continue;
        slotRelease(t->dev->parts[p].arena(n->clSel[p]), n->clSlot[p]);
        n->clSlot[p] = -1;
        t->dev->slotStamp++;
    }
}

int nodeDeviceCreate(Node *n)
{
    Tree *t = n->tree;
    n->clSlot.assign(t->nParts, -1);
    n->clSel.assign(t->nParts, 0);
    n->clStamp.assign(t->nParts, 0);
    n->clResident.assign(t->nParts, 1);
    n->pStamp.assign(t->nParts, 0);
    n->pKey.assign(t->nParts, std::vector<double>());
    n->clKey.assign(t->nParts, std::vector<uint64_t>());
    if (!n->isLeaf)
        for (int p = 0; p < t->nParts; p++)
            if (nodeEnsureCLSlot(n, p)) return 1;
    return 0;
}

// The data parts can change under a living tree: p4 simulates new sequences into them and re-compresses
// (Tree.simulate -> pf.p4_simulate, pf.makePatterns; p4/tree.py:9617-9626).  The reference's arrays are sized by
// nChar and survive that; here the pattern shard, its stride and the CL arenas depend on the pattern count, so
// the tree's device state is laid out again the next time it is used (and every P deck recomputed).
int treeCalculateAllBigPDecks(Tree *t);
static int ensureFresh(Tree *t, bool needPatterns)
{
    TreeDevice *d = t->dev;
    if (!d) return 0;
    bool stale = false;
    for (int p = 0; p < t->nParts; p++) {
        Part *dp = t->data->parts[p];
        // a mirror built before p4b_setShard / p4b_commInitRank moved this process to another pattern range
        if (dp->dev.tips && (dp->dev.rank != G.rank || dp->dev.world != G.world)) dp->version++;
        if (dp->version != d->dataVersion[p]) stale = true;
    }
    if (!stale) return 0;
    for (int p = 0; p < t->nParts; p++)
        if (t->data->parts[p]->nPatterns <= 0) {
            if (!needPatterns) return 0;      // nothing that needs patterns is being asked for yet
            setError("part %d has no patterns (pf.makePatterns not called after the sequences changed?)", p);
            return 1;
        }
    treeDeviceDestroy(t);
    // on any failure leave NO device state behind: later calls then report "tree has no device state"
    // instead of walking a half-built layout
    if (treeDeviceCreate(t)) { treeDeviceDestroy(t); return 1; }
    for (Node *n : t->nodes)
        if (n && nodeDeviceCreate(n)) { treeDeviceDestroy(t); return 1; }
    return treeCalculateAllBigPDecks(t);
}

static inline double *nodeCL(Node *n, int p)
{
    PartLayout &L = n->tree->dev->parts[p];
    Arena *A = L.arena(n->clSel[p]);
    return A->base + A->slotDoubles * (size_t)n->clSlot[p];
}
static inline int *nodeSC(Node *n, int p)   // the buffer's exponents sit right behind its CL
{
    PartLayout &L = n->tree->dev->parts[p];
    return L.scalers ? reinterpret_cast<int *>(nodeCL(n, p) + L.clNodeDoubles) : nullptr;
}
// Base address the whole-tree kernel addresses this tree's CL buffers from: the lower of its own
// arena and its twin's.
static inline double *arenaBase(const PartLayout &L)
{
    return (L.twin && L.twin->base < L.own->base) ? L.twin->base : L.own->base;
}
// (buffer - base) / 256 bytes: the form the whole-tree kernel addresses a CL buffer by (30 bits)
static inline unsigned nodeSlotCode(Node *n, int p)
{
    return (unsigned)((nodeCL(n, p) - arenaBase(n->tree->dev->parts[p])) / 32);
}
static inline double *nodeP(Node *n, int p)
{
    TreeDevice *d = n->tree->dev;
    return d->P + d->pNodeDoubles * (size_t)n->nodeNum + d->parts[p].pOff;
}
static inline double *nodeAux(Node *n, int p)
{
    TreeDevice *d = n->tree->dev;
    return d->aux + d->auxNodeDoubles * (size_t)n->nodeNum + d->parts[p].auxOff;
}
static inline double *nodeTbl(Node *n, int p)
{
    TreeDevice *d = n->tree->dev;
    return d->tbl + d->tblNodeDoubles * (size_t)n->nodeNum + d->parts[p].tblOff;
}

// ---------------------------------------------------------------------------
// setPrams: host part (Pf/p4_tree.c:218-501) then the batched P(t) kernel
// ---------------------------------------------------------------------------
static int eigEnsureUploaded(Tree *t, int p, int cNum, int rNum)
{
    ModelPart *mp = t->model->parts[p];
    PartLayout &L = t->dev->parts[p];
    const int idx = cNum * mp->nRMatrices + rNum;
    Eig &e = mp->bqe[idx];
    if (!e.allocated) { setError("part %d comp %d rMatrix %d has no eigensystem", p, cNum, rNum); return 1; }
    if (L.eigUploaded[idx] == e.version) return 0;
    if (flushPJobs()) return 1;   // queued jobs must still see the eigensystem they were issued with
    const int dim = L.dim;
    std::vector<double> buf(L.eigStride);
    memcpy(buf.data(), e.V.data(), sizeof(double) * dim * dim);
    memcpy(buf.data() + dim * dim, e.Vinv.data(), sizeof(double) * dim * dim);
    memcpy(buf.data() + 2 * dim * dim, e.lam.data(), sizeof(double) * dim);
    void *dsrc = nullptr;
    if (stage(buf.data(), buf.size() * sizeof(double), &dsrc)) return 1;
    CUDA_TRY(cudaMemcpyAsync(t->dev->eig + L.eigOff + L.eigStride * idx, dsrc, buf.size() * sizeof(double),
                             cudaMemcpyDeviceToDevice, G.stream));
    L.eigUploaded[idx] = e.version;
    return 0;
}

static int hostSetPramsPart(Tree *t, int p)
{
    Model *m = t->model;
    ModelPart *mp = m->parts[p];
    // gamma rates of free gdasrvs (Pf/p4_tree.c:236-258)
    if (mp->nCat > 1)
        for (Gdasrv *g : mp->gdasrvs)
            if (g && g->isFree) discreteGamma(g->freqs, g->rates, g->val[0], g->val[0], mp->nCat, 0);
    // 2-parameter rate matrices (:297-322)
    for (RMatrix &r : mp->rMatrices)
        if (r.isFree && r.spec == 5) setKappaBigR(r);
    // composition checks (:424-452)
    for (int i = 0; i < mp->nComps; i++) {
        const double *v = mp->comps[i].val;
        if (!v) { setError("p4_setPramsPart() part %d, comp %d was never created", p, i); return 1; }
        for (int j = 0; j < mp->dim; j++)
            if (v[j] < m->PIVEC_MIN[0]) {
                setError("p4_setPramsPart()  part %d, comp %d, value %d is %g   Bad.", p, i, j, v[j]);
                return 1;
            }
        double sum = 0.0;
        for (int j = 0; j < mp->dim; j++) sum += v[j];
        if (fabs(sum - 1.0) > 1e-14) {
            setError("**p4_setPramsPart()  part %d, comp %d, values do not sum to 1.0.  sum=%g  sum - 1.0 = %g", p, i, sum, sum - 1.0);
            return 1;
        }
    }
    // mark every (comp,rMatrix) stale, then rebuild the ones a non-root node uses (:455-501)
    for (int i = 0; i < mp->nComps * mp->nRMatrices; i++) mp->bQETneedsReset[i] = 1;
    for (Node *n : t->nodes) {
        if (!n || n == t->root) continue;
        const int c = n->compNums[p], r = n->rMatrixNums[p];
        if (c < 0 || c >= mp->nComps || r < 0 || r >= mp->nRMatrices) { setError("node %d part %d uses comp %d rMatrix %d which do not exist", n->nodeNum, p, c, r); return 1; }
        if (mp->bQETneedsReset[c * mp->nRMatrices + r])
            if (resetBQET(m, p, c, r)) return 1;
    }
    return 0;
}

// Append the P(t) job of (node, part) to the batch.
static int buildPJob(Node *n, int p, bool allowSkip)
{
    std::vector<PJob> &jobs = G.pJobs;
    std::vector<double> &tvals = G.pT;
    Tree *t = n->tree;
    ModelPart *mp = t->model->parts[p];
    TreeDevice *d = t->dev;
    PartLayout &L = d->parts[p];
    const int c = n->compNums[p], r = n->rMatrixNums[p];
    if (c < 0 || c >= mp->nComps || r < 0 || r >= mp->nRMatrices) { setError("node %d part %d uses comp %d rMatrix %d which do not exist", n->nodeNum, p, c, r); return 1; }
    if (mp->bQETneedsReset[c * mp->nRMatrices + r]) {   // Pf/p4_node.c:310-315: self-heal with a warning
        printf("p4_calculateBigPDecksPart() pNum=%i, compNum=%i, rMatrixNum=%i, needsReset. Fix me.\n", p, c, r);
        if (resetBQET(t->model, p, c, r)) return 1;
    }
    if (eigEnsureUploaded(t, p, c, r)) return 1;
    const Gdasrv *g = nullptr;
    if (mp->nGdasrvs) {
        const int gi = n->gdasrvNums[p];
        if (gi < 0 || gi >= mp->nGdasrvs || !mp->gdasrvs[gi]) { setError("node %d part %d uses gdasrv %d which does not exist", n->nodeNum, p, gi); return 1; }
        g = mp->gdasrvs[gi];
    }
    PJob j;
    j.P = d->P + d->pNodeDoubles * (size_t)n->nodeNum + L.pOff;
    j.tbl = d->tbl + d->tblNodeDoubles * (size_t)n->nodeNum + L.tblOff;
    j.eig = d->eig + L.eigOff + L.eigStride * (size_t)(c * mp->nRMatrices + r);
    j.eq = d->eqMasks + L.eqOff;
    j.aux = L.auxDoubles ? d->aux + d->auxNodeDoubles * (size_t)n->nodeNum + L.auxOff : nullptr;
    j.dim = L.dim;
    j.nCat = L.nCat;
    j.auxDP = L.auxDP;
    j.pad = 0;
    j.tblW = n->isLeaf ? L.W : 0;
    j.nRealEq = L.W - L.dim - 1;
    j.tOff = (long long)tvals.size();
    std::vector<double> key(1 + mp->nCat);
    key[0] = (double)mp->bqe[c * mp->nRMatrices + r].content;     // which solve of which (pi, R)
    for (int cat = 0; cat < mp->nCat; cat++) {   // Pf/p4_node.c:321-345, same expressions
        double tt;
        if (mp->pInvar == 0.0) tt = g ? (n->brLen * g->rates[cat] * mp->relRate) : (n->brLen * mp->relRate);
        else tt = g ? (n->brLen * g->rates[cat] * mp->relRate) / (1.0 - mp->pInvar) : (n->brLen * mp->relRate) / (1.0 - mp->pInvar);
        key[1 + cat] = tt;
    }
    // P(t) is a pure function of (eigensystem, t per category): the same inputs again give the deck already there
    if (g_memoize && allowSkip && n->pStamp[p] != 0 && key[0] != 0.0 && n->pKey[p] == key) return 0;
    for (int cat = 0; cat < mp->nCat; cat++) tvals.push_back(key[1 + cat]);
    jobs.push_back(j);
    n->pKey[p].swap(key);
    n->pStamp[p] = ++G.stamp;
    return 0;
}

static int flushPJobs()
{
    std::vector<PJob> &jobs = G.pJobs;
    std::vector<double> &tvals = G.pT;
    if (jobs.empty()) return 0;
    // one staged block: [tvals | jobs]; job.tOff is relative to the block start
    const size_t tBytes = ((tvals.size() * sizeof(double)) + 15) & ~(size_t)15;
    std::vector<char> blk(tBytes + jobs.size() * sizeof(PJob));
    memcpy(blk.data(), tvals.data(), tvals.size() * sizeof(double));
    memcpy(blk.data() + tBytes, jobs.data(), jobs.size() * sizeof(PJob));
    // shared memory: exp(lambda t) per category for every job, and for the jobs that fit 64 KB (every 4- and 20-state part)
    // the eigensystem, the P deck and the leaf table as well (kernels.cuh, pmatrixStageDoubles)
    size_t maxSm = 0;
    static const bool stageP = [] { const char *e = getenv("P4B_PMATRIX_STAGE"); return e ? atoi(e) != 0 : true; }();
    for (auto &j : jobs) {
        size_t need = pmatrixStageDoubles(j.dim, j.nCat, j.tblW);
        if (!stageP || need * sizeof(double) > 64 * 1024) need = (size_t)j.nCat * j.dim;
        if (need > maxSm) maxSm = need;
    }
    static bool pAttr = false;
    if (!pAttr) {
        CUDA_TRY(cudaFuncSetAttribute(pmatrix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        pAttr = true;
    }
    const unsigned nJobs = (unsigned)jobs.size();
    jobs.clear();
    tvals.clear();
    void *dblk = nullptr;
    if (stage(blk.data(), blk.size(), &dblk)) return 1;
    pmatrix_kernel<<<nJobs, 128, maxSm * sizeof(double), G.stream>>>(
        reinterpret_cast<const PJob *>((char *)dblk + tBytes), reinterpret_cast<const double *>(dblk), (int)maxSm);
    CUDA_TRY(cudaGetLastError());
    G.launches++;
    return 0;
}

static int treeFlushAllPending(Tree *t);

int treeSetPrams(Tree *t, int pNum)
{
    if (!t->dev) { setError("tree has no device state"); return 1; }
    if (ensureFresh(t, false)) return 1;
    if (pNum < -1 || pNum >= t->nParts) { setError("p4_setPrams: bad part %d", pNum); return 1; }
    if (treeFlushAllPending(t)) return 1;   // queued CLs were issued against the old P decks
    for (int p = 0; p < t->nParts; p++) {
        if (pNum >= 0 && p != pNum) continue;
        if (hostSetPramsPart(t, p)) return 1;
        // mirror every changed eigensystem first (an upload flushes the job queue), then queue all jobs
        ModelPart *mp = t->model->parts[p];
        for (Node *n : t->nodes)
            if (n && n != t->root && !mp->bQETneedsReset[n->compNums[p] * mp->nRMatrices + n->rMatrixNums[p]])
                if (eigEnsureUploaded(t, p, n->compNums[p], n->rMatrixNums[p])) return 1;
        for (Node *n : t->nodes)
            if (n && n != t->root)
                if (buildPJob(n, p, true)) return 1;
    }
    return 0;
}

int nodeCalculateBigPDecks(Node *n)
{
    Tree *t = n->tree;
    if (treeFlushAllPending(t)) return 1;
    for (int p = 0; p < t->nParts; p++)
        if (buildPJob(n, p, true)) return 1;
    return 0;
}

int treeCalculateAllBigPDecks(Tree *t)   // Pf/p4_tree.c p4_calculateAllBigPDecksAllParts
{
    if (treeFlushAllPending(t)) return 1;
    for (int p = 0; p < t->nParts; p++)
        for (Node *n : t->nodes)
            if (n && n != t->root)
                if (buildPJob(n, p, true)) return 1;
    return 0;
}

// ---------------------------------------------------------------------------
// Conditional likelihoods
// ---------------------------------------------------------------------------
static bool g_dmmaEnabled = true;
static const bool g_newtDmma = [] { const char *e = getenv("P4B_NEWT_DMMA"); return e ? atoi(e) != 0 : true; }();   // 0: the FMA derivative kernel for 20-state internal nodes too
void setDmmaEnabled(int on) { g_dmmaEnabled = on != 0; }
static bool g_fusedEnabled = true;
static bool g_deferCL = true;
void setMemoizeEnabled(int on) { g_memoize = on != 0; }
void setDeferEnabled(int on) { g_deferCL = on != 0; }

static bool g_fusedAAEnabled = true;
void setFusedAAEnabled(int on) { g_fusedAAEnabled = on != 0; }

// second-generation 4-state kernel in use?  (P4B_FUSED2=0 or a forced first-generation launch shape switch it off)
static int g_fusedVariant = -1;   // p4b_setFusedVariant: -1 = by shard size
static bool g_fused2On()
{
    static int env2 = -1;
    if (env2 < 0) { const char *e = getenv("P4B_FUSED2"); env2 = e ? atoi(e) != 0 : 1; }
    return env2 && !(g_fusedVariant >= 0 && g_fusedVariant < 10);
}

// Parts the whole-tree kernels serve: 4 states (FMA kernels), 20 states with 4 categories and 21..64 states (tensor-core kernels).
static bool fusedEligible(const PartLayout &L)
{
    if (!g_fusedEnabled) return false;
    if (L.dim == 4) return L.nCat == 4 || L.nCat == 1 || (!L.scalers && L.nCat >= 2 && L.nCat <= 8 && g_fused2On());
    if (L.dim == 20) return g_fusedAAEnabled && g_dmmaEnabled && L.nCat <= 16 && !L.scalers && L.W <= 64;
    if (L.dim > 20 && L.dim <= 64)    // two children x two buffers of max(P^T fragments, transposed leaf table) must fit shared memory
        return g_fusedAAEnabled && g_dmmaEnabled && !L.scalers && L.auxDP > 0 && L.nCat <= 64 &&
               4 * sizeof(double) * std::max(dmmaFragDoubles(L.auxDP), (size_t)L.W * L.auxDP) + 64 <= 200 * 1024;
    return false;
}

// Shared memory of the tensor-core kernel: A fragments of the internal children
// (15 fragments x 32 lanes per category) and the leaf children's lookup tables.
static size_t dmmaSmemBytes(const CLArgs &a)
{
    int nInt = 0, nLeaf = 0;
    for (int c = 0; c < a.nChildren; c++) (a.ch[c].tips ? nLeaf : nInt)++;
    return ((size_t)nInt * a.nCat * 15 * 32 + (size_t)nLeaf * a.nCat * 20 * a.tblW) * sizeof(double);
}

static int launchCL(const CLArgs &a)
{
    if (flushPJobs()) return 1;
    const int maxPer = a.dim * (a.tblW > a.dim ? a.tblW : a.dim);
    if (a.dim == 4 && a.nCat >= 1 && a.nCat <= 8) {
        const int K = a.nCat * 4;
        const size_t sm = (size_t)a.nChildren * K * (a.tblW > 4 ? a.tblW : 4) * sizeof(double);
        const int pairs = a.ps / 2;
        const dim3 grid((pairs + 255) / 256);
        switch (a.nCat) {
        case 1: cl_dna_kernel<1><<<grid, 256, sm, G.stream>>>(a); break;
        case 2: cl_dna_kernel<2><<<grid, 256, sm, G.stream>>>(a); break;
        case 3: cl_dna_kernel<3><<<grid, 256, sm, G.stream>>>(a); break;
        case 4: cl_dna_kernel<4><<<grid, 256, sm, G.stream>>>(a); break;
        case 5: cl_dna_kernel<5><<<grid, 256, sm, G.stream>>>(a); break;
        case 6: cl_dna_kernel<6><<<grid, 256, sm, G.stream>>>(a); break;
        case 7: cl_dna_kernel<7><<<grid, 256, sm, G.stream>>>(a); break;
        default: cl_dna_kernel<8><<<grid, 256, sm, G.stream>>>(a); break;
        }
    } else if (a.dim == 20 && g_dmmaEnabled && dmmaSmemBytes(a) <= 200 * 1024) {
        static bool attrSet = false;
        if (!attrSet) {
            CUDA_TRY(cudaFuncSetAttribute(cl_dmma20_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attrSet = true;
        }
        int nGrid = a.ps / kDmmaGroup;
        if (nGrid > G.numSMs * 3) nGrid = G.numSMs * 3;
        cl_dmma20_kernel<<<nGrid, 128, dmmaSmemBytes(a), G.stream>>>(a);
    } else if (a.dim == 20) {
        const size_t sm = (size_t)a.nChildren * maxPer * sizeof(double);
        const dim3 grid((a.ps + 127) / 128, a.nCat);
        static bool attrSet = false;
        if (!attrSet) {
            CUDA_TRY(cudaFuncSetAttribute(cl_dim_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            attrSet = true;
        }
        if (sm > 96 * 1024) { setError("leaf table too wide for the 20-state kernel"); return 1; }
        cl_dim_kernel<20><<<grid, 128, sm, G.stream>>>(a);
    } else {
        const dim3 grid((a.ps + 127) / 128, a.nCat);
        cl_generic_kernel<<<grid, 128, 0, G.stream>>>(a);
    }
    CUDA_TRY(cudaGetLastError());
    snprintf(g_lastKernel, sizeof(g_lastKernel), "per-node CL kernels (dim %d, nCat %d)", a.dim, a.nCat);
    G.launches++;
    return 0;
}

// What a node's CL is a pure function of: its children in order, each child's CL (by computation id) or tip
// row, each child's P deck (by computation id), and the data.
static void makeClKey(Node *n, int p, std::vector<uint64_t> &key)
{
    key.clear();
    key.push_back(n->tree->data->parts[p]->version);
    for (Node *c = n->leftChild; c; c = c->sibling) {
        key.push_back((uint64_t)c->nodeNum);
        key.push_back(c->isLeaf ? (uint64_t)c->seqNum : c->clStamp[p]);
        key.push_back(c->pStamp[p]);
    }
}
static bool clIsCurrent(Node *n, int p, const std::vector<uint64_t> &key)
{
    return g_memoize && n->clSlot[p] >= 0 && n->clResident[p] && n->clStamp[p] != 0 && n->clKey[p] == key;
}

static int nodeSetCLImpl(Node *n, int p, bool memo);
int nodeSetCL(Node *n, int p) { return nodeSetCLImpl(n, p, true); }

static int nodeSetCLImpl(Node *n, int p, bool memo)
{
    Tree *t = n->tree;
    if (ensureFresh(t, true)) return 1;
    TreeDevice *d = t->dev;
    if (p < 0 || p >= t->nParts) { setError("p4_setConditionalLikelihoodsOfInternalNodePart: bad part %d", p); return 1; }
    if (!n->leftChild) { setError("node %d has no children; cannot set its conditional likelihoods", n->nodeNum); return 1; }
    PartLayout &L = d->parts[p];
    if (memo && g_deferCL && fusedEligible(L)) {
        if (nodeEnsureCLSlot(n, p)) return 1;   // a root that is a leaf gets its CL lazily, Pf/p4_node.c:608-626
        // The callers issue node-level calls in dependency order (the dirty set is decided in
        // Python): queue, and run the whole queue as one step-list launch when its result is needed.
        d->pending[p].push_back(n);
        n->clNeedsUpdating = 0;
        return 0;
    }
    if (treeEnsureResident(t, p)) return 1;
    std::vector<uint64_t> key;
    makeClKey(n, p, key);
    if (memo && clIsCurrent(n, p, key)) { n->clNeedsUpdating = 0; return 0; }   // same inputs as last time: the CL in memory is the answer
    if (nodeMakeWritable(n, p)) return 1;
    Part *dp = t->data->parts[p];
    CLArgs a;
    memset(&a, 0, sizeof(a));
    a.out = nodeCL(n, p);
    a.ps = L.ps;
    a.dim = L.dim;
    a.nCat = L.nCat;
    a.tblW = L.W;
    a.accumulate = 0;
    int k = 0;
    for (Node *c = n->leftChild; c; c = c->sibling) {
        CLChild &ch = a.ch[k];
        if (c->isLeaf) {
            if (c->seqNum < 0 || c->seqNum >= dp->nTax) { setError("leaf node %d has seqNum %d", c->nodeNum, c->seqNum); return 1; }
            ch.tips = dp->dev.tips + (size_t)c->seqNum * L.ps;
            ch.tbl = nodeTbl(c, p);
            ch.cl = nullptr;
            ch.P = nullptr;
        } else {
            if (c->clSlot[p] < 0) { setError("internal node %d has no conditional likelihoods", c->nodeNum); return 1; }
            ch.cl = nodeCL(c, p);
            ch.P = nodeP(c, p);
            ch.tips = nullptr;
            ch.tbl = nullptr;
        }
        k++;
        if (k == kMaxChildren && c->sibling) {   // polytomy wider than one launch: chain
            a.nChildren = k;
            if (launchCL(a)) return 1;
            d->lastCLLaunches++;
            a.accumulate = 1;
            k = 0;
        }
    }
    if (k > 0) {
        a.nChildren = k;
        if (launchCL(a)) return 1;
        d->lastCLLaunches++;
    }
    if (L.scalers) {   // scalers on: one extra pass rescales the node and sums the children's exponents
        RescaleArgs r;
        memset(&r, 0, sizeof(r));
        r.cl = nodeCL(n, p);
        r.scOut = nodeSC(n, p);
        r.nRows = L.nCat * L.dim;
        r.ps = L.ps;
        for (Node *c = n->leftChild; c; c = c->sibling)
            if (!c->isLeaf) {
                if (r.nScChild == 16) { setError("scalers: node %d has more than 16 internal children", n->nodeNum); return 1; }
                r.scChild[r.nScChild++] = nodeSC(c, p);
            }
        rescale_kernel<<<(L.ps + 255) / 256, 256, 0, G.stream>>>(r);
        CUDA_TRY(cudaGetLastError());
        G.launches++;
    }
    n->clStamp[p] = ++G.stamp;
    n->clKey[p].swap(key);
    n->clResident[p] = 1;
    n->clNeedsUpdating = 0;
    return 0;
}

// ---------------------------------------------------------------------------
// Whole-tree recursion in one launch (4-state parts)
// ---------------------------------------------------------------------------

static int enqueueRootLike(Tree *t, int p, bool wantPatLikes, double *resultDev);

// One job of a batched whole-tree launch: the nodes of one tree to compute, in dependency order.
struct FusedJob {
    Tree *t;
    const std::vector<Node *> *order;
    bool withLike;       // fuse the root reduction (order must end at the root)
    bool wantPatLikes;
    bool storeAll;       // false: lnL-only evaluation, keep only the CLs the launch itself re-reads
    bool memo;           // queued node-level calls: a node whose inputs are those of its current CL is skipped
};

// Launch shape of the whole-tree kernel: threads per CTA x CTAs per SM.  0: 128x3, 1: 64x6, 2: 32x12 (all
// 166-168 registers, no spills); 3: 128x4 and 4: 256x2 are kept for comparison.  Measured on B200
// (tools/sweep_shapes.sh, 200 taxa): 1 M patterns 5.14 / 5.24 / 6.33 ms for shapes 0 / 1 / 2, 500 k
// 2.70 / 2.82 / 3.21, 250 k 1.56 / 1.42 / 1.65, 125 k 0.99 / 0.90 / 0.86.  What decides is how many waves
// of 128-thread CTAs the launch makes: the last, partial wave leaves SMs idle unless the CTAs are small.
// P4B_FUSED_VARIANT overrides the choice (tuning).
int setFusedVariant(int v)
{
    if (v < -1 || (v > 8 && v < 10) || v > 16) { setError("p4b_setFusedVariant: launch shape %d does not exist (-1, 0..8, 10..16)", v); return 1; }
    g_fusedVariant = v;
    return 0;
}
static int fusedVariant(int ps, int nTrees)
{
    static int forced = -2;
    if (forced == -2) {
        const char *e = getenv("P4B_FUSED_VARIANT");
        forced = e ? atoi(e) : -1;
        if (forced < -1 || forced > 8) forced = -1;
    }
    if (g_fusedVariant >= 0 && g_fusedVariant <= 8) return g_fusedVariant;
    if (forced >= 0) return forced;
    const double waves = (double)nTrees * (double)(ps / 2) / (128.0 * 3.0 * G.numSMs);
    return waves >= 3.0 ? 0 : (waves >= 1.6 ? 1 : 2);
}

// Append the steps of one job to the argument block.  Returns the number of steps written, or -1 on
// error, or -2 when they do not fit into `room` steps (nothing is modified in that case beyond a.steps).
static int buildSteps(TreeArgs &a, int base, int room, const FusedJob &job, int p, bool *overflowOk, size_t *resumeAt, int maxKids)
{
    Tree *t = job.t;
    Part *dp = t->data->parts[p];
    const std::vector<Node *> &order = *job.order;
    int ns = 0;
    Node *prev = nullptr;   // node whose CL the previous step of THIS launch left in registers
    std::unordered_map<Node *, int> lastStepOf;   // node -> index of the step that finishes it (this launch)
    std::unordered_set<Node *> needsMemory;       // nodes some step loads from the arena
    size_t oi = resumeAt ? *resumeAt : 0;
    for (; oi < order.size(); oi++) {
        Node *n = order[oi];
        if (!n->leftChild) { setError("node %d has no children; cannot set its conditional likelihoods", n->nodeNum); return -1; }
        std::vector<uint64_t> key;
        makeClKey(n, p, key);
        // never skip the node the fused root reduction reads from registers
        if (job.memo && !(job.withLike && oi + 1 == order.size()) && clIsCurrent(n, p, key)) { n->clNeedsUpdating = 0; continue; }
        int nKids = 0;
        for (Node *c = n->leftChild; c; c = c->sibling) nKids++;
        const int chunks = (nKids + maxKids - 1) / maxKids;
        if (ns + chunks > room) {
            if (!overflowOk || !*overflowOk) return -2;
            break;   // the caller launches what we have and calls again from *resumeAt
        }
        if (nodeMakeWritable(n, p)) return -1;   // a buffer shared with the twin tree is left to the twin
        StepC *st = &a.steps[base + ns];
        st->outSlot = (int)nodeSlotCode(n, p);
        st->first = 1;
        int k = 0;
        bool prevUsed = false;
        for (Node *c = n->leftChild; c; c = c->sibling) {
            unsigned kind, index;
            if (c->isLeaf) {
                if (c->seqNum < 0 || c->seqNum >= dp->nTax) { setError("leaf node %d has seqNum %d", c->nodeNum, c->seqNum); return -1; }
                kind = 2u;
                index = (unsigned)c->seqNum;
            } else {
                if (c->clSlot[p] < 0) { setError("internal node %d has no conditional likelihoods", c->nodeNum); return -1; }
                kind = (c == prev && !prevUsed && st->first) ? 1u : 0u;
                if (kind == 1u) prevUsed = true;
                else needsMemory.insert(c);
                index = nodeSlotCode(c, p);
            }
            st->ch[k].a = (int)((kind << 30) | index);
            st->ch[k].b = c->nodeNum;
            k++;
            if (k == maxKids && c->sibling) {   // node wider than one step: continue in the next
                st->nChildren = (short)k;
                st->store = 0;
                ns++;
                st = &a.steps[base + ns];
                st->outSlot = (int)nodeSlotCode(n, p);
                st->first = 0;
                k = 0;
            }
        }
        st->nChildren = (short)k;
        st->store = 1;
        if (k == 2 && st->first && maxKids > 2) {
            // two-children step (4-state kernel): its one re-loaded internal child (if any) goes through the prefetch buffer
            const unsigned k0 = (unsigned)st->ch[0].a >> 30, k1 = (unsigned)st->ch[1].a >> 30;
            if (k0 == 0u && k1 != 0u) st->ch[0].a = (int)((3u << 30) | ((unsigned)st->ch[0].a & 0x3fffffffu));
            else if (k1 == 0u && k0 != 0u) st->ch[1].a = (int)((3u << 30) | ((unsigned)st->ch[1].a & 0x3fffffffu));
        }
        lastStepOf[n] = ns;
        ns++;
        prev = n;
        n->clStamp[p] = ++G.stamp;
        n->clKey[p].swap(key);
        n->clResident[p] = job.storeAll ? 1 : 0;
        n->clNeedsUpdating = 0;
    }
    if (resumeAt) *resumeAt = oi;
    if (overflowOk) *overflowOk = oi < order.size();
    if (!job.storeAll) {
        // lnL-only evaluation: keep the store only for nodes that a later step of this launch
        // re-reads from memory (kind 0); everything else lives and dies in registers.
        for (int i = 0; i < ns; i++) a.steps[base + i].store = 0;
        for (auto &kv : lastStepOf)
            if (needsMemory.count(kv.first)) {
                a.steps[base + kv.second].store = 1;
                kv.first->clResident[p] = 1;
            }
    }
    return ns;
}

// ---------------------------------------------------------------------------
// Second-generation whole-tree kernel for 4-state parts (tree_dna.cuh): step planning and launch
// ---------------------------------------------------------------------------
static bool g_fused2 = true;           // P4B_FUSED2=0 / p4b_setFusedVariant(0..8) select the first-generation kernel
static bool g_heavyFirst = true;       // P4B_HEAVY_FIRST=0 keeps the caller's post-order
static bool g_pushBuf = true;          // P4B_PUSH=0: siblings always travel through global memory (prefetched)

// Any order in which children precede parents computes the same numbers.  The caller's post-order leaves a node's
// FIRST-visited internal child waiting in memory while the other child's whole subtree is computed; visiting the
// heavier subtree first makes that wait as short as possible, so the re-read (or the push buffer) finds it close:
// in L2, or in shared memory.  Only the order of the launch's steps changes, never the order of a node's children.
static void reorderHeavyFirst(std::vector<Node *> &order)
{
    const size_t n = order.size();
    if (n < 3) return;
    std::unordered_map<Node *, int> w;
    w.reserve(n * 2);
    for (Node *x : order) w[x] = 0;
    for (Node *x : order) {           // children precede parents in a valid order
        int sum = 1;
        for (Node *c = x->leftChild; c; c = c->sibling) {
            auto it = w.find(c);
            if (it != w.end()) sum += it->second;
        }
        w[x] = sum;
    }
    std::vector<Node *> out;
    out.reserve(n);
    struct Frame { Node *n; std::vector<Node *> kids; size_t next; };
    std::vector<Frame> stack;
    auto open = [&](Node *x) {
        Frame f{x, {}, 0};
        for (Node *c = x->leftChild; c; c = c->sibling)
            if (w.count(c)) f.kids.push_back(c);
        std::stable_sort(f.kids.begin(), f.kids.end(), [&](Node *a, Node *b) { return w[a] > w[b]; });
        stack.push_back(std::move(f));
    };
    for (Node *r : order) {
        if (r->parent && w.count(r->parent)) continue;      // not a root of the set
        open(r);
        while (!stack.empty()) {
            Frame &f = stack.back();
            if (f.next < f.kids.size()) { Node *c = f.kids[f.next++]; open(c); }
            else { out.push_back(f.n); stack.pop_back(); }
        }
    }
    if (out.size() == n) order.swap(out);
}

// Steps of one job, appended to `steps`; fills the job's header fields that depend on them.  Returns the number of
// steps, or -1 on error.
// mode 0: 4-state kernel (P decks / leaf tables, shared-memory buffer); mode 1: 20-state kernel (operand decks, no buffer).
static int buildSteps2(std::vector<Step2> &steps, TreeHdr2 &h, const FusedJob &job, int p, int mode = 0)
{
    Tree *t = job.t;
    Part *dp = t->data->parts[p];
    std::vector<Node *> order = *job.order;
    if (g_heavyFirst) reorderHeavyFirst(order);
    const size_t base = steps.size();
    std::unordered_map<Node *, int> doneAt;          // node -> step (relative) that finishes it in this launch
    std::vector<std::array<unsigned, 2>> tipRows;    // per step: tip rows of its leaf children
    std::vector<std::array<Node *, 2>> kidsOf;       // per step: the children themselves
    std::vector<Node *> nodeOf;
    Node *prev = nullptr;
    for (size_t oi = 0; oi < order.size(); oi++) {
        Node *n = order[oi];
        if (!n->leftChild) { setError("node %d has no children; cannot set its conditional likelihoods", n->nodeNum); return -1; }
        std::vector<uint64_t> key;
        makeClKey(n, p, key);
        // never skip the node the fused root reduction reads from registers
        if (job.memo && !(job.withLike && n == order.back()) && clIsCurrent(n, p, key)) { n->clNeedsUpdating = 0; continue; }
        if (nodeMakeWritable(n, p)) return -1;   // a buffer shared with the twin tree is left to the twin
        int k = 0;
        bool firstChunk = true, prevUsed = false;
        Step2 st;
        auto begin = [&]() {
            memset(&st, 0, sizeof(st));
            st.out = nodeSlotCode(n, p);
            st.c0 = st.c1 = 0;
            st.nt0 = st.nt1 = st.pf = kNone;
            tipRows.push_back({kNone, kNone});
            kidsOf.push_back({nullptr, nullptr});
            nodeOf.push_back(n);
        };
        auto finish = [&](bool last) {
            st.flags |= (unsigned)k | (firstChunk ? kStepFirst : 0u) | (last ? kStepStore : 0u);
            steps.push_back(st);
        };
        begin();
        for (Node *c = n->leftChild; c; c = c->sibling) {
            unsigned kind;
            if (c->isLeaf) {
                if (c->seqNum < 0 || c->seqNum >= dp->nTax) { setError("leaf node %d has seqNum %d", c->nodeNum, c->seqNum); return -1; }
                kind = 2u;
                tipRows.back()[k] = (unsigned)c->seqNum;
            } else {
                if (c->clSlot[p] < 0) { setError("internal node %d has no conditional likelihoods", c->nodeNum); return -1; }
                kind = (c == prev && !prevUsed && firstChunk) ? 1u : 0u;
                if (kind == 1u) prevUsed = true;
                (k == 0 ? st.c0 : st.c1) = nodeSlotCode(c, p);
            }
            (k == 0 ? st.n0 : st.n1) = (unsigned)c->nodeNum;
            st.flags |= kind << (4 + 2 * k);
            kidsOf.back()[k] = c;
            k++;
            if (k == 2 && c->sibling) {          // node wider than one step: continue in the next
                finish(false);
                firstChunk = false;
                k = 0;
                begin();
            }
        }
        finish(true);
        doneAt[n] = (int)(steps.size() - base) - 1;
        prev = n;
        n->clStamp[p] = ++G.stamp;
        n->clKey[p].swap(key);
        n->clResident[p] = 1;
        n->clNeedsUpdating = 0;
    }
    const int ns = (int)(steps.size() - base);
    Step2 *S = steps.data() + base;
    // operand addresses as 32-bit offsets (doubles) into the tree's leaf tables / P decks
    {
        TreeDevice *d = t->dev;
        for (int j = 0; j < ns; j++) {
            const unsigned k0 = (S[j].flags >> 4) & 3u, k1 = (S[j].flags >> 6) & 3u;
            unsigned long long o0, o1;
            if (mode == 1) {       // operand deck of the node; a leaf's transposed tables lie behind the P^T fragments
                const PartLayout &L = d->parts[p];
                const unsigned long long leafOff = (unsigned long long)L.nCat * kAAFrag;
                o0 = (unsigned long long)d->auxNodeDoubles * S[j].n0 + (k0 == 2u ? leafOff : 0ull);
                o1 = (unsigned long long)d->auxNodeDoubles * S[j].n1 + (k1 == 2u ? leafOff : 0ull);
            } else {
                o0 = (unsigned long long)(k0 == 2u ? d->tblNodeDoubles : d->pNodeDoubles) * S[j].n0;
                o1 = (unsigned long long)(k1 == 2u ? d->tblNodeDoubles : d->pNodeDoubles) * S[j].n1;
            }
            if (o0 >> 32 || o1 >> 32) { setError("internal: operand decks larger than 32 GB"); return -1; }
            S[j].n0 = (unsigned)o0;
            S[j].n1 = (unsigned)o1;
        }
    }
    h.t0 = h.t1 = h.pf0 = kNone;
    for (int j = 0; j < ns; j++) { S[j].nt0 = tipRows[j][0]; S[j].nt1 = tipRows[j][1]; }   // a leaf child's tip codes travel with its lookup table
    // The per-thread shared-memory buffer: in consumption order, give each step's first in-memory child the buffer if it
    // is free -- pushed by the step that produces it when nothing else needs the buffer in between (no global re-read),
    // else prefetched from global memory up to three steps ahead; a child that gets neither is loaded when it is needed.
    std::unordered_set<Node *> needsMemory;
    int lastUse = -1;                    // step whose computation last reads the buffer
    for (int j = 0; j < ns; j++) {
        bool taken = mode == 1;          // the 20-state kernel has no buffer: in-memory children are loaded when they are needed
        for (int i = 0; i < (int)(S[j].flags & 3u); i++) {
            const unsigned kind = (S[j].flags >> (4 + 2 * i)) & 3u;
            if (kind != 0u) continue;
            Node *x = kidsOf[j][i];
            auto it = doneAt.find(x);
            const int s = it == doneAt.end() ? -1 : it->second;
            const unsigned code = i == 0 ? S[j].c0 : S[j].c1;
            bool viaBuffer = false;
            if (!taken) {
                if (g_pushBuf && s >= 0 && lastUse <= s) {
                    S[s].flags |= kStepPush;
                    viaBuffer = true;
                } else if (j == 0) {
                    h.pf0 = code;
                    needsMemory.insert(x);
                    viaBuffer = true;
                } else {
                    int e = s + 1;
                    if (e < lastUse) e = lastUse;
                    if (e < j - 3) e = j - 3;
                    if (e < 0) e = 0;
                    if (e <= j - 1 && S[e].pf == kNone) {
                        S[e].pf = code;
                        if (e == lastUse) S[e].flags |= kStepPfLate;     // step e reads the buffer itself: refill it afterwards
                        needsMemory.insert(x);
                        viaBuffer = true;
                    }
                }
            }
            if (viaBuffer) {
                S[j].flags = (S[j].flags & ~(3u << (4 + 2 * i))) | (3u << (4 + 2 * i));
                taken = true;
                lastUse = j;
            } else {
                needsMemory.insert(x);
            }
        }
    }
    // Canonical child order for the kernel's straight-line bodies.  The product of the two children's factors is
    // commutative bit for bit, so the children of a two-children first step can be handed over in either order: the
    // one in registers first, else the one in the buffer, leaves last -- four combinations instead of seven.
    for (int j = 0; j < ns; j++) {
        if ((S[j].flags & (3u | kStepFirst)) != (2u | kStepFirst)) continue;
        const unsigned k0 = (S[j].flags >> 4) & 3u, k1 = (S[j].flags >> 6) & 3u;
        auto rank = [mode](unsigned k) { return mode == 1 ? (k == 1u ? 0 : (k == 0u ? 1 : 2)) : (k == 1u ? 0 : (k == 3u ? 1 : (k == 2u ? 2 : 3))); };
        if (rank(k1) < rank(k0)) {
            std::swap(S[j].c0, S[j].c1);
            std::swap(S[j].nt0, S[j].nt1);
            std::swap(S[j].n0, S[j].n1);
            S[j].flags = (S[j].flags & ~0xf0u) | (k1 << 4) | (k0 << 6);
        }
    }
    if (!job.storeAll) {
        // lnL-only evaluation: a node is written to its buffer only if a later step of this launch reads it from there
        for (int j = 0; j < ns; j++) S[j].flags &= ~kStepStore;
        for (auto &kv : doneAt) {
            if (needsMemory.count(kv.first)) { S[kv.second].flags |= kStepStore; kv.first->clResident[p] = 1; }
            else kv.first->clResident[p] = 0;
        }
    }
    h.nSteps = ns;
    return ns;
}

// Launch shapes of the second-generation kernel: {categories per thread, warps per CTA, CTAs per SM}.
struct Shape2 { int ct, cw, minb; };
static const Shape2 kShapes2[] = {
    {4, 4, 3},    // 10: 128 threads x 3
    {4, 2, 6},    // 11:  64 threads x 6
    {4, 1, 12},   // 12:  32 threads x 12
    {2, 4, 4},    // 13: categories split over 2 warps, 128 threads x 4
    {2, 2, 8},    // 14:  64 threads x 8
    {1, 4, 6},    // 15: one category per warp, 128 threads x 6
    {2, 8, 2},    // 16: 256 threads x 2
};
constexpr int kNumShapes2 = (int)(sizeof(kShapes2) / sizeof(kShapes2[0]));
typedef void (*Kernel2Fn)(const TreeArgs2);
// Rate-category counts other than 1 and 4 (the reference is generic in nCat, Pf/p4_node.c:652-654): one shape each,
// with the categories split over warps where that keeps a thread at four categories or fewer.
static bool shapeForNCat(int nCat, Shape2 *sh)
{
    switch (nCat) {
    case 2: *sh = {2, 4, 4}; return true;
    case 3: *sh = {3, 4, 3}; return true;
    case 5: *sh = {1, 5, 4}; return true;
    case 6: *sh = {3, 4, 3}; return true;
    case 7: *sh = {1, 7, 3}; return true;
    case 8: *sh = {4, 4, 3}; return true;
    default: return false;
    }
}
static Kernel2Fn kernel2For(int nCat, int shape)
{
    static const Kernel2Fn k4[kNumShapes2] = {cl_tree_dna2_kernel<4, 4, 4, 3>, cl_tree_dna2_kernel<4, 4, 2, 6>, cl_tree_dna2_kernel<4, 4, 1, 12>,
                                              cl_tree_dna2_kernel<4, 2, 4, 4>, cl_tree_dna2_kernel<4, 2, 2, 8>, cl_tree_dna2_kernel<4, 1, 4, 6>,
                                              cl_tree_dna2_kernel<4, 2, 8, 2>};
    static const Kernel2Fn k1[3] = {cl_tree_dna2_kernel<1, 1, 4, 4>, cl_tree_dna2_kernel<1, 1, 2, 8>, cl_tree_dna2_kernel<1, 1, 1, 16>};
    switch (nCat) {
    case 4: return k4[shape];
    case 1: return k1[shape < 3 ? shape : 0];
    case 2: return cl_tree_dna2_kernel<2, 2, 4, 4>;
    case 3: return cl_tree_dna2_kernel<3, 3, 4, 3>;
    case 5: return cl_tree_dna2_kernel<5, 1, 5, 4>;
    case 6: return cl_tree_dna2_kernel<6, 3, 4, 3>;
    case 7: return cl_tree_dna2_kernel<7, 1, 7, 3>;
    case 8: return cl_tree_dna2_kernel<8, 4, 4, 3>;
    default: return nullptr;
    }
}

static int fused2Shape(int ps, int nTrees)
{
    if (g_fusedVariant >= 10 && g_fusedVariant < 10 + kNumShapes2) return g_fusedVariant - 10;
    static int forced = -2;
    if (forced == -2) {
        const char *e = getenv("P4B_FUSED2_SHAPE");
        forced = e ? atoi(e) : -1;
        if (forced < -1 || forced >= kNumShapes2) forced = -1;
    }
    if (forced >= 0) return forced;
    // Measured on B200 (tools/sweep_dna.py, 200 taxa; profiles/r2_sweep_dna.txt): shape 16 -- two rate categories per
    // thread, the four categories of a pattern block split over two warps, 256 threads x 2 CTAs/SM = 16 warps/SM at 128
    // registers -- is the fastest at every shard size: 1 M patterns 4.61 ms (10: 5.08, 13: 4.73), 250 k 1.21 (13: 1.25),
    // 125 k 0.67 (13: 0.70, 14: 0.71); the first-generation kernel 4.98 / 1.35 / 0.82.
    (void)ps;
    (void)nTrees;
    return 6;
}

// Make `steps` the content of the device step buffer (skipped when it already is).
static int uploadSteps2(const std::vector<Step2> &steps)
{
    const size_t n = steps.size(), bytes = n * sizeof(Step2);
    if (n == 0) return 0;
    if (G.stepShadow.size() == n && memcmp(G.stepShadow.data(), steps.data(), bytes) == 0) return 0;
    if (G.stepDevCap < n) {
        if (G.stepDev) { CUDA_TRY(cudaStreamSynchronize(G.stream)); cudaFree(G.stepDev); G.stepDev = nullptr; }
        size_t cap = n < 1024 ? 1024 : n * 2;
        CUDA_TRY(cudaMalloc(&G.stepDev, cap * sizeof(Step2)));
        G.stepDevCap = cap;
    }
    const int k = G.stepSlabNext;
    G.stepSlabNext = (k + 1) & 3;
    if (!G.stepEv[k]) CUDA_TRY(cudaEventCreateWithFlags(&G.stepEv[k], cudaEventDisableTiming));
    else CUDA_TRY(cudaEventSynchronize(G.stepEv[k]));       // the slab's previous upload has left the host
    if (G.stepSlabCap[k] < n) {
        if (G.stepSlab[k]) cudaFreeHost(G.stepSlab[k]);
        size_t cap = n < 1024 ? 1024 : n * 2;
        CUDA_TRY(cudaMallocHost(&G.stepSlab[k], cap * sizeof(Step2)));
        G.stepSlabCap[k] = cap;
    }
    memcpy(G.stepSlab[k], steps.data(), bytes);
    CUDA_TRY(cudaMemcpyAsync(G.stepDev, G.stepSlab[k], bytes, cudaMemcpyHostToDevice, G.stream));
    CUDA_TRY(cudaEventRecord(G.stepEv[k], G.stream));
    G.stepShadow = steps;
    return 0;
}

static int launchFused2Batch(const FusedJob *jobs, int nJobs, int p, double *resultDev)
{
    if (flushPJobs()) return 1;
    Tree *t0 = jobs[0].t;
    TreeDevice *d0 = t0->dev;
    PartLayout &L = d0->parts[p];
    Part *dp = t0->data->parts[p];
    static TreeArgs2 a;
    memset(&a, 0, sizeof(a));
    a.ps = L.ps;
    a.nPat = L.nPat;
    a.tblW = L.W;
    a.nTrees = nJobs;
    a.tileLog = 0;
    a.pNodeDoubles = (long long)d0->pNodeDoubles;
    a.tblNodeDoubles = (long long)d0->tblNodeDoubles;
    a.tips = dp->dev.tips;
    a.counts = dp->dev.counts;
    a.invarMask = dp->dev.invarMask;
    a.eqMask = dp->dev.equateMask;
    if (dp->nTax >= 65535) { setError("internal: more than 65534 sequences"); return 1; }
    const int shape = fused2Shape(L.ps, nJobs);
    const int shape1 = shape == 1 ? 1 : (shape == 2 ? 2 : 0);      // one rate category: 128 threads unless a small-CTA shape is forced
    Shape2 sh = L.nCat == 4 ? kShapes2[shape] : Shape2{1, shape1 == 0 ? 4 : (shape1 == 1 ? 2 : 1), shape1 == 0 ? 4 : (shape1 == 1 ? 8 : 16)};
    if (L.nCat != 4 && L.nCat != 1 && !shapeForNCat(L.nCat, &sh)) { setError("internal: no whole-tree kernel for %d rate categories", L.nCat); return 1; }
    const int csplit = L.nCat / sh.ct;
    const int patsPerCta = (sh.cw / csplit) * 64;
    const int blocks = (L.ps + patsPerCta - 1) / patsPerCta;
    static std::vector<Step2> steps;
    steps.clear();
    int maxSteps = 1;
    for (int i = 0; i < nJobs; i++) {
        Tree *t = jobs[i].t;
        TreeDevice *d = t->dev;
        PartLayout &Li = d->parts[p];
        ModelPart *mp = t->model->parts[p];
        if (t->data->parts[p] != dp || Li.ps != L.ps || Li.nCat != L.nCat || Li.dim != L.dim || Li.W != L.W || d->pNodeDoubles != d0->pNodeDoubles ||
            d->tblNodeDoubles != d0->tblNodeDoubles || Li.scalers != L.scalers) {
            setError("batched evaluation: the trees do not share the data part and model shape");
            return 1;
        }
        TreeHdr2 &h = a.hdr[i];
        h.arena = arenaBase(Li);
        h.Pdeck = d->P + Li.pOff;
        h.tbl = d->tbl + Li.tblOff;
        h.result = resultDev + 2 * i;
        h.ticket = nJobs == 1 ? d->tickets + p : G.batchTickets + i;
        h.partials = d->partials + (size_t)2 * d->maxLikeBlocks * 8 * p;
        if (blocks > d->maxLikeBlocks * 8) { setError("internal: partial buffer too small"); return 1; }
        if (jobs[i].withLike) {
            Node *root = t->root;
            if (!root || jobs[i].order->empty() || jobs[i].order->back() != root) { setError("fused evaluation: the last node must be the root"); return 1; }
            const int rc = root->compNums[p];
            if (rc < 0 || rc >= mp->nComps || !mp->comps[rc].val) { setError("root uses comp %d which does not exist", rc); return 1; }
            h.rootTips = root->isLeaf ? dp->dev.tips + (size_t)root->seqNum * L.ps : nullptr;
            h.pInvar = mp->pInvar;
            if (h.pInvar != 0.0 && !a.invarMask) { setError("pInvar is set but pf.setGlobalInvarSitesVec was not called on part %d", p); return 1; }
            for (int s = 0; s < 4; s++) h.pi[s] = mp->comps[rc].val[s];
            if (jobs[i].wantPatLikes) {
                if (d->patLikesCap < L.ps) {
                    if (d->patLikes) cudaFree(d->patLikes);
                    CUDA_TRY(cudaMalloc(&d->patLikes, sizeof(double) * L.ps));
                    d->patLikesCap = L.ps;
                }
                h.patLikes = d->patLikes;
            }
            h.doLike = 1;
        }
        h.stepBase = (int)steps.size();
        int ns = -1;
        PlanCache *pc = (nJobs == 1 && !jobs[i].memo) ? &d->plan[p] : nullptr;
        if (pc && pc->valid && pc->topo == t->topoStamp && pc->withLike == jobs[i].withLike && pc->storeAll == jobs[i].storeAll && pc->order == *jobs[i].order) {
            // the plan of the last launch may still hold: every node must keep the buffer it had
            bool ok = true;
            for (Node *n : pc->nodes)
                if (nodeMakeWritable(n, p)) return 1;
            ok = pc->slot == d->slotStamp;
            if (ok) {
                steps.insert(steps.end(), pc->steps.begin(), pc->steps.end());
                for (size_t k = 0; k < pc->nodes.size(); k++) {
                    Node *n = pc->nodes[k];
                    n->clStamp[p] = ++G.stamp;
                    n->clKey[p].clear();          // (no input record: a later node-level call on it recomputes rather than skips)
                    n->clResident[p] = pc->resident[k];
                    n->clNeedsUpdating = 0;
                }
                h.pf0 = pc->pf0;
                h.t0 = h.t1 = kNone;
                h.nSteps = ns = (int)pc->steps.size();
            }
        }
        if (ns < 0) {
            ns = buildSteps2(steps, h, jobs[i], p);
            if (ns < 0) return 1;
            if (pc) {
                pc->valid = true;
                pc->topo = t->topoStamp;
                pc->slot = d->slotStamp;
                pc->withLike = jobs[i].withLike;
                pc->storeAll = jobs[i].storeAll;
                pc->order = *jobs[i].order;
                pc->steps.assign(steps.begin() + h.stepBase, steps.end());
                pc->pf0 = h.pf0;
                pc->nodes.clear();
                pc->resident.clear();
                for (Node *n : *jobs[i].order) {       // every node of a non-memoised job has steps
                    pc->nodes.push_back(n);
                    pc->resident.push_back(n->clResident[p]);
                }
            }
        }
        if (ns > maxSteps) maxSteps = ns;
    }
    bool anything = false;
    for (int i = 0; i < nJobs; i++)
        if (a.hdr[i].nSteps > 0 || a.hdr[i].doLike) anything = true;
    if (!anything) return 0;
    a.maxSteps = maxSteps;
    const size_t smem = treeDna2SmemBytes(L.nCat, L.W, sh.ct, sh.cw, maxSteps);
    if (smem > 200 * 1024) { setError("internal: step list of %d steps does not fit the whole-tree kernel's shared memory", maxSteps); return 1; }
    if (uploadSteps2(steps)) return 1;
    a.steps = G.stepDev;
    bool anyLike2 = false;
    for (int i = 0; i < nJobs; i++) anyLike2 = anyLike2 || a.hdr[i].doLike;
    if (anyLike2) a.mail = nextMail();
    else a.mail.world = 1;
    Kernel2Fn fn = kernel2For(L.nCat, L.nCat == 1 ? shape1 : shape);
    static std::unordered_set<void *> attrSet;
    if (!attrSet.count((void *)fn)) {
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attrSet.insert((void *)fn);
    }
    if (blocks * patsPerCta != L.ps) { setError("internal: pattern stride %d is not a multiple of the CTA tile", L.ps); return 1; }
    fn<<<dim3(blocks, nJobs), sh.cw * 32, smem, G.stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    snprintf(g_lastKernel, sizeof(g_lastKernel), "cl_tree_dna2_kernel<%d,%d,%d,%d>", L.nCat, sh.ct, sh.cw, sh.minb);
    G.launches++;
    for (int i = 0; i < nJobs; i++) jobs[i].t->dev->lastCLLaunches++;
    return 0;
}

// ---------------------------------------------------------------------------
// 20 states: the tensor-core whole-tree kernel (tree_aa.cuh)
// ---------------------------------------------------------------------------
// jobs[i] is (tree, part partOf[i]) -- or part p for all of them when partOf is NULL (the batched trees of one part); the root
// reduction of job i writes resultOf[i] (NULL: resultDev + 2 i).
static int launchFusedAABatch(const FusedJob *jobs, int nJobs, int p, double *resultDev, const int *partOf = nullptr, double *const *resultOf = nullptr)
{
    if (flushPJobs()) return 1;
    if (nJobs > kMaxBatchTrees) { setError("internal: more than %d (tree, part) pairs in one launch", kMaxBatchTrees); return 1; }
    static TreeArgsAA a;
    memset(&a, 0, sizeof(a));
    a.nTrees = nJobs;
    // launch shape "warps per CTA,ring depth,CTAs per SM": 8,8,2 unless P4B_AA_SHAPE says otherwise
    static int shapeRead = 0, dbgNoStore = 0, groups3 = 8, ring3 = 8, minb3 = 2;
    if (!shapeRead) {
        shapeRead = 1;
        const char *e = getenv("P4B_AA_SHAPE");
        if (e && sscanf(e, "%d,%d,%d", &groups3, &ring3, &minb3) != 3) { groups3 = 8; ring3 = 8; minb3 = 2; }
        e = getenv("P4B_AA_NOSTORE");      // measurement only: the CLs of an earlier evaluation stay in the arena
        dbgNoStore = e ? atoi(e) : 0;
    }
    a.pad0 = dbgNoStore;
    static std::vector<Step2> steps;
    steps.clear();
    int maxSteps = 1, maxW = 1, maxPs = 0, maxCat = 1;
    for (int i = 0; i < nJobs; i++) {
        Tree *t = jobs[i].t;
        TreeDevice *d = t->dev;
        const int pi = partOf ? partOf[i] : p;
        PartLayout &Li = d->parts[pi];
        Part *dp = t->data->parts[pi];
        if (Li.dim != 20 || !d->aux || Li.auxDP != 0 || !Li.auxDoubles) { setError("internal: 20-state whole-tree kernel without operand decks"); return 1; }
        if (dp->nTax >= 65535) { setError("internal: more than 65534 sequences"); return 1; }
        if (jobs[i].withLike && (!t->root || jobs[i].order->empty() || jobs[i].order->back() != t->root)) { setError("fused evaluation: the last node must be the root"); return 1; }
        TreeArgsAA::Hdr &hd = a.hdr[i];
        hd.arena = arenaBase(Li);
        hd.aux = d->aux + Li.auxOff;
        hd.tips = dp->dev.tips;
        hd.ps = Li.ps;
        hd.tblW = Li.W;
        hd.nCat = Li.nCat;
        hd.stepBase = (int)steps.size();
        TreeHdr2 h;
        memset(&h, 0, sizeof(h));
        const int ns = buildSteps2(steps, h, jobs[i], pi, 1);
        if (ns < 0) return 1;
        if (!jobs[i].storeAll && jobs[i].withLike && ns > 0) {      // lnL-only: the root reduction reads the root's CL from memory
            steps.back().flags |= kStepStore;
            t->root->clResident[pi] = 1;
        }
        hd.nSteps = ns;
        if (ns > maxSteps) maxSteps = ns;
        if (ns > 0) {
            if (Li.W > maxW) maxW = Li.W;
            if (Li.ps > maxPs) maxPs = Li.ps;
            if (Li.nCat > maxCat) maxCat = Li.nCat;
        }
    }
    if (maxPs > 0) {
        a.maxSteps = maxSteps;
        if (uploadSteps2(steps)) return 1;
        a.steps = G.stepDev;
        typedef void (*Fn)(const TreeArgsAA);
        static std::unordered_set<void *> attrSet;
        auto prepare = [&](Fn fn) {
            if (attrSet.count((void *)fn)) return 0;
            CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            attrSet.insert((void *)fn);
            return 0;
        };
        struct Shape { int g, r, b; Fn fn; };
        static const Shape shapes[] = {
            {8, 8, 2, (Fn)cl_tree_aa_kernel<8, 8, 2>}, {8, 4, 2, (Fn)cl_tree_aa_kernel<8, 4, 2>}, {4, 4, 4, (Fn)cl_tree_aa_kernel<4, 4, 4>}, {16, 8, 1, (Fn)cl_tree_aa_kernel<16, 8, 1>},
        };
        const Shape *sh = nullptr;
        for (const Shape &c : shapes)
            if (c.g == groups3 && c.r == ring3 && c.b == minb3) sh = &c;
        if (!sh) { setError("P4B_AA_SHAPE: no such launch shape of the 20-state whole-tree kernel"); return 1; }
        size_t smem = aaSmemBytes(maxW, sh->g, sh->r, maxSteps);
        if ((smem + 1024) * sh->b > 227 * 1024 && sh->g == 8 && sh->r == 8) {      // wide leaf tables or a long step list: the shallower ring
            sh = &shapes[1];
            smem = aaSmemBytes(maxW, sh->g, sh->r, maxSteps);
        }
        if (smem > 220 * 1024) { setError("internal: leaf tables or step list too large for the 20-state whole-tree kernel's shared memory"); return 1; }
        if (prepare(sh->fn)) return 1;
        for (int i = 0; i < nJobs; i++) a.hdr[i].nBlocks = a.hdr[i].nSteps > 0 ? a.hdr[i].ps / (sh->g * 16) : 0;
        const int blocks = maxPs / (sh->g * 16);
        sh->fn<<<dim3(blocks, nJobs, maxCat), sh->g * 32, smem, G.stream>>>(a);
        CUDA_TRY(cudaGetLastError());
        snprintf(g_lastKernel, sizeof(g_lastKernel), "cl_tree_aa_kernel<%d,%d,%d> x %d categories", sh->g, sh->r, sh->b, maxCat);
        G.launches++;
        for (int i = 0; i < nJobs; i++)
            if (i == 0 || jobs[i].t != jobs[i - 1].t) jobs[i].t->dev->lastCLLaunches++;
    }
    for (int i = 0; i < nJobs; i++)
        if (jobs[i].withLike)
            if (enqueueRootLike(jobs[i].t, partOf ? partOf[i] : p, jobs[i].wantPatLikes, resultOf ? resultOf[i] : resultDev + 2 * i)) return 1;
    return 0;
}

// ---------------------------------------------------------------------------
// 21..64 states: the generic tensor-core whole-tree kernel (tree_dmma.cuh)
// ---------------------------------------------------------------------------
static int buildSteps(TreeArgs &a, int base, int room, const FusedJob &job, int p, bool *overflowOk, size_t *resumeAt, int maxKids);
static int launchFusedDmmaBatch(const FusedJob *jobs, int nJobs, int p, double *resultDev)
{
    if (flushPJobs()) return 1;
    Tree *t0 = jobs[0].t;
    TreeDevice *d0 = t0->dev;
    PartLayout &L = d0->parts[p];
    Part *dp = t0->data->parts[p];
    static TreeArgs a;
    memset(&a, 0, offsetof(TreeArgs, steps));
    a.ps = L.ps;
    a.nPat = L.nPat;
    a.tblW = L.W;
    a.pNodeDoubles = (long long)d0->pNodeDoubles;
    a.tblNodeDoubles = (long long)d0->tblNodeDoubles;
    a.auxNodeDoubles = (long long)d0->auxNodeDoubles;
    a.tips = dp->dev.tips;
    a.counts = dp->dev.counts;
    a.invarMask = dp->dev.invarMask;
    a.eqMask = dp->dev.equateMask;
    if (!d0->aux || !L.auxDP) { setError("internal: generic tensor-core kernel without operand decks"); return 1; }
    static int mt = -1, warps = 8;
    if (mt < 0) {
        const char *e = getenv("P4B_DMMA_MT");
        mt = e ? atoi(e) : 2;
        if (mt != 1 && mt != 2) mt = 2;
        warps = mt == 1 ? 16 : 8;
    }
    const int DP = L.auxDP;
    const int patsPerCta = warps * 8 * mt;
    const int blocks = (L.ps + patsPerCta - 1) / patsPerCta;
    for (int i = 0; i < nJobs; i++) {
        Tree *t = jobs[i].t;
        TreeDevice *d = t->dev;
        PartLayout &Li = d->parts[p];
        if (t->data->parts[p] != dp || Li.ps != L.ps || Li.nCat != L.nCat || Li.dim != L.dim || Li.W != L.W || d->auxNodeDoubles != d0->auxNodeDoubles || Li.auxDP != L.auxDP) {
            setError("batched evaluation: the trees do not share the data part and model shape");
            return 1;
        }
        TreeHdr &h = a.hdr[i];
        h.arena = arenaBase(Li);
        h.Pdeck = d->P + Li.pOff;
        h.tbl = d->tbl + Li.tblOff;
        h.aux = d->aux + Li.auxOff;
        if (jobs[i].withLike && (!t->root || jobs[i].order->empty() || jobs[i].order->back() != t->root)) { setError("fused evaluation: the last node must be the root"); return 1; }
    }
    const size_t slot = std::max(dmmaFragDoubles(DP), (size_t)L.W * DP);
    const size_t smem = 2 * kAAKids * slot * sizeof(double) + 64;
    typedef void (*Fn)(const TreeArgs, const int, const int);
    Fn fn = DP == 32 ? (mt == 1 ? (Fn)cl_tree_dmma_kernel<32, 1, 16> : (Fn)cl_tree_dmma_kernel<32, 2, 8>)
                     : (mt == 1 ? (Fn)cl_tree_dmma_kernel<64, 1, 16> : (Fn)cl_tree_dmma_kernel<64, 2, 8>);
    static std::unordered_set<void *> attrSet;
    if (!attrSet.count((void *)fn)) {
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attrSet.insert((void *)fn);
    }
    auto launch = [&](int nTrees) -> int {
        a.nTrees = nTrees;
        fn<<<dim3(blocks, nTrees, L.nCat), warps * 32, smem, G.stream>>>(a, L.dim, L.nCat);
        CUDA_TRY(cudaGetLastError());
        snprintf(g_lastKernel, sizeof(g_lastKernel), "cl_tree_dmma_kernel<%d,%d,%d>", DP, mt, warps);
        G.launches++;
        for (int i = 0; i < nJobs; i++) jobs[i].t->dev->lastCLLaunches++;
        return 0;
    };
    if (nJobs == 1) {
        size_t at = 0;
        for (;;) {
            bool more = true;
            const int ns = buildSteps(a, 0, kMaxSteps, jobs[0], p, &more, &at, kAAKids);
            if (ns < 0) return 1;
            a.hdr[0].stepBase = 0;
            a.hdr[0].nSteps = ns;
            if (ns > 0 && launch(1)) return 1;
            if (!more) break;
        }
    } else {
        int base = 0;
        for (int i = 0; i < nJobs; i++) {
            const int ns = buildSteps(a, base, kMaxSteps - base, jobs[i], p, nullptr, nullptr, kAAKids);
            if (ns == -2) { setError("batched evaluation: the step lists of the trees do not fit one launch"); return 1; }
            if (ns < 0) return 1;
            a.hdr[i].stepBase = base;
            a.hdr[i].nSteps = ns;
            base += ns;
        }
        if (launch(nJobs)) return 1;
    }
    for (int i = 0; i < nJobs; i++)
        if (jobs[i].withLike)
            if (enqueueRootLike(jobs[i].t, p, jobs[i].wantPatLikes, resultDev + 2 * i)) return 1;
    return 0;
}

// Whole-tree kernel over one or several trees that share data part p (same shard, same model shape).
// With withLike jobs the per-tree results land in resultDev[2*i] (sum of count*log like) and
// resultDev[2*i+1] (count of non-positive likelihoods) for job i.
static int launchFusedBatch(const FusedJob *jobs, int nJobs, int p, double *resultDev)
{
    if (nJobs < 1 || nJobs > kMaxBatchTrees) { setError("internal: bad batch size %d", nJobs); return 1; }
    {
        const PartLayout &L0 = jobs[0].t->dev->parts[p];
        if (L0.dim == 4 && !L0.scalers && g_fused2 && g_fused2On()) return launchFused2Batch(jobs, nJobs, p, resultDev);
        if (L0.dim == 4 && L0.nCat != 4 && L0.nCat != 1) { setError("internal: the first-generation whole-tree kernel serves 1 or 4 rate categories"); return 1; }
        if (L0.dim > 20) return launchFusedDmmaBatch(jobs, nJobs, p, resultDev);
        if (L0.dim == 20) return launchFusedAABatch(jobs, nJobs, p, resultDev);
    }
    if (flushPJobs()) return 1;
    Tree *t0 = jobs[0].t;
    TreeDevice *d0 = t0->dev;
    PartLayout &L = d0->parts[p];
    Part *dp = t0->data->parts[p];
    static TreeArgs a;   // ~30 KB: too big for the stack of a small thread; the engine is single-threaded
    memset(&a, 0, offsetof(TreeArgs, steps));
    a.ps = L.ps;
    a.nPat = L.nPat;
    a.tblW = L.W;
    a.pNodeDoubles = (long long)d0->pNodeDoubles;
    a.tblNodeDoubles = (long long)d0->tblNodeDoubles;
    a.auxNodeDoubles = (long long)d0->auxNodeDoubles;
    a.tips = dp->dev.tips;
    a.counts = dp->dev.counts;
    a.invarMask = dp->dev.invarMask;
    a.eqMask = dp->dev.equateMask;
    // (20-state parts have their own launch path: launchFusedAABatch)
    const int variant = fusedVariant(L.ps, nJobs);
    static const int kThreads[9] = {128, 64, 32, 128, 256, 64, 32, 64, 32};
    const int THREADS = kThreads[variant];
    const int blocks = (L.ps / 2 + THREADS - 1) / THREADS;
    const int maxKids = kMaxChildren;
    bool anyLike = false;
    for (int i = 0; i < nJobs; i++) {
        Tree *t = jobs[i].t;
        TreeDevice *d = t->dev;
        PartLayout &Li = d->parts[p];
        ModelPart *mp = t->model->parts[p];
        if (t->data->parts[p] != dp || Li.ps != L.ps || Li.nCat != L.nCat || Li.dim != L.dim || Li.W != L.W || d->pNodeDoubles != d0->pNodeDoubles ||
            d->tblNodeDoubles != d0->tblNodeDoubles || d->auxNodeDoubles != d0->auxNodeDoubles || Li.scalers != L.scalers) {
            setError("batched evaluation: the trees do not share the data part and model shape");
            return 1;
        }
        TreeHdr &h = a.hdr[i];
        h.arena = arenaBase(Li);
        h.Pdeck = d->P + Li.pOff;
        h.tbl = d->tbl + Li.tblOff;
        h.aux = d->aux ? d->aux + Li.auxOff : nullptr;
        if (jobs[i].withLike) {
            anyLike = true;
            Node *root = t->root;
            if (!root || jobs[i].order->empty() || jobs[i].order->back() != root) { setError("fused evaluation: the last node must be the root"); return 1; }
            const int rc = root->compNums[p];
            if (rc < 0 || rc >= mp->nComps || !mp->comps[rc].val) { setError("root uses comp %d which does not exist", rc); return 1; }
            h.rootTips = root->isLeaf ? dp->dev.tips + (size_t)root->seqNum * L.ps : nullptr;
            h.pInvar = mp->pInvar;
            if (h.pInvar != 0.0 && !a.invarMask) { setError("pInvar is set but pf.setGlobalInvarSitesVec was not called on part %d", p); return 1; }
            for (int s = 0; s < 4; s++) h.pi[s] = mp->comps[rc].val[s];
            if (jobs[i].wantPatLikes) {
                if (d->patLikesCap < L.ps) {
                    if (d->patLikes) cudaFree(d->patLikes);
                    CUDA_TRY(cudaMalloc(&d->patLikes, sizeof(double) * L.ps));
                    d->patLikesCap = L.ps;
                }
                h.patLikes = d->patLikes;
            }
            if (blocks > d->maxLikeBlocks * 8) { setError("internal: partial buffer too small"); return 1; }
            h.partials = d->partials + (size_t)2 * d->maxLikeBlocks * 8 * p;
        }
    }
    const int K = L.nCat * 4;
    size_t smem = (size_t)2 * kMaxChildren * K * (L.W > 4 ? L.W : 4) * sizeof(double) + (size_t)K * THREADS * sizeof(double) * 2;
    if (smem > 100 * 1024) { setError("leaf tables too wide for the fused kernel"); return 1; }
    typedef void (*KernelFn)(const TreeArgs);
    static const KernelFn kFn4[9] = {cl_tree_dna_kernel<4, 128, 3, false>, cl_tree_dna_kernel<4, 64, 6, false>,
                                     cl_tree_dna_kernel<4, 32, 12, false>, cl_tree_dna_kernel<4, 128, 4, false>,
                                     cl_tree_dna_kernel<4, 256, 2, false>, cl_tree_dna_kernel<4, 64, 8, false>,
                                     cl_tree_dna_kernel<4, 32, 16, false>, cl_tree_dna_kernel_r<4, 64, 144>,
                                     cl_tree_dna_kernel_r<4, 32, 144>};
    static const KernelFn kFn1[9] = {cl_tree_dna_kernel<1, 128, 4, false>, cl_tree_dna_kernel<1, 64, 8, false>,
                                     cl_tree_dna_kernel<1, 32, 16, false>, cl_tree_dna_kernel<1, 128, 4, false>,
                                     cl_tree_dna_kernel<1, 256, 2, false>, cl_tree_dna_kernel<1, 64, 8, false>,
                                     cl_tree_dna_kernel<1, 32, 16, false>, cl_tree_dna_kernel<1, 64, 8, false>,
                                     cl_tree_dna_kernel<1, 32, 16, false>};
    static const KernelFn kFn4s[3] = {cl_tree_dna_kernel<4, 128, 3, true>, cl_tree_dna_kernel<4, 64, 6, true>,
                                      cl_tree_dna_kernel<4, 32, 12, true>};
    static const KernelFn kFn1s[3] = {cl_tree_dna_kernel<1, 128, 4, true>, cl_tree_dna_kernel<1, 64, 8, true>,
                                      cl_tree_dna_kernel<1, 32, 16, true>};
    KernelFn fn = L.nCat == 4 ? kFn4[variant] : kFn1[variant];
    if (L.scalers) {
        if (variant > 2) { setError("scalers need one of the default launch shapes"); return 1; }
        fn = L.nCat == 4 ? kFn4s[variant] : kFn1s[variant];
    }
    static bool attrSet = false;
    if (!attrSet) {
        for (int v = 0; v < 9; v++) {
            CUDA_TRY(cudaFuncSetAttribute(kFn4[v], cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            CUDA_TRY(cudaFuncSetAttribute(kFn1[v], cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        }
        for (int v = 0; v < 3; v++) {
            CUDA_TRY(cudaFuncSetAttribute(kFn4s[v], cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            CUDA_TRY(cudaFuncSetAttribute(kFn1s[v], cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        }
        attrSet = true;
    }
    auto launch = [&](int nTrees) -> int {
        a.nTrees = nTrees;
        size_t smemNow = smem;
        // staging buffers sized by the widest step actually present (2 or 3 in a binary tree), not by kMaxChildren
        int mk = 1;
        for (int i = 0; i < nTrees; i++)
            for (int k = 0; k < a.hdr[i].nSteps; k++) mk = a.steps[a.hdr[i].stepBase + k].nChildren > mk ? a.steps[a.hdr[i].stepBase + k].nChildren : mk;
        a.maxKids = mk;
        smemNow = (size_t)2 * mk * K * (L.W > 4 ? L.W : 4) * sizeof(double) + (size_t)K * THREADS * sizeof(double) * 2;
        // Residency of the small-CTA shapes is set on purpose, through the shared-memory request: 5 CTAs of 64
        // threads, 7 of 32.  A step costs a fixed latency whatever the occupancy, so what matters for a shard of
        // one to three waves is that the LAST wave is nearly full: 125 k patterns are 1.9 waves of 7 x 32 threads
        // per SM (0.86 ms) but 1.2 waves of 11 x 32 (0.93 ms); 250 k: 64 x 5 1.39 ms, 64 x 6 1.49 ms.
        static const int kResident[3] = {0, 5, 7};
        if (variant >= 1 && variant <= 2 && !getenv("P4B_FUSED_NOCAP")) {
            const size_t need = (size_t)233472 / (kResident[variant] + 1) - 1024 + 256;
            if (smemNow < need) smemNow = need;
        }
        fn<<<dim3(blocks, nTrees), THREADS, smemNow, G.stream>>>(a);
        CUDA_TRY(cudaGetLastError());
        snprintf(g_lastKernel, sizeof(g_lastKernel), "cl_tree_dna_kernel<%d,%d> shape %d", L.nCat, THREADS, variant);
        G.launches++;
        for (int i = 0; i < nJobs; i++) jobs[i].t->dev->lastCLLaunches++;
        return 0;
    };

    if (nJobs == 1) {
        // one tree: a step list longer than the argument block is cut into several launches
        size_t at = 0;
        for (;;) {
            bool more = jobs[0].storeAll;   // in: cutting allowed (not for lnL-only evaluations); out: steps remain
            const int ns = buildSteps(a, 0, kMaxSteps, jobs[0], p, &more, &at, maxKids);
            if (ns == -2) { setError("internal: lnL-only evaluation needs the whole tree in one launch"); return 1; }
            if (ns < 0) return 1;
            a.hdr[0].stepBase = 0;
            a.hdr[0].nSteps = ns;
            a.hdr[0].doLike = (jobs[0].withLike && !more) ? 1 : 0;
            if (ns > 0 || a.hdr[0].doLike)
                if (launch(1)) return 1;
            if (!more) break;
        }
    } else {
        int base = 0;
        for (int i = 0; i < nJobs; i++) {
            const int ns = buildSteps(a, base, kMaxSteps - base, jobs[i], p, nullptr, nullptr, maxKids);
            if (ns == -2) { setError("batched evaluation: the step lists of the trees do not fit one launch"); return 1; }
            if (ns < 0) return 1;
            a.hdr[i].stepBase = base;
            a.hdr[i].nSteps = ns;
            a.hdr[i].doLike = jobs[i].withLike ? 1 : 0;
            base += ns;
        }
        if (launch(nJobs)) return 1;
    }
    if (anyLike) {
        if (nJobs == 1) {
            like_final_kernel<<<1, 256, 0, G.stream>>>(a.hdr[0].partials, blocks, resultDev, nextMail());
        } else {
            FinalBatchArgs f;
            memset(&f, 0, sizeof(f));
            f.mail = nextMail();
            for (int i = 0; i < nJobs; i++) f.partials[i] = jobs[i].withLike ? a.hdr[i].partials : a.hdr[0].partials;
            f.result = resultDev;
            f.nBlocks = blocks;
            like_final_batch_kernel<<<nJobs, 256, 0, G.stream>>>(f);
        }
        CUDA_TRY(cudaGetLastError());
        G.launches++;
    }
    return 0;
}

static int launchFusedTree(Tree *t, int p, const std::vector<Node *> &order, bool withLike, bool wantPatLikes, bool storeAll = true, bool memo = false)
{
    FusedJob job = {t, &order, withLike, wantPatLikes, storeAll, memo};
    return launchFusedBatch(&job, 1, p, t->dev->result + 2 * p);
}

void setFusedEnabled(int on) { g_fusedEnabled = on != 0; }

// After an lnL-only evaluation some CLs of part p are not in memory: one storing whole-tree pass
// over the nodes of the evaluation order makes them current again.
static int residentPass(Tree *t, int p)
{
    bool all = true;
    for (Node *n : t->nodes)
        if (n && n->clSlot[p] >= 0 && !n->clResident[p]) { all = false; break; }
    if (all) return 0;
    std::vector<Node *> order;
    for (int j = 0; j < t->nNodes; j++) {
        const int i = t->postOrder[j];
        if (i == P4B_NO_ORDER) continue;
        Node *n = t->nodes[i];
        if (n && (!n->isLeaf || n == t->root)) order.push_back(n);
    }
    if (!fusedEligible(t->dev->parts[p])) { setError("internal: non-resident CLs on a part without the whole-tree kernel"); return 1; }
    const int keep = t->dev->lastCLLaunches;
    if (launchFusedTree(t, p, order, false, false, true)) return 1;
    t->dev->lastCLLaunches = keep;
    for (Node *n : t->nodes)
        if (n && n->clSlot[p] >= 0) n->clResident[p] = 1;
    return 0;
}

// Bring part p of the tree to the state the reference would be in now: every CL in memory and every
// queued node-level CL call executed.
int treeEnsureResident(Tree *t, int p)
{
    if (residentPass(t, p)) return 1;
    std::vector<Node *> &q = t->dev->pending[p];
    if (q.empty()) return 0;
    std::vector<Node *> order;
    order.swap(q);
    return launchFusedTree(t, p, order, false, false, true, true);
}

static int treeFlushAllPending(Tree *t)
{
    if (!t->dev) return 0;
    for (int p = 0; p < t->nParts; p++)
        if (!t->dev->pending[p].empty())
            if (treeEnsureResident(t, p)) return 1;
    return 0;
}
int treeFlushPending(Tree *t) { return treeFlushAllPending(t); }
bool treeHasPending(Tree *t)
{
    if (!t || !t->dev) return false;
    for (auto &q : t->dev->pending)
        if (!q.empty()) return true;
    return false;
}

// ---------------------------------------------------------------------------
// Log-likelihood
// ---------------------------------------------------------------------------
// like_kernel + fold on the root CL as it is in memory now (stream order).
static int enqueueRootLike(Tree *t, int p, bool wantPatLikes, double *resultDev)
{
    TreeDevice *d = t->dev;
    PartLayout &L = d->parts[p];
    Part *dp = t->data->parts[p];
    ModelPart *mp = t->model->parts[p];
    Node *root = t->root;
    if (!root) { setError("tree has no root"); return 1; }
    if (root->clSlot[p] < 0) { setError("the root has no conditional likelihoods"); return 1; }
    const int rc = root->compNums[p];
    if (rc < 0 || rc >= mp->nComps || !mp->comps[rc].val) { setError("root uses comp %d which does not exist", rc); return 1; }
    LikeArgs a;
    memset(&a, 0, sizeof(a));
    a.cl = nodeCL(root, p);
    a.counts = dp->dev.counts;
    a.invarMask = dp->dev.invarMask;
    a.rootTips = root->isLeaf ? dp->dev.tips + (size_t)root->seqNum * L.ps : nullptr;
    a.eqMask = d->eqMasks + L.eqOff;
    a.rootScale = nodeSC(root, p);
    a.ps = L.ps;
    a.nPat = L.nPat;
    a.dim = L.dim;
    a.nCat = L.nCat;
    a.pInvar = mp->pInvar;
    if (a.pInvar != 0.0 && !a.invarMask) { setError("pInvar is set but pf.setGlobalInvarSitesVec was not called on part %d", p); return 1; }
    for (int s = 0; s < L.dim; s++) a.pi[s] = mp->comps[rc].val[s];
    if (wantPatLikes) {
        if (d->patLikesCap < L.ps) {
            if (d->patLikes) cudaFree(d->patLikes);
            CUDA_TRY(cudaMalloc(&d->patLikes, sizeof(double) * L.ps));
            d->patLikesCap = L.ps;
        }
        a.patLikes = d->patLikes;
    }
    const int blocks = (L.ps + 255) / 256;
    a.partials = d->partials + (size_t)2 * d->maxLikeBlocks * 8 * p;
    like_kernel<<<blocks, 256, 0, G.stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    like_final_kernel<<<1, 256, 0, G.stream>>>(a.partials, blocks, resultDev, nextMail());
    CUDA_TRY(cudaGetLastError());
    G.launches += 2;
    return 0;
}

static int enqueuePartLike(Tree *t, int p, bool wantPatLikes)
{
    if (!t->root) { setError("tree has no root"); return 1; }
    if (t->root->clSlot[p] < 0) { setError("the root has no conditional likelihoods"); return 1; }
    if (treeEnsureResident(t, p)) return 1;
    return enqueueRootLike(t, p, wantPatLikes, t->dev->result + 2 * p);
}

static int fetchResults(Tree *t, int p0, int p1)
{
    TreeDevice *d = t->dev;
    const int n = 2 * (p1 - p0);
    if (allReduceIfNeeded(d->result + 2 * p0, n)) return 1;
    CUDA_TRY(cudaMemcpyAsync(d->hResult + 2 * p0, d->result + 2 * p0, n * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
    return streamSync();
}

static int fillSiteLikes(Tree *t, int p)
{
    // Pf/p4_tree.c:1015-1021: expand pattern likelihoods to sites.
    TreeDevice *d = t->dev;
    PartLayout &L = d->parts[p];
    Part *dp = t->data->parts[p];
    if (G.world > 1) { setError("getSiteLikes is not available when patterns are sharded over several processes"); return 1; }
    std::vector<double> pl(L.ps);
    CUDA_TRY(cudaMemcpyAsync(pl.data(), d->patLikes, sizeof(double) * L.ps, cudaMemcpyDeviceToHost, G.stream));
    if (streamSync()) return 1;
    dp->siteLikes.assign(dp->nChar, 0.0);
    for (int i = 0; i < dp->nChar; i++) dp->siteLikes[i] = pl[dp->sequencePositionPatternIndex[i]];
    return 0;
}

// Launch everything the log-likelihood of part p needs, results into d->result[2p..2p+1]; no synchronisation.
static int launchPartLike(Tree *t, int p, int getSiteLikes)
{
    std::vector<Node *> &q = t->dev->pending[p];
    if (!q.empty() && q.back() == t->root && fusedEligible(t->dev->parts[p])) {
        // the usual end of a proposal: the queued dirty path ends at the root, so the path and the
        // root reduction are one launch
        if (residentPass(t, p)) return 1;
        std::vector<Node *> order;
        order.swap(q);
        return launchFusedTree(t, p, order, true, getSiteLikes != 0, true, true);
    }
    return enqueuePartLike(t, p, getSiteLikes != 0);
}

// Wait for an evaluation started by p4b_partLogLikeBegin (its result is already on its way to the pinned
// host buffer) -- on its own event, not on the stream, so work queued behind it keeps running.
static int collectBegun(Tree *t, int p, double *lnL)
{
    TreeDevice *d = t->dev;
    CUDA_TRY(cudaEventSynchronize(d->evLike[p]));
    d->likeBegun[p] = 0;
    double v = d->hResult[2 * p];
    if (d->hResult[2 * p + 1] > 0.0) v = P4B_BAD_LIKE;
    t->partLikes[p] = v;
    *lnL = v;
    return 0;
}

double treePartLogLike(Tree *t, Part *dpArg, int p, int getSiteLikes)
{
    if (!t->dev) { setError("tree has no device state"); return NAN; }
    if (ensureFresh(t, true)) return NAN;
    if (p < 0 || p >= t->nParts) { setError("p4_partLogLike: bad part %d", p); return NAN; }
    (void)dpArg;   // the reference passes the part explicitly; it is data->parts[pNum]
    if (t->dev->likeBegun[p] && !getSiteLikes && t->dev->pending[p].empty()) {
        double v = NAN;
        if (collectBegun(t, p, &v)) return NAN;
        return v;
    }
    t->dev->likeBegun[p] = 0;
    if (launchPartLike(t, p, getSiteLikes)) return NAN;
    if (getSiteLikes && fillSiteLikes(t, p)) return NAN;
    if (fetchResults(t, p, p + 1)) return NAN;
    double lnL = t->dev->hResult[2 * p];
    if (t->dev->hResult[2 * p + 1] > 0.0) lnL = P4B_BAD_LIKE;
    t->partLikes[p] = lnL;
    return lnL;
}

// p4b_partLogLikeBegin: start the evaluation of part p (queued P(t) jobs, queued CL calls, root reduction,
// all-reduce, device->host copy) and return at once; the value is collected by p4b_partLogLike /
// p4b_treesPartLogLike, which wait on this evaluation's own event.  Lets the host prepare the next chain's
// proposal, and finish the previous chain's generation, while the GPU evaluates this one.
int treePartLogLikeBegin(Tree *t, int p)
{
    if (!t->dev) { setError("tree has no device state"); return 1; }
    if (ensureFresh(t, true)) return 1;
    if (p < 0 || p >= t->nParts) { setError("p4b_partLogLikeBegin: bad part %d", p); return 1; }
    TreeDevice *d = t->dev;
    if (launchPartLike(t, p, 0)) return 1;
    if (allReduceIfNeeded(d->result + 2 * p, 2)) return 1;
    CUDA_TRY(cudaMemcpyAsync(d->hResult + 2 * p, d->result + 2 * p, 2 * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
    if (!d->evLike[p]) CUDA_TRY(cudaEventCreateWithFlags(&d->evLike[p], cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(d->evLike[p], G.stream));
    d->likeBegun[p] = 1;
    return 0;
}

double treeLogLike(Tree *t, int getSiteLikes)
{
    if (!t->dev) { setError("tree has no device state"); return NAN; }
    if (ensureFresh(t, true)) return NAN;
    TreeDevice *d = t->dev;
    if (treeFlushAllPending(t)) return NAN;
    d->lastCLLaunches = 0;
    // nodes whose CL is recomputed, in the caller's post-order (Pf/p4_tree.c:875-887)
    std::vector<Node *> order;
    for (int j = 0; j < t->nNodes; j++) {
        const int i = t->postOrder[j];
        if (i == P4B_NO_ORDER) continue;
        if (i < 0 || i >= (int)t->nodes.size() || !t->nodes[i]) { setError("postOrder[%d] = %d is not a node", j, i); return NAN; }
        Node *n = t->nodes[i];
        if (!n->isLeaf || n == t->root) order.push_back(n);
    }
    if (order.empty() || order.back() != t->root) { setError("p4_treeLogLike: postOrder does not end at the root"); return NAN; }
    cudaEventRecord(d->evCLa, G.stream);
    std::vector<char> likeDone(t->nParts, 0);
    {   // the 20-state parts of a partitioned alignment go through ONE launch of the whole-tree kernel: no launch gaps, and one
        // last, partly filled wave of CTAs instead of one per part
        std::vector<int> aaParts;
        for (int p = 0; p < t->nParts; p++)
            if (d->parts[p].dim == 20 && fusedEligible(d->parts[p])) aaParts.push_back(p);
        if (!getSiteLikes && aaParts.size() >= 2 && aaParts.size() <= (size_t)kMaxBatchTrees && order.size() + 8 <= (size_t)kMaxSteps) {      // (site likelihoods: one part at a time, they share a buffer)
            const bool storeAll = t->storeCL != 0;
            std::vector<FusedJob> jobs(aaParts.size(), FusedJob{t, &order, true, false, storeAll, false});
            std::vector<double *> res;
            for (int p : aaParts) res.push_back(d->result + 2 * p);
            if (launchFusedAABatch(jobs.data(), (int)jobs.size(), aaParts[0], nullptr, aaParts.data(), res.data())) return NAN;
            for (int p : aaParts) likeDone[p] = 1;
        }
    }
    for (int p = 0; p < t->nParts; p++) {
        if (likeDone[p]) continue;
        if (fusedEligible(d->parts[p])) {
            const bool storeAll = t->storeCL != 0 || (d->parts[p].dim != 4 && d->parts[p].dim != 20) || order.size() + 8 > (size_t)kMaxSteps;
            if (launchFusedTree(t, p, order, true, getSiteLikes != 0, storeAll)) return NAN;
            likeDone[p] = 1;
            if (getSiteLikes && fillSiteLikes(t, p)) return NAN;
        } else {
            for (Node *n : order)
                if (nodeSetCLImpl(n, p, false)) return NAN;   // a whole-tree evaluation recomputes every node, as the reference does
        }
    }
    cudaEventRecord(d->evCLb, G.stream);
    d->clTimed = true;
    for (int p = 0; p < t->nParts; p++) {
        if (likeDone[p]) continue;
        if (enqueuePartLike(t, p, getSiteLikes != 0)) return NAN;
        if (getSiteLikes && fillSiteLikes(t, p)) return NAN;
    }
    if (fetchResults(t, 0, t->nParts)) return NAN;
    double lnL = 0.0;
    for (int p = 0; p < t->nParts; p++) {
        double v = d->hResult[2 * p];
        if (d->hResult[2 * p + 1] > 0.0) v = P4B_BAD_LIKE;
        t->partLikes[p] = v;
        lnL += v;
    }
    t->logLike = lnL;
    return lnL;
}

// p4_partLogLike for several trees at once (the prop trees of Metropolis-coupled chains after their
// proposals): when every tree's queued node-level calls end at its root, all dirty paths and all root
// reductions run as one launch, followed by one fold, one all-reduce and one device->host copy.
int treesPartLogLike(Tree **trees, int n, int p, double *out)
{
    if (n <= 0) return 0;
    for (int i = 0; i < n; i++)
        if (trees[i] && trees[i]->dev && ensureFresh(trees[i], true)) return 1;
    {   // evaluations already started by p4b_partLogLikeBegin: their results are on the way, wait for each
        bool allBegun = n <= kMaxBatchTrees;
        for (int i = 0; i < n && allBegun; i++) {
            Tree *t = trees[i];
            if (!t || !t->dev || p < 0 || p >= t->nParts) { setError("p4b_treesPartLogLike: bad tree or part"); return 1; }
            if (!t->dev->likeBegun[p] || !t->dev->pending[p].empty()) allBegun = false;
        }
        if (allBegun) {
            for (int i = 0; i < n; i++)
                if (collectBegun(trees[i], p, &out[i])) return 1;
            return 0;
        }
    }
    bool batchable = true;
    for (int i = 0; i < n && batchable; i++) {
        Tree *t = trees[i];
        if (!t || !t->dev || p < 0 || p >= t->nParts) { setError("p4b_treesPartLogLike: bad tree or part"); return 1; }
        const std::vector<Node *> &q = t->dev->pending[p];
        if (q.empty() || q.back() != t->root || !fusedEligible(t->dev->parts[p]) || t->data->parts[p] != trees[0]->data->parts[p] || t->dev->likeBegun[p]) batchable = false;
        for (int j = 0; j < i && batchable; j++)
            if (trees[j] == t) batchable = false;
    }
    if (!batchable) {
        for (int i = 0; i < n; i++) {
            out[i] = treePartLogLike(trees[i], nullptr, p, 0);
            if (out[i] != out[i]) return 1;
        }
        return 0;
    }
    const int stepKids = trees[0]->dev->parts[p].dim != 4 ? kAAKids : 2;     // every whole-tree kernel takes at most two children per step
    int i0 = 0;
    while (i0 < n) {
        // greedy group: at most kMaxBatchTrees trees and kMaxSteps steps
        int i1 = i0, steps = 0;
        while (i1 < n && i1 - i0 < kMaxBatchTrees) {
            int need = 0;
            for (Node *nd : trees[i1]->dev->pending[p]) {
                int kids = 0;
                for (Node *c = nd->leftChild; c; c = c->sibling) kids++;
                need += (kids + stepKids - 1) / stepKids;     // the kernel's own step width: an upper bound on what buildSteps writes
            }
            if (i1 > i0 && steps + need > kMaxSteps) break;
            steps += need;
            i1++;
        }
        const int m = i1 - i0;
        if (m == 1) {
            out[i0] = treePartLogLike(trees[i0], nullptr, p, 0);
            if (out[i0] != out[i0]) return 1;
            i0 = i1;
            continue;
        }
        std::vector<std::vector<Node *>> orders(m);
        std::vector<FusedJob> jobs(m);
        for (int k = 0; k < m; k++) {
            Tree *t = trees[i0 + k];
            if (residentPass(t, p)) return 1;
            orders[k].swap(t->dev->pending[p]);
            jobs[k] = FusedJob{t, &orders[k], true, false, true, true};
        }
        if (launchFusedBatch(jobs.data(), m, p, G.dBatch)) {
            // nothing was launched: the calls go back into their queues, and no node of them may pass for
            // computed (buildSteps stamps nodes as it writes their steps)
            for (int k = 0; k < m; k++) {
                Tree *t = trees[i0 + k];
                for (Node *nd : orders[k]) {
                    nd->clStamp[p] = ++G.stamp;
                    nd->clKey[p].clear();
                    nd->clNeedsUpdating = 1;
                }
                std::vector<Node *> &q = t->dev->pending[p];
                q.insert(q.begin(), orders[k].begin(), orders[k].end());
            }
            return 1;
        }
        if (allReduceIfNeeded(G.dBatch, 2 * m)) return 1;
        CUDA_TRY(cudaMemcpyAsync(G.hBatch, G.dBatch, 2 * m * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
        if (streamSync()) return 1;
        for (int k = 0; k < m; k++) {
            double v = G.hBatch[2 * k];
            if (G.hBatch[2 * k + 1] > 0.0) v = P4B_BAD_LIKE;
            trees[i0 + k]->partLikes[p] = v;
            out[i0 + k] = v;
        }
        i0 = i1;
    }
    return 0;
}

// ---------------------------------------------------------------------------
// cur/prop state transfer
// ---------------------------------------------------------------------------
static int checkTwins(Tree *a, Tree *b)
{
    if (!a->dev || !b->dev) { setError("tree has no device state"); return 1; }
    if (a->nNodes != b->nNodes || a->nParts != b->nParts) { setError("the two trees differ in node or part count"); return 1; }
    if (ensureFresh(a, false) || ensureFresh(b, false)) return 1;   // data parts re-compressed since: lay both out again first
    for (int p = 0; p < a->nParts; p++) {
        const PartLayout &A = a->dev->parts[p], &B = b->dev->parts[p];
        if (A.clNodeDoubles != B.clNodeDoubles || A.pDoubles != B.pDoubles || A.W != B.W || A.scalers != B.scalers || A.ps != B.ps) { setError("the two trees differ in part %d layout", p); return 1; }
    }
    return 0;
}

static bool g_shareCL = true;
void setShareEnabled(int on) { g_shareCL = on != 0; }

// Can nodes of b reference CL buffers that nodes of a hold (part p)?  Arenas pair up exclusively.
static bool canShare(Tree *a, Tree *b, int p)
{
    if (!g_shareCL || a == b) return false;
    PartLayout &LA = a->dev->parts[p], &LB = b->dev->parts[p];
    Arena *A = LA.own.get(), *B = LB.own.get();
    if (A->slotDoubles != B->slotDoubles || A->ps != B->ps || A->scalers != B->scalers) return false;
    if ((A->partner && A->partner != B) || (B->partner && B->partner != A)) return false;
    {   // both arenas must be addressable from one base with 30 bits of 256-byte units
        const char *lo = (const char *)(A->base < B->base ? A->base : B->base);
        const char *hiA = (const char *)(A->base + A->slotDoubles * A->nSlots), *hiB = (const char *)(B->base + B->slotDoubles * B->nSlots);
        const char *hi = hiA > hiB ? hiA : hiB;
        if ((size_t)(hi - lo) >= ((size_t)1 << 30) * 256) return false;
    }
    if (!A->partner) {
        A->partner = B;
        B->partner = A;
        LA.twin = LB.own;
        LB.twin = LA.own;
        a->dev->slotStamp++;
        b->dev->slotStamp++;
    }
    return LA.twin.get() == B && LB.twin.get() == A;
}

int treeCopyCondLikes(Tree *a, Tree *b, int doAll)
{
    if (checkTwins(a, b)) return 1;
    if (treeFlushAllPending(b)) return 1;   // b's queued calls were issued before this copy
    for (int p = 0; p < a->nParts; p++)
        if (treeEnsureResident(a, p)) return 1;
    std::vector<char> share(a->nParts);
    for (int p = 0; p < a->nParts; p++) share[p] = canShare(a, b, p) ? 1 : 0;
    for (int j = 0; j < a->nNodes; j++) {
        const int i = a->preOrder[j];
        if (i == P4B_NO_ORDER) continue;
        Node *nA = a->nodes[i], *nB = b->nodes[i];
        if (!nA || !nB || nA->isLeaf) continue;
        if (!doAll && !(nA->clNeedsUpdating || nB->clNeedsUpdating)) continue;
        for (int p = 0; p < a->nParts; p++) {
            if (nA->clSlot[p] < 0) continue;
            // Content that is already identical (same computation id) is not moved again.
            if (nA->clStamp[p] == nB->clStamp[p] && nB->clResident[p] && nB->clSlot[p] >= 0) continue;
            if (share[p]) {
                // b's node now references a's buffer; b's old slot goes back to its arena
                PartLayout &LA = a->dev->parts[p], &LB = b->dev->parts[p];
                Arena *src = LA.arena(nA->clSel[p]);
                const int slot = nA->clSlot[p];
                if (nB->clSlot[p] >= 0) slotRelease(LB.arena(nB->clSel[p]), nB->clSlot[p]);
                src->refs[slot]++;
                nB->clSel[p] = (char)(src == LB.own.get() ? 0 : 1);
                nB->clSlot[p] = slot;
                b->dev->slotStamp++;
            } else {
                if (nodeMakeWritable(nB, p)) return 1;
                CUDA_TRY(cudaMemcpyAsync(nodeCL(nB, p), nodeCL(nA, p), a->dev->parts[p].own->slotDoubles * sizeof(double),
                                         cudaMemcpyDeviceToDevice, G.stream));   // CL and, behind it, the exponents
            }
            nB->clStamp[p] = nA->clStamp[p];
            nB->clKey[p] = nA->clKey[p];
            nB->clResident[p] = 1;
        }
        if (!doAll) nA->clNeedsUpdating = nB->clNeedsUpdating = 0;
    }
    return 0;
}

int treeCopyBigPDecks(Tree *a, Tree *b, int doAll)
{
    if (!doAll) { setError("p4_copyBigPDecks() doAll is not set. Programming error?"); return 1; }
    if (checkTwins(a, b)) return 1;
    if (treeFlushAllPending(a) || treeFlushAllPending(b) || flushPJobs()) return 1;
    std::vector<std::pair<Node *, Node *>> todo;
    for (int j = 0; j < a->nNodes; j++) {
        const int i = a->preOrder[j];
        if (i == P4B_NO_ORDER) continue;
        Node *nA = a->nodes[i], *nB = b->nodes[i];
        if (!nA || !nB || nA == a->root) continue;
        for (int p = 0; p < a->nParts; p++)
            if (nA->pStamp[p] != nB->pStamp[p]) { todo.emplace_back(nA, nB); break; }
    }
    TreeDevice *da = a->dev, *db = b->dev;
    const bool sameShape = da->pNodeDoubles == db->pNodeDoubles && da->tblNodeDoubles == db->tblNodeDoubles &&
                           da->auxNodeDoubles == db->auxNodeDoubles && a->nNodes == b->nNodes;
    if (todo.size() > 4 && sameShape) {
        // the decks are small (cfg 2: 100 KB of P, 380 KB of leaf tables per tree): two copies of
        // everything beat hundreds of per-node copies on launch overhead alone
        CUDA_TRY(cudaMemcpyAsync(db->P, da->P, da->pNodeDoubles * sizeof(double) * (size_t)a->nNodes, cudaMemcpyDeviceToDevice, G.stream));
        if (da->tblNodeDoubles)
            CUDA_TRY(cudaMemcpyAsync(db->tbl, da->tbl, da->tblNodeDoubles * sizeof(double) * (size_t)a->nNodes, cudaMemcpyDeviceToDevice, G.stream));
        if (da->auxNodeDoubles)
            CUDA_TRY(cudaMemcpyAsync(db->aux, da->aux, da->auxNodeDoubles * sizeof(double) * (size_t)a->nNodes, cudaMemcpyDeviceToDevice, G.stream));
        for (int i = 0; i < a->nNodes; i++)
            if (a->nodes[i] && b->nodes[i])
                for (int p = 0; p < a->nParts; p++) {
                    b->nodes[i]->pStamp[p] = a->nodes[i]->pStamp[p];
                    b->nodes[i]->pKey[p] = a->nodes[i]->pKey[p];
                }
        return 0;
    }
    for (auto &ab : todo) {
        Node *nA = ab.first, *nB = ab.second;
        for (int p = 0; p < a->nParts; p++) {
            if (nA->pStamp[p] == nB->pStamp[p]) continue;
            CUDA_TRY(cudaMemcpyAsync(nodeP(nB, p), nodeP(nA, p), a->dev->parts[p].pDoubles * sizeof(double), cudaMemcpyDeviceToDevice, G.stream));
            if (a->dev->parts[p].tblDoubles)
                CUDA_TRY(cudaMemcpyAsync(nodeTbl(nB, p), nodeTbl(nA, p), a->dev->parts[p].tblDoubles * sizeof(double), cudaMemcpyDeviceToDevice, G.stream));
            if (a->dev->parts[p].auxDoubles && sameShape)
                CUDA_TRY(cudaMemcpyAsync(nodeAux(nB, p), nodeAux(nA, p), a->dev->parts[p].auxDoubles * sizeof(double), cudaMemcpyDeviceToDevice, G.stream));
            nB->pStamp[p] = nA->pStamp[p];
            nB->pKey[p] = nA->pKey[p];
        }
    }
    return 0;
}

int treeVerifyDevice(Tree *a, Tree *b)
{
    if (checkTwins(a, b)) return -1;
    if (flushPJobs()) return -1;
    for (int p = 0; p < a->nParts; p++)
        if (treeEnsureResident(a, p) || treeEnsureResident(b, p)) return -1;
    TreeDevice *d = a->dev;
    if (cudaMemsetAsync(d->flag, 0, sizeof(int), G.stream) != cudaSuccess) { setError("memset failed"); return -1; }
    int result = 0;
    for (int pass = 0; pass < 2; pass++) {   // 0: CL, 1: P decks
        for (int j = 0; j < a->nNodes; j++) {
            const int i = a->preOrder[j];
            if (i == P4B_NO_ORDER) continue;
            Node *nA = a->nodes[i], *nB = b->nodes[i];
            if (!nA || !nB) continue;
            if (pass == 0 && nA->isLeaf) continue;
            if (pass == 1 && nA == a->root) continue;
            for (int p = 0; p < a->nParts; p++) {
                const double *x, *y;
                size_t n;
                if (pass == 0) {
                    if (nA->clSlot[p] < 0 || nB->clSlot[p] < 0) { if (nA->clSlot[p] != nB->clSlot[p]) result = 1; continue; }
                    x = nodeCL(nA, p); y = nodeCL(nB, p); n = d->parts[p].clNodeDoubles;
                    if (x == y) continue;   // one buffer referenced by both trees
                } else {
                    x = nodeP(nA, p); y = nodeP(nB, p); n = d->parts[p].pDoubles;
                }
                int blocks = (int)((n + 255) / 256);
                if (blocks > G.numSMs * 8) blocks = G.numSMs * 8;
                diff_kernel<<<blocks, 256, 0, G.stream>>>(x, y, n, 1.e-15, d->flag);
                G.launches++;
            }
        }
        int h = 0;
        if (cudaMemcpyAsync(&h, d->flag, sizeof(int), cudaMemcpyDeviceToHost, G.stream) != cudaSuccess || streamSync()) { setError("verify: copy failed"); return -1; }
        if (h) {
            printf(pass == 0 ? "Verify: cond likes are different.  Bad.\n" : "Verify: bigPDecks are different.  Bad.\n");
            result = 1;
            cudaMemsetAsync(d->flag, 0, sizeof(int), G.stream);
        }
    }
    return result;
}

// ---------------------------------------------------------------------------
// Inspection / timing
// ---------------------------------------------------------------------------
int nodeGetCL(Node *n, int p, double *out)
{
    Tree *t = n->tree;
    if (!t->dev || p < 0 || p >= t->nParts) { setError("p4b_getNodeCL: bad part"); return 1; }
    if (n->clSlot[p] < 0) { setError("node %d has no conditional likelihoods", n->nodeNum); return 1; }
    if (treeEnsureResident(t, p)) return 1;
    PartLayout &L = t->dev->parts[p];
    std::vector<double> h(L.clNodeDoubles);
    CUDA_TRY(cudaMemcpyAsync(h.data(), nodeCL(n, p), L.clNodeDoubles * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
    if (streamSync()) return 1;
    for (int k = 0; k < L.nCat * L.dim; k++) memcpy(out + (size_t)k * L.nPat, h.data() + (size_t)k * L.ps, sizeof(double) * L.nPat);
    return 0;
}

int nodeGetBigP(Node *n, int p, double *out)
{
    Tree *t = n->tree;
    if (!t->dev || p < 0 || p >= t->nParts) { setError("p4b_getNodeBigP: bad part"); return 1; }
    if (flushPJobs()) return 1;
    CUDA_TRY(cudaMemcpyAsync(out, nodeP(n, p), t->dev->parts[p].pDoubles * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
    return streamSync();
}

// Test hook: overwrite one node's P deck (and its leaf lookup table) with given
// values, so that the CL kernels can be checked against the reference with the
// transition matrices held identical.
int nodeSetBigP(Node *n, int p, const double *in)
{
    Tree *t = n->tree;
    if (!t->dev || p < 0 || p >= t->nParts) { setError("p4b_setNodeBigP: bad part"); return 1; }
    if (treeFlushAllPending(t) || flushPJobs()) return 1;
    PartLayout &L = t->dev->parts[p];
    Part *dp = t->data->parts[p];
    CUDA_TRY(cudaMemcpyAsync(nodeP(n, p), in, L.pDoubles * sizeof(double), cudaMemcpyHostToDevice, G.stream));
    std::vector<double> T(L.tblDoubles);
    const int dim = L.dim, W = L.W;
    std::vector<uint64_t> em(dp->nRealEquates > 0 ? dp->nRealEquates : 1, 0);
    for (int e = 0; e < dp->nEquates; e++) {
        const int j = dp->equateColumn(e);
        if (j < 0) continue;
        for (int s = 0; s < dim; s++)
            if (dp->equates[(size_t)e * dim + s]) em[j] |= 1ull << s;
    }
    for (int k = 0; k < L.nCat * dim; k++)
        for (int w = 0; w < W; w++) {
            double v;
            if (w < dim) v = in[(size_t)k * dim + w];
            else if (w == dim) v = 1.0;
            else {
                v = 0.0;
                for (int x = 0; x < dim; x++)
                    if ((em[w - dim - 1] >> x) & 1ull) v += in[(size_t)k * dim + x];
            }
            T[(size_t)k * W + w] = v;
        }
    CUDA_TRY(cudaMemcpyAsync(nodeTbl(n, p), T.data(), T.size() * sizeof(double), cudaMemcpyHostToDevice, G.stream));
    std::vector<double> A(L.auxDoubles);
    if (L.auxDoubles && L.auxDP > 0) {   // the generic tensor-core kernel's decks, as pmatrix_kernel derives them from P
        const int DP = L.auxDP, NT = DP / 8, KS = DP / 4, F = KS * NT * 32;
        for (int i = 0; i < L.nCat * F; i++) {
            const int l = i & 31, nt = (i >> 5) % NT, kk = (i / (32 * NT)) % KS, ct = i / F;
            const int s = 8 * nt + (l >> 2), x = 8 * (kk >> 1) + 2 * (l & 3) + (kk & 1);
            A[i] = (s < dim && x < dim) ? in[((size_t)ct * dim + s) * dim + x] : 0.0;
        }
        const size_t nF = (size_t)L.nCat * F;
        for (int i = 0; i < L.nCat * W * DP; i++) {
            const int st = i % DP, w = (i / DP) % W, ct = i / (DP * W);
            A[nF + i] = st < dim ? T[((size_t)ct * dim + st) * W + w] : 0.0;
        }
        CUDA_TRY(cudaMemcpyAsync(nodeAux(n, p), A.data(), A.size() * sizeof(double), cudaMemcpyHostToDevice, G.stream));
    } else if (L.auxDoubles) {   // the same two decks pmatrix_kernel derives from P
        const int nF = L.nCat * kAAFrag;
        for (int i = 0; i < nF; i++) {
            const int l = i & 31, nt = (i >> 5) % 3, kk = (i / 96) % 5, ct = i / kAAFrag;
            const int s = 8 * nt + (l >> 2), x = kk < 4 ? 8 * (kk >> 1) + 2 * (l & 3) + (kk & 1) : 16 + (l & 3);
            A[i] = (s < dim && x < dim) ? in[((size_t)ct * dim + s) * dim + x] : 0.0;
        }
        for (int i = 0; i < L.nCat * W * kAATblStates; i++) {
            const int st = i % kAATblStates, w = (i / kAATblStates) % W, ct = i / (kAATblStates * W);
            A[nF + i] = st < dim ? T[((size_t)ct * dim + st) * W + w] : 0.0;
        }
        CUDA_TRY(cudaMemcpyAsync(nodeAux(n, p), A.data(), A.size() * sizeof(double), cudaMemcpyHostToDevice, G.stream));
    }
    if (streamSync()) return 1;
    n->pStamp[p] = ++G.stamp;
    n->pKey[p].clear();
    return 0;
}

int treeShardRangeOf(Tree *t, int p, int *lo, int *hi)
{
    if (!t->data || p < 0 || p >= t->nParts) { setError("bad part"); return 1; }
    *lo = t->data->parts[p]->dev.lo;
    *hi = t->data->parts[p]->dev.hi;
    return 0;
}

int treeSync(Tree *t)
{
    if (!G.ready) return 0;
    if (t && treeFlushAllPending(t)) return 1;
    if (flushPJobs()) return 1;
    return streamSync();
}

int treeTimerBegin(Tree *t)
{
    if (!t->dev) { setError("tree has no device state"); return 1; }
    CUDA_TRY(cudaEventRecord(t->dev->evA, G.stream));
    return 0;
}

double treeTimerEnd(Tree *t)
{
    if (!t->dev) { setError("tree has no device state"); return -1.0; }
    if (cudaEventRecord(t->dev->evB, G.stream) != cudaSuccess || cudaEventSynchronize(t->dev->evB) != cudaSuccess) { setError("timer failed"); return -1.0; }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, t->dev->evA, t->dev->evB) != cudaSuccess) { setError("timer failed"); return -1.0; }
    return (double)ms;
}

int treeLastCLTiming(Tree *t, double *ms, int *nLaunches)
{
    if (!t->dev || !t->dev->clTimed) { setError("no p4b_treeLogLike has run on this tree"); return 1; }
    CUDA_TRY(cudaEventSynchronize(t->dev->evCLb));
    float f = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&f, t->dev->evCLa, t->dev->evCLb));
    if (ms) *ms = (double)f;
    if (nLaunches) *nLaunches = t->dev->lastCLLaunches;
    return 0;
}

long long treeDeviceBytes(Tree *t) { return t->dev ? t->dev->bytes : 0; }

int treeFlushL2(Tree *t)
{
    (void)t;
    if (engineInit()) return 1;
    if (!G.flushBuf) {
        G.flushN = (size_t)(256u << 20) / sizeof(double);   // 256 MiB > 126 MB L2
        CUDA_TRY(cudaMalloc(&G.flushBuf, G.flushN * sizeof(double)));
    }
    flush_kernel<<<G.numSMs * 4, 256, 0, G.stream>>>(G.flushBuf, G.flushN, 1.0);
    CUDA_TRY(cudaGetLastError());
    return 0;
}


// ---------------------------------------------------------------------------
// Newton-Raphson on the branch lengths (Pf/p4_treeNewt.c)
//
// The reference keeps, beside every node's CL (everything below the node), a second array cl2
// (everything on the other side of the node's branch), so that the likelihood and its first two
// derivatives in ONE branch length are a single pass over two arrays -- no walk to the root.  It goes
// round the tree in post-order; thanks to that order every cl2 and every CL is computed once per round
// (Pf/p4_treeNewt.c:81-172).  Here cl2 arrays live in an arena of their own, are computed by the per-node
// CL kernels (the parent's contribution is a child with a transposed deck, or a table of pi at the root),
// and each Newton iteration is newt_deck_kernel + newt_kernel + fold + one 24-byte read-back per part.
// ---------------------------------------------------------------------------
struct NewtState {
    int nSlots = 0;                         // one per non-leaf node; the leaves share the root's (the root has no cl2)
    std::vector<int> slotOf;                // per node number
    std::vector<double *> cl2;              // per part: [nSlots][clNodeDoubles]
    double *PT = nullptr;                   // [pNodeDoubles]: transposed deck of the parent whose child's cl2 is being set
    double *rootTbl = nullptr;              // [tblNodeDoubles]: per part T[k][w] = pi_root[state of k] in every column
    double *decks = nullptr;                // per part, at deckOff[p]: newt_deck_kernel's output
    std::vector<size_t> deckOff;
    double *partials = nullptr;             // per part, at 3*maxBlocks*p
    int maxBlocks = 0;
    double *result = nullptr, *hResult = nullptr;   // [3*nParts]
    unsigned *ticket = nullptr;             // newt_dna_kernel: which CTA finished last (returns to 0 after every launch)
    std::vector<char> cl2NeedsUpdating;
    long long iters = 0;                    // derivative evaluations so far
};

static void newtStateFree(TreeDevice *d)
{
    NewtState *s = d->newt;
    if (!s) return;
    for (double *b : s->cl2)
        if (b) cudaFree(b);
    if (s->PT) cudaFree(s->PT);
    if (s->rootTbl) cudaFree(s->rootTbl);
    if (s->decks) cudaFree(s->decks);
    if (s->partials) cudaFree(s->partials);
    if (s->result) cudaFree(s->result);
    if (s->ticket) cudaFree(s->ticket);
    if (s->hResult) cudaFreeHost(s->hResult);
    delete s;
    d->newt = nullptr;
}

// pf.p4_newtSetup (Pf/p4_treeNewt.c:11-75): allocate cl2 and the derivative decks, once per tree.
int treeNewtSetup(Tree *t)
{
    if (!t->dev) { setError("tree has no device state"); return 1; }
    // the data parts may have been re-compressed since the tree was laid out: cl2 must be sized for the
    // layout the evaluation will use, not for one that is about to be discarded
    if (ensureFresh(t, true)) return 1;
    if (!t->dev) { setError("tree has no device state"); return 1; }
    TreeDevice *d = t->dev;
    if (d->newt) return 0;
    if (!t->root) { setError("p4_newtSetup: the tree has no root"); return 1; }
    if (t->root->isLeaf) { setError("p4_newtSetup: the root is a leaf; the reference's cl2 recursion does not handle that either (Pf/p4_node.c:905-908)"); return 1; }
    if (d->scalers) { setError("p4_newtSetup: not available on trees created with per-pattern scalers"); return 1; }
    std::unique_ptr<NewtState> s(new NewtState());
    s->slotOf.assign(t->nNodes, -1);
    int next = 0;
    for (Node *n : t->nodes)
        if (n && !n->isLeaf) s->slotOf[n->nodeNum] = next++;
    if (next == 0) { setError("p4_newtSetup: the tree has no internal node"); return 1; }
    s->nSlots = next;
    const int shared = s->slotOf[t->root->nodeNum];
    for (Node *n : t->nodes)
        if (n && n->isLeaf) s->slotOf[n->nodeNum] = shared;
    s->cl2.assign(t->nParts, nullptr);
    s->deckOff.assign(t->nParts, 0);
    size_t deckTotal = 0;
    long long bytes = 0;
    d->newt = s.release();
    NewtState *S = d->newt;
    for (int p = 0; p < t->nParts; p++) {
        PartLayout &L = d->parts[p];
        if (L.nCat > 16) { setError("p4_newtSetup: more than 16 rate categories"); newtStateFree(d); return 1; }
        const size_t b = L.clNodeDoubles * sizeof(double) * (size_t)S->nSlots;
        if (cudaMalloc(&S->cl2[p], b) != cudaSuccess) {
            cudaGetLastError();
            setError("p4_newtSetup: cannot allocate %zu bytes for cl2 of part %d", b, p);
            newtStateFree(d);
            return 1;
        }
        bytes += (long long)b;
        S->deckOff[p] = deckTotal;
        deckTotal += (size_t)3 * L.nCat * L.dim * (L.dim + L.W);
    }
    S->maxBlocks = G.numSMs * 16;           // the derivative kernels run persistent CTAs
    bool ok = cudaMalloc(&S->ticket, sizeof(unsigned)) == cudaSuccess &&
              cudaMemsetAsync(S->ticket, 0, sizeof(unsigned), G.stream) == cudaSuccess &&
              cudaMalloc(&S->PT, d->pNodeDoubles * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&S->rootTbl, d->tblNodeDoubles * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&S->decks, deckTotal * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&S->partials, sizeof(double) * 3 * (size_t)S->maxBlocks * t->nParts) == cudaSuccess &&
              cudaMalloc(&S->result, sizeof(double) * 3 * t->nParts) == cudaSuccess &&
              cudaMallocHost(&S->hResult, sizeof(double) * 3 * t->nParts) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        setError("p4_newtSetup: cannot allocate the derivative decks");
        newtStateFree(d);
        return 1;
    }
    S->cl2NeedsUpdating.assign(t->nNodes, 0);
    d->bytes += bytes;
    return 0;
}

static inline double *nodeCL2(Node *n, int p)
{
    TreeDevice *d = n->tree->dev;
    return d->newt->cl2[p] + d->parts[p].clNodeDoubles * (size_t)d->newt->slotOf[n->nodeNum];
}

// The root's composition as a leaf lookup table (every column the same), refreshed at the start of every
// round: the compositions are borrowed buffers and may have changed since.
static int newtUploadRootTables(Tree *t)
{
    TreeDevice *d = t->dev;
    for (int p = 0; p < t->nParts; p++) {
        PartLayout &L = d->parts[p];
        ModelPart *mp = t->model->parts[p];
        const int rc = t->root->compNums[p];
        if (rc < 0 || rc >= mp->nComps || !mp->comps[rc].val) { setError("root uses comp %d which does not exist", rc); return 1; }
        std::vector<double> T(L.tblDoubles);
        for (int k = 0; k < L.nCat * L.dim; k++)
            for (int w = 0; w < L.W; w++) T[(size_t)k * L.W + w] = mp->comps[rc].val[k % L.dim];
        void *src = nullptr;
        if (stage(T.data(), T.size() * sizeof(double), &src)) return 1;
        CUDA_TRY(cudaMemcpyAsync(d->newt->rootTbl + L.tblOff, src, T.size() * sizeof(double), cudaMemcpyDeviceToDevice, G.stream));
    }
    return 0;
}

// p4_setNodeCL2 (Pf/p4_treeNewt.c:603-624): cl2 of n = what arrives at n's parent from everywhere but n.
static int newtSetCL2(Tree *t, Node *n)
{
    TreeDevice *d = t->dev;
    NewtState *S = d->newt;
    Node *par = n->parent;
    if (!par) { setError("p4_setNodeCL2: node %d has no parent", n->nodeNum); return 1; }
    if (flushPJobs()) return 1;
    for (int p = 0; p < t->nParts; p++) {
        if (treeEnsureResident(t, p)) return 1;
        PartLayout &L = d->parts[p];
        Part *dp = t->data->parts[p];
        CLArgs a;
        memset(&a, 0, sizeof(a));
        a.out = nodeCL2(n, p);
        a.ps = L.ps;
        a.dim = L.dim;
        a.nCat = L.nCat;
        a.tblW = L.W;
        int k = 0;
        if (par == t->root) {        // p4_initializeCL2ToRootComp, Pf/p4_node.c:860-881
            a.ch[k].tips = dp->dev.tips;
            a.ch[k].tbl = S->rootTbl + L.tblOff;
        } else {                     // p4_setCL2Up, Pf/p4_node.c:883-928: sum_from P_par[from][s] * cl2_par[from]
            double *PT = S->PT + L.pOff;
            transpose_deck_kernel<<<(L.nCat * L.dim * L.dim + 255) / 256, 256, 0, G.stream>>>(nodeP(par, p), PT, L.dim, L.nCat);
            CUDA_TRY(cudaGetLastError());
            G.launches++;
            a.ch[k].cl = nodeCL2(par, p);
            a.ch[k].P = PT;
        }
        k++;
        for (Node *c = par->leftChild; c; c = c->sibling) {   // p4_setCL2Down for every sibling, :614-620
            if (c == n) continue;
            if (k == kMaxChildren) {
                a.nChildren = k;
                if (launchCL(a)) return 1;
                a.accumulate = 1;
                k = 0;
            }
            CLChild &ch = a.ch[k];
            memset(&ch, 0, sizeof(ch));
            if (c->isLeaf) {
                if (c->seqNum < 0 || c->seqNum >= dp->nTax) { setError("leaf node %d has seqNum %d", c->nodeNum, c->seqNum); return 1; }
                ch.tips = dp->dev.tips + (size_t)c->seqNum * L.ps;
                ch.tbl = nodeTbl(c, p);
            } else {
                if (c->clSlot[p] < 0) { setError("internal node %d has no conditional likelihoods", c->nodeNum); return 1; }
                ch.cl = nodeCL(c, p);
                ch.P = nodeP(c, p);
            }
            k++;
        }
        if (k > 0) {
            a.nChildren = k;
            if (launchCL(a)) return 1;
        }
    }
    S->cl2NeedsUpdating[n->nodeNum] = 0;
    return 0;
}

// lnL, d lnL / dv, d2 lnL / dv2 in the length v of n's branch, everything else fixed (the body of the
// loop of p4_newtNode, Pf/p4_treeNewt.c:238-520).  cl2 of n and the CL of n must be current.
static int newtDerivs(Tree *t, Node *n, double out[3])
{
    TreeDevice *d = t->dev;
    NewtState *S = d->newt;
    if (flushPJobs()) return 1;
    for (int p = 0; p < t->nParts; p++) {
        if (treeEnsureResident(t, p)) return 1;
        PartLayout &L = d->parts[p];
        ModelPart *mp = t->model->parts[p];
        Part *dp = t->data->parts[p];
        const int c = n->compNums[p], r = n->rMatrixNums[p];
        if (c < 0 || c >= mp->nComps || r < 0 || r >= mp->nRMatrices) { setError("node %d part %d uses comp %d rMatrix %d which do not exist", n->nodeNum, p, c, r); return 1; }
        if (mp->bQETneedsReset[c * mp->nRMatrices + r])
            if (resetBQET(t->model, p, c, r)) return 1;
        if (eigEnsureUploaded(t, p, c, r)) return 1;
        const Gdasrv *g = nullptr;
        if (mp->nGdasrvs) {
            const int gi = n->gdasrvNums[p];
            if (gi < 0 || gi >= mp->nGdasrvs || !mp->gdasrvs[gi]) { setError("node %d part %d uses gdasrv %d which does not exist", n->nodeNum, p, gi); return 1; }
            g = mp->gdasrvs[gi];
        }
        NewtDeckJob j;
        memset(&j, 0, sizeof(j));
        j.eig = d->eig + L.eigOff + L.eigStride * (size_t)(c * mp->nRMatrices + r);
        j.eq = d->eqMasks + L.eqOff;
        j.decks = S->decks + S->deckOff[p];
        j.dim = L.dim;
        j.nCat = L.nCat;
        j.tblW = n->isLeaf ? L.W : 0;
        for (int cat = 0; cat < mp->nCat; cat++) {
            // P itself as p4_calculateBigPDecksPart forms its argument (Pf/p4_node.c:321-345); the derivative
            // decks as p4_calculateBigPDecks_1stD / _2ndD form theirs (:456-485, :505-534) -- including the
            // second derivative's extra relRate factor in the gamma, no-pInvar branch (:510-514)
            if (mp->pInvar == 0.0) {
                if (g) {
                    const double temp = g->rates[cat] * mp->relRate;
                    j.t0[cat] = n->brLen * g->rates[cat] * mp->relRate;
                    j.t1[cat] = n->brLen * temp;
                    j.r1[cat] = temp;
                    j.r2[cat] = temp * mp->relRate;
                } else {
                    j.t0[cat] = j.t1[cat] = n->brLen * mp->relRate;
                    j.r1[cat] = j.r2[cat] = mp->relRate;
                }
            } else {
                if (g) {
                    const double temp = g->rates[cat] * mp->relRate;
                    j.t0[cat] = (n->brLen * g->rates[cat] * mp->relRate) / (1.0 - mp->pInvar);
                    j.t1[cat] = (n->brLen * temp) / (1.0 - mp->pInvar);
                    j.r1[cat] = j.r2[cat] = temp / (1.0 - mp->pInvar);
                } else {
                    j.t0[cat] = j.t1[cat] = (n->brLen * mp->relRate) / (1.0 - mp->pInvar);
                    j.r1[cat] = j.r2[cat] = mp->relRate / (1.0 - mp->pInvar);
                }
            }
        }
        const bool dna = L.dim == 4 && (L.nCat == 4 || L.nCat == 1);
        if (!dna) {
            const size_t deckSm = ((size_t)5 * L.dim * L.dim + 3 * L.dim) * sizeof(double);      // 161 KB at 64 states
            static bool deckAttr = false;
            if (!deckAttr) {
                CUDA_TRY(cudaFuncSetAttribute(newt_deck_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                deckAttr = true;
            }
            newt_deck_kernel<<<L.nCat, 256, deckSm, G.stream>>>(j);
            CUDA_TRY(cudaGetLastError());
            G.launches++;
        }
        NewtArgs a;
        memset(&a, 0, sizeof(a));
        a.cl2 = nodeCL2(n, p);
        if (n->isLeaf) {
            if (n->seqNum < 0 || n->seqNum >= dp->nTax) { setError("leaf node %d has seqNum %d", n->nodeNum, n->seqNum); return 1; }
            a.tips = dp->dev.tips + (size_t)n->seqNum * L.ps;
        } else {
            if (n->clSlot[p] < 0) { setError("internal node %d has no conditional likelihoods", n->nodeNum); return 1; }
            a.cl = nodeCL(n, p);
        }
        a.decks = j.decks;
        a.counts = dp->dev.counts;
        a.invarMask = dp->dev.invarMask;
        a.ps = L.ps;
        a.nPat = L.nPat;
        a.dim = L.dim;
        a.nCat = L.nCat;
        a.tblW = L.W;
        a.pInvar = mp->pInvar;
        if (a.pInvar != 0.0 && !a.invarMask) { setError("pInvar is set but pf.setGlobalInvarSitesVec was not called on part %d", p); return 1; }
        const int rc = t->root->compNums[p];
        if (rc < 0 || rc >= mp->nComps || !mp->comps[rc].val) { setError("root uses comp %d which does not exist", rc); return 1; }
        for (int s = 0; s < L.dim; s++) a.pi[s] = mp->comps[rc].val[s];
        a.partials = S->partials + (size_t)3 * S->maxBlocks * p;
        if (dna) {
            // one launch: decks in the prologue, two patterns per thread, the last CTA folds
            const size_t smBytes = ((size_t)3 * L.nCat * 16 + (n->isLeaf ? (size_t)3 * L.nCat * 4 * L.W : 0)) * sizeof(double);
            static int resident4 = 0, resident1 = 0;
            int &resident = L.nCat == 4 ? resident4 : resident1;
            if (!resident) {
                const size_t worst = ((size_t)3 * L.nCat * 16 + (size_t)3 * L.nCat * 4 * 64) * sizeof(double);
                if (L.nCat == 4) {
                    CUDA_TRY(cudaFuncSetAttribute(newt_dna_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)worst));
                    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, newt_dna_kernel<4>, 256, smBytes));
                } else {
                    CUDA_TRY(cudaFuncSetAttribute(newt_dna_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)worst));
                    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, newt_dna_kernel<1>, 256, smBytes));
                }
                if (resident < 1) resident = 1;
                if (resident > 8) resident = 8;
            }
            if (L.W > 64) { setError("leaf table too wide for the 4-state derivative kernel"); return 1; }
            int grid = (L.ps / 2 + 255) / 256;
            if (grid > G.numSMs * resident) grid = G.numSMs * resident;
            if (L.nCat == 4) newt_dna_kernel<4><<<grid, 256, smBytes, G.stream>>>(a, j, S->ticket, S->result + 3 * p);
            else newt_dna_kernel<1><<<grid, 256, smBytes, G.stream>>>(a, j, S->ticket, S->result + 3 * p);
            CUDA_TRY(cudaGetLastError());
            G.launches++;
            continue;
        }
        const size_t sm = (size_t)3 * L.nCat * L.dim * (n->isLeaf ? L.W : L.dim) * sizeof(double);
        a.useSmem = sm <= 96 * 1024 ? 1 : 0;
        const size_t smBytes = a.useSmem ? sm : 0;
        int blocks = (L.ps + 127) / 128;
        if (blocks > S->maxBlocks) blocks = S->maxBlocks;
        static bool attrSet = false;
        if (!attrSet) {
            CUDA_TRY(cudaFuncSetAttribute(newt_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            CUDA_TRY(cudaFuncSetAttribute(newt_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            attrSet = true;
        }
        if (L.dim == 20 && !n->isLeaf && g_dmmaEnabled && g_newtDmma && (size_t)3 * L.nCat * 480 * sizeof(double) <= 200 * 1024 && S->maxBlocks >= 3 * G.numSMs) {
            // an internal node on the FP64 tensor cores: persistent CTAs, 16 patterns per warp and tile
            const size_t fragBytes = (size_t)3 * L.nCat * 480 * sizeof(double);
            typedef void (*NFn)(const NewtArgs, unsigned *, double *);
            static int shape = -1;          // 0: 4 warps x 3 CTAs per SM (168 registers); 1: 8 warps x 2 (128 registers)
            static const NFn fns[2][2] = {{(NFn)newt_aa_dmma_kernel<4, 3, 4>, (NFn)newt_aa_dmma_kernel<4, 3, 0>},
                                          {(NFn)newt_aa_dmma_kernel<8, 2, 4>, (NFn)newt_aa_dmma_kernel<8, 2, 0>}};
            if (shape < 0) {
                const char *e = getenv("P4B_NEWT_DMMA_SHAPE");
                shape = e ? (atoi(e) != 0) : 0;
                for (int i = 0; i < 4; i++) CUDA_TRY(cudaFuncSetAttribute(fns[i >> 1][i & 1], cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            }
            const int warps = shape == 0 ? 4 : 8, want = shape == 0 ? 3 : 2;
            int perSM = (int)((227 * 1024) / (fragBytes + 2048));
            if (perSM > want) perSM = want;
            if (perSM < 1) perSM = 1;
            int grid = (L.ps / 16 + warps - 1) / warps;
            if (grid > G.numSMs * perSM) grid = G.numSMs * perSM;
            fns[shape][L.nCat == 4 ? 0 : 1]<<<grid, warps * 32, fragBytes, G.stream>>>(a, S->ticket, S->result + 3 * p);
            CUDA_TRY(cudaGetLastError());
            G.launches++;
            continue;
        }
        if (L.dim == 20 && a.useSmem) {
            // two patterns per thread, the last CTA folds: deck launch + this one
            static int residentAA = 0;
            if (!residentAA) {
                CUDA_TRY(cudaFuncSetAttribute(newt_aa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
                CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&residentAA, newt_aa_kernel, 128, smBytes));
                if (residentAA < 1) residentAA = 1;
                if (residentAA > 8) residentAA = 8;
            }
            int grid = (L.ps / 2 + 127) / 128;
            if (grid > G.numSMs * residentAA) grid = G.numSMs * residentAA;
            newt_aa_kernel<<<grid, 128, smBytes, G.stream>>>(a, S->ticket, S->result + 3 * p);
            CUDA_TRY(cudaGetLastError());
            G.launches++;
            continue;
        }
        if (L.dim == 20) newt_kernel<20><<<blocks, 128, smBytes, G.stream>>>(a);
        else newt_kernel<0><<<blocks, 128, smBytes, G.stream>>>(a);
        CUDA_TRY(cudaGetLastError());
        newt_final_kernel<<<1, 256, 0, G.stream>>>(a.partials, blocks, S->result + 3 * p);
        CUDA_TRY(cudaGetLastError());
        G.launches += 2;
    }
    const int nRes = 3 * t->nParts;
    if (commActive())
        if (commAllReduceSum(S->result, nRes, (void *)G.stream)) return 1;
    CUDA_TRY(cudaMemcpyAsync(S->hResult, S->result, nRes * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
    if (streamSync()) return 1;
    out[0] = out[1] = out[2] = 0.0;
    for (int p = 0; p < t->nParts; p++)
        for (int k = 0; k < 3; k++) out[k] += S->hResult[3 * p + k];
    S->iters++;
    return 0;
}

// p4_newtNode (Pf/p4_treeNewt.c:210-600): Newton-Raphson on one branch length, the reference's guards.
static int newtNode(Tree *t, Node *n, double epsilon)
{
    const double BRLEN_MIN = t->model->BRLEN_MIN[0], BRLEN_MAX = t->model->BRLEN_MAX[0];
    const double oldBrLen = n->brLen;
    double currentGuess = oldBrLen, nextGuess = 0.0;
    int iter = 0;
    while (true) {
        double r[3];
        if (newtDerivs(t, n, r)) return 1;
        const double firstD = r[1], secondD = r[2];
        nextGuess = currentGuess - (firstD / secondD);
        if (secondD >= 0.0) nextGuess = currentGuess / 5.0;
        if (nextGuess < BRLEN_MIN) { n->brLen = BRLEN_MIN; break; }
        if (nextGuess >= 5.0 * oldBrLen) { n->brLen = 5.0 * oldBrLen; break; }
        if (nextGuess > BRLEN_MAX) { n->brLen = BRLEN_MAX; break; }
        if (iter > 20) { n->brLen = currentGuess; break; }
        iter++;
        if (fabs(firstD) < epsilon) { n->brLen = currentGuess; break; }
        currentGuess = nextGuess;
        n->brLen = nextGuess;
    }
    return nodeCalculateBigPDecks(n);
}

// p4_newtAround (Pf/p4_treeNewt.c:78-205): rounds over all branches in post-order until the
// log-likelihood moves by less than likeDelta (at most 20 rounds).  Returns the last log-likelihood.
double treeNewtAround(Tree *t, double epsilon, double likeDelta)
{
    if (!t->dev) { setError("tree has no device state"); return NAN; }
    if (!t->dev->newt) { setError("p4_newtAround: call p4_newtSetup first"); return NAN; }
    // A data-version change (pf.makePatterns, setGlobalInvarSitesVec, p4_simulate) lays the tree out again
    // and drops the Newton work arrays with the old layout: do that NOW and set them up again for the new
    // one, so that no pointer into the old state is held across it.
    if (ensureFresh(t, true)) return NAN;
    if (!t->dev) { setError("tree has no device state"); return NAN; }
    if (!t->dev->newt && treeNewtSetup(t)) return NAN;
    if (t->root && t->root->isLeaf) { setError("p4_newtAround: the root is a leaf"); return NAN; }
    double previous = treeLogLike(t, 0);
    if (previous != previous) return NAN;
    NewtState *S = t->dev ? t->dev->newt : nullptr;
    if (!S) { setError("p4_newtAround: the Newton state was lost during the evaluation"); return NAN; }
    for (Node *n : t->nodes)
        if (n) n->clNeedsUpdating = 0;
    std::vector<Node *> order;
    for (int j = 0; j < t->nNodes; j++) {
        const int i = t->postOrder[j];
        if (i == P4B_NO_ORDER) continue;
        Node *n = t->nodes[i];
        if (n && n != t->root) order.push_back(n);
    }
    std::vector<Node *> path;
    double thisLike = previous;
    for (int round = 0; round < 20; round++) {
        if (newtUploadRootTables(t)) return NAN;
        for (Node *n : order) S->cl2NeedsUpdating[n->nodeNum] = 1;
        for (Node *n : order) {
            if (n->clNeedsUpdating && n->leftChild) {       // at most the one node below the branches just changed
                for (int p = 0; p < t->nParts; p++)
                    if (nodeSetCLImpl(n, p, false)) return NAN;
            }
            if (S->cl2NeedsUpdating[n->nodeNum]) {          // from the highest stale ancestor down to n
                path.clear();
                for (Node *q = n; q->parent && S->cl2NeedsUpdating[q->nodeNum]; q = q->parent) path.push_back(q);
                for (size_t k = path.size(); k-- > 0;)
                    if (newtSetCL2(t, path[k])) return NAN;
            }
            if (newtNode(t, n, epsilon)) return NAN;
            for (Node *q = n->parent; q; q = q->parent) q->clNeedsUpdating = 1;
        }
        thisLike = treeLogLike(t, 0);
        if (thisLike != thisLike) return NAN;
        for (Node *n : t->nodes)
            if (n) n->clNeedsUpdating = 0;
        const double diff = thisLike - previous;
        if (fabs(diff) < likeDelta) break;
        previous = thisLike;
    }
    return thisLike;
}

// Test / inspection hooks: the derivatives at the node's current branch length, and its cl2.
int nodeNewtDerivs(Node *n, double out[3])
{
    Tree *t = n->tree;
    if (!t->dev || !t->dev->newt) { setError("p4b_newtDerivs: call p4_newtSetup first"); return 1; }
    if (ensureFresh(t, true)) return 1;
    if (!t->dev) { setError("tree has no device state"); return 1; }
    if (!t->dev->newt && treeNewtSetup(t)) return 1;
    if (n == t->root || !n->parent) { setError("p4b_newtDerivs: the root has no branch"); return 1; }
    if (treeFlushAllPending(t)) return 1;
    if (newtUploadRootTables(t)) return 1;
    // cl2 of every node from the root's child down to n (nothing is assumed current)
    std::vector<Node *> path;
    for (Node *q = n; q->parent; q = q->parent) path.push_back(q);
    for (size_t k = path.size(); k-- > 0;)
        if (newtSetCL2(t, path[k])) return 1;
    return newtDerivs(t, n, out);
}

int nodeGetCL2(Node *n, int p, double *out)
{
    Tree *t = n->tree;
    if (!t->dev || !t->dev->newt || p < 0 || p >= t->nParts) { setError("p4b_getNodeCL2: no Newton state or bad part"); return 1; }
    if (n == t->root) { setError("p4b_getNodeCL2: the root has no cl2"); return 1; }
    PartLayout &L = t->dev->parts[p];
    std::vector<double> h(L.clNodeDoubles);
    CUDA_TRY(cudaMemcpyAsync(h.data(), nodeCL2(n, p), L.clNodeDoubles * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
    if (streamSync()) return 1;
    for (int k = 0; k < L.nCat * L.dim; k++) memcpy(out + (size_t)k * L.nPat, h.data() + (size_t)k * L.ps, sizeof(double) * L.nPat);
    return 0;
}

long long treeNewtIterations(Tree *t) { return (t->dev && t->dev->newt) ? t->dev->newt->iters : 0; }


// ---------------------------------------------------------------------------
// p4_simulate, device part (Pf/p4_treeSim.c:315-360): picker decks from the current P decks, then the nodes in
// preOrder, a chunk at a time; the uniforms of a chunk are produced by `fill` in the reference's stream order
// and shipped through one pinned buffer.  Leaf states come back into part->sequences.
// ---------------------------------------------------------------------------
int treeSimulateDevice(Tree *t, int p, const uint8_t *cats, const uint8_t *rootStates, const uint8_t *invar, const int *rank,
                       int nVar, const std::function<void(double *, size_t)> &fill)
{
    if (G.world > 1) { setError("p4_simulate is not available when patterns are sharded over several processes"); return 1; }
    TreeDevice *d = t->dev;
    PartLayout &L = d->parts[p];
    Part *dp = t->data->parts[p];
    const int nChar = dp->nChar, dim = L.dim, nCat = L.nCat;
    if (flushPJobs()) return 1;
    uint8_t *dStates = nullptr, *dCats = nullptr, *dInv = nullptr;
    int *dRank = nullptr;
    double *dPicker = nullptr, *dU = nullptr, *hU = nullptr;
    std::vector<Node *> order;
    for (int j = 0; j < t->nNodes; j++) {
        const int i = t->preOrder[j];
        if (i == P4B_NO_ORDER) continue;
        Node *n = t->nodes[i];
        if (n && n != t->root) order.push_back(n);
    }
    // chunk size: at most kSimChunk nodes and about 64 MB of uniforms
    size_t perNode = (size_t)(nVar > 0 ? nVar : 1);
    int chunk = (int)((64u << 20) / (perNode * sizeof(double)));
    if (chunk < 1) chunk = 1;
    if (chunk > kSimChunk) chunk = kSimChunk;
    int rc = 1;
    do {
        if (cudaMalloc(&dStates, (size_t)t->nNodes * nChar) != cudaSuccess) break;
        if (cudaMalloc(&dCats, nChar) != cudaSuccess || cudaMalloc(&dInv, nChar) != cudaSuccess) break;
        if (cudaMalloc(&dRank, sizeof(int) * (size_t)nChar) != cudaSuccess) break;
        if (cudaMalloc(&dPicker, sizeof(double) * (size_t)t->nNodes * nCat * dim * dim) != cudaSuccess) break;
        if (cudaMalloc(&dU, sizeof(double) * perNode * chunk) != cudaSuccess) break;
        if (cudaMallocHost(&hU, sizeof(double) * perNode * chunk) != cudaSuccess) break;
        if (cudaMemcpyAsync(dCats, cats, nChar, cudaMemcpyHostToDevice, G.stream) != cudaSuccess) break;
        if (cudaMemcpyAsync(dInv, invar, nChar, cudaMemcpyHostToDevice, G.stream) != cudaSuccess) break;
        if (cudaMemcpyAsync(dRank, rank, sizeof(int) * (size_t)nChar, cudaMemcpyHostToDevice, G.stream) != cudaSuccess) break;
        if (cudaMemcpyAsync(dStates + (size_t)t->root->nodeNum * nChar, rootStates, nChar, cudaMemcpyHostToDevice, G.stream) != cudaSuccess) break;
        picker_kernel<<<G.numSMs * 2, 256, 0, G.stream>>>(d->P, dPicker, (long long)d->pNodeDoubles, (int)L.pOff, dim, nCat, t->nNodes);
        if (cudaGetLastError() != cudaSuccess) break;
        G.launches++;
        bool ok = true;
        for (size_t j0 = 0; j0 < order.size() && ok; j0 += chunk) {
            const int n = (int)std::min((size_t)chunk, order.size() - j0);
            if (cudaStreamSynchronize(G.stream) != cudaSuccess) { ok = false; break; }   // the pinned buffer is free again
            fill(hU, (size_t)n * nVar);
            SimArgs a;
            memset(&a, 0, sizeof(a));
            a.states = dStates;
            a.cats = dCats;
            a.invar = dInv;
            a.rank = dRank;
            a.picker = dPicker;
            a.U = dU;
            a.nChar = nChar;
            a.nVar = nVar;
            a.dim = dim;
            a.nCat = nCat;
            a.n = n;
            for (int j = 0; j < n; j++) {
                a.node[j] = order[j0 + j]->nodeNum;
                a.parent[j] = order[j0 + j]->parent ? order[j0 + j]->parent->nodeNum : t->root->nodeNum;
            }
            if (nVar > 0 && cudaMemcpyAsync(dU, hU, sizeof(double) * (size_t)n * nVar, cudaMemcpyHostToDevice, G.stream) != cudaSuccess) { ok = false; break; }
            simulate_kernel<<<(nChar + 255) / 256, 256, 0, G.stream>>>(a);
            if (cudaGetLastError() != cudaSuccess) { ok = false; break; }
            G.launches++;
        }
        if (!ok) break;
        std::vector<uint8_t> row(nChar);
        for (Node *n : t->nodes) {
            if (!n || !n->isLeaf) continue;
            if (n->seqNum < 0 || n->seqNum >= dp->nTax) { ok = false; break; }
            if (cudaMemcpyAsync(row.data(), dStates + (size_t)n->nodeNum * nChar, nChar, cudaMemcpyDeviceToHost, G.stream) != cudaSuccess) { ok = false; break; }
            if (cudaStreamSynchronize(G.stream) != cudaSuccess) { ok = false; break; }
            int *seq = &dp->sequences[(size_t)n->seqNum * nChar];
            for (int k = 0; k < nChar; k++) seq[k] = row[k];
        }
        if (!ok) break;
        rc = 0;
    } while (0);
    if (rc) {
        cudaError_t e = cudaGetLastError();
        setError("p4_simulate: device step failed (%s)", e == cudaSuccess ? "allocation, copy, or a leaf without a sequence" : cudaGetErrorString(e));
    }
    cudaStreamSynchronize(G.stream);
    if (dStates) cudaFree(dStates);
    if (dCats) cudaFree(dCats);
    if (dInv) cudaFree(dInv);
    if (dRank) cudaFree(dRank);
    if (dPicker) cudaFree(dPicker);
    if (dU) cudaFree(dU);
    if (hU) cudaFreeHost(hU);
    G.stageDirtySinceSync = false;
    G.stageHead = 0;
    return rc;
}


// Host copy of the root's CL of part p, refreshed when the root's CL was recomputed since (p4_drawAncState reads one
// column per call; p4 calls it once per site).  Returns NULL on error.
const double *treeRootCLHost(Tree *t, int p, int *psOut)
{
    if (!t->dev) { setError("tree has no device state"); return nullptr; }
    if (p < 0 || p >= t->nParts) { setError("bad part %d", p); return nullptr; }
    if (G.world > 1) { setError("p4_drawAncState is not available when patterns are sharded over several processes"); return nullptr; }
    if (ensureFresh(t, true)) return nullptr;
    TreeDevice *d = t->dev;
    Node *root = t->root;
    if (!root || root->clSlot[p] < 0) { setError("the root has no conditional likelihoods (calculate the likelihood first)"); return nullptr; }
    if (treeEnsureResident(t, p)) return nullptr;
    PartLayout &L = d->parts[p];
    if (d->rootCLHost.size() != (size_t)t->nParts) {
        d->rootCLHost.assign(t->nParts, std::vector<double>());
        d->rootCLStamp.assign(t->nParts, 0);
    }
    if (d->rootCLStamp[p] != root->clStamp[p] || root->clStamp[p] == 0 || d->rootCLHost[p].size() != L.clNodeDoubles) {
        d->rootCLHost[p].resize(L.clNodeDoubles);
        if (cudaMemcpyAsync(d->rootCLHost[p].data(), nodeCL(root, p), L.clNodeDoubles * sizeof(double), cudaMemcpyDeviceToHost, G.stream) != cudaSuccess ||
            streamSync()) {
            setError("p4_drawAncState: cannot read the root's conditional likelihoods");
            return nullptr;
        }
        d->rootCLStamp[p] = root->clStamp[p];
    }
    *psOut = L.ps;
    return d->rootCLHost[p].data();
}

void *engineStream() { return (void *)G.stream; }
int engineInitPublic() { return engineInit(); }

}  // namespace p4b
