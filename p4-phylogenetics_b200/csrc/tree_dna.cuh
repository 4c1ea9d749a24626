// tree_dna.cuh -- whole-tree CL recursion for 4-state parts, second generation (sm_100a, FP64).
//
// Computes what Pf/p4_node.c:636-857 (p4_setConditionalLikelihoodsOfInternalNodePart) computes for every node
// of a step list, and -- when the list ends at the root -- what Pf/p4_tree.c:1029-1378 (p4_partLogLikeLoop /
// ...RootLeaf) computes from the root's CL, in ONE launch.
//
// A thread owns two adjacent patterns (16-byte accesses) and CT of the part's rate categories; it walks the
// step list by itself, the CL of the node it has just computed staying in registers.  What the first
// generation (cl_tree_dna_kernel, kernels.cuh) paid per step and per warp -- decoding the step from
// runtime-indexed constant memory, staging the children's P decks / leaf tables with per-thread cp.async, a
// CTA-wide barrier -- is taken off the path:
//   * the tree's whole step list, pre-decoded by the host (48 bytes per step), is copied into shared memory
//     once, in the prologue: a step's descriptor is two 16-byte shared-memory loads;
//   * the operands of a step's (at most two) children -- P decks, leaf tables, and the CTA's tip codes of a leaf
//     child -- arrive in a ring of slots by bulk copies (TMA, cp.async.bulk) that complete on the slot's
//     mbarrier.  Nobody is a dedicated producer: the LAST warp to finish step s (a shared-memory counter per
//     slot) issues the copies of step s + RING into the slot it has just freed.  Warps wait on the slot's
//     mbarrier only (try_wait, normally already satisfied) and drift up to RING - 1 steps apart: no CTA barrier
//     inside the step loop, no per-thread global load on the step's path.
// The one internal child of a step that is not in registers comes from a per-thread shared-memory buffer that
// the host's step planner fills in one of two ways: the producing step PUSHES its result there (no global
// re-read at all; short-lived siblings), or the thread PREFETCHES it from its own earlier global store with
// cp.async one or more steps ahead.
// Rate categories can be split over CSPLIT = NCAT / CT warps (more warps per SM at fewer registers each).
// After the last step the root CL (in registers) is folded into site likelihoods and the CTA's partial lnL;
// the LAST CTA to finish (atomic ticket) folds the partials in a fixed order: no second launch.
#pragma once
#include "kernels.cuh"

namespace p4b {

constexpr unsigned kNone = 0xffffffffu;

// flags of a step
constexpr unsigned kStepFirst = 4u, kStepStore = 8u, kStepPush = 256u, kStepPfLate = 512u;

struct Step2 {            // 48 bytes in global memory, written by the host's step planner
    unsigned out;         // the node's CL buffer: (address - hdr.arena) / 256 bytes
    unsigned flags;       // bits 0-1 children (1..2) | kStepFirst | kStepStore | kind of child 0 << 4 | kind of child 1 << 6 | kStepPush | kStepPfLate
                          //   kind 0 internal child, loaded from its buffer now; 1 internal child in registers (previous step);
                          //   2 leaf child; 3 internal child in the thread's shared-memory buffer (pushed or prefetched)
    unsigned c0, c1;      // kind 0: the child's CL buffer (256-byte units)
    unsigned nt0, nt1;    // tip rows (sequence numbers) of this step's leaf children 0 / 1 (kNone: not a leaf)
    unsigned pf;          // CL buffer to prefetch into the shared-memory buffer during this step (kNone: none)
    unsigned pad;
    unsigned n0, n1;      // where the children's operands lie: offset (doubles) of the leaf table from hdr.tbl / of the P deck from hdr.Pdeck
    unsigned pad2[2];
};
static_assert(sizeof(Step2) == 48, "Step2 layout");

struct TreeHdr2 {
    double *arena;            // base address the tree's CL buffers are addressed from
    const double *Pdeck;      // tree's P decks, already offset to this part
    const double *tbl;        // tree's leaf tables, already offset to this part
    double *patLikes;         // optional [ps]
    double *partials;         // [2*gridDim.x]
    double *result;           // [2]: sum of count*log(like), count of like <= 0 -- written by the last CTA
    unsigned *ticket;
    const uint8_t *rootTips;  // non-NULL when the root is a leaf
    double pInvar;
    double pi[4];
    int stepBase, nSteps;
    unsigned t0, t1, pf0;     // pf0: buffer to prefetch for step 0 in the prologue (t0, t1 unused)
    int doLike;
};

struct TreeArgs2 {
    int ps, nPat, tblW, nTrees;
    int tileLog, maxSteps;    // maxSteps: most steps any tree of the launch has (sizes the shared-memory step list)
                              // CL layout: 0 = rows [k][ps]; else tiles of 2^tileLog patterns, [tile][k][2^tileLog]
    long long pNodeDoubles;   // stride between nodes in a P deck
    long long tblNodeDoubles;
    const uint8_t *tips;      // part's tip rows [nTax][ps]
    const int *counts;
    const uint64_t *invarMask;
    const uint64_t *eqMask;
    const Step2 *steps;
    MailArgs mail;            // world > 1: the last CTA exchanges the folded sums with the other pattern shards (kernels.cuh)
    TreeHdr2 hdr[kMaxBatchTrees];
};

// Shared memory of one CTA (all offsets multiples of 16 bytes):
//   step digests [maxSteps] x 16 B, node numbers of the children [maxSteps] x 8 B |
//   operand ring [RING][ops child 0 | ops child 1 | tips child 0 | tips child 1] | per-thread buffers [KT][CW*32] double2 |
//   category hand-over [CW*32] double2 | mbarriers full[RING] | slot counters [RING] | reduction scratch
// A digest is what a warp needs of a step on its fast path: {out, pf, flags, -}; the node numbers address the operand
// copies; children loaded directly from global memory (rare) are looked up in the global record.
__host__ __device__ inline int treeDna2Ring(int CW) { return CW >= 4 ? 8 : 4; }
__host__ __device__ inline size_t treeDna2OpsDoubles(int K, int W) { return (size_t)K * (W > 4 ? W : 4); }
__host__ __device__ inline size_t treeDna2StepBytes(int maxSteps) { return ((size_t)maxSteps * 24 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t treeDna2SlotBytes(int K, int W, int PB) { return 2 * treeDna2OpsDoubles(K, W) * 8 + 2 * (size_t)PB * 64; }
__host__ inline size_t treeDna2SmemBytes(int nCat, int W, int CT, int CW, int maxSteps)
{
    const int K = nCat * 4, PB = CW / (nCat / CT), RING = treeDna2Ring(CW);
    return treeDna2StepBytes(maxSteps) + RING * treeDna2SlotBytes(K, W, PB) + (size_t)CT * 4 * CW * 32 * 16 + (size_t)CW * 32 * 16 +
           RING * 8 + RING * 8 + 16 + (2 * CW + 2) * 8;
}

__device__ __forceinline__ double2 lds2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ bool mbar_test(uint64_t *bar, unsigned parity)      // one try_wait: has the phase completed?
{
    unsigned done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity)
                 : "memory");
    return done != 0u;
}

// Leaving a ring slot: a plain arrival on the slot's "empty" mbarrier.  (A shared-memory counter needs a fence in front of its
// atomicAdd to order the warp's reads of the slot before it -- MEMBAR.SC.CTA, which also waits for every global STORE the
// lane has in flight: the full latency of a store to L2 on every warp's path, once per step.  An mbarrier arrival carries
// that ordering for shared memory by itself: one SYNCS.ARRIVE, no MEMBAR.)
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_poll(uint64_t *bar, unsigned parity)      // non-blocking test: has the phase completed?
{
    unsigned done;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity)
                 : "memory");
    return done != 0u;
}

// Factor contributed by one child to the 4 states of category `cat` (same association as kernels.cuh child_factor).
//   KIND 1: the child's CL is in `cur` (rows c*4..c*4+3)      KIND 3: in the thread's shared-memory buffer
//   KIND 2: leaf, table lookup with the two tip codes          KIND 0: in global memory at `mem` (row stride rs)
template <int KIND, int CTH>
__device__ __forceinline__ void child2(int c, int cat, const double *__restrict__ s, int W, unsigned cx, unsigned cy, const double2 *cur,
                                       const double2 *buf, const double *__restrict__ mem, unsigned rs, double2 f[4])
{
    if (KIND == 2) {
#pragma unroll
        for (int s4 = 0; s4 < 4; s4++) {
            const int k = cat * 4 + s4;
            f[s4].x = s[k * W + cx];
            f[s4].y = s[k * W + cy];
        }
    } else {
        double2 v0, v1, v2, v3;
        if (KIND == 1) { v0 = cur[c * 4 + 0]; v1 = cur[c * 4 + 1]; v2 = cur[c * 4 + 2]; v3 = cur[c * 4 + 3]; }
        else if (KIND == 3) { v0 = buf[(c * 4 + 0) * CTH]; v1 = buf[(c * 4 + 1) * CTH]; v2 = buf[(c * 4 + 2) * CTH]; v3 = buf[(c * 4 + 3) * CTH]; }
        else {
            v0 = ld2(mem + (size_t)(c * 4 + 0) * rs); v1 = ld2(mem + (size_t)(c * 4 + 1) * rs);
            v2 = ld2(mem + (size_t)(c * 4 + 2) * rs); v3 = ld2(mem + (size_t)(c * 4 + 3) * rs);
        }
#pragma unroll
        for (int s4 = 0; s4 < 4; s4++) {
            const double2 p01 = lds2(s + cat * 16 + s4 * 4), p23 = lds2(s + cat * 16 + s4 * 4 + 2);
            double2 sum;
            sum.x = p01.x * v0.x;
            sum.y = p01.x * v0.y;
            sum.x = fma(p01.y, v1.x, sum.x);
            sum.y = fma(p01.y, v1.y, sum.y);
            sum.x = fma(p23.x, v2.x, sum.x);
            sum.y = fma(p23.x, v2.y, sum.y);
            sum.x = fma(p23.y, v3.x, sum.x);
            sum.y = fma(p23.y, v3.y, sum.y);
            f[s4] = sum;
        }
    }
}

// A first step with exactly two children -- nearly every step of a binary tree -- as straight-line code for one
// combination of child kinds.
template <int CT, int CTH, int K0, int K1>
__device__ __forceinline__ void step2(bool store, double2 *cur, const double2 *buf, int opOff, const double *__restrict__ s0,
                                      const double *__restrict__ s1, int W, unsigned code0, unsigned code1, unsigned rs, double *__restrict__ outp)
{
    const unsigned c0x = code0 & 0xffu, c0y = (code0 >> 8) & 0xffu, c1x = code1 & 0xffu, c1y = (code1 >> 8) & 0xffu;
#pragma unroll
    for (int c = 0; c < CT; c++) {
        double2 f[4], g[4];
        child2<K0, CTH>(c, opOff + c, s0, W, c0x, c0y, cur, buf, nullptr, rs, f);
        child2<K1, CTH>(c, opOff + c, s1, W, c1x, c1y, cur, buf, nullptr, rs, g);
#pragma unroll
        for (int s4 = 0; s4 < 4; s4++) {
            double2 r;
            r.x = f[s4].x * g[s4].x;      // (left child) * (sibling), the reference's order
            r.y = f[s4].y * g[s4].y;
            cur[c * 4 + s4] = r;
            if (store) st2(outp + (size_t)(c * 4 + s4) * rs, r);
        }
    }
}

template <int NCAT, int CT, int CW, int MINB>
__global__ void __launch_bounds__(CW * 32, MINB)
cl_tree_dna2_kernel(const __grid_constant__ TreeArgs2 a)
{
    constexpr int K = NCAT * 4;            // rows of a CL buffer
    constexpr int KT = CT * 4;             // rows this thread owns
    constexpr int CSPLIT = NCAT / CT;      // warps sharing a block of 64 patterns
    constexpr int CTH = CW * 32;           // threads
    constexpr int PB = CW / CSPLIT;        // pattern blocks (64 patterns each) per CTA
    static_assert(NCAT % CT == 0 && CW % CSPLIT == 0, "shape");
    constexpr int RING = CW >= 4 ? 8 : 4;
    const TreeHdr2 &hd = a.hdr[blockIdx.y];
    extern __shared__ __align__(16) unsigned char smraw[];
    const int W = a.tblW;
    const int nSteps = hd.nSteps;
    const unsigned opsD = (unsigned)treeDna2OpsDoubles(K, W);
    const unsigned slotB = (unsigned)treeDna2SlotBytes(K, W, PB);     // bytes per ring slot
    uint4 *sSteps = reinterpret_cast<uint4 *>(smraw);                 // [maxSteps] digests
    uint2 *sNodes = reinterpret_cast<uint2 *>(sSteps + a.maxSteps);   // [maxSteps] where the children's operands lie (offsets into the decks)
    unsigned char *ring = smraw + treeDna2StepBytes(a.maxSteps);
    double2 *bufAll = reinterpret_cast<double2 *>(ring + RING * slotB);   // [KT][CTH]
    double2 *sA = bufAll + KT * CTH;                                  // hand-over of the category sum between the warps of a pattern block
    uint64_t *full = reinterpret_cast<uint64_t *>(sA + CTH);
    uint64_t *empty = full + RING;                                    // arrivals of the CW x 32 lanes that have left the slot
    double *sRed = reinterpret_cast<double *>(empty + RING + 1);      // [2][CW] + flag, 8-byte aligned
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t ps = (size_t)a.ps;
    const uint8_t *ctaTips = a.tips + (size_t)blockIdx.x * (PB * 64);     // the CTA's PB*64 patterns of tip row 0

    // one lane: the bulk copies (TMA) of step j's operands into its ring slot, completing on the slot's mbarrier.
    // A leaf child's copy is its lookup table plus the CTA's tip codes of its sequence; `tipRow` = its sequence number.
    auto produce = [&](int j) {
        const int slot = j & (RING - 1);
        const uint4 dg = sSteps[j];
        const uint2 nn = sNodes[j];
        const unsigned flags = dg.z;
        const unsigned pBytes = K * 4 * 8, tBytes = (unsigned)(K * W * 8), tipBytes = PB * 64;
        const unsigned nc = flags & 3u, k0 = (flags >> 4) & 3u, k1 = (flags >> 6) & 3u;
        const bool l0 = k0 == 2u, l1 = nc == 2u && k1 == 2u;
        unsigned char *sl = ring + slot * slotB;
        mbar_expect_tx(full + slot, (l0 ? tBytes + tipBytes : pBytes) + (nc == 2u ? (l1 ? tBytes + tipBytes : pBytes) : 0u));
        bulk_g2s(sl, (l0 ? hd.tbl : hd.Pdeck) + nn.x, l0 ? tBytes : pBytes, full + slot);
        if (l0) bulk_g2s(sl + 2 * opsD * 8, ctaTips + (size_t)(dg.w & 0xffffu) * ps, tipBytes, full + slot);
        if (nc == 2u) {
            bulk_g2s(sl + opsD * 8, (l1 ? hd.tbl : hd.Pdeck) + nn.y, l1 ? tBytes : pBytes, full + slot);
            if (l1) bulk_g2s(sl + 2 * opsD * 8 + tipBytes, ctaTips + (size_t)(dg.w >> 16) * ps, tipBytes, full + slot);
        }
    };

    {
        const Step2 *gSteps = a.steps + hd.stepBase;
        for (int i = threadIdx.x; i < nSteps; i += CTH) {
            const uint4 dA = __ldg(reinterpret_cast<const uint4 *>(gSteps + i)), dB = __ldg(reinterpret_cast<const uint4 *>(gSteps + i) + 1);
            // {out, pf, flags, tip row (sequence number) of a leaf child 0 | of a leaf child 1 << 16}
            sSteps[i] = make_uint4(dA.x, dB.z, dA.y, (dB.x & 0xffffu) | (dB.y << 16));
            sNodes[i] = make_uint2(__ldg(&gSteps[i].n0), __ldg(&gSteps[i].n1));
        }
        if (threadIdx.x == 0) {
            for (int i = 0; i < RING; i++) { mbar_init(full + i, 1); mbar_init(empty + i, CW * 32); }
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int j = 0; j < RING && j < nSteps; j++) produce(j);

    const int cg = warp % CSPLIT, pb = warp / CSPLIT;       // category group, pattern block of this warp
    const int pat = ((blockIdx.x * PB + pb) * 32 + lane) * 2;
    // this thread's two patterns inside a CL buffer: element offset of its first row, and the row stride (both < 2^31)
    unsigned off, rs;
    if (a.tileLog == 0) { off = (unsigned)pat; rs = (unsigned)a.ps; }
    else {
        const unsigned T = 1u << a.tileLog;
        off = (unsigned)(pat >> a.tileLog) * ((unsigned)K << a.tileLog) + ((unsigned)pat & (T - 1u));
        rs = T;
    }
    off += (unsigned)(cg * KT) * rs;
    double2 *buf = bufAll + threadIdx.x;                     // [KT][CTH]: row stride CTH double2
    const int opOff = cg * CT;                               // first category of this thread inside a P deck / leaf table
    const unsigned tipOff = 2 * opsD * 8 + (unsigned)(pb * 64 + lane * 2);   // this thread's two tip codes of child 0 inside a slot

    double2 cur[KT];
#pragma unroll
    for (int k = 0; k < KT; k++) cur[k] = make_double2(1.0, 1.0);

    auto prefetch = [&](unsigned code) {
        const double *cl = hd.arena + (size_t)code * 32 + off;
#pragma unroll
        for (int k = 0; k < KT; k++) cp_async16(buf + k * CTH, cl + (size_t)k * rs);
        cp_async_commit();
    };
    if (hd.pf0 != kNone) prefetch(hd.pf0);

    // Refilling the ring is shared out: warp w refills the slots of the steps w, w + CW, w + 2 CW, ... -- each as soon as
    // all CW warps have left that step (the slot's counter) -- so no warp carries the producer's work alone and falls
    // behind the others.  `done` = the last step this warp has finished.
    int myNext = warp;
    auto service = [&](int done) {      // lane 0 only
        while (myNext <= done && myNext + RING < nSteps) {
            if (!mbar_poll(empty + (myNext & (RING - 1)), (unsigned)(myNext / RING) & 1u)) break;   // somebody is still working from that slot
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // the warps' reads of the slot precede the bulk copy's writes
            produce(myNext + RING);
            myNext += CW;
        }
    };

    for (int si = 0; si < nSteps; si++) {
        const int slot = si & (RING - 1);
        {   // this step's operands; while they are not there, lane 0 keeps up with its refills (nobody can be starved)
            const unsigned parity = (unsigned)(si / RING) & 1u;
            int spins = 0;
            while (!mbar_test(full + slot, parity)) {
                if (lane == 0) service(si - 1);
                if (++spins > (1 << 22)) __trap();       // a lost copy must be an error, never a hang
            }
        }
        {
            const uint4 d = sSteps[si];
            const unsigned flags = d.z, pf = d.y;
            const unsigned k0 = (flags >> 4) & 3u, k1 = (flags >> 6) & 3u;
            const unsigned char *sl = ring + slot * slotB;
            const unsigned code0 = *reinterpret_cast<const unsigned short *>(sl + tipOff);              // meaningful for a leaf child only
            const unsigned code1 = *reinterpret_cast<const unsigned short *>(sl + tipOff + PB * 64);
            if (k0 == 3u || k1 == 3u) cp_async_wait_all();   // a prefetched child has landed (a pushed one is already there)
            if (pf != kNone && !(flags & kStepPfLate)) prefetch(pf);
            const double *s0 = reinterpret_cast<const double *>(sl), *s1 = s0 + opsD;
            double *outp = hd.arena + (size_t)d.x * 32 + off;
            const bool store = (flags & kStepStore) != 0u, push = (flags & kStepPush) != 0u;
            if ((flags & (3u | kStepFirst)) == (2u | kStepFirst) && ((k0 == 1u && k1 >= 2u) || (k0 >= 2u && k1 == 2u))) {
                // the planner hands the children over in canonical order: registers, buffer, leaf
                if (k0 == 1u) {
                    if (k1 == 2u) step2<CT, CTH, 1, 2>(store, cur, buf, opOff, s0, s1, W, code0, code1, rs, outp);
                    else step2<CT, CTH, 1, 3>(store, cur, buf, opOff, s0, s1, W, code0, code1, rs, outp);
                } else if (k0 == 3u) step2<CT, CTH, 3, 2>(store, cur, buf, opOff, s0, s1, W, code0, code1, rs, outp);
                else step2<CT, CTH, 2, 2>(store, cur, buf, opOff, s0, s1, W, code0, code1, rs, outp);
            } else {
                // any other shape: one child, a continuation of a node with more than two children, children loaded
                // straight from global memory -- kinds decided at run time
                const unsigned nc = flags & 3u;
                const bool first = (flags & kStepFirst) != 0u;
                const Step2 *gs = a.steps + hd.stepBase + si;
                const double *m0 = hd.arena + (size_t)__ldg(&gs->c0) * 32 + off, *m1 = hd.arena + (size_t)__ldg(&gs->c1) * 32 + off;
                const unsigned c0x = code0 & 0xffu, c0y = (code0 >> 8) & 0xffu, c1x = code1 & 0xffu, c1y = (code1 >> 8) & 0xffu;
#pragma unroll
                for (int c = 0; c < CT; c++) {
                    const int cat = opOff + c;
                    double2 f[4];
                    if (k0 == 2u) child2<2, CTH>(c, cat, s0, W, c0x, c0y, cur, buf, m0, rs, f);
                    else if (k0 == 1u) child2<1, CTH>(c, cat, s0, W, c0x, c0y, cur, buf, m0, rs, f);
                    else if (k0 == 3u) child2<3, CTH>(c, cat, s0, W, c0x, c0y, cur, buf, m0, rs, f);
                    else child2<0, CTH>(c, cat, s0, W, c0x, c0y, cur, buf, m0, rs, f);
                    if (!first) {
#pragma unroll
                        for (int s4 = 0; s4 < 4; s4++) { f[s4].x = cur[c * 4 + s4].x * f[s4].x; f[s4].y = cur[c * 4 + s4].y * f[s4].y; }
                    }
                    if (nc == 2u) {
                        double2 g[4];
                        if (k1 == 2u) child2<2, CTH>(c, cat, s1, W, c1x, c1y, cur, buf, m1, rs, g);
                        else if (k1 == 1u) child2<1, CTH>(c, cat, s1, W, c1x, c1y, cur, buf, m1, rs, g);
                        else if (k1 == 3u) child2<3, CTH>(c, cat, s1, W, c1x, c1y, cur, buf, m1, rs, g);
                        else child2<0, CTH>(c, cat, s1, W, c1x, c1y, cur, buf, m1, rs, g);
#pragma unroll
                        for (int s4 = 0; s4 < 4; s4++) { f[s4].x *= g[s4].x; f[s4].y *= g[s4].y; }
                    }
#pragma unroll
                    for (int s4 = 0; s4 < 4; s4++) {
                        cur[c * 4 + s4] = f[s4];
                        if (store) st2(outp + (size_t)(c * 4 + s4) * rs, f[s4]);
                    }
                }
            }
            if (push) {      // a later step takes this node from the thread's shared-memory buffer
#pragma unroll
                for (int k = 0; k < KT; k++) buf[k * CTH] = cur[k];
            }
            if (pf != kNone && (flags & kStepPfLate)) prefetch(pf);   // the buffer was in use by this step: refill it now
        }
        // this warp has left the slot: every lane arrives for itself (CW x 32 arrivals per phase), so each lane's reads of
        // the slot are ordered before the refill by its own arrival
        mbar_arrive(empty + slot);
        if (lane == 0) service(si);
    }

    if (!hd.doLike) return;
    // ---- root reduction from registers (Pf/p4_tree.c:1029-1378) -------------------------------------------------
    // The sum over categories runs in category order as ONE fma chain, handed from warp to warp of the pattern
    // block when the categories are split: the result does not depend on CSPLIT.
    double2 A = make_double2(0.0, 0.0);
    uint64_t mask0 = ~0ull, mask1 = ~0ull;
    if (hd.rootTips) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (pat + h < a.nPat) {
                const int w = hd.rootTips[pat + h];
                uint64_t m = ~0ull;
                if (w < 4) m = 1ull << w;
                else if (w > 4) m = a.eqMask[w - 5];
                if (h) mask1 = m; else mask0 = m;
            }
        }
    }
#pragma unroll
    for (int g = 0; g < CSPLIT; g++) {
        if (cg == g) {
            if (g > 0) A = sA[pb * 32 + lane];
#pragma unroll
            for (int k = 0; k < KT; k++) {
                if ((mask0 >> (k & 3)) & 1ull) A.x = fma(hd.pi[k & 3], cur[k].x, A.x);
                if ((mask1 >> (k & 3)) & 1ull) A.y = fma(hd.pi[k & 3], cur[k].y, A.y);
            }
            if (g < CSPLIT - 1) sA[pb * 32 + lane] = A;
        }
        if (g < CSPLIT - 1) named_barrier(1, CTH);
    }
    double term = 0.0, bad = 0.0;
    if (cg == CSPLIT - 1) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int p1 = pat + h;
            if (p1 < a.nPat) {
                const uint64_t im = (hd.pInvar != 0.0 && a.invarMask) ? a.invarMask[p1] : 0ull;
                double t1 = 0.0, like = 0.0;
                if (like_term(h ? A.y : A.x, 0, hd.pInvar, NCAT, im, hd.pi, 4, a.counts[p1], &t1, &like)) term += t1;
                else bad += 1.0;
                if (hd.patLikes) hd.patLikes[p1] = like;
            }
        }
    }
    term = warpSum(term);
    bad = warpSum(bad);
    if (lane == 0) { sRed[warp] = term; sRed[CW + warp] = bad; }
    named_barrier(1, CTH);
    int *sLast = reinterpret_cast<int *>(sRed + 2 * CW);
    if (threadIdx.x == 0) {
        double t = 0.0, b = 0.0;
        for (int i = 0; i < CW; i++) { t += sRed[i]; b += sRed[CW + i]; }
        hd.partials[2 * blockIdx.x] = t;
        hd.partials[2 * blockIdx.x + 1] = b;
        __threadfence();
        const unsigned tk = atomicInc(hd.ticket, gridDim.x - 1);   // wraps to 0 with the last CTA: ready for the next launch
        *sLast = (tk == gridDim.x - 1) ? 1 : 0;
    }
    named_barrier(1, CTH);
    if (*sLast) {
        // the last CTA folds every CTA's partial in a fixed order (deterministic for a given launch shape)
        __threadfence();
        double t = 0.0, b = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += CTH) {
            t += __ldcg(hd.partials + 2 * i);
            b += __ldcg(hd.partials + 2 * i + 1);
        }
        t = warpSum(t);
        b = warpSum(b);
        named_barrier(1, CTH);          // sRed was read by thread 0 before the previous barrier
        if (lane == 0) { sRed[warp] = t; sRed[CW + warp] = b; }
        named_barrier(1, CTH);
        if (threadIdx.x == 0) {
            double tt = 0.0, bb = 0.0;
            for (int i = 0; i < CW; i++) { tt += sRed[i]; bb += sRed[CW + i]; }
            if (a.mail.world > 1) mail_allreduce(a.mail, blockIdx.y, tt, bb);   // the other shards' sums, over NVLink
            hd.result[0] = tt;
            hd.result[1] = bb;
        }
    }
}

}  // namespace p4b
