// tree_aa.cuh -- whole-tree CL recursion for 20-state parts, second generation (sm_100a, FP64 tensor cores).
//
// Same arithmetic as cl_tree_aa_kernel (kernels.cuh; Pf/p4_node.c:636-857 for 20 states): out^T = cl_child^T x P^T on
// mma.sync.m8n8k4 (DMMA) with the summation index dealt to the k-steps so that a node's result, as it leaves the tensor
// core, IS the next node's operand -- the running CL stays in the accumulator registers from step to step; 15 DMMAs per
// child and 8 patterns.  What changes is everything around the DMMAs, where the first generation spent 14 of every 15
// warp instructions:
//   * the step list arrives pre-decoded (Step2 records, the host's planner of tree_dna.cuh) and is copied into shared
//     memory once: a step's descriptor is one 16-byte shared-memory load, not a walk through runtime-indexed constant
//     memory; the children come in canonical order (registers, memory, leaf) and the step body is straight-line code
//     for one combination of kinds;
//   * operands -- P^T fragments or the transposed leaf table of all NCAT categories of a child in ONE bulk copy (TMA),
//     plus the CTA's tip codes of a leaf child -- arrive in a ring of slots shared by the CTA, each completing on its
//     mbarrier; the warps refill the slots in turn (warp w the steps w, w + NW, ...): no barrier in the step loop, no
//     thread that stages for the others;
//   * the transposed leaf tables are padded to 24 states with zeros, so the lanes that hold the padding states 20..23
//     read their zeros like everybody else: no per-lane selects.
// A warp owns one rate category of 16 patterns (two m-tiles; a lane's two patterns are one 16-byte access); a CTA is
// NCAT x GROUPS warps.  The root reduction stays like_kernel (+ the last block's fold and the shard exchange).
#pragma once
#include <type_traits>

#include "tree_dna.cuh"

namespace p4b {

constexpr int kAA2TblStates = 24;      // states per code in the transposed leaf table of the operand deck (20 + 4 zeros)
__host__ __device__ inline size_t aa2CatDoubles(int W) { return (size_t)W * kAA2TblStates > (size_t)kAAFrag ? (size_t)W * kAA2TblStates : (size_t)kAAFrag; }
// ring slot: [child 0: NCAT x catDoubles][child 1: ...][tips child 0: GROUPS*16 B][tips child 1]
__host__ __device__ inline size_t aa2SlotBytes(int nCat, int W, int GROUPS) { return 2 * (size_t)nCat * aa2CatDoubles(W) * 8 + 2 * (size_t)GROUPS * 16; }
__host__ inline size_t aa2SmemBytes(int nCat, int W, int GROUPS, int RING, int maxSteps)
{
    return treeDna2StepBytes(maxSteps) + RING * aa2SlotBytes(nCat, W, GROUPS) + RING * 8 + RING * 4 + 32;
}

struct TreeArgsAA2 {
    int ps, nPat, tblW, nTrees;
    int maxSteps, pad0;       // pad0: measurement switch (no stores)
    int nCat, pad1;
    const uint8_t *tips;      // part's tip rows [nTax][ps]
    const Step2 *steps;       // n0 / n1: offset (doubles) of the child's operands from hdr.aux -- P^T fragments or transposed leaf tables of category 0
    struct Hdr {
        double *arena;
        const double *aux;    // tree's operand decks, already offset to this part
        int stepBase, nSteps;
    } hdr[kMaxBatchTrees];
};

template <int NCAT, int GROUPS, int RING, int MINB>
__global__ void __launch_bounds__(NCAT * GROUPS * 32, MINB)
cl_tree_aa2_kernel(const __grid_constant__ TreeArgsAA2 a)
{
    constexpr int DIM = 20, MT = 2, NW = NCAT * GROUPS, THREADS = NW * 32;
    const TreeArgsAA2::Hdr &hd = a.hdr[blockIdx.y];
    extern __shared__ __align__(16) unsigned char smraw[];
    const int W = a.tblW, nSteps = hd.nSteps;
    const unsigned catD = (unsigned)aa2CatDoubles(W);            // doubles per category inside a child's slot half
    const unsigned childB = NCAT * catD * 8;                      // bytes of one child's operands in a slot
    const unsigned slotB = (unsigned)aa2SlotBytes(NCAT, W, GROUPS);
    uint4 *sSteps = reinterpret_cast<uint4 *>(smraw);
    uint2 *sNodes = reinterpret_cast<uint2 *>(sSteps + a.maxSteps);
    unsigned char *ring = smraw + treeDna2StepBytes(a.maxSteps);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + RING * slotB);
    unsigned *cnt = reinterpret_cast<unsigned *>(full + RING);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    const int cat = warp / GROUPS, grp = warp % GROUPS;
    const size_t ps = (size_t)a.ps;
    const int pat0 = (blockIdx.x * GROUPS + grp) * (8 * MT);
    const size_t rowBase = (size_t)cat * DIM * ps + pat0 + MT * g;       // + state * ps: this lane's two patterns of a row
    const uint8_t *ctaTips = a.tips + (size_t)blockIdx.x * (GROUPS * 16);
    const bool tail = q >= 2;                                            // n-tile 2 of these lanes holds the padding states 20..23

    // one lane: the bulk copies of step j's operands (all categories of each child at once) into its ring slot
    auto produce = [&](int j) {
        const int slot = j & (RING - 1);
        const uint4 dg = sSteps[j];
        const uint2 nn = sNodes[j];
        const unsigned flags = dg.z, nc = flags & 3u, k0 = (flags >> 4) & 3u, k1 = (flags >> 6) & 3u;
        const bool l0 = k0 == 2u, l1 = nc == 2u && k1 == 2u;
        const unsigned fBytes = NCAT * kAAFrag * 8, tBytes = (unsigned)(NCAT * W * kAA2TblStates * 8), tipBytes = GROUPS * 16;
        unsigned char *sl = ring + slot * slotB;
        mbar_expect_tx(full + slot, (l0 ? tBytes + tipBytes : fBytes) + (nc == 2u ? (l1 ? tBytes + tipBytes : fBytes) : 0u));
        bulk_g2s(sl, hd.aux + nn.x, l0 ? tBytes : fBytes, full + slot);
        if (l0) bulk_g2s(sl + 2 * childB, ctaTips + (size_t)(dg.w & 0xffffu) * ps, tipBytes, full + slot);
        if (nc == 2u) {
            bulk_g2s(sl + childB, hd.aux + nn.y, l1 ? tBytes : fBytes, full + slot);
            if (l1) bulk_g2s(sl + 2 * childB + tipBytes, ctaTips + (size_t)(dg.w >> 16) * ps, tipBytes, full + slot);
        }
    };
    {
        const Step2 *gSteps = a.steps + hd.stepBase;
        for (int i = threadIdx.x; i < nSteps; i += THREADS) {
            const uint4 dA = __ldg(reinterpret_cast<const uint4 *>(gSteps + i)), dB = __ldg(reinterpret_cast<const uint4 *>(gSteps + i) + 1);
            // {out, CL buffer of the child that is read from memory (child 0 if it is one, else child 1), flags, tip rows}
            sSteps[i] = make_uint4(dA.x, ((dA.y >> 4) & 3u) == 0u ? dA.z : dA.w, dA.y, (dB.x & 0xffffu) | (dB.y << 16));
            sNodes[i] = make_uint2(__ldg(&gSteps[i].n0), __ldg(&gSteps[i].n1));
        }
        if (threadIdx.x == 0) {
            for (int i = 0; i < RING; i++) { mbar_init(full + i, 1); cnt[i] = 0u; }
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int j = 0; j < RING && j < nSteps; j++) produce(j);

    // acc (=|*=) A x B for one internal child: A in the C layout of the child (a4 = the fifth k-step's operand: state
    // 16 + q), B = this category's P^T fragments.  The three n-tiles advance together: 6 independent accumulator chains.
    auto contract = [&](const double (&A)[MT][3][2], const double (&a4)[MT], const double *__restrict__ Bc, double (&out)[MT][3][2], bool assign) {
        double acc[MT][3][2];
#pragma unroll
        for (int j = 0; j < MT; j++)
#pragma unroll
            for (int nt = 0; nt < 3; nt++) acc[j][nt][0] = acc[j][nt][1] = 0.0;
#pragma unroll
        for (int kk = 0; kk < 5; kk++) {
#pragma unroll
            for (int nt = 0; nt < 3; nt++) {
                const double b = Bc[(kk * 3 + nt) * 32];
#pragma unroll
                for (int j = 0; j < MT; j++) dmma884(acc[j][nt][0], acc[j][nt][1], kk < 4 ? A[j][kk >> 1][kk & 1] : a4[j], b);
            }
        }
#pragma unroll
        for (int j = 0; j < MT; j++)
#pragma unroll
            for (int nt = 0; nt < 3; nt++) {
                if (assign) { out[j][nt][0] = acc[j][nt][0]; out[j][nt][1] = acc[j][nt][1]; }
                else { out[j][nt][0] *= acc[j][nt][0]; out[j][nt][1] *= acc[j][nt][1]; }
            }
    };
    const int quadSrc = (lane & ~3) | (q >> 1);   // the lane of this quad that holds state 16 + q (as its element q & 1 of n-tile 2)

    // an internal child's CL from global memory (written earlier by this same lane), in the operand layout
    auto loadSib = [&](unsigned slotCode, double (&sib)[MT][3][2], double (&a4)[MT]) {
        const double *cl = hd.arena + (size_t)slotCode * 32 + rowBase;
#pragma unroll
        for (int r = 0; r < 4; r++) {            // states 8t + 2q + i, t = r >> 1, i = r & 1
            const double2 v = ld2(cl + (size_t)(8 * (r >> 1) + 2 * q + (r & 1)) * ps);
            sib[0][r >> 1][r & 1] = v.x;
            sib[1][r >> 1][r & 1] = v.y;
        }
        const double2 v = ld2(cl + (size_t)(16 + q) * ps);    // state 16 + q: the fifth k-step's operand, straight from its row
        a4[0] = v.x;
        a4[1] = v.y;
#pragma unroll
        for (int j = 0; j < MT; j++) sib[j][2][0] = sib[j][2][1] = 0.0;   // not read
    };
    // the factor of one child folded into `out`; KIND 1: the child is `in` (registers), 0: its CL is in global memory, 2: leaf
    auto child = [&](auto kindTag, bool assign, const double (&in)[MT][3][2], double (&out)[MT][3][2], const unsigned char *opsB, unsigned code, unsigned slotCode) {
        constexpr int KIND = decltype(kindTag)::value;
        const double *ops = reinterpret_cast<const double *>(opsB) + (size_t)cat * (KIND == 2 ? (unsigned)(W * kAA2TblStates) : (unsigned)kAAFrag);
        if (KIND == 2) {
            // leaf: out (=|*=) T[code][state]; the table is [code][24], so the lane's states 8t+2q, 8t+2q+1 are one 16-byte load
            const double *T = ops + 2 * q;
#pragma unroll
            for (int j = 0; j < MT; j++) {
                const double *Tj = T + ((code >> (8 * j)) & 0xffu) * kAA2TblStates;
#pragma unroll
                for (int t = 0; t < 3; t++) {
                    const double2 v = *reinterpret_cast<const double2 *>(Tj + 8 * t);
                    if (assign) { out[j][t][0] = v.x; out[j][t][1] = v.y; }
                    else { out[j][t][0] *= v.x; out[j][t][1] *= v.y; }
                }
            }
        } else if (KIND == 1) {
            double a4[MT];
#pragma unroll
            for (int j = 0; j < MT; j++) {
                const double v0 = __shfl_sync(0xffffffffu, in[j][2][0], quadSrc), v1 = __shfl_sync(0xffffffffu, in[j][2][1], quadSrc);
                a4[j] = (q & 1) ? v1 : v0;
            }
            contract(in, a4, ops + lane, out, assign);
        } else {
            double sib[MT][3][2], a4[MT];
            loadSib(slotCode, sib, a4);
            contract(sib, a4, ops + lane, out, assign);
        }
    };
    auto storeCL = [&](unsigned slotCode, const double (&out)[MT][3][2]) {
        double *o = hd.arena + (size_t)slotCode * 32 + rowBase + (size_t)(2 * q) * ps;
#pragma unroll
        for (int r = 0; r < 6; r++) {
            if (r < 4 || !tail) st2(o + (size_t)(8 * (r >> 1) + (r & 1)) * ps, make_double2(out[0][r >> 1][r & 1], out[1][r >> 1][r & 1]));
        }
    };

    int myNext = warp;      // ring refills are shared out: warp w refills the slots of the steps w, w + NW, ... (see tree_dna.cuh)
    auto service = [&](int done) {
        while (myNext <= done && myNext + RING < nSteps) {
            volatile unsigned *c = cnt + (myNext & (RING - 1));
            if (*c != (unsigned)NW) break;
            *c = 0u;
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            produce(myNext + RING);
            myNext += NW;
        }
    };

    // One step: `in` holds the CL of the node computed by the previous step, `out` receives this node's.
    auto step = [&](int si, const double (&in)[MT][3][2], double (&out)[MT][3][2]) {
        const int slot = si & (RING - 1);
        {
            const unsigned parity = (unsigned)(si / RING) & 1u;
            int spins = 0;
            while (!mbar_test(full + slot, parity)) {
                if (lane == 0) service(si - 1);
                if (++spins > (1 << 22)) __trap();
            }
        }
        const uint4 d = sSteps[si];
        const unsigned flags = d.z, nc = flags & 3u, k0 = (flags >> 4) & 3u, k1 = (flags >> 6) & 3u;
        const unsigned char *sl = ring + slot * slotB;
        const unsigned tipOff = 2 * childB + (unsigned)(grp * 16 + MT * g);
        const unsigned code0 = *reinterpret_cast<const unsigned short *>(sl + tipOff);
        const unsigned code1 = *reinterpret_cast<const unsigned short *>(sl + tipOff + GROUPS * 16);
        const std::integral_constant<int, 0> MEM;
        const std::integral_constant<int, 1> REG;
        const std::integral_constant<int, 2> LEAF;
        if ((flags & (3u | kStepFirst)) == (2u | kStepFirst)) {
            // two children in canonical order (registers, memory, leaf): straight-line code per combination
            if (k0 == 1u) {
                if (k1 == 2u) { child(REG, true, in, out, sl, 0u, 0u); child(LEAF, false, in, out, sl + childB, code1, 0u); }
                else {       // the sibling's rows are requested before the contraction of the child in registers: their latency hides behind it
                    double sib[MT][3][2], a4[MT];
                    loadSib(d.y, sib, a4);
                    child(REG, true, in, out, sl, 0u, 0u);
                    contract(sib, a4, reinterpret_cast<const double *>(sl + childB) + (size_t)cat * kAAFrag + lane, out, false);
                }
            } else if (k0 == 0u) {
                if (k1 == 2u) { child(MEM, true, in, out, sl, 0u, d.y); child(LEAF, false, in, out, sl + childB, code1, 0u); }
                else { child(MEM, true, in, out, sl, 0u, d.y); child(MEM, false, in, out, sl + childB, 0u, __ldg(&(a.steps + hd.stepBase + si)->c1)); }
            } else { child(LEAF, true, in, out, sl, code0, 0u); child(LEAF, false, in, out, sl + childB, code1, 0u); }
        } else {
            // one child, or the continuation of a node with more than two children
            const bool first = (flags & kStepFirst) != 0u;
            if (!first) {
#pragma unroll
                for (int j = 0; j < MT; j++)
#pragma unroll
                    for (int t = 0; t < 3; t++) { out[j][t][0] = in[j][t][0]; out[j][t][1] = in[j][t][1]; }
            }
            if (k0 == 1u) child(REG, first, in, out, sl, 0u, 0u);
            else if (k0 == 0u) child(MEM, first, in, out, sl, 0u, d.y);
            else child(LEAF, first, in, out, sl, code0, 0u);
            if (nc == 2u) {
                if (k1 == 0u) child(MEM, false, in, out, sl + childB, 0u, k0 == 0u ? __ldg(&(a.steps + hd.stepBase + si)->c1) : d.y);
                else child(LEAF, false, in, out, sl + childB, code1, 0u);
            }
        }
        if ((flags & kStepStore) && !a.pad0) storeCL(d.x, out);
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            atomicAdd(cnt + slot, 1u);
            service(si);
        }
    };

    double cA[MT][3][2], cB[MT][3][2];
#pragma unroll
    for (int j = 0; j < MT; j++)
#pragma unroll
        for (int t = 0; t < 3; t++) cA[j][t][0] = cA[j][t][1] = cB[j][t][0] = cB[j][t][1] = 0.0;
    for (int si = 0; si < nSteps; si += 2) {
        step(si, cA, cB);
        if (si + 1 < nSteps) step(si + 1, cB, cA);
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Third generation: one rate category per CTA.
//
// A CTA works on ONE rate category (blockIdx.z) of GROUPS x 16 patterns: a ring slot holds 2 x 3.8 KB instead of
// 2 x 15.4 KB -- a deeper ring in less shared memory, several independent CTAs per SM -- and any number of categories
// is served.  Same arithmetic, same fragment scheme, same results as the second generation.
// (Measured and dropped: handing the node's 20 x 16 tile to the bulk-copy engine through shared memory -- the
// fence.proxy.async every lane needs between its shared-memory stores and the copy costs more than the stores'
// address translation it was meant to take off the warp's path: 5.0 ms against 3.7 ms without any store.)
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t aa3SlotBytes(int W, int GROUPS) { return 2 * aa2CatDoubles(W) * 8 + 2 * (size_t)GROUPS * 16; }
__host__ inline size_t aa3SmemBytes(int W, int GROUPS, int RING, int maxSteps)
{
    return treeDna2StepBytes(maxSteps) + RING * aa3SlotBytes(W, GROUPS) + RING * 8 + RING * 8 + 32;
}

// PM: which two of the warp's 16 patterns a lane owns -- 0: the adjacent patterns 2g, 2g+1 (one 16-byte access per row);
// 1: the patterns g and g+8 (two 8-byte accesses per row, each straight from / into the accumulator registers: no packing
// moves, and a store's source registers are not written again for a whole step)
template <int GROUPS, int RING, int MINB, int PM>
__global__ void __launch_bounds__(GROUPS * 32, MINB)
cl_tree_aa3_kernel(const __grid_constant__ TreeArgsAA2 a)
{
    constexpr int DIM = 20, MT = 2, NW = GROUPS;
    const TreeArgsAA2::Hdr &hd = a.hdr[blockIdx.y];
    extern __shared__ __align__(16) unsigned char smraw[];
    const int W = a.tblW, nSteps = hd.nSteps;
    const int cat = blockIdx.z;
    const unsigned catD = (unsigned)aa2CatDoubles(W);
    const unsigned childB = catD * 8;                               // bytes of one child's operands (this category) in a slot
    const unsigned slotB = (unsigned)aa3SlotBytes(W, GROUPS);
    const unsigned opsStride = (unsigned)cat * kAAFrag, tblStride = (unsigned)(cat * W * kAA2TblStates);
    uint4 *sSteps = reinterpret_cast<uint4 *>(smraw);
    uint2 *sNodes = reinterpret_cast<uint2 *>(sSteps + a.maxSteps);
    unsigned char *ring = smraw + treeDna2StepBytes(a.maxSteps);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + RING * slotB);
    uint64_t *empty = full + RING;      // arrivals of the warps that have left the slot (see mbar_arrive, tree_dna.cuh)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    const size_t ps = (size_t)a.ps;
    const int pat0 = (blockIdx.x * GROUPS + warp) * (8 * MT);
    const size_t rowBase = (size_t)cat * DIM * ps + pat0 + (PM ? g : MT * g);       // + state * ps: this lane's (first) pattern of a row
    const uint8_t *ctaTips = a.tips + (size_t)blockIdx.x * (GROUPS * 16);
    const bool tail = q >= 2;                                            // n-tile 2 of these lanes holds the padding states 20..23

    auto produce = [&](int j) {
        const int slot = j % RING;
        const uint4 dg = sSteps[j];
        const uint2 nn = sNodes[j];
        const unsigned flags = dg.z, nc = flags & 3u, k0 = (flags >> 4) & 3u, k1 = (flags >> 6) & 3u;
        const bool l0 = k0 == 2u, l1 = nc == 2u && k1 == 2u;
        const unsigned fBytes = kAAFrag * 8, tBytes = (unsigned)(W * kAA2TblStates * 8), tipBytes = GROUPS * 16;
        unsigned char *sl = ring + slot * slotB;
        mbar_expect_tx(full + slot, (l0 ? tBytes + tipBytes : fBytes) + (nc == 2u ? (l1 ? tBytes + tipBytes : fBytes) : 0u));
        bulk_g2s(sl, hd.aux + nn.x + (l0 ? tblStride : opsStride), l0 ? tBytes : fBytes, full + slot);
        if (l0) bulk_g2s(sl + 2 * childB, ctaTips + (size_t)(dg.w & 0xffffu) * ps, tipBytes, full + slot);
        if (nc == 2u) {
            bulk_g2s(sl + childB, hd.aux + nn.y + (l1 ? tblStride : opsStride), l1 ? tBytes : fBytes, full + slot);
            if (l1) bulk_g2s(sl + 2 * childB + tipBytes, ctaTips + (size_t)(dg.w >> 16) * ps, tipBytes, full + slot);
        }
    };
    {
        const Step2 *gSteps = a.steps + hd.stepBase;
        for (int i = threadIdx.x; i < nSteps; i += GROUPS * 32) {
            const uint4 dA = __ldg(reinterpret_cast<const uint4 *>(gSteps + i)), dB = __ldg(reinterpret_cast<const uint4 *>(gSteps + i) + 1);
            // {out, CL buffer of the child that is read from memory (child 0 if it is one, else child 1), flags, tip rows}
            sSteps[i] = make_uint4(dA.x, ((dA.y >> 4) & 3u) == 0u ? dA.z : dA.w, dA.y, (dB.x & 0xffffu) | (dB.y << 16));
            sNodes[i] = make_uint2(__ldg(&gSteps[i].n0), __ldg(&gSteps[i].n1));
        }
        if (threadIdx.x == 0) {
            for (int i = 0; i < RING; i++) { mbar_init(full + i, 1); mbar_init(empty + i, NW); }
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int j = 0; j < RING && j < nSteps; j++) produce(j);

    auto contract = [&](const double (&A)[MT][3][2], const double (&a4)[MT], const double *__restrict__ Bc, double (&out)[MT][3][2], bool assign) {
        double acc[MT][3][2];
#pragma unroll
        for (int j = 0; j < MT; j++)
#pragma unroll
            for (int nt = 0; nt < 3; nt++) acc[j][nt][0] = acc[j][nt][1] = 0.0;
#pragma unroll
        for (int kk = 0; kk < 5; kk++) {
#pragma unroll
            for (int nt = 0; nt < 3; nt++) {
                const double b = Bc[(kk * 3 + nt) * 32];
#pragma unroll
                for (int j = 0; j < MT; j++) dmma884(acc[j][nt][0], acc[j][nt][1], kk < 4 ? A[j][kk >> 1][kk & 1] : a4[j], b);
            }
        }
#pragma unroll
        for (int j = 0; j < MT; j++)
#pragma unroll
            for (int nt = 0; nt < 3; nt++) {
                if (assign) { out[j][nt][0] = acc[j][nt][0]; out[j][nt][1] = acc[j][nt][1]; }
                else { out[j][nt][0] *= acc[j][nt][0]; out[j][nt][1] *= acc[j][nt][1]; }
            }
    };
    const int quadSrc = (lane & ~3) | (q >> 1);   // the lane of this quad that holds state 16 + q (as its element q & 1 of n-tile 2)

    auto loadSib = [&](unsigned slotCode, double (&sib)[MT][3][2], double (&a4)[MT]) {
        const double *cl = hd.arena + (size_t)slotCode * 32 + rowBase;
#pragma unroll
        for (int r = 0; r < 4; r++) {            // states 8t + 2q + i, t = r >> 1, i = r & 1
            const double *row = cl + (size_t)(8 * (r >> 1) + 2 * q + (r & 1)) * ps;
            if (PM) { sib[0][r >> 1][r & 1] = __ldg(row); sib[1][r >> 1][r & 1] = __ldg(row + 8); }
            else { const double2 v = ld2(row); sib[0][r >> 1][r & 1] = v.x; sib[1][r >> 1][r & 1] = v.y; }
        }
        {
            const double *row = cl + (size_t)(16 + q) * ps;    // state 16 + q: the fifth k-step's operand, straight from its row
            if (PM) { a4[0] = __ldg(row); a4[1] = __ldg(row + 8); }
            else { const double2 v = ld2(row); a4[0] = v.x; a4[1] = v.y; }
        }
#pragma unroll
        for (int j = 0; j < MT; j++) sib[j][2][0] = sib[j][2][1] = 0.0;   // not read
    };
    auto child = [&](auto kindTag, bool assign, const double (&in)[MT][3][2], double (&out)[MT][3][2], const unsigned char *opsB, unsigned code, unsigned slotCode) {
        constexpr int KIND = decltype(kindTag)::value;
        const double *ops = reinterpret_cast<const double *>(opsB);
        if (KIND == 2) {
            const double *T = ops + 2 * q;
#pragma unroll
            for (int j = 0; j < MT; j++) {
                const double *Tj = T + ((code >> (8 * j)) & 0xffu) * kAA2TblStates;
#pragma unroll
                for (int t = 0; t < 3; t++) {
                    const double2 v = *reinterpret_cast<const double2 *>(Tj + 8 * t);
                    if (assign) { out[j][t][0] = v.x; out[j][t][1] = v.y; }
                    else { out[j][t][0] *= v.x; out[j][t][1] *= v.y; }
                }
            }
        } else if (KIND == 1) {
            double a4[MT];
#pragma unroll
            for (int j = 0; j < MT; j++) {
                const double v0 = __shfl_sync(0xffffffffu, in[j][2][0], quadSrc), v1 = __shfl_sync(0xffffffffu, in[j][2][1], quadSrc);
                a4[j] = (q & 1) ? v1 : v0;
            }
            contract(in, a4, ops + lane, out, assign);
        } else {
            double sib[MT][3][2], a4[MT];
            loadSib(slotCode, sib, a4);
            contract(sib, a4, ops + lane, out, assign);
        }
    };
    auto storeCL = [&](unsigned slotCode, const double (&out)[MT][3][2]) {
        double *o = hd.arena + (size_t)slotCode * 32 + rowBase + (size_t)(2 * q) * ps;
#pragma unroll
        for (int r = 0; r < 6; r++) {
            if (r < 4 || !tail) {
                double *row = o + (size_t)(8 * (r >> 1) + (r & 1)) * ps;
                if (PM) { row[0] = out[0][r >> 1][r & 1]; row[8] = out[1][r >> 1][r & 1]; }
                else st2(row, make_double2(out[0][r >> 1][r & 1], out[1][r >> 1][r & 1]));
            }
        }
    };

    int myNext = warp;
    auto service = [&](int done) {
        while (myNext <= done && myNext + RING < nSteps) {
            if (!mbar_poll(empty + (myNext % RING), (unsigned)(myNext / RING) & 1u)) break;
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            produce(myNext + RING);
            myNext += NW;
        }
    };

    auto step = [&](int si, const double (&in)[MT][3][2], double (&out)[MT][3][2]) {
        const int slot = si % RING;
        {
            const unsigned parity = (unsigned)(si / RING) & 1u;
            int spins = 0;
            while (!mbar_test(full + slot, parity)) {
                if (lane == 0) service(si - 1);
                if (++spins > (1 << 22)) __trap();
            }
        }
        const uint4 d = sSteps[si];
        const unsigned flags = d.z, nc = flags & 3u, k0 = (flags >> 4) & 3u, k1 = (flags >> 6) & 3u;
        const unsigned char *sl = ring + slot * slotB;
        const unsigned tipOff = 2 * childB + (unsigned)(warp * 16 + (PM ? g : MT * g));
        unsigned code0, code1;      // the lane's two tip codes of each leaf child, one per byte
        if (PM) {
            code0 = (unsigned)sl[tipOff] | ((unsigned)sl[tipOff + 8] << 8);
            code1 = (unsigned)sl[tipOff + GROUPS * 16] | ((unsigned)sl[tipOff + GROUPS * 16 + 8] << 8);
        } else {
            code0 = *reinterpret_cast<const unsigned short *>(sl + tipOff);
            code1 = *reinterpret_cast<const unsigned short *>(sl + tipOff + GROUPS * 16);
        }
        const std::integral_constant<int, 0> MEM;
        const std::integral_constant<int, 1> REG;
        const std::integral_constant<int, 2> LEAF;
        if ((flags & (3u | kStepFirst)) == (2u | kStepFirst)) {
            if (k0 == 1u) {
                if (k1 == 2u) { child(REG, true, in, out, sl, 0u, 0u); child(LEAF, false, in, out, sl + childB, code1, 0u); }
                else {
                    double sib[MT][3][2], a4[MT];
                    loadSib(d.y, sib, a4);
                    child(REG, true, in, out, sl, 0u, 0u);
                    contract(sib, a4, reinterpret_cast<const double *>(sl + childB) + lane, out, false);
                }
            } else if (k0 == 0u) {
                if (k1 == 2u) { child(MEM, true, in, out, sl, 0u, d.y); child(LEAF, false, in, out, sl + childB, code1, 0u); }
                else { child(MEM, true, in, out, sl, 0u, d.y); child(MEM, false, in, out, sl + childB, 0u, __ldg(&(a.steps + hd.stepBase + si)->c1)); }
            } else { child(LEAF, true, in, out, sl, code0, 0u); child(LEAF, false, in, out, sl + childB, code1, 0u); }
        } else {
            const bool first = (flags & kStepFirst) != 0u;
            if (!first) {
#pragma unroll
                for (int j = 0; j < MT; j++)
#pragma unroll
                    for (int t = 0; t < 3; t++) { out[j][t][0] = in[j][t][0]; out[j][t][1] = in[j][t][1]; }
            }
            if (k0 == 1u) child(REG, first, in, out, sl, 0u, 0u);
            else if (k0 == 0u) child(MEM, first, in, out, sl, 0u, d.y);
            else child(LEAF, first, in, out, sl, code0, 0u);
            if (nc == 2u) {
                if (k1 == 0u) child(MEM, false, in, out, sl + childB, 0u, k0 == 0u ? __ldg(&(a.steps + hd.stepBase + si)->c1) : d.y);
                else child(LEAF, false, in, out, sl + childB, code1, 0u);
            }
        }
        if ((flags & kStepStore) && !a.pad0) storeCL(d.x, out);
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(empty + slot);
            service(si);
        }
    };

    double cA[MT][3][2], cB[MT][3][2];
#pragma unroll
    for (int j = 0; j < MT; j++)
#pragma unroll
        for (int t = 0; t < 3; t++) cA[j][t][0] = cA[j][t][1] = cB[j][t][0] = cB[j][t][1] = 0.0;
    for (int si = 0; si < nSteps; si += 2) {
        step(si, cA, cB);
        if (si + 1 < nSteps) step(si + 1, cB, cA);
    }
}

}  // namespace p4b
