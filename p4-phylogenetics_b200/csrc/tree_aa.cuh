// tree_aa.cuh -- whole-tree CL recursion for 20-state parts (sm_100a, FP64 tensor cores).
//
// What Pf/p4_node.c:636-857 computes for every node of a step list, for dim = 20: out^T = cl_child^T x P^T on
// mma.sync.m8n8k4 (DMMA) with the summation index dealt to the k-steps so that a node's result, as it leaves the
// tensor core, IS the next node's operand -- the running CL stays in the accumulator registers from step to step;
// 15 DMMAs per child and 8 patterns.  Around the DMMAs (the machinery of tree_dna.cuh):
//   * the step list arrives pre-decoded (Step2 records, the host's planner) and is copied into shared memory once;
//     the children come in canonical order (registers, memory, leaf), the step body is straight-line code per
//     combination of kinds;
//   * a CTA works on ONE rate category (blockIdx.z) of GROUPS x 16 patterns -- any number of categories is served;
//     a step's operands (this category's P^T fragments or transposed leaf table, 3.8 KB per child, plus the CTA's tip
//     codes of a leaf child) arrive by bulk copies (TMA) in a ring of slots, each completing on its "full" mbarrier;
//     the warps leave a slot by arriving on its "empty" mbarrier and refill the slots in turn (warp w the steps
//     w, w + NW, ...): no barrier in the step loop, no thread that stages for the others; two independent CTAs per SM;
//   * the transposed leaf tables are padded to 24 states with zeros: the lanes that hold the padding states 20..23
//     read their zeros like everybody else;
//   * stores: the warp transposes its 20 x 16 tile with one fixed lane permutation (24 shuffles) so that eight
//     consecutive lanes write one 128-byte line (see storeCL).
// History, measured on cfg 3 (100 taxa x 200 k patterns, LG+G4): first generation (runtime-indexed step decode, generic
// child loops, named barriers) 5.17 ms; second (4 categories per CTA, 16 warps, ring of 4 x 31 KB) 4.53 ms; this one
// 3.64 ms.  Dropped on the way: handing the tile to the bulk-copy engine through shared memory (the fence.proxy.async
// between the shared-memory stores and the copy costs more than the stores: 5.0 ms), fewer registers for more warps
// (spills: 4.5 ms at 96, 4.6 ms at 80).
// The root reduction stays like_kernel (+ the last block's fold and the shard exchange).
#pragma once
#include <type_traits>

#include "tree_dna.cuh"

namespace p4b {

constexpr int kAATblStates = 24;      // states per code in the transposed leaf table of the operand deck (20 + 4 zeros)
__host__ __device__ inline size_t aaCatDoubles(int W) { return (size_t)W * kAATblStates > (size_t)kAAFrag ? (size_t)W * kAATblStates : (size_t)kAAFrag; }
// ring slot: [child 0: catDoubles][child 1: catDoubles][tips child 0: GROUPS*16 B][tips child 1]

// One launch serves up to kMaxBatchTrees (tree, part) pairs (blockIdx.y): the trees of a batched MCMC evaluation of one part, or
// the 20-state parts of one tree -- each with its own data part, pattern count, leaf-table width and number of categories; the grid
// is sized for the largest, CTAs beyond a pair's own extent leave at once.
struct TreeArgsAA {
    int nTrees, maxSteps;
    int pad0, pad1;           // pad0: measurement switch (no stores)
    const Step2 *steps;       // n0 / n1: offset (doubles) of the child's operands from hdr.aux -- P^T fragments or transposed leaf tables of category 0
    struct Hdr {
        double *arena;
        const double *aux;    // tree's operand decks, already offset to this part
        const uint8_t *tips;  // part's tip rows [nTax][ps]
        int stepBase, nSteps;
        int ps, tblW;         // pattern stride of the part's shard; width of its leaf tables
        int nCat, nBlocks;    // rate categories; CTAs along x that have patterns (ps / (GROUPS * 16))
    } hdr[kMaxBatchTrees];
};

__host__ __device__ inline size_t aaSlotBytes(int W, int GROUPS) { return 2 * aaCatDoubles(W) * 8 + 2 * (size_t)GROUPS * 16; }
__host__ inline size_t aaSmemBytes(int W, int GROUPS, int RING, int maxSteps)
{
    return treeDna2StepBytes(maxSteps) + RING * aaSlotBytes(W, GROUPS) + RING * 8 + RING * 8 + 32;
}

// A warp owns 16 patterns of the CTA's category (two m-tiles); lane (g, q) holds the adjacent patterns 2g, 2g+1 (one
// 16-byte access per row) of the states 8t + 2q + i.
template <int GROUPS, int RING, int MINB>
__global__ void __launch_bounds__(GROUPS * 32, MINB)
cl_tree_aa_kernel(const __grid_constant__ TreeArgsAA a)
{
    constexpr int DIM = 20, MT = 2, NW = GROUPS;
    const TreeArgsAA::Hdr &hd = a.hdr[blockIdx.y];
    extern __shared__ __align__(16) unsigned char smraw[];
    if ((int)blockIdx.x >= hd.nBlocks || (int)blockIdx.z >= hd.nCat) return;
    const int W = hd.tblW, nSteps = hd.nSteps;
    const int cat = blockIdx.z;
    const unsigned catD = (unsigned)aaCatDoubles(W);
    const unsigned childB = catD * 8;                               // bytes of one child's operands (this category) in a slot
    const unsigned slotB = (unsigned)aaSlotBytes(W, GROUPS);
    const unsigned opsStride = (unsigned)cat * kAAFrag, tblStride = (unsigned)(cat * W * kAATblStates);
    uint4 *sSteps = reinterpret_cast<uint4 *>(smraw);
    uint2 *sNodes = reinterpret_cast<uint2 *>(sSteps + a.maxSteps);
    unsigned char *ring = smraw + treeDna2StepBytes(a.maxSteps);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + RING * slotB);
    uint64_t *empty = full + RING;      // arrivals of the lanes that have left the slot (see mbar_arrive, tree_dna.cuh)
    // the thread index through a shuffle: an opaque copy, which keeps ptxas from re-reading the special register (S2R, a
    // 25-cycle scoreboard wait) wherever the step loop needs the lane under register pressure
    const unsigned tidx = __shfl_sync(0xffffffffu, threadIdx.x, threadIdx.x & 31u);
    const int lane = tidx & 31, warp = tidx >> 5, g = lane >> 2, q = lane & 3;
    const size_t ps = (size_t)hd.ps;
    const int pat0 = (blockIdx.x * GROUPS + warp) * (8 * MT);
    const size_t rowBase = (size_t)cat * DIM * ps + pat0 + MT * g;       // + state * ps: this lane's two patterns of a row
    const uint8_t *ctaTips = hd.tips + (size_t)blockIdx.x * (GROUPS * 16);

    auto produce = [&](int j) {
        const int slot = j % RING;
        const uint4 dg = sSteps[j];
        const uint2 nn = sNodes[j];
        const unsigned flags = dg.z, nc = flags & 3u, k0 = (flags >> 4) & 3u, k1 = (flags >> 6) & 3u;
        const bool l0 = k0 == 2u, l1 = nc == 2u && k1 == 2u;
        const unsigned fBytes = kAAFrag * 8, tBytes = (unsigned)(W * kAATblStates * 8), tipBytes = GROUPS * 16;
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
        for (int i = tidx; i < nSteps; i += GROUPS * 32) {
            const uint4 dA = __ldg(reinterpret_cast<const uint4 *>(gSteps + i)), dB = __ldg(reinterpret_cast<const uint4 *>(gSteps + i) + 1);
            // {out, CL buffer of the child that is read from memory (child 0 if it is one, else child 1), flags, tip rows}
            sSteps[i] = make_uint4(dA.x, ((dA.y >> 4) & 3u) == 0u ? dA.z : dA.w, dA.y, (dB.x & 0xffffu) | (dB.y << 16));
            sNodes[i] = make_uint2(__ldg(&gSteps[i].n0), __ldg(&gSteps[i].n1));
        }
        if (tidx == 0) {
            for (int i = 0; i < RING; i++) { mbar_init(full + i, 1); mbar_init(empty + i, NW * 32); }
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
    }
    __syncthreads();
    if (tidx == 0)
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
            const double2 v = ld2(row);
            sib[0][r >> 1][r & 1] = v.x;
            sib[1][r >> 1][r & 1] = v.y;
        }
        {
            const double *row = cl + (size_t)(16 + q) * ps;    // state 16 + q: the fifth k-step's operand, straight from its row
            const double2 v = ld2(row);
            a4[0] = v.x;
            a4[1] = v.y;
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
                const double *Tj = T + ((code >> (8 * j)) & 0xffu) * kAATblStates;
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
        // A warp-wide global store costs the load/store pipe one pass per 32-byte sector when the lanes of a quad write
        // different rows -- the accumulator layout: quad = one pattern, its lanes = four states -- and one pass per 128
        // bytes when eight consecutive lanes write one line (tools/microbench/stwave.cu: 16 against 4 passes per 512
        // bytes).  At 86 % of that pipe's capacity (ncu), half of it stores, the pipe is this kernel's bound.  So the
        // warp transposes its tile first: lane L takes the values of lane 4 (L & 7) + (L >> 3) -- one fixed permutation
        // for all twelve values, 24 shuffles -- after which lanes 8r' .. 8r'+7 hold one row's 16 patterns.
        const int src = 4 * (lane & 7) + (lane >> 3);
        double2 w[6];
#pragma unroll
        for (int r = 0; r < 6; r++) {
            w[r].x = __shfl_sync(0xffffffffu, out[0][r >> 1][r & 1], src);
            w[r].y = __shfl_sync(0xffffffffu, out[1][r >> 1][r & 1], src);
        }
        double *o = hd.arena + (size_t)slotCode * 32 + ((size_t)cat * DIM + 2 * (lane >> 3)) * ps + pat0 + 2 * (lane & 7);
        double *r1 = o + ps, *r2 = o + 8 * ps, *r3 = o + 9 * ps, *r4 = o + 16 * ps, *r5 = o + 17 * ps;
        asm volatile(
            "{\n.reg .pred pt;\nsetp.eq.u32 pt, %18, 0;\n"
            "st.global.v2.f64 [%0], {%6, %7};\nst.global.v2.f64 [%1], {%8, %9};\n"
            "st.global.v2.f64 [%2], {%10, %11};\nst.global.v2.f64 [%3], {%12, %13};\n"
            "@pt st.global.v2.f64 [%4], {%14, %15};\n@pt st.global.v2.f64 [%5], {%16, %17};\n}\n" ::"l"(o),
            "l"(r1), "l"(r2), "l"(r3), "l"(r4), "l"(r5), "d"(w[0].x), "d"(w[0].y), "d"(w[1].x), "d"(w[1].y), "d"(w[2].x), "d"(w[2].y), "d"(w[3].x),
            "d"(w[3].y), "d"(w[4].x), "d"(w[4].y), "d"(w[5].x), "d"(w[5].y), "r"((unsigned)(lane >> 4))
            : "memory");
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
        const unsigned tipOff = 2 * childB + (unsigned)(warp * 16 + MT * g);
        const unsigned code0 = *reinterpret_cast<const unsigned short *>(sl + tipOff);      // the lane's two tip codes of a leaf child, one per byte
        const unsigned code1 = *reinterpret_cast<const unsigned short *>(sl + tipOff + GROUPS * 16);
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
        // every lane leaves the slot for itself (NW x 32 arrivals per phase): each lane's reads of the slot are ordered before
        // the refill by its OWN arrival -- no reliance on a __syncwarp in between (compute-sanitizer's racecheck does not follow that
        // edge, and an arrival costs nothing measurable: 3.61 ms either way)
        mbar_arrive(empty + slot);
        if (lane == 0) service(si);
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
