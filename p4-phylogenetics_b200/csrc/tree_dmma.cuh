// tree_dmma.cuh -- whole-tree CL recursion for parts with 21..64 states (61-state codon-like data), FP64 tensor cores.
//
// The reference's loop (Pf/p4_node.c:636-857, generic `dim`: for pattern / cat / parent state / child / child state,
// :727-743 the dot product with an internal child, :685-716 the lookups for a leaf child) costs 2*dim^2 flops per
// internal-child edge, pattern and category: at 61 states that is dense matrix work, so it runs on the FP64 tensor
// cores.  tcgen05 has no f64 kind; FP64 tensor work on sm_100a is mma.sync.m8n8k4 (SASS DMMA.8x8x4).
//
// Same scheme as cl_tree_aa_kernel (kernels.cuh), for any padded state count DP (a multiple of 8):
//     out^T = cl_child^T x P^T      A (8 x 4) = child CL, rows = patterns      B (4 x 8) = P^T      C (8 x 8) = parent CL
// with the summation index dealt to the k-steps so that k-step (t, i) covers the child states {8t + 2q + i}: lane
// (g, q) needs as its A element exactly what it holds as C element i of n-tile t of the child.  A node's result is the
// next node's operand with NO data movement; the running CL stays in the accumulator registers from step to step.
// DP/8 n-tiles x DP/4 k-steps = DP^2/32 DMMAs per child and 8 patterns (128 at DP = 64).
//
// What differs from 20 states is the size of the operands: P^T in fragment order is DP^2*8 bytes per (child, category)
// -- 32 KB at DP = 64 -- so a CTA works on ONE rate category (blockIdx.z) and all its warps share that category's
// operands: two children x two buffers x 32 KB of shared memory, filled one step ahead by bulk copies (TMA,
// cp.async.bulk) that complete on an mbarrier, straight from the operand decks the P(t) kernel leaves in fragment
// order (pmatrix_kernel).  A warp owns 8*MT patterns; a B fragment read from shared memory feeds MT DMMAs.
// Padding states (>= dim) have zero columns in P^T and zero entries in the transposed leaf tables: they are computed
// as zeros and never stored.  The root reduction is like_kernel (the categories of a pattern live in different CTAs).
#pragma once
#include "kernels.cuh"

namespace p4b {

// doubles of one (node, part) operand deck for padded state count DP: per category P^T fragments [DP/4][DP/8][32],
// then per category the transposed leaf table [W][DP]
__host__ __device__ inline size_t dmmaFragDoubles(int DP) { return (size_t)(DP / 4) * (DP / 8) * 32; }
__host__ __device__ inline size_t dmmaAuxDoubles(int DP, int nCat, int W) { return (size_t)nCat * (dmmaFragDoubles(DP) + (size_t)W * DP); }
__host__ inline int dmmaPaddedDim(int dim) { return dim <= 32 ? 32 : 64; }

template <int DP, int MT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
cl_tree_dmma_kernel(const __grid_constant__ TreeArgs a, const int dim, const int nCat)
{
    constexpr int NT = DP / 8, KS = DP / 4, THREADS = WARPS * 32;
    constexpr int WP = 8 * MT;                      // patterns per warp
    const TreeHdr &hd = a.hdr[blockIdx.y];
    const int cat = blockIdx.z;
    extern __shared__ double sm[];                  // [2 buffers][kAAKids][slot], then 2 mbarriers
    const int W = a.tblW;
    const size_t frag = dmmaFragDoubles(DP), tbl = (size_t)W * DP;
    const size_t slot = frag > tbl ? frag : tbl;
    const size_t bufSize = kAAKids * slot;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + 2 * bufSize);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    const int pat0 = (blockIdx.x * WARPS + warp) * WP;
    const bool active = pat0 < a.ps;
    const size_t ps = (size_t)a.ps;
    const size_t rowBase = (size_t)cat * dim * ps + pat0 + MT * g;       // + state * ps: this lane's MT patterns of a row
    const size_t auxFragOff = (size_t)cat * frag, auxLeafOff = (size_t)nCat * frag + (size_t)cat * tbl;
    const int nSteps = hd.nSteps, stepBase = hd.stepBase;

    auto stage = [&](int stepIdx, int b) {          // one thread
        const StepC &st = a.steps[stepBase + stepIdx];
        const int nc = st.nChildren;
        unsigned total = 0;
        for (int c = 0; c < nc; c++) total += (unsigned)((((unsigned)st.ch[c].a >> 30) == 2u ? tbl : frag) * 8);
        mbar_expect_tx(bars + b, total);
        for (int c = 0; c < nc; c++) {
            const bool leaf = ((unsigned)st.ch[c].a >> 30) == 2u;
            const double *src = hd.aux + a.auxNodeDoubles * st.ch[c].b + (leaf ? auxLeafOff : auxFragOff);
            bulk_g2s(sm + b * bufSize + c * slot, src, (unsigned)((leaf ? tbl : frag) * 8), bars + b);
        }
    };
    auto tipCode = [&](int stepIdx, int c) -> unsigned {
        if (!active || stepIdx >= nSteps) return 0u;
        const StepC &st = a.steps[stepBase + stepIdx];
        const unsigned av = (unsigned)st.ch[c].a;
        if (c >= st.nChildren || (av >> 30) != 2u) return 0u;
        const uint8_t *tp = a.tips + (size_t)(av & 0x3fffffffu) * ps + pat0 + MT * g;
        return MT == 4 ? *reinterpret_cast<const unsigned *>(tp) : (MT == 2 ? (unsigned)*reinterpret_cast<const unsigned short *>(tp) : (unsigned)*tp);
    };

    // out (=|*=) A x B for one child; A in the C layout of the child; Bc = this lane's column of the fragment block
    auto contract = [&](const double (&A)[MT][NT][2], const double *__restrict__ Bc, double (&out)[MT][NT][2], bool assign) {
        double acc[MT][NT][2];
#pragma unroll
        for (int j = 0; j < MT; j++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++) acc[j][nt][0] = acc[j][nt][1] = 0.0;
#pragma unroll
        for (int kk = 0; kk < KS; kk++) {
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                const double b = Bc[(kk * NT + nt) * 32];
#pragma unroll
                for (int j = 0; j < MT; j++) dmma884(acc[j][nt][0], acc[j][nt][1], A[j][kk >> 1][kk & 1], b);
            }
        }
#pragma unroll
        for (int j = 0; j < MT; j++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                if (assign) { out[j][nt][0] = acc[j][nt][0]; out[j][nt][1] = acc[j][nt][1]; }
                else { out[j][nt][0] *= acc[j][nt][0]; out[j][nt][1] *= acc[j][nt][1]; }
            }
    };

    unsigned next0 = 0u, next1 = 0u;
    auto step = [&](int si, const double (&in)[MT][NT][2], double (&out)[MT][NT][2]) {
        named_barrier(1, THREADS);               // every warp is done with step si-1: its buffer is free
        if (threadIdx.x == 0 && si + 1 < nSteps) stage(si + 1, (si + 1) & 1);
        const unsigned code0 = next0, code1 = next1;
        next0 = tipCode(si + 1, 0);
        next1 = tipCode(si + 1, 1);
        mbar_wait(bars + (si & 1), (unsigned)(si >> 1) & 1u);
        if (!active) return;
        const double *buf = sm + (si & 1) * bufSize;
        const StepC &st = a.steps[stepBase + si];
        const int nc = st.nChildren;
        int regChild = -1;
        for (int c = 0; c < nc; c++)
            if (((unsigned)st.ch[c].a >> 30) == 1u) regChild = c;
        // The two warps that share a scheduler (w and w + 4) take a step's children in opposite orders: while one runs the
        // DMMAs of the child in registers the other is in its table lookups, instead of both queueing for the tensor pipe and
        // then both leaving it idle (the steps are in lock step: one CTA barrier each).  A product of factors is the same
        // number in either order.
        const bool regLast = regChild >= 0 && nc > 1 && st.first && ((warp >> 2) & 1);
        if (regLast) {
#pragma unroll
            for (int j = 0; j < MT; j++)
#pragma unroll
                for (int t = 0; t < NT; t++) out[j][t][0] = out[j][t][1] = 1.0;
        } else if (regChild >= 0) {
            contract(in, buf + regChild * slot + lane, out, true);
        } else if (!st.first) {
#pragma unroll
            for (int j = 0; j < MT; j++)
#pragma unroll
                for (int t = 0; t < NT; t++) { out[j][t][0] = in[j][t][0]; out[j][t][1] = in[j][t][1]; }
        } else {
#pragma unroll
            for (int j = 0; j < MT; j++)
#pragma unroll
                for (int t = 0; t < NT; t++) out[j][t][0] = out[j][t][1] = 1.0;
        }
        for (int c = 0; c < nc; c++) {
            if (c == regChild) continue;
            const unsigned av = (unsigned)st.ch[c].a, kind = av >> 30;
            if (kind == 2u) {
                // leaf: out *= T[state][code]; the table is [code][DP], so the lane's states 8t+2q, 8t+2q+1 are one 16-byte load
                const unsigned cw = c == 0 ? code0 : code1;
                const double *T = buf + c * slot + 2 * q;
#pragma unroll
                for (int j = 0; j < MT; j++) {
                    const double *Tj = T + (size_t)((cw >> (8 * j)) & 0xffu) * DP;
#pragma unroll
                    for (int t = 0; t < NT; t++) {
                        const double2 v = *reinterpret_cast<const double2 *>(Tj + 8 * t);
                        out[j][t][0] *= v.x;
                        out[j][t][1] *= v.y;
                    }
                }
            } else {                             // internal child in memory (written earlier by this same lane)
                const double *cl = hd.arena + (size_t)(av & 0x3fffffffu) * 32 + rowBase;
                double sib[MT][NT][2];
#pragma unroll
                for (int t = 0; t < NT; t++)
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        const int s = 8 * t + 2 * q + i;
                        if (s < dim) {
                            const double *row = cl + (size_t)s * ps;
                            if (MT == 1) sib[0][t][i] = row[0];
                            else {
#pragma unroll
                                for (int h = 0; h < MT / 2; h++) {
                                    const double2 v = ld2(row + 2 * h);
                                    sib[2 * h][t][i] = v.x;
                                    sib[2 * h + (MT > 1 ? 1 : 0)][t][i] = v.y;
                                }
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < MT; j++) sib[j][t][i] = 0.0;
                        }
                    }
                contract(sib, buf + c * slot + lane, out, false);
            }
        }
        if (regLast) contract(in, buf + regChild * slot + lane, out, false);
        if (st.store) {
            double *o = hd.arena + (size_t)(unsigned)st.outSlot * 32 + rowBase;
#pragma unroll
            for (int t = 0; t < NT; t++)
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const int s = 8 * t + 2 * q + i;
                    if (s < dim) {
                        double *orow = o + (size_t)s * ps;
                        if (MT == 1) orow[0] = out[0][t][i];
                        else {
#pragma unroll
                            for (int h = 0; h < MT / 2; h++) st2(orow + 2 * h, make_double2(out[2 * h][t][i], out[2 * h + (MT > 1 ? 1 : 0)][t][i]));
                        }
                    }
                }
        }
    };

    double cA[MT][NT][2], cB[MT][NT][2];
#pragma unroll
    for (int j = 0; j < MT; j++)
#pragma unroll
        for (int t = 0; t < NT; t++) cA[j][t][0] = cA[j][t][1] = cB[j][t][0] = cB[j][t][1] = 0.0;

    if (threadIdx.x == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    named_barrier(1, THREADS);
    if (threadIdx.x == 0 && nSteps > 0) stage(0, 0);
    next0 = tipCode(0, 0);
    next1 = tipCode(0, 1);
    for (int si = 0; si < nSteps; si += 2) {
        step(si, cA, cB);
        if (si + 1 < nSteps) step(si + 1, cB, cA);
    }
}

}  // namespace p4b
