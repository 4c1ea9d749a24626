// newt.cuh -- sm_100a kernels of the Newton-Raphson branch-length step and of the
// simulation down the tree (8f rank 4).
//
//   newt_deck_kernel   P(t), dP/dv, d2P/dv2 of ONE branch at a trial length, plus the three leaf lookup
//                      tables when the node is a leaf; one CTA per rate category
//                                                       (Pf/eig.c:163-191, 294-318, 346-371;
//                                                        Pf/p4_node.c:296-346, 442-540)
//   transpose_deck_kernel   P^T of a node's deck: the "up" term of a child's cl2 is a CL-kernel child
//                      whose matrix is the parent's P transposed (Pf/p4_node.c:883-928 p4_setCL2Up)
//   newt_dna_kernel    4 states, ONE launch per derivative evaluation: decks in the prologue, two patterns
//                      per thread with the next category's rows in flight, the last CTA folds
//   newt_aa_kernel     20 states: two patterns per thread, 16-byte shared-memory operand reads, last CTA folds
//   newt_kernel<DIM>   any dim (decks in shared memory when they fit, else read through L1) + newt_final_kernel
//                      per pattern: like, d like, d2 like through the branch from cl2 (everything on the
//                      far side of the branch) and the node's own CL or tip; folds
//                      sum count*log(like), sum count*(f/l), sum count*((s*l - f*f)/(l*l))
//                                                       (Pf/p4_treeNewt.c:210-533 p4_newtNode)
//   picker_kernel, simulate_kernel   p4_simulate: running row sums of the P decks, one thread per site walking
//                      the nodes in preOrder on the caller's stream of uniforms (Pf/p4_treeSim.c:315-360)
//
// cl2 arrays have the layout of CL arrays ([cat*dim + state][pattern], row stride ps) and are computed by
// the per-node CL kernels of kernels.cuh: cl2(n) = up(parent) * prod over siblings (P_s x cl_s), where
// up(parent) is pi_root when the parent is the root (a "leaf" child whose table holds pi in every column,
// Pf/p4_node.c:860-881) and P_parent^T x cl2(parent) otherwise.
#pragma once
#include <cstdint>

namespace p4b {

struct NewtDeckJob {
    const double *eig;     // V | Vinv | lambda
    const uint64_t *eq;    // masks of the part's non-N-like equates
    double *decks;         // out: [3][nCat][dim][dim], then (tblW > 0) [3][nCat][dim][tblW]
    int dim, nCat, tblW;
    double t0[16];         // effective branch length per category as P(t) forms it (Pf/p4_node.c:321-345)
    double t1[16];         // ... as the derivative decks form it (:456-485): the same number up to association
    double r1[16];         // factor of the first derivative per category (Pf/p4_node.c:456-485)
    double r2[16];         // factor of the second derivative per category, squared by the kernel (:505-534)
};

// grid.x = rate category: the CTA of category c forms that category's slices of the three decks and tables.  The
// eigensystem is staged in shared memory first (every entry of a deck reads a row of V and a column of V^-1), and the
// category's deck slices stay there for the table pass: nothing on the 20-term sums' path comes from global memory.
// Shared memory: V | V^-1 | lambda | exp(lambda t0) | exp(lambda t1) | the three deck slices = (5 dim^2 + 3 dim) doubles.
__global__ void __launch_bounds__(256)
newt_deck_kernel(const NewtDeckJob job)
{
    extern __shared__ double sDk[];
    const int dim = job.dim, nCat = job.nCat, c = blockIdx.x;
    const int nc = dim * dim;
    double *V = sDk, *Vi = V + nc, *lam = Vi + nc, *sExp = lam + dim, *sExpD = sExp + dim, *sDeck = sExpD + dim;   // sDeck: [3][dim*dim]
    for (int i = threadIdx.x; i < 2 * nc + dim; i += blockDim.x) sDk[i] = job.eig[i];
    __syncthreads();
    for (int i = threadIdx.x; i < dim; i += blockDim.x) {
        sExp[i] = exp(lam[i] * job.t0[c]);
        sExpD[i] = exp(lam[i] * job.t1[c]);
    }
    __syncthreads();
    const int n = nCat * nc;
    double *D0 = job.decks, *D1 = D0 + n, *D2 = D1 + n;
    const double r1 = job.r1[c], r2 = job.r2[c];
    for (int ij = threadIdx.x; ij < nc; ij += blockDim.x) {
        const int i = ij / dim, j = ij - i * dim;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int k = 0; k < dim; k++) {
            // the reference's association: V[i][k] * Vinv[k][j] * lambda * rate * exp(lambda t)
            const double vv = __dmul_rn(V[i * dim + k], Vi[k * dim + j]);
            const double e = sExpD[k], l = lam[k];
            s0 = fma(vv, sExp[k], s0);                             // identical to pmatrix_kernel
            s1 += __dmul_rn(__dmul_rn(__dmul_rn(vv, l), r1), e);
            s2 += __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(vv, l), l), r2), r2), e);
        }
        D0[c * nc + ij] = sDeck[ij] = s0;
        D1[c * nc + ij] = sDeck[nc + ij] = s1;
        D2[c * nc + ij] = sDeck[2 * nc + ij] = s2;
    }
    if (job.tblW > 0) {
        __syncthreads();
        const int W = job.tblW;
        const int nT = nCat * dim * W, ncT = dim * W;
        double *T = D2 + n;
        for (int idx = threadIdx.x; idx < 3 * ncT; idx += blockDim.x) {
            const int d = idx / ncT, r = idx - d * ncT, f = r / W, w = r - f * W;
            const int k = c * dim + f;                             // row of the table: cat*dim + from
            const double *D = sDeck + (size_t)d * nc + (size_t)f * dim;
            double v = 0.0;
            if (w < dim) v = D[w];
            else if (w == dim) {
                // gap, '?', N-like equates: the likelihood term is 1 (Pf/p4_treeNewt.c:292-306), the
                // derivative terms are the row sums of the derivative decks
                if (d == 0) v = 1.0;
                else
                    for (int x = 0; x < dim; x++) v += D[x];
            } else {
                const uint64_t m = job.eq[w - dim - 1];
                for (int x = 0; x < dim; x++)
                    if ((m >> x) & 1ull) v += D[x];
            }
            T[(size_t)d * nT + (size_t)k * W + w] = v;
        }
    }
}

// PT[cat][to][from] = P[cat][from][to]
__global__ void __launch_bounds__(256)
transpose_deck_kernel(const double *__restrict__ P, double *__restrict__ PT, int dim, int nCat)
{
    const int n = nCat * dim * dim;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int c = idx / (dim * dim), ij = idx - c * dim * dim, i = ij / dim, j = ij - i * dim;
        PT[((size_t)c * dim + j) * dim + i] = P[idx];
    }
}

struct NewtArgs {
    const double *cl2;         // [nCat*dim][ps] of the node
    const double *cl;          // internal node: its CL; NULL for a leaf
    const uint8_t *tips;       // leaf: its tip code indices
    const double *decks;       // newt_deck_kernel's output
    const int *counts;
    const uint64_t *invarMask;
    double *partials;          // [3*gridDim.x]
    int ps, nPat, dim, nCat, tblW, useSmem;
    double pInvar;
    double pi[64];             // the root's composition (constant-site term, Pf/p4_treeNewt.c:385-391)
};

// DIM > 0: the node's CL of one category is held in registers; DIM == 0: any dim, re-read through L1.
template <int DIM>
__global__ void __launch_bounds__(128)
newt_kernel(const NewtArgs a)
{
    extern __shared__ double sD[];
    __shared__ double sRed[3][4];
    const int dim = DIM ? DIM : a.dim, nCat = a.nCat, W = a.tblW;
    const bool leaf = a.cl == nullptr;
    // a leaf needs the three tables only, an internal node the three decks only
    const int per = leaf ? nCat * dim * W : nCat * dim * dim;
    const double *src = leaf ? a.decks + (size_t)3 * nCat * dim * dim : a.decks;
    const double *D = src;
    if (a.useSmem) {
        for (int i = threadIdx.x; i < 3 * per; i += blockDim.x) sD[i] = src[i];
        __syncthreads();
        D = sD;
    }
    double tL = 0.0, tF = 0.0, tS = 0.0;
    for (int pat = blockIdx.x * blockDim.x + threadIdx.x; pat < a.nPat; pat += gridDim.x * blockDim.x) {
        const size_t ps = (size_t)a.ps;
        double likeS = 0.0, firstS = 0.0, secondS = 0.0;
        const int code = leaf ? a.tips[pat] : 0;
        for (int c = 0; c < nCat; c++) {
            double like = 0.0, first = 0.0, second = 0.0;
            const double *z = a.cl2 + (size_t)c * dim * ps + pat;
            if (leaf) {
                const double *T0 = D + (size_t)c * dim * W + code, *T1 = T0 + per, *T2 = T1 + per;
                for (int f = 0; f < dim; f++) {
                    const double zz = z[f * ps];
                    like = fma(zz, T0[f * W], like);
                    first = fma(zz, T1[f * W], first);
                    second = fma(zz, T2[f * W], second);
                }
            } else {
                const double *x = a.cl + (size_t)c * dim * ps + pat;
                const double *D0 = D + (size_t)c * dim * dim, *D1 = D0 + per, *D2 = D1 + per;
                if (DIM) {
                    double xr[DIM ? DIM : 1];
#pragma unroll
                    for (int t = 0; t < DIM; t++) xr[t] = x[t * ps];
                    for (int f = 0; f < DIM; f++) {
                        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
                        for (int t = 0; t < DIM; t++) {
                            a0 = fma(D0[f * DIM + t], xr[t], a0);
                            a1 = fma(D1[f * DIM + t], xr[t], a1);
                            a2 = fma(D2[f * DIM + t], xr[t], a2);
                        }
                        const double zz = z[f * ps];
                        like = fma(zz, a0, like);
                        first = fma(zz, a1, first);
                        second = fma(zz, a2, second);
                    }
                } else {
                    for (int f = 0; f < dim; f++) {
                        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
                        for (int t = 0; t < dim; t++) {
                            const double xt = x[t * ps];
                            a0 = fma(D0[f * dim + t], xt, a0);
                            a1 = fma(D1[f * dim + t], xt, a1);
                            a2 = fma(D2[f * dim + t], xt, a2);
                        }
                        const double zz = z[f * ps];
                        like = fma(zz, a0, like);
                        first = fma(zz, a1, first);
                        second = fma(zz, a2, second);
                    }
                }
            }
            likeS += like;
            firstS += first;
            secondS += second;
        }
        if (a.pInvar != 0.0) {                      // Pf/p4_treeNewt.c:380-391
            const double f = (1.0 - a.pInvar) / (double)nCat;
            likeS *= f;
            firstS *= f;
            secondS *= f;
            const uint64_t im = a.invarMask ? a.invarMask[pat] : 0ull;
            if (im)
                for (int s = 0; s < dim; s++)
                    if ((im >> s) & 1ull) likeS += a.pi[s] * a.pInvar;
        } else if (nCat > 1) {                       // :501-505
            likeS /= (double)nCat;
            firstS /= (double)nCat;
            secondS /= (double)nCat;
        }
        const double cnt = (double)a.counts[pat];
        if (likeS < 1.0e-300) {                      // :508-512
            tL += cnt * -100000.0;
            tF += cnt * 1000000.0;
            tS += cnt * 10000000.0;
        } else {                                     // :513-517
            tL += cnt * log(likeS);
            tF += cnt * (firstS / likeS);
            tS += cnt * ((secondS * likeS - firstS * firstS) / (likeS * likeS));
        }
    }
    tL = warpSum(tL);
    tF = warpSum(tF);
    tS = warpSum(tS);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sRed[0][w] = tL; sRed[1][w] = tF; sRed[2][w] = tS; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) v += sRed[threadIdx.x][i];
        a.partials[3 * blockIdx.x + threadIdx.x] = v;
    }
}

// ---------------------------------------------------------------------------
// 4 states: ONE launch per derivative evaluation.
//   prologue  every CTA forms the three decks (and, for a leaf, the three lookup tables) of the branch in
//             shared memory itself -- 3 x NCAT x 16 entries, the expressions of newt_deck_kernel;
//   body      a thread owns two adjacent patterns (16-byte loads, all 8 rows of a category in flight),
//             persistent CTAs stride over the pattern pairs;
//   epilogue  per-CTA partials, then the LAST CTA to finish (atomic ticket) folds them in a fixed order,
//             so the result is deterministic and no second launch is needed.
// ---------------------------------------------------------------------------
template <int NCAT>
__global__ void __launch_bounds__(256)
newt_dna_kernel(const NewtArgs a, const NewtDeckJob job, unsigned *__restrict__ ticket, double *__restrict__ result)
{
    extern __shared__ double sD[];          // decks [3][NCAT][4][4], then (leaf) tables [3][NCAT*4][W]
    __shared__ double sExp[2][NCAT * 4];
    __shared__ double sRed[3][8];
    __shared__ bool sLast;
    constexpr int ND = NCAT * 16;
    const bool leaf = a.cl == nullptr;
    const int W = a.tblW;
    {
        const double *V = job.eig, *Vi = V + 16, *lam = Vi + 16;
        if (threadIdx.x < NCAT * 4) {
            sExp[0][threadIdx.x] = exp(lam[threadIdx.x & 3] * job.t0[threadIdx.x >> 2]);
            sExp[1][threadIdx.x] = exp(lam[threadIdx.x & 3] * job.t1[threadIdx.x >> 2]);
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < ND; idx += blockDim.x) {
            const int c = idx >> 4, i = (idx >> 2) & 3, j = idx & 3;
            const double r1 = job.r1[c], r2 = job.r2[c];
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const double vv = __dmul_rn(V[i * 4 + k], Vi[k * 4 + j]);
                const double e = sExp[1][c * 4 + k], l = lam[k];
                s0 = fma(vv, sExp[0][c * 4 + k], s0);
                s1 += __dmul_rn(__dmul_rn(__dmul_rn(vv, l), r1), e);
                s2 += __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(vv, l), l), r2), r2), e);
            }
            sD[idx] = s0;
            sD[ND + idx] = s1;
            sD[2 * ND + idx] = s2;
        }
        __syncthreads();
        if (leaf) {
            const int nT = NCAT * 4 * W;
            double *T = sD + 3 * ND;
            for (int idx = threadIdx.x; idx < 3 * nT; idx += blockDim.x) {
                const int d = idx / nT, r = idx - d * nT, k = r / W, w = r - k * W;
                const double *D = sD + d * ND + k * 4;
                double v = 0.0;
                if (w < 4) v = D[w];
                else if (w == 4) {
                    if (d == 0) v = 1.0;
                    else
                        for (int x = 0; x < 4; x++) v += D[x];
                } else {
                    const uint64_t m = job.eq[w - 5];
                    for (int x = 0; x < 4; x++)
                        if ((m >> x) & 1ull) v += D[x];
                }
                T[idx] = v;
            }
            __syncthreads();
        }
    }
    const size_t ps = (size_t)a.ps;
    const int nPairs = a.ps >> 1;
    const int stride = gridDim.x * blockDim.x;
    double tL = 0.0, tF = 0.0, tS = 0.0;
    // The rows of the NEXT rate category (or of the thread's next pattern pair) are requested before the
    // current category is worked on: 8 (leaf: 4) 16-byte loads are always in flight behind the arithmetic.
    int pair = blockIdx.x * blockDim.x + threadIdx.x;
    double2 zc[4], xc[4];
    uchar2 codeC = make_uchar2(0, 0);
#pragma unroll
    for (int f = 0; f < 4; f++) zc[f] = xc[f] = make_double2(0.0, 0.0);
    if (pair < nPairs) {
#pragma unroll
        for (int f = 0; f < 4; f++) zc[f] = ld2(a.cl2 + (size_t)f * ps + 2 * pair);
        if (leaf) codeC = *reinterpret_cast<const uchar2 *>(a.tips + 2 * pair);
        else {
#pragma unroll
            for (int f = 0; f < 4; f++) xc[f] = ld2(a.cl + (size_t)f * ps + 2 * pair);
        }
    }
    for (; pair < nPairs; pair += stride) {
        const int pat = pair * 2;
        const int nextPat = (pair + stride) * 2;
        const bool more = pair + stride < nPairs;
        double2 likeS = make_double2(0.0, 0.0), firstS = likeS, secondS = likeS;
        uchar2 codeN = codeC;
#pragma unroll
        for (int c = 0; c < NCAT; c++) {
            double2 zn[4], xn[4];
#pragma unroll
            for (int f = 0; f < 4; f++) { zn[f] = zc[f]; xn[f] = xc[f]; }
            if (c + 1 < NCAT) {
#pragma unroll
                for (int f = 0; f < 4; f++) zn[f] = ld2(a.cl2 + (size_t)((c + 1) * 4 + f) * ps + pat);
                if (!leaf) {
#pragma unroll
                    for (int f = 0; f < 4; f++) xn[f] = ld2(a.cl + (size_t)((c + 1) * 4 + f) * ps + pat);
                }
            } else if (more) {
#pragma unroll
                for (int f = 0; f < 4; f++) zn[f] = ld2(a.cl2 + (size_t)f * ps + nextPat);
                if (leaf) codeN = *reinterpret_cast<const uchar2 *>(a.tips + nextPat);
                else {
#pragma unroll
                    for (int f = 0; f < 4; f++) xn[f] = ld2(a.cl + (size_t)f * ps + nextPat);
                }
            }
            double2 like = make_double2(0.0, 0.0), first = like, second = like;
            if (leaf) {
                const int nT = NCAT * 4 * W;
                const double *T0 = sD + 3 * ND + c * 4 * W, *T1 = T0 + nT, *T2 = T1 + nT;
#pragma unroll
                for (int f = 0; f < 4; f++) {
                    like.x = fma(zc[f].x, T0[f * W + codeC.x], like.x);
                    like.y = fma(zc[f].y, T0[f * W + codeC.y], like.y);
                    first.x = fma(zc[f].x, T1[f * W + codeC.x], first.x);
                    first.y = fma(zc[f].y, T1[f * W + codeC.y], first.y);
                    second.x = fma(zc[f].x, T2[f * W + codeC.x], second.x);
                    second.y = fma(zc[f].y, T2[f * W + codeC.y], second.y);
                }
            } else {
                const double *D0 = sD + c * 16, *D1 = D0 + ND, *D2 = D1 + ND;
#pragma unroll
                for (int f = 0; f < 4; f++) {
                    double2 a0 = make_double2(0.0, 0.0), a1 = a0, a2 = a0;
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const double d0 = D0[f * 4 + t], d1 = D1[f * 4 + t], d2 = D2[f * 4 + t];
                        a0.x = fma(d0, xc[t].x, a0.x);
                        a0.y = fma(d0, xc[t].y, a0.y);
                        a1.x = fma(d1, xc[t].x, a1.x);
                        a1.y = fma(d1, xc[t].y, a1.y);
                        a2.x = fma(d2, xc[t].x, a2.x);
                        a2.y = fma(d2, xc[t].y, a2.y);
                    }
                    like.x = fma(zc[f].x, a0.x, like.x);
                    like.y = fma(zc[f].y, a0.y, like.y);
                    first.x = fma(zc[f].x, a1.x, first.x);
                    first.y = fma(zc[f].y, a1.y, first.y);
                    second.x = fma(zc[f].x, a2.x, second.x);
                    second.y = fma(zc[f].y, a2.y, second.y);
                }
            }
            likeS.x += like.x;
            likeS.y += like.y;
            firstS.x += first.x;
            firstS.y += first.y;
            secondS.x += second.x;
            secondS.y += second.y;
#pragma unroll
            for (int f = 0; f < 4; f++) { zc[f] = zn[f]; xc[f] = xn[f]; }
        }
        codeC = codeN;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int pp = pat + h;
            if (pp >= a.nPat) continue;
            double l = h ? likeS.y : likeS.x, f1 = h ? firstS.y : firstS.x, f2 = h ? secondS.y : secondS.x;
            if (a.pInvar != 0.0) {                      // Pf/p4_treeNewt.c:380-391
                const double f = (1.0 - a.pInvar) / (double)NCAT;
                l *= f;
                f1 *= f;
                f2 *= f;
                const uint64_t im = a.invarMask ? a.invarMask[pp] : 0ull;
                if (im)
                    for (int s = 0; s < 4; s++)
                        if ((im >> s) & 1ull) l += a.pi[s] * a.pInvar;
            } else if (NCAT > 1) {                       // :501-505
                l /= (double)NCAT;
                f1 /= (double)NCAT;
                f2 /= (double)NCAT;
            }
            const double cnt = (double)a.counts[pp];
            if (l < 1.0e-300) {                          // :508-512
                tL += cnt * -100000.0;
                tF += cnt * 1000000.0;
                tS += cnt * 10000000.0;
            } else {                                     // :513-517
                tL += cnt * log(l);
                tF += cnt * (f1 / l);
                tS += cnt * ((f2 * l - f1 * f1) / (l * l));
            }
        }
    }
    tL = warpSum(tL);
    tF = warpSum(tF);
    tS = warpSum(tS);
    const int w = threadIdx.x >> 5, ln = threadIdx.x & 31;
    if (ln == 0) { sRed[0][w] = tL; sRed[1][w] = tF; sRed[2][w] = tS; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double v[3] = {0.0, 0.0, 0.0};
        for (int i = 0; i < 8; i++) { v[0] += sRed[0][i]; v[1] += sRed[1][i]; v[2] += sRed[2][i]; }
        a.partials[3 * blockIdx.x] = v[0];
        a.partials[3 * blockIdx.x + 1] = v[1];
        a.partials[3 * blockIdx.x + 2] = v[2];
        __threadfence();
        const unsigned t = atomicInc(ticket, gridDim.x - 1);   // wraps to 0 with the last CTA: ready for the next launch
        sLast = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (sLast) {
        __threadfence();
        double v[3] = {0.0, 0.0, 0.0};
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
            v[0] += __ldcg(a.partials + 3 * i);
            v[1] += __ldcg(a.partials + 3 * i + 1);
            v[2] += __ldcg(a.partials + 3 * i + 2);
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            v[k] = warpSum(v[k]);
            if (ln == 0) sRed[k][w] = v[k];
        }
        __syncthreads();
        if (threadIdx.x < 3) {
            double s2 = 0.0;
            for (int i = 0; i < 8; i++) s2 += sRed[threadIdx.x][i];
            result[threadIdx.x] = s2;
        }
    }
}

// Mixing over categories, constant-site term and the three per-pattern terms (Pf/p4_treeNewt.c:380-391, 501-517).
__device__ __forceinline__ void newt_finish(const NewtArgs &a, int pp, int dim, int nCat, double l, double f1, double f2,
                                            double &tL, double &tF, double &tS)
{
    if (a.pInvar != 0.0) {
        const double f = (1.0 - a.pInvar) / (double)nCat;
        l *= f;
        f1 *= f;
        f2 *= f;
        const uint64_t im = a.invarMask ? a.invarMask[pp] : 0ull;
        if (im)
            for (int s = 0; s < dim; s++)
                if ((im >> s) & 1ull) l += a.pi[s] * a.pInvar;
    } else if (nCat > 1) {
        l /= (double)nCat;
        f1 /= (double)nCat;
        f2 /= (double)nCat;
    }
    const double cnt = (double)a.counts[pp];
    if (l < 1.0e-300) {
        tL += cnt * -100000.0;
        tF += cnt * 1000000.0;
        tS += cnt * 10000000.0;
    } else {
        tL += cnt * log(l);
        tF += cnt * (f1 / l);
        tS += cnt * ((f2 * l - f1 * f1) / (l * l));
    }
}

// ---------------------------------------------------------------------------
// 20 states.  1260 FMAs per pattern and category against 320 bytes: the arithmetic and the shared-memory
// operand reads bound this kernel, not HBM.  A thread owns TWO adjacent patterns, so every deck entry read
// from shared memory (16-byte reads, two entries each) feeds two FMAs per deck; the node's CL of the current
// category sits in registers.  Persistent CTAs; the last CTA to finish folds the partials (fixed order).
// The decks come from newt_deck_kernel (one small launch before this one).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
newt_aa_kernel(const NewtArgs a, unsigned *__restrict__ ticket, double *__restrict__ result)
{
    constexpr int DIM = 20;
    extern __shared__ double sD[];
    __shared__ double sRed[3][4];
    __shared__ bool sLast;
    const int nCat = a.nCat, W = a.tblW;
    const bool leaf = a.cl == nullptr;
    const int per = leaf ? nCat * DIM * W : nCat * DIM * DIM;
    {
        const double *src = leaf ? a.decks + (size_t)3 * nCat * DIM * DIM : a.decks;
        for (int i = threadIdx.x; i < 3 * per; i += blockDim.x) sD[i] = src[i];
        __syncthreads();
    }
    const size_t ps = (size_t)a.ps;
    const int nPairs = a.ps >> 1;
    double tL = 0.0, tF = 0.0, tS = 0.0;
    for (int pair = blockIdx.x * blockDim.x + threadIdx.x; pair < nPairs; pair += gridDim.x * blockDim.x) {
        const int pat = pair * 2;
        double2 likeS = make_double2(0.0, 0.0), firstS = likeS, secondS = likeS;
        uchar2 code = make_uchar2(0, 0);
        if (leaf) code = *reinterpret_cast<const uchar2 *>(a.tips + pat);
        for (int c = 0; c < nCat; c++) {
            const double *z = a.cl2 + (size_t)c * DIM * ps + pat;
            double2 like = make_double2(0.0, 0.0), first = like, second = like;
            if (leaf) {
                const double *T0 = sD + (size_t)c * DIM * W, *T1 = T0 + per, *T2 = T1 + per;
                double2 zr[DIM];          // all twenty rows requested before the first is used: the loads in flight are what
#pragma unroll                            // bounds a leaf evaluation (128 MB read, 120 FMAs per pattern pair and category)
                for (int f = 0; f < DIM; f++) zr[f] = ld2(z + (size_t)f * ps);
#pragma unroll
                for (int f = 0; f < DIM; f++) {
                    const double2 zz = zr[f];
                    like.x = fma(zz.x, T0[f * W + code.x], like.x);
                    like.y = fma(zz.y, T0[f * W + code.y], like.y);
                    first.x = fma(zz.x, T1[f * W + code.x], first.x);
                    first.y = fma(zz.y, T1[f * W + code.y], first.y);
                    second.x = fma(zz.x, T2[f * W + code.x], second.x);
                    second.y = fma(zz.y, T2[f * W + code.y], second.y);
                }
            } else {
                const double *x = a.cl + (size_t)c * DIM * ps + pat;
                double2 xr[DIM];
#pragma unroll
                for (int t = 0; t < DIM; t++) xr[t] = ld2(x + (size_t)t * ps);
                const double *D0 = sD + (size_t)c * DIM * DIM, *D1 = D0 + per, *D2 = D1 + per;
#pragma unroll 2
                for (int f = 0; f < DIM; f++) {
                    const double2 zz = ld2(z + (size_t)f * ps);
                    double2 a0 = make_double2(0.0, 0.0), a1 = a0, a2 = a0;
#pragma unroll
                    for (int t = 0; t < DIM; t += 2) {
                        const double2 d0 = *reinterpret_cast<const double2 *>(D0 + f * DIM + t);
                        const double2 d1 = *reinterpret_cast<const double2 *>(D1 + f * DIM + t);
                        const double2 d2 = *reinterpret_cast<const double2 *>(D2 + f * DIM + t);
                        a0.x = fma(d0.x, xr[t].x, a0.x);
                        a0.y = fma(d0.x, xr[t].y, a0.y);
                        a1.x = fma(d1.x, xr[t].x, a1.x);
                        a1.y = fma(d1.x, xr[t].y, a1.y);
                        a2.x = fma(d2.x, xr[t].x, a2.x);
                        a2.y = fma(d2.x, xr[t].y, a2.y);
                        a0.x = fma(d0.y, xr[t + 1].x, a0.x);
                        a0.y = fma(d0.y, xr[t + 1].y, a0.y);
                        a1.x = fma(d1.y, xr[t + 1].x, a1.x);
                        a1.y = fma(d1.y, xr[t + 1].y, a1.y);
                        a2.x = fma(d2.y, xr[t + 1].x, a2.x);
                        a2.y = fma(d2.y, xr[t + 1].y, a2.y);
                    }
                    like.x = fma(zz.x, a0.x, like.x);
                    like.y = fma(zz.y, a0.y, like.y);
                    first.x = fma(zz.x, a1.x, first.x);
                    first.y = fma(zz.y, a1.y, first.y);
                    second.x = fma(zz.x, a2.x, second.x);
                    second.y = fma(zz.y, a2.y, second.y);
                }
            }
            likeS.x += like.x;
            likeS.y += like.y;
            firstS.x += first.x;
            firstS.y += first.y;
            secondS.x += second.x;
            secondS.y += second.y;
        }
        if (pat < a.nPat) newt_finish(a, pat, DIM, nCat, likeS.x, firstS.x, secondS.x, tL, tF, tS);
        if (pat + 1 < a.nPat) newt_finish(a, pat + 1, DIM, nCat, likeS.y, firstS.y, secondS.y, tL, tF, tS);
    }
    tL = warpSum(tL);
    tF = warpSum(tF);
    tS = warpSum(tS);
    const int w = threadIdx.x >> 5, ln = threadIdx.x & 31;
    if (ln == 0) { sRed[0][w] = tL; sRed[1][w] = tF; sRed[2][w] = tS; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double v[3] = {0.0, 0.0, 0.0};
        for (int i = 0; i < 4; i++) { v[0] += sRed[0][i]; v[1] += sRed[1][i]; v[2] += sRed[2][i]; }
        a.partials[3 * blockIdx.x] = v[0];
        a.partials[3 * blockIdx.x + 1] = v[1];
        a.partials[3 * blockIdx.x + 2] = v[2];
        __threadfence();
        const unsigned t = atomicInc(ticket, gridDim.x - 1);
        sLast = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (sLast) {
        __threadfence();
        double v[3] = {0.0, 0.0, 0.0};
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
            v[0] += __ldcg(a.partials + 3 * i);
            v[1] += __ldcg(a.partials + 3 * i + 1);
            v[2] += __ldcg(a.partials + 3 * i + 2);
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            v[k] = warpSum(v[k]);
            if (ln == 0) sRed[k][w] = v[k];
        }
        __syncthreads();
        if (threadIdx.x < 3) {
            double s2 = 0.0;
            for (int i = 0; i < 4; i++) s2 += sRed[threadIdx.x][i];
            result[threadIdx.x] = s2;
        }
    }
}

// ---------------------------------------------------------------------------
// 20 states, internal node, on the FP64 tensor cores.  The three sums of a pattern and category are
//     like / first / second = SUM_f z[f] * (D_d x)[f],   D_0 = P(v), D_1 = dP/dv, D_2 = d2P/dv2 (20 x 20 each)
// (Pf/p4_treeNewt.c:238-520): three matrix products with the node's CL -- the contraction of the whole-tree kernel
// (tree_aa.cuh) with D_d in the place of P: Y^T = x^T D_d^T on mma.sync.m8n8k4, patterns as rows, the summation index
// dealt to the k-steps so that k-step (t, i) covers the states {8t + 2q + i} and a fifth one the states 16 + q; 15 DMMAs
// per deck and 8 patterns, 45 for the three decks, where the FMA kernel above spends 1260 FMAs and 15 shared-memory reads
// per pattern.  A lane then holds Y[pattern g][f = 8nt + 2q + i]; it weights them with z[f] (read in the same layout),
// adds up over its f's and over the categories, and the quad (q = 0..3) folds its four partial sums: the order of the
// additions differs from the reference's f-then-category order by rounding only.
// A warp owns 16 patterns (two m-tiles; a lane's two patterns are one 16-byte access) and takes tiles round-robin
// (persistent CTAs); the CTA's prologue lays the three decks out in shared memory in fragment order
// [deck][cat][k-step][n-tile][lane] (zero rows for the padding states 20..23).  The last CTA folds the partials.
// ---------------------------------------------------------------------------
template <int WARPS, int MINB, int NCAT>
__global__ void __launch_bounds__(WARPS * 32, MINB)
newt_aa_dmma_kernel(const NewtArgs a, unsigned *__restrict__ ticket, double *__restrict__ result)
{
    constexpr int DIM = 20, FRAG = 15 * 32, MT = 2;
    extern __shared__ double sF[];            // [3][nCat][FRAG]
    __shared__ double sRed[3][WARPS];
    __shared__ bool sLast;
    const int nCat = NCAT ? NCAT : a.nCat;
    {
        const int per = nCat * DIM * DIM, total = 3 * nCat * FRAG;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int r = i % FRAG, ct = (i / FRAG) % nCat, d = i / (FRAG * nCat);
            const int l = r & 31, nt = (r >> 5) % 3, kk = r / 96;
            const int f = 8 * nt + (l >> 2), x = kk < 4 ? 8 * (kk >> 1) + 2 * (l & 3) + (kk & 1) : 16 + (l & 3);
            sF[i] = f < DIM ? a.decks[(size_t)d * per + (size_t)ct * DIM * DIM + f * DIM + x] : 0.0;
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    const size_t ps = (size_t)a.ps;
    const int nTiles = a.ps >> 4, tileStride = gridDim.x * WARPS;
    double tL = 0.0, tF = 0.0, tS = 0.0;

    struct Ops { double A[5][MT], Z[3][2][MT]; };
    // the lane's operands of one (tile, category): x in the A layout (states 8t + 2q + i and 16 + q), z in the C layout
    auto load = [&](int tile, int c, Ops &o) {
        const size_t base = (size_t)c * DIM * ps + (size_t)tile * 16 + MT * g;
        const double *x = a.cl + base, *z = a.cl2 + base;
#pragma unroll
        for (int r = 0; r < 4; r++) {            // states 8t + 2q + i, t = r >> 1, i = r & 1
            const size_t row = (size_t)(8 * (r >> 1) + 2 * q + (r & 1)) * ps;
            const double2 xv = ld2(x + row), zv = ld2(z + row);
            o.A[r][0] = xv.x;
            o.A[r][1] = xv.y;
            o.Z[r >> 1][r & 1][0] = zv.x;
            o.Z[r >> 1][r & 1][1] = zv.y;
        }
        {
            const double2 xv = ld2(x + (size_t)(16 + q) * ps);     // the fifth k-step's operand: state 16 + q
            o.A[4][0] = xv.x;
            o.A[4][1] = xv.y;
        }
#pragma unroll
        for (int i = 0; i < 2; i++) {            // states 16 + 2q + i exist for q < 2; the padding states weigh nothing
            double2 zv = make_double2(0.0, 0.0);
            if (q < 2) zv = ld2(z + (size_t)(16 + 2 * q + i) * ps);
            o.Z[2][i][0] = zv.x;
            o.Z[2][i][1] = zv.y;
        }
    };
    auto compute = [&](int c, const Ops &o, double (&sum)[3][MT]) {
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double *Bc = sF + ((size_t)d * nCat + c) * FRAG + lane;
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
                    for (int j = 0; j < MT; j++) dmma884(acc[j][nt][0], acc[j][nt][1], o.A[kk][j], b);
                }
            }
#pragma unroll
            for (int j = 0; j < MT; j++)
#pragma unroll
                for (int nt = 0; nt < 3; nt++) {
                    sum[d][j] = fma(o.Z[nt][0][j], acc[j][nt][0], sum[d][j]);
                    sum[d][j] = fma(o.Z[nt][1][j], acc[j][nt][1], sum[d][j]);
                }
        }
    };
    auto finish = [&](int tile, double (&sum)[3][MT]) {
        // the quad's four lanes hold the sums over their own f's: fold them (every lane ends with the total)
#pragma unroll
        for (int d = 0; d < 3; d++)
#pragma unroll
            for (int j = 0; j < MT; j++) {
                double v = sum[d][j];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                sum[d][j] = v;
            }
        // lane q = 0 finishes the quad's first pattern, lane q = 1 the second
        if (q < MT) {
            const int pp = tile * 16 + MT * g + q;
            if (pp < a.nPat) newt_finish(a, pp, DIM, nCat, q ? sum[0][1] : sum[0][0], q ? sum[1][1] : sum[1][0], q ? sum[2][1] : sum[2][0], tL, tF, tS);
        }
    };

    if (NCAT > 0 && (NCAT & 1) == 0) {
        // an even number of categories known at compile time: the operands of the next (tile, category) are requested
        // before the DMMAs of the current one, in two operand sets that swap roles (the unrolled loop indexes them statically)
        Ops o[2];
        int tile = blockIdx.x * WARPS + warp;
        if (tile < nTiles) load(tile, 0, o[0]);
        for (; tile < nTiles; tile += tileStride) {
            double sum[3][MT];
#pragma unroll
            for (int d = 0; d < 3; d++) sum[d][0] = sum[d][1] = 0.0;
#pragma unroll
            for (int c = 0; c < (NCAT ? NCAT : 1); c++) {
                if (c + 1 < NCAT) load(tile, c + 1, o[(c + 1) & 1]);
                else if (tile + tileStride < nTiles) load(tile + tileStride, 0, o[(c + 1) & 1]);
                compute(c, o[c & 1], sum);
            }
            finish(tile, sum);
        }
    } else {
        for (int tile = blockIdx.x * WARPS + warp; tile < nTiles; tile += tileStride) {
            double sum[3][MT];
#pragma unroll
            for (int d = 0; d < 3; d++) sum[d][0] = sum[d][1] = 0.0;
            for (int c = 0; c < nCat; c++) {
                Ops o;
                load(tile, c, o);
                compute(c, o, sum);
            }
            finish(tile, sum);
        }
    }
    tL = warpSum(tL);
    tF = warpSum(tF);
    tS = warpSum(tS);
    if (lane == 0) { sRed[0][warp] = tL; sRed[1][warp] = tF; sRed[2][warp] = tS; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double v[3] = {0.0, 0.0, 0.0};
        for (int i = 0; i < WARPS; i++) { v[0] += sRed[0][i]; v[1] += sRed[1][i]; v[2] += sRed[2][i]; }
        a.partials[3 * blockIdx.x] = v[0];
        a.partials[3 * blockIdx.x + 1] = v[1];
        a.partials[3 * blockIdx.x + 2] = v[2];
        __threadfence();
        const unsigned t = atomicInc(ticket, gridDim.x - 1);   // wraps to 0 with the last CTA: ready for the next launch
        sLast = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (sLast) {
        __threadfence();
        double v[3] = {0.0, 0.0, 0.0};
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
            v[0] += __ldcg(a.partials + 3 * i);
            v[1] += __ldcg(a.partials + 3 * i + 1);
            v[2] += __ldcg(a.partials + 3 * i + 2);
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            v[k] = warpSum(v[k]);
            if (lane == 0) sRed[k][warp] = v[k];
        }
        __syncthreads();
        if (threadIdx.x < 3) {
            double s2 = 0.0;
            for (int i = 0; i < WARPS; i++) s2 += sRed[threadIdx.x][i];
            result[threadIdx.x] = s2;
        }
    }
}

// One block folds the per-block partials in a fixed order (deterministic): result[0..2].
__global__ void __launch_bounds__(256)
newt_final_kernel(const double *__restrict__ partials, int n, double *__restrict__ result)
{
    __shared__ double sRed[3][8];
    double v[3] = {0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        v[0] += partials[3 * i];
        v[1] += partials[3 * i + 1];
        v[2] += partials[3 * i + 2];
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        v[k] = warpSum(v[k]);
        if (l == 0) sRed[k][w] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int i = 0; i < 8; i++) s += sRed[threadIdx.x][i];
        result[threadIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------
// Simulation down the tree (p4_simulate, Pf/p4_treeSim.c:14-420): integer / byte work, one thread per SITE.
//   picker_kernel    picker[cat][from][to] = running sum of the P deck's row, last entry 1.0
//                    (p4_calculatePickerDecks, Pf/p4_node.c)
//   simulate_kernel  a chunk of nodes in preOrder: a site's state at a node is its parent's state at an
//                    invariant site, else the first `to` with u < picker[cat][parentState][to], u being the
//                    site's own uniform of that node in the reference's stream order (:330-360).
// States live as one byte per (node, site); a thread reads back only bytes it wrote itself.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
picker_kernel(const double *__restrict__ Pdeck, double *__restrict__ picker, long long pNodeDoubles, int pOff, int dim, int nCat, int nNodes)
{
    // one thread per (node, cat, from) row
    const int rows = nNodes * nCat * dim;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
        const int node = r / (nCat * dim), k = r - node * nCat * dim;
        const double *src = Pdeck + (size_t)pNodeDoubles * node + pOff + (size_t)k * dim;
        double *dst = picker + ((size_t)node * nCat * dim + k) * dim;
        double run = src[0];
        dst[0] = run;
        for (int l = 1; l < dim - 1; l++) {
            run = run + src[l];
            dst[l] = run;
        }
        dst[dim - 1] = 1.0;
    }
}

constexpr int kSimChunk = 64;
struct SimArgs {
    uint8_t *states;            // [nNodes][nChar]
    const uint8_t *cats;        // [nChar]
    const uint8_t *invar;       // [nChar]
    const int *rank;            // [nChar]: index of the site among the variable sites
    const double *picker;       // [nNodes][nCat][dim][dim]
    const double *U;            // [nChunkNodes][nVar] uniforms of this chunk
    int nChar, nVar, dim, nCat, n;
    int node[kSimChunk], parent[kSimChunk];
};

__global__ void __launch_bounds__(256)
simulate_kernel(const SimArgs a)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.nChar) return;
    const bool inv = a.invar[k] != 0;
    const int cat = a.cats[k], rk = a.rank[k], dim = a.dim;
    for (int j = 0; j < a.n; j++) {
        const int parentState = a.states[(size_t)a.parent[j] * a.nChar + k];
        int st = parentState;
        if (!inv) {
            const double u = a.U[(size_t)j * a.nVar + rk];
            const double *row = a.picker + (((size_t)a.node[j] * a.nCat + cat) * dim + parentState) * dim;
            st = dim - 1;
            for (int l = 0; l < dim; l++)
                if (u < __ldg(row + l)) { st = l; break; }
        }
        a.states[(size_t)a.node[j] * a.nChar + k] = (uint8_t)st;
    }
}

}  // namespace p4b
