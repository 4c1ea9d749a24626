// kernels.cuh -- sm_100a FP64 kernels of the likelihood path.
//
//   pmatrix_kernel   batched P(t) = V exp(Lt) V^-1       (Pf/eig.c:163-191, Pf/p4_node.c:296-346)
//   cl_dna_kernel    CL update, 4 states, pattern pairs  (Pf/p4_node.c:636-857)
//   cl_dim_kernel    CL update, DIM states in registers
//   cl_generic_kernel CL update, any dim <= 64
//   like_kernel / like_final_kernel  per-part lnL        (Pf/p4_tree.c:1029-1197, 1199-1378)
//
// Memory layout (all in HBM, per tree and part):
//   CL of a node      cl[k][pat], k = cat*dim + state, row stride `ps` doubles
//                     (ps = shard pattern count padded to a multiple of 32, so
//                     rows are 256-byte aligned).  Same index order as the
//                     reference's cl[part][cat][state][pattern] (Pf/pftypes.h:133).
//   P deck of a node  P[cat][from][to] row-major, as bigPDecks (Pf/pftypes.h:130).
//   leaf table        T[k][w], w = tip code index: w < dim a state (P[s][w]),
//                     w == dim "matches all" (1.0), w > dim an ambiguity
//                     (sum of P[s][x] over its states, Pf/p4_node.c:705-716).
//   tips              one byte per (taxon, pattern): the tip code index.
#pragma once
#include <cstdint>

namespace p4b {

constexpr int kMaxChildren = 6;    // children folded into one CL launch; more are chained

struct PJob {
    double *P;             // the node's P deck for this part: [cat][from][to]
    double *tbl;           // the node's leaf lookup table; used when tblW > 0
    const double *eig;     // device mirror of the eigensystem: V | Vinv | lambda
    const uint64_t *eq;    // masks of the part's non-N-like equates
    double *aux;           // 20-state parts served by the whole-tree kernel: P^T in fragment order + transposed leaf table
    int dim, nCat, tblW, nRealEq;
    int auxDP, pad;        // aux layout: 0 = the 20-state kernel's; else the padded state count of cl_tree_dmma_kernel
    long long tOff;        // into the staged doubles: t[cat] effective branch lengths
};

struct CLChild {
    const double *cl;     // internal child: its CL
    const double *P;      // internal child: its P deck
    const uint8_t *tips;  // leaf child: tip code indices of its sequence
    const double *tbl;    // leaf child: its lookup table
};

struct CLArgs {
    double *out;
    int nChildren;
    int accumulate;   // multiply into the existing CL (continuation of a >kMaxChildren node)
    int ps;           // pattern stride (multiple of 32)
    int dim, nCat, tblW;
    CLChild ch[kMaxChildren];
};

struct LikeArgs {
    const double *cl;          // root CL
    const int *counts;
    const uint64_t *invarMask;
    const uint8_t *rootTips;   // non-NULL when the root is a leaf
    const uint64_t *eqMask;    // masks of the non-N-like equates
    double *patLikes;          // optional [ps]
    double *partials;          // [2*gridDim.x]: sum count*log(like), count of like <= 0
    const int *rootScale;      // optional [ps]: the root CL of pattern p carries a factor 2^(256*rootScale[p])
    int ps, nPat, dim, nCat;
    double pInvar;
    double pi[64];
};

// ---------------------------------------------------------------------------
// Per-pattern scalers (opt-in, p4b_setScalers).  The reference has none and
// returns -1e99 once a site likelihood underflows (Pf/p4_tree.c:1182).  When
// enabled, a node whose largest CL entry for a pattern falls below 2^-256 has
// that pattern's entries multiplied by 2^256 (exact in binary floating point)
// and the pattern's exponent count e incremented; e is summed up the tree and
// the root turns it back into  log(like) - e*256*ln2.  Patterns that never
// trigger a rescale produce bit-identical likelihoods with scalers on or off.
// ---------------------------------------------------------------------------
#define P4B_SCALE_THRESHOLD 8.636168555094445e-78    /* 2^-256 */
#define P4B_SCALE_FACTOR 1.157920892373162e+77       /* 2^256  */
#define P4B_SCALE_LOG 177.445678223345993            /* 256*ln(2) */

// count * log-likelihood of one pattern from its (possibly scaled) mixture sum A.
// Returns false when the site likelihood is not positive.
__device__ __forceinline__ bool like_term(double A, int e, double pInvar, int nCat, uint64_t invarBits, const double *pi, int dim,
                                          int count, double *term, double *likeOut)
{
    double like;
    if (pInvar != 0.0) {
        A *= (1.0 - pInvar) / (double)nCat;          // freqsTimesOneMinusPInvar[0], Pf/p4_tree.c:992-996
        if (invarBits) {
            // the constant-site term is unscaled: bring A back down first, then add the terms one
            // by one in state order like the reference (Pf/p4_tree.c:1073-1091)
            like = e ? scalbn(A, -256 * e) : A;
            for (int s = 0; s < dim; s++)
                if ((invarBits >> s) & 1ull) like += pi[s] * pInvar;
            *likeOut = like;
            if (like <= 0.0) return false;
            *term = (double)count * log(like);
            return true;
        }
        like = A;
    } else {
        like = nCat > 1 ? A / (double)nCat : A;
    }
    *likeOut = e ? scalbn(like, -256 * e) : like;
    if (like <= 0.0) return false;
    *term = (double)count * (e ? log(like) - (double)e * P4B_SCALE_LOG : log(like));
    return true;
}

struct RescaleArgs {
    double *cl;            // the node's CL, rescaled in place
    int *scOut;            // its exponent counts [ps]
    const int *scChild[16];
    int nScChild, nRows, ps;
};

// Node-level paths with scalers on: one extra pass over the node's CL.
__global__ void __launch_bounds__(256)
rescale_kernel(const RescaleArgs a)
{
    const int pat = blockIdx.x * blockDim.x + threadIdx.x;
    if (pat >= a.ps) return;
    int e = 0;
    for (int c = 0; c < a.nScChild; c++) e += a.scChild[c][pat];
    double m = 0.0;
    for (int k = 0; k < a.nRows; k++) m = fmax(m, a.cl[(size_t)k * a.ps + pat]);
    if (m < P4B_SCALE_THRESHOLD && m > 0.0) {
        for (int k = 0; k < a.nRows; k++) a.cl[(size_t)k * a.ps + pat] *= P4B_SCALE_FACTOR;
        e += 1;
    }
    a.scOut[pat] = e;
}

// ---------------------------------------------------------------------------
// P(t)
// ---------------------------------------------------------------------------
// Doubles of shared memory a job wants so that nothing on its paths comes from global memory: exp(lambda t) per category, the
// eigensystem, the P deck and the leaf table (the table and operand-deck passes read P and the table back).
__host__ __device__ inline size_t pmatrixStageDoubles(int dim, int nCat, int W) { return (size_t)nCat * dim + 2 * (size_t)dim * dim + dim + (size_t)nCat * dim * dim + (size_t)nCat * dim * W; }

__global__ void __launch_bounds__(128)
pmatrix_kernel(const PJob *__restrict__ jobs, const double *__restrict__ staged, const int smCap)
{
    extern __shared__ double sExp[];   // [nCat][dim] exp(lambda_k * t_cat) | (staged jobs) V | Vinv | lambda | P | table
    const PJob job = jobs[blockIdx.x];
    const int dim = job.dim, nCat = job.nCat;
    // a job that fits the launch's shared memory (every 4- and 20-state part does) keeps its eigensystem, its P deck and its
    // leaf table there: a lone 20-state job -- one changed branch of an MCMC proposal, one Newton step -- took 35 us with
    // every pass going through global memory
    const bool st = pmatrixStageDoubles(dim, nCat, job.tblW) <= (size_t)smCap;
    double *sEig = sExp + nCat * dim, *sP = sEig + 2 * dim * dim + dim, *sT = sP + nCat * dim * dim;
    if (st) {
        for (int i = threadIdx.x; i < 2 * dim * dim + dim; i += blockDim.x) sEig[i] = job.eig[i];
        __syncthreads();
    }
    const double *V = st ? sEig : job.eig;
    const double *Vi = V + dim * dim;
    const double *lam = Vi + dim * dim;
    const double *t = staged + job.tOff;
    for (int i = threadIdx.x; i < nCat * dim; i += blockDim.x) sExp[i] = exp(lam[i % dim] * t[i / dim]);
    __syncthreads();
    double *Pg = job.P;
    const int n = nCat * dim * dim;
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const int c = idx / (dim * dim), ij = idx - c * dim * dim, i = ij / dim, j = ij - i * dim;
        double sum = 0.0;
        // same association as the reference: sum += (V[i][k] * Vinv[k][j]) * exp(.)
        for (int k = 0; k < dim; k++) sum = fma(__dmul_rn(V[i * dim + k], Vi[k * dim + j]), sExp[c * dim + k], sum);
        Pg[idx] = sum;
        if (st) sP[idx] = sum;
    }
    const double *P = st ? sP : Pg;        // what the later passes read
    const double *T = st ? sT : job.tbl;
    if (job.tblW > 0) {
        __syncthreads();   // the block's own writes of P are visible after the barrier
        const int W = job.tblW;
        double *Tg = job.tbl;
        const uint64_t *em = job.eq;
        const int nT = nCat * dim * W;
        for (int idx = threadIdx.x; idx < nT; idx += blockDim.x) {
            const int k = idx / W, w = idx - k * W;   // k = cat*dim + s
            double v;
            if (w < dim) v = P[k * dim + w];
            else if (w == dim) v = 1.0;
            else {
                const uint64_t m = em[w - dim - 1];
                v = 0.0;
                for (int x = 0; x < dim; x++)
                    if ((m >> x) & 1ull) v += P[k * dim + x];
            }
            Tg[idx] = v;
            if (st) sT[idx] = v;
        }
    }
    if (job.aux && job.auxDP > 0) {
        // Operand decks of cl_tree_dmma_kernel (tree_dmma.cuh), padded state count DP:
        //   aux[0 .. nCat*F), F = (DP/4)*(DP/8)*32   P^T in mma fragment order [cat][kk][nt][lane]: lane (g,q) holds
        //                                            P[cat][8nt+g][x], x = 8*(kk>>1) + 2q + (kk&1); zero beyond dim
        //   aux[nCat*F .. +nCat*W*DP)                leaf table transposed and padded, [cat][w][DP]; zero beyond dim
        __syncthreads();
        const int DP = job.auxDP, NT = DP / 8, KS = DP / 4, F = KS * NT * 32;
        double *A = job.aux;
        for (int i = threadIdx.x; i < nCat * F; i += blockDim.x) {
            const int l = i & 31, nt = (i >> 5) % NT, kk = (i / (32 * NT)) % KS, ct = i / F;
            const int s = 8 * nt + (l >> 2), x = 8 * (kk >> 1) + 2 * (l & 3) + (kk & 1);
            A[i] = (s < dim && x < dim) ? P[(ct * dim + s) * dim + x] : 0.0;
        }
        if (job.tblW > 0) {
            const int W = job.tblW;
            double *TT = A + (size_t)nCat * F;
            const int nT = nCat * W * DP;
            for (int i = threadIdx.x; i < nT; i += blockDim.x) {
                const int st = i % DP, w = (i / DP) % W, ct = i / (DP * W);
                TT[i] = st < dim ? T[(ct * dim + st) * W + w] : 0.0;
            }
        }
    } else if (job.aux) {
        // Operand decks of cl_tree_aa_kernel, so that its per-step staging is a straight copy:
        //   aux[0 .. nCat*480)            P^T in mma fragment order [cat][kk][nt][lane]: lane (g,q) holds P[cat][8nt+g][x],
        //                                 x = 8t+2q+i for kk = 2t+i < 4, x = 16+q for kk = 4; zero where the parent state is >= 20
        //   aux[nCat*480 .. +nCat*W*24)   leaf table transposed and padded with zeros, [cat][w][24]
        __syncthreads();
        double *A = job.aux;
        const int nF = nCat * 480;
        for (int i = threadIdx.x; i < nF; i += blockDim.x) {
            const int l = i & 31, nt = (i >> 5) % 3, kk = (i / 96) % 5, ct = i / 480;
            const int s = 8 * nt + (l >> 2), x = kk < 4 ? 8 * (kk >> 1) + 2 * (l & 3) + (kk & 1) : 16 + (l & 3);
            A[i] = (s < dim && x < dim) ? P[(ct * dim + s) * dim + x] : 0.0;
        }
        if (job.tblW > 0) {
            const int W = job.tblW;
            double *TT = A + nF;
            const int nT = nCat * W * 24;
            for (int i = threadIdx.x; i < nT; i += blockDim.x) {
                const int st = i % 24, w = (i / 24) % W, ct = i / (24 * W);
                TT[i] = st < dim ? T[(ct * dim + st) * W + w] : 0.0;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// CL, 4 states.  One thread owns two adjacent patterns and all NCAT*4 entries;
// every global access is a 16-byte vector, 512 contiguous bytes per warp.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double warpSum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------
// In-kernel all-reduce of (partial lnL, count of non-positive likelihoods) across the pattern shards of a multi-GPU run
// (the path's one exchange step).  Every rank owns a small mailbox in its HBM that every other rank has
// mapped through CUDA IPC (csrc/comm.cpp commOpenPeerMailboxes).  The thread that has folded its rank's partials writes
// them straight into every peer's mailbox over NVLink (plain stores to peer memory, a system-scope fence, then the
// sequence number as the "ready" flag), waits until the entries of all ranks carry this evaluation's sequence number
// in its own mailbox, and sums them in rank order -- the same order on every rank, so all ranks return the same bits.
// No NCCL launch, no extra kernel: the exchange rides in the epilogue of the kernel that produced the partials.
// Entry = {sum, bad, seq, -}; index ((seq % 4) * kMaxBatchTrees + tree) * world + source rank.
// ---------------------------------------------------------------------------
constexpr int kMailSlots = 4, kMailTrees = 16, kMailMaxWorld = 8;
struct MailArgs {
    double *mine;
    double *peer[kMailMaxWorld];
    unsigned long long seq;
    int rank, world;      // world <= 1: no exchange
};
__device__ __noinline__ void mail_allreduce(const MailArgs &m, int tree, double &sum, double &bad)   // ONE thread calls this
{
    const size_t base = ((size_t)(m.seq & (kMailSlots - 1)) * kMailTrees + tree) * m.world;
    for (int r = 0; r < m.world; r++) {
        volatile double *dst = m.peer[r] + (base + m.rank) * 4;
        dst[0] = sum;
        dst[1] = bad;
    }
    __threadfence_system();
    for (int r = 0; r < m.world; r++)
        *(reinterpret_cast<volatile unsigned long long *>(m.peer[r] + (base + m.rank) * 4) + 2) = m.seq;
    double s = 0.0, b = 0.0;
    for (int r = 0; r < m.world; r++) {
        volatile double *src = m.mine + (base + r) * 4;
        volatile unsigned long long *flag = reinterpret_cast<volatile unsigned long long *>(m.mine + (base + r) * 4) + 2;
        long spins = 0;
        while (*flag != m.seq)
            if (++spins > (1L << 28)) __trap();      // a rank that never evaluates must be an error, not a hang
        __threadfence_system();
        s += src[0];
        b += src[1];
    }
    sum = s;
    bad = b;
}

__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }

template <int NCAT>
__global__ void __launch_bounds__(256, 3)
cl_dna_kernel(const CLArgs a)
{
    constexpr int K = NCAT * 4;
    extern __shared__ double sm[];   // per child: P[NCAT][4][4] (internal) or T[K][tblW] (leaf)
    const int W = a.tblW;
    const int perChild = (K * W > K * 4) ? K * W : K * 4;
    for (int c = 0; c < a.nChildren; c++) {
        const bool leaf = a.ch[c].tips != nullptr;
        const double *src = leaf ? a.ch[c].tbl : a.ch[c].P;
        const int n = leaf ? K * W : K * 4;
        for (int i = threadIdx.x; i < n; i += blockDim.x) sm[c * perChild + i] = src[i];
    }
    __syncthreads();
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    const int pat = pair * 2;
    if (pat >= a.ps) return;
    const size_t ps = (size_t)a.ps;

    // tip codes of the leaf children, two patterns per thread
    uchar2 code[kMaxChildren];
#pragma unroll
    for (int c = 0; c < kMaxChildren; c++)
        if (c < a.nChildren && a.ch[c].tips != nullptr) code[c] = *reinterpret_cast<const uchar2 *>(a.ch[c].tips + pat);

    // rate category outermost: 4 accumulators live at a time, stored as soon as done
#pragma unroll
    for (int cat = 0; cat < NCAT; cat++) {
        double2 acc[4];
        if (a.accumulate) {
#pragma unroll
            for (int st = 0; st < 4; st++) acc[st] = ld2(a.out + (cat * 4 + st) * ps + pat);
        }
#pragma unroll
        for (int c = 0; c < kMaxChildren; c++) {
            if (c < a.nChildren) {
                const bool first = (c == 0) && !a.accumulate;
                const double *s = sm + c * perChild;
                if (a.ch[c].tips == nullptr) {
                    const double *cl = a.ch[c].cl + pat;
                    const double2 v0 = ld2(cl + (cat * 4 + 0) * ps);
                    const double2 v1 = ld2(cl + (cat * 4 + 1) * ps);
                    const double2 v2 = ld2(cl + (cat * 4 + 2) * ps);
                    const double2 v3 = ld2(cl + (cat * 4 + 3) * ps);
#pragma unroll
                    for (int st = 0; st < 4; st++) {
                        const double2 p01 = *reinterpret_cast<const double2 *>(s + cat * 16 + st * 4);
                        const double2 p23 = *reinterpret_cast<const double2 *>(s + cat * 16 + st * 4 + 2);
                        double2 sum;
                        sum.x = p01.x * v0.x;
                        sum.y = p01.x * v0.y;
                        sum.x = fma(p01.y, v1.x, sum.x);
                        sum.y = fma(p01.y, v1.y, sum.y);
                        sum.x = fma(p23.x, v2.x, sum.x);
                        sum.y = fma(p23.x, v2.y, sum.y);
                        sum.x = fma(p23.y, v3.x, sum.x);
                        sum.y = fma(p23.y, v3.y, sum.y);
                        if (first) acc[st] = sum;
                        else { acc[st].x *= sum.x; acc[st].y *= sum.y; }
                    }
                } else {
#pragma unroll
                    for (int st = 0; st < 4; st++) {
                        const int k = cat * 4 + st;
                        const double fx = s[k * W + code[c].x], fy = s[k * W + code[c].y];
                        if (first) { acc[st].x = fx; acc[st].y = fy; }
                        else { acc[st].x *= fx; acc[st].y *= fy; }
                    }
                }
            }
        }
#pragma unroll
        for (int st = 0; st < 4; st++) st2(a.out + (cat * 4 + st) * ps + pat, acc[st]);
    }
}

// ---------------------------------------------------------------------------
// Whole-tree CL recursion, 4 states, ONE launch.
//
// Patterns are independent, so a thread that owns two patterns can walk the
// whole list of internal nodes (in the caller's post-order) by itself: every
// child CL it needs was written earlier by the same thread.  A CTA owns a tile
// of 2*blockDim.x patterns and executes the step list; per step it
//   - stages the children's P decks / leaf tables into shared memory
//     (double-buffered, one __syncthreads per step),
//   - takes the CL of the child computed in the previous step straight from
//     registers (in post-order a node follows its last internal child),
//   - loads any other internal child from global memory -- recently written by
//     this CTA, so mostly served by the 126 MB L2 rather than HBM,
//   - writes the node's CL to HBM (every node's CL stays resident, exactly the
//     state the reference leaves behind), and
//   - after the last step optionally folds the root CL into the site
//     likelihoods and the block's partial lnL (Pf/p4_tree.c:1029-1197).
// The same kernel serves any ordered subset of nodes (a dirty path).
// ---------------------------------------------------------------------------
// One step = one node (or one chunk of a node with more than kMaxChildren
// children).  Compact, because the whole step list travels in kernel-parameter
// (constant) memory: no descriptor ever has to be fetched from global memory.
struct StepC {
    int outSlot;                 // CL arena slot of the node
    short nChildren;
    signed char first;           // 1: the running product starts at 1; 0: continues from the previous chunk
    signed char store;           // 1: write the CL at the end of the step
    struct { int a, b; } ch[kMaxChildren];
    // outSlot and an internal child's index address a CL buffer as (buffer - hdr.arena) / 256 bytes: buffers are
    // 256-byte aligned, and 30 bits of such units span 274 GB -- more than the device has.  With scalers on, the
    // buffer's per-pattern exponents (int32[ps]) sit right behind its K*ps doubles.
    // a = kind << 30 | index; kind 0: internal child, load CL slot `index`
    //                         kind 1: internal child, CL in registers (computed by the previous step)
    //                         kind 2: leaf child, tip row `index` (its seqNum)
    // b = node number, addressing its P deck / leaf table
};

// Several trees that share a data part (the cur/prop trees of Metropolis-coupled chains) can be
// evaluated by ONE launch: blockIdx.y selects the tree, i.e. its header and its range of the step list.
struct TreeHdr {
    double *arena;            // base address the tree's CL buffers are addressed from (covers its own arena and its twin's)
    const double *Pdeck;      // tree's P decks, already offset to this part
    const double *tbl;        // tree's leaf tables, already offset to this part
    const double *aux;        // 20-state kernel: tree's fragment-order decks, already offset to this part
    double *patLikes;         // optional
    double *partials;         // [2*gridDim.x]
    const uint8_t *rootTips;  // non-NULL when the root is a leaf
    double pInvar;
    double pi[4];
    int stepBase, nSteps;     // this tree's steps are steps[stepBase .. stepBase+nSteps)
    int doLike, pad;          // fused root reduction
};
constexpr int kMaxBatchTrees = 16;
constexpr int kMaxSteps = 500;   // 56 B each; with the headers the argument block stays under the 32 KB limit

struct TreeArgs {
    int ps, nPat, tblW, nTrees;
    int maxKids, pad0;        // most children any step of this launch has: sizes the staging buffers
    long long pNodeDoubles;   // stride between nodes in a P deck
    long long tblNodeDoubles;
    long long auxNodeDoubles;
    const uint8_t *tips;      // part's tip rows [nTax][ps]
    const int *counts;
    const uint64_t *invarMask;
    const uint64_t *eqMask;
    TreeHdr hdr[kMaxBatchTrees];
    StepC steps[kMaxSteps];
};
static_assert(sizeof(TreeArgs) <= 32764, "kernel argument block too large");

__device__ __forceinline__ void cp_async8(double *smemDst, const double *gmemSrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmemSrc));
}
__device__ __forceinline__ void cp_async16(void *smemDst, const void *gmemSrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmemSrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Factor contributed by one child to the 4 states of one rate category, for the
// thread's two patterns.  KIND is a compile-time constant: 0 internal child read
// from global memory, 1 internal child whose CL is in `cur`, 2 leaf child,
// 3 internal child prefetched into this thread's shared-memory slots.
template <int KIND>
__device__ __forceinline__ void child_factor(int cat, const double *__restrict__ s, int W, unsigned cx, unsigned cy,
                                             const double2 *cur, const double *__restrict__ cl, unsigned ps, double2 f[4])
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
        if (KIND == 1) {
            v0 = cur[cat * 4 + 0]; v1 = cur[cat * 4 + 1]; v2 = cur[cat * 4 + 2]; v3 = cur[cat * 4 + 3];
        } else if (KIND == 3) {   // `cl` points at this thread's slot of the prefetch buffer, row stride `ps` double2
            const double2 *pre = reinterpret_cast<const double2 *>(cl);
            v0 = pre[(cat * 4 + 0) * ps]; v1 = pre[(cat * 4 + 1) * ps]; v2 = pre[(cat * 4 + 2) * ps]; v3 = pre[(cat * 4 + 3) * ps];
        } else {
            v0 = ld2(cl + (cat * 4 + 0) * ps);
            v1 = ld2(cl + (cat * 4 + 1) * ps);
            v2 = ld2(cl + (cat * 4 + 2) * ps);
            v3 = ld2(cl + (cat * 4 + 3) * ps);
        }
#pragma unroll
        for (int s4 = 0; s4 < 4; s4++) {
            const double2 p01 = *reinterpret_cast<const double2 *>(s + cat * 16 + s4 * 4);
            const double2 p23 = *reinterpret_cast<const double2 *>(s + cat * 16 + s4 * 4 + 2);
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

// A node with exactly two children -- nearly every node of a binary tree --
// as straight-line code for one combination of child kinds.
template <int NCAT, int K0, int K1, int THREADS, bool STORE>
__device__ __forceinline__ void step_two_children(double2 *cur, const double *__restrict__ s0, const double *__restrict__ s1, int W,
                                                  unsigned c0, unsigned c1, const double *__restrict__ cl0,
                                                  const double *__restrict__ cl1, unsigned ps, double *__restrict__ out)
{
#pragma unroll
    for (int cat = 0; cat < NCAT; cat++) {
        double2 f0[4], f1[4];
        child_factor<K0>(cat, s0, W, c0 & 0xffu, (c0 >> 8) & 0xffu, cur, cl0, K0 == 3 ? (unsigned)THREADS : ps, f0);
        child_factor<K1>(cat, s1, W, c1 & 0xffu, (c1 >> 8) & 0xffu, cur, cl1, K1 == 3 ? (unsigned)THREADS : ps, f1);
#pragma unroll
        for (int s4 = 0; s4 < 4; s4++) {
            double2 r;
            r.x = f0[s4].x * f1[s4].x;     // (left child) * (sibling), the reference's order
            r.y = f0[s4].y * f1[s4].y;
            cur[cat * 4 + s4] = r;
            if (STORE) st2(out + (cat * 4 + s4) * ps, r);
        }
    }
}

template <int NCAT, int THREADS, bool SCALE>
__device__ __forceinline__ void cl_tree_dna_body(const TreeArgs &a)
{
    constexpr int K = NCAT * 4;
    const TreeHdr &hd = a.hdr[blockIdx.y];     // uniform across the CTA: stays in parameter (constant) memory
    extern __shared__ double sm[];            // 2 buffers x maxKids x perChild, then the prefetch slots
    __shared__ double sSum[THREADS / 32], sBad[THREADS / 32];
    const int W = a.tblW;
    const int perChild = K * (W > 4 ? W : 4);
    const int bufSize = a.maxKids * perChild;
    const int pat = (blockIdx.x * THREADS + threadIdx.x) * 2;
    const bool active = pat < a.ps;
    const unsigned ps = (unsigned)a.ps;

    // children's P decks / leaf tables of one step -> shared memory, asynchronously
    auto stage = [&](int stepIdx, double *buf) {
        const StepC &st = a.steps[hd.stepBase + stepIdx];
        const int nc = st.nChildren;
        for (int c = 0; c < nc; c++) {
            const bool leaf = ((unsigned)st.ch[c].a >> 30) == 2u;
            const double *src = leaf ? hd.tbl + a.tblNodeDoubles * st.ch[c].b : hd.Pdeck + a.pNodeDoubles * st.ch[c].b;
            const int n = leaf ? K * W : K * 4;
            for (int i = threadIdx.x; i < n; i += THREADS) cp_async8(buf + c * perChild + i, src + i);
        }
        cp_async_commit();
    };
    // tip codes (two patterns = 16 bits) of child c of a step, if it is a leaf
    auto tipCode = [&](int stepIdx, int c) -> unsigned {
        if (!active || stepIdx >= hd.nSteps) return 0u;
        const StepC &st = a.steps[hd.stepBase + stepIdx];
        const unsigned av = (unsigned)st.ch[c].a;
        if (c >= st.nChildren || (av >> 30) != 2u) return 0u;
        return *reinterpret_cast<const unsigned short *>(a.tips + (size_t)(av & 0x3fffffffu) * ps + pat);
    };

    // Prefetch buffer: the one internal child of a two-children step that is not in registers
    // (kind 3) is copied by cp.async into slots private to each thread, one step ahead, so its
    // L2 / HBM latency is paid while the previous step computes.  [K][THREADS] double2.
    double2 *pre = reinterpret_cast<double2 *>(sm + 2 * bufSize) + threadIdx.x;
    auto usesPre = [&](int stepIdx) -> bool {
        const StepC &st = a.steps[hd.stepBase + stepIdx];
        return st.nChildren == 2 && ((((unsigned)st.ch[0].a >> 30) == 3u) || (((unsigned)st.ch[1].a >> 30) == 3u));
    };
    auto prefetch = [&](int stepIdx) {
        if (!active) return;
        const StepC &st = a.steps[hd.stepBase + stepIdx];
        const unsigned av = (((unsigned)st.ch[0].a >> 30) == 3u) ? (unsigned)st.ch[0].a : (unsigned)st.ch[1].a;
        const double *cl = hd.arena + (size_t)(av & 0x3fffffffu) * 32 + pat;
#pragma unroll
        for (int k = 0; k < K; k++) cp_async16(pre + k * THREADS, cl + (size_t)k * ps);
    };

    double2 cur[K];   // CL of the node computed by the previous step (this thread's two patterns)
#pragma unroll
    for (int k = 0; k < K; k++) cur[k] = make_double2(1.0, 1.0);
    int2 ecur = make_int2(0, 0);   // its scaler exponents (SCALE only)

    if (usesPre(0)) prefetch(0);
    stage(0, sm);
    unsigned next0 = tipCode(0, 0), next1 = tipCode(0, 1);   // children 0 and 1 are prefetched one step ahead
    for (int si = 0; si < hd.nSteps; si++) {
        cp_async_wait_all();
        __syncthreads();   // buffer si&1 is complete; every thread is done with step si-1
        const double *buf = sm + (si & 1) * bufSize;
        const bool nextWantsPre = si + 1 < hd.nSteps && usesPre(si + 1), curUsesPre = usesPre(si);
        if (nextWantsPre && !curUsesPre) prefetch(si + 1);    // joins the commit group of stage() below
        if (si + 1 < hd.nSteps) stage(si + 1, sm + ((si + 1) & 1) * bufSize);
        else cp_async_commit();
        const unsigned code0 = next0, code1 = next1;
        next0 = tipCode(si + 1, 0);          // in flight while this step computes
        next1 = tipCode(si + 1, 1);
        if (active) {
            const StepC &st = a.steps[hd.stepBase + si];
            const int nc = st.nChildren;
            double *out = hd.arena + (size_t)(unsigned)st.outSlot * 32 + pat;
            const unsigned a0 = (unsigned)st.ch[0].a, a1 = (unsigned)st.ch[1].a;
            const unsigned k0 = a0 >> 30, k1 = a1 >> 30;
            int2 esum = make_int2(0, 0);
            if (SCALE) {     // exponents of the internal children add up
                if (!st.first) esum = ecur;
                for (int c = 0; c < nc; c++) {
                    const unsigned av = (unsigned)st.ch[c].a, kind = av >> 30;
                    if (kind == 1u) { esum.x += ecur.x; esum.y += ecur.y; }
                    else if (kind != 2u) {
                        const int2 ec = *reinterpret_cast<const int2 *>(reinterpret_cast<const int *>(hd.arena + (size_t)(av & 0x3fffffffu) * 32 + (size_t)K * ps) + pat);
                        esum.x += ec.x;
                        esum.y += ec.y;
                    }
                }
            }
            if (nc == 2 && st.first && st.store && k0 != 0u && k1 != 0u) {
                const double *pre0 = reinterpret_cast<const double *>(pre);
                const double *s0 = buf, *s1 = buf + perChild;
                switch (k0 * 4 + k1) {   // uniform across the CTA; kinds 1 (registers), 2 (leaf), 3 (prefetched)
                case 1 * 4 + 2: step_two_children<NCAT, 1, 2, THREADS, !SCALE>(cur, s0, s1, W, code0, code1, nullptr, nullptr, ps, out); break;
                case 2 * 4 + 1: step_two_children<NCAT, 2, 1, THREADS, !SCALE>(cur, s0, s1, W, code0, code1, nullptr, nullptr, ps, out); break;
                case 1 * 4 + 3: step_two_children<NCAT, 1, 3, THREADS, !SCALE>(cur, s0, s1, W, code0, code1, nullptr, pre0, ps, out); break;
                case 3 * 4 + 1: step_two_children<NCAT, 3, 1, THREADS, !SCALE>(cur, s0, s1, W, code0, code1, pre0, nullptr, ps, out); break;
                case 2 * 4 + 3: step_two_children<NCAT, 2, 3, THREADS, !SCALE>(cur, s0, s1, W, code0, code1, nullptr, pre0, ps, out); break;
                case 3 * 4 + 2: step_two_children<NCAT, 3, 2, THREADS, !SCALE>(cur, s0, s1, W, code0, code1, pre0, nullptr, ps, out); break;
                default: step_two_children<NCAT, 2, 2, THREADS, !SCALE>(cur, s0, s1, W, code0, code1, nullptr, nullptr, ps, out); break;
                }
            } else {
                // any other shape (1 child, 3+ children, chunks of a wide polytomy): generic loop
                const bool first = st.first != 0, store = st.store != 0;
#pragma unroll
                for (int cat = 0; cat < NCAT; cat++) {
                    double2 acc[4];
#pragma unroll
                    for (int s4 = 0; s4 < 4; s4++) acc[s4] = first ? make_double2(1.0, 1.0) : cur[cat * 4 + s4];
                    for (int c = 0; c < nc; c++) {
                        const unsigned av = (unsigned)st.ch[c].a;
                        const unsigned kind = av >> 30;
                        const double *s = buf + c * perChild;
                        double2 f[4];
                        if (kind == 2u) {
                            const unsigned cc = c == 0 ? code0 : (c == 1 ? code1 : tipCode(si, c));
                            child_factor<2>(cat, s, W, cc & 0xffu, (cc >> 8) & 0xffu, cur, nullptr, ps, f);
                        } else if (kind == 1u) {
                            child_factor<1>(cat, s, W, 0u, 0u, cur, nullptr, ps, f);
                        } else {
                            child_factor<0>(cat, s, W, 0u, 0u, cur, hd.arena + (size_t)(av & 0x3fffffffu) * 32 + pat, ps, f);
                        }
#pragma unroll
                        for (int s4 = 0; s4 < 4; s4++) { acc[s4].x *= f[s4].x; acc[s4].y *= f[s4].y; }
                    }
#pragma unroll
                    for (int s4 = 0; s4 < 4; s4++) {
                        cur[cat * 4 + s4] = acc[s4];
                        if (!SCALE && store) st2(out + (cat * 4 + s4) * ps, acc[s4]);
                    }
                }
            }
            if (SCALE) {
                double mx = 0.0, my = 0.0;
#pragma unroll
                for (int k = 0; k < K; k++) { mx = fmax(mx, cur[k].x); my = fmax(my, cur[k].y); }
                if (mx < P4B_SCALE_THRESHOLD && mx > 0.0) {
#pragma unroll
                    for (int k = 0; k < K; k++) cur[k].x *= P4B_SCALE_FACTOR;
                    esum.x += 1;
                }
                if (my < P4B_SCALE_THRESHOLD && my > 0.0) {
#pragma unroll
                    for (int k = 0; k < K; k++) cur[k].y *= P4B_SCALE_FACTOR;
                    esum.y += 1;
                }
                ecur = esum;
                if (st.store) {
#pragma unroll
                    for (int k = 0; k < K; k++) st2(out + k * ps, cur[k]);
                    *reinterpret_cast<int2 *>(reinterpret_cast<int *>(out - pat + (size_t)K * ps) + pat) = ecur;
                }
            }
        }
        if (nextWantsPre && curUsesPre) {   // the buffer was in use by this step: refill it now
            prefetch(si + 1);
            cp_async_commit();
        }
    }

    if (!hd.doLike) return;
    // ---- root reduction from registers ----------------------------------------
    double term = 0.0, bad = 0.0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int p1 = pat + h;
        if (active && p1 < a.nPat) {
            uint64_t mask = ~0ull;
            if (hd.rootTips) {
                const int w = hd.rootTips[p1];
                if (w < 4) mask = 1ull << w;
                else if (w > 4) mask = a.eqMask[w - 5];
            }
            double A = 0.0;
#pragma unroll
            for (int k = 0; k < K; k++) {
                const double v = h ? cur[k].y : cur[k].x;
                if ((mask >> (k & 3)) & 1ull) A = fma(hd.pi[k & 3], v, A);
            }
            const int e = SCALE ? (h ? ecur.y : ecur.x) : 0;
            const uint64_t im = (hd.pInvar != 0.0 && a.invarMask) ? a.invarMask[p1] : 0ull;
            double t1 = 0.0, like = 0.0;
            if (like_term(A, e, hd.pInvar, NCAT, im, hd.pi, 4, a.counts[p1], &t1, &like)) term += t1;
            else bad += 1.0;
            if (hd.patLikes) hd.patLikes[p1] = like;
        }
    }
    term = warpSum(term);
    bad = warpSum(bad);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sSum[w] = term; sBad[w] = bad; }
    __syncthreads();
    if (w == 0) {
        term = (l < THREADS / 32) ? sSum[l] : 0.0;
        bad = (l < THREADS / 32) ? sBad[l] : 0.0;
        term = warpSum(term);
        bad = warpSum(bad);
        if (l == 0) {
            hd.partials[2 * blockIdx.x] = term;
            hd.partials[2 * blockIdx.x + 1] = bad;
        }
    }
}

// Two ways of fixing the residency of the same body: MINB CTAs per SM (the compiler derives the register
// budget), or an explicit register budget MAXREG (144 registers = 14 warps per SM: one wave for a 125 k-pattern shard).
template <int NCAT, int THREADS, int MINB, bool SCALE>
__global__ void __launch_bounds__(THREADS, MINB)
cl_tree_dna_kernel(const __grid_constant__ TreeArgs a)
{
    cl_tree_dna_body<NCAT, THREADS, SCALE>(a);
}
template <int NCAT, int THREADS, int MAXREG>
__global__ void __launch_bounds__(THREADS) __maxnreg__(MAXREG)
cl_tree_dna_kernel_r(const __grid_constant__ TreeArgs a)
{
    cl_tree_dna_body<NCAT, THREADS, false>(a);
}

// ---------------------------------------------------------------------------
// CL, DIM states held in registers.  grid.y = rate category; one thread owns
// one pattern of one category; P / leaf tables of that category sit in shared
// memory and are read as warp-wide broadcasts.
// ---------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(128)
cl_dim_kernel(const CLArgs a)
{
    extern __shared__ double sm[];   // per child: P[DIM][DIM] or T[DIM][tblW] of this category
    const int cat = blockIdx.y;
    const int W = a.tblW;
    const int perChild = (DIM * W > DIM * DIM) ? DIM * W : DIM * DIM;
    for (int c = 0; c < a.nChildren; c++) {
        const bool leaf = a.ch[c].tips != nullptr;
        const double *src = leaf ? a.ch[c].tbl + (size_t)cat * DIM * W : a.ch[c].P + (size_t)cat * DIM * DIM;
        const int n = leaf ? DIM * W : DIM * DIM;
        for (int i = threadIdx.x; i < n; i += blockDim.x) sm[c * perChild + i] = src[i];
    }
    __syncthreads();
    const int pat = blockIdx.x * blockDim.x + threadIdx.x;
    if (pat >= a.ps) return;
    const size_t ps = (size_t)a.ps;
    const size_t base = (size_t)cat * DIM * ps + pat;

    double acc[DIM];
    if (a.accumulate) {
#pragma unroll
        for (int s = 0; s < DIM; s++) acc[s] = a.out[base + s * ps];
    }
    for (int c = 0; c < a.nChildren; c++) {
        const bool first = (c == 0) && !a.accumulate;
        const double *sp = sm + c * perChild;
        if (a.ch[c].tips == nullptr) {
            double v[DIM];
            const double *cl = a.ch[c].cl + base;
#pragma unroll
            for (int x = 0; x < DIM; x++) v[x] = cl[x * ps];
#pragma unroll
            for (int s = 0; s < DIM; s++) {
                double sum = sp[s * DIM] * v[0];
#pragma unroll
                for (int x = 1; x < DIM; x++) sum = fma(sp[s * DIM + x], v[x], sum);
                acc[s] = first ? sum : acc[s] * sum;
            }
        } else {
            const int code = a.ch[c].tips[pat];
#pragma unroll
            for (int s = 0; s < DIM; s++) {
                const double f = sp[s * W + code];
                acc[s] = first ? f : acc[s] * f;
            }
        }
    }
#pragma unroll
    for (int s = 0; s < DIM; s++) a.out[base + s * ps] = acc[s];
}

// ---------------------------------------------------------------------------
// CL, 20 states, FP64 tensor cores (mma.sync m8n8k4 -- tcgen05 has no f64 kind).
//
// out[s][pat] = prod_children  sum_x P[s][x] * cl_child[x][pat]  is a (24 x 20) x
// (20 x 8) product per child, rate category and tile of 8 patterns:
//   A = P of the child, rows padded 20 -> 24 (3 m-tiles), 5 k-steps of 4 states;
//       its 15 fragments stay in registers for every pattern tile of the CTA
//   B = child CL tile, read straight from global memory in fragment order
//       (4 rows x 64 contiguous bytes per load)
//   C = the factor; factors of all children are multiplied element-wise in the
//       accumulator layout and stored (2 adjacent patterns per lane).
// Warp w handles rate categories w, w+4, ...; CTAs are persistent (3 per SM) and take
// groups of 32 patterns round-robin.
// At most two internal children per launch (their A fragments live in
// registers); wider nodes are chained by the host like in the other kernels.
// ---------------------------------------------------------------------------
constexpr int kDmmaGroup = 32;    // patterns a warp handles per iteration (4 n-tiles of 8)

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// Pattern <-> fragment mapping inside a group of 32 patterns starting at pat0:
// n-tile j (0..3), tile column n (0..7)  <->  pattern pat0 + 4*n + j.
//   B fragment of tile j: lane (g,q) holds row x0+q, column g -> pattern pat0 + 4g + j:
//       the 4 tiles of a lane are 4 CONSECUTIVE patterns (two 16-byte loads per row).
//   C fragment of tile j: lane (g,q) holds row s0+g, columns 2q+i -> patterns pat0 + 8q + 4i + j:
//       over j and i a lane owns the 8 consecutive patterns pat0 + 8q .. +7 (64-byte stores).
__global__ void __launch_bounds__(128, 3)
cl_dmma20_kernel(const __grid_constant__ CLArgs a)
{
    constexpr int DIM = 20;
    extern __shared__ double sm[];   // A fragments [intIdx][cat][15][32], then leaf tables [leafIdx][cat][DIM][W]
    const int W = a.tblW, nCat = a.nCat;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const size_t ps = (size_t)a.ps;
    int nLeaf = 0, nInt = 0;
    for (int c = 0; c < a.nChildren; c++) (a.ch[c].tips != nullptr ? nLeaf : nInt)++;
    const int fragSize = nCat * 15 * 32, tblSize = nCat * DIM * W;
    double *sT = sm + nInt * fragSize;
    {
        int ii = 0, li = 0;
        for (int c = 0; c < a.nChildren; c++) {
            if (a.ch[c].tips == nullptr) {
                const double *P = a.ch[c].P;
                for (int i = threadIdx.x; i < fragSize; i += blockDim.x) {
                    const int l = i & 31, f = (i >> 5) % 15, cat = i / (15 * 32);
                    const int mt = f / 5, ks = f - mt * 5, srow = mt * 8 + (l >> 2), x = ks * 4 + (l & 3);
                    sm[ii * fragSize + i] = srow < DIM ? __ldg(P + ((size_t)cat * DIM + srow) * DIM + x) : 0.0;
                }
                ii++;
            } else {
                for (int i = threadIdx.x; i < tblSize; i += blockDim.x) sT[li * tblSize + i] = __ldg(a.ch[c].tbl + i);
                li++;
            }
        }
    }
    __syncthreads();

    // persistent CTAs: groups of 32 patterns are dealt round-robin, so the staging above is
    // paid once per CTA and every SM ends within one group of the others
    const int nGroups = a.ps / kDmmaGroup;
    for (int cat = warp; cat < nCat; cat += 4) {
        for (int tg = blockIdx.x; tg < nGroups; tg += gridDim.x) {
            const int pat0 = tg * kDmmaGroup;
            double acc[24];   // [mt][j][i] -> mt*8 + j*2 + i
            if (a.accumulate) {
#pragma unroll
                for (int mt = 0; mt < 3; mt++) {
                    const int srow = mt * 8 + g;
                    const double *o = a.out + ((size_t)cat * DIM + (srow < DIM ? srow : 0)) * ps + pat0 + 8 * q;
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        const double2 v = ld2(o + 2 * h);   // patterns 8q + 2h, 8q + 2h + 1
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int off = 2 * h + e, i = off >> 2, j = off & 3;
                            acc[mt * 8 + j * 2 + i] = e ? v.y : v.x;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < 24; k++) acc[k] = 1.0;
            }
            int ii = 0, li = 0;
            for (int c = 0; c < a.nChildren; c++) {      // uniform runtime loop
                double f[24];
                if (a.ch[c].tips == nullptr) {
#pragma unroll
                    for (int k = 0; k < 24; k++) f[k] = 0.0;
                    const double *cl = a.ch[c].cl + (size_t)cat * DIM * ps + pat0 + 4 * g;
                    const double *As = sm + ii * fragSize + (size_t)cat * 15 * 32 + lane;
                    double2 b01[5], b23[5];   // all ten 16-byte loads of the tile group in flight together
#pragma unroll
                    for (int ks = 0; ks < 5; ks++) {
                        const double *row = cl + (size_t)(ks * 4 + q) * ps;
                        b01[ks] = ld2(row);
                        b23[ks] = ld2(row + 2);
                    }
#pragma unroll
                    for (int ks = 0; ks < 5; ks++) {
                        const double b[4] = {b01[ks].x, b01[ks].y, b23[ks].x, b23[ks].y};
#pragma unroll
                        for (int mt = 0; mt < 3; mt++) {
                            const double aF = As[(mt * 5 + ks) * 32];
#pragma unroll
                            for (int j = 0; j < 4; j++) dmma884(f[mt * 8 + j * 2], f[mt * 8 + j * 2 + 1], aF, b[j]);
                        }
                    }
                    ii++;
                } else {
                    const uint2 cw = *reinterpret_cast<const uint2 *>(a.ch[c].tips + pat0 + 8 * q);   // 8 tip codes
                    const double *T = sT + li * tblSize + (size_t)cat * DIM * W;
#pragma unroll
                    for (int mt = 0; mt < 3; mt++) {
                        const int srow = mt * 8 + g;
                        const double *Tr = T + (srow < DIM ? srow : 0) * W;
#pragma unroll
                        for (int off = 0; off < 8; off++) {
                            const unsigned code = ((off < 4 ? cw.x : cw.y) >> (8 * (off & 3))) & 0xffu;
                            const int i = off >> 2, j = off & 3;
                            f[mt * 8 + j * 2 + i] = Tr[code];
                        }
                    }
                    li++;
                }
#pragma unroll
                for (int k = 0; k < 24; k++) acc[k] *= f[k];
            }
#pragma unroll
            for (int mt = 0; mt < 3; mt++) {
                const int srow = mt * 8 + g;
                if (srow < DIM) {
                    double *o = a.out + ((size_t)cat * DIM + srow) * ps + pat0 + 8 * q;
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        const int o0 = 2 * h, o1 = 2 * h + 1;
                        st2(o + 2 * h, make_double2(acc[mt * 8 + (o0 & 3) * 2 + (o0 >> 2)], acc[mt * 8 + (o1 & 3) * 2 + (o1 >> 2)]));
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Whole-tree CL recursion, 20 states, FP64 tensor cores, ONE launch: the scheme (the kernel itself is cl_tree_aa2_kernel, tree_aa.cuh).
//
// The per-node kernel above computes out = P x cl_child with P as the A operand, so its result
// comes out of the tensor core in the C layout and would have to be transposed (shared memory or
// shuffles) before it could be the B operand of the parent's product.  Here the product is taken
// the other way round,  out^T = cl_child^T x P^T :
//   A (8 x 4)  = child CL, rows = patterns, columns = child states
//   B (4 x 8)  = P^T,      rows = child states, columns = parent states   (pre-arranged in shared memory)
//   C (8 x 8)  = parent CL, rows = patterns, columns = parent states
// and the summation index is dealt to the k-steps so that k-step (t, i), t = 0..1, i = 0..1, covers the
// child states {8t + 2q + i : q = 0..3}.  Lane (g, q) then needs, as its A element of k-step (t, i), the
// value (pattern g, state 8t + 2q + i) -- exactly what it holds as C element i of n-tile t after the child
// was computed.  A node's result is therefore the next node's operand with no data movement: a warp walks
// the whole step list with the running CL in registers, like the 4-state kernel does with FMAs.  The last
// four states (16..19, held two per lane by the lanes q = 0, 1 of n-tile 2) form a fifth k-step of their
// own, lane q <-> state 16 + q, gathered inside each quad by one shuffle.  Parent states are padded
// 20 -> 24 (zero columns in B): 15 DMMAs per child and 8 patterns, the same as an unpadded 24 x 20 product.
//
// A warp owns one rate category of one group of 8*MT patterns (categories never mix below the root);
// m-tile j, row g  <->  pattern pat0 + MT*g + j, so a lane's MT m-tiles are MT consecutive patterns of a
// row (16-byte accesses), and the eight g-lanes cover 64*MT contiguous bytes.  A CTA is NCAT x GROUPS
// warps; P^T fragments and leaf tables of a step are staged into shared memory one step ahead (cp.async,
// double-buffered) as straight copies: the P(t) kernel leaves P^T in fragment order and the leaf table
// transposed ([code][state]) in an operand deck of its own.  Steps have at most two children (wider nodes
// are chained by the host).  The root reduction is a separate kernel (like_kernel).
// ---------------------------------------------------------------------------
constexpr int kAAKids = 2;
constexpr int kAAFrag = 15 * 32;     // doubles of P^T fragments per child and category: 5 k-steps x 3 n-tiles x 32 lanes

__device__ __forceinline__ void named_barrier(int id, int nThreads)
{
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nThreads) : "memory");
}
// mbarrier + bulk (TMA) copy: one thread arms the barrier with the byte count and issues the copies;
// everybody waits on the barrier's phase parity.
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smemDst, const void *gmemSrc, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smemDst)),
                 "l"(gmemSrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done = 0;
    int spins = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done)
                     : "r"(addr), "r"(parity)
                     : "memory");
        if (!done && ++spins > (1 << 22)) __trap();   // a lost copy must be an error, never a hang
    } while (!done);
}

// ---------------------------------------------------------------------------
// CL, any dim.  grid.y = rate category; one thread per pattern; child values
// are re-read per parent state (served by L1).  Correctness path for unusual
// state counts (2, 6-state recodings, ...).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
cl_generic_kernel(const CLArgs a)
{
    const int cat = blockIdx.y;
    const int pat = blockIdx.x * blockDim.x + threadIdx.x;
    if (pat >= a.ps) return;
    const int dim = a.dim, W = a.tblW;
    const size_t ps = (size_t)a.ps;
    const size_t base = (size_t)cat * dim * ps + pat;
    for (int s = 0; s < dim; s++) {
        double prod = a.accumulate ? a.out[base + s * ps] : 1.0;
        for (int c = 0; c < a.nChildren; c++) {
            double f;
            if (a.ch[c].tips == nullptr) {
                const double *P = a.ch[c].P + ((size_t)cat * dim + s) * dim;
                const double *cl = a.ch[c].cl + base;
                f = __ldg(P) * cl[0];
                for (int x = 1; x < dim; x++) f = fma(__ldg(P + x), cl[x * ps], f);
            } else {
                f = __ldg(a.ch[c].tbl + ((size_t)cat * dim + s) * W + a.ch[c].tips[pat]);
            }
            prod = (c == 0 && !a.accumulate) ? f : prod * f;
        }
        a.out[base + s * ps] = prod;
    }
}

// ---------------------------------------------------------------------------
// Site likelihoods and the per-part log-likelihood.
// ---------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
like_kernel(const LikeArgs a)
{
    __shared__ double sSum[8], sBad[8];
    const int pat = blockIdx.x * blockDim.x + threadIdx.x;
    double term = 0.0, bad = 0.0;
    if (pat < a.nPat) {
        const int dim = a.dim, nCat = a.nCat;
        const size_t ps = (size_t)a.ps;
        uint64_t mask = ~0ull;    // states the root may be in
        if (a.rootTips) {         // the root is a leaf: only its observed state(s) (Pf/p4_tree.c:1199-1378)
            const int w = a.rootTips[pat];
            if (w < dim) mask = 1ull << w;
            else if (w > dim) mask = a.eqMask[w - dim - 1];
        }
        double A = 0.0;
        for (int c = 0; c < nCat; c++)
            for (int s = 0; s < dim; s++)
                if ((mask >> s) & 1ull) A = fma(a.pi[s], a.cl[((size_t)c * dim + s) * ps + pat], A);
        const int e = a.rootScale ? a.rootScale[pat] : 0;
        const uint64_t im = (a.pInvar != 0.0 && a.invarMask) ? a.invarMask[pat] : 0ull;
        double t1 = 0.0, like = 0.0;
        if (like_term(A, e, a.pInvar, nCat, im, a.pi, dim, a.counts[pat], &t1, &like)) term = t1;
        else bad = 1.0;
        if (a.patLikes) a.patLikes[pat] = like;
    }
    term = warpSum(term);
    bad = warpSum(bad);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sSum[w] = term; sBad[w] = bad; }
    __syncthreads();
    if (w == 0) {
        term = (l < (int)(blockDim.x >> 5)) ? sSum[l] : 0.0;
        bad = (l < (int)(blockDim.x >> 5)) ? sBad[l] : 0.0;
        term = warpSum(term);
        bad = warpSum(bad);
        if (l == 0) {
            a.partials[2 * blockIdx.x] = term;
            a.partials[2 * blockIdx.x + 1] = bad;
        }
    }
}

// One block folds the per-block partials in a fixed order (deterministic).
__global__ void __launch_bounds__(256)
like_final_kernel(const double *__restrict__ partials, int n, double *__restrict__ result, const MailArgs mail)
{
    __shared__ double sSum[8], sBad[8];
    double term = 0.0, bad = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        term += partials[2 * i];
        bad += partials[2 * i + 1];
    }
    term = warpSum(term);
    bad = warpSum(bad);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sSum[w] = term; sBad[w] = bad; }
    __syncthreads();
    if (w == 0) {
        term = (l < 8) ? sSum[l] : 0.0;
        bad = (l < 8) ? sBad[l] : 0.0;
        term = warpSum(term);
        bad = warpSum(bad);
        if (l == 0) {
            if (mail.world > 1) mail_allreduce(mail, 0, term, bad);
            result[0] = term;
            result[1] = bad;
        }
    }
}

// The same fold for several trees evaluated by one batched launch: block b folds tree b.
struct FinalBatchArgs {
    const double *partials[kMaxBatchTrees];
    double *result;     // [2*nTrees]
    int nBlocks;
    MailArgs mail;
};
__global__ void __launch_bounds__(256)
like_final_batch_kernel(const FinalBatchArgs a)
{
    __shared__ double sSum[8], sBad[8];
    const double *partials = a.partials[blockIdx.x];
    double term = 0.0, bad = 0.0;
    for (int i = threadIdx.x; i < a.nBlocks; i += blockDim.x) {
        term += partials[2 * i];
        bad += partials[2 * i + 1];
    }
    term = warpSum(term);
    bad = warpSum(bad);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sSum[w] = term; sBad[w] = bad; }
    __syncthreads();
    if (w == 0) {
        term = (l < 8) ? sSum[l] : 0.0;
        bad = (l < 8) ? sBad[l] : 0.0;
        term = warpSum(term);
        bad = warpSum(bad);
        if (l == 0) {
            if (a.mail.world > 1) mail_allreduce(a.mail, blockIdx.x, term, bad);
            a.result[2 * blockIdx.x] = term;
            a.result[2 * blockIdx.x + 1] = bad;
        }
    }
}

// max |a-b| > eps ?  (p4_verifyCondLikes / p4_verifyBigPDecks, epsilon 1e-15)
__global__ void __launch_bounds__(256)
diff_kernel(const double *__restrict__ x, const double *__restrict__ y, size_t n, double eps, int *__restrict__ flag)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    int bad = 0;
    for (; i < n; i += stride)
        if (fabs(x[i] - y[i]) > eps) bad = 1;
    if (bad) atomicOr(flag, 1);
}

__global__ void flush_kernel(double *buf, size_t n, double v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) buf[i] = v;
}

}  // namespace p4b
