/* oracle/pf_port.c -- TEST INFRASTRUCTURE ONLY.  Never linked into the product.
 *
 * A plain-C, single-threaded restatement of the reference's likelihood path, used
 * as a checker by tests/, __graft_entry__.smoke() and (where oracle/_ref is not
 * available) bench.py's cpu_baseline leg.  Each function names the reference code
 * whose algorithm it follows (paths relative to /root/reference).
 *
 * PINNING.  The port is checked against (a) the golden fixtures in tests/golden/,
 * which were produced by the reference's own Python package driving its own Pf
 * engine on the reference's example inputs (tests/golden/make_golden.py), and
 * (b) oracle/_ref -- the reference engine itself -- on seeded synthetic inputs
 * (tests/test_oracle.py).  So parity is pinned by executing the reference; the
 * reference records no known answers of its own.
 *
 * One deliberate difference: the eigensystem of Q.  The reference runs an
 * EISPACK-style general real solver (Pf/linalg.c:249) and an LU inverse; the
 * port symmetrises the reversible Q and runs cyclic Jacobi.  P(t) does not depend
 * on the eigenbasis, and the fixtures confirm agreement to ~1e-15.
 */
#include "pf_port.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define GAP_CODE (-1)     /* Pf/defines.h:33 */
#define QMARK_CODE (-2)   /* Pf/defines.h:34 */
#define N_LIKE (-3)       /* Pf/defines.h:35 */
#define EQUATES_BASE (-64) /* Pf/defines.h:36 */

/* ---- data: Pf/part.c ------------------------------------------------------ */

/* pokeSequences, Pf/part.c:127-275: symbols first, then '-', '?', then equates. */
int pfport_poke_sequences(const char *s, int nTax, int nChar, const char *symbols, int dim,
                          const char *equateSymbols, int nEquates, int *sequences)
{
    long k = 0;
    for (int i = 0; i < nTax; i++)
        for (int j = 0; j < nChar; j++, k++) {
            const char c = s[k];
            int coded = 0;
            for (int m = 0; m < dim && !coded; m++)
                if (c == symbols[m]) { sequences[k] = m; coded = 1; }
            if (!coded && c == '-') { sequences[k] = GAP_CODE; coded = 1; }
            else if (!coded && c == '?') { sequences[k] = QMARK_CODE; coded = 1; }
            else if (!coded)
                for (int m = 0; m < nEquates && !coded; m++)
                    if (c == equateSymbols[m]) { sequences[k] = EQUATES_BASE + m; coded = 1; }
            if (!coded) return 1;
        }
    return 0;
}

/* makePatterns, Pf/part.c:317-448: unique columns in first-occurrence order, found
 * by scanning every earlier pattern -- quadratic, exactly like the reference. */
int pfport_make_patterns(const int *sequences, int nTax, int nChar, int *patterns, int *patternCounts,
                         int *sequencePositionPatternIndex)
{
    int nPat = 0;
    for (int i = 0; i < nChar; i++) { patternCounts[i] = 0; sequencePositionPatternIndex[i] = 0; }
    for (int i = 0; i < nChar; i++) {
        int already = -1;
        for (int p = 0; p < nPat && already < 0; p++) {
            int same = 1;
            for (int j = 0; j < nTax; j++)
                if (sequences[(long)j * nChar + i] != patterns[(long)j * nChar + p]) { same = 0; break; }
            if (same) already = p;
        }
        if (already < 0) {
            for (int j = 0; j < nTax; j++) patterns[(long)j * nChar + nPat] = sequences[(long)j * nChar + i];
            already = nPat++;
        }
        patternCounts[already]++;
        sequencePositionPatternIndex[i] = already;
    }
    return nPat;
}

/* setGlobalInvarSitesVec, Pf/part.c:716-848: per pattern, the states every taxon is compatible with. */
void pfport_invar_sites(const int *patterns, int nTax, int nChar, int nPatterns, int dim, const int *equates,
                        int *vec, int *array)
{
    for (int p = 0; p < nPatterns; p++) {
        int sum = 0;
        for (int s = 0; s < dim; s++) {
            int ok = 1;
            for (int t = 0; t < nTax && ok; t++) {
                const int c = patterns[(long)t * nChar + p];
                if (c >= 0) ok = (c == s);
                else if (c == N_LIKE || c == GAP_CODE || c == QMARK_CODE) ok = 1;
                else ok = equates[(c - EQUATES_BASE) * dim + s] != 0;
            }
            array[(long)s * nChar + p] = ok;
            sum += ok;
        }
        vec[p] = sum;
    }
}

/* ---- discrete gamma: Pf/gamma.c ---------------------------------------------- */
static double ln_gamma(double alpha)   /* LnGamma, Pf/gamma.c:240-262 */
{
    double x = alpha, f = 0.0, z;
    if (x < 7) {
        f = 1.0;
        z = x - 1.0;
        while (++z < 7.0) f *= z;
        x = z;
        f = -log(f);
    }
    z = 1.0 / (x * x);
    return f + (x - 0.5) * log(x) - x + 0.918938533204673 +
           (((-0.000595238095238 * z + 0.000793650793651) * z - 0.002777777777778) * z + 0.083333333333333) / x;
}

static double point_normal(double prob)   /* PointNormal, Pf/gamma.c:282-299 */
{
    double a0 = -0.322232431088, a1 = -1.0, a2 = -0.342242088547, a3 = -0.0204231210245, a4 = -0.453642210148e-4,
           b0 = 0.0993484626060, b1 = 0.588581570495, b2 = 0.531103462366, b3 = 0.103537752850, b4 = 0.0038560700634, y, z,
           p = prob, p1;
    p1 = (p < 0.5 ? p : 1 - p);
    if (p1 < 1e-20) return -9999;
    y = sqrt(log(1 / (p1 * p1)));
    z = y + ((((y * a4 + a3) * y + a2) * y + a1) * y + a0) / ((((y * b4 + b3) * y + b2) * y + b1) * y + b0);
    return (p < 0.5 ? -z : z);
}

static double incomplete_gamma(double x, double alpha, double lnGammaAlpha)   /* IncompleteGamma, Pf/gamma.c:150-224 */
{
    double p = alpha, g = lnGammaAlpha, accurate = 1e-8, overflow = 1e30, factor, gin, rn, a, b, an, dif, term, pn[6];
    int i;
    if (x == 0.0) return 0.0;
    if (x < 0 || p <= 0) return -1.0;
    factor = exp(p * log(x) - x - g);
    if (!(x > 1 && x >= p)) {
        gin = 1.0; term = 1.0; rn = p;
        do { rn++; term *= x / rn; gin += term; } while (term > accurate);
        gin *= factor / p;   /* same association as the reference: gin * (factor / p) */
        return gin;
    }
    a = 1.0 - p; b = a + x + 1.0; term = 0.0;
    pn[0] = 1.0; pn[1] = x; pn[2] = x + 1; pn[3] = x * b;
    gin = pn[2] / pn[3];
    for (;;) {
        a++; b += 2.0; term++;
        an = a * term;
        for (i = 0; i < 2; i++) pn[i + 4] = b * pn[i + 2] - an * pn[i];
        if (pn[5] != 0) {
            rn = pn[4] / pn[5];
            dif = fabs(gin - rn);
            if (dif <= accurate && dif <= accurate * rn) break;
            gin = rn;
        }
        for (i = 0; i < 4; i++) pn[i] = pn[i + 2];
        if (fabs(pn[4]) >= overflow)
            for (i = 0; i < 4; i++) pn[i] /= overflow;
    }
    return 1.0 - factor * gin;
}

static double point_chi2(double prob, double v)   /* PointChi2, Pf/gamma.c:66-133 */
{
    double e = 0.5e-6, aa = 0.6931471805, p = prob, g, xx, c, ch, a, q, p1, p2, t, x, b, s1, s2, s3, s4, s5, s6;
    if (p < 0.000002 || p > 0.999998 || v <= 0.0) return -1.0;
    g = ln_gamma(v / 2.0);
    xx = v / 2.0;
    c = xx - 1.0;
    if (!(v >= -1.24 * log(p))) {
        ch = pow((p * xx * exp(g + xx * aa)), 1.0 / xx);
        if (ch - e < 0) return ch;
    } else if (v > 0.32) {
        x = point_normal(p);
        p1 = 0.222222 / v;
        ch = v * pow((x * sqrt(p1) + 1.0 - p1), 3.0);
        if (ch > 2.2 * v + 6.0) ch = -2.0 * (log(1.0 - p) - c * log(0.5 * ch) + g);
    } else {
        ch = 0.4;
        a = log(1.0 - p);
        do {
            q = ch;
            p1 = 1.0 + ch * (4.67 + ch);
            p2 = ch * (6.73 + ch * (6.66 + ch));
            t = -0.5 + (4.67 + 2.0 * ch) / p1 - (6.73 + ch * (13.32 + 3.0 * ch)) / p2;
            ch -= (1.0 - exp(a + g + 0.5 * ch + c * aa) * p2 / p1) / t;
        } while (!(fabs(q / ch - 1.0) - 0.01 <= 0.0));
    }
    do {
        q = ch;
        p1 = 0.5 * ch;
        if ((t = incomplete_gamma(p1, xx, g)) < 0.0) return -1.0;
        p2 = p - t;
        t = p2 * exp(xx * aa + g + p1 - c * log(ch));
        b = t / ch;
        a = 0.5 * t - b * c;
        s1 = (210.0 + a * (140.0 + a * (105.0 + a * (84.0 + a * (70.0 + 60.0 * a))))) / 420.0;
        s2 = (420.0 + a * (735.0 + a * (966.0 + a * (1141.0 + 1278.0 * a)))) / 2520.0;
        s3 = (210.0 + a * (462.0 + a * (707.0 + 932.0 * a))) / 2520.0;
        s4 = (252.0 + a * (672.0 + 1182.0 * a) + c * (294.0 + a * (889.0 + 1740.0 * a))) / 5040.0;
        s5 = (84.0 + 264.0 * a + c * (175.0 + 606.0 * a)) / 2520.0;
        s6 = (120.0 + c * (346.0 + 127.0 * c)) / 5040.0;
        ch += t * (1 + 0.5 * t * s1 - b * c * (s1 - b * (s2 - b * (s3 - b * (s4 - b * (s5 - b * s6))))));
    } while (fabs(q / ch - 1.0) > e);
    return ch;
}

/* DiscreteGamma with median = 0, alfa = beta, Pf/gamma.c:17-50. */
void pfport_discrete_gamma(double alpha, int K, double *freqK, double *rK)
{
    const double beta = alpha, factor = alpha / beta * K, lnga1 = ln_gamma(alpha + 1);
    int i;
    for (i = 0; i < K - 1; i++) freqK[i] = point_chi2((i + 1.0) / K, 2.0 * alpha) / (2.0 * beta);
    for (i = 0; i < K - 1; i++) freqK[i] = incomplete_gamma(freqK[i] * beta, alpha + 1, lnga1);
    rK[0] = freqK[0] * factor;
    rK[K - 1] = (1 - freqK[K - 2]) * factor;
    for (i = 1; i < K - 1; i++) rK[i] = (freqK[i] - freqK[i - 1]) * factor;
    for (i = 0; i < K; i++) freqK[i] = 1.0 / K;
}

/* ---- Q and P(t): Pf/util.c:44-115, Pf/eig.c:63-223 ------------------------------- */
void pfport_big_q(const double *R, const double *pi, int dim, double *Q)
{
    double sumODE = 0.0;
    for (int col = 0; col < dim; col++)
        for (int row = 0; row < dim; row++) Q[row * dim + col] = R[row * dim + col] * pi[col];
    for (int row = 0; row < dim; row++) {
        double sum = 0.0;
        for (int col = 0; col < dim; col++)
            if (row != col) sum = sum + Q[row * dim + col];
        Q[row * dim + row] = -sum;
    }
    for (int row = 0; row < dim; row++)
        for (int col = 0; col < dim; col++)
            if (row != col) sumODE = sumODE + (pi[row] * Q[row * dim + col]);
    sumODE = 1.0 / sumODE;
    for (int i = 0; i < dim * dim; i++) Q[i] = Q[i] * sumODE;
}

/* Eigensystem of a reversible Q through its symmetric similarity transform (see header note). */
static void eigen_reversible(const double *Q, const double *pi, int n, double *V, double *Vi, double *lam)
{
    double *S = malloc(sizeof(double) * n * n), *U = malloc(sizeof(double) * n * n), *sp = malloc(sizeof(double) * n);
    for (int i = 0; i < n; i++) sp[i] = sqrt(pi[i]);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) S[i * n + j] = 0.5 * (Q[i * n + j] * sp[i] / sp[j] + Q[j * n + i] * sp[j] / sp[i]);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) U[i * n + j] = (i == j);
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) off += S[p * n + q] * S[p * n + q];
        if (off < 1e-300) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = S[p * n + q];
                if (fabs(apq) < 1e-300) continue;
                const double theta = (S[q * n + q] - S[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) {
                    const double a = S[k * n + p], b = S[k * n + q];
                    S[k * n + p] = c * a - s * b;
                    S[k * n + q] = s * a + c * b;
                }
                for (int k = 0; k < n; k++) {
                    const double a = S[p * n + k], b = S[q * n + k];
                    S[p * n + k] = c * a - s * b;
                    S[q * n + k] = s * a + c * b;
                }
                for (int k = 0; k < n; k++) {
                    const double a = U[k * n + p], b = U[k * n + q];
                    U[k * n + p] = c * a - s * b;
                    U[k * n + q] = s * a + c * b;
                }
            }
    }
    /* Gram-Schmidt pass: rotations leave U orthogonal only to ~1e-14 */
    for (int k = 0; k < n; k++) {
        for (int pass = 0; pass < 2; pass++)
            for (int j = 0; j < k; j++) {
                double d = 0.0;
                for (int i = 0; i < n; i++) d += U[i * n + k] * U[i * n + j];
                for (int i = 0; i < n; i++) U[i * n + k] -= d * U[i * n + j];
            }
        double nrm = 0.0;
        for (int i = 0; i < n; i++) nrm += U[i * n + k] * U[i * n + k];
        nrm = sqrt(nrm);
        for (int i = 0; i < n; i++) U[i * n + k] /= nrm;
    }
    for (int k = 0; k < n; k++) {   /* Rayleigh quotients on the original symmetric matrix */
        double num = 0.0;
        for (int i = 0; i < n; i++) {
            double su = 0.0;
            for (int j = 0; j < n; j++) su += 0.5 * (Q[i * n + j] * sp[i] / sp[j] + Q[j * n + i] * sp[j] / sp[i]) * U[j * n + k];
            num += U[i * n + k] * su;
        }
        lam[k] = num;
    }
    for (int i = 0; i < n; i++)
        for (int k = 0; k < n; k++) {
            V[i * n + k] = U[i * n + k] / sp[i];
            Vi[k * n + i] = U[i * n + k] * sp[i];
        }
    free(S); free(U); free(sp);
}

/* matrixExpTimesBranchLength, Pf/eig.c:163-191 */
static void matrix_exp(const double *V, const double *Vi, const double *lam, int n, double t, double *P)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            double r = 0.0;
            for (int k = 0; k < n; k++) r = r + (V[i * n + k] * Vi[k * n + j] * exp(lam[k] * t));
            P[i * n + j] = r;
        }
}

/* The CL recursion and the log-likelihood are instantiated twice: in double, like the reference, and in
 * long double (x87 80-bit, exponent range 1e-4932), which does not underflow for thousands of taxa and
 * serves as the independent check of the engine's optional scalers. */
#define REAL double
#define FN pfport_part_loglike
#define LOGF log
#include "pf_port_cl.inc"
#undef REAL
#undef FN
#undef LOGF
#define REAL long double
#define FN pfport_part_loglike_ld
#define LOGF logl
#include "pf_port_cl.inc"
#undef REAL
#undef FN
#undef LOGF


/* ---- Newton-Raphson on one branch length: Pf/p4_treeNewt.c, Pf/p4_node.c:860-1018 -------------------------
 * out = { sum n log l,  sum n l'/l,  sum n (l'' l - l'^2)/l^2 }  in the length of `node`'s branch, the three sums
 * p4_newtNode forms per iteration (:508-517).  cl2 (everything on the far side of a branch) is built top-down along
 * the path from the root's child to `node` (p4_setNodeCL2 :603-624: p4_initializeCL2ToRootComp or p4_setCL2Up for
 * the parent's side, p4_setCL2Down for every sibling); the derivative decks follow
 * first/secondDerivativeOfMatrixExpTimesBranchLength (Pf/eig.c:294-371) with the rate factors of
 * p4_calculateBigPDecks_1stD / _2ndD (Pf/p4_node.c:442-540, including the second derivative's extra relRate in the
 * gamma, no-pInvar branch). */
static double leaf_term(const pfport_part *D, int dim, int code, const double *row /* deck row [dim] */, int plainIsOne)
{
    if (code >= 0) return row[code];
    if (code == GAP_CODE || code == QMARK_CODE) {
        if (plainIsOne) return 1.0;
        double s = 0.0;
        for (int x = 0; x < dim; x++) s += row[x];
        return s;
    }
    const int *eq = D->equates + (code - EQUATES_BASE) * dim;
    double s = 0.0;
    for (int x = 0; x < dim; x++)
        if (eq[x]) s += row[x];
    return s;
}

int pfport_branch_derivs(const pfport_tree *T, const pfport_part *D, const pfport_model *M, int node, double out[3])
{
    const int dim = M->dim, nCat = M->nCat, nPat = D->nPatterns, nN = T->nNodes;
    const long clSize = (long)nCat * dim * nPat, pSize = (long)nCat * dim * dim;
    if (node < 0 || node >= nN || node == T->root || T->parent[node] < 0) return 1;
    double *cl = malloc(sizeof(double) * clSize * nN), *P = malloc(sizeof(double) * pSize * nN);
    pfport_part_loglike(T, D, M, cl, P, NULL);
    /* the path root's child ... node */
    int *path = malloc(sizeof(int) * nN), nPath = 0;
    for (int q = node; T->parent[q] >= 0; q = T->parent[q]) path[nPath++] = q;
    double *cl2 = malloc(sizeof(double) * clSize * nPath);
    const double *piRoot = M->comps + (long)T->compNum[T->root] * dim;
    for (int k = nPath - 1; k >= 0; k--) {
        const int q = path[k], par = T->parent[q];
        double *mine = cl2 + (long)k * clSize;
        const double *up = (par == T->root) ? NULL : cl2 + (long)(k + 1) * clSize;
        for (int pat = 0; pat < nPat; pat++)
            for (int cat = 0; cat < nCat; cat++)
                for (int s = 0; s < dim; s++) {
                    double v;
                    if (!up) v = piRoot[s];                                          /* p4_initializeCL2ToRootComp */
                    else {                                                           /* p4_setCL2Up */
                        const double *Pp = P + (long)par * pSize + (long)cat * dim * dim;
                        v = 0.0;
                        for (int f = 0; f < dim; f++) v += Pp[f * dim + s] * up[((long)cat * dim + f) * nPat + pat];
                    }
                    for (int sib = T->leftChild[par]; sib >= 0; sib = T->sibling[sib]) {   /* p4_setCL2Down */
                        if (sib == q) continue;
                        const double *Pc = P + (long)sib * pSize + (long)cat * dim * dim + (long)s * dim;
                        double f;
                        if (T->isLeaf[sib]) {
                            const int code = D->patterns[(long)T->seqNum[sib] * D->stride + pat];
                            int isN = 0;
                            if (code <= EQUATES_BASE + D->nEquates - 1 && code >= EQUATES_BASE) {
                                const int *eq = D->equates + (code - EQUATES_BASE) * dim;
                                isN = 1;
                                for (int x = 0; x < dim; x++)
                                    if (!eq[x]) { isN = 0; break; }
                            }
                            f = isN ? 1.0 : leaf_term(D, dim, code, Pc, 1);
                        } else {
                            const double *cc = cl + (long)sib * clSize + (long)cat * dim * nPat + pat;
                            f = 0.0;
                            for (int x = 0; x < dim; x++) f += Pc[x] * cc[(long)x * nPat];
                        }
                        v *= f;
                    }
                    mine[((long)cat * dim + s) * nPat + pat] = v;
                }
    }
    /* decks of the branch: P (already in P), first and second derivative */
    double *V = malloc(sizeof(double) * dim * dim), *Vi = malloc(sizeof(double) * dim * dim), *lam = malloc(sizeof(double) * dim),
           *Q = malloc(sizeof(double) * dim * dim), *D1 = malloc(sizeof(double) * pSize), *D2 = malloc(sizeof(double) * pSize);
    pfport_big_q(M->bigR + (long)T->rMatrixNum[node] * dim * dim, M->comps + (long)T->compNum[node] * dim, dim, Q);
    eigen_reversible(Q, M->comps + (long)T->compNum[node] * dim, dim, V, Vi, lam);
    const double *rates = M->nGdasrvs ? M->rates + (long)T->gdasrvNum[node] * nCat : NULL;
    for (int cat = 0; cat < nCat; cat++) {
        double t, r1, r2;
        if (M->pInvar == 0.0) {
            if (rates) { const double temp = rates[cat] * M->relRate; t = T->brLen[node] * temp; r1 = temp; r2 = temp * M->relRate; }
            else { t = T->brLen[node] * M->relRate; r1 = r2 = M->relRate; }
        } else {
            if (rates) { const double temp = rates[cat] * M->relRate; t = (T->brLen[node] * temp) / (1.0 - M->pInvar); r1 = r2 = temp / (1.0 - M->pInvar); }
            else { t = (T->brLen[node] * M->relRate) / (1.0 - M->pInvar); r1 = r2 = M->relRate / (1.0 - M->pInvar); }
        }
        for (int i = 0; i < dim; i++)
            for (int j = 0; j < dim; j++) {
                double a = 0.0, b = 0.0;
                for (int k = 0; k < dim; k++) {
                    a = a + (V[i * dim + k] * Vi[k * dim + j] * lam[k] * r1 * exp(lam[k] * t));
                    b = b + (V[i * dim + k] * Vi[k * dim + j] * lam[k] * lam[k] * r2 * r2 * exp(lam[k] * t));
                }
                D1[(long)cat * dim * dim + i * dim + j] = a;
                D2[(long)cat * dim * dim + i * dim + j] = b;
            }
    }
    /* the sums of p4_newtNode, :258-517 */
    const double *z = cl2, *x = cl + (long)node * clSize, *P0 = P + (long)node * pSize;
    double lnL = 0.0, firstD = 0.0, secondD = 0.0;
    for (int pat = 0; pat < nPat; pat++) {
        double likeS = 0.0, firstS = 0.0, secondS = 0.0;
        for (int cat = 0; cat < nCat; cat++) {
            double like = 0.0, first = 0.0, second = 0.0;
            for (int f = 0; f < dim; f++) {
                const double zz = z[((long)cat * dim + f) * nPat + pat];
                const double *r0 = P0 + (long)cat * dim * dim + (long)f * dim, *r1 = D1 + (long)cat * dim * dim + (long)f * dim,
                             *r2 = D2 + (long)cat * dim * dim + (long)f * dim;
                if (T->isLeaf[node]) {
                    const int code = D->patterns[(long)T->seqNum[node] * D->stride + pat];
                    like += zz * leaf_term(D, dim, code, r0, code == GAP_CODE || code == QMARK_CODE);
                    first += zz * leaf_term(D, dim, code, r1, 0);
                    second += zz * leaf_term(D, dim, code, r2, 0);
                } else {
                    for (int t = 0; t < dim; t++) {
                        const double temp2 = zz * x[((long)cat * dim + t) * nPat + pat];
                        like += temp2 * r0[t];
                        first += temp2 * r1[t];
                        second += temp2 * r2[t];
                    }
                }
            }
            likeS += like;
            firstS += first;
            secondS += second;
        }
        if (M->pInvar != 0.0) {
            const double f0 = (1.0 - M->pInvar) / (double)nCat;
            likeS *= f0;
            firstS *= f0;
            secondS *= f0;
            if (D->invarVec[pat] > 0)
                for (int s = 0; s < dim; s++)
                    if (D->invarArray[(long)s * D->stride + pat]) likeS += piRoot[s] * M->pInvar;
        } else if (nCat > 1) {
            likeS /= (double)nCat;
            firstS /= (double)nCat;
            secondS /= (double)nCat;
        }
        const double cnt = D->patternCounts[pat];
        if (likeS < 1.0e-300) { lnL += cnt * -100000; firstD += cnt * 1000000; secondD += cnt * 10000000; }
        else {
            lnL += cnt * log(likeS);
            firstD += cnt * (firstS / likeS);
            secondD += cnt * ((secondS * likeS - firstS * firstS) / (likeS * likeS));
        }
    }
    out[0] = lnL;
    out[1] = firstD;
    out[2] = secondD;
    free(cl); free(P); free(path); free(cl2); free(V); free(Vi); free(lam); free(Q); free(D1); free(D2);
    return 0;
}
