// model.cpp -- substitution-model side of the engine (host, tiny, feeds the
// P(t) kernel): discrete-gamma rates, rate matrix Q and its eigensystem.
//
// Reference behaviour being matched:
//   DiscreteGamma and helpers        Pf/gamma.c:17-50, 66-133, 150-224, 240-262, 282-299
//   setBigQFromRMatrixDotCharFreq    Pf/util.c:44-74
//   normalizeBigQ                    Pf/util.c:76-115
//   p4_resetBQET / eigensystem       Pf/p4_model.c:513-543, Pf/eig.c:63-161
//   p4_newRMatrix spec dispatch      Pf/p4_model.c:346-435
//
// Eigensystem: the reference runs an EISPACK-style real *general* solver
// (Pf/linalg.c:249) plus an LU inverse.  Every exchangeability matrix that can
// reach the engine through the pf API is symmetric (pf.p4_setRMatrixBigR pokes
// [i][j] and [j][i] together, Pf/pfmodule.c:1735-1736; the empirical tables
// and the 2-parameter matrix are symmetric), so Q = R.diag(pi) is reversible:
// S = diag(sqrt pi) Q diag(1/sqrt pi) is symmetric, S = U L U^T, and
//   V = diag(1/sqrt pi) U,   V^-1 = U^T diag(sqrt pi).
// A cyclic Jacobi solver on S gives eigenvectors to machine precision and
// needs no explicit inverse.  P(t) = V exp(Lt) V^-1 does not depend on which
// eigenbasis is used, so results agree with the reference to rounding.  The
// reference's complex-eigenvalue branch (Pf/eig.c:109-136) is unreachable for
// reversible Q and has no counterpart here; a non-symmetric R is rejected.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "engine.h"

namespace p4b {

#include "protein_rmatrices.inc"

// ---------------------------------------------------------------------------
// Discrete gamma (Yang 1994).  The arithmetic is written in the same order as
// the published routines the reference uses, because the inner solver stops at
// a 0.5e-6 tolerance (Pf/gamma.c:70): only the same sequence of roundings
// reproduces the reference's rates to the last digits.
// ---------------------------------------------------------------------------
static double lnGammaPH(double alpha)   // Pike & Hill 1966, Algorithm 291
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

static double pointNormalOE(double prob)   // Odeh & Evans 1974, AS70
{
    const double a0 = -0.322232431088, a1 = -1.0, a2 = -0.342242088547, a3 = -0.0204231210245,
                 a4 = -0.453642210148e-4, b0 = 0.0993484626060, b1 = 0.588581570495, b2 = 0.531103462366,
                 b3 = 0.103537752850, b4 = 0.0038560700634;
    const double p = prob;
    const double p1 = (p < 0.5 ? p : 1 - p);
    if (p1 < 1e-20) return -9999;
    const double y = sqrt(log(1 / (p1 * p1)));
    const double z = y + ((((y * a4 + a3) * y + a2) * y + a1) * y + a0) / ((((y * b4 + b3) * y + b2) * y + b1) * y + b0);
    return (p < 0.5 ? -z : z);
}

static double incompleteGammaB(double x, double alpha, double lnGammaAlpha)   // Bhattacharjee 1970, AS32
{
    const double p = alpha, g = lnGammaAlpha, accurate = 1e-8, overflow = 1e30;
    if (x == 0.0) return 0.0;
    if (x < 0 || p <= 0) return -1.0;
    const double factor = exp(p * log(x) - x - g);
    if (!(x > 1 && x >= p)) {
        // series expansion
        double gin = 1.0, term = 1.0, rn = p;
        do {
            rn++;
            term *= x / rn;
            gin += term;
        } while (term > accurate);
        gin *= factor / p;
        return gin;
    }
    // continued fraction
    double pn[6];
    double a = 1.0 - p, b = a + x + 1.0, term = 0.0, gin, rn, an, dif;
    pn[0] = 1.0;
    pn[1] = x;
    pn[2] = x + 1;
    pn[3] = x * b;
    gin = pn[2] / pn[3];
    for (;;) {
        a++;
        b += 2.0;
        term++;
        an = a * term;
        for (int i = 0; i < 2; i++) pn[i + 4] = b * pn[i + 2] - an * pn[i];
        if (pn[5] != 0) {
            rn = pn[4] / pn[5];
            dif = fabs(gin - rn);
            if (dif <= accurate && dif <= accurate * rn) break;
            gin = rn;
        }
        for (int i = 0; i < 4; i++) pn[i] = pn[i + 2];
        if (fabs(pn[4]) >= overflow)
            for (int i = 0; i < 4; i++) pn[i] /= overflow;
    }
    return 1.0 - factor * gin;
}

static double pointChi2BR(double prob, double v)   // Best & Roberts 1975, AS91
{
    const double e = 0.5e-6, aa = 0.6931471805, p = prob;
    double ch, a, q, p1, p2, t, x, b, s1, s2, s3, s4, s5, s6;
    if (p < 0.000002 || p > 0.999998 || v <= 0.0) return -1.0;
    const double g = lnGammaPH(v / 2.0);
    const double xx = v / 2.0;
    const double c = xx - 1.0;
    if (!(v >= -1.24 * log(p))) {
        ch = pow((p * xx * exp(g + xx * aa)), 1.0 / xx);
        if (ch - e < 0) return ch;
    } else if (v > 0.32) {
        x = pointNormalOE(p);
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
        if ((t = incompleteGammaB(p1, xx, g)) < 0.0) return -1.0;
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

static inline double pointGamma(double prob, double alpha, double beta)
{
    return pointChi2BR(prob, 2.0 * (alpha)) / (2.0 * (beta));
}

int discreteGamma(double *freqK, double *rK, double alfa, double beta, int K, int median)
{
    const double gap05 = 1.0 / (2.0 * K), factor = alfa / beta * K;
    if (median) {
        double t = 0;
        for (int i = 0; i < K; i++) rK[i] = pointGamma((i * 2.0 + 1) * gap05, alfa, beta);
        for (int i = 0; i < K; i++) t += rK[i];
        for (int i = 0; i < K; i++) rK[i] *= factor / t;
    } else {
        const double lnga1 = lnGammaPH(alfa + 1);
        for (int i = 0; i < K - 1; i++) freqK[i] = pointGamma((i + 1.0) / K, alfa, beta);
        for (int i = 0; i < K - 1; i++) freqK[i] = incompleteGammaB(freqK[i] * beta, alfa + 1, lnga1);
        rK[0] = freqK[0] * factor;
        rK[K - 1] = (1 - freqK[K - 2]) * factor;
        for (int i = 1; i < K - 1; i++) rK[i] = (freqK[i] - freqK[i - 1]) * factor;
    }
    for (int i = 0; i < K; i++) freqK[i] = 1.0 / K;
    return 0;
}

// ---------------------------------------------------------------------------
// Rate matrices
// ---------------------------------------------------------------------------
int proteinBigR(int spec, double *out400)
{
    const int n = (int)(sizeof(kProteinSpecs) / sizeof(kProteinSpecs[0]));
    for (int i = 0; i < n; i++)
        if (kProteinSpecs[i] == spec) {
            memcpy(out400, kProteinBigR[i], 400 * sizeof(double));
            return 0;
        }
    return 1;
}

void setKappaBigR(RMatrix &r)   // Pf/pfmodule.c:1755-1774, Pf/p4_tree.c:297-322
{
    const double alpha = 1.0 / 3.0;
    const double beta = alpha * r.kappa;
    const double m[16] = {0.0, alpha, beta, alpha, alpha, 0.0, alpha, beta, beta, alpha, 0.0, alpha, alpha, beta, alpha, 0.0};
    for (int i = 0; i < 16; i++) r.bigR[i] = m[i];
}

// Q = R.diag(pi), diagonal = -rowsum, scaled so that the expected rate is 1.
static int buildNormalisedQ(double *Q, const double *R, const double *pi, int dim)
{
    for (int col = 0; col < dim; col++)
        for (int row = 0; row < dim; row++) Q[row * dim + col] = R[row * dim + col] * pi[col];
    for (int row = 0; row < dim; row++) {
        double sum = 0.0;
        for (int col = 0; col < dim; col++)
            if (row != col) sum = sum + Q[row * dim + col];
        Q[row * dim + row] = -sum;
    }
    double sumPi = 0.0;
    for (int row = 0; row < dim; row++) sumPi = sumPi + pi[row];
    if (sumPi < 0.999 || sumPi > 1.001) {
        setError("Model: normalizeBigQ: Something wrong with the charFreq.  sumOfCharFreqElements is %f", sumPi);
        return 1;
    }
    double sumODE = 0.0;
    for (int row = 0; row < dim; row++)
        for (int col = 0; col < dim; col++)
            if (row != col) sumODE = sumODE + (pi[row] * Q[row * dim + col]);
    sumODE = 1.0 / sumODE;
    for (int i = 0; i < dim * dim; i++) Q[i] = Q[i] * sumODE;
    return 0;
}

// Cyclic Jacobi eigen-decomposition of a symmetric matrix.  A is destroyed
// (its diagonal ends up holding the eigenvalues); U receives eigenvectors as
// columns.
static int jacobiSym(double *A, double *U, double *lam, int n)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) U[i * n + j] = (i == j) ? 1.0 : 0.0;
    double frob = 0.0;
    for (int i = 0; i < n * n; i++) frob += A[i] * A[i];
    if (frob == 0.0) {
        for (int i = 0; i < n; i++) lam[i] = 0.0;
        return 0;
    }
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0.0;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) off += A[p * n + q] * A[p * n + q];
        if (off <= 1e-36 * frob) break;
        for (int p = 0; p < n - 1; p++) {
            for (int q = p + 1; q < n; q++) {
                const double apq = A[p * n + q];
                if (apq == 0.0) continue;
                const double app = A[p * n + p], aqq = A[q * n + q];
                if (fabs(apq) < 1e-300) { A[p * n + q] = A[q * n + p] = 0.0; continue; }
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) {   // columns p,q
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {   // rows p,q
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                A[p * n + q] = A[q * n + p] = 0.0;
                for (int k = 0; k < n; k++) {
                    const double ukp = U[k * n + p], ukq = U[k * n + q];
                    U[k * n + p] = c * ukp - s * ukq;
                    U[k * n + q] = s * ukp + c * ukq;
                }
            }
        }
    }
    for (int i = 0; i < n; i++) lam[i] = A[i * n + i];
    return 0;
}

// Hundreds of plane rotations leave U orthogonal only to ~1e-14.  One
// Newton-Schulz step U <- U (3I - U^T U) / 2 restores orthogonality to rounding
// level (the iteration converges quadratically and U starts 1e-14 away), and
// Rayleigh quotients u_k^T S u_k then give eigenvalues accurate to second order
// in the remaining eigenvector error.  V V^-1 = I then holds to ~1e-15, which
// is what bounds the absolute error of P(t).
static void polishEigenvectors(const double *S, double *U, double *lam, int n)
{
    std::vector<double> G((size_t)n * n), T((size_t)n * n);
    for (int pass = 0; pass < 2; pass++) {
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                double g = 0.0;
                for (int k = 0; k < n; k++) g += U[k * n + i] * U[k * n + j];
                G[i * n + j] = (i == j ? 3.0 : 0.0) - g;
            }
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                double t = 0.0;
                for (int k = 0; k < n; k++) t += U[i * n + k] * G[k * n + j];
                T[i * n + j] = 0.5 * t;
            }
        for (int i = 0; i < n * n; i++) U[i] = T[i];
    }
    for (int k = 0; k < n; k++) {
        double num = 0.0;
        for (int i = 0; i < n; i++) {
            double su = 0.0;
            for (int j = 0; j < n; j++) su += S[i * n + j] * U[j * n + k];
            num += U[i * n + k] * su;
        }
        lam[k] = num;
    }
}

int resetBQET(Model *m, int pNum, int cNum, int rNum)
{
    if (!m || pNum < 0 || pNum >= m->nParts || !m->parts[pNum]) { setError("p4_resetBQET: bad part %d", pNum); return 1; }
    ModelPart *mp = m->parts[pNum];
    if (cNum < 0 || cNum >= mp->nComps || rNum < 0 || rNum >= mp->nRMatrices || !mp->compSet[cNum] || !mp->rMatrixSet[rNum]) {
        setError("p4_resetBQET: part %d has no comp %d / rMatrix %d", pNum, cNum, rNum);
        return 1;
    }
    const int dim = mp->dim;
    Eig &e = mp->bqe[(size_t)cNum * mp->nRMatrices + rNum];
    if (!e.allocated) {
        e.Q.assign((size_t)dim * dim, 0.0);
        e.V.assign((size_t)dim * dim, 0.0);
        e.Vinv.assign((size_t)dim * dim, 0.0);
        e.lam.assign(dim, 0.0);
        e.allocated = true;
    }
    const double *pi = mp->comps[cNum].val;
    const double *R = mp->rMatrices[rNum].bigR.data();
    // The reference recomputes every pair a non-root node uses on each p4_setPrams (Pf/p4_tree.c:455-501) --
    // 118 eigensystems of 20x20 per part for a composition-per-node model.  The eigensystem is a pure function
    // of (pi, R): when neither changed since it was last solved, the cached one IS the answer.
    if (e.inPi.size() == (size_t)dim && memcmp(e.inPi.data(), pi, sizeof(double) * dim) == 0 &&
        memcmp(e.inR.data(), R, sizeof(double) * dim * dim) == 0) {
        if (mp->bQETneedsReset) mp->bQETneedsReset[cNum * mp->nRMatrices + rNum] = 0;
        return 0;
    }
    e.inPi.clear();   // invalid until the solve below succeeds
    if (buildNormalisedQ(e.Q.data(), R, pi, dim)) return 1;

    std::vector<double> S((size_t)dim * dim), U((size_t)dim * dim), sp(dim);
    for (int i = 0; i < dim; i++) {
        if (!(pi[i] > 0.0)) { setError("p4_resetBQET: part %d comp %d value %d is %g", pNum, cNum, i, pi[i]); return 1; }
        sp[i] = sqrt(pi[i]);
    }
    double scale = 0.0, asym = 0.0;
    for (int i = 0; i < dim; i++)
        for (int j = 0; j < dim; j++) {
            S[i * dim + j] = e.Q[i * dim + j] * sp[i] / sp[j];
            if (fabs(S[i * dim + j]) > scale) scale = fabs(S[i * dim + j]);
        }
    for (int i = 0; i < dim; i++)
        for (int j = i + 1; j < dim; j++) {
            const double d = fabs(S[i * dim + j] - S[j * dim + i]);
            if (d > asym) asym = d;
            const double avg = 0.5 * (S[i * dim + j] + S[j * dim + i]);
            S[i * dim + j] = S[j * dim + i] = avg;
        }
    if (asym > 1e-9 * (scale > 0 ? scale : 1.0)) {
        setError("p4_resetBQET: part %d comp %d rMatrix %d: the exchangeability matrix is not symmetric "
                 "(non-reversible Q is not supported by this engine)", pNum, cNum, rNum);
        return 1;
    }
    const std::vector<double> S0(S);   // jacobiSym destroys its input
    if (jacobiSym(S.data(), U.data(), e.lam.data(), dim)) { setError("There is a problem with the eigensystem."); return 1; }
    polishEigenvectors(S0.data(), U.data(), e.lam.data(), dim);
    for (int i = 0; i < dim; i++)
        for (int k = 0; k < dim; k++) {
            e.V[i * dim + k] = U[i * dim + k] / sp[i];
            e.Vinv[k * dim + i] = U[i * dim + k] * sp[i];
        }
    e.version++;
    static uint64_t solveCounter = 0;
    e.content = ++solveCounter;
    e.inPi.assign(pi, pi + dim);
    e.inR.assign(R, R + (size_t)dim * dim);
    if (mp->bQETneedsReset) mp->bQETneedsReset[cNum * mp->nRMatrices + rNum] = 0;
    return 0;
}

}  // namespace p4b
