/* oracle/pf_port.h -- TEST INFRASTRUCTURE ONLY (see pf_port.c). */
#ifndef PF_PORT_H
#define PF_PORT_H

typedef struct {
    int nNodes, root, nPost;
    const int *parent, *leftChild, *sibling, *isLeaf, *seqNum;   /* [nNodes], -1 = none */
    const double *brLen;                                         /* [nNodes] */
    const int *postOrder;                                        /* [nPost], negative entries skipped */
    const int *compNum, *rMatrixNum, *gdasrvNum;                 /* [nNodes], for the part being evaluated */
} pfport_tree;

typedef struct {
    int nTax, nPatterns, stride;      /* patterns is [nTax][stride], columns < nPatterns valid */
    const int *patterns, *patternCounts;
    int nEquates;
    const int *equates;               /* [nEquates][dim] */
    const int *invarVec;              /* [stride] */
    const int *invarArray;            /* [dim][stride] */
} pfport_part;

typedef struct {
    int dim, nCat, nComps, nRMatrices, nGdasrvs;
    const double *comps;              /* [nComps][dim] */
    const double *bigR;               /* [nRMatrices][dim][dim] */
    const double *rates;              /* [nGdasrvs][nCat] */
    double pInvar, relRate;
} pfport_model;

int pfport_poke_sequences(const char *s, int nTax, int nChar, const char *symbols, int dim,
                          const char *equateSymbols, int nEquates, int *sequences);
int pfport_make_patterns(const int *sequences, int nTax, int nChar, int *patterns, int *patternCounts,
                         int *sequencePositionPatternIndex);
void pfport_invar_sites(const int *patterns, int nTax, int nChar, int nPatterns, int dim, const int *equates,
                        int *vec, int *array);
void pfport_discrete_gamma(double alpha, int K, double *freqK, double *rK);
void pfport_big_q(const double *R, const double *pi, int dim, double *Q);
/* Log-likelihood of one part.  clOut ([nNodes][nCat][dim][nPatterns]), pOut ([nNodes][nCat][dim][dim]) and
 * patLikes ([nPatterns]) are optional outputs. */
double pfport_part_loglike(const pfport_tree *T, const pfport_part *D, const pfport_model *M, double *clOut, double *pOut,
                           double *patLikes);
/* Same, with the CL recursion carried in long double (no underflow for thousands of taxa). */
double pfport_part_loglike_ld(const pfport_tree *T, const pfport_part *D, const pfport_model *M, double *clOut, double *pOut,
                              double *patLikes);
/* The three sums of one Newton-Raphson iteration on the branch of `node` (Pf/p4_treeNewt.c:238-520): lnL, d lnL/dv, d2 lnL/dv2. */
int pfport_branch_derivs(const pfport_tree *T, const pfport_part *D, const pfport_model *M, int node, double out[3]);
#endif
