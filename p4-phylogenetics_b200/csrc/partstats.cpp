// partstats.cpp -- the data part's own statistics and views (Pf/part.c), the wrappers p4's callers use around the
// likelihood path: Part.composition() for empirical compositions (p4/part.py:62-70 -> partComposition),
// Data.resetSequencesFromParts after a simulation (symbolSequences, p4/alignment.py:5947-5957), and the
// composition / model-fit tests that follow Tree.simulate (partSequenceSitesCount, singleSequenceBaseCounts,
// partBigXSquared, partSimpleConstantSitesCount, partMeanNCharsPerSite; p4/tree.py:8119-9052, p4/data.py:339-766).
// Host code on the part's integer arrays; every result is checked for equality with the reference's on the CPU
// (tests/test_host.py).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../include/p4b200.h"
#include "engine.h"

using namespace p4b;

extern "C" {

/* pf.singleSequenceBaseCounts(part, seqNum) -> list, Pf/part.c:556-602: from the patterns and their counts when the
 * part has patterns, else from the sequences. */
int p4b_singleSequenceBaseCounts(p4b_part part, int seqNum, int *outDim)
{
    Part *p = (Part *)part;
    if (!p || !outDim) { setError("singleSequenceBaseCounts: NULL argument"); return 1; }
    if (seqNum < 0 || seqNum >= p->nTax) { setError("singleSequenceBaseCounts: bad sequence number %d", seqNum); return 1; }
    for (int i = 0; i < p->dim; i++) outDim[i] = 0;
    if (p->nPatterns) {
        const int *row = &p->patterns[(size_t)seqNum * p->nChar];
        for (int i = 0; i < p->nPatterns; i++)
            if (row[i] >= 0) outDim[row[i]] = outDim[row[i]] + p->patternCounts[i];
    } else if (p->nChar) {
        const int *row = &p->sequences[(size_t)seqNum * p->nChar];
        for (int i = 0; i < p->nChar; i++)
            if (row[i] >= 0) outDim[row[i]] = outDim[row[i]] + 1;
    } else {
        setError("part: singleSequenceBaseCounts: no sequences.");
        return 1;
    }
    return 0;
}

/* pf.symbolSequences(part) -> one string of nTax*nChar characters, Pf/part.c:604-680. */
int p4b_symbolSequences(p4b_part part, char *outNTaxTimesNCharPlus1)
{
    Part *p = (Part *)part;
    if (!p || !outNTaxTimesNCharPlus1) { setError("symbolSequences: NULL argument"); return 1; }
    if (!p->nChar) { setError("part.c: symbolSequences: no sequences."); return 1; }
    const size_t cells = (size_t)p->nTax * p->nChar;
    for (size_t k = 0; k < cells; k++) {
        const int c = p->sequences[k];
        char ch;
        if (c >= 0 && c < p->dim) ch = p->symbols[c];
        else if (c == P4B_GAP_CODE) ch = '-';
        else if (c == P4B_QMARK_CODE) ch = '?';
        else if (c >= P4B_EQUATES_BASE && c < P4B_EQUATES_BASE + p->nEquates) ch = p->equateSymbols[c - P4B_EQUATES_BASE];
        else { setError("part.c: symbolSequences: character number %d is not recognized.", c); return 1; }
        outNTaxTimesNCharPlus1[k] = ch;
    }
    outNTaxTimesNCharPlus1[cells] = '\0';
    return 0;
}

/* pf.partSequenceSitesCount(part, seqNum), Pf/part.c:1068-1082: sites that are neither gap nor '?'. */
int p4b_partSequenceSitesCount(p4b_part part, int seqNum)
{
    Part *p = (Part *)part;
    if (!p || seqNum < 0 || seqNum >= p->nTax) { setError("partSequenceSitesCount: bad argument"); return -1; }
    int n = 0;
    const int *row = &p->sequences[(size_t)seqNum * p->nChar];
    for (int i = 0; i < p->nChar; i++)
        if (row[i] == P4B_GAP_CODE || row[i] == P4B_QMARK_CODE) n++;
    return p->nChar - n;
}

/* pf.pokePartTaxListAtIndex(part, val, index), Pf/pfmodule.c:326-346: which sequences partComposition looks at. */
int p4b_pokePartTaxListAtIndex(p4b_part part, int val, int index)
{
    Part *p = (Part *)part;
    if (!p) { setError("pokePartTaxListAtIndex: NULL part"); return 1; }
    if (index < 0 || index >= p->nTax) { setError("pokePartTaxListAtIndex: index %d out of range", index); return 1; }
    if ((int)p->taxList.size() != p->nTax) p->taxList.assign(p->nTax, 0);
    p->taxList[index] = val;
    return 0;
}

/* pf.partComposition(part) -> list, Pf/part.c:850-1066: composition over the sequences selected by the tax list; an
 * ambiguity's count is shared among its states in proportion to the composition (iterated to 1e-12, at most 1000
 * times, per sequence), sequences weighted by their number of sites. */
int p4b_partComposition(p4b_part part, double *outDim)
{
    Part *p = (Part *)part;
    if (!p || !outDim) { setError("partComposition: NULL argument"); return 1; }
    const int dim = p->dim, nEq = p->nEquates;
    if ((int)p->taxList.size() != p->nTax) p->taxList.assign(p->nTax, 0);
    // the states each ambiguity code stands for, ascending (the order the reference's loops visit them in)
    std::vector<std::vector<int>> statesOf(nEq);
    for (int e = 0; e < nEq; e++)
        for (int k = 0; k < dim; k++)
            if (p->equates[(size_t)e * dim + k]) statesOf[e].push_back(k);
    std::vector<double> plain(dim), ambig(nEq), freq(dim), share(dim), weighted(dim, 0.0);
    long totalSites = 0;
    for (int seq = 0; seq < p->nTax; seq++) {
        if (!p->taxList[seq]) continue;
        // counts of this sequence: plain states, ambiguity codes, and the sites that carry neither
        std::fill(plain.begin(), plain.end(), 0.0);
        std::fill(ambig.begin(), ambig.end(), 0.0);
        int blanks = 0;
        const int *row = &p->sequences[(size_t)seq * p->nChar];
        for (int j = 0; j < p->nChar; j++) {
            const int c = row[j];
            if (c >= 0) plain[c] = plain[c] + 1.0;
            else if (c == P4B_QMARK_CODE || c == P4B_GAP_CODE) blanks += 1;
            else ambig[c - P4B_EQUATES_BASE] += 1.0;
        }
        const int nSites = p->nChar - blanks;
        totalSites += nSites;
        if (!nSites) continue;
        // fixed point of: share every ambiguity's count among its states in proportion to the current composition
        // (start uniform -- the reference divides by a float dim --, stop when the composition moves by < 1e-12, 1000 rounds at most)
        const double start = 1.0 / ((float)dim);
        for (int k = 0; k < dim; k++) freq[k] = start;
        for (int round = 0; round < 1000; round++) {
            for (int k = 0; k < dim; k++) share[k] = plain[k];
            for (int e = 0; e < nEq; e++) {
                if (!(ambig[e] > 0.0)) continue;
                double mass = 0.0;
                for (int k : statesOf[e]) mass = mass + freq[k];
                for (int k : statesOf[e]) share[k] = share[k] + (ambig[e] * (freq[k] / mass));
            }
            double total = 0.0;
            for (int k = 0; k < dim; k++) total = total + share[k];
            double moved = 0.0;
            for (int k = 0; k < dim; k++) {
                const double before = freq[k];
                freq[k] = share[k] / total;
                moved = moved + fabs(freq[k] - before);
            }
            if (moved < 1.0e-12) break;
        }
        for (int k = 0; k < dim; k++) weighted[k] = weighted[k] + (freq[k] * (double)nSites);   // sequences weigh by their sites
    }
    for (int k = 0; k < dim; k++) outDim[k] = totalSites ? weighted[k] / (double)totalSites : 0.0;
    return 0;
}

/* pf.partMeanNCharsPerSite(part), Pf/part.c:1419-1457: mean number of different states per site. */
double p4b_partMeanNCharsPerSite(p4b_part part)
{
    Part *p = (Part *)part;
    if (!p) { setError("partMeanNCharsPerSite: NULL part"); return NAN; }
    std::vector<int> counts(p->dim, 0);
    int nSites = 0, theTotal = 0;
    for (int c = 0; c < p->nChar; c++) {
        for (int t = 0; t < p->nTax; t++) {
            const int s = p->sequences[(size_t)t * p->nChar + c];
            if (s >= 0 && s < p->dim) counts[s] += 1;
        }
        int here = 0;
        for (int s = 0; s < p->dim; s++)
            if (counts[s]) { here += 1; counts[s] = 0; }
        theTotal += here;
        nSites += 1;
    }
    return (double)theTotal / (double)nSites;
}

/* pf.partSimpleConstantSitesCount(part), Pf/part.c:1459-1488: sites where all unambiguous states are the same. */
int p4b_partSimpleConstantSitesCount(p4b_part part)
{
    Part *p = (Part *)part;
    if (!p) { setError("partSimpleConstantSitesCount: NULL part"); return -1; }
    int counts = 0;
    for (int c = 0; c < p->nChar; c++) {
        int top = -1, t = 0;
        bool isConstant = true;
        for (t = 0; t < p->nTax; t++) {
            const int s = p->sequences[(size_t)t * p->nChar + c];
            if (s >= 0 && s < p->dim) { top = s; break; }
        }
        for (int t2 = t; t2 < p->nTax; t2++) {
            const int s = p->sequences[(size_t)t2 * p->nChar + c];
            if (s >= 0 && s < p->dim && s != top) { isConstant = false; break; }
        }
        if (top != -1 && isConstant) counts += 1;
    }
    return counts;
}

/* pf.partBigXSquared(part), Pf/part.c:1490-1571: the X^2 statistic of compositional homogeneity over the sequences;
 * -2.0 (with the reference's message) when the data hold gaps or ambiguities. */
double p4b_partBigXSquared(p4b_part part)
{
    Part *p = (Part *)part;
    if (!p) { setError("partBigXSquared: NULL part"); return NAN; }
    const int dim = p->dim, nTax = p->nTax;
    std::vector<double> obs((size_t)nTax * dim, 0.0), exp(dim, 0.0), soc(dim, 0.0);
    const double oneOverNTax = 1.0 / nTax;
    for (int c = 0; c < p->nChar; c++)
        for (int t = 0; t < nTax; t++) {
            const int s = p->sequences[(size_t)t * p->nChar + c];
            if (s >= 0 && s < dim) obs[(size_t)t * dim + s] += 1;
            else {
                printf("This function cannot handle gaps and ambiguities.  Returning -2.0\n");
                return -2.0;
            }
        }
    for (int s = 0; s < dim; s++) {
        soc[s] = 0.0;
        for (int t = 0; t < nTax; t++) soc[s] += obs[(size_t)t * dim + s];
    }
    for (int s = 0; s < dim; s++) exp[s] = oneOverNTax * soc[s];
    double xSq = 0.0;
    for (int t = 0; t < nTax; t++)
        for (int s = 0; s < dim; s++)
            if (exp[s]) xSq += ((obs[(size_t)t * dim + s] - exp[s]) * (obs[(size_t)t * dim + s] - exp[s])) / exp[s];
    return xSq;
}

}  // extern "C"
