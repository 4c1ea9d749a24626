// data.cpp -- data parts: character coding, site-pattern compression and
// constant-site masks (host side, setup time).
//
// Behaviour follows Pf/part.c of the reference and must be bit-exact with it:
//   pokeEquatesTable   Pf/part.c:279-315
//   pokeSequences      Pf/part.c:127-275
//   makePatterns       Pf/part.c:317-448   (unique columns, first-occurrence order)
//   setGlobalInvarSitesVec  Pf/part.c:716-848
// The reference finds a site's pattern by scanning every earlier pattern
// (quadratic).  Here columns are gathered into a site-major byte matrix, hashed
// and looked up in an open-addressing table; the result (order, counts, index)
// is the same because a column is appended exactly when no earlier identical
// column exists, whatever the search method.
#include "engine.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/p4b200.h"

namespace p4b {

static thread_local char g_err[1024];
void setError(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *lastError() { return g_err; }

static const int kInvalidCode = -9999;

Part *newPart(int nTax, int nChar, const char *equateSymbols, int nEquates, const char *symbols, int dim)
{
    if (nTax <= 0 || nChar <= 0 || dim <= 0 || nEquates < 0 || !symbols) {
        setError("newPart: bad arguments nTax=%d nChar=%d dim=%d nEquates=%d", nTax, nChar, dim, nEquates);
        return nullptr;
    }
    if (dim > 64) {
        setError("newPart: dim=%d; this engine supports at most 64 character states", dim);
        return nullptr;
    }
    if (nEquates > 60) {   // equate codes are -64+k and must stay below QMARK_CODE (Pf/defines.h:33-36)
        setError("newPart: nEquates=%d exceeds 60", nEquates);
        return nullptr;
    }
    Part *p = new Part();
    p->dim = dim;
    p->nTax = nTax;
    p->nChar = nChar;
    p->nEquates = nEquates;
    p->symbols.assign(symbols, strnlen(symbols, dim));
    p->symbols.resize(dim, '\0');
    if (nEquates > 0 && equateSymbols) p->equateSymbols.assign(equateSymbols, strnlen(equateSymbols, nEquates));
    p->equateSymbols.resize(nEquates, '\0');
    const size_t cells = (size_t)nTax * (size_t)nChar;
    p->sequences.assign(cells, 0);
    p->patterns.assign(cells, 0);
    p->patternCounts.assign(nChar, 0);
    p->sequencePositionPatternIndex.assign(nChar, 0);
    p->equates.assign((size_t)nEquates * dim, 0);
    p->realEquateOfEquate.assign(nEquates, -1);
    return p;
}

void freePart(Part *p)
{
    if (!p) return;
    partDeviceFree(p);
    delete p;
}

// Which equates are not "N-like" (all ones)?  Only those need a column of their
// own in the leaf lookup tables; N-like ones behave exactly like a gap
// (Pf/p4_node.c:691-716).
static void classifyEquates(Part *p)
{
    p->nRealEquates = 0;
    for (int e = 0; e < p->nEquates; e++) {
        bool isN = true;
        for (int s = 0; s < p->dim; s++)
            if (!p->equates[(size_t)e * p->dim + s]) { isN = false; break; }
        p->realEquateOfEquate[e] = isN ? -1 : p->nRealEquates++;
    }
}

int pokeEquatesTable(Part *p, const char *table)
{
    if (p->nEquates == 0) return 0;           // reference: no equates array, nothing to do
    if (!table) { setError("pokeEquatesTable: NULL table"); return 1; }
    size_t k = 0;
    for (int i = 0; i < p->nEquates; i++)
        for (int j = 0; j < p->dim; j++) {
            if (table[k] == '\0') { setError("pokeEquatesTable: table shorter than nEquates*dim"); return 1; }
            if (table[k] == '1') p->equates[(size_t)i * p->dim + j] = 1;   // only ever sets, like the reference
            k++;
        }
    classifyEquates(p);
    p->version++;
    return 0;
}

int pokeSequences(Part *p, const char *s)
{
    if (!s) { setError("pokeSequences: NULL string"); return 1; }
    // Precedence of the reference: symbols first, then '-', then '?', then equates.
    int lut[256];
    for (int i = 0; i < 256; i++) lut[i] = kInvalidCode;
    for (int m = p->nEquates - 1; m >= 0; m--) lut[(unsigned char)p->equateSymbols[m]] = P4B_EQUATES_BASE + m;
    lut[(unsigned char)'?'] = P4B_QMARK_CODE;
    lut[(unsigned char)'-'] = P4B_GAP_CODE;
    for (int m = p->dim - 1; m >= 0; m--) lut[(unsigned char)p->symbols[m]] = m;
    lut[0] = kInvalidCode;
    const size_t cells = (size_t)p->nTax * (size_t)p->nChar;
    int *dst = p->sequences.data();
    for (size_t k = 0; k < cells; k++) {
        const int c = lut[(unsigned char)s[k]];
        if (c == kInvalidCode) {
            setError("part.c pokeSequences.  Got character '%c' (0x%02x) at taxon %zu, site %zu.  "
                     "It is neither in the symbols nor in the equates.",
                     s[k] ? s[k] : '0', (unsigned)(unsigned char)s[k], k / (size_t)p->nChar, k % (size_t)p->nChar);
            return 1;
        }
        dst[k] = c;
    }
    p->version++;
    return 0;
}

static inline uint64_t hashBytes(const int8_t *d, int n)
{
    // 64-bit multiply-xorshift over 8-byte words; quality only affects speed,
    // never the result (equal hashes are confirmed by a full compare).
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)n;
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        memcpy(&w, d + i, 8);
        h = (h ^ w) * 0xFF51AFD7ED558CCDull;
        h ^= h >> 32;
    }
    uint64_t w = 0;
    if (i < n) memcpy(&w, d + i, (size_t)(n - i));
    h = (h ^ w) * 0xC4CEB9FE1A85EC53ull;
    h ^= h >> 29;
    return h;
}

int makePatterns(Part *p)
{
    const int nTax = p->nTax, nChar = p->nChar;
    const int *seq = p->sequences.data();
    for (int i = 0; i < nChar; i++) {
        p->patternCounts[i] = 0;
        p->sequencePositionPatternIndex[i] = 0;
    }
    // 1. site-major byte copy of the alignment (codes fit in int8: -64..63).
    std::vector<int8_t> cols((size_t)nChar * nTax);
    const int B = 256;
    for (int i0 = 0; i0 < nChar; i0 += B) {
        const int i1 = i0 + B < nChar ? i0 + B : nChar;
        for (int j = 0; j < nTax; j++) {
            const int *row = seq + (size_t)j * nChar;
            for (int i = i0; i < i1; i++) cols[(size_t)i * nTax + j] = (int8_t)row[i];
        }
    }
    // 2. first-occurrence de-duplication.
    size_t cap = 16;
    while (cap < (size_t)nChar * 2) cap <<= 1;
    std::vector<int> table(cap, -1);          // pattern index, -1 = empty
    std::vector<int> firstSite;               // pattern -> first site carrying it
    firstSite.reserve(nChar);
    for (int i = 0; i < nChar; i++) {
        const int8_t *c = cols.data() + (size_t)i * nTax;
        size_t slot = (size_t)hashBytes(c, nTax) & (cap - 1);
        int found = -1;
        while (table[slot] >= 0) {
            const int cand = table[slot];
            if (memcmp(cols.data() + (size_t)firstSite[cand] * nTax, c, (size_t)nTax) == 0) { found = cand; break; }
            slot = (slot + 1) & (cap - 1);
        }
        if (found < 0) {
            found = (int)firstSite.size();
            firstSite.push_back(i);
            table[slot] = found;
        }
        p->patternCounts[found]++;
        p->sequencePositionPatternIndex[i] = found;
    }
    p->nPatterns = (int)firstSite.size();
    // 3. taxon-major pattern matrix, as the reference stores it.
    for (int j = 0; j < nTax; j++) {
        const int *row = seq + (size_t)j * nChar;
        int *out = p->patterns.data() + (size_t)j * nChar;
        for (int k = 0; k < p->nPatterns; k++) out[k] = row[firstSite[k]];
        for (int k = p->nPatterns; k < nChar; k++) out[k] = 0;
    }
    // Pf/part.c:419-427
    for (int j = 0; j < nTax; j++)
        for (int k = 0; k < p->nPatterns; k++)
            if (p->patterns[(size_t)j * nChar + k] >= p->dim) {
                setError("makePatterns: bad character %d", p->patterns[(size_t)j * nChar + k]);
                return 1;
            }
    p->version++;
    return 0;
}

int setGlobalInvarSitesVec(Part *p)
{
    const int nChar = p->nChar, dim = p->dim;
    if (p->globalInvarSitesVec.empty()) p->globalInvarSitesVec.assign(nChar, 0);
    if (p->globalInvarSitesArray.empty()) p->globalInvarSitesArray.assign((size_t)dim * nChar, 0);
    const uint64_t all = dim == 64 ? ~0ull : ((1ull << dim) - 1ull);
    std::vector<uint64_t> eqMask(p->nEquates, 0);
    for (int e = 0; e < p->nEquates; e++)
        for (int s = 0; s < dim; s++)
            if (p->equates[(size_t)e * dim + s]) eqMask[e] |= 1ull << s;
    std::vector<uint64_t> acc(p->nPatterns, all);
    for (int t = 0; t < p->nTax; t++) {
        const int *row = p->patterns.data() + (size_t)t * nChar;
        for (int k = 0; k < p->nPatterns; k++) {
            const int c = row[k];
            uint64_t m;
            if (c >= 0) m = 1ull << c;
            else if (c == -3 || c == P4B_GAP_CODE || c == P4B_QMARK_CODE) m = all;   // N_LIKE, gap, '?'
            else {
                const int e = c - P4B_EQUATES_BASE;
                if (e < 0 || e >= p->nEquates) { setError("setGlobalInvarSitesVec: bad code %d", c); return 1; }
                m = eqMask[e];
            }
            acc[k] &= m;
        }
    }
    for (int k = 0; k < p->nPatterns; k++) {
        p->globalInvarSitesVec[k] = __builtin_popcountll(acc[k]);
        for (int s = 0; s < dim; s++) p->globalInvarSitesArray[(size_t)s * nChar + k] = (int)((acc[k] >> s) & 1ull);
    }
    p->version++;
    return 0;
}

}  // namespace p4b
