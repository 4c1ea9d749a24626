"""CPU tests of the product's host side (no GPU, no compute calls): the C-ABI library
loads and exports every declared symbol; character coding, pattern compression,
constant-site masks, gamma rates, Q and the eigensystem match the golden fixtures
and the reference engine bit for bit where they are integer work."""
import ctypes
import os
import re

import numpy as np
import pytest

import golden_io
import ref_peek
from util import rel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "p4b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(p4b_\w+)\s*\(", hdr)))
    assert len(names) > 70
    lib = ctypes.CDLL(pkg.pf.lib_path)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the compute path must fail loudly, not fall back."""
    pf = pkg.pf
    if pf.deviceCount() > 0:
        pytest.skip("a CUDA device is present")
    tree = pkg.synth.build_config(pf, 1, nTax=6, nPatterns=40)
    with pytest.raises(pf.P4bFatal) as e:
        tree.calcLogLike()
    assert "no CPU fallback" in str(e.value)


def test_bad_character_is_fatal(pkg):
    pf = pkg.pf
    p = pf.newPart(2, 4, "ry", 2, "acgt", 4)
    pf.pokeEquatesTable(p, "10100101")
    with pytest.raises(pf.P4bFatal):
        pf.pokeSequences(p, "acgtacgz")
    pf.freePart(p)


@pytest.mark.parametrize("name", golden_io.case_names())
def test_data_path_matches_golden(pkg, name):
    meta, arr = golden_io.load(name)
    tree = golden_io.build_tree(pkg, pkg.pf, meta)
    for pNum, part in enumerate(tree.data.parts):
        A = pkg.pf.partArrays(part.cPart)
        n = int(arr["p%d_nPatterns" % pNum])
        assert A["nPatterns"] == n
        assert np.array_equal(A["patterns"][:, :n], arr["p%d_patterns" % pNum])
        assert np.array_equal(A["patternCounts"][:n], arr["p%d_patternCounts" % pNum])
        assert np.array_equal(A["sequencePositionPatternIndex"], arr["p%d_sequencePositionPatternIndex" % pNum])
        assert np.array_equal(A["globalInvarSitesVec"][:n], arr["p%d_globalInvarSitesVec" % pNum])
        assert np.array_equal(A["globalInvarSitesArray"][:, :n], arr["p%d_globalInvarSitesArray" % pNum])
    tree.data.free()


@pytest.mark.parametrize("seed,nTax,nPat,kind", [(0, 12, 300, "dna"), (1, 7, 50, "dna"), (2, 20, 2000, "protein"), (3, 5, 3, "dna"),
                                                 (4, 40, 20000, "dna")])
def test_pattern_compression_bit_exact_vs_reference(pkg, ref_pf, seed, nTax, nPat, kind):
    P = pkg
    rng = np.random.Generator(np.random.PCG64(seed))
    t = P.synth.random_tree(P.pf, nTax, rng)
    mp = P.synth.dna_model_part(0, rng, 4, pInvar=0.2) if kind == "dna" else P.synth.protein_model_part(0, rng)
    aln = P.synth.make_alignment(P.pf, t, mp, nPat, rng, kind, gap_frac=0.05, ambig_frac=0.05)
    mine = aln._initParts()
    theirs = P.host.Alignment(ref_pf, aln.sequences, aln.symbols, aln.equates)._initParts()
    A, B = P.pf.partArrays(mine.cPart), ref_peek.part_arrays(theirs.cPart)
    n = A["nPatterns"]
    assert n == B["nPatterns"] == nPat
    for k in ("sequences", "patternCounts", "sequencePositionPatternIndex"):
        assert np.array_equal(A[k], B[k]), k
    assert np.array_equal(A["patterns"][:, :n], B["patterns"][:, :n])
    assert np.array_equal(A["globalInvarSitesVec"][:n], B["globalInvarSitesVec"][:n])
    assert np.array_equal(A["globalInvarSitesArray"][:, :n], B["globalInvarSitesArray"][:, :n])
    P.pf.freePart(mine.cPart)
    ref_pf.freePart(theirs.cPart)


EDGE_ALIGNMENTS = {
    "one column": ["a", "c", "g"],
    "two taxa": ["acgtacgt", "acgtaagt"],
    "every column the same": ["aaaaaaaa", "cccccccc", "gggggggg", "tttttttt"],
    "every column distinct": ["acgtacgtacgtacgt", "aaaaccccggggtttt", "acgtcagtgtactgca"],
    "all-gap, all-? and all-n columns between data": ["a-?na-c", "c-?nc-c", "g-?ng-t", "t-?nt-a"],
    "ambiguity codes only": ["rynrynry", "yrnyrnyr", "nnnnnnnn"],
    "gaps make equal columns unequal": ["aa-a", "cccc", "gg-g"],
    "first occurrence order (later columns repeat earlier ones out of order)": ["gatcgatcctag", "gatcgatcctag", "ccccaaaagggg"],
    "constant except for gaps and ambiguities": ["aaaaaa", "a-anra", "aa?ana", "aaaaya"],
}


@pytest.mark.parametrize("name", sorted(EDGE_ALIGNMENTS))
def test_pattern_compression_edge_cases_vs_reference(pkg, ref_pf, name):
    """Hand-made alignments at the edges of Pf/part.c:127-448 (makePatterns) and :716-848 (setGlobalInvarSitesVec): one column,
    one pattern, no repeated pattern, columns of gaps / missing / fully ambiguous characters, columns that differ only by a gap."""
    P = pkg
    seqs = EDGE_ALIGNMENTS[name]
    mine = P.host.Alignment(P.pf, seqs, P.host.DNA_SYMBOLS, P.host.DNA_EQUATES)._initParts()
    theirs = P.host.Alignment(ref_pf, seqs, P.host.DNA_SYMBOLS, P.host.DNA_EQUATES)._initParts()
    A, B = P.pf.partArrays(mine.cPart), ref_peek.part_arrays(theirs.cPart)
    n = A["nPatterns"]
    assert n == B["nPatterns"]
    for k in ("sequences", "sequencePositionPatternIndex"):
        assert np.array_equal(A[k], B[k]), k
    assert np.array_equal(A["patternCounts"][:n], B["patternCounts"][:n])
    assert np.array_equal(A["patterns"][:, :n], B["patterns"][:, :n])
    assert np.array_equal(A["globalInvarSitesVec"][:n], B["globalInvarSitesVec"][:n])
    assert np.array_equal(A["globalInvarSitesArray"][:, :n], B["globalInvarSitesArray"][:, :n])
    assert int(A["patternCounts"][:n].sum()) == len(seqs[0])
    P.pf.freePart(mine.cPart)
    ref_pf.freePart(theirs.cPart)


def test_pattern_compression_random_ragged_inputs_vs_reference(pkg, ref_pf):
    """300 random alignments -- 1 to 8 taxa, 1 to 400 columns, DNA and protein, drawn from small random subsets of the symbols, gaps,
    missing and every equate (so columns repeat heavily, in any order) -- compressed by this engine's hashed makePatterns and by the
    reference's scan (Pf/part.c:127-448): every array identical."""
    P = pkg
    rng = np.random.default_rng(11)
    for it in range(300):
        sym, eq = (P.host.DNA_SYMBOLS, P.host.DNA_EQUATES) if it % 3 else (P.host.PROTEIN_SYMBOLS, P.host.PROTEIN_EQUATES)
        alphabet = list(sym + "-?" + "".join(sorted(eq)))
        nTax, nChar = int(rng.integers(1, 9)), int(rng.integers(1, 400))
        letters = rng.choice(alphabet, size=int(rng.integers(1, len(alphabet) + 1)), replace=False)
        seqs = ["".join(rng.choice(letters, size=nChar)) for _ in range(nTax)]
        mine = P.host.Alignment(P.pf, seqs, sym, eq)._initParts()
        theirs = P.host.Alignment(ref_pf, seqs, sym, eq)._initParts()
        A, B = P.pf.partArrays(mine.cPart), ref_peek.part_arrays(theirs.cPart)
        n = A["nPatterns"]
        assert n == B["nPatterns"], (it, seqs)
        for k in ("sequences", "sequencePositionPatternIndex"):
            assert np.array_equal(A[k], B[k]), (it, k, seqs)
        for k in ("patternCounts", "globalInvarSitesVec"):
            assert np.array_equal(A[k][:n], B[k][:n]), (it, k, seqs)
        for k in ("patterns", "globalInvarSitesArray"):
            assert np.array_equal(A[k][:, :n], B[k][:, :n]), (it, k, seqs)
        P.pf.freePart(mine.cPart)
        ref_pf.freePart(theirs.cPart)


def test_unconstrained_loglike_vs_reference(pkg, ref_pf):
    """pf.getUnconstrainedLogLike (Pf/part.c:682-714): same number on clean data, fatal with any gap or ambiguity."""
    P = pkg
    rng = np.random.Generator(np.random.PCG64(8))
    t = P.synth.random_tree(P.pf, 9, rng)
    mp = P.synth.dna_model_part(0, rng, 4)
    aln = P.synth.make_alignment(P.pf, t, mp, 400, rng, "dna", gap_frac=0.0, ambig_frac=0.0)
    mine = aln._initParts()
    theirs = P.host.Alignment(ref_pf, aln.sequences, aln.symbols, aln.equates)._initParts()
    assert P.pf.getUnconstrainedLogLike(mine.cPart) == ref_pf.getUnconstrainedLogLike(theirs.cPart)
    P.pf.freePart(mine.cPart)
    ref_pf.freePart(theirs.cPart)
    gappy = P.synth.make_alignment(P.pf, t, mp, 50, rng, "dna", gap_frac=0.1, ambig_frac=0.0)._initParts()
    with pytest.raises(SystemExit):
        P.pf.getUnconstrainedLogLike(gappy.cPart)
    P.pf.freePart(gappy.cPart)


def test_rng_stream_is_mt19937_with_gsl_seeding(pkg, ref_pf):
    """pf.gsl_rng_* (the stream p4 hands to pf.p4_simulate): MT19937's published first output for its reference seed, and
    the reference engine's stream for the same seeds (seed 0 means 4357, GSL's convention)."""
    pf = pkg.pf
    g, r = pf.gsl_rng_get(), ref_pf.gsl_rng_get()
    pf.gsl_rng_set(g, 5489)
    assert int(pf.gsl_rng_uniform(g) * 4294967296) == 3499211612        # Matsumoto & Nishimura's mt19937ar.c, init_genrand(5489)
    for seed in (0, 1, 4357, 123456789):
        pf.gsl_rng_set(g, seed)
        ref_pf.gsl_rng_set(r, seed)
        assert [pf.gsl_rng_uniform(g) for _ in range(1500)] == [ref_pf.gsl_rng_uniform(r) for _ in range(1500)]
    # the bulk path the simulation draws from: same stream, whatever the block boundaries (624 words per state block)
    for n0, n1 in ((0, 5000), (1, 623), (623, 2), (624, 624), (100, 1248)):
        pf.gsl_rng_set(g, 7)
        ref_pf.gsl_rng_set(r, 7)
        got = [pf.gsl_rng_uniform(g) for _ in range(n0)] + list(pf.gsl_rng_uniform_array(g, n1)) + [pf.gsl_rng_uniform(g) for _ in range(3)]
        assert got == [ref_pf.gsl_rng_uniform(r) for _ in range(n0 + n1 + 3)]
    pf.gsl_rng_free(g)


@pytest.mark.parametrize("kind,gaps", [("dna", 0.05), ("dna", 0.0), ("protein", 0.04)])
def test_part_statistics_equal_the_reference(pkg, ref_pf, kind, gaps):
    """The data part's own statistics and views (csrc/partstats.cpp vs Pf/part.c): equal values, not close ones."""
    P = pkg
    rng = np.random.Generator(np.random.PCG64(31))
    t = P.synth.random_tree(P.pf, 9, rng)
    mp = P.synth.dna_model_part(0, rng, 4) if kind == "dna" else P.synth.protein_model_part(0, rng)
    aln = P.synth.make_alignment(P.pf, t, mp, 300, rng, kind, gap_frac=gaps, ambig_frac=gaps)
    mine = aln._initParts()
    theirs = P.host.Alignment(ref_pf, aln.sequences, aln.symbols, aln.equates)._initParts()
    a, b, pf = mine.cPart, theirs.cPart, P.pf
    assert pf.symbolSequences(a) == ref_pf.symbolSequences(b)
    for k in range(9):
        assert pf.singleSequenceBaseCounts(a, k) == list(ref_pf.singleSequenceBaseCounts(b, k))
        assert pf.partSequenceSitesCount(a, k) == ref_pf.partSequenceSitesCount(b, k)
    for sel in ([1] * 9, [1, 0, 1, 0, 0, 1, 1, 0, 1], [0] * 8 + [1]):
        for i, v in enumerate(sel):
            pf.pokePartTaxListAtIndex(a, v, i)
            ref_pf.pokePartTaxListAtIndex(b, v, i)
        assert pf.partComposition(a) == list(ref_pf.partComposition(b))
    assert pf.partMeanNCharsPerSite(a) == ref_pf.partMeanNCharsPerSite(b)
    assert pf.partSimpleConstantSitesCount(a) == ref_pf.partSimpleConstantSitesCount(b)
    assert pf.partBigXSquared(a) == ref_pf.partBigXSquared(b)
    P.pf.freePart(a)
    ref_pf.freePart(b)


def test_bootstrap_equals_the_reference_for_the_same_seed(pkg, ref_pf):
    """pf.bootstrapData: the same mt19937 seed resamples the same columns (gsl_rng_uniform_int), patterns re-made."""
    P, H = pkg, pkg.host
    rng = np.random.Generator(np.random.PCG64(12))
    t = P.synth.random_tree(P.pf, 7, rng)
    mp = P.synth.dna_model_part(0, rng, 4)
    alns = [P.synth.make_alignment(P.pf, t, mp, n, rng, "dna", gap_frac=0.03, ambig_frac=0.02) for n in (150, 90)]
    out = []
    for pf in (P.pf, ref_pf):
        src = H.Data(pf, [H.Alignment(pf, a.sequences, a.symbols, a.equates) for a in alns])
        dst = H.Data(pf, [H.Alignment(pf, a.sequences, a.symbols, a.equates) for a in alns])
        src._setCStuff()
        dst._setCStuff()
        g = pf.gsl_rng_get()
        pf.gsl_rng_set(g, 99)
        pf.bootstrapData(src.cData, dst.cData, g)
        out.append([(pf.symbolSequences(p.cPart), pf.partPatternCount(p.cPart)) for p in dst.parts])
        assert out[-1][0][0] != pf.symbolSequences(src.parts[0].cPart)
        pf.gsl_rng_free(g)
        src.free()
        dst.free()
    assert out[0] == out[1]


@pytest.mark.parametrize("cfg,kw,pInvarFree", [(1, dict(nTax=10, nPatterns=250), 1), (3, dict(nTax=7, nPatterns=120), 0)])
def test_draw_anc_state_logic_vs_reference(pkg, ref_pf, cfg, kw, pInvarFree):
    """The host logic of pf.p4_drawAncState (csrc/sim.cpp drawFromRootCL) fed with the REFERENCE's root CL: the same
    srandom seed gives the reference's draws at every site (the device only supplies the root CL; GPU parity in tests/test_gpu_sim.py)."""
    P = pkg
    twin = P.synth.build_config(ref_pf, cfg, **kw)
    mp = twin.model.parts[0]
    mp.pInvar.free = pInvarFree
    twin.calcLogLike()
    rp = ref_peek.part_arrays(twin.data.parts[0].cPart)
    cl = ref_peek.node_cl(twin.root.cNode, 0, mp.nGammaCat, mp.dim, rp["nChar"], rp["nPatterns"])
    aln = twin.data.alignments[0]
    mine = P.host.Alignment(P.pf, aln.sequences, aln.symbols, aln.equates)._initParts()
    pi = np.array(mp.comps[twin.root.parts[0].compNum].val, dtype=np.float64)
    pInvar = float(mp.pInvar.val)
    n = rp["nChar"]
    P.pf.reseedCRandomizer(123)
    got = [P.pf.drawAncStateFromCL(mine.cPart, k, mp.nGammaCat, pInvar, pInvarFree, pi, cl) for k in range(n)]
    ref_pf.reseedCRandomizer(123)
    d = np.empty(4, dtype=np.int32)
    want = []
    for k in range(n):
        ref_pf.p4_drawAncState(twin.cTree, 0, k, d)
        want.append(d.tolist())
    assert got == want
    assert any(w[2] for w in want) == (cfg == 1)          # invariant draws do occur in the pInvar case
    P.pf.freePart(mine.cPart)


def test_gamma_rates_bit_identical_to_reference(pkg, ref_pf):
    for alpha in (0.1, 0.2, 0.5, 0.73, 1.0, 2.7, 10.0, 100.0, 299.0):
        for K in (2, 3, 4, 5, 8, 16):
            f1, r1, f2, r2 = np.zeros(K), np.zeros(K), np.zeros(K), np.zeros(K)
            pkg.pf.gdasrvCalcRates_np(K, alpha, f1, r1)
            ref_pf.gdasrvCalcRates_np(K, alpha, f2, r2)
            assert np.array_equal(r1, r2) and np.array_equal(f1, f2), (alpha, K)
    rng = np.random.default_rng(5)              # and 1000 random shapes between 0.001 and 300, 2 to 12 categories
    for _ in range(1000):
        alpha, K = float(np.exp(rng.uniform(np.log(1e-3), np.log(300.0)))), int(rng.integers(2, 13))
        f1, r1, f2, r2 = np.zeros(K), np.zeros(K), np.zeros(K), np.zeros(K)
        pkg.pf.gdasrvCalcRates_np(K, alpha, f1, r1)
        ref_pf.gdasrvCalcRates_np(K, alpha, f2, r2)
        assert np.array_equal(r1, r2) and np.array_equal(f1, f2), (alpha, K)


def _expm_longdouble(Q, t):
    A = Q.astype(np.longdouble) * np.longdouble(t)
    n = A.shape[0]
    s = max(0, int(np.ceil(np.log2(max(1e-30, float(np.abs(A).sum(1).max()))))) + 8)
    A = A / np.longdouble(2.0) ** s
    E = np.eye(n, dtype=np.longdouble)
    term = np.eye(n, dtype=np.longdouble)
    for k in range(1, 30):
        term = term @ A / np.longdouble(k)
        E = E + term
    for _ in range(s):
        E = E @ E
    return E


@pytest.mark.parametrize("name", ["navidi_gtr_g4", "protein_lg_i_g4", "navidi_hetero_ndch2", "grouped_aas_dayhoff6"])
def test_q_and_eigensystem(pkg, name):
    """Q is bit-identical to the reference's; V diag(exp(lambda t)) V^-1 reproduces
    exp(Qt) (80-bit Taylor reference) to 5e-15, at least as well as the reference's P decks do."""
    pf = pkg.pf
    meta, arr = golden_io.load(name)
    tree = golden_io.build_tree(pkg, pf, meta)
    tree.model.allocCStuff()
    tree.model.setCStuff()
    mp = tree.model.parts[0]
    n1 = tree.nodes[1]
    c, r = n1.parts[0].compNum, n1.br.parts[0].rMatrixNum
    pf.p4_resetBQET(tree.model.cModel, 0, c, r)
    Q = np.zeros((mp.dim, mp.dim))
    pf.getBigQ(tree.model.cModel, mp.dim, 0, c, r, Q)
    assert np.array_equal(Q, arr["p0_bigQ_%d_%d" % (c, r)])
    V, Vi, lam = pf.getEig(tree.model.cModel, mp.dim, 0, c, r)
    assert np.max(np.abs(V @ Vi - np.eye(mp.dim))) < 5e-15
    assert np.max(np.abs(V @ np.diag(lam) @ Vi - Q)) < 5e-14
    worst_mine = worst_ref = 0.0
    for n in tree.nodes:
        if n is tree.root or n.parts[0].compNum != c or n.br.parts[0].rMatrixNum != r:
            continue
        g = mp.gdasrvs[n.br.parts[0].gdasrvNum] if mp.gdasrvs else None
        rates = meta["parts"][0]["gdasrvs"][n.br.parts[0].gdasrvNum]["rates"] if g else [1.0]
        for cat, rate in enumerate(rates):
            t = n.br.len * rate * mp.relRate / (1.0 - mp.pInvar.val)
            truth = _expm_longdouble(Q, t)
            mine = (V * np.exp(lam * t)[None, :]) @ Vi
            worst_mine = max(worst_mine, float(np.max(np.abs(mine - truth))))
            worst_ref = max(worst_ref, float(np.max(np.abs(arr["p0_bigP_%d" % n.nodeNum][cat] - truth))))
    assert worst_mine < 5e-15, (worst_mine, worst_ref)
    tree.model.free()
    tree.data.free()


def test_every_empirical_protein_model_equals_the_reference(pkg, ref_pf):
    """All eighteen empirical protein rate matrices (Pf/proteinModels.c through pf.getBigR, Pf/pfmodule.c:1270; p4/var.py:245-263) and
    the normalised Q each of them gives with its own composition (Pf/p4_tree.c:340-450): bit-identical to the reference's."""
    P, pf = pkg, pkg.pf
    assert len(P.host.RMATRIX_PROTEIN_SPEC) == 18
    for spec, code in sorted(P.host.RMATRIX_PROTEIN_SPEC.items(), key=lambda kv: kv[1]):
        A, B = np.zeros((20, 20)), np.zeros((20, 20))
        pf.getBigR(code, A)
        ref_pf.getBigR(code, B)
        assert np.array_equal(A, B) and np.array_equal(B, B.T) and B.max() > 0, spec
        rng = np.random.Generator(np.random.PCG64(code))
        tree = P.synth.random_tree(pf, 5, rng)
        mp = P.synth.protein_model_part(0, rng, spec, 2)
        aln = P.synth.make_alignment(pf, tree, mp, 20, rng, "protein")
        tree.attach(P.host.Data(pf, [aln]), P.host.Model(pf, [mp]))
        twin = P.host.clone_tree(tree, ref_pf)
        twin.calcLogLike()                              # the reference builds its Q on the way (CPU)
        tree.model.allocCStuff()
        tree.model.setCStuff()
        pf.p4_resetBQET(tree.model.cModel, 0, 0, 0)      # host side only: no device needed
        pf.getBigQ(tree.model.cModel, 20, 0, 0, 0, A)
        ref_pf.getBigQ(twin.model.cModel, 20, 0, 0, 0, B)
        assert np.array_equal(A, B), spec
        V, Vi, lam = pf.getEig(tree.model.cModel, 20, 0, 0, 0)
        assert np.max(np.abs(V @ np.diag(lam) @ Vi - A)) < 5e-14 and np.max(np.abs(V @ Vi - np.eye(20))) < 5e-15, spec
        tree.model.free()
        tree.data.free()


def test_q_and_eigensystem_on_skewed_models_of_every_size(pkg, ref_pf):
    """Sixty random models with 2 to 61 states, frequencies down to 1e-7 and exchangeabilities down to 1e-9: Q bit-identical to the
    reference's, and P(t) = V exp(lambda t) V^-1 from this engine's eigensystem (symmetric Jacobi, csrc/model.cpp) never further from
    exp(Qt) (80-bit Taylor) than 10x the error of the reference's own P decks (EISPACK general, Pf/eig.c:63-161)."""
    P, pf = pkg, pkg.pf
    syms = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789"
    rng = np.random.default_rng(3)
    for it in range(60):
        dim = int(rng.choice([2, 3, 4, 4, 5, 6, 11, 20, 21, 33, 61]))
        tree = P.synth.build_generic(pf, syms[:dim], 4, 12, 1, 1000 + it)
        mp = tree.model.parts[0]
        v = np.maximum(rng.dirichlet(float(rng.choice([0.05, 0.3, 1.0])) * np.ones(dim)), 1e-7)
        mp.comps[0].val[:] = P.synth.normalise_comp(v)
        r = np.maximum(rng.dirichlet(float(rng.choice([0.1, 1.0, 5.0])) * np.ones(dim * (dim - 1) // 2)), 1e-9)
        mp.rMatrices[0].val[:] = r / r.sum()
        twin = P.host.clone_tree(tree, ref_pf)
        twin.calcLogLike()
        tree.model.allocCStuff()
        tree.model.setCStuff()
        pf.p4_resetBQET(tree.model.cModel, 0, 0, 0)
        Q, Qref = np.zeros((dim, dim)), np.zeros((dim, dim))
        pf.getBigQ(tree.model.cModel, dim, 0, 0, 0, Q)
        ref_pf.getBigQ(twin.model.cModel, dim, 0, 0, 0, Qref)
        assert np.array_equal(Q, Qref), (it, dim)
        V, Vi, lam = pf.getEig(tree.model.cModel, dim, 0, 0, 0)
        mine = theirs = 0.0
        for a, b in zip(tree.nodes, twin.nodes):
            if a is tree.root:
                continue
            truth = _expm_longdouble(Q, a.br.len)
            mine = max(mine, float(np.max(np.abs((V * np.exp(lam * a.br.len)[None, :]) @ Vi - truth))))
            theirs = max(theirs, float(np.max(np.abs(ref_peek.node_bigP(b.cNode, 0, 1, dim)[0] - truth))))
        assert mine <= max(10.0 * theirs, 5e-14), (it, dim, mine, theirs)
        tree.model.free()
        tree.data.free()


def test_fast_bindings_mirror_the_ctypes_wrappers(pkg):
    """csrc/pfhot.c: the METH_FASTCALL bindings of the per-node calls are installed over the ctypes wrappers and keep
    their contract: same names, arity checked, engine errors raise P4bFatal (no GPU needed: NULL handles)."""
    pf = pkg.pf
    assert pf.use_fast_bindings(True), "_pfhot was not built (g.build())"
    for name in pf._HOT_NAMES:
        assert getattr(pf, name) is getattr(pf._hot, name)
    with pytest.raises(pf.P4bFatal) as e:
        pf.p4_setBrLen(0, 0.1)
    assert "NULL handle" in str(e.value)
    with pytest.raises(TypeError):
        pf.p4_setBrLen(0)
    with pytest.raises(pf.P4bFatal):
        pf.p4_setNodeRelation(0, 0, np.int32(3))        # numpy integers are accepted as integers
    with pytest.raises(TypeError):
        pf.p4_setNodeRelation(0, 0, "x")
    try:
        assert not pf.use_fast_bindings(False)
        assert pf.p4_setBrLen is pf._ctypes_versions["p4_setBrLen"]
        with pytest.raises(pf.P4bFatal):
            pf.p4_setBrLen(0, 0.1)
    finally:
        pf.use_fast_bindings(True)
