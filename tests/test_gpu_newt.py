"""GPU parity of the Newton-Raphson branch-length step against the reference's own
Pf/p4_treeNewt.c, reached through ctypes on oracle/_ref (its pf module wraps only p4_newtSetup and the two
drivers): cl2 arrays node by node, the derivatives against finite differences of the engine's own lnL, whole
p4_newtAround rounds branch length by branch length, and the four-round schedule of p4_newtAndBrentPowellOpt.

Tolerances: cl2 like CL arrays (1e-9 of the pattern's largest entry, each engine with its own P decks);
log-likelihoods 1e-9 relative; branch lengths after the same Newton rounds 1e-6 relative (each is the fixed
point of an iteration on derivatives that agree to ~1e-12, stopped by the same tests)."""
import numpy as np
import pytest

import ref_peek
from util import build_pair, rel

pytestmark = pytest.mark.gpu


def _perturb(trees, seed, sd=0.8):
    rng = np.random.default_rng(seed)
    f = {n.nodeNum: float(np.exp(rng.normal(0.0, sd))) for n in trees[0].iterNodesNoRoot()}
    for t in trees:
        for n in t.iterNodesNoRoot():
            n.br.len = min(max(n.br.len * f[n.nodeNum], 1e-4), 2.0)


def _top_down(tree):
    out, stack = [], [tree.root]
    while stack:
        n = stack.pop()
        if n is not tree.root:
            out.append(n)
        stack.extend(reversed(list(n.iterChildren())))
    return out


def _hetero_pair(pkg, ref_pf):
    """Two parts with different relRates, pInvar in one of them, compositions / rate matrices / gamma
    shapes that vary over the tree."""
    P, H = pkg, pkg.host
    rng = np.random.Generator(np.random.PCG64(77))
    tree = P.synth.random_tree(P.pf, 11, rng)
    mps, alns = [], []
    for pNum in range(2):
        mp = H.ModelPart(pNum, 4, 4)
        for _ in range(3):
            mp.comps.append(H.Comp(P.synth.normalise_comp(rng.dirichlet(8.0 * np.ones(4))), free=0))
        for _ in range(2):
            r = rng.dirichlet(4.0 * np.ones(6))
            mp.rMatrices.append(H.RMatrix("specified", r / r.sum(), free=0))
        for a in (0.3, 1.7):
            mp.gdasrvs.append(H.Gdasrv(4, a, free=0))
        mp.pInvar = H.PInvar(0.1 if pNum == 0 else 0.0)
        mp.relRate = 0.7 if pNum == 0 else 1.4
        mp.isHet = 1
        mps.append(mp)
        sim = P.synth.dna_model_part(0, rng, 4)
        alns.append(P.synth.make_alignment(P.pf, tree, sim, 300 + 100 * pNum, rng, "dna", gap_frac=0.02, ambig_frac=0.02))
    model = H.Model(P.pf, mps)
    model.doRelRates = 1
    tree.attach(H.Data(P.pf, alns), model)
    for n in tree.nodes:
        for pNum in range(2):
            n.parts[pNum].compNum = int(rng.integers(3))
            n.br.parts[pNum].rMatrixNum = int(rng.integers(2))
            n.br.parts[pNum].gdasrvNum = int(rng.integers(2))
    return tree, H.clone_tree(tree, ref_pf)


CASES = [
    (1, dict(nTax=12, nPatterns=500)),          # DNA GTR+I+G4: constant-site term
    (2, dict(nTax=24, nPatterns=3000)),         # DNA GTR+G4
    (3, dict(nTax=10, nPatterns=400)),          # protein LG+G4: cl2 through the tensor-core node kernel
    (4, dict(nTax=8, nPatterns=200)),           # protein NDCH2, 4 parts: a composition per node
]


@pytest.mark.parametrize("cfg,kw", CASES)
def test_cl2_matches_reference(pkg, ref_pf, cfg, kw):
    pf = pkg.pf
    mine, twin = build_pair(pkg, ref_pf, cfg, **kw)
    assert rel(mine.calcLogLike(), twin.calcLogLike()) <= 1e-9
    pf.p4_newtSetup(mine.cTree)
    ref_pf.p4_newtSetup(twin.cTree)
    lib = ref_peek.newt_lib()
    for b in _top_down(twin):                    # a node's cl2 needs its parent's
        lib.p4_setNodeCL2(twin.cTree, b.cNode)
    lnL = mine.logLike
    for a, b in zip(_top_down(mine), _top_down(twin)):
        assert a.nodeNum == b.nodeNum
        d = pf.newtDerivs(a.cNode)               # recomputes cl2 from the root's child down to a
        assert rel(d[0], lnL) <= 1e-10, "lnL through the branch of node %d" % a.nodeNum
        for pNum, mp in enumerate(mine.model.parts):
            rp = ref_peek.part_arrays(twin.data.parts[pNum].cPart)
            c1 = pf.getNodeCL2(mine.cTree, a.cNode, pNum, mp.nGammaCat, mp.dim)
            c0 = ref_peek.node_cl2(b.cNode, pNum, mp.nGammaCat, mp.dim, rp["nChar"], rp["nPatterns"])
            assert c1.shape == c0.shape
            scale = np.max(np.abs(c0), axis=(0, 1), keepdims=True)
            assert np.max(np.abs(c1 - c0) / scale) < 1e-9, "cl2 of node %d part %d" % (a.nodeNum, pNum)


@pytest.mark.parametrize("cfg,kw", CASES[:3])
def test_derivatives_match_finite_differences(pkg, cfg, kw):
    """d lnL/dv and d2 lnL/dv2 from the device (analytic, through cl2 and the derivative decks) against central
    differences of the engine's own full-tree log-likelihood; leaves with gaps and ambiguity codes included."""
    pf = pkg.pf
    tree = pkg.synth.build_config(pf, cfg, **kw)
    tree.calcLogLike()
    pf.p4_newtSetup(tree.cTree)
    nodes = [n for n in tree.iterNodesNoRoot()]
    picks = [n for n in nodes if n.isLeaf][:3] + [n for n in nodes if not n.isLeaf][:3]
    for n in picks:
        v = n.br.len
        l0, d1, d2 = pf.newtDerivs(n.cNode)
        h = max(1e-3 * v, min(1e-5, 0.1 * v))
        n.br.len = v + h
        lp = tree.calcLogLike()
        n.br.len = v - h
        lm = tree.calcLogLike()
        n.br.len = v
        base = tree.calcLogLike()
        assert rel(l0, base) <= 1e-10
        fd1 = (lp - lm) / (2 * h)
        fd2 = (lp - 2 * base + lm) / (h * h)
        assert abs(fd1 - d1) <= 1e-4 * max(1.0, abs(d1)), (n.nodeNum, d1, fd1)
        assert abs(fd2 - d2) <= 2e-3 * max(1.0, abs(d2)), (n.nodeNum, d2, fd2)


@pytest.mark.parametrize("cfg,kw", CASES)
def test_newt_around_matches_reference(pkg, ref_pf, cfg, kw):
    """One p4_newtAround(1e-5, 1e-7) from perturbed branch lengths on both engines: every branch goes through
    the same Newton iterations, guards included, so the lengths agree branch by branch."""
    pf = pkg.pf
    mine, twin = build_pair(pkg, ref_pf, cfg, **kw)
    _perturb((mine, twin), seed=cfg)
    start = mine.calcLogLike()
    assert rel(start, twin.calcLogLike()) <= 1e-9
    pf.p4_newtSetup(mine.cTree)
    ref_pf.p4_newtSetup(twin.cTree)
    got = pf.newtAround(mine.cTree, 1.0e-5, 1.0e-7)
    ref_peek.newt_lib().p4_newtAround(twin.cTree, 1.0e-5, 1.0e-7)
    want = ref_pf.p4_treeLogLike(twin.cTree, 0)
    assert got > start
    assert rel(got, want) <= 1e-9, (got, want)
    mineLens = pf.p4_getBrLens(mine.cTree)
    for a, b in zip(mine.iterNodesNoRoot(), twin.iterNodesNoRoot()):
        wantLen = ref_peek.node_brlen(b.cNode)
        assert abs(mineLens[a.nodeNum] - wantLen) <= 1e-6 * max(wantLen, 1e-3), (a.nodeNum, mineLens[a.nodeNum], wantLen)
    assert pf.newtIterations(mine.cTree) >= len(mineLens) - 1
    # the state left behind is a consistent tree: a plain evaluation of those lengths reproduces the value
    for n in mine.iterNodesNoRoot():
        n.br.len = mineLens[n.nodeNum]
    assert rel(mine.calcLogLike(), got) <= 1e-10


@pytest.mark.parametrize("nCat", [1, 3])
def test_newt_around_protein_other_category_counts(pkg, ref_pf, nCat):
    """The tensor-core derivative kernel for 20-state internal nodes (newt_aa_dmma_kernel) has a compile-time path for four
    categories and a run-time one for any other number; pInvar exercises the constant-site term of the finishing step."""
    P, pf = pkg, pkg.pf
    rng = np.random.Generator(np.random.PCG64(300 + nCat))
    tree = P.synth.random_tree(pf, 9, rng)
    mp = P.synth.protein_model_part(0, rng, "lg", nCat)
    mp.pInvar = P.host.PInvar(0.15)
    aln = P.synth.make_alignment(pf, tree, mp, 350, rng, "protein", gap_frac=0.02, ambig_frac=0.02)
    tree.attach(P.host.Data(pf, [aln]), P.host.Model(pf, [mp]))
    twin = P.host.clone_tree(tree, ref_pf)
    _perturb((tree, twin), seed=nCat)
    start = tree.calcLogLike()
    assert rel(start, twin.calcLogLike()) <= 1e-9
    pf.p4_newtSetup(tree.cTree)
    ref_pf.p4_newtSetup(twin.cTree)
    got = pf.newtAround(tree.cTree, 1.0e-5, 1.0e-7)
    ref_peek.newt_lib().p4_newtAround(twin.cTree, 1.0e-5, 1.0e-7)
    want = ref_pf.p4_treeLogLike(twin.cTree, 0)
    assert got > start and rel(got, want) <= 1e-9, (got, want)
    lens = pf.p4_getBrLens(tree.cTree)
    for a, b in zip(tree.iterNodesNoRoot(), twin.iterNodesNoRoot()):
        wantLen = ref_peek.node_brlen(b.cNode)
        assert abs(lens[a.nodeNum] - wantLen) <= 1e-6 * max(wantLen, 1e-3), (a.nodeNum, lens[a.nodeNum], wantLen)


def test_newt_around_two_parts_relrates_pinvar_hetero(pkg, ref_pf):
    """Two parts with relRates 0.7 / 1.4 (the reference's second-derivative factor carries an extra relRate in the
    gamma, no-pInvar branch, Pf/p4_node.c:510-514 -- step sizes depend on it), pInvar in one part, a model that
    varies over the tree."""
    pf = pkg.pf
    mine, twin = _hetero_pair(pkg, ref_pf)
    _perturb((mine, twin), seed=11, sd=0.5)
    assert rel(mine.calcLogLike(), twin.calcLogLike()) <= 1e-9
    pf.p4_newtSetup(mine.cTree)
    ref_pf.p4_newtSetup(twin.cTree)
    for eps, delta in ((1.0, 10.0), (1.0e-5, 1.0e-7)):
        got = pf.newtAround(mine.cTree, eps, delta)
        ref_peek.newt_lib().p4_newtAround(twin.cTree, eps, delta)
        want = ref_pf.p4_treeLogLike(twin.cTree, 0)
        assert rel(got, want) <= 1e-9, (eps, got, want)
        mineLens = pf.p4_getBrLens(mine.cTree)
        for a, b in zip(mine.iterNodesNoRoot(), twin.iterNodesNoRoot()):
            wantLen = ref_peek.node_brlen(b.cNode)
            assert abs(mineLens[a.nodeNum] - wantLen) <= 1e-6 * max(wantLen, 1e-3), (eps, a.nodeNum, mineLens[a.nodeNum], wantLen)


def test_newt_and_brent_powell_no_free_parameters_is_the_reference_schedule(pkg, ref_pf):
    """With no free model parameter p4_newtAndBrentPowellOpt is four p4_newtAround calls (Pf/p4_treeOpt.c:1214-1226):
    Tree.optLogLike(method='newtAndBrentPowell') then gives the reference's branch lengths, not just its optimum."""
    mine, twin = build_pair(pkg, ref_pf, 2, nTax=9, nPatterns=400)
    _perturb((mine, twin), seed=5)
    got = mine.optLogLike(method="newtAndBrentPowell")
    want = twin.optLogLike(method="newtAndBrentPowell")
    assert rel(got, want) <= 1e-9, (got, want)
    for a, b in zip(mine.iterNodesNoRoot(), twin.iterNodesNoRoot()):
        assert abs(a.br.len - b.br.len) <= 1e-6 * max(b.br.len, 1e-3), (a.nodeNum, a.br.len, b.br.len)


def test_newt_wide_polytomy_and_many_siblings(pkg, ref_pf):
    """A star tree: every leaf's cl2 is pi times the factors of ALL the other leaves (more siblings than one
    CL launch folds, so the launches chain)."""
    P = pkg
    pf = P.pf
    rng = np.random.Generator(np.random.PCG64(15))
    nTax = 15
    nodes = [P.host.Node(i) for i in range(nTax + 1)]
    root = nodes[0]
    for i in range(1, nTax + 1):
        nodes[i].isLeaf, nodes[i].seqNum, nodes[i].parent = 1, i - 1, root
        nodes[i].br.len = float(rng.uniform(0.01, 0.3))
        if i < nTax:
            nodes[i].sibling = nodes[i + 1]
    root.leftChild = nodes[1]
    tree = P.host.Tree(pf, nodes, root)
    tree.setPreAndPostOrder()
    mp = P.synth.dna_model_part(0, rng, 4, pInvar=0.1)
    aln = P.synth.make_alignment(pf, tree, mp, 400, rng, "dna", gap_frac=0.03, ambig_frac=0.03)
    tree.attach(P.host.Data(pf, [aln]), P.host.Model(pf, [mp]))
    twin = P.host.clone_tree(tree, ref_pf)
    _perturb((tree, twin), seed=3, sd=0.5)
    assert rel(tree.calcLogLike(), twin.calcLogLike()) <= 1e-9
    pf.p4_newtSetup(tree.cTree)
    ref_pf.p4_newtSetup(twin.cTree)
    got = pf.newtAround(tree.cTree, 1.0e-5, 1.0e-7)
    ref_peek.newt_lib().p4_newtAround(twin.cTree, 1.0e-5, 1.0e-7)
    assert rel(got, ref_pf.p4_treeLogLike(twin.cTree, 0)) <= 1e-9
    mineLens = pf.p4_getBrLens(tree.cTree)
    for a, b in zip(tree.iterNodesNoRoot(), twin.iterNodesNoRoot()):
        wantLen = ref_peek.node_brlen(b.cNode)
        assert abs(mineLens[a.nodeNum] - wantLen) <= 1e-6 * max(wantLen, 1e-3), (a.nodeNum, mineLens[a.nodeNum], wantLen)


def test_newt_setup_rejects_a_leaf_root(pkg):
    pf = pkg.pf
    rng = np.random.Generator(np.random.PCG64(4))
    tree = pkg.synth.random_tree(pf, 6, rng, root_is_leaf=True)
    mp = pkg.synth.dna_model_part(0, rng, 4)
    sim_tree = pkg.synth.random_tree(pf, 6, np.random.Generator(np.random.PCG64(5)))
    aln = pkg.synth.make_alignment(pf, sim_tree, mp, 100, rng, "dna")
    tree.attach(pkg.host.Data(pf, [aln]), pkg.host.Model(pf, [mp]))
    tree.calcLogLike()
    with pytest.raises(SystemExit):
        pf.p4_newtSetup(tree.cTree)


def test_newt_around_61_states(pkg, ref_pf):
    """61-symbol data (the any-dim derivative kernel with its decks read from global memory: 3 x 2 x 61 x 61 doubles
    do not fit the shared-memory budget), two rate categories, gaps at the tips."""
    P, H = pkg, pkg.host
    pf = P.pf
    rng = np.random.Generator(np.random.PCG64(61))
    symbols = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ012345678"
    tree = P.synth.random_tree(pf, 7, rng)
    lut = np.frombuffer(symbols.encode(), dtype=np.uint8)
    base = rng.integers(61, size=150)
    seqs = []
    for _ in range(7):
        s = base.copy()
        m = rng.random(150) < 0.4
        s[m] = rng.integers(61, size=int(m.sum()))
        chars = lut[s].copy()
        chars[rng.random(150) < 0.03] = ord("-")
        seqs.append(chars.tobytes())
    aln = H.Alignment(pf, seqs, symbols, {})
    mp = H.ModelPart(0, 61, 2)
    mp.comps.append(H.Comp(P.synth.normalise_comp(rng.dirichlet(20.0 * np.ones(61)))))
    mp.rMatrices.append(H.RMatrix("ones"))
    mp.gdasrvs.append(H.Gdasrv(2, 0.8))
    tree.attach(H.Data(pf, [aln]), H.Model(pf, [mp]))
    twin = H.clone_tree(tree, ref_pf)
    _perturb((tree, twin), seed=61, sd=0.5)
    assert rel(tree.calcLogLike(), twin.calcLogLike()) <= 1e-9
    pf.p4_newtSetup(tree.cTree)
    ref_pf.p4_newtSetup(twin.cTree)
    got = pf.newtAround(tree.cTree, 1.0e-5, 1.0e-7)
    ref_peek.newt_lib().p4_newtAround(twin.cTree, 1.0e-5, 1.0e-7)
    assert rel(got, ref_pf.p4_treeLogLike(twin.cTree, 0)) <= 1e-9
    mineLens = pf.p4_getBrLens(tree.cTree)
    for a, b in zip(tree.iterNodesNoRoot(), twin.iterNodesNoRoot()):
        wantLen = ref_peek.node_brlen(b.cNode)
        assert abs(mineLens[a.nodeNum] - wantLen) <= 1e-6 * max(wantLen, 1e-3), (a.nodeNum, mineLens[a.nodeNum], wantLen)
