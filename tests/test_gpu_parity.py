"""GPU parity: this engine, through the C ABI / pf mirror, against the reference's
own Pf engine (oracle/_ref) on the same seeded inputs.

Tolerances: log-likelihoods 1e-9 relative (BASELINE.json north_star).  P decks
2e-14 absolute: both engines rebuild P from an eigensystem, which is accurate to
about one ulp of 1.0 in absolute terms whatever the size of the entry.
An entry of 1e-8 thus carries a relative error near 1e-8 in BOTH engines, and
conditional likelihoods -- products of such entries -- inherit it: with each
engine's own P they are compared at 1e-9 relative to the largest entry of the
pattern.  test_cl_kernels_with_identical_P removes that effect by loading the
reference's P decks into this engine: the CL kernels alone then agree with the
reference to 1e-13 element by element.
"""
import numpy as np
import pytest

import ref_peek
from util import build_pair, max_rel_err, rel

pytestmark = pytest.mark.gpu

LNL_TOL = 1e-9
ARR_TOL = 1e-9


def _check_arrays(pkg, mine, twin):
    pf = pkg.pf
    for pNum, mp in enumerate(mine.model.parts):
        rp = ref_peek.part_arrays(twin.data.parts[pNum].cPart)
        nPat = rp["nPatterns"]
        for a, b in zip(mine.nodes, twin.nodes):
            if a is not mine.root:
                P1 = pf.getNodeBigP(a.cNode, pNum, mp.nGammaCat, mp.dim)
                P0 = ref_peek.node_bigP(b.cNode, pNum, mp.nGammaCat, mp.dim)
                assert np.max(np.abs(P1 - P0)) < 2e-14, "P deck of node %d" % a.nodeNum
            if not a.isLeaf:
                c1 = pf.getNodeCL(mine.cTree, a.cNode, pNum, mp.nGammaCat, mp.dim)
                c0 = ref_peek.node_cl(b.cNode, pNum, mp.nGammaCat, mp.dim, rp["nChar"], nPat)
                assert c1.shape == c0.shape
                scale = np.max(np.abs(c0), axis=(0, 1), keepdims=True)
                assert np.max(np.abs(c1 - c0) / scale) < ARR_TOL, "CL of node %d" % a.nodeNum


@pytest.mark.parametrize("cfg,kw", [
    (1, dict(nTax=12, nPatterns=500)),          # DNA GTR+I+G4
    (2, dict(nTax=24, nPatterns=3000)),         # DNA GTR+G4
    (3, dict(nTax=10, nPatterns=400)),          # protein LG+G4
    (4, dict(nTax=8, nPatterns=200)),           # protein NDCH2, 4 parts
])
def test_tree_loglike_matches_reference(pkg, ref_pf, cfg, kw):
    mine, twin = build_pair(pkg, ref_pf, cfg, **kw)
    got = mine.calcLogLike()
    want = twin.calcLogLike()
    assert rel(got, want) <= LNL_TOL
    for g, w in zip(mine.partLikes, twin.partLikes):
        assert rel(g, w) <= LNL_TOL
    _check_arrays(pkg, mine, twin)


@pytest.mark.parametrize("cfg,kw", [
    (1, dict(nTax=12, nPatterns=700)),
    (3, dict(nTax=10, nPatterns=300)),
    (4, dict(nTax=8, nPatterns=150)),
])
def test_cl_kernels_with_identical_P(pkg, ref_pf, cfg, kw):
    """CL recursion alone: P decks copied from the reference, so the only differences
    left are FMA contraction and summation order inside one dot product."""
    pf = pkg.pf
    mine, twin = build_pair(pkg, ref_pf, cfg, **kw)
    want = twin.calcLogLike()
    mine.calcLogLike()
    for pNum, mp in enumerate(mine.model.parts):
        for a, b in zip(mine.nodes, twin.nodes):
            if a is not mine.root:
                pf.setNodeBigP(a.cNode, pNum, ref_peek.node_bigP(b.cNode, pNum, mp.nGammaCat, mp.dim))
    got = pf.p4_treeLogLike(mine.cTree, 0)
    assert rel(got, want) <= 1e-13
    for pNum, mp in enumerate(mine.model.parts):
        rp = ref_peek.part_arrays(twin.data.parts[pNum].cPart)
        for a, b in zip(mine.nodes, twin.nodes):
            if not a.isLeaf:
                c1 = pf.getNodeCL(mine.cTree, a.cNode, pNum, mp.nGammaCat, mp.dim)
                c0 = ref_peek.node_cl(b.cNode, pNum, mp.nGammaCat, mp.dim, rp["nChar"], rp["nPatterns"])
                assert max_rel_err(c1, c0) < 1e-13, "CL of node %d" % a.nodeNum


def test_config1_full_size(pkg, ref_pf):
    """BASELINE config 1 at full size: 32 taxa, 10k patterns, GTR+I+G4."""
    mine, twin = build_pair(pkg, ref_pf, 1)
    assert pkg.pf.partPatternCount(mine.data.parts[0].cPart) == 10000
    assert rel(mine.calcLogLike(), twin.calcLogLike()) <= LNL_TOL


def test_site_likes(pkg, ref_pf):
    mine, twin = build_pair(pkg, ref_pf, 1, nTax=9, nPatterns=300)
    a = np.array(mine.getSiteLikes())
    b = np.array(twin.getSiteLikes())
    assert a.shape == b.shape
    assert max_rel_err(a, b) < 1e-11
    assert rel(mine.logLike, twin.logLike) <= LNL_TOL


def test_dirty_path_matches_full_recompute(pkg, ref_pf):
    """Chain.proposeSp's node-level path (p4/chain.py:668-688) against a full
    recompute (the reference's own self-check, p4/chain.py:265-286) and the reference."""
    mine, twin = build_pair(pkg, ref_pf, 2, nTax=20, nPatterns=1500)
    mine.calcLogLike()
    twin.calcLogLike()
    rng = np.random.default_rng(5)
    for _ in range(5):
        i = int(rng.integers(1, len(mine.nodes)))
        new = float(rng.uniform(0.001, 0.4))
        for t in (mine, twin):
            t.nodes[i].br.len = new
            t.nodes[i].br.lenChanged = True
        got = mine.recalcAfterBranchChange()
        want = twin.recalcAfterBranchChange()
        assert rel(got, want) <= LNL_TOL
        assert abs(got - mine.calcLogLike()) <= 1e-9 * abs(got)


def test_root_is_leaf(pkg, ref_pf):
    """A root that is itself a leaf (Pf/p4_tree.c:1199-1378)."""
    P = pkg
    rng = np.random.Generator(np.random.PCG64(77))
    tree = P.synth.random_tree(P.pf, 9, rng, root_is_leaf=True)
    assert tree.root.isLeaf
    mp = P.synth.dna_model_part(0, rng, 4, pInvar=0.1)
    sim_tree = P.synth.random_tree(P.pf, 9, np.random.Generator(np.random.PCG64(78)))
    aln = P.synth.make_alignment(P.pf, sim_tree, mp, 300, rng, "dna", gap_frac=0.05, ambig_frac=0.05)
    tree.attach(P.host.Data(P.pf, [aln]), P.host.Model(P.pf, [mp]))
    twin = P.host.clone_tree(tree, ref_pf)
    assert rel(tree.calcLogLike(), twin.calcLogLike()) <= LNL_TOL


def test_nonpositive_site_like_gives_sentinel(pkg, ref_pf):
    """like <= 0 -> the part returns -1.0e99 (Pf/p4_tree.c:1182).  1200 taxa of DNA
    underflow double precision (the reference has no scalers)."""
    P = pkg
    rng = np.random.Generator(np.random.PCG64(3))
    tree = P.synth.random_tree(P.pf, 1200, rng)
    mp = P.synth.dna_model_part(0, rng, 4)
    for n in tree.nodes:
        n.br.len = 0.5
    aln = P.synth.make_alignment(P.pf, tree, mp, 64, rng, "dna", gap_frac=0.0, ambig_frac=0.0, repeat=False)
    tree.attach(P.host.Data(P.pf, [aln]), P.host.Model(P.pf, [mp]))
    twin = P.host.clone_tree(tree, ref_pf)
    want = twin.calcLogLike()
    got = tree.calcLogLike()
    assert want == -1.0e99
    assert got == -1.0e99


def test_cur_prop_copy_and_verify(pkg):
    """p4_copyCondLikes / p4_copyBigPDecks / p4_copyModelPrams / p4_verifyIdentityOfTwoTrees
    as Chain.__init__ uses them (p4/chain.py:24-71)."""
    P = pkg
    cur = P.synth.build_config(P.pf, 2, nTax=14, nPatterns=800)
    prop = P.host.clone_tree(cur, P.pf, data=cur.data)
    cur.calcLogLike()
    prop.calcLogLike()
    pf = P.pf
    pf.p4_copyCondLikes(cur.cTree, prop.cTree, 1)
    pf.p4_copyBigPDecks(cur.cTree, prop.cTree, 1)
    pf.p4_copyModelPrams(cur.cTree, prop.cTree)
    prop.calcLogLike()
    assert pf.p4_verifyIdentityOfTwoTrees(cur.cTree, prop.cTree) == 0
    # a proposal on prop makes them differ; copying prop -> cur restores identity
    prop.nodes[3].br.len *= 1.7
    prop.nodes[3].br.lenChanged = True
    l1 = prop.recalcAfterBranchChange()
    assert pf.p4_verifyIdentityOfTwoTrees(cur.cTree, prop.cTree) == 1
    cur.nodes[3].br.len = prop.nodes[3].br.len
    cur.setCStuff()
    pf.p4_copyCondLikes(prop.cTree, cur.cTree, 1)
    pf.p4_copyBigPDecks(prop.cTree, cur.cTree, 1)
    pf.p4_copyModelPrams(prop.cTree, cur.cTree)
    assert pf.p4_verifyIdentityOfTwoTrees(cur.cTree, prop.cTree) == 0
    for pNum in range(cur.model.nParts):
        pf.p4_partLogLike(cur.cTree, cur.data.parts[pNum].cPart, pNum, 0)
    assert abs(float(sum(cur.partLikes)) - l1) <= 1e-12 * abs(l1)


@pytest.mark.parametrize("cfg,kw", [
    (1, dict(nTax=13, nPatterns=900)),
    (2, dict(nTax=40, nPatterns=5000)),
])
def test_fused_tree_kernel_equals_per_node_kernels(pkg, cfg, kw):
    """The one-launch whole-tree kernel does the same arithmetic in the same order
    as the per-node kernels: every CL must be bit-identical."""
    pf = pkg.pf
    tree = pkg.synth.build_config(pf, cfg, **kw)
    mp = tree.model.parts[0]
    try:
        pf.setFusedTreeKernel(0)
        a = tree.calcLogLike()
        cl_a = {n.nodeNum: pf.getNodeCL(tree.cTree, n.cNode, 0, mp.nGammaCat, mp.dim) for n in tree.nodes if not n.isLeaf}
        la = pf.kernelLaunchCount()
        pf.setFusedTreeKernel(1)
        b = tree.calcLogLike()
        lb = pf.kernelLaunchCount() - la
        cl_b = {n.nodeNum: pf.getNodeCL(tree.cTree, n.cNode, 0, mp.nGammaCat, mp.dim) for n in tree.nodes if not n.isLeaf}
    finally:
        pf.setFusedTreeKernel(1)
    assert lb == 1      # whole-tree CL + site likelihoods + the last CTA's fold: ONE launch (no P(t) job: every deck already has these inputs)
    assert rel(b, a) <= 1e-13
    for k in cl_a:
        assert np.array_equal(cl_a[k], cl_b[k]), "CL of node %d" % k


def test_star_tree_wide_polytomy(pkg, ref_pf):
    """A root with more children than one kernel step folds (chained steps)."""
    P = pkg
    rng = np.random.Generator(np.random.PCG64(11))
    nTax = 15
    nodes = [P.host.Node(i) for i in range(nTax + 1)]
    root = nodes[0]
    for i in range(1, nTax + 1):
        nodes[i].isLeaf, nodes[i].seqNum, nodes[i].parent = 1, i - 1, root
        nodes[i].br.len = float(rng.uniform(0.01, 0.3))
        if i < nTax:
            nodes[i].sibling = nodes[i + 1]
    root.leftChild = nodes[1]
    tree = P.host.Tree(P.pf, nodes, root)
    tree.setPreAndPostOrder()
    mp = P.synth.dna_model_part(0, rng, 4, pInvar=0.1)
    aln = P.synth.make_alignment(P.pf, tree, mp, 400, rng, "dna", gap_frac=0.03, ambig_frac=0.03)
    tree.attach(P.host.Data(P.pf, [aln]), P.host.Model(P.pf, [mp]))
    twin = P.host.clone_tree(tree, ref_pf)
    want = twin.calcLogLike()
    assert rel(tree.calcLogLike(), want) <= LNL_TOL
    P.pf.setFusedTreeKernel(0)
    try:
        assert rel(tree.calcLogLike(), want) <= LNL_TOL
    finally:
        P.pf.setFusedTreeKernel(1)


@pytest.mark.parametrize("cfg,kw", [
    (3, dict(nTax=14, nPatterns=700)),
    (4, dict(nTax=9, nPatterns=260)),
])
def test_tensor_core_kernel_equals_fma_kernel(pkg, ref_pf, cfg, kw):
    """20-state CL through FP64 mma.sync tiles against the plain FMA kernel and the reference."""
    pf = pkg.pf
    mine, twin = build_pair(pkg, ref_pf, cfg, **kw)
    want = twin.calcLogLike()
    try:
        pf.setTensorCoreKernel(0)
        a = mine.calcLogLike()
        cl_a = {(p, n.nodeNum): pf.getNodeCL(mine.cTree, n.cNode, p, mp.nGammaCat, mp.dim)
                for p, mp in enumerate(mine.model.parts) for n in mine.nodes if not n.isLeaf}
        pf.setTensorCoreKernel(1)
        b = mine.calcLogLike()
        cl_b = {(p, n.nodeNum): pf.getNodeCL(mine.cTree, n.cNode, p, mp.nGammaCat, mp.dim)
                for p, mp in enumerate(mine.model.parts) for n in mine.nodes if not n.isLeaf}
    finally:
        pf.setTensorCoreKernel(1)
    assert rel(a, want) <= LNL_TOL and rel(b, want) <= LNL_TOL
    assert rel(a, b) <= 1e-13
    for k in cl_a:
        assert max_rel_err(cl_b[k], cl_a[k]) < 1e-12, k


@pytest.mark.parametrize("cfg,kw", [(2, dict(nTax=30, nPatterns=2500)), (3, dict(nTax=21, nPatterns=900))])
def test_lnl_only_mode_is_transparent(pkg, ref_pf, cfg, kw):
    """p4b_setTreeStoresCL(0): whole-tree evaluations keep only the CLs they re-read; anything that
    later needs a CL gets it recomputed.  lnL, CLs and the dirty path must be unaffected (4 and 20 states)."""
    pf = pkg.pf
    mine, twin = build_pair(pkg, ref_pf, cfg, **kw)
    want = twin.calcLogLike()
    full = mine.calcLogLike()
    mp = mine.model.parts[0]
    cl_full = {n.nodeNum: pf.getNodeCL(mine.cTree, n.cNode, 0, mp.nGammaCat, mp.dim) for n in mine.nodes if not n.isLeaf}
    pf.setTreeStoresCL(mine.cTree, 0)
    lean = pf.p4_treeLogLike(mine.cTree, 0)
    assert lean == full and rel(lean, want) <= LNL_TOL
    # reading a CL after an lnL-only evaluation triggers the storing pass
    for n in mine.nodes:
        if not n.isLeaf:
            assert np.array_equal(pf.getNodeCL(mine.cTree, n.cNode, 0, mp.nGammaCat, mp.dim), cl_full[n.nodeNum])
    # lnL-only evaluation, then the node-level dirty path
    rng = np.random.default_rng(9)
    for _ in range(3):
        assert pf.p4_treeLogLike(mine.cTree, 0) == full
        i = int(rng.integers(1, len(mine.nodes)))
        old = mine.nodes[i].br.len
        for t in (mine, twin):
            t.nodes[i].br.len = old * 1.3
            t.nodes[i].br.lenChanged = True
        assert rel(mine.recalcAfterBranchChange(), twin.recalcAfterBranchChange()) <= LNL_TOL
        for t in (mine, twin):
            t.nodes[i].br.len = old
            t.nodes[i].br.lenChanged = True
        assert rel(mine.recalcAfterBranchChange(), want) <= LNL_TOL
        mine.calcLogLike()


def test_tree_heterogeneous_dna_ndch_ndrh(pkg, ref_pf):
    """Composition, rate matrix AND gamma shape vary over the tree (NDCH + NDRH, p4/model.py isHet): every node
    draws its own (comp, rMatrix, gdasrv) numbers; two data parts with different relative rates."""
    P, H = pkg, pkg.host
    rng = np.random.Generator(np.random.PCG64(99))
    tree = P.synth.random_tree(P.pf, 13, rng)
    mps, alns = [], []
    for pNum in range(2):
        mp = H.ModelPart(pNum, 4, 4)
        for _ in range(3):
            mp.comps.append(H.Comp(P.synth.normalise_comp(rng.dirichlet(8.0 * np.ones(4))), free=1))
        for _ in range(2):
            r = rng.dirichlet(4.0 * np.ones(6))
            mp.rMatrices.append(H.RMatrix("specified", r / r.sum(), free=1))
        for a in (0.3, 1.7):
            mp.gdasrvs.append(H.Gdasrv(4, a, free=1))
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
    twin = H.clone_tree(tree, ref_pf)
    got, want = tree.calcLogLike(), twin.calcLogLike()
    assert rel(got, want) <= LNL_TOL
    for g, w in zip(tree.partLikes, twin.partLikes):
        assert rel(g, w) <= LNL_TOL
    _check_arrays(pkg, tree, twin)
    # a dirty path through the heterogeneous model
    for t in (tree, twin):
        t.nodes[4].br.len *= 3.0
        t.nodes[4].br.lenChanged = True
    assert rel(tree.recalcAfterBranchChange(), twin.recalcAfterBranchChange()) <= LNL_TOL


def test_61_state_data_generic_kernel(pkg, ref_pf):
    """north_star names 61-state codon data; the reference has no codon model, so such data are a 'standard'
    datatype with 61 symbols: the any-dim path (SURVEY.md section 2 note)."""
    P, H = pkg, pkg.host
    rng = np.random.Generator(np.random.PCG64(61))
    symbols = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ012345678"
    assert len(symbols) == 61
    tree = P.synth.random_tree(P.pf, 7, rng)
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
    aln = H.Alignment(P.pf, seqs, symbols, {})
    mp = H.ModelPart(0, 61, 2)
    mp.comps.append(H.Comp(P.synth.normalise_comp(rng.dirichlet(20.0 * np.ones(61)))))
    mp.rMatrices.append(H.RMatrix("ones"))
    mp.gdasrvs.append(H.Gdasrv(2, 0.8))
    tree.attach(H.Data(P.pf, [aln]), H.Model(P.pf, [mp]))
    twin = H.clone_tree(tree, ref_pf)
    assert rel(tree.calcLogLike(), twin.calcLogLike()) <= LNL_TOL
