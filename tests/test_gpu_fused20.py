"""The whole-tree 20-state kernel (cl_tree_aa_kernel, tree_aa.cuh: FP64 tensor cores, running CL in accumulator
registers, one rate category per CTA) against the one-launch-per-node kernels and the reference engine."""
import numpy as np
import pytest

import ref_peek
from util import build_pair, rel

pytestmark = pytest.mark.gpu
LNL_TOL = 1e-9


def _cls(pf, tree, pNum=0):
    mp = tree.model.parts[pNum]
    return {n.nodeNum: pf.getNodeCL(tree.cTree, n.cNode, pNum, mp.nGammaCat, mp.dim) for n in tree.iterInternalsPostOrder()}


@pytest.mark.parametrize("cfg,kw", [
    (3, dict(nTax=17, nPatterns=1000)),      # 1000 patterns: 32 groups of 32, the last one partly padding
    (3, dict(nTax=6, nPatterns=33)),
    (4, dict(nTax=9, nPatterns=300)),
])
def test_whole_tree_kernel_equals_per_node_kernels(pkg, ref_pf, cfg, kw):
    pf = pkg.pf
    mine, twin = build_pair(pkg, ref_pf, cfg, **kw)
    want = twin.calcLogLike()
    n0 = pf.kernelLaunchCount()
    got = mine.calcLogLike()
    fusedLaunches = pf.kernelLaunchCount() - n0
    a = [_cls(pf, mine, p) for p in range(mine.model.nParts)]
    pf.setFusedTreeKernel20(0)
    try:
        n0 = pf.kernelLaunchCount()
        got2 = mine.calcLogLike()
        nodeLaunches = pf.kernelLaunchCount() - n0
        b = [_cls(pf, mine, p) for p in range(mine.model.nParts)]
    finally:
        pf.setFusedTreeKernel20(1)
    assert rel(got, want) <= LNL_TOL
    assert rel(got, got2) <= 1e-12
    assert fusedLaunches < nodeLaunches
    for x, y in zip(a, b):
        for k in x:
            scale = np.max(np.abs(y[k]), axis=(0, 1), keepdims=True)
            assert np.max(np.abs(x[k] - y[k]) / scale) < 1e-13, "CL of node %d" % k


@pytest.mark.parametrize("nCat", [1, 2, 3, 6, 8])
def test_any_number_of_rate_categories(pkg, ref_pf, nCat):
    """The reference is generic in nCat (Pf/p4_node.c:652-654); the whole-tree kernel takes a category per CTA, so any
    number is served: lnL and EVERY node's CL array against the reference engine, gaps and ambiguity codes included."""
    P, pf = pkg, pkg.pf
    rng = np.random.Generator(np.random.PCG64(40 + nCat))
    tree = P.synth.random_tree(pf, 13, rng)
    mp = P.synth.protein_model_part(0, rng, "lg", nCat)
    aln = P.synth.make_alignment(pf, tree, mp, 700, rng, "protein", gap_frac=0.03, ambig_frac=0.03)
    tree.attach(P.host.Data(pf, [aln]), P.host.Model(pf, [mp]))
    twin = P.host.clone_tree(tree, ref_pf)
    got, want = tree.calcLogLike(), twin.calcLogLike()
    assert pf.lastCLKernelName().startswith("cl_tree_aa_kernel"), pf.lastCLKernelName()
    assert rel(got, want) <= LNL_TOL
    rp = ref_peek.part_arrays(twin.data.parts[0].cPart)
    for a, b in zip(tree.nodes, twin.nodes):
        if a.isLeaf:
            continue
        c1 = pf.getNodeCL(tree.cTree, a.cNode, 0, nCat, 20)
        c0 = ref_peek.node_cl(b.cNode, 0, nCat, 20, rp["nChar"], rp["nPatterns"])
        scale = np.max(np.abs(c0), axis=(0, 1), keepdims=True)
        assert np.max(np.abs(c1 - c0) / scale) < 1e-9, "CL of node %d" % a.nodeNum


def test_unequal_protein_parts_in_one_launch(pkg, ref_pf):
    """p4_treeLogLike sends all 20-state parts of a tree through ONE launch; the parts may differ in pattern count, leaf-table
    width (ambiguity codes present) and number of rate categories, and a DNA part may sit between them."""
    P, pf = pkg, pkg.pf
    rng = np.random.Generator(np.random.PCG64(77))
    tree = P.synth.random_tree(pf, 14, rng)
    mps = [P.synth.protein_model_part(0, rng, "lg", 4), P.synth.dna_model_part(1, rng, 4, pInvar=0.1),
           P.synth.protein_model_part(2, rng, "lg", 2), P.synth.protein_model_part(3, rng, "lg", 1)]
    alns = [P.synth.make_alignment(pf, tree, mps[0], 300, rng, "protein", gap_frac=0.02, ambig_frac=0.02),
            P.synth.make_alignment(pf, tree, mps[1], 500, rng, "dna"),
            P.synth.make_alignment(pf, tree, mps[2], 1000, rng, "protein", gap_frac=0.0, ambig_frac=0.0),
            P.synth.make_alignment(pf, tree, mps[3], 150, rng, "protein", gap_frac=0.05, ambig_frac=0.0)]
    tree.attach(P.host.Data(pf, alns), P.host.Model(pf, mps))
    twin = P.host.clone_tree(tree, ref_pf)
    n0 = pf.kernelLaunchCount()
    got, want = tree.calcLogLike(), twin.calcLogLike()
    merged = pf.kernelLaunchCount() - n0
    assert rel(got, want) <= LNL_TOL
    for a, b in zip(tree.partLikes, twin.partLikes):
        assert rel(a, b) <= LNL_TOL
    rp = [ref_peek.part_arrays(p.cPart) for p in twin.data.parts]
    for pNum in (0, 2, 3):
        nCat = mps[pNum].nGammaCat
        for a, b in zip(tree.nodes, twin.nodes):
            if a.isLeaf:
                continue
            c1 = pf.getNodeCL(tree.cTree, a.cNode, pNum, nCat, 20)
            c0 = ref_peek.node_cl(b.cNode, pNum, nCat, 20, rp[pNum]["nChar"], rp[pNum]["nPatterns"])
            scale = np.max(np.abs(c0), axis=(0, 1), keepdims=True)
            assert np.max(np.abs(c1 - c0) / scale) < 1e-9, "part %d, CL of node %d" % (pNum, a.nodeNum)
    pf.setFusedTreeKernel20(0)
    try:
        n0 = pf.kernelLaunchCount()
        assert rel(tree.calcLogLike(), got) <= 1e-12
        assert merged < pf.kernelLaunchCount() - n0
    finally:
        pf.setFusedTreeKernel20(1)
    # lnL-only evaluations take the same route
    pf.setTreeStoresCL(tree.cTree, 0)
    assert rel(pf.p4_treeLogLike(tree.cTree, 0), want) <= LNL_TOL
    pf.setTreeStoresCL(tree.cTree, 1)
    assert rel(tree.calcLogLike(), want) <= LNL_TOL


def test_cl_arrays_match_reference(pkg, ref_pf):
    pf = pkg.pf
    mine, twin = build_pair(pkg, ref_pf, 3, nTax=11, nPatterns=400)
    mine.calcLogLike()
    twin.calcLogLike()
    rp = ref_peek.part_arrays(twin.data.parts[0].cPart)
    for a, b in zip(mine.nodes, twin.nodes):
        if a.isLeaf:
            continue
        c1 = pf.getNodeCL(mine.cTree, a.cNode, 0, 4, 20)
        c0 = ref_peek.node_cl(b.cNode, 0, 4, 20, rp["nChar"], rp["nPatterns"])
        scale = np.max(np.abs(c0), axis=(0, 1), keepdims=True)
        assert np.max(np.abs(c1 - c0) / scale) < 1e-9, "CL of node %d" % a.nodeNum


def test_dirty_path_protein(pkg, ref_pf):
    mine, twin = build_pair(pkg, ref_pf, 3, nTax=15, nPatterns=500)
    mine.calcLogLike()
    twin.calcLogLike()
    rng = np.random.default_rng(8)
    for _ in range(6):
        i = int(rng.integers(1, len(mine.nodes)))
        new = float(rng.uniform(0.001, 0.4))
        for t in (mine, twin):
            t.nodes[i].br.len = new
            t.nodes[i].br.lenChanged = True
        got = mine.recalcAfterBranchChange()
        want = twin.recalcAfterBranchChange()
        assert rel(got, want) <= LNL_TOL
        assert rel(got, mine.calcLogLike()) <= 1e-12


def test_polytomy_protein(pkg, ref_pf):
    """A 7-way star: the root's children go through the kernel two at a time (chained steps)."""
    P = pkg
    rng = np.random.Generator(np.random.PCG64(5))
    base = P.synth.random_tree(P.pf, 7, rng)
    mp = P.synth.protein_model_part(0, rng, "lg", 4)
    aln = P.synth.make_alignment(P.pf, base, mp, 200, rng, "protein", gap_frac=0.03, ambig_frac=0.03)
    H = P.host
    nodes = [H.Node(i) for i in range(8)]
    root = nodes[0]
    for k in range(7):
        leaf = nodes[1 + k]
        leaf.isLeaf, leaf.seqNum, leaf.parent = 1, k, root
        leaf.br.len = 0.02 + 0.03 * k
        if k:
            nodes[k].sibling = leaf
    root.leftChild = nodes[1]
    tree = H.Tree(P.pf, nodes, root)
    tree.attach(H.Data(P.pf, [aln]), H.Model(P.pf, [mp]))
    twin = H.clone_tree(tree, ref_pf)
    assert rel(tree.calcLogLike(), twin.calcLogLike()) <= LNL_TOL


@pytest.mark.parametrize("mode", ["pipelined", True, False])
def test_ndch2_chain_matches_reference(pkg, ref_pf, mode):
    """A short MCMC on the tree-heterogeneous protein config: allCompsDir proposals re-solve the changed
    eigensystems, topology moves shuffle which node uses which composition.  All three ways of reading the chains'
    likelihoods (pipelined, one batched launch per part, one chain after the other)."""
    def make(pf):
        tree = pkg.synth.build_config(pf, 4, nTax=8, nPatterns=120)
        return pkg.mcmc.Mcmc(tree, nChains=3, seed=3)
    a = make(pkg.pf).run(40, batched=mode)
    b = make(ref_pf).run(40)
    for (ga, la), (gb, lb) in zip(a, b):
        for x, y in zip(la, lb):
            assert rel(x, y) <= LNL_TOL
