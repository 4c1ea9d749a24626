"""Parity AT THE SHAPES THAT ARE BENCHMARKED ("What's weak" item 1).

The whole-tree kernels pick their launch shape from the shard size, so a test at 3 k patterns never runs the code a
1 M-pattern evaluation (or its 250 k / 125 k shards on 4 / 8 GPUs) runs.  Here every launch shape of the 4-state
kernel is forced (`pf.setFusedVariant`) at a small size, at the two shard sizes of the multi-GPU runs, and the
BASELINE configs 3, 4 and 5 are evaluated at their FULL size; each time the comparison surface is the reference's
own: per-site likelihoods (Pf/p4_tree.c:1015-1021) and conditional likelihoods (Pf/p4_node.c:636-857) of the
reference engine (oracle/_ref) on a sample of the alignment's columns -- site likelihoods do not depend on the
other columns, so the reference only has to evaluate the sample.
"""
import numpy as np
import pytest

import ref_peek
from util import rel

pytestmark = pytest.mark.gpu


def _sample_twin(pkg, ref_pf, tree, nCols, seed):
    """The reference engine on `nCols` sampled columns of every alignment of `tree`: (twin tree, [cols per part])."""
    H = pkg.host
    rng = np.random.default_rng(seed)
    alns, colsPer = [], []
    for aln in tree.data.alignments:
        cols = np.sort(rng.choice(aln.length, size=min(nCols, aln.length), replace=False))
        seqs = [np.frombuffer(s if isinstance(s, (bytes, bytearray)) else s.encode(), dtype=np.uint8)[cols].tobytes() for s in aln.sequences]
        alns.append(H.Alignment(ref_pf, seqs, aln.symbols, aln.equates))
        colsPer.append(cols)
    twin = tree.dupe()
    twin.pf = ref_pf
    twin.data = H.Data(ref_pf, alns)
    twin.model = H.clone_model(tree.model, ref_pf)
    return twin, colsPer


def _free_twin(twin):
    twin.deleteCStuff()
    twin.model.free()
    twin.data.free()


def check_against_reference_sample(pkg, ref_pf, tree, nCols=2000, seed=7, clNodes=3, siteTol=1e-9, clTol=1e-9):
    """Site likelihoods of `tree` (this engine, every column) vs the reference on a column sample; CLs of a few
    internal nodes, pattern by pattern, through the two engines' site->pattern indices."""
    pf = pkg.pf
    site = np.array(tree.getSiteLikes())
    twin, colsPer = _sample_twin(pkg, ref_pf, tree, nCols, seed)
    want = np.array(twin.getSiteLikes())
    off = woff = 0
    rng = np.random.default_rng(seed + 1)
    internals = [n for n in tree.nodes if not n.isLeaf]
    worstSite = worstCL = 0.0
    for pNum, (aln, cols) in enumerate(zip(tree.data.alignments, colsPer)):
        got = site[off + cols]
        w = want[woff:woff + len(cols)]
        worstSite = max(worstSite, float(np.max(np.abs(got - w) / w)))
        mp = tree.model.parts[pNum]
        mine = pf.partArrays(tree.data.parts[pNum].cPart)
        theirs = ref_peek.part_arrays(twin.data.parts[pNum].cPart)
        patMine = mine["sequencePositionPatternIndex"][cols]
        patRef = theirs["sequencePositionPatternIndex"][:len(cols)]
        for n in rng.choice(len(internals), size=min(clNodes, len(internals)), replace=False):
            a = internals[int(n)]
            b = twin.nodes[a.nodeNum]
            c1 = pf.getNodeCL(tree.cTree, a.cNode, pNum, mp.nGammaCat, mp.dim)[:, :, patMine]
            c0 = ref_peek.node_cl(b.cNode, pNum, mp.nGammaCat, mp.dim, theirs["nChar"], theirs["nPatterns"])[:, :, patRef]
            scale = np.max(np.abs(c0), axis=(0, 1), keepdims=True)
            worstCL = max(worstCL, float(np.max(np.abs(c1 - c0) / scale)))
        off += aln.length
        woff += len(cols)
    _free_twin(twin)
    assert worstSite <= siteTol, "site likelihoods differ from the reference's: %.3e" % worstSite
    assert worstCL <= clTol, "conditional likelihoods differ from the reference's: %.3e" % worstCL
    # the log-likelihood is the sum of the site log-likelihoods
    assert rel(float(np.sum(np.log(site))), tree.logLike) <= 1e-12
    return worstSite, worstCL


@pytest.fixture
def auto_variant(pkg):
    yield
    pkg.pf.setFusedVariant(-1)


# ---- every launch shape of the 4-state whole-tree kernel, small case, full comparison with the reference ----------
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 10, 11, 12, 13, 14, 15, 16])
def test_every_launch_shape_full_comparison_20k(pkg, ref_pf, auto_variant, variant):
    pf, H = pkg.pf, pkg.host
    tree = pkg.synth.build_config(pf, 2, nTax=48, nPatterns=20000)
    twin = H.clone_tree(tree, ref_pf)
    want = twin.calcLogLike()
    pf.setFusedVariant(variant)
    got = tree.calcLogLike()
    assert rel(got, want) <= 1e-9
    mp = tree.model.parts[0]
    rp = ref_peek.part_arrays(twin.data.parts[0].cPart)
    for a, b in zip(tree.nodes, twin.nodes):
        if a.isLeaf:
            continue
        c1 = pf.getNodeCL(tree.cTree, a.cNode, 0, mp.nGammaCat, mp.dim)
        c0 = ref_peek.node_cl(b.cNode, 0, mp.nGammaCat, mp.dim, rp["nChar"], rp["nPatterns"])
        scale = np.max(np.abs(c0), axis=(0, 1), keepdims=True)
        assert np.max(np.abs(c1 - c0) / scale) <= 1e-9, "CL of node %d, launch shape %d" % (a.nodeNum, variant)
    site = np.array(tree.getSiteLikes())
    ws = np.array(twin.getSiteLikes())
    assert np.max(np.abs(site - ws) / ws) <= 1e-9
    tree.deleteCStuff()
    _free_twin(twin)


# ---- the shard sizes of the 4- and 8-GPU runs of BASELINE configs[1]: 250 k and 125 k patterns, 200 taxa ----------
@pytest.fixture(scope="module", params=[250000, 125000])
def shard_tree(request, pkg):
    tree = pkg.synth.build_config(pkg.pf, 2, nPatterns=request.param)
    tree.calcLogLike()
    yield tree
    tree.deleteCStuff()
    tree.model.free()
    tree.data.free()


@pytest.mark.parametrize("variant", [-1, 0, 1, 2, 10, 11, 12, 13, 14, 15, 16])
def test_shard_sizes_every_launch_shape(pkg, ref_pf, shard_tree, auto_variant, variant):
    pkg.pf.setFusedVariant(variant)
    check_against_reference_sample(pkg, ref_pf, shard_tree, nCols=1500, seed=11 + variant)


def test_shard_sizes_shapes_are_bit_identical(pkg, shard_tree, auto_variant):
    """Every launch shape of both kernel generations computes the same per-pattern numbers, bit for bit; only the order
    in which the per-pattern terms are summed differs (CTA size), so the totals agree to rounding."""
    pf = pkg.pf
    vals, sites = [], []
    for v in (0, 1, 2, 10, 11, 12, 13, 14, 15, 16):
        pf.setFusedVariant(v)
        sites.append(np.array(shard_tree.getSiteLikes()))
        vals.append(shard_tree.logLike)
    for s in sites[1:]:
        assert np.array_equal(s, sites[0])
    for v in vals[1:]:
        assert rel(v, vals[0]) <= 1e-13


# ---- BASELINE configs 3, 4, 5 at full size --------------------------------------------------------------------
def test_cfg3_full_size_site_likes_and_cls(pkg, ref_pf):
    """100 taxa x 200 k patterns, LG+G4: the 20-state whole-tree tensor-core kernel at the shape the bench quotes."""
    tree = pkg.synth.build_config(pkg.pf, 3)
    tree.calcLogLike()
    check_against_reference_sample(pkg, ref_pf, tree, nCols=300, seed=3, clNodes=2)
    tree.deleteCStuff()
    tree.model.free()
    tree.data.free()


def test_cfg4_full_size_site_likes_and_dirty_path(pkg, ref_pf):
    """60 taxa, 4 parts x 50 k patterns, a composition per node (NDCH2): full evaluation, then p4's whole-part
    protocol after ONE composition changed (p4/chain.py:305-380) against a full recompute and the reference."""
    pf = pkg.pf
    tree = pkg.synth.build_config(pf, 4)
    tree.calcLogLike()
    check_against_reference_sample(pkg, ref_pf, tree, nCols=150, seed=4, clNodes=2)
    # one leaf composition changed: whole-part protocol (memoised dirty path) == full evaluation == reference sample
    leaf = next(n for n in tree.nodes if n.isLeaf)
    mp = tree.model.parts[1]
    v = mp.comps[leaf.nodeNum].val
    v[0] += 0.01
    v[1] -= 0.01
    pf.p4_setPrams(tree.cTree, 1)
    for n in tree.iterInternalsPostOrder():
        pf.p4_setConditionalLikelihoodsOfInternalNodePart(n.cNode, 1)
    dirty = pf.p4_partLogLike(tree.cTree, tree.data.parts[1].cPart, 1, 0)
    full = tree.calcLogLike()
    assert rel(dirty, tree.partLikes[1]) <= 1e-13
    assert full == full
    check_against_reference_sample(pkg, ref_pf, tree, nCols=150, seed=5, clNodes=1)
    tree.deleteCStuff()
    tree.model.free()
    tree.data.free()


def test_cfg5_shape_batched_eight_trees(pkg, ref_pf):
    """100 taxa x 500 k patterns, GTR+G4, EIGHT trees sharing the data part evaluated by batched launches
    (`pf.treesPartLogLike`, what the 8 Metropolis-coupled chains of BASELINE configs[4] do; blockIdx.y = tree):
    every tree's value equals its own unbatched evaluation, and the site and conditional likelihoods of the first
    and the last tree match the reference's on a column sample."""
    pf, H = pkg.pf, pkg.host
    base = pkg.synth.build_config(pf, 5)
    trees = [base]
    rng = np.random.default_rng(55)
    try:
        for k in range(7):
            t = H.clone_tree(base, pf, data=base.data)
            for n in t.iterNodesNoRoot():
                n.br.len = float(n.br.len * rng.uniform(0.8, 1.25))
            trees.append(t)
        alone = [t.calcLogLike() for t in trees]
        pf.setMemoize(0)            # every queued call does its full work
        try:
            for t in trees:
                for n in t.iterInternalsPostOrder():
                    pf.p4_setConditionalLikelihoodsOfInternalNodePart(n.cNode, 0)
            n0 = pf.kernelLaunchCount()
            batched = pf.treesPartLogLike([t.cTree for t in trees], 0)
            assert pf.kernelLaunchCount() - n0 <= 4     # 8 x 98 steps: two batched launches (+ their folds), not eight
        finally:
            pf.setMemoize(1)
        for a, b in zip(alone, batched):
            assert rel(b, a) <= 1e-13
        for i in (0, len(trees) - 1):
            check_against_reference_sample(pkg, ref_pf, trees[i], nCols=1000, seed=50 + i, clNodes=2)
    finally:
        for t in trees[1:]:
            t.deleteCStuff()
            t.model.free()
        base.deleteCStuff()
        base.model.free()
        base.data.free()


# ---- any number of rate categories (the reference is generic in nCat, Pf/p4_node.c:652-654) -----------------------
@pytest.mark.parametrize("nCat", [2, 3, 5, 6, 7, 8])
def test_dna_whole_tree_kernel_any_ncat(pkg, ref_pf, nCat):
    P, H, pf = pkg, pkg.host, pkg.pf
    rng = np.random.Generator(np.random.PCG64(800 + nCat))
    tree = P.synth.random_tree(pf, 14, rng)
    mp = P.synth.dna_model_part(0, rng, nCat, pInvar=0.1 if nCat % 2 else 0.0)
    aln = P.synth.make_alignment(pf, tree, mp, 1500, rng, "dna", gap_frac=0.02, ambig_frac=0.02)
    tree.attach(H.Data(pf, [aln]), H.Model(pf, [mp]))
    twin = H.clone_tree(tree, ref_pf)
    want = twin.calcLogLike()
    got = tree.calcLogLike()
    assert pf.lastCLKernelName().startswith("cl_tree_dna2_kernel<%d," % nCat), pf.lastCLKernelName()
    assert rel(got, want) <= 1e-9
    rp = ref_peek.part_arrays(twin.data.parts[0].cPart)
    for a, b in zip(tree.nodes, twin.nodes):
        if a.isLeaf:
            continue
        c1 = pf.getNodeCL(tree.cTree, a.cNode, 0, nCat, 4)
        c0 = ref_peek.node_cl(b.cNode, 0, nCat, 4, rp["nChar"], rp["nPatterns"])
        scale = np.max(np.abs(c0), axis=(0, 1), keepdims=True)
        assert np.max(np.abs(c1 - c0) / scale) <= 1e-9
    # the per-node kernels (cl_dna_kernel<nCat>) leave the same CLs, bit for bit
    whole = [pf.getNodeCL(tree.cTree, a.cNode, 0, nCat, 4) for a in tree.nodes if not a.isLeaf]
    pf.setFusedTreeKernel(0)
    try:
        perNode = tree.calcLogLike()
        for a, w in zip([a for a in tree.nodes if not a.isLeaf], whole):
            assert np.array_equal(pf.getNodeCL(tree.cTree, a.cNode, 0, nCat, 4), w)
    finally:
        pf.setFusedTreeKernel(1)
    assert rel(perNode, got) <= 1e-13
    # a dirty path through the queue
    for t in (tree, twin):
        n = [x for x in t.iterNodesNoRoot()][5]
        n.br.len *= 1.7
        n.br.lenChanged = True
    assert rel(tree.recalcAfterBranchChange(), twin.recalcAfterBranchChange()) <= 1e-9
