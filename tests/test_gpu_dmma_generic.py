"""GPU: the generic tensor-core whole-tree kernel (csrc/tree_dmma.cuh, 21..64 states) against the reference engine.

north_star names "61-state codon" data.  The reference has no codon model (SURVEY.md section 2 note): such data are a
'standard' datatype whose symbols give dim = 61, and the reference's CL loop is generic in dim (Pf/p4_node.c:636-857).
Compared here: log-likelihood, every node's CL array and the site likelihoods, with the whole-tree kernel and with the
per-node kernels (`pf.setFusedTreeKernel20(0)` switches the tensor-core whole-tree kernels off), on
  * 61 symbols, 12 taxa, 10,000+ patterns, 2 rate categories, gaps and two ambiguity codes (equates);
  * 61 symbols with 4 categories and pInvar, a 5-way polytomy (chained steps) and a dirty path;
  * 24 symbols (the padded-to-32 instantiation), 3 categories.
"""
import numpy as np
import pytest

import ref_peek
from util import rel

pytestmark = pytest.mark.gpu

SYM61 = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ012345678"      # = synth.SYMBOLS_61
SYM24 = "abcdefghijklmnopqrstuvwx"


def _build(pkg, symbols, nTax, nSites, nCat, seed, equates=None, pInvar=0.0, polytomy=False):
    P, H = pkg, pkg.host
    tree = None
    if polytomy:        # collapse internal branches below the root until the root has 5 children
        rng = np.random.Generator(np.random.PCG64(seed + 1000))
        tree = P.synth.random_tree(P.pf, nTax, rng)
        while sum(1 for _ in tree.root.iterChildren()) < 5:
            c = next(k for k in tree.root.iterChildren() if not k.isLeaf)
            kids = list(c.iterChildren())
            sibs = [k for k in tree.root.iterChildren() if k is not c] + kids
            for k in kids:
                k.parent = tree.root
            tree.root.leftChild = sibs[0]
            for a, b in zip(sibs, sibs[1:] + [None]):
                a.sibling = b
            c.parent = c.leftChild = c.sibling = None
            tree.nodes = [n for n in tree.nodes if n is not c]
            for k, n in enumerate(tree.nodes):
                n.nodeNum = k
        tree = H.Tree(P.pf, tree.nodes, tree.root)
    return P.synth.build_generic(P.pf, symbols, nTax, nSites, nCat, seed, equates=equates, pInvar=pInvar, tree=tree)


def _compare(pkg, ref_pf, tree, clTol=1e-9):
    pf, H = pkg.pf, pkg.host
    twin = H.clone_tree(tree, ref_pf)
    want = twin.calcLogLike()
    got = tree.calcLogLike()
    assert pf.lastCLKernelName().startswith("cl_tree_dmma_kernel"), pf.lastCLKernelName()
    assert rel(got, want) <= 1e-9
    mp = tree.model.parts[0]
    rp = ref_peek.part_arrays(twin.data.parts[0].cPart)
    for a, b in zip(tree.nodes, twin.nodes):
        if a.isLeaf:
            continue
        c1 = pf.getNodeCL(tree.cTree, a.cNode, 0, mp.nGammaCat, mp.dim)
        c0 = ref_peek.node_cl(b.cNode, 0, mp.nGammaCat, mp.dim, rp["nChar"], rp["nPatterns"])
        scale = np.max(np.abs(c0), axis=(0, 1), keepdims=True)
        assert np.max(np.abs(c1 - c0) / scale) <= clTol, "CL of node %d" % a.nodeNum
    site = np.array(tree.getSiteLikes())
    ws = np.array(twin.getSiteLikes())
    assert np.max(np.abs(site - ws) / ws) <= 1e-9
    # the per-node kernels agree with the whole-tree kernel
    pf.setFusedTreeKernel20(0)
    try:
        perNode = tree.calcLogLike()
        assert not pf.lastCLKernelName().startswith("cl_tree_dmma_kernel")
    finally:
        pf.setFusedTreeKernel20(1)
    assert rel(perNode, got) <= 1e-12
    return twin


def test_61_states_10k_patterns(pkg, ref_pf):
    tree = _build(pkg, SYM61, 12, 16000, 2, 61, equates={"!": "abcd", "#": "XYZ012"})
    assert pkg.pf.partPatternCount(tree.data.parts[0].cPart) >= 10000
    _compare(pkg, ref_pf, tree)


def test_61_states_polytomy_pinvar_and_dirty_path(pkg, ref_pf):
    tree = _build(pkg, SYM61, 11, 1500, 4, 62, pInvar=0.15, polytomy=True)
    assert sum(1 for _ in tree.root.iterChildren()) == 5
    twin = _compare(pkg, ref_pf, tree)
    for t in (tree, twin):
        n = [x for x in t.iterNodesNoRoot() if not x.isLeaf][0]
        n.br.len *= 2.5
        n.br.lenChanged = True
    a, b = tree.recalcAfterBranchChange(), twin.recalcAfterBranchChange()
    assert rel(a, b) <= 1e-9
    assert rel(tree.calcLogLike(), a) <= 1e-12


def test_24_states_padded_to_32(pkg, ref_pf):
    tree = _build(pkg, SYM24, 9, 2500, 3, 24, equates={"!": SYM24[:5]})
    _compare(pkg, ref_pf, tree)
