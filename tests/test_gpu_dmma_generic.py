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

SYM61 = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ012345678"
SYM24 = "abcdefghijklmnopqrstuvwx"


def _evolve(rng, tree, dim, nSites, mut=0.35):
    """Sequences with phylogenetic signal: states copied down the tree with per-branch change probability."""
    states = {}
    root = tree.root
    states[root.nodeNum] = rng.integers(dim, size=nSites)
    out = {}
    for i in tree.preOrder:
        if i < 0 or i == root.nodeNum:
            continue
        n = tree.nodes[i]
        s = states[n.parent.nodeNum].copy()
        m = rng.random(nSites) < min(0.9, mut * (0.3 + 5.0 * n.br.len))
        s[m] = rng.integers(dim, size=int(m.sum()))
        states[i] = s
        if n.isLeaf:
            out[n.seqNum] = s
    return [out[k] for k in sorted(out)]


def _build(pkg, symbols, nTax, nSites, nCat, seed, equates=None, pInvar=0.0, polytomy=False):
    P, H = pkg, pkg.host
    rng = np.random.Generator(np.random.PCG64(seed))
    dim = len(symbols)
    tree = P.synth.random_tree(P.pf, nTax, rng)
    if polytomy:        # collapse internal branches below the root until the root has 5 children
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
        tree.setPreAndPostOrder()
    lut = np.frombuffer(symbols.encode(), dtype=np.uint8)
    rows = _evolve(rng, tree, dim, nSites)
    seqs = []
    eqChars = sorted((equates or {}).keys())
    for s in rows:
        chars = lut[s].copy()
        chars[rng.random(nSites) < 0.02] = ord("-")
        for e in eqChars:
            chars[rng.random(nSites) < 0.01] = ord(e)
        seqs.append(chars.tobytes())
    aln = H.Alignment(P.pf, seqs, symbols, equates or {})
    mp = H.ModelPart(0, dim, nCat)
    mp.comps.append(H.Comp(P.synth.normalise_comp(rng.dirichlet(20.0 * np.ones(dim)))))
    r = rng.dirichlet(3.0 * np.ones(dim * (dim - 1) // 2))
    mp.rMatrices.append(H.RMatrix("specified", r / r.sum()))
    if nCat > 1:
        mp.gdasrvs.append(H.Gdasrv(nCat, 0.7))
    mp.pInvar = H.PInvar(pInvar)
    tree.attach(H.Data(P.pf, [aln]), H.Model(P.pf, [mp]))
    return tree


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
