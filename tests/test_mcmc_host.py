"""The Mcmc / Chain call protocol (p4_phylogenetics_b200/mcmc.py) driven through the reference's own
Pf engine on the CPU: the protocol itself must be sound before it is used to compare engines."""
import numpy as np
import pytest


def _mcmc(pkg, pf, nChains, seed, **kw):
    tree = pkg.synth.build_config(pf, 5, nTax=kw.pop("nTax", 10), nPatterns=kw.pop("nPatterns", 200))
    return pkg.mcmc.Mcmc(tree, nChains=nChains, seed=seed, **kw)


def test_protocol_keeps_cur_and_prop_identical(pkg, ref_pf):
    """After every generation the prop tree must equal the cur tree (the reference's own check,
    p4/chain.py:1542-1560), and the running lnL must equal a full recompute (p4/chain.py:265-286)."""
    m = _mcmc(pkg, ref_pf, 2, 11)
    for _ in range(40):
        m.run(1)
        for c in m.chains:
            assert ref_pf.p4_verifyIdentityOfTwoTrees(c.curTree.cTree, c.propTree.cTree) == 0
    for c in m.chains:
        was = c.curTree.logLike
        assert abs(c.curTree.calcLogLike() - was) <= 1e-9 * abs(was)
    assert sum(p.nAcceptances for p in m.proposals) > 0
    assert {p.name for p in m.proposals} >= {"local", "eTBR", "allBrLens", "allCompsDir", "allRMatricesDir", "gdasrv"}


def test_topology_moves_keep_a_valid_tree(pkg, ref_pf):
    m = _mcmc(pkg, ref_pf, 1, 3, nTax=9, nPatterns=100)
    only = [p for p in m.proposals if p.name in ("local", "eTBR")]
    m.proposals = only
    w = np.array([p.weight for p in only])
    m._cum = np.cumsum(w / w.sum())
    m.run(60)
    t = m.chains[0].curTree
    leaves = sorted(n.seqNum for n in t.nodes if n.isLeaf)
    assert leaves == list(range(9))
    seen = set()
    stack = [t.root]
    while stack:
        n = stack.pop()
        assert n.nodeNum not in seen
        seen.add(n.nodeNum)
        for c in n.iterChildren():
            assert c.parent is n
            stack.append(c)
    assert len(seen) == len(t.nodes)
    assert len(list(t.root.iterChildren())) == 3
    for n in t.nodes:
        if not n.isLeaf and n is not t.root:
            assert len(list(n.iterChildren())) == 2


def test_same_seed_same_chain(pkg, ref_pf):
    a = _mcmc(pkg, ref_pf, 2, 5).run(25)
    b = _mcmc(pkg, ref_pf, 2, 5).run(25)
    assert a == b
