"""CPU tests of the oracle: the plain-C restatement (oracle/pf_port.c) against

  (a) the golden fixtures, which were produced by the reference's own p4 package
      on the reference's example inputs (tests/golden/make_golden.py), and
  (b) the reference's own Pf engine (oracle/_ref) on seeded synthetic inputs.

These pin the oracle before anything is allowed to trust it."""
import numpy as np
import pytest

import golden_io
import pf_port
from util import max_rel_err, rel


@pytest.mark.parametrize("name", golden_io.case_names())
def test_port_matches_golden(pkg, name):
    meta, arr = golden_io.load(name)
    tree = golden_io.build_tree(pkg, None, meta)
    lnL, partLikes, extra = pf_port.tree_loglike(tree, want_arrays=True)
    assert rel(lnL, meta["lnL"]) <= 1e-12
    for g, w in zip(partLikes, meta["partLikes"]):
        assert rel(g, w) <= 1e-12
    site = []
    for pNum, ex in enumerate(extra):
        c = ex["compress"]
        n = int(arr["p%d_nPatterns" % pNum])
        assert c["nPatterns"] == n
        assert np.array_equal(c["patterns"][:, :n], arr["p%d_patterns" % pNum])
        assert np.array_equal(c["patternCounts"][:n], arr["p%d_patternCounts" % pNum])
        assert np.array_equal(c["sequencePositionPatternIndex"], arr["p%d_sequencePositionPatternIndex" % pNum])
        assert np.array_equal(c["globalInvarSitesVec"][:n], arr["p%d_globalInvarSitesVec" % pNum])
        assert np.array_equal(c["globalInvarSitesArray"][:, :n], arr["p%d_globalInvarSitesArray" % pNum])
        for gi, g in enumerate(meta["parts"][pNum]["gdasrvs"]):
            assert np.array_equal(ex["rates"][gi], np.array(g["rates"])), "gamma rates must be bit-identical"
        for n_ in tree.nodes:
            if n_ is not tree.root:
                assert np.max(np.abs(ex["P"][n_.nodeNum] - arr["p%d_bigP_%d" % (pNum, n_.nodeNum)])) < 1e-14
            if not n_.isLeaf:
                want = arr["p%d_cl_%d" % (pNum, n_.nodeNum)]
                scale = np.max(np.abs(want), axis=(0, 1), keepdims=True)
                assert np.max(np.abs(ex["cl"][n_.nodeNum] - want) / scale) < 1e-9
        site.append(ex["patLikes"][c["sequencePositionPatternIndex"]])
    assert max_rel_err(np.concatenate(site), np.array(meta["siteLikes"])) < 1e-11


def test_port_bigq_matches_golden(pkg):
    meta, arr = golden_io.load("navidi_gtr_g4")
    tree = golden_io.build_tree(pkg, None, meta)
    mp = tree.model.parts[0]
    Q = pf_port.big_q(pf_port._big_r(mp.rMatrices[0], 4), mp.comps[0].val)
    assert np.array_equal(Q, arr["p0_bigQ_0_0"])


@pytest.mark.parametrize("cfg,kw", [
    (1, dict(nTax=10, nPatterns=300)),
    (3, dict(nTax=7, nPatterns=120)),
    (4, dict(nTax=6, nPatterns=60)),
])
def test_port_matches_reference_engine(pkg, ref_pf, cfg, kw):
    twin = pkg.synth.build_config(ref_pf, cfg, **kw)
    want = twin.calcLogLike()
    got = pf_port.tree_loglike(twin)
    assert rel(got, want) <= 1e-12


def test_port_root_leaf_and_sentinel(pkg, ref_pf):
    P = pkg
    rng = np.random.Generator(np.random.PCG64(77))
    tree = P.synth.random_tree(ref_pf, 9, rng, root_is_leaf=True)
    mp = P.synth.dna_model_part(0, rng, 4, pInvar=0.1)
    sim_tree = P.synth.random_tree(ref_pf, 9, np.random.Generator(np.random.PCG64(78)))
    aln = P.synth.make_alignment(ref_pf, sim_tree, mp, 200, rng, "dna", gap_frac=0.05, ambig_frac=0.05)
    tree.attach(P.host.Data(ref_pf, [aln]), P.host.Model(ref_pf, [mp]))
    assert rel(pf_port.tree_loglike(tree), tree.calcLogLike()) <= 1e-12


def test_long_double_port_agrees_with_double_port(pkg):
    """The 80-bit variant of the port (used to check the engine's scalers) equals the double one in range."""
    tree = pkg.synth.build_config(None, 1, nTax=12, nPatterns=200)
    a = pf_port.tree_loglike(tree)
    b = pf_port.tree_loglike(tree, long_double=True)
    assert rel(a, b) <= 1e-13


@pytest.mark.parametrize("cfg,kw", [(1, dict(nTax=9, nPatterns=200)), (3, dict(nTax=6, nPatterns=80))])
def test_port_newton_step_matches_reference_engine(pkg, ref_pf, cfg, kw):
    """The port's restatement of the Newton-Raphson sums (pfport_branch_derivs): lnL through every branch equals the tree's
    lnL, the first derivative agrees with a central difference of the port's own lnL, and p4_newtNode driven by the port's
    derivatives lands on the branch length the reference's own p4_newtNode lands on (leaf and internal branches)."""
    import ref_peek
    P = pkg
    twin = P.synth.build_config(ref_pf, cfg, **kw)
    rng = np.random.default_rng(cfg)
    for n in twin.iterNodesNoRoot():
        n.br.len = float(min(max(n.br.len * np.exp(rng.normal(0.0, 0.6)), 1e-4), 1.0))
    base = twin.calcLogLike()
    d = pf_port.branch_derivs(twin)
    for k, (l0, d1, d2) in d.items():
        assert rel(l0, base) <= 1e-10
    n0 = [n for n in twin.iterNodesNoRoot() if not n.isLeaf][0]
    v = n0.br.len
    h = 1e-4 * v
    n0.br.len = v + h
    lp = pf_port.tree_loglike(twin)
    n0.br.len = v - h
    lm = pf_port.tree_loglike(twin)
    n0.br.len = v
    assert abs((lp - lm) / (2 * h) - d[n0.nodeNum][1]) <= 1e-4 * max(1.0, abs(d[n0.nodeNum][1]))
    # the reference's p4_newtNode on the same branches: cl2 down the whole tree first, then one node at a time from the same state
    ref_pf.p4_newtSetup(twin.cTree)
    lib = ref_peek.newt_lib()
    stack, order = [twin.root], []
    while stack:
        n = stack.pop()
        if n is not twin.root:
            order.append(n)
        stack.extend(reversed(list(n.iterChildren())))
    picks = [n for n in order if n.isLeaf][:2] + [n for n in order if not n.isLeaf][:2]
    for n in picks:
        start = n.br.len
        twin.calcLogLike()                                   # CLs and P decks for the current lengths
        for q in order:
            lib.p4_setNodeCL2(twin.cTree, q.cNode)
        lib.p4_newtNode(n.cNode, 1.0e-5, 1.0e-8, 3.0)
        want = ref_peek.node_brlen(n.cNode)
        n.br.len = start
        got = pf_port.newt_node(twin, n, 1.0e-5)
        assert abs(got - want) <= 1e-7 * max(want, 1e-3), (n.nodeNum, start, got, want)
        n.br.len = start
