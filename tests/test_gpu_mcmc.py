"""MCMC protocol on the GPU engine: queued node-level calls, batched chains, cur/prop transfer --
against the reference's own Pf engine running the same chain (same seed, same proposals)."""
import numpy as np
import pytest

from util import rel

pytestmark = pytest.mark.gpu

LNL_TOL = 1e-9


def _mcmc(pkg, pf, nChains, seed, nTax=14, nPatterns=700, bulk=False):
    tree = pkg.synth.build_config(pf, 5, nTax=nTax, nPatterns=nPatterns)
    tree.bulkSetCStuff = bulk
    return pkg.mcmc.Mcmc(tree, nChains=nChains, seed=seed)


def _same_trace(a, b, tol):
    assert len(a) == len(b)
    for (ga, la), (gb, lb) in zip(a, b):
        assert ga == gb
        for x, y in zip(la, lb):
            assert rel(x, y) <= tol, "generation %d: %r vs %r" % (ga, x, y)


def test_chain_matches_reference_engine(pkg, ref_pf):
    """The same 3-chain MCMCMC, 120 generations, on both engines: every chain's lnL after every generation
    agrees to 1e-9 relative, hence the same accept / reject / swap decisions throughout."""
    mine = _mcmc(pkg, pkg.pf, 3, 42)
    ref = _mcmc(pkg, ref_pf, 3, 42)
    ta = mine.run(120)
    tb = ref.run(120)
    _same_trace(ta, tb, LNL_TOL)
    for p, q in zip(mine.proposals, ref.proposals):
        assert (p.name, p.nProposals, p.nAcceptances) == (q.name, q.nProposals, q.nAcceptances)
    assert (mine.nSwapAttempts, mine.nSwaps) == (ref.nSwapAttempts, ref.nSwaps)
    for c in mine.chains:
        assert pkg.pf.p4_verifyIdentityOfTwoTrees(c.curTree.cTree, c.propTree.cTree) == 0
        was = c.curTree.logLike
        assert rel(c.curTree.calcLogLike(), was) <= 1e-12


def test_queued_calls_equal_immediate_calls(pkg):
    """p4b_setDeferredNodeCalls(0) launches every node-level call at once (one kernel per node); the default
    queues them into one step-list launch.  Same chain either way."""
    pf = pkg.pf
    a = _mcmc(pkg, pf, 2, 7).run(60, batched=False)
    n0 = pf.kernelLaunchCount()
    _mcmc(pkg, pf, 2, 7).run(10, batched=False)
    queued = pf.kernelLaunchCount() - n0
    pf.setDeferredNodeCalls(0)
    try:
        b = _mcmc(pkg, pf, 2, 7).run(60, batched=False)
        n0 = pf.kernelLaunchCount()
        _mcmc(pkg, pf, 2, 7).run(10, batched=False)
        immediate = pf.kernelLaunchCount() - n0
    finally:
        pf.setDeferredNodeCalls(1)
    _same_trace(a, b, 1e-12)
    assert queued < immediate


def test_batched_chains_equal_sequential_chains(pkg):
    """Mcmc.run(batched=True): one pf.treesPartLogLike per part and generation for all chains."""
    pf = pkg.pf
    a = _mcmc(pkg, pf, 4, 9).run(50, batched=True)
    b = _mcmc(pkg, pf, 4, 9).run(50, batched=False)
    _same_trace(a, b, 1e-12)


def test_trees_part_loglike_mixed_batch(pkg):
    """pf.treesPartLogLike with trees whose queues do not all end at the root falls back per tree."""
    pf = pkg.pf
    m = _mcmc(pkg, pf, 3, 1)
    trees = [c.propTree for c in m.chains]
    want = [t.calcLogLike() for t in trees]
    # tree 0: a dirty path to the root queued; tree 1: nothing queued; tree 2: the whole tree queued
    n = [x for x in trees[0].iterInternalsPostOrder()][0]
    q = n
    path = [n]
    while q.parent:
        q = q.parent
        path.append(q)
    for x in path:
        pf.p4_setConditionalLikelihoodsOfInternalNodePart(x.cNode, 0)
    for x in trees[2].iterInternalsPostOrder():
        pf.p4_setConditionalLikelihoodsOfInternalNodePart(x.cNode, 0)
    got = pf.treesPartLogLike([t.cTree for t in trees], 0)
    for g, w in zip(got, want):
        assert rel(g, w) <= 1e-12
    # all three with a queue ending at the root: the batched launch proper
    for t in trees:
        for x in t.iterInternalsPostOrder():
            pf.p4_setConditionalLikelihoodsOfInternalNodePart(x.cNode, 0)
    n0 = pf.kernelLaunchCount()
    got = pf.treesPartLogLike([t.cTree for t in trees], 0)
    assert pf.kernelLaunchCount() - n0 == 1          # ONE launch for the three trees: CL steps, site likelihoods and each tree's fold
    for g, w, t in zip(got, want, trees):
        assert rel(g, w) <= 1e-12
        assert t.partLikes[0] == g


def test_queue_is_flushed_before_state_is_read_or_changed(pkg, ref_pf):
    pf = pkg.pf
    mine = pkg.synth.build_config(pf, 2, nTax=12, nPatterns=500)
    twin = pkg.host.clone_tree(mine, ref_pf)
    mine.calcLogLike()
    twin.calcLogLike()
    import ref_peek
    # change a branch, queue the path, then READ a CL on the path before any partLogLike
    for t in (mine, twin):
        n = t.nodes[5]
        n.br.len *= 2.5
        t.setCStuff()
        t.pf.p4_calculateBigPDecks(n.cNode)
        q = n
        while q.parent:
            q = q.parent
            if not q.isLeaf:
                t.pf.p4_setConditionalLikelihoodsOfInternalNodePart(q.cNode, 0)
    root = mine.root
    c1 = pf.getNodeCL(mine.cTree, root.cNode, 0, 4, 4)
    rp = ref_peek.part_arrays(twin.data.parts[0].cPart)
    c0 = ref_peek.node_cl(twin.root.cNode, 0, 4, 4, rp["nChar"], rp["nPatterns"])
    scale = np.max(np.abs(c0), axis=(0, 1), keepdims=True)
    assert np.max(np.abs(c1 - c0) / scale) < 1e-9
    got = pf.p4_partLogLike(mine.cTree, mine.data.parts[0].cPart, 0, 0)
    want = ref_pf.p4_partLogLike(twin.cTree, twin.data.parts[0].cPart, 0, 0)
    assert rel(got, want) <= LNL_TOL


def test_bulk_setcstuff_equals_per_node_calls(pkg):
    pf = pkg.pf
    a = _mcmc(pkg, pf, 2, 13, bulk=True).run(40)
    b = _mcmc(pkg, pf, 2, 13, bulk=False).run(40)
    assert a == b


def test_shared_buffers_equal_real_copies(pkg):
    """p4b_setSharedCondLikes(0): p4_copyCondLikes memcpys; default: twins reference one buffer."""
    pf = pkg.pf
    a = _mcmc(pkg, pf, 3, 21).run(60)
    pf.setSharedCondLikes(0)
    try:
        b = _mcmc(pkg, pf, 3, 21).run(60)
    finally:
        pf.setSharedCondLikes(1)
    assert a == b


def test_copy_on_write_keeps_the_twin_intact(pkg, ref_pf):
    """After p4_copyCondLikes(cur, prop) the two trees reference the same buffers; recomputing a path in
    prop must leave every CL of cur as it was, and vice versa."""
    pf = pkg.pf
    cur = pkg.synth.build_config(pf, 2, nTax=12, nPatterns=600)
    prop = pkg.host.clone_tree(cur, pf, data=cur.data)
    cur.calcLogLike()
    prop.calcLogLike()
    pf.p4_copyCondLikes(cur.cTree, prop.cTree, 1)
    pf.p4_copyBigPDecks(cur.cTree, prop.cTree, 1)
    pf.p4_copyModelPrams(cur.cTree, prop.cTree)
    internals = [n.nodeNum for n in cur.iterInternalsPostOrder()]
    before = {i: pf.getNodeCL(cur.cTree, cur.nodes[i].cNode, 0, 4, 4) for i in internals}
    l0 = cur.logLike
    for k, t, other in ((3, prop, cur), (7, cur, prop), (5, prop, cur)):
        t.nodes[k].br.len *= 1.9
        t.nodes[k].br.lenChanged = True
        lnew = t.recalcAfterBranchChange()
        assert lnew != l0
        # the other tree still evaluates to its own value from its own (partly shared) buffers
        for pNum in range(other.model.nParts):
            pf.p4_partLogLike(other.cTree, other.data.parts[pNum].cPart, pNum, 0)
        twin = pkg.host.clone_tree(other, ref_pf)
        assert rel(float(sum(other.partLikes)), twin.calcLogLike()) <= LNL_TOL
        if other is cur and k == 3:
            for i in internals:
                assert np.array_equal(pf.getNodeCL(cur.cTree, cur.nodes[i].cNode, 0, 4, 4), before[i])
    # many rounds of copy / recompute never run out of slots
    rng = np.random.default_rng(0)
    for _ in range(30):
        src, dst = (cur, prop) if rng.random() < 0.5 else (prop, cur)
        k = int(rng.integers(1, len(src.nodes)))
        src.nodes[k].br.len = float(rng.uniform(0.01, 0.3))
        src.nodes[k].br.lenChanged = True
        src.recalcAfterBranchChange()
        for x, y in zip(src.nodes, dst.nodes):
            y.br.len = x.br.len
        dst.setCStuff()
        pf.p4_copyCondLikes(src.cTree, dst.cTree, 1)
        pf.p4_copyBigPDecks(src.cTree, dst.cTree, 1)
        pf.p4_copyModelPrams(src.cTree, dst.cTree)
        assert pf.p4_verifyIdentityOfTwoTrees(src.cTree, dst.cTree) == 0
    assert rel(cur.calcLogLike(), pkg.host.clone_tree(cur, ref_pf).calcLogLike()) <= LNL_TOL


def test_pipelined_chains_equal_batched_chains(pkg):
    """Mcmc.run(batched="pipelined"): pf.partLogLikeBegin per chain, one pf.treesPartLogLike to collect."""
    pf = pkg.pf
    a = _mcmc(pkg, pf, 4, 9).run(50, batched="pipelined")
    b = _mcmc(pkg, pf, 4, 9).run(50, batched=False)
    _same_trace(a, b, 1e-12)


def test_begin_then_plain_part_loglike(pkg):
    pf = pkg.pf
    t = pkg.synth.build_config(pf, 2, nTax=10, nPatterns=300)
    want = t.calcLogLike()
    t.nodes[3].br.len *= 1.3
    t.nodes[3].br.lenChanged = True
    new = t.recalcAfterBranchChange()
    # the same proposal again, this time started asynchronously and collected by the ordinary call
    t.nodes[3].br.len /= 1.3
    t.setCStuff()
    pf.p4_calculateBigPDecks(t.nodes[3].cNode)
    q = t.nodes[3]
    while q.parent:
        q = q.parent
        pf.p4_setConditionalLikelihoodsOfInternalNodePart(q.cNode, 0)
    pf.partLogLikeBegin(t.cTree, 0)
    got = pf.p4_partLogLike(t.cTree, t.data.parts[0].cPart, 0, 0)
    assert rel(got, want) <= 1e-12 and new != want


def test_memoized_calls_equal_full_work(pkg, ref_pf):
    """p4b_setMemoize: after ONE composition of a composition-per-node model changes, p4's protocol recomputes the
    whole part (p4_setPrams + every node, p4/chain.py:305-380); the engine recognises which P decks and CLs have the
    inputs they were computed from and redoes only the others.  Same numbers, fewer steps."""
    pf, H = pkg.pf, pkg.host
    mine, twin = pkg.synth.build_config(pf, 4, nTax=10, nPatterns=200), None
    twin = H.clone_tree(mine, ref_pf)
    mine.calcLogLike()
    twin.calcLogLike()
    rng = np.random.default_rng(4)

    def whole_part(t, pNum):
        p = t.pf
        p.p4_setPrams(t.cTree, pNum)
        for n in t.iterInternalsPostOrder():
            p.p4_setConditionalLikelihoodsOfInternalNodePart(n.cNode, pNum)
        return p.p4_partLogLike(t.cTree, t.data.parts[pNum].cPart, pNum, 0)

    leaf = [n for n in mine.nodes if n.isLeaf][3]
    for trial, pNum in enumerate((0, 2, 0)):
        old = mine.model.parts[pNum].comps[leaf.nodeNum].val
        new = pkg.synth.normalise_comp(old * np.exp(rng.normal(0.0, 0.2, size=old.shape)))
        for t in (mine, twin):
            t.model.parts[pNum].comps[leaf.nodeNum].val[:] = new
        pf.treeSync(mine.cTree)
        n0 = pf.kernelLaunchCount()
        got = whole_part(mine, pNum)
        want = whole_part(twin, pNum)
        assert rel(got, want) <= LNL_TOL
        # and against this engine doing its full work on a fresh copy of the state
        pf.setMemoize(0)
        try:
            full = whole_part(mine, pNum)
        finally:
            pf.setMemoize(1)
        assert full == got
    # nothing changed at all: the calls are answered without recomputing any node but the root
    before = [pf.getNodeCL(mine.cTree, n.cNode, 0, 4, 20) for n in mine.iterInternalsPostOrder()]
    again = whole_part(mine, 0)
    assert again == got
    for b, n in zip(before, mine.iterInternalsPostOrder()):
        assert np.array_equal(b, pf.getNodeCL(mine.cTree, n.cNode, 0, 4, 20))
    # the whole-tree evaluation still recomputes everything and agrees
    assert rel(pf.p4_treeLogLike(mine.cTree, 0), float(sum(mine.partLikes))) <= 1e-12
