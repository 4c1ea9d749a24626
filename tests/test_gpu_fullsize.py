"""BASELINE config 2 at its full size (200 taxa, 1,000,000 patterns, GTR+G4) on one GPU: properties that do not
need the reference to evaluate a million patterns.

  * site likelihoods are independent of the other columns: the reference's own engine, given a sample of the
    alignment's columns with the same tree and model, must reproduce the GPU's site likelihoods at those sites;
  * additivity over pattern shards: the two halves of the pattern range, evaluated as shards 0/2 and 1/2 of
    fresh copies of the data, sum to the unsharded log-likelihood;
  * the queued dirty path equals a full recompute.
"""
import numpy as np
import pytest

from util import rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big(pkg):
    tree = pkg.synth.build_config(pkg.pf, 2)
    tree.calcLogLike()
    yield tree
    tree.deleteCStuff()


def test_site_likelihoods_match_reference_on_a_column_sample(pkg, ref_pf, big):
    pf, H = pkg.pf, pkg.host
    aln = big.data.alignments[0]
    site = np.array(big.getSiteLikes())
    assert site.shape[0] == aln.length
    rng = np.random.default_rng(123)
    cols = np.sort(rng.choice(aln.length, size=3000, replace=False))
    seqs = [np.frombuffer(s if isinstance(s, (bytes, bytearray)) else s.encode(), dtype=np.uint8)[cols].tobytes() for s in aln.sequences]
    small = H.Alignment(ref_pf, seqs, aln.symbols, aln.equates)
    twin = big.dupe()
    twin.pf = ref_pf
    twin.data = H.Data(ref_pf, [small])
    twin.model = H.clone_model(big.model, ref_pf)
    want = np.array(twin.getSiteLikes())
    got = site[cols]
    assert np.max(np.abs(got - want) / want) <= 1e-9
    # and the total is the count-weighted sum of the site log-likelihoods
    assert rel(float(np.sum(np.log(site))), big.logLike) <= 1e-12


def test_pattern_shards_add_up(pkg, big):
    pf, H = pkg.pf, pkg.host
    total = big.logLike
    parts = []
    try:
        for rank in (0, 1):
            pf.setShard(rank, 2)
            t = H.clone_tree(big, pf)          # fresh data objects: a part's device mirror is tied to its shard
            parts.append(t.calcLogLike())
            lo, hi = pf.treeShardRange(t.cTree, 0)
            assert (lo, hi) == pf.shardRangeFor(pf.partPatternCount(t.data.parts[0].cPart), rank, 2)
            t.deleteCStuff()
            t.model.free()
            t.data.free()
    finally:
        pf.setShard(0, 1)
    assert rel(parts[0] + parts[1], total) <= 1e-12


def test_dirty_path_equals_full_recompute(pkg, big):
    rng = np.random.default_rng(9)
    for _ in range(3):
        i = int(rng.integers(1, len(big.nodes)))
        big.nodes[i].br.len = float(rng.uniform(0.001, 0.3))
        big.nodes[i].br.lenChanged = True
        a = big.recalcAfterBranchChange()
        b = big.calcLogLike()
        assert rel(a, b) <= 1e-12


def test_newton_step_at_full_size(pkg, big):
    """The Newton-Raphson quantities at 1 M patterns, by properties: the likelihood seen THROUGH any branch (cl2 on
    one side, the node's CL or tip on the other) is the tree's likelihood; the analytic first derivative agrees with a
    central difference of the full-tree evaluation; one Newton round over all 397 branches does not lower lnL and
    leaves a consistent tree."""
    pf = pkg.pf
    base = big.calcLogLike()
    pf.p4_newtSetup(big.cTree)
    nodes = list(big.iterNodesNoRoot())
    for n in (nodes[0], nodes[len(nodes) // 2], nodes[-1]):
        l0, d1, d2 = pf.newtDerivs(n.cNode)
        assert rel(l0, base) <= 1e-11
        v = n.br.len
        h = max(1e-3 * v, min(1e-5, 0.1 * v))
        n.br.len = v + h
        lp = big.calcLogLike()
        n.br.len = v - h
        lm = big.calcLogLike()
        n.br.len = v
        assert rel(big.calcLogLike(), base) <= 1e-13      # the engine is back at the base lengths for the next node
        # a central difference of two sums of 10^6 terms near -1.7e8: truncation O(h^2) plus rounding ~ ulp(lnL) / h
        assert abs((lp - lm) / (2 * h) - d1) <= 2e-3 * max(1.0, abs(d1)) + 64 * np.spacing(abs(base)) / h
        assert d2 == d2
    assert rel(big.calcLogLike(), base) <= 1e-13
    after = pf.newtAround(big.cTree, 1.0, 1.0e9)        # likeDelta so large that exactly one round runs
    assert after >= base - 1e-6
    lens = pf.p4_getBrLens(big.cTree)
    for n in nodes:
        n.br.len = lens[n.nodeNum]
    assert rel(big.calcLogLike(), after) <= 1e-12
