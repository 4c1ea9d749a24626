"""GPU parity of the P-deck consumers beside the likelihood: the composition a
tree-heterogeneous model expects at every tip, pf.p4_expectedComposition / p4_expectedCompositionCounts
(Pf/p4_treeSim.c:859-1045), against the reference engine on the same tree."""
import numpy as np
import pytest

from util import build_pair

pytestmark = pytest.mark.gpu


def _compare(mine, twin, pkg, ref_pf):
    got = pkg.pf.p4_expectedComposition(mine.cTree)
    want = ref_pf.p4_expectedComposition(twin.cTree)
    assert len(got) == len(want) == mine.model.nParts
    for g, w in zip(got, want):
        g, w = np.array(g), np.array(w)
        assert g.shape == w.shape
        assert np.max(np.abs(g - w)) < 1e-13
        assert np.allclose(g.sum(axis=1), 1.0, atol=1e-12)
    for pNum in range(mine.model.nParts):
        g = np.array(pkg.pf.p4_expectedCompositionCounts(mine.cTree, pNum))
        w = np.array(ref_pf.p4_expectedCompositionCounts(twin.cTree, pNum))
        assert g.shape == w.shape
        assert np.max(np.abs(g - w)) < 1e-9 * max(1.0, np.max(np.abs(w)))


@pytest.mark.parametrize("cfg,kw", [
    (1, dict(nTax=12, nPatterns=500)),          # homogeneous DNA with pInvar: the constant-site share
    (4, dict(nTax=8, nPatterns=200)),           # NDCH2: a composition per node, 4 parts
])
def test_expected_composition_matches_reference(pkg, ref_pf, cfg, kw):
    mine, twin = build_pair(pkg, ref_pf, cfg, **kw)
    mine.calcLogLike()
    twin.calcLogLike()
    _compare(mine, twin, pkg, ref_pf)


def test_expected_composition_follows_the_model(pkg, ref_pf):
    """Heterogeneous compositions: the expected tip compositions differ between tips, and move when a branch does."""
    mine, twin = build_pair(pkg, ref_pf, 4, nTax=8, nPatterns=200)
    rng = np.random.default_rng(3)
    for t in (mine, twin):
        r = np.random.default_rng(3)
        for mp in t.model.parts:
            for c in mp.comps:
                c.val[:] = pkg.synth.normalise_comp(r.dirichlet(2.0 * np.ones(len(c.val))))
    mine.calcLogLike()
    twin.calcLogLike()
    _compare(mine, twin, pkg, ref_pf)
    before = np.array(pkg.pf.p4_expectedComposition(mine.cTree)[0])
    assert np.max(np.abs(before[0] - before[1])) > 1e-3
    for t in (mine, twin):
        n = [x for x in t.iterNodesNoRoot() if x.isLeaf][0]
        n.br.len *= 4.0
    mine.calcLogLike()
    twin.calcLogLike()
    _compare(mine, twin, pkg, ref_pf)
    after = np.array(pkg.pf.p4_expectedComposition(mine.cTree)[0])
    assert np.max(np.abs(after - before)) > 1e-5
    del rng
