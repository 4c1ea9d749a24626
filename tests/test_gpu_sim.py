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


# ---- p4_simulate -------------------------------------------------------------------------------------------------
import ref_peek  # noqa: E402


def _sequences(pf, tree, ref=False):
    out = []
    for p in tree.data.parts:
        A = ref_peek.part_arrays(p.cPart) if ref else pf.partArrays(p.cPart)
        out.append((np.array(A["sequences"]), int(A["nPatterns"]), np.array(A["patternCounts"]), np.array(A["globalInvarSitesVec"])))
    return out


@pytest.mark.parametrize("cfg,kw", [
    (1, dict(nTax=12, nPatterns=500)),          # DNA GTR+I+G4: categories, invariant sites (no draw at those)
    (2, dict(nTax=24, nPatterns=3000)),         # DNA GTR+G4
    (3, dict(nTax=10, nPatterns=400)),          # protein LG+G4
    (4, dict(nTax=8, nPatterns=200)),           # NDCH2, 4 parts: a composition per node, the stream runs over the parts
])
def test_simulate_gives_the_reference_sequences_for_the_same_seed(pkg, ref_pf, cfg, kw):
    """pf.p4_simulate on the device consumes the caller's mt19937 stream in the reference's order (Pf/p4_treeSim.c:235-360):
    with the same seed every simulated site of every taxon is the reference's, and so are the patterns made from them."""
    pf = pkg.pf
    mine, twin = build_pair(pkg, ref_pf, cfg, **kw)
    for seed in (1, 20240):
        mine.simulate(seed=seed)
        twin.simulate(seed=seed)
        for (s1, n1, c1, v1), (s0, n0, c0, v0) in zip(_sequences(pf, mine), _sequences(ref_pf, twin, ref=True)):
            assert s1.shape == s0.shape
            assert np.array_equal(s1, s0)
            assert n1 == n0 and np.array_equal(c1[:n1], c0[:n0])
            assert np.array_equal(v1[:n1], v0[:n0])
        # the tree re-lays its device state for the new patterns: the likelihood of the simulated data agrees too
        got, want = mine.calcLogLike(), twin.calcLogLike()
        assert abs(got - want) <= 1e-9 * abs(want)
    assert len(np.unique(_sequences(pf, mine)[0][0])) > 1


@pytest.mark.parametrize("cfg,kw,pInvarFree", [
    # pInvar > 0: the reference looks at the invariant share only when pInvar is FREE (Pf/p4_treeSim.c:780) and exits
    # with "gotIt is zero" when a fixed pInvar's share is hit -- this engine raises the same error, so only free here
    (1, dict(nTax=12, nPatterns=300), 1),
    (3, dict(nTax=8, nPatterns=150), 0),        # protein, no pInvar
    (3, dict(nTax=8, nPatterns=150), 1),
])
def test_draw_anc_state_matches_reference(pkg, ref_pf, cfg, kw, pInvarFree):
    """pf.p4_drawAncState: one draw per site from the root's posterior, from the C library's random() like the reference
    (pf.reseedCRandomizer seeds it): the same seed gives the same draws, site by site."""
    pf = pkg.pf
    mine, twin = build_pair(pkg, ref_pf, cfg, **kw)
    for t in (mine, twin):
        t.model.parts[0].pInvar.free = pInvarFree
    pf.reseedCRandomizer(77)
    got = mine.ancestralStateDraw()
    ref_pf.reseedCRandomizer(77)
    want = twin.ancestralStateDraw()
    assert len(got) == len(want) == mine.data.parts[0].nChar
    assert got == want
    d1, d0 = np.empty(4, np.int32), np.empty(4, np.int32)
    pf.reseedCRandomizer(5)
    a = []
    for k in range(50):
        pf.p4_drawAncState(mine.cTree, 0, k, d1)
        a.append(d1.tolist())
    ref_pf.reseedCRandomizer(5)
    b = []
    for k in range(50):
        ref_pf.p4_drawAncState(twin.cTree, 0, k, d0)
        b.append(d0.tolist())
    assert a == b


def test_simulate_with_a_ref_tree(pkg, ref_pf):
    """Tree.simulate(refTree=...): root states, categories and invariant flags from the posterior at refTree's root
    (C library stream), the walk down the tree on the mt19937 stream; same seeds, same sequences as the reference."""
    pf = pkg.pf
    mine, twin = build_pair(pkg, ref_pf, 1, nTax=10, nPatterns=300)
    refMine, refTwin = build_pair(pkg, ref_pf, 1, nTax=10, nPatterns=300)      # same tree + model on its own data objects
    for t in (mine, twin, refMine, refTwin):
        t.model.parts[0].pInvar.free = 1
    refMine.calcLogLike()
    refTwin.calcLogLike()
    pf.reseedCRandomizer(9)
    mine.simulate(seed=31, refTree=refMine)
    ref_pf.reseedCRandomizer(9)
    twin.simulate(seed=31, refTree=refTwin)
    (s1, n1, c1, v1), (s0, n0, c0, v0) = _sequences(pf, mine)[0], _sequences(ref_pf, twin, ref=True)[0]
    assert np.array_equal(s1, s0)
    assert n1 == n0 and np.array_equal(c1[:n1], c0[:n0])
    got, want = mine.calcLogLike(), twin.calcLogLike()
    assert abs(got - want) <= 1e-9 * abs(want)
