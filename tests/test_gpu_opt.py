"""GPU: Tree.optLogLike on the engine (bounded Powell over the reference's parameter vector, objective
on the GPU) against the reference's own self-contained optimiser (allBrentPowell, Pf/brent.c)."""
import numpy as np
import pytest

from util import build_pair, rel

pytestmark = pytest.mark.gpu


def test_parameter_vector_round_trip(pkg):
    pf = pkg.pf
    tree = pkg.synth.build_config(pf, 1, nTax=8, nPatterns=200)
    mp = tree.model.parts[0]
    mp.comps[0].free = mp.rMatrices[0].free = mp.gdasrvs[0].free = 1
    mp.pInvar.free = 1
    base = tree.calcLogLike()
    x, lo, hi = pf.windUpParameters(tree.cTree, 1)
    assert len(x) == 3 + 5 + 1 + 1 + (len(tree.nodes) - 1)
    assert np.all(x >= lo) and np.all(x <= hi)
    assert rel(pf.logLikeForParameters(tree.cTree, 1, x), base) <= 1e-12     # unwinding what was wound changes nothing
    x2 = x.copy()
    x2[-1] *= 1.5
    assert pf.logLikeForParameters(tree.cTree, 1, x2) != base


def test_optloglike_reaches_the_reference_optimum(pkg, ref_pf):
    mine, twin = build_pair(pkg, ref_pf, 1, nTax=7, nPatterns=250)
    for t in (mine, twin):
        mp = t.model.parts[0]
        mp.comps[0].free = mp.rMatrices[0].free = mp.gdasrvs[0].free = 1
        mp.pInvar.free = 1
        t.model.nFreePrams = 3 + 5 + 1 + 1
    start = mine.calcLogLike()
    got = mine.optLogLike(method="BOBYQA")
    want = twin.optLogLike(method="allBrentPowell")
    assert got > start + 1.0
    assert abs(got - want) < 0.05, (got, want)
    # the optimised state is a consistent tree: a plain evaluation reproduces it in both engines
    assert rel(mine.calcLogLike(), got) <= 1e-10
    check = pkg.host.clone_tree(mine, ref_pf)
    assert rel(check.calcLogLike(), got) <= 1e-9


def test_branch_lengths_through_the_dirty_path(pkg, ref_pf):
    """p4b_optimizeBrLens: every branch maximised by Brent's method on the dirty-path objective.  With no free
    model parameter the reference's newtAndBrentPowell (Newton-Raphson on the branch lengths) must land on the
    same optimum."""
    pf = pkg.pf
    mine, twin = build_pair(pkg, ref_pf, 2, nTax=9, nPatterns=400)
    rng = np.random.default_rng(2)
    for a, b in zip(mine.nodes, twin.nodes):
        a.br.len = b.br.len = float(a.br.len * np.exp(rng.normal(0.0, 0.8)))
    start = mine.calcLogLike()
    n0 = pf.kernelLaunchCount()
    got, nEvals = pf.optimizeBrLens(mine.cTree, maxPasses=60, tol=1e-9)
    launches = pf.kernelLaunchCount() - n0
    assert got > start
    assert launches <= 3 * nEvals + 8            # P(t) + step list (+ fold) per evaluation, not one launch per node
    brLens = pf.p4_getBrLens(mine.cTree)
    for n in mine.iterNodesNoRoot():
        n.br.len = brLens[n.nodeNum]
    assert rel(mine.calcLogLike(), got) <= 1e-10   # the state left behind is the optimum it reports
    want = twin.optLogLike(method="newtAndBrentPowell")
    assert abs(got - want) < 1e-3 * max(1.0, abs(want) * 1e-6) + 2e-3, (got, want)
    for a, b in zip(mine.iterNodesNoRoot(), twin.iterNodesNoRoot()):
        assert abs(a.br.len - b.br.len) < 5e-3 + 0.05 * b.br.len


def test_newt_and_brent_powell_method(pkg, ref_pf):
    mine, twin = build_pair(pkg, ref_pf, 1, nTax=7, nPatterns=250)
    for t in (mine, twin):
        mp = t.model.parts[0]
        mp.comps[0].free = mp.rMatrices[0].free = mp.gdasrvs[0].free = 1
        mp.pInvar.free = 1
        t.model.nFreePrams = 3 + 5 + 1 + 1
    got = mine.optLogLike(method="newtAndBrentPowell")
    want = twin.optLogLike(method="newtAndBrentPowell")
    assert abs(got - want) < 0.05, (got, want)
    assert rel(pkg.host.clone_tree(mine, ref_pf).calcLogLike(), got) <= 1e-9
