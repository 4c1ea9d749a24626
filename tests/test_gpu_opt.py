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
