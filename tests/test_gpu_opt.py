"""GPU: Tree.optLogLike on the engine -- the reference's four optimiser entry points, native behind the C ABI
(csrc/opt.cpp, csrc/praxis.cpp), every objective evaluation on the GPU -- against the reference's own optimisers
(Pf/p4_treeOpt.c, Pf/brent.c) from the same start.  lnL at the optimum to 1e-6 relative, and EVERY optimised
parameter: the likelihood surface is flat at its top, so two optimisers that both stop when lnL moves by less than
1e-6 agree in a parameter only to about sqrt(1e-6 / curvature); the parameter tolerances below say that."""
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


def _free_everything(t):
    mp = t.model.parts[0]
    mp.comps[0].free = mp.rMatrices[0].free = mp.gdasrvs[0].free = 1
    mp.pInvar.free = 1
    t.model.nFreePrams = 3 + 5 + 1 + 1


def _compare_optima(pkg, ref_pf, mine, twin, got, want, lnlTol=1e-6):
    assert rel(got, want) <= lnlTol, (got, want)
    # every optimised parameter, not only lnL
    xm = pkg.pf.windUpParameters(mine.cTree, 1)[0]
    twinOnMine = pkg.host.clone_tree(twin, pkg.pf)       # the reference's optimum, wound up by this engine's packing
    twinOnMine.calcLogLike()
    xr = pkg.pf.windUpParameters(twinOnMine.cTree, 1)[0]
    assert len(xm) == len(xr)
    nModel = len(pkg.pf.windUpParameters(mine.cTree, 0)[0])
    for k, (a, b) in enumerate(zip(xm, xr)):
        tol = 2e-3 + 2e-2 * abs(b) if k < nModel else 1e-3 + 2e-2 * abs(b)
        assert abs(a - b) <= tol, "parameter %d: %g vs the reference's %g" % (k, a, b)
    # the optimised state is a consistent tree: a plain evaluation reproduces it in both engines
    assert rel(mine.calcLogLike(), got) <= 1e-10
    check = pkg.host.clone_tree(mine, ref_pf)
    assert rel(check.calcLogLike(), got) <= 1e-9
    twinOnMine.deleteCStuff()


def test_all_brent_powell_lands_on_the_reference_optimum(pkg, ref_pf):
    """p4_allBrentPowellOptimize: Brent's praxis, native, against the reference's praxis (Pf/brent.c) from the same start."""
    mine, twin = build_pair(pkg, ref_pf, 1, nTax=7, nPatterns=250)
    _free_everything(mine)
    _free_everything(twin)
    start = mine.calcLogLike()
    got = mine.optLogLike(method="allBrentPowell")
    want = twin.optLogLike(method="allBrentPowell")
    assert got > start + 1.0
    _compare_optima(pkg, ref_pf, mine, twin, got, want)


def test_bounded_optimiser_lands_on_the_reference_optimum(pkg, ref_pf):
    """p4_allBOBYQAOptimize: the bounded method (the oracle build has no nlopt, so the truth is the reference's praxis)."""
    mine, twin = build_pair(pkg, ref_pf, 1, nTax=7, nPatterns=250)
    _free_everything(mine)
    _free_everything(twin)
    start = mine.calcLogLike()
    got = mine.optLogLike(method="BOBYQA")
    want = twin.optLogLike(method="allBrentPowell")
    assert got > start + 1.0
    _compare_optima(pkg, ref_pf, mine, twin, got, want)
    x, lo, hi = pkg.pf.windUpParameters(mine.cTree, 1)
    assert np.all(x >= lo) and np.all(x <= hi)


def test_bounded_optimiser_respects_an_active_bound(pkg, ref_pf):
    """With GAMMA_SHAPE_MAX pulled below the unconstrained optimum the result sits ON the bound, and lnL is below the free one."""
    pf = pkg.pf
    mine = pkg.synth.build_config(pf, 1, nTax=7, nPatterns=250)
    _free_everything(mine)
    free = mine.optLogLike(method="BOBYQA")
    alphaFree = float(mine.model.parts[0].gdasrvs[0].val[0])
    other = pkg.synth.build_config(pf, 1, nTax=7, nPatterns=250)
    _free_everything(other)
    pkg.host.var._GAMMA_SHAPE_MAX[0] = 0.5 * alphaFree
    try:
        bound = other.optLogLike(method="BOBYQA")
        alpha = float(other.model.parts[0].gdasrvs[0].val[0])
    finally:
        pkg.host.var._GAMMA_SHAPE_MAX[0] = 300.0
    assert alpha <= 0.5 * alphaFree * (1 + 1e-12)
    assert abs(alpha - 0.5 * alphaFree) <= 1e-3 * alphaFree
    assert bound < free


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
    _compare_optima(pkg, ref_pf, mine, twin, got, want)


def test_newt_and_bounded_method(pkg, ref_pf):
    mine, twin = build_pair(pkg, ref_pf, 1, nTax=7, nPatterns=250)
    _free_everything(mine)
    _free_everything(twin)
    got = mine.optLogLike(method="newtAndBOBYQA")
    want = twin.optLogLike(method="newtAndBrentPowell")
    _compare_optima(pkg, ref_pf, mine, twin, got, want)


def test_newt_and_one_free_parameter(pkg, ref_pf):
    """One free model parameter: the reference's p4_newtAnd1DBrent branch (Pf/p4_treeOpt.c:1395-1436)."""
    mine, twin = build_pair(pkg, ref_pf, 2, nTax=8, nPatterns=300)
    for t in (mine, twin):
        t.model.parts[0].gdasrvs[0].free = 1
        t.model.nFreePrams = 1
    got = mine.optLogLike(method="newtAndBrentPowell")
    want = twin.optLogLike(method="newtAndBrentPowell")
    _compare_optima(pkg, ref_pf, mine, twin, got, want)
