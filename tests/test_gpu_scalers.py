"""GPU: optional per-pattern log-scalers (p4b_setScalers).  Off by default -- the reference has none
and answers -1e99 on underflow.  On: identical results in range, finite lnL beyond it, checked against
the oracle port carried in 80-bit long double (which does not underflow)."""
import numpy as np
import pytest

import pf_port
from util import rel

pytestmark = pytest.mark.gpu


@pytest.fixture
def scalers(pkg):
    pkg.pf.setScalers(1)
    yield
    pkg.pf.setScalers(0)


def _deep_tree(P, nTax, seed, kind="dna", brlen=0.5):
    rng = np.random.Generator(np.random.PCG64(seed))
    tree = P.synth.random_tree(P.pf, nTax, rng)
    mp = P.synth.dna_model_part(0, rng, 4, pInvar=0.1) if kind == "dna" else P.synth.protein_model_part(0, rng)
    for n in tree.nodes:
        n.br.len = brlen
    aln = P.synth.make_alignment(P.pf, tree, mp, 96, rng, kind, gap_frac=0.01, ambig_frac=0.01, repeat=False)
    tree.attach(P.host.Data(P.pf, [aln]), P.host.Model(P.pf, [mp]))
    return tree


@pytest.mark.parametrize("fused", [1, 0])
def test_scalers_rescue_underflow_dna(pkg, scalers, fused):
    tree = _deep_tree(pkg, 1500, 3)
    pkg.pf.setFusedTreeKernel(fused)
    try:
        got = tree.calcLogLike()
    finally:
        pkg.pf.setFusedTreeKernel(1)
    want = pf_port.tree_loglike(tree, long_double=True)
    assert want > -1e98 and np.isfinite(got)
    assert rel(got, want) <= 1e-9


def test_scalers_rescue_underflow_protein(pkg, scalers):
    tree = _deep_tree(pkg, 400, 5, kind="protein")
    got = tree.calcLogLike()
    want = pf_port.tree_loglike(tree, long_double=True)
    assert pf_port.tree_loglike(tree) == -1.0e99          # plain double underflows here
    assert rel(got, want) <= 1e-9


def test_scalers_are_neutral_in_range(pkg, ref_pf):
    """Where nothing underflows, scalers on/off give the same log-likelihood and site likelihoods."""
    pf = pkg.pf
    off = pkg.synth.build_config(pf, 1, nTax=20, nPatterns=1500)
    a = off.calcLogLike()
    sa = np.array(off.getSiteLikes())
    pf.setScalers(1)
    try:
        on = pkg.host.clone_tree(off, pf, data=off.data)
        b = on.calcLogLike()
        sb = np.array(on.getSiteLikes())
    finally:
        pf.setScalers(0)
    assert a == b
    assert np.array_equal(sa, sb)
    twin = pkg.host.clone_tree(off, ref_pf)
    assert rel(b, twin.calcLogLike()) <= 1e-9


def test_scalers_dirty_path(pkg, scalers):
    tree = _deep_tree(pkg, 900, 8)
    full = tree.calcLogLike()
    tree.nodes[5].br.len = 0.3
    tree.nodes[5].br.lenChanged = True
    part = tree.recalcAfterBranchChange()
    assert rel(part, tree.calcLogLike()) <= 1e-12
    assert rel(part, pf_port.tree_loglike(tree, long_double=True)) <= 1e-9
    assert part != full
