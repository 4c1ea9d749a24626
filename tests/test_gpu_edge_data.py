"""GPU parity on hand-made alignments at the edges of the data path: one pattern, one column, columns of gaps / missing /
fully ambiguous characters, ambiguity codes only.  The whole path (pattern compression -> leaf tables -> whole-tree kernel ->
pInvar / constant-site terms -> lnL fold) against the reference engine (Pf/p4_tree.c:868-1378) on the same tree and model."""
import numpy as np
import pytest

from test_host import EDGE_ALIGNMENTS
from util import max_rel_err, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(k for k, v in EDGE_ALIGNMENTS.items() if len(v) >= 3))
@pytest.mark.parametrize("pInvar", [0.0, 0.25])
def test_edge_alignments_match_reference(pkg, ref_pf, name, pInvar):
    P, pf = pkg, pkg.pf
    seqs = EDGE_ALIGNMENTS[name]
    rng = np.random.Generator(np.random.PCG64(len(name)))
    tree = P.synth.random_tree(pf, len(seqs), rng)
    mp = P.synth.dna_model_part(0, rng, 4, pInvar=pInvar)
    aln = P.host.Alignment(pf, seqs, P.host.DNA_SYMBOLS, P.host.DNA_EQUATES)
    tree.attach(P.host.Data(pf, [aln]), P.host.Model(pf, [mp]))
    twin = P.host.clone_tree(tree, ref_pf)
    got, want = tree.calcLogLike(), twin.calcLogLike()
    assert rel(got, want) <= 1e-9, (got, want)
    assert max_rel_err(tree.getSiteLikes(), twin.getSiteLikes()) <= 1e-9
    # a second evaluation after a branch-length change goes down the dirty path of the same launch plan
    for t in (tree, twin):
        next(iter(t.iterNodesNoRoot())).br.len = 0.37       # calcLogLike's setCStuff passes every branch length down again
    assert rel(tree.calcLogLike(), twin.calcLogLike()) <= 1e-9
