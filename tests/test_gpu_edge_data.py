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


PROTEIN_EDGE = {
    "one column": ["a", "r", "n", "w"],
    "every column the same": ["llllll", "kkkkkk", "llllll", "vvvvvv", "aaaaaa"],
    "gap, missing and x columns between data": ["a-?xw-c", "r-?xw-c", "n-?xy-d", "d-?xy-e"],
    "ambiguity codes only": ["bzxbzx", "zbxzbx", "xxxxxx"],
    "constant except for gaps and ambiguities": ["dddddd", "d-dbxd", "dd?dbd", "ddddzd"],
}


@pytest.mark.parametrize("name", sorted(PROTEIN_EDGE))
@pytest.mark.parametrize("nCat,pInvar", [(4, 0.0), (4, 0.2), (1, 0.2), (3, 0.0)])
def test_protein_edge_alignments_match_reference(pkg, ref_pf, name, nCat, pInvar):
    """The same edges through the 20-state whole-tree kernel (FP64 tensor cores; a 256-pattern tile that is almost all padding)."""
    P, pf = pkg, pkg.pf
    seqs = PROTEIN_EDGE[name]
    rng = np.random.Generator(np.random.PCG64(100 + len(name)))
    tree = P.synth.random_tree(pf, len(seqs), rng)
    mp = P.synth.protein_model_part(0, rng, "lg", nCat)
    if pInvar:
        mp.pInvar = P.host.PInvar(pInvar)
    aln = P.host.Alignment(pf, seqs, P.host.PROTEIN_SYMBOLS, P.host.PROTEIN_EQUATES)
    tree.attach(P.host.Data(pf, [aln]), P.host.Model(pf, [mp]))
    twin = P.host.clone_tree(tree, ref_pf)
    got, want = tree.calcLogLike(), twin.calcLogLike()
    assert rel(got, want) <= 1e-9, (got, want)
    assert max_rel_err(tree.getSiteLikes(), twin.getSiteLikes()) <= 1e-9
    for t in (tree, twin):
        next(iter(t.iterNodesNoRoot())).br.len = 0.21
    assert rel(tree.calcLogLike(), twin.calcLogLike()) <= 1e-9
