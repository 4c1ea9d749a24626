"""GPU: the engine against the golden fixtures -- inputs and answers frozen from the
reference's own p4 package running its own example data (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import golden_io
from util import max_rel_err, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fused", [1, 0])
@pytest.mark.parametrize("name", golden_io.case_names())
def test_engine_matches_golden(pkg, name, fused):
    pf = pkg.pf
    meta, arr = golden_io.load(name)
    tree = golden_io.build_tree(pkg, pf, meta)
    pf.setFusedTreeKernel(fused)
    try:
        lnL = tree.calcLogLike()
        assert rel(lnL, meta["lnL"]) <= 1e-9
        for g, w in zip(tree.partLikes, meta["partLikes"]):
            assert rel(g, w) <= 1e-9
        for pNum, mp in enumerate(tree.model.parts):
            for n in tree.nodes:
                if n is not tree.root:
                    P1 = pf.getNodeBigP(n.cNode, pNum, mp.nGammaCat, mp.dim)
                    assert np.max(np.abs(P1 - arr["p%d_bigP_%d" % (pNum, n.nodeNum)])) < 1e-14
                if not n.isLeaf:
                    c1 = pf.getNodeCL(tree.cTree, n.cNode, pNum, mp.nGammaCat, mp.dim)
                    want = arr["p%d_cl_%d" % (pNum, n.nodeNum)]
                    scale = np.max(np.abs(want), axis=(0, 1), keepdims=True)
                    assert np.max(np.abs(c1 - want) / scale) < 1e-9
        site = np.array(tree.getSiteLikes())
        assert max_rel_err(site, np.array(meta["siteLikes"])) < 1e-11
    finally:
        pf.setFusedTreeKernel(1)
        tree.deleteCStuff()
        tree.model.free()
        tree.data.free()
