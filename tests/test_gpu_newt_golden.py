"""GPU: p4_newtAround on the engine against the committed golden answers (tests/golden/newt_around.json, frozen by
tests/golden/make_newt_golden.py from the reference's own p4_newtAround on the reference's example data): after each
call of the schedule, lnL within 1e-9 relative and every branch length within 1e-6 relative of the reference's.
Covers 4 states with and without pInvar and gamma, two parts, a composition per node, 6-state recoded protein
(the any-dim derivative kernel) and 20-state protein."""
import json
import os

import pytest

import golden_io
from util import rel

pytestmark = pytest.mark.gpu

sys_path_golden = os.path.join(golden_io.GOLDEN, "newt_around.json")
with open(sys_path_golden) as f:
    NEWT = json.load(f)["cases"]


def fixed(meta):
    for p in meta["parts"]:
        for k in ("comps", "rMatrices", "gdasrvs"):
            for m in p[k]:
                m["free"] = 0
        p["pInvarFree"] = 0
    meta["relRatesAreFree"] = 0
    return meta


def replay(pkg, pf, name, around):
    meta, _ = golden_io.load(name)
    tree = golden_io.build_tree(pkg, pf, fixed(meta))
    want = NEWT[name]
    try:
        assert rel(tree.calcLogLike(), want["lnL_start"]) <= 1e-9
        pf.p4_newtSetup(tree.cTree)
        for call in want["calls"]:
            around(tree.cTree, call["epsilon"], call["likeDelta"])
            assert rel(pf.p4_treeLogLike(tree.cTree, 0), call["lnL"]) <= 1e-9, (name, call["epsilon"])
            lens = pf.p4_getBrLens(tree.cTree)
            for n in tree.iterNodesNoRoot():
                w = call["brLens"][str(n.nodeNum)]
                assert abs(lens[n.nodeNum] - w) <= 1e-6 * max(w, 1e-3), (name, call["epsilon"], n.nodeNum, lens[n.nodeNum], w)
    finally:
        tree.deleteCStuff()
        tree.model.free()
        tree.data.free()


@pytest.mark.parametrize("name", sorted(NEWT))
def test_newt_around_matches_golden(pkg, name):
    pf = pkg.pf
    replay(pkg, pf, name, pf.newtAround)
