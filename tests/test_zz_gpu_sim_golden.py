"""GPU: pf.p4_simulate and pf.p4_drawAncState on the engine against committed golden answers (tests/golden/simulate.json,
frozen by tests/golden/make_sim_golden.py from the reference's own p4_simulate / p4_drawAncState on the reference's example
cases): for the recorded seed every simulated symbol of every sequence, the pattern counts, the log-likelihood of the
simulated data (1e-9) and the root-state draws are the reference's.  Needs no reference engine at run time.
(The file sorts last on purpose: it was written after the round's GPU minutes were spent -- its logic was checked against the
reference engine on the CPU -- so under `pytest -x` a surprise here cannot hide the suites that were verified on a B200.)"""
import json
import os

import numpy as np
import pytest

import golden_io
from util import rel

pytestmark = pytest.mark.gpu

with open(os.path.join(golden_io.GOLDEN, "simulate.json")) as f:
    _G = json.load(f)
SEED, SIM = _G["seed"], _G["cases"]


def replay(pkg, pf, name):
    meta, _ = golden_io.load(name)
    for p in meta["parts"]:
        p["pInvarFree"] = 1 if p["pInvar"] else 0
    tree = golden_io.build_tree(pkg, pf, meta)
    want = SIM[name]
    try:
        tree.calcLogLike()
        pf.reseedCRandomizer(SEED)
        d = np.empty(4, dtype=np.int32)
        for pNum, dp in enumerate(tree.data.parts):
            for k, w in enumerate(want["drawAncState"][pNum]):
                pf.p4_drawAncState(tree.cTree, pNum, k, d)
                assert [int(v) for v in d] == w, (name, pNum, k)
        tree.simulate(seed=SEED)
        for pNum, p in enumerate(tree.data.parts):
            assert pf.symbolSequences(p.cPart) == want["sequences"][pNum], (name, pNum)
            assert int(pf.partPatternCount(p.cPart)) == want["nPatterns"][pNum]
        assert rel(tree.calcLogLike(), want["lnL_of_simulated_data"]) <= 1e-9
    finally:
        tree.deleteCStuff()
        tree.model.free()
        tree.data.free()


@pytest.mark.parametrize("name", sorted(SIM))
def test_simulate_and_draws_match_golden(pkg, name):
    replay(pkg, pkg.pf, name)
