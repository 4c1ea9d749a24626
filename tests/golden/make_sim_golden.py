"""tests/golden/make_sim_golden.py -- freeze the reference's simulated sequences and root-state draws.

Runs in the BUILD container only.  For golden cases already committed here (inputs frozen from the reference's own
example data by make_golden.py) the reference's Pf engine (oracle/_ref) runs its own p4_simulate (Pf/p4_treeSim.c:14-420)
on its own mt19937 stream seeded with SEED, and p4_drawAncState (Pf/p4_treeSim.c:591-857) for the first sites after
srandom(SEED); the simulated sequences (as symbol strings, pf.symbolSequences) and the draws go to
tests/golden/simulate.json.  tests/test_zz_gpu_sim_golden.py replays them on the GPU.

Usage: python tests/golden/make_sim_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

import golden_io  # noqa: E402
import ref_loader  # noqa: E402

SEED = 20240
N_DRAWS = 40
CASES = ["a_simple_jc", "navidi_hky", "navidi_hky_i_g4", "navidi_gtr_g4", "navidi_two_parts", "navidi_hetero_ndch2",
         "grouped_aas_dayhoff6", "protein_lg_i_g4"]


def run(pkg, pf, meta):
    """(symbol sequences per part after Tree.simulate(seed=SEED), draws per part) on engine ``pf``."""
    for p in meta["parts"]:
        p["pInvarFree"] = 1 if p["pInvar"] else 0      # p4_drawAncState looks at the invariant share of a FREE pInvar only
    tree = golden_io.build_tree(pkg, pf, meta)
    tree.calcLogLike()
    draws = []
    pf.reseedCRandomizer(SEED)
    d = np.empty(4, dtype=np.int32)
    for pNum, dp in enumerate(tree.data.parts):
        rows = []
        for k in range(min(N_DRAWS, dp.nChar)):
            pf.p4_drawAncState(tree.cTree, pNum, k, d)
            rows.append([int(v) for v in d])
        draws.append(rows)
    tree.simulate(seed=SEED)
    seqs = [pf.symbolSequences(p.cPart) for p in tree.data.parts]
    nPat = [int(pf.partPatternCount(p.cPart)) for p in tree.data.parts]
    lnL = tree.calcLogLike()
    tree.deleteCStuff()
    tree.model.free()
    tree.data.free()
    return {"sequences": seqs, "nPatterns": nPat, "lnL_of_simulated_data": lnL, "drawAncState": draws}


def main():
    import p4_phylogenetics_b200 as P
    rpf = ref_loader.load_ref_pf()
    res = {}
    for name in CASES:
        meta, _ = golden_io.load(name)
        res[name] = run(P, rpf, meta)
        print("%-24s parts %d, patterns %s, lnL %.6f" % (name, len(res[name]["sequences"]), res[name]["nPatterns"], res[name]["lnL_of_simulated_data"]))
    with open(os.path.join(HERE, "simulate.json"), "w") as f:
        json.dump({"_comment": "written by make_sim_golden.py from the reference's p4_simulate / p4_drawAncState (oracle/_ref)",
                   "seed": SEED, "cases": res}, f, indent=0)


if __name__ == "__main__":
    main()
