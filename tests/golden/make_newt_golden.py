"""tests/golden/make_newt_golden.py -- freeze the reference's Newton-Raphson branch lengths.

Runs in the BUILD container only.  For every golden case already committed here (inputs frozen from the
reference's own example data by make_golden.py) the reference's Pf engine (oracle/_ref, Pf/*.c compiled
unmodified) runs its own p4_newtAround (Pf/p4_treeNewt.c:78-205, reached through ctypes: the pf module wraps
only the drivers) from the fixture's branch lengths, every model parameter held fixed:

    p4_newtAround(1.0, 10.0)  ->  lnL, branch lengths      (the first call of p4_newtAndBrentPowellOpt)
    p4_newtAround(1e-5, 1e-7) ->  lnL, branch lengths      (its last call)

and the results go to tests/golden/newt_around.json.  tests/test_gpu_newt_golden.py replays them on the GPU.

Usage: python tests/golden/make_newt_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import golden_io  # noqa: E402
import ref_loader  # noqa: E402
import ref_peek  # noqa: E402

SCHEDULE = [(1.0, 10.0), (1.0e-5, 1.0e-7)]


def fixed(meta):
    """The fixture with every model parameter held fixed (Newton-Raphson moves branch lengths only)."""
    for p in meta["parts"]:
        for k in ("comps", "rMatrices", "gdasrvs"):
            for m in p[k]:
                m["free"] = 0
        p["pInvarFree"] = 0
    meta["relRatesAreFree"] = 0
    return meta


def run(pkg, pf, meta, around):
    """[(lnL, {nodeNum: brLen})] after each call of the schedule, on engine ``pf``."""
    tree = golden_io.build_tree(pkg, pf, fixed(meta))
    start = tree.calcLogLike()
    pf.p4_newtSetup(tree.cTree)
    out = []
    for eps, delta in SCHEDULE:
        around(tree.cTree, eps, delta)
        lens = pf.p4_getBrLens(tree.cTree)
        out.append({"epsilon": eps, "likeDelta": delta, "lnL": pf.p4_treeLogLike(tree.cTree, 0),
                    "brLens": {str(n.nodeNum): float(lens[n.nodeNum]) for n in tree.iterNodesNoRoot()}})
    tree.deleteCStuff()
    tree.model.free()
    tree.data.free()
    return start, out


def main():
    import p4_phylogenetics_b200 as P
    rpf = ref_loader.load_ref_pf()
    lib = ref_peek.newt_lib()
    res = {}
    for name in golden_io.case_names():
        if name == "newt_around":
            continue
        meta, _ = golden_io.load(name)
        if "nodes" not in meta:
            continue
        start, out = run(P, rpf, meta, lambda t, e, d: lib.p4_newtAround(t, e, d))
        res[name] = {"lnL_start": start, "calls": out}
        print("%-28s %.6f -> %.6f -> %.6f" % (name, start, out[0]["lnL"], out[1]["lnL"]))
    with open(os.path.join(HERE, "newt_around.json"), "w") as f:
        json.dump({"_comment": "written by make_newt_golden.py from the reference's p4_newtAround (oracle/_ref)", "cases": res}, f, indent=0)


if __name__ == "__main__":
    main()
