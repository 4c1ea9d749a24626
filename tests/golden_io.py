"""Load a golden fixture (tests/golden/*.json + *.npz, written by make_golden.py from
the reference's own p4 package) into host-side objects driven by a given ``pf``."""
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    # a case is a .json + .npz pair (newt_around.json holds the Newton-Raphson answers of all cases)
    return sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLDEN, "*.json"))
                  if os.path.exists(p[:-5] + ".npz"))


def load(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        meta = json.load(f)
    arrays = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return meta, arrays


def build_tree(pkg, pf, meta):
    """host.Tree with data and model attached, exactly as the fixture describes them."""
    host = pkg.host
    nodes = [host.Node(n["nodeNum"]) for n in meta["nodes"]]
    for h, n in zip(nodes, meta["nodes"]):
        h.isLeaf, h.seqNum = n["isLeaf"], n["seqNum"]
        h.parent = nodes[n["parent"]] if n["parent"] >= 0 else None
        h.leftChild = nodes[n["leftChild"]] if n["leftChild"] >= 0 else None
        h.sibling = nodes[n["sibling"]] if n["sibling"] >= 0 else None
        if n["brLen"] is not None:
            h.br.len = n["brLen"]
    tree = host.Tree(pf, nodes, nodes[meta["root"]])
    tree.preOrder[:] = meta["preOrder"]
    tree.postOrder[:] = meta["postOrder"]
    tree.preAndPostOrderAreValid = True
    alns, mps = [], []
    for pNum, p in enumerate(meta["parts"]):
        alns.append(host.Alignment(pf, p["sequences"], p["symbols"], p["equates"]))
        mp = host.ModelPart(pNum, p["dim"], p["nGammaCat"])
        mp.comps = [host.Comp(c["val"], c["free"]) for c in p["comps"]]
        mp.rMatrices = [host.RMatrix(r["spec"], r["val"], r["free"]) for r in p["rMatrices"]]
        mp.gdasrvs = [host.Gdasrv(g["nGammaCat"], g["val"], g["free"]) for g in p["gdasrvs"]]
        mp.pInvar = host.PInvar(p["pInvar"], p["pInvarFree"])
        mp.relRate = p["relRate"]
        mp.isHet = p["isHet"]
        mps.append(mp)
    data = host.Data(pf, alns)
    model = host.Model(pf, mps)
    model.doRelRates, model.relRatesAreFree = meta["doRelRates"], meta["relRatesAreFree"]
    tree.attach(data, model)
    for h, n in zip(nodes, meta["nodes"]):
        for pNum in range(len(mps)):
            h.parts[pNum].compNum = n["compNum"][pNum]
            if n["rMatrixNum"]:
                h.br.parts[pNum].rMatrixNum = n["rMatrixNum"][pNum]
                h.br.parts[pNum].gdasrvNum = n["gdasrvNum"][pNum]
    return tree
