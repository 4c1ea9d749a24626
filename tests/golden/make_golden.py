"""tests/golden/make_golden.py -- regenerate the golden fixtures (build container only).

The reference ships no tests and records no expected log-likelihoods
(SURVEY.md section 4), so known answers have to be produced by running the
reference itself.  This script imports the reference's real Python package
``p4`` from /root/reference on top of its own Pf engine (oracle/_ref, built from
the unmodified Pf/*.c), runs the reference's OWN example inputs through its own
public API (``Tree.calcLogLike``, ``Tree.getSiteLikes``), and freezes inputs and
outputs as small JSON/NPZ files:

  inputs   sequences, symbols, equates, tree (relations, branch lengths, seqNums),
           model (comps, rMatrices, gdasrvs, pInvar, relRate, per-node usage)
  outputs  pattern arrays, lnL, partLikes, site likelihoods, every P deck,
           every conditional-likelihood array

Cases come from share/Examples/H_calcLike (A_simple, F_YangAndRoberts with the
Navidi SSU rRNA alignment, I_groupedAAs) plus one multi-part tree-heterogeneous
case assembled through the same p4 API.

Usage: python tests/golden/make_golden.py      (writes tests/golden/*.json, *.npz)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_loader  # noqa: E402
import ref_peek  # noqa: E402

p4 = ref_loader.load_ref_p4()
from p4 import Data, func, read, var  # noqa: E402

EX = os.path.join(ref_loader.REF_ROOT, "share", "Examples", "H_calcLike")
var.verboseRead = 0
var.warnReadNoFile = 0


def fresh():
    var.alignments = []
    var.trees = []
    var.sequenceLists = []
    var.nexusSets = None


def dump_case(name, t, note):
    """Freeze tree ``t`` (data + model attached, calcLogLike done) as a fixture."""
    t.calcLogLike(verbose=0)
    lnL = float(t.logLike)
    partLikes = [float(x) for x in t.partLikes]
    t.getSiteLikes()          # fills t.siteLikes (p4/tree.py:9679)
    siteLikes = [float(x) for x in t.siteLikes]
    assert abs(t.logLike - lnL) <= 1e-9 * abs(lnL)
    nodes = []
    for n in t.nodes:
        nodes.append({
            "nodeNum": n.nodeNum, "parent": n.parent.nodeNum if n.parent else -1,
            "leftChild": n.leftChild.nodeNum if n.leftChild else -1,
            "sibling": n.sibling.nodeNum if n.sibling else -1,
            "isLeaf": int(n.isLeaf), "seqNum": int(n.seqNum), "name": n.name,
            "brLen": float(n.br.len) if n.br else None,
            "compNum": [int(p.compNum) for p in n.parts],
            "rMatrixNum": [int(p.rMatrixNum) for p in n.br.parts] if n.br else [],
            "gdasrvNum": [int(p.gdasrvNum) for p in n.br.parts] if n.br else [],
        })
    parts, arrays = [], {}
    for pNum, (dp, mp) in enumerate(zip(t.data.parts, t.model.parts)):
        a = dp.alignment
        # the sequences of THIS part (a char partition subsets the alignment)
        seqs = pf_part_sequences(dp)
        parts.append({
            "symbols": dp.symbols, "equates": dp.equates, "dim": dp.dim, "nTax": dp.nTax, "nChar": dp.nChar,
            "sequences": seqs,
            "comps": [{"val": [float(x) for x in c.val], "free": int(c.free)} for c in mp.comps],
            "rMatrices": [{"spec": r.spec, "free": int(r.free), "val": None if r.val is None else [float(x) for x in np.atleast_1d(r.val)]}
                          for r in mp.rMatrices],
            "gdasrvs": [{"val": float(g.val[0]), "free": int(g.free), "nGammaCat": int(g.nGammaCat),
                         "rates": [float(x) for x in g.rates]} for g in mp.gdasrvs],
            "nGammaCat": int(mp.nGammaCat), "pInvar": float(mp.pInvar.val), "pInvarFree": int(mp.pInvar.free),
            "relRate": float(mp.relRate), "isHet": int(mp.isHet),
        })
        rp = ref_peek.part_arrays(dp.cPart)
        nPat = rp["nPatterns"]
        arrays["p%d_nPatterns" % pNum] = np.array(nPat)
        arrays["p%d_patterns" % pNum] = rp["patterns"][:, :nPat].astype(np.int8)
        arrays["p%d_patternCounts" % pNum] = rp["patternCounts"][:nPat]
        arrays["p%d_sequencePositionPatternIndex" % pNum] = rp["sequencePositionPatternIndex"]
        arrays["p%d_globalInvarSitesVec" % pNum] = rp["globalInvarSitesVec"][:nPat]
        arrays["p%d_globalInvarSitesArray" % pNum] = rp["globalInvarSitesArray"][:, :nPat].astype(np.int8)
        for n in t.nodes:
            if n is not t.root:
                arrays["p%d_bigP_%d" % (pNum, n.nodeNum)] = ref_peek.node_bigP(n.cNode, pNum, mp.nGammaCat, mp.dim)
            if not n.isLeaf:
                arrays["p%d_cl_%d" % (pNum, n.nodeNum)] = ref_peek.node_cl(n.cNode, pNum, mp.nGammaCat, mp.dim, rp["nChar"], nPat)
        # the normalised Q of the (comp, rMatrix) pair used by node 1 (unused pairs have no Q in the reference)
        n1 = t.nodes[1]
        cN, rN = int(n1.parts[pNum].compNum), int(n1.br.parts[pNum].rMatrixNum)
        Q = np.zeros((mp.dim, mp.dim))
        p4.pf.getBigQ(t.model.cModel, mp.dim, pNum, cN, rN, Q)
        arrays["p%d_bigQ_%d_%d" % (pNum, cN, rN)] = Q
    meta = {
        "name": name, "note": note, "root": t.root.nodeNum, "nodes": nodes, "parts": parts,
        "preOrder": [int(x) for x in t.preOrder], "postOrder": [int(x) for x in t.postOrder],
        "doRelRates": int(t.model.doRelRates), "relRatesAreFree": int(t.model.relRatesAreFree),
        "isHet": int(t.model.isHet),
        "lnL": lnL, "partLikes": partLikes, "siteLikes": siteLikes,
    }
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(meta, f)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print("%-28s lnL %.10f  parts %d  patterns %s" % (name, lnL, len(parts), [int(arrays["p%d_nPatterns" % i]) for i in range(len(parts))]))


def pf_part_sequences(dp):
    """The character strings the part was built from (what pf.pokeSequences received)."""
    rp = ref_peek.part_arrays(dp.cPart)
    eqSymb = "".join(sorted(dp.equates.keys())) if dp.equates else ""
    out = []
    for row in rp["sequences"]:
        chars = []
        for c in row:
            if c >= 0:
                chars.append(dp.symbols[c])
            elif c == -1:
                chars.append("-")
            elif c == -2:
                chars.append("?")
            else:
                chars.append(eqSymb[c + 64])
        out.append("".join(chars))
    return out


def case_a_simple():
    fresh()
    read(os.path.join(EX, "A_simple", "t.nex"))
    t = var.trees[0]
    read(os.path.join(EX, "A_simple", "d.nex"))
    t.data = Data()
    t.newComp(free=0, spec="equal")
    t.newRMatrix(free=0, spec="ones")
    t.setPInvar(free=0, val=0.0)
    t.setNGammaCat(nGammaCat=1)
    dump_case("a_simple_jc", t, "share/Examples/H_calcLike/A_simple/s.py: JC, 5 taxa x 200 sites")


def navidi(treeIdx=0):
    fresh()
    read(os.path.join(EX, "F_YangAndRoberts", "A_NavidiAlignFromYang", "navidiSSRNA.nex"))
    d = Data()
    read(os.path.join(EX, "F_YangAndRoberts", "C_homogAnalysis", "t.nex"))
    t = var.trees[treeIdx]
    t.data = d
    return t


def case_navidi_hky():
    t = navidi()
    t.newComp(free=1, spec="empirical")
    t.newRMatrix(free=1, spec="2p", val=2.0)
    t.setPInvar(free=0, val=0.0)
    dump_case("navidi_hky", t, "F_YangAndRoberts/C_homogAnalysis/s.py, first model: HKY kappa=2, empirical comp, no rate variation")


def case_navidi_hky_ig():
    t = navidi(1)
    t.newComp(free=1, spec="empirical")
    t.newRMatrix(free=1, spec="2p", val=3.7)
    t.setNGammaCat(nGammaCat=4)
    t.newGdasrv(free=1, val=0.5)
    t.setPInvar(free=1, val=0.15)
    dump_case("navidi_hky_i_g4", t, "F_YangAndRoberts/C_homogAnalysis/s.py, second model plus pInvar: HKY+I+G4")


def case_navidi_gtr_g():
    t = navidi(2)
    t.newComp(free=1, spec="specified", val=[0.19, 0.31, 0.27, 0.23])
    t.newRMatrix(free=1, spec="specified", val=[0.11, 0.33, 0.07, 0.12, 0.28, 0.09])
    t.setNGammaCat(nGammaCat=4)
    t.newGdasrv(free=1, val=1.3)
    t.setPInvar(free=0, val=0.0)
    dump_case("navidi_gtr_g4", t, "Navidi alignment, GTR+G4 with specified comp and rates")


def case_navidi_hetero():
    fresh()
    read(os.path.join(EX, "F_YangAndRoberts", "A_NavidiAlignFromYang", "navidiSSRNA.nex"))
    d = Data()
    read(os.path.join(EX, "F_YangAndRoberts", "D_Figure1", "t.nex"))
    t = var.trees[0]
    t.data = d
    rng = np.random.default_rng(4)
    nNodes = len(t.nodes)
    for i in range(nNodes):
        v = rng.dirichlet(20.0 * np.ones(4))
        v = v / v.sum()
        t.newComp(free=1, spec="specified", val=[float(x) for x in v])
    t.newRMatrix(free=1, spec="2p", val=2.0)
    t.newRMatrix(free=1, spec="2p", val=5.0)
    t.setNGammaCat(nGammaCat=4)
    t.newGdasrv(free=1, val=0.5)
    t.newGdasrv(free=1, val=1.5)
    t.setPInvar(free=0, val=0.0)
    mp = t.model.parts[0]
    for i, n in enumerate(t.nodes):
        t.setModelComponentOnNode(mp.comps[i], node=n, clade=0)
        if n is not t.root:
            t.setModelComponentOnNode(mp.rMatrices[i % 2], node=n, clade=0)
            t.setModelComponentOnNode(mp.gdasrvs[(i // 2) % 2], node=n, clade=0)
    dump_case("navidi_hetero_ndch2", t, "F_YangAndRoberts/D_Figure1 tree: a composition on every node (NDCH2 pattern), 2 rMatrices, 2 gdasrvs")


def case_grouped_aas():
    fresh()
    read(os.path.join(EX, "I_groupedAAs", "protein.nex"))
    a = var.alignments[0]
    a.recodeDayhoff()
    read("(((A:0.4, (B:0.4, C:0.4):0.05):0.05, D:0.4):0.1, E:0.4, F:0.4);")
    t = var.trees[0]
    t.data = Data()
    t.newComp(free=1, spec="empirical")
    t.newRMatrix(free=1, spec="ones")
    t.setNGammaCat(nGammaCat=1)
    t.setPInvar(free=0, val=0)
    dump_case("grouped_aas_dayhoff6", t, "share/Examples/H_calcLike/I_groupedAAs/s.py: Dayhoff-recoded protein, 6 states")


def case_protein_lg():
    fresh()
    read(os.path.join(EX, "I_groupedAAs", "protein.nex"))
    read("(((A:0.3, (B:0.2, C:0.1):0.05):0.05, D:0.4):0.1, E:0.25, F:0.15);")
    t = var.trees[0]
    t.data = Data()
    t.newComp(free=0, spec="lg")
    t.newRMatrix(free=0, spec="lg")
    t.setNGammaCat(nGammaCat=4)
    t.newGdasrv(free=1, val=0.7)
    t.setPInvar(free=1, val=0.05)
    dump_case("protein_lg_i_g4", t, "I_groupedAAs/protein.nex unrecoded: LG+I+G4, 20 states, with b/z/x ambiguities if present")


def case_two_parts():
    fresh()
    read(os.path.join(EX, "F_YangAndRoberts", "A_NavidiAlignFromYang", "navidiSSRNA.nex"))
    a = var.alignments[0]
    half = a.length // 2
    read("#nexus\nbegin sets;\n charset c1 = 1-%d;\n charset c2 = %d-.;\n charpartition cp1 = c1:c1, c2:c2;\nend;\n" % (half, half + 1))
    a.setCharPartition("cp1")
    d = Data()
    read(os.path.join(EX, "F_YangAndRoberts", "C_homogAnalysis", "t.nex"))
    t = var.trees[0]
    t.data = d
    t.newComp(partNum=0, free=1, spec="empirical")
    t.newRMatrix(partNum=0, free=1, spec="ones")
    t.setPInvar(partNum=0, free=1, val=0.1)
    t.setNGammaCat(partNum=0, nGammaCat=1)
    t.setRelRate(partNum=0, val=0.6)
    t.newComp(partNum=1, free=1, spec="empirical")
    t.newRMatrix(partNum=1, free=1, spec="2p", val=2.5)
    t.setPInvar(partNum=1, free=0, val=0.0)
    t.setNGammaCat(partNum=1, nGammaCat=4)
    t.newGdasrv(partNum=1, free=1, val=2.0)
    t.setRelRate(partNum=1, val=1.4)
    t.model.relRatesAreFree = 1
    dump_case("navidi_two_parts", t, "C_3_data_partitions/sOpt.py pattern: 2 char partitions with different models and relRates")


if __name__ == "__main__":
    case_a_simple()
    case_navidi_hky()
    case_navidi_hky_ig()
    case_navidi_gtr_g()
    case_navidi_hetero()
    case_grouped_aas()
    case_protein_lg()
    case_two_parts()
