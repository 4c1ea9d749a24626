"""tools/bench_cfg.py -- device-timed full-tree evaluations for any BASELINE config
(parity-test cases 1, 3, 4; config 2 is bench.py's headline).  Prints one JSON line.

Usage: python tools/bench_cfg.py --cfg 3 [--taxa N] [--patterns N] [--steps K]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import p4_phylogenetics_b200 as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=3)
    ap.add_argument("--taxa", type=int, default=None)
    ap.add_argument("--patterns", type=int, default=None)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--check", action="store_true", help="compare lnL with the reference engine on the same inputs (slow)")
    ap.add_argument("--dirty", action="store_true",
                    help="config 4's scenario (SURVEY.md 8d): evaluate after changing one leaf composition, one internal "
                         "composition, one branch length, with p4's call protocol; with and without p4b_setMemoize")
    a = ap.parse_args()
    pf = P.pf
    pf.setMemoize(0)      # every call does its full work
    t0 = time.perf_counter()
    tree = P.synth.build_config(pf, a.cfg, nTax=a.taxa, nPatterns=a.patterns)
    setup = time.perf_counter() - t0
    lnL = tree.calcLogLike()
    for _ in range(3):
        pf.p4_treeLogLike(tree.cTree, 0)
    pf.treeTimerBegin(tree.cTree)
    for _ in range(a.steps):
        pf.p4_treeLogLike(tree.cTree, 0)
    ms = pf.treeTimerEnd(tree.cTree) / a.steps
    e0 = time.perf_counter()
    for _ in range(a.steps):
        tree.calcLogLike()
    e2e = (time.perf_counter() - e0) * 1e3 / a.steps
    nInt = sum(1 for _ in tree.iterInternalsPostOrder())
    bpp = 0
    flops = 0
    for pNum, mp in enumerate(tree.model.parts):
        nPat = pf.partPatternCount(tree.data.parts[pNum].cPart)
        unit = 8 * mp.dim * mp.nGammaCat
        for n in tree.iterInternalsPostOrder():
            kids = list(n.iterChildren())
            k_int = sum(1 for c in kids if not c.isLeaf)
            bpp += (unit * (1 + k_int) + (len(kids) - k_int)) * nPat
            flops += 2 * mp.dim * mp.dim * mp.nGammaCat * k_int * nPat
    out = {"cfg": a.cfg, "taxa": len([n for n in tree.nodes if n.isLeaf]), "parts": tree.model.nParts,
           "patterns": [pf.partPatternCount(p.cPart) for p in tree.data.parts], "lnL": lnL,
           "ms_per_eval": ms, "evals_per_s": 1000.0 / ms, "e2e_ms_calcLogLike": e2e,
           "algorithmic_GB_per_eval": bpp / 1e9, "algorithmic_GBps": bpp / ms / 1e6,
           "GFLOP_per_eval": flops / 1e9, "TFLOPs": flops / ms / 1e9, "setup_s": setup,
           "device_GB": pf.treeDeviceBytes(tree.cTree) / 1e9}
    if a.dirty:
        import numpy as np
        rng = np.random.default_rng(0)

        def whole_part(pNum):     # what Chain.proposeSp issues after a composition proposal (p4/chain.py:305-380)
            pf.p4_setPrams(tree.cTree, pNum)
            for n in tree.iterInternalsPostOrder():
                pf.p4_setConditionalLikelihoodsOfInternalNodePart(n.cNode, pNum)
            return pf.p4_partLogLike(tree.cTree, tree.data.parts[pNum].cPart, pNum, 0)

        def comp_change(node, pNum):
            c = tree.model.parts[pNum].comps[node.parts[pNum].compNum]
            c.val[:] = P.synth.normalise_comp(c.val * np.exp(rng.normal(0.0, 0.05, size=c.val.shape)))

        leaves = [n for n in tree.nodes if n.isLeaf]
        internals = [n for n in tree.nodes if not n.isLeaf and n is not tree.root]
        res = {}
        for memo in (1, 0):
            pf.setMemoize(memo)
            tree.calcLogLike()
            for name, pool in (("one_leaf_comp", leaves), ("one_internal_comp", internals)):
                ts = []
                for k in range(a.steps):
                    comp_change(pool[k % len(pool)], k % tree.model.nParts)
                    t1 = time.perf_counter()
                    whole_part(k % tree.model.nParts)
                    ts.append((time.perf_counter() - t1) * 1e3)
                res["%s_ms_memo%d" % (name, memo)] = sum(ts) / len(ts)
            ts = []
            for k in range(a.steps):
                n = tree.nodes[1 + k % (len(tree.nodes) - 1)]
                n.br.len *= 1.05
                n.br.lenChanged = True
                t1 = time.perf_counter()
                tree.recalcAfterBranchChange()
                ts.append((time.perf_counter() - t1) * 1e3)
            res["one_brlen_ms_memo%d" % memo] = sum(ts) / len(ts)
        pf.setMemoize(0)
        out["dirty"] = res
    if a.check:
        import ref_loader
        twin = P.host.clone_tree(tree, ref_loader.load_ref_pf())
        t1 = time.perf_counter()
        want = twin.calcLogLike()
        out["reference_lnL"] = want
        out["reference_s_calcLogLike"] = time.perf_counter() - t1
        out["rel_diff"] = abs(lnL - want) / abs(want)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
