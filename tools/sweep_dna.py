"""tools/sweep_dna.py -- launch shapes of the 4-state whole-tree kernels at the shard sizes of 1 / 2 / 4 / 8 GPUs.

For every pattern count: one tree (BASELINE configs[1] shape, 200 taxa), every launch shape of both kernel generations
(pf.setFusedVariant), device-timed p4_treeLogLike.  Prints one line per (size, shape) and a JSON summary.

Usage: python tools/sweep_dna.py [--sizes 1000000,250000,125000] [--variants 0,1,2,10,11,12,13,14,15,16] [--steps 30] [--lean]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import p4_phylogenetics_b200 as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1000000,250000,125000")
    ap.add_argument("--variants", default="0,1,2,10,11,12,13,14,15,16")
    ap.add_argument("--taxa", type=int, default=200)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--lean", action="store_true", help="also time lnL-only evaluations (p4b_setTreeStoresCL(0))")
    a = ap.parse_args()
    pf = P.pf
    pf.setMemoize(0)
    out = []
    for n in [int(x) for x in a.sizes.split(",")]:
        t0 = time.perf_counter()
        tree = P.synth.build_config(pf, 2, nTax=a.taxa, nPatterns=n)
        base = tree.calcLogLike()
        print("# %d patterns: setup %.1f s, lnL %.6f" % (n, time.perf_counter() - t0, base), flush=True)
        for v in [int(x) for x in a.variants.split(",")]:
            try:
                pf.setFusedVariant(v)
                modes = [(1, "store")] + ([(0, "lnl-only")] if a.lean else [])
                for storeCL, label in modes:
                    pf.setTreeStoresCL(tree.cTree, storeCL)
                    for _ in range(3):
                        lnL = pf.p4_treeLogLike(tree.cTree, 0)
                    pf.treeTimerBegin(tree.cTree)
                    for _ in range(a.steps):
                        lnL = pf.p4_treeLogLike(tree.cTree, 0)
                    ms = pf.treeTimerEnd(tree.cTree) / a.steps
                    cl = pf.treeLastCLTiming(tree.cTree)[0]
                    rec = {"patterns": n, "variant": v, "mode": label, "kernel": pf.lastCLKernelName(), "ms": ms, "cl_ms": cl, "lnL_rel": abs(lnL - base) / abs(base)}
                    out.append(rec)
                    print("%8d  v%-2d %-9s %-34s %8.4f ms  (kernel %8.4f)  rel %.1e" % (n, v, label, rec["kernel"], ms, cl, rec["lnL_rel"]), flush=True)
                pf.setTreeStoresCL(tree.cTree, 1)
            except SystemExit as e:
                print("%8d  v%-2d failed: %s" % (n, v, e), flush=True)
        pf.setFusedVariant(-1)
        tree.deleteCStuff()
        tree.model.free()
        tree.data.free()
    print("JSON" + json.dumps(out))


if __name__ == "__main__":
    main()
