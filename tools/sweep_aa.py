"""tools/sweep_aa.py -- one device-timed line for the 20-state whole-tree kernel on BASELINE config 3 (or 4).

The launch shape is chosen by environment variables read once per process (P4B_AA2_GROUPS, P4B_AA2_RING, P4B_AA2_KERNEL,
P4B_AA2_NOSTORE), so a sweep is one process per shape:  for g in 4 2 1; do P4B_AA2_GROUPS=$g python tools/sweep_aa.py; done
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import p4_phylogenetics_b200 as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=3)
    ap.add_argument("--taxa", type=int, default=None)
    ap.add_argument("--patterns", type=int, default=None)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--want", type=float, default=None, help="lnL expected (printed with the relative difference)")
    a = ap.parse_args()
    pf = P.pf
    pf.setMemoize(0)
    tree = P.synth.build_config(pf, a.cfg, nTax=a.taxa, nPatterns=a.patterns)
    lnL = tree.calcLogLike()
    for _ in range(3):
        pf.p4_treeLogLike(tree.cTree, 0)
    pf.treeTimerBegin(tree.cTree)
    for _ in range(a.steps):
        lnL = pf.p4_treeLogLike(tree.cTree, 0)
    ms = pf.treeTimerEnd(tree.cTree) / a.steps
    env = {k: v for k, v in os.environ.items() if k.startswith("P4B_AA2")}
    rec = {"cfg": a.cfg, "env": env, "kernel": pf.lastCLKernelName(), "ms": ms, "cl_ms": pf.treeLastCLTiming(tree.cTree)[0], "lnL": lnL}
    if a.want is not None:
        rec["rel"] = abs(lnL - a.want) / abs(a.want)
    print("SWEEPAA" + json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
