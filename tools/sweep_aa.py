"""tools/sweep_aa.py -- one device-timed line for the 20-state whole-tree kernel on BASELINE config 3 (or 4).

The launch shape is chosen by environment variables read once per process (P4B_AA_SHAPE = "warps per CTA,ring depth,CTAs per SM",
P4B_AA_NOSTORE), so a sweep is one process per shape:  for s in 8,8,2 8,4,2 4,4,4 16,8,1; do P4B_AA_SHAPE=$s python tools/sweep_aa.py; done
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
    ap.add_argument("--sustain", type=float, default=0.0, help="seconds of back-to-back evaluations with SM clock / power sampling")
    ap.add_argument("--lean", action="store_true", help="lnL-only evaluations (p4b_setTreeStoresCL(0))")
    ap.add_argument("--want", type=float, default=None, help="lnL expected (printed with the relative difference)")
    a = ap.parse_args()
    pf = P.pf
    pf.setMemoize(0)
    if a.cfg == 61:      # the 61-state case of bench.py's codon61 block (generic tensor-core kernel, tree_dmma.cuh)
        tree = P.synth.build_generic(pf, P.synth.SYMBOLS_61, a.taxa or 32, a.patterns or 60000, 4, 6161, equates={"!": "abcd"})
    else:
        tree = P.synth.build_config(pf, a.cfg, nTax=a.taxa, nPatterns=a.patterns)
    lnL = tree.calcLogLike()
    if a.lean:
        pf.setTreeStoresCL(tree.cTree, 0)
    for _ in range(3):
        pf.p4_treeLogLike(tree.cTree, 0)
    pf.treeTimerBegin(tree.cTree)
    for _ in range(a.steps):
        lnL = pf.p4_treeLogLike(tree.cTree, 0)
    ms = pf.treeTimerEnd(tree.cTree) / a.steps
    clocks = None
    if a.sustain > 0:      # the same call back to back for a while, SM clock and power sampled through NVML every 5 ms
        import threading
        import time
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        sm, pw, stop = [], [], []

        def loop():
            while not stop:
                sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                time.sleep(0.005)
        th = threading.Thread(target=loop)
        n = max(1, int(a.sustain * 1000.0 / ms))
        th.start()
        pf.treeTimerBegin(tree.cTree)
        for _ in range(n):
            pf.p4_treeLogLike(tree.cTree, 0)
        ms2 = pf.treeTimerEnd(tree.cTree) / n
        stop.append(1)
        th.join()
        sm.sort()
        pw.sort()
        clocks = {"sustained_ms": ms2, "steps": n, "sm_mhz_median": sm[len(sm) // 2], "sm_mhz_min": sm[0], "power_w_median": pw[len(pw) // 2], "power_w_max": pw[-1],
                  "sm_max_mhz": pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM), "power_limit_w": pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0}
    env = {k: v for k, v in os.environ.items() if k.startswith("P4B_AA")}
    rec = {"cfg": a.cfg, "env": env, "kernel": pf.lastCLKernelName(), "ms": ms, "cl_ms": pf.treeLastCLTiming(tree.cTree)[0], "lnL": lnL}
    if clocks:
        rec["clocks"] = clocks
    if a.want is not None:
        rec["rel"] = abs(lnL - a.want) / abs(a.want)
    print("SWEEPAA" + json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
