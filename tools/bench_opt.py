"""tools/bench_opt.py -- branch-length optimisation on the device.

One p4_newtAround(epsilon, likeDelta) (Newton-Raphson through cl2, Pf/p4_treeNewt.c) on a config-2 shaped
tree from perturbed branch lengths, beside (a) the same branches maximised one at a time by Brent's method
on the dirty-path objective (p4b_optimizeBrLens, one pass) and (b) the reference's own p4_newtAround on one
host core, on a bounded pattern sample of the same alignment, scaled by patterns.  Prints one JSON line.

Usage: python tools/bench_opt.py [--taxa 200] [--patterns 1000000] [--cpu-sample 2048] [--no-brent] [--no-cpu]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import p4_phylogenetics_b200 as P  # noqa: E402


def build(pf, taxa, patterns, sd, cfg=2):
    tree = P.synth.build_config(pf, cfg, nTax=taxa, nPatterns=patterns)
    rng = np.random.default_rng(7)
    start = {}
    for n in tree.iterNodesNoRoot():
        n.br.len = float(min(max(n.br.len * np.exp(rng.normal(0.0, sd)), 1e-4), 2.0))
        start[n.nodeNum] = n.br.len
    return tree, start


def reset(tree, start):
    for n in tree.iterNodesNoRoot():
        n.br.len = start[n.nodeNum]
    return tree.calcLogLike()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=2, help="2: DNA GTR+G4 (200 taxa x 1M patterns); 3: protein LG+G4 (100 x 200k)")
    ap.add_argument("--taxa", type=int, default=None)
    ap.add_argument("--patterns", type=int, default=None)
    ap.add_argument("--cpu-sample", type=int, default=2048)
    ap.add_argument("--sd", type=float, default=0.3, help="log-normal perturbation of the starting branch lengths")
    ap.add_argument("--epsilon", type=float, default=1.0e-5)
    ap.add_argument("--like-delta", type=float, default=1.0e-7)
    ap.add_argument("--no-brent", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    pf = P.pf
    a.taxa = a.taxa or (200 if a.cfg == 2 else 100)
    a.patterns = a.patterns or (1000000 if a.cfg == 2 else 200000)
    t0 = time.perf_counter()
    tree, start = build(pf, a.taxa, a.patterns, a.sd, a.cfg)
    lnL0 = tree.calcLogLike()
    setup = time.perf_counter() - t0
    nBranches = len(start)
    nPat = pf.partPatternCount(tree.data.parts[0].cPart)
    out = {"workload": "cfg%d shape: %d taxa, %d patterns, %s; branch lengths perturbed by exp(N(0, %.2f))" % (a.cfg, a.taxa, nPat, "GTR+G4" if a.cfg == 2 else "LG+G4", a.sd),
           "branches": nBranches, "lnL_start": lnL0, "setup_s": setup}

    pf.p4_newtSetup(tree.cTree)
    out["device_GB_with_cl2"] = pf.treeDeviceBytes(tree.cTree) / 1e9
    pf.newtAround(tree.cTree, a.epsilon, a.like_delta)          # warm-up (allocations, attribute calls)
    reset(tree, start)
    it0, k0 = pf.newtIterations(tree.cTree), pf.kernelLaunchCount()
    pf.treeTimerBegin(tree.cTree)
    w0 = time.perf_counter()
    lnL1 = pf.newtAround(tree.cTree, a.epsilon, a.like_delta)
    wall = time.perf_counter() - w0
    dev_ms = pf.treeTimerEnd(tree.cTree)
    iters = pf.newtIterations(tree.cTree) - it0
    out["newtAround"] = {"lnL": lnL1, "wall_s": wall, "device_ms": dev_ms, "derivative_evaluations": iters,
                         "per_branch_evaluations": iters / float(nBranches), "kernel_launches": pf.kernelLaunchCount() - k0,
                         "us_per_derivative_evaluation_wall": wall * 1e6 / max(iters, 1)}
    lens = pf.p4_getBrLens(tree.cTree)

    if not a.no_brent:
        reset(tree, start)
        w0 = time.perf_counter()
        lnL2, nEvals = pf.optimizeBrLens(tree.cTree, maxPasses=1, tol=1e-6)
        out["brent_dirty_path_one_pass"] = {"lnL": lnL2, "wall_s": time.perf_counter() - w0, "likelihood_evaluations": nEvals}

    if not a.no_cpu:
        import ref_loader
        import ref_peek
        if ref_loader.have_ref_pf():
            rpf = ref_loader.load_ref_pf()
            small, startS = build(rpf, a.taxa, min(a.cpu_sample, a.patterns), a.sd, a.cfg)
            small.calcLogLike()
            rpf.p4_newtSetup(small.cTree)
            w0 = time.perf_counter()
            ref_peek.newt_lib().p4_newtAround(small.cTree, a.epsilon, a.like_delta)
            t = time.perf_counter() - w0
            nS = rpf.partPatternCount(small.data.parts[0].cPart)
            out["reference_newtAround_1core"] = {"sample_patterns": nS, "wall_s_sample": t,
                                                 "wall_s_scaled_to_workload": t * nPat / float(nS),
                                                 "note": "a different (smaller) alignment of the same shape; scaled by patterns"}
            out["speedup_vs_reference_1core"] = (t * nPat / float(nS)) / wall
    out["brLen_mean_after"] = float(np.mean([lens[i] for i in start]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
