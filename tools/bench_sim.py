"""tools/bench_sim.py -- pf.p4_simulate on the device at the config-2 size: 200 taxa, 1,000,000
sites, GTR+G4, beside the reference's own p4_simulate on one host core on a bounded number of sites (scaled by sites).
The stream of uniforms is sequential by contract (the same seed must give the reference's sequences), so the host's
MT19937 is part of the timed call.  Prints one JSON line.

Usage: python tools/bench_sim.py [--taxa 200] [--sites 1000000] [--cpu-sites 20000] [--no-cpu]
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


def build(pf, taxa, sites):
    rng = np.random.Generator(np.random.PCG64(20242))
    tree = P.synth.random_tree(pf, taxa, rng)
    mp = P.synth.dna_model_part(0, rng, 4, pInvar=0.0)
    blank = b"a" * sites                       # one pattern; the simulation overwrites every site
    aln = P.host.Alignment(pf, [blank] * taxa, P.host.DNA_SYMBOLS, P.host.DNA_EQUATES)
    tree.attach(P.host.Data(pf, [aln]), P.host.Model(pf, [mp]))
    return tree


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--taxa", type=int, default=200)
    ap.add_argument("--sites", type=int, default=1000000)
    ap.add_argument("--cpu-sites", type=int, default=20000)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    pf = P.pf
    tree = build(pf, a.taxa, a.sites)
    tree.simulate(seed=3, calculatePatterns=False, resetSequences=False)         # warm-up: allocations
    k0 = pf.kernelLaunchCount()
    t0 = time.perf_counter()
    tree.simulate(seed=4, calculatePatterns=False, resetSequences=False)
    sim_s = time.perf_counter() - t0
    launches = pf.kernelLaunchCount() - k0
    t0 = time.perf_counter()
    for p in tree.data.parts:
        pf.makePatterns(p.cPart)
        pf.setGlobalInvarSitesVec(p.cPart)
    pat_s = time.perf_counter() - t0
    nPat = pf.partPatternCount(tree.data.parts[0].cPart)
    t0 = time.perf_counter()
    lnL = tree.calcLogLike()
    like_s = time.perf_counter() - t0
    out = {"workload": "p4_simulate: %d taxa x %d sites, GTR+G4 (%d draws of the mt19937 stream)" % (a.taxa, a.sites, a.sites * (2 * a.taxa - 1)),
           "simulate_s": sim_s, "site_states_per_s": a.sites * (2 * a.taxa - 2) / sim_s, "kernel_launches": launches,
           "makePatterns_s": pat_s, "patterns": nPat, "first_calcLogLike_on_the_new_data_s": like_s, "lnL": lnL}
    if not a.no_cpu:
        import ref_loader
        if ref_loader.have_ref_pf():
            rpf = ref_loader.load_ref_pf()
            small = build(rpf, a.taxa, a.cpu_sites)
            small.simulate(seed=3, calculatePatterns=False, resetSequences=False)
            t0 = time.perf_counter()
            small.simulate(seed=4, calculatePatterns=False, resetSequences=False)
            t = time.perf_counter() - t0
            out["reference_1core"] = {"sites": a.cpu_sites, "simulate_s": t, "scaled_to_workload_s": t * a.sites / a.cpu_sites}
            out["speedup_vs_reference_1core"] = (t * a.sites / a.cpu_sites) / sim_s
    print(json.dumps(out))


if __name__ == "__main__":
    main()
