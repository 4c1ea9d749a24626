"""tools/bench_mcmc.py -- BASELINE config 5: Mcmc.run with Metropolis-coupled chains, generations/sec.

100-taxon synthetic DNA, 500k patterns, GTR+G4 with every model parameter free, nChains chains, the
default proposal mix with the reference's weights (p4_phylogenetics_b200/mcmc.py).  A *generation* is
one proposal + accept/reject + cur/prop transfer on every chain, plus one swap attempt
(p4/mcmc.py:2830-2975).  Timed wall-clock around Mcmc.run(): host proposal code, the pf calls, kernels,
the device->host reads -- everything a user of Mcmc.run waits for.

  --engine b200       this repository's engine (default)
  --engine reference  the reference's own Pf engine (oracle/_ref) on one host core, on --ref-patterns
                      patterns; its generations/sec are scaled by patterns (every hot loop and every
                      copy is linear in patterns)

Prints one JSON line.
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
    ap.add_argument("--taxa", type=int, default=100)
    ap.add_argument("--patterns", type=int, default=500000)
    ap.add_argument("--chains", type=int, default=8)
    ap.add_argument("--gens", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--engine", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-patterns", type=int, default=2048)
    ap.add_argument("--mode", default="pipelined", choices=["pipelined", "batched", "sequential"],
                    help="pipelined: each chain's evaluation starts when its proposal is issued and is collected on its own event; "
                         "batched: one launch for all chains; sequential: the reference's order, one chain after the other")
    ap.add_argument("--no-defer", action="store_true", help="node-level calls launch at once")
    ap.add_argument("--no-bulk", action="store_true", help="Tree.setCStuff through the per-node calls")
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    if a.engine == "reference":
        import ref_loader
        pf = ref_loader.load_ref_pf()
        nPat = a.ref_patterns
    else:
        pf = P.pf
        nPat = a.patterns
        if a.no_defer:
            pf.setDeferredNodeCalls(0)
    t0 = time.perf_counter()
    tree = P.synth.build_config(pf, 5, nTax=a.taxa, nPatterns=nPat)
    tree.bulkSetCStuff = not a.no_bulk
    m = P.mcmc.Mcmc(tree, nChains=a.chains, seed=a.seed)
    setup = time.perf_counter() - t0
    batched = False
    if a.engine == "b200":
        batched = {"pipelined": "pipelined", "batched": True, "sequential": False}[a.mode]
    m.run(a.warmup, batched=batched)
    n0 = pf.kernelLaunchCount() if a.engine == "b200" else 0
    t0 = time.perf_counter()
    m.run(a.gens, batched=batched)
    if a.engine == "b200":
        pf.treeSync(m.chains[0].curTree.cTree)
    dt = time.perf_counter() - t0
    out = {"cfg": 5, "engine": a.engine, "taxa": a.taxa, "patterns": nPat, "chains": a.chains, "gens": a.gens,
           "gens_per_s": a.gens / dt, "ms_per_gen": 1e3 * dt / a.gens, "setup_s": setup,
           "mode": "sequential" if batched is False else ("batched" if batched is True else batched), "deferred": not a.no_defer, "bulk_setCStuff": not a.no_bulk,
           "proposals": {p.name: [p.nProposals, p.nAcceptances] for p in m.proposals},
           "swaps": [m.nSwapAttempts, m.nSwaps], "lnL_cold_chain": m.trace[-1][1][0]}
    if a.engine == "b200":
        out["launches_per_gen"] = (pf.kernelLaunchCount() - n0) / a.gens
        out["device_GB"] = sum(pf.treeDeviceBytes(t.cTree) for c in m.chains for t in (c.curTree, c.propTree)) / 1e9
    else:
        out["scaled_to_patterns"] = a.patterns
        out["gens_per_s_scaled"] = out["gens_per_s"] * nPat / a.patterns
        out["cores"] = 1
    print(json.dumps(out))


if __name__ == "__main__":
    main()
