"""tests/dropin/p4_like_side.py -- run by tests/test_gpu_real_p4.py (and tools/bench_real_p4.py) in a process of its own.

The reference's REAL ``p4`` package -- its own Alignment / Data / Tree / Model / Mcmc / Chain code, unmodified (from
/root/reference/p4 in the build container, from the copy staged under oracle/_ref/p4 on the GPU box) -- with either
pf module installed as ``p4.pf``:

    ref    the reference's own Pf engine (oracle/_ref/pf*.so)
    mine   this repository's pf module: every likelihood call lands on the B200 engine through the C ABI

and runs, on the same synthetic alignment and tree,
    Tree.calcLogLike()                                   p4/tree.py:9406
    Tree.optLogLike(method="newtAndBrentPowell")         p4/tree.py:9417  (and "allBrentPowell" on request)
    Mcmc(t, nChains=4).run(n)                            p4/mcmc.py:2496  with p4's own proposals (p4/chain.py)
printing one RESULT line of JSON.  No pass-through: the GSL wrappers p4's proposals use are this module's own.

usage: p4_like_side.py ref|mine [--taxa N] [--patterns N] [--gens N] [--chains N] [--skip-opt] [--opt-method M] [--time]
"""
import argparse
import json
import os
import random
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

import ref_loader

ap = argparse.ArgumentParser()
ap.add_argument("which", choices=["ref", "mine"])
ap.add_argument("--taxa", type=int, default=12)
ap.add_argument("--patterns", type=int, default=600)
ap.add_argument("--gens", type=int, default=200)
ap.add_argument("--chains", type=int, default=4)
ap.add_argument("--skip-opt", action="store_true")
ap.add_argument("--opt-method", default="newtAndBrentPowell")
ap.add_argument("--seed", type=int, default=7)
ap.add_argument("--continue-gens", type=int, default=0,
                help="after the run, time a second Mcmc.run(N) of the same object: generations/s without chain construction")
args = ap.parse_args()

import p4_phylogenetics_b200 as P          # the synthetic-input generator lives in the package (numpy only)
if args.which == "ref":
    pfm = ref_loader.load_ref_pf()
else:
    pfm = P.pf
p4 = ref_loader.load_ref_p4(pf_module=pfm)
from p4 import Data, Mcmc, func, read, var

var.verboseRead = 0
var.warnReadNoFile = False

# ---- the same synthetic inputs for both engines: a random tree, a GTR+I+G4 alignment simulated down it (numpy) ----
rng = np.random.Generator(np.random.PCG64(20240 + args.seed))
htree = P.synth.random_tree(None, args.taxa, rng)
hmp = P.synth.dna_model_part(0, rng, 4, pInvar=0.2)
aln = P.synth.make_alignment(None, htree, hmp, args.patterns, rng, "dna")
names = ["t%03d" % i for i in range(args.taxa)]


def newick(n):
    if n.isLeaf:
        return "%s:%.10f" % (names[n.seqNum], n.br.len)
    inner = ",".join(newick(c) for c in n.iterChildren())
    return "(%s)" % inner if n.parent is None else "(%s):%.10f" % (inner, n.br.len)


tmp = tempfile.mkdtemp(prefix="p4like_")
cwd = os.getcwd()
os.chdir(tmp)                      # Mcmc writes its sample and log files into the working directory
with open("d.phy", "w") as f:
    f.write("%d %d\n" % (args.taxa, aln.length))
    for nm, s in zip(names, aln.sequences):
        f.write("%s  %s\n" % (nm, s.decode("latin-1") if isinstance(s, (bytes, bytearray)) else s))
with open("t.nwk", "w") as f:
    f.write(newick(htree.root) + ";\n")
read("d.phy")
d = Data()
read("t.nwk")
t = var.trees[0]
t.taxNames = d.taxNames
t.data = d
t.newComp(free=1, spec="empirical")
t.newRMatrix(free=1, spec="ones")
t.setNGammaCat(nGammaCat=4)
t.newGdasrv(free=1, val=0.5)
t.setPInvar(free=1, val=0.2)

out = {"which": args.which, "nPatterns": int(pfm.partPatternCount(d.parts[0].cPart)) if hasattr(pfm, "partPatternCount") else None}
t0 = time.perf_counter()
t.calcLogLike(verbose=0)
out["lnL0"] = float(t.logLike)
out["calc_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
for _ in range(3):
    t.calcLogLike(verbose=0)
out["calc_again_s"] = (time.perf_counter() - t0) / 3

if not args.skip_opt:
    t0 = time.perf_counter()
    t.optLogLike(verbose=0, method=args.opt_method)
    out["opt_s"] = time.perf_counter() - t0
    out["lnLopt"] = float(t.logLike)
    out["brLens"] = [float(n.br.len) for n in t.iterNodesNoRoot()]
    mp = t.model.parts[0]
    out["comp"] = [float(v) for v in mp.comps[0].val]
    out["rMatrix"] = [float(v) for v in np.asarray(mp.rMatrices[0].val).ravel()]
    out["shape"] = float(mp.gdasrvs[0].val[0])
    out["pInvar"] = float(mp.pInvar.val)

if args.gens > 0:
    random.seed(args.seed)
    var.gsl_rng = pfm.gsl_rng_get()
    pfm.gsl_rng_set(var.gsl_rng, args.seed)
    m = Mcmc(t, nChains=args.chains, runNum=0, sampleInterval=10, checkPointInterval=None, verbose=False)
    t0 = time.perf_counter()
    m.run(args.gens, verbose=False)
    out["mcmc_s"] = time.perf_counter() - t0
    out["gens_per_s"] = args.gens / out["mcmc_s"]
    likes = []
    with open("mcmc_likes_0") as f:
        for line in f:
            parts = line.split()
            if len(parts) == 2:
                likes.append(float(parts[1]))
    out["mcmc_likes"] = likes
    out["final_likes"] = [float(c.curTree.logLike) for c in m.chains]
    out["accepted"] = [[p.name, int(sum(p.nAcceptances)), int(sum(p.nProposals))] for p in m.props.proposals]
    if args.continue_gens > 0:
        # Mcmc.run builds its chains (two trees each: device state) on the first call; a second call continues the run
        try:
            t0 = time.perf_counter()
            m.run(args.continue_gens, verbose=False)
            out["mcmc_continued_s"] = time.perf_counter() - t0
            out["gens_per_s_continued"] = args.continue_gens / out["mcmc_continued_s"]
        except Exception as e:     # the first run's numbers above stand on their own
            out["mcmc_continued_error"] = repr(e)[:200]
os.chdir(cwd)
print("RESULT" + json.dumps(out))
