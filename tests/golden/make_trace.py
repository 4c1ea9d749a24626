"""tests/golden/make_trace.py -- record pf-boundary traces of the reference's REAL Mcmc (build container only).

Imports the reference's Python package ``p4`` from /root/reference with ``oracle/pf_trace.Recorder`` installed as ``p4.pf``
(on top of the reference's own Pf engine, oracle/_ref), runs the reference's own MCMC example
(share/Examples/L_mcmc/A_simple/sMcmc.py: GTR+I+G on d.nex, Metropolis-coupled chains, the default proposal mix of
p4/mcmc.py -- local, eTBR, allCompsDir, allRMatricesDir, gdasrv, pInvar, allBrLens -- and chain swaps) and freezes every call across the boundary
with its arguments, borrowed-buffer updates and results.  tests/test_gpu_trace.py replays the trace on the GPU engine.

Usage: python tests/golden/make_trace.py      (writes tests/golden/trace_*.json.gz)
"""
import os
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pf_trace  # noqa: E402
import ref_loader  # noqa: E402

rec = pf_trace.Recorder(ref_loader.load_ref_pf())
p4 = ref_loader.load_ref_p4(pf_module=rec)
from p4 import Data, Mcmc, func, read, var  # noqa: E402

EX = os.path.join(ref_loader.REF_ROOT, "share", "Examples", "L_mcmc")
var.verboseRead = 0


def fresh():
    import gc
    rec.recording = False       # objects of an earlier run are freed here, outside any trace
    var.alignments = []
    var.trees = []
    var.sequenceLists = []
    var.nexusSets = None
    gc.collect()


def mcmc_simple(nChains, nGens, seed, name):
    fresh()
    rec.events = []
    rec._handles.clear()
    rec._arrays.clear()
    rec._nHandles = rec._nArrays = 0
    rec.recording = True
    random.seed(seed)
    read(os.path.join(EX, "d.nex"))
    d = Data()
    t = func.randomTree(taxNames=d.taxNames)
    t.data = d
    t.newComp(free=1, spec="empirical")
    t.newRMatrix(free=1, spec="ones")
    t.setNGammaCat(nGammaCat=4)
    t.newGdasrv(free=1, val=0.5)
    t.setPInvar(free=1, val=0.2)
    m = Mcmc(t, nChains=nChains, runNum=0, sampleInterval=10, checkPointInterval=None)
    m.run(nGens)
    rec.recording = False
    counts = {}
    for ev in rec.events:
        if ev[0] == "call":
            counts[ev[1]] = counts.get(ev[1], 0) + 1
    props = {p.name: [int(sum(p.nProposals)), int(sum(p.nAcceptances))] for p in m.props.proposals}
    meta = {"what": "reference p4 Mcmc.run(%d), %d chains, GTR+I+G4 on share/Examples/L_mcmc/d.nex" % (nGens, nChains),
            "seed": seed, "calls": counts, "proposals [made, accepted]": props,
            "final_cur_lnL": [float(c.curTree.logLike) for c in m.chains]}
    out = os.path.join(HERE, name)
    rec.save(out, meta)
    print(out, os.path.getsize(out), "bytes;", sum(counts.values()), "calls;", props)


def _begin(seed):
    fresh()
    rec.events = []
    rec._handles.clear()
    rec._arrays.clear()
    rec._nHandles = rec._nArrays = 0
    rec.recording = True
    random.seed(seed)


def _finish(m, what, seed, name):
    rec.recording = False
    counts = {}
    for ev in rec.events:
        if ev[0] == "call":
            counts[ev[1]] = counts.get(ev[1], 0) + 1
    props = {p.name: [int(sum(p.nProposals)), int(sum(p.nAcceptances))] for p in m.props.proposals}
    meta = {"what": what, "seed": seed, "calls": counts, "proposals [made, accepted]": props,
            "final_cur_lnL": [float(c.curTree.logLike) for c in m.chains]}
    out = os.path.join(HERE, name)
    rec.save(out, meta)
    print(out, os.path.getsize(out), "bytes;", sum(counts.values()), "calls;", props)


def mcmc_two_parts(nChains, nGens, seed, name):
    """Two character partitions with different models and free relative rates (share/Examples/H_calcLike/
    C_3_data_partitions pattern): per-part proposals, the relRate proposal, p4_calculateAllBigPDecksAllParts."""
    _begin(seed)
    read(os.path.join(EX, "d.nex"))
    a = var.alignments[0]
    half = a.length // 2
    read("#nexus\nbegin sets;\n charset c1 = 1-%d;\n charset c2 = %d-.;\n charpartition cp1 = c1:c1, c2:c2;\nend;\n" % (half, half + 1))
    a.setCharPartition("cp1")
    d = Data()
    t = func.randomTree(taxNames=d.taxNames)
    t.data = d
    t.newComp(partNum=0, free=1, spec="empirical")
    t.newRMatrix(partNum=0, free=1, spec="ones")
    t.setPInvar(partNum=0, free=1, val=0.1)
    t.setNGammaCat(partNum=0, nGammaCat=1)
    t.setRelRate(partNum=0, val=0.6)
    t.newComp(partNum=1, free=1, spec="empirical")
    t.newRMatrix(partNum=1, free=0, spec="2p", val=2.5)     # a FIXED 2-parameter matrix (kappa through p4_setKappa)
    t.setPInvar(partNum=1, free=0, val=0.0)
    t.setNGammaCat(partNum=1, nGammaCat=4)
    t.newGdasrv(partNum=1, free=1, val=2.0)
    t.setRelRate(partNum=1, val=1.4)
    t.model.relRatesAreFree = 1
    m = Mcmc(t, nChains=nChains, runNum=2, sampleInterval=10, checkPointInterval=None)
    m.run(nGens)
    _finish(m, "reference p4 Mcmc.run(%d), %d chains, 2 partitions (GTR+I / fixed 2-parameter matrix +G4), free relRates, on L_mcmc/d.nex" % (nGens, nChains), seed, name)


def mcmc_locations_polytomy(nGens, seed, name):
    """Two compositions and two rate matrices placed on the tree (NDCH + NDRH) with the location proposals, root3 and the
    brLen proposals switched on: model assignments move over the tree, the root changes.  (The reference's polytomy proposal does not
    run under Python 3.12 -- random.randrange(float), p4/chain.py:6393 -- so it is not in the trace.)"""
    _begin(seed)
    read(os.path.join(EX, "d.nex"))
    d = Data()
    t = func.randomTree(taxNames=d.taxNames)
    t.data = d
    c0 = t.newComp(free=1, spec="empirical")
    c1 = t.newComp(free=1, spec="empirical")
    r0 = t.newRMatrix(free=1, spec="ones")
    r1 = t.newRMatrix(free=1, spec="ones")
    t.setModelComponentsOnNodesRandomly()
    t.setNGammaCat(nGammaCat=4)
    t.newGdasrv(free=1, val=0.7)
    t.setPInvar(free=0, val=0.0)
    m = Mcmc(t, nChains=1, runNum=3, sampleInterval=10, checkPointInterval=None)
    m.prob.compLocation = 1.0
    m.prob.rMatrixLocation = 1.0
    m.prob.root3 = 1.0
    m.prob.brLen = 1.0
    m.run(nGens)
    _finish(m, "reference p4 Mcmc.run(%d), 1 chain, 2 comps + 2 rMatrices on the tree, compLocation / rMatrixLocation / root3 / "
            "brLen proposals on, GTR+G4 on L_mcmc/d.nex" % nGens, seed, name)


def mcmc_grouped_aa(nChains, nGens, seed, name):
    """share/Examples/L_mcmc/B_grouped_aa: protein recoded to the 6 Dayhoff groups (a 6-state 'standard' datatype)."""
    _begin(seed)
    read(os.path.join(EX, "B_grouped_aa", "protein.nex"))
    a = var.alignments[0]
    a.recodeDayhoff()
    d = Data()
    t = func.randomTree(taxNames=d.taxNames)
    t.data = d
    t.newComp(free=1, spec="empirical")
    t.newRMatrix(free=1, spec="ones")
    t.setNGammaCat(nGammaCat=4)
    t.newGdasrv(free=1, val=0.5)
    t.setPInvar(free=1, val=0.2)
    m = Mcmc(t, nChains=nChains, runNum=4, sampleInterval=10, checkPointInterval=None)
    m.run(nGens)
    _finish(m, "reference p4 Mcmc.run(%d), %d chains, Dayhoff-6 recoded protein (L_mcmc/B_grouped_aa), 6-state GTR+I+G4" % (nGens, nChains), seed, name)


def mcmc_protein(nChains, nGens, seed, name, nGammaCat):
    """Protein, LG exchangeabilities, free composition, gamma rates: with 4 categories the 20-state whole-tree
    tensor-core kernel serves both the whole-part recomputations and the dirty paths of the reference's proposals;
    with 1 category the per-node kernels do.  Ends with Tree.getSiteLikes on the cold chain's tree."""
    _begin(seed)
    read(os.path.join(EX, "B_grouped_aa", "protein.nex"))
    d = Data()
    t = func.randomTree(taxNames=d.taxNames)
    t.data = d
    t.newComp(free=1, spec="empirical")
    t.newRMatrix(free=0, spec="lg")
    t.setNGammaCat(nGammaCat=nGammaCat)
    if nGammaCat > 1:
        t.newGdasrv(free=1, val=0.5)
    t.setPInvar(free=1, val=0.1)
    m = Mcmc(t, nChains=nChains, runNum=5 + nGammaCat, sampleInterval=10, checkPointInterval=None)
    m.run(nGens)
    m.chains[0].curTree.getSiteLikes()
    _finish(m, "reference p4 Mcmc.run(%d), %d chains, protein LG+F+I%s (L_mcmc/B_grouped_aa/protein.nex), then Tree.getSiteLikes"
            % (nGens, nChains, "+G4" if nGammaCat > 1 else ""), seed, name)


def mcmc_ndch2(nChains, nGens, seed, name):
    """The composition-per-node model of share/Examples/W_recipes/sMcmcNDCH2.py on the same alignment: one free
    composition on every node (NDCH2 proposals change all leaf or all internal compositions; topology moves carry the
    compositions around; bQETneedsReset travels both ways)."""
    fresh()
    rec.events = []
    rec._handles.clear()
    rec._arrays.clear()
    rec._nHandles = rec._nArrays = 0
    rec.recording = True
    random.seed(seed)
    var.PIVEC_MIN = 1.e-6
    var.RATE_MIN = 1.e-6
    var.BRLEN_MIN = 1.e-5
    var.GAMMA_SHAPE_MIN = 0.15
    read(os.path.join(EX, "d.nex"))
    d = Data()
    t = func.randomTree(taxNames=d.taxNames)
    t.data = d
    for n in t.iterNodes():
        c = t.newComp(free=1, spec="empirical", symbol="-")
        t.setModelComponentOnNode(c, node=n, clade=0)
    t.newRMatrix(free=1, spec="ones")
    t.setNGammaCat(nGammaCat=4)
    t.newGdasrv(free=1, val=0.5)
    t.setPInvar(free=0, val=0.0)
    t.model.parts[0].ndch2 = True
    t.model.parts[0].ndch2_writeComps = False
    m = Mcmc(t, nChains=nChains, runNum=1, sampleInterval=10, checkPointInterval=None)
    m.run(nGens)
    rec.recording = False
    counts = {}
    for ev in rec.events:
        if ev[0] == "call":
            counts[ev[1]] = counts.get(ev[1], 0) + 1
    props = {p.name: [int(sum(p.nProposals)), int(sum(p.nAcceptances))] for p in m.props.proposals}
    meta = {"what": "reference p4 Mcmc.run(%d), %d chains, NDCH2 (a free composition per node) GTR+G4 on share/Examples/L_mcmc/d.nex" % (nGens, nChains),
            "seed": seed, "calls": counts, "proposals [made, accepted]": props,
            "final_cur_lnL": [float(c.curTree.logLike) for c in m.chains]}
    out = os.path.join(HERE, name)
    rec.save(out, meta)
    print(out, os.path.getsize(out), "bytes;", sum(counts.values()), "calls;", props)


def opt_newt(seed, name):
    """The reference's real Tree.optLogLike(method="newtAndBrentPowell") and (method="newtAndBOBYQA") (p4/tree.py:9417-9499)
    with every model parameter fixed: both drivers are then the four-call p4_newtAround schedule (Pf/p4_treeOpt.c:755-775,
    1214-1226), a deterministic function of the tree -- F81+I+G4 on L_mcmc/d.nex from a random tree with default branch lengths,
    then a second optimisation from doubled branch lengths."""
    _begin(seed)
    read(os.path.join(EX, "d.nex"))
    d = Data()
    t = func.randomTree(taxNames=d.taxNames)
    t.data = d
    t.newComp(free=0, spec="empirical")
    t.newRMatrix(free=0, spec="ones")
    t.setNGammaCat(nGammaCat=4)
    t.newGdasrv(free=0, val=0.5)
    t.setPInvar(free=0, val=0.2)
    t.calcLogLike(verbose=0)
    start = t.logLike
    t.optLogLike(verbose=0, method="newtAndBrentPowell")
    first = t.logLike
    for n in t.iterNodesNoRoot():
        n.br.len *= 2.0
    t.optLogLike(verbose=0, method="newtAndBOBYQA")
    rec.recording = False
    counts = {}
    for ev in rec.events:
        if ev[0] == "call":
            counts[ev[1]] = counts.get(ev[1], 0) + 1
    meta = {"what": "reference p4 Tree.optLogLike(newtAndBrentPowell) then (newtAndBOBYQA), no free model parameter, F81+I+G4 on L_mcmc/d.nex",
            "seed": seed, "calls": counts, "lnL": [float(start), float(first), float(t.logLike)],
            "brLens": [float(n.br.len) for n in t.iterNodesNoRoot()]}
    out = os.path.join(HERE, name)
    rec.save(out, meta)
    print(out, os.path.getsize(out), "bytes;", sum(counts.values()), "calls;", meta["lnL"])


def sim_tree(seed, name):
    """The reference's real Tree.simulate() (p4/tree.py:9527-9637): p4 makes its mt19937 stream (pf.gsl_rng_get / gsl_rng_set),
    simulates into the tree's own data (pf.p4_simulate), re-compresses (pf.makePatterns, pf.setGlobalInvarSitesVec), brings the
    sequences back (pf.symbolSequences) -- then evaluates the new data on the same tree.  GTR+I+G4 on L_mcmc/d.nex, twice."""
    _begin(seed)
    var.gsl_rng = None
    read(os.path.join(EX, "d.nex"))
    d = Data()
    t = func.randomTree(taxNames=d.taxNames)
    t.data = d
    t.newComp(free=0, spec="empirical")
    t.newRMatrix(free=0, spec="specified", val=[1.2, 3.1, 0.8, 0.9, 3.5, 1.0])
    t.setNGammaCat(nGammaCat=4)
    t.newGdasrv(free=0, val=0.6)
    t.setPInvar(free=0, val=0.15)
    lnLs = []
    t.calcLogLike(verbose=0)
    lnLs.append(float(t.logLike))
    var.gsl_rng = pf_trace_rng_seed(4242)
    for _ in range(2):
        t.simulate()
        t.calcLogLike(verbose=0)
        lnLs.append(float(t.logLike))
    rec.recording = False
    var.gsl_rng = None
    counts = {}
    for ev in rec.events:
        if ev[0] == "call":
            counts[ev[1]] = counts.get(ev[1], 0) + 1
    meta = {"what": "reference p4 Tree.calcLogLike(), then twice Tree.simulate() + Tree.calcLogLike() on the simulated data; GTR+I+G4, L_mcmc/d.nex",
            "seed": seed, "calls": counts, "lnL": lnLs}
    out = os.path.join(HERE, name)
    rec.save(out, meta)
    print(out, os.path.getsize(out), "bytes;", sum(counts.values()), "calls;", lnLs)


def pf_trace_rng_seed(seed):
    """What Tree.simulate does when var.gsl_rng is unset (p4/tree.py:9596-9598), with a fixed seed instead of the clock."""
    g = rec.gsl_rng_get()
    rec.gsl_rng_set(g, seed)
    return g


if __name__ == "__main__":
    os.chdir(tempfile.mkdtemp())
    if len(sys.argv) > 1 and sys.argv[1] == "opt":      # only the optimisation trace
        opt_newt(18, "trace_opt_newt.json.gz")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sim":      # only the simulation trace
        sim_tree(19, "trace_sim_gtr_i_g4.json.gz")
        sys.exit(0)
    mcmc_simple(2, 120, 11, "trace_mcmc_gtr_i_g4.json.gz")
    mcmc_ndch2(2, 100, 12, "trace_mcmc_ndch2.json.gz")
    mcmc_two_parts(2, 100, 13, "trace_mcmc_two_parts_relrate.json.gz")
    mcmc_locations_polytomy(150, 14, "trace_mcmc_locations_root3.json.gz")
    mcmc_grouped_aa(2, 80, 15, "trace_mcmc_grouped_aa.json.gz")
    mcmc_protein(2, 80, 16, "trace_mcmc_protein_lg_i_g4.json.gz", 4)
    mcmc_protein(1, 60, 17, "trace_mcmc_protein_lg_i.json.gz", 1)
    opt_newt(18, "trace_opt_newt.json.gz")
    sim_tree(19, "trace_sim_gtr_i_g4.json.gz")
