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


if __name__ == "__main__":
    os.chdir(tempfile.mkdtemp())
    mcmc_simple(2, 120, 11, "trace_mcmc_gtr_i_g4.json.gz")
    mcmc_ndch2(2, 100, 12, "trace_mcmc_ndch2.json.gz")
