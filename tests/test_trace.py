"""pf-boundary traces of the reference's REAL Mcmc, of its REAL Tree.optLogLike (Newton-Raphson drivers) and of its REAL
Tree.simulate (tests/golden/make_trace.py, oracle/pf_trace.py).

CPU: the trace replays exactly on the reference's own engine (the harness is sound), and this repository's ``pf`` module
offers every function the reference's callers used.  GPU: the trace replays on the B200 engine -- every log-likelihood
the reference's Mcmc.run saw (its real proposals: local, eTBR, allBrLens, allCompsDir, allRMatricesDir, gdasrv, pInvar; chain
swaps; cur/prop transfer) and every buffer the engine writes, within 1e-9 relative; flag matrices and pattern counts exactly.
"""
import glob
import os

import pytest

import pf_trace

HERE = os.path.dirname(os.path.abspath(__file__))
TRACES = sorted(glob.glob(os.path.join(HERE, "golden", "trace_*.json.gz")))


@pytest.mark.parametrize("path", TRACES, ids=[os.path.basename(p)[6:-8] for p in TRACES])
def test_trace_replays_on_the_reference_engine(ref_pf, path):
    trace = pf_trace.load(path)
    stats = pf_trace.replay(ref_pf, trace, tol=1e-14)
    assert stats["calls"] == sum(trace["meta"]["calls"].values())
    base = os.path.basename(path)
    # an optimisation trace returns two vectors of branch lengths, a simulation trace three log-likelihoods and two sets of sequences
    floor = 20 if base.startswith("trace_opt_") else (4 if base.startswith("trace_sim_") else 100)
    assert stats["checked_values"] > floor and stats["worst_rel_diff"] <= 1e-14


@pytest.mark.parametrize("path", TRACES, ids=[os.path.basename(p)[6:-8] for p in TRACES])
def test_engine_offers_every_call_the_reference_made(pkg, path):
    trace = pf_trace.load(path)
    for name in trace["meta"]["calls"]:
        assert callable(getattr(pkg.pf, name)), name


# the simulation trace was recorded after the round's GPU minutes were spent: it replays last (see the end of the file)
SIM_TRACES = [p for p in TRACES if os.path.basename(p).startswith("trace_sim_")]
GPU_TRACES = [p for p in TRACES if p not in SIM_TRACES]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["default", "immediate"])
@pytest.mark.parametrize("path", GPU_TRACES, ids=[os.path.basename(p)[6:-8] for p in GPU_TRACES])
def test_trace_replays_on_the_gpu_engine(pkg, path, mode):
    pf = pkg.pf
    trace = pf_trace.load(path)
    if mode == "immediate":          # no queueing, no buffer sharing, no memoisation: every call does its work at once
        pf.setDeferredNodeCalls(0)
        pf.setSharedCondLikes(0)
        pf.setMemoize(0)
    try:
        # every log-likelihood, engine-written float and -- in the optimisation trace -- branch length within 1e-9
        # (observed: 1.2e-13 for the branch lengths after the reference's Newton-Raphson schedule)
        stats = pf_trace.replay(pf, trace, tol=1e-9)
    finally:
        pf.setDeferredNodeCalls(1)
        pf.setSharedCondLikes(1)
        pf.setMemoize(1)
    assert stats["calls"] == sum(trace["meta"]["calls"].values())
    print("replayed %(calls)d calls, checked %(checked_values)d returned values and %(engine_written_buffers_checked)d "
          "engine-written buffers, worst rel. diff %(worst_rel_diff).3e" % stats)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["default", "immediate"])
@pytest.mark.parametrize("path", SIM_TRACES, ids=[os.path.basename(p)[6:-8] for p in SIM_TRACES])
def test_simulation_trace_replays_on_the_gpu_engine(pkg, path, mode):
    """The reference's real Tree.simulate() + Tree.calcLogLike(): same sequences (pf.symbolSequences) and log-likelihoods."""
    test_trace_replays_on_the_gpu_engine(pkg, path, mode)
