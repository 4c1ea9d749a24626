"""oracle/pf_trace.py -- TEST INFRASTRUCTURE ONLY.

Record, and replay, the traffic across the ``p4.pf`` boundary.

``Recorder`` wraps a ``pf`` module.  Installed as ``p4.pf`` under the reference's real Python package it sees every call
the reference's own code makes on the likelihood path -- ``Tree.calcLogLike``, ``Chain.proposeSp`` with the reference's
real proposals, ``Chain.gen``'s cur/prop transfer, ``Mcmc.run`` -- and writes it down: function name, arguments (handles
and NumPy buffers by symbolic id), return value.  The NumPy buffers the engine borrows (comp.val, gdasrv val/freqs/rates,
bQETneedsReset, preOrder/postOrder, partLikes, the var limits) are compared with their last known content around every
call: a change found BEFORE a call was made by Python and becomes an update event; a change found right AFTER a call was
made by the engine and becomes an expectation.

``replay`` feeds such a trace to another ``pf`` module -- this repository's engine -- call for call, applies the updates
to its own buffers, and checks every returned number and every engine-written buffer against the recording.  That is the
drop-in claim tested at the boundary itself: the reference's callers, unmodified, with their real call sequences.
"""
import gzip
import json

import numpy as np

# reference wrappers on the likelihood path (Pf/pfmodule.c method table)
RECORDED = """newData freeData pokePartInData newPart freePart pokeEquatesTable pokeSequences makePatterns
setGlobalInvarSitesVec partPatternCount getSiteLikes p4_newModel p4_freeModel p4_newModelPart p4_newComp p4_newRMatrix
p4_newGdasrv gdasrvCalcRates gdasrvCalcRates_np p4_setRMatrixBigR p4_setKappa p4_setPInvarVal p4_setRelRateVal
p4_resetBQET p4_getRelRate p4_newTree p4_freeTree p4_newNode p4_freeNode p4_setNodeRelation p4_setTreeRoot p4_setBrLen
p4_getTreeLen p4_setCompNum p4_setRMatrixNum p4_setGdasrvNum p4_setPrams p4_calculateBigPDecks
p4_calculateAllBigPDecksAllParts p4_setConditionalLikelihoodsOfInternalNodePart p4_partLogLike p4_treeLogLike
p4_copyCondLikes p4_copyBigPDecks p4_copyModelPrams p4_verifyIdentityOfTwoTrees
p4_newtSetup p4_newtAndBrentPowellOpt p4_newtAndBOBYQAOpt p4_getBrLens p4_getFreePrams
gsl_rng_get gsl_rng_set p4_simulate symbolSequences""".split()
MAKES_HANDLE = {"newData", "newPart", "p4_newModel", "p4_newTree", "p4_newNode", "p4_newGdasrv", "gsl_rng_get"}


class Recorder:
    """Module-like proxy: ``Recorder(ref_pf)`` behaves as ``ref_pf`` and logs the RECORDED calls in ``self.events``."""

    def __init__(self, target):
        self._t = target
        self.events = []
        self._handles = {}      # actual handle value -> symbolic id
        self._nHandles = 0      # ids are never reused, addresses are (free, then malloc)
        self._arrays = {}       # id(ndarray) -> (symbolic id, ndarray, last known content); holds the array, so id() stays unique
        self._nArrays = 0
        self.recording = True

    def __getattr__(self, name):
        f = getattr(self._t, name)
        if name not in RECORDED or not callable(f):
            return f

        def wrapped(*args):
            if not self.recording:
                return f(*args)
            self._scan("upd")
            enc = [self._enc(x) for x in args]
            ret = f(*args)
            if name in MAKES_HANDLE:
                hid = self._nHandles
                self._nHandles += 1
                self._handles[int(ret)] = hid
                r = {"h": hid}
            elif isinstance(ret, str):
                r = {"str": ret}
            elif isinstance(ret, (list, tuple)):
                r = [float(v) for v in ret]
            elif isinstance(ret, (int, float, np.integer, np.floating)):
                r = float(ret) if isinstance(ret, (float, np.floating)) else int(ret)
            else:
                r = None
            self.events.append(["call", name, enc, r])
            self._scan("cw")
            return ret
        return wrapped

    def _enc(self, x):
        if isinstance(x, np.ndarray):
            key = id(x)
            if key not in self._arrays:
                aid = self._nArrays
                self._nArrays += 1
                self._arrays[key] = [aid, x, x.copy()]
                self.events.append(["arr", aid, str(x.dtype), list(x.shape), x.ravel().tolist()])
            return {"a": self._arrays[key][0]}
        if isinstance(x, (bytes, bytearray)):
            return {"s": x.decode("latin-1")}
        if isinstance(x, str):
            return {"s": x}
        if isinstance(x, (bool, np.bool_)):
            return int(x)
        if isinstance(x, (int, np.integer)):
            v = int(x)
            return {"h": self._handles[v]} if v in self._handles and v > 4096 else v
        if isinstance(x, (float, np.floating)):
            return float(x)
        raise TypeError("pf_trace: cannot record argument %r" % (x,))

    def _scan(self, kind):
        for rec in self._arrays.values():
            aid, arr, last = rec
            if not np.array_equal(arr, last):
                self.events.append([kind, aid, arr.ravel().tolist()])
                rec[2] = arr.copy()

    def save(self, path, meta=None):
        with gzip.open(path, "wt") as fh:
            json.dump({"meta": meta or {}, "events": self.events}, fh)


def load(path):
    with gzip.open(path, "rt") as fh:
        return json.load(fh)


def replay(pf, trace, tol=1e-9):
    """Run the trace against ``pf``.  Returns a dict of statistics; raises AssertionError on the first mismatch."""
    handles, arrays = {}, {}
    worst, nCalls, nChecked, nWrites = 0.0, 0, 0, 0

    def close(got, want, what):
        nonlocal worst
        if want == got:
            return
        d = abs(got - want) / max(abs(want), 1e-300)
        if abs(want) < 1e-12:
            d = abs(got - want)
        worst = max(worst, d)
        assert d <= tol, "%s: got %r, recorded %r (rel. diff %.3e)" % (what, got, want, d)

    for k, ev in enumerate(trace["events"]):
        kind = ev[0]
        if kind == "arr":
            _, aid, dtype, shape, content = ev
            arrays[aid] = np.array(content, dtype=np.dtype(dtype)).reshape(shape)
        elif kind == "upd":
            _, aid, content = ev
            arrays[aid].ravel()[:] = content
        elif kind == "cw":
            _, aid, content = ev
            got = arrays[aid].ravel()
            nWrites += 1
            if got.dtype.kind in "iu":
                assert got.tolist() == content, "event %d: engine-written int buffer %d differs: %r vs recorded %r" % (k, aid, got.tolist(), content)
            else:
                for g, w in zip(got.tolist(), content):
                    close(g, w, "event %d: engine-written buffer %d" % (k, aid))
        else:
            _, name, enc, want = ev
            args = []
            for x in enc:
                if isinstance(x, dict):
                    args.append(handles[x["h"]] if "h" in x else arrays[x["a"]] if "a" in x else x["s"])
                else:
                    args.append(x)
            got = getattr(pf, name)(*args)
            nCalls += 1
            if isinstance(want, dict) and "str" in want:
                assert got == want["str"], "event %d: %s returned a different string" % (k, name)
                nChecked += 1
            elif isinstance(want, dict):
                handles[want["h"]] = got
            elif isinstance(want, list):
                assert len(got) == len(want), "event %d: %s returned %d values, recorded %d" % (k, name, len(got), len(want))
                for g, w in zip(got, want):
                    close(float(g), w, "event %d: %s" % (k, name))
                nChecked += len(want)
            elif isinstance(want, float):
                close(float(got), want, "event %d: %s" % (k, name))
                nChecked += 1
            elif isinstance(want, int) and name in ("partPatternCount", "p4_verifyIdentityOfTwoTrees"):
                assert int(got) == want, "event %d: %s returned %r, recorded %r" % (k, name, got, want)
                nChecked += 1
    return {"calls": nCalls, "checked_values": nChecked, "engine_written_buffers_checked": nWrites, "worst_rel_diff": worst}
