"""pf -- drop-in mirror of the reference's ``p4.pf`` extension module for the
likelihood path, backed by the B200 engine (libp4b200.so, include/p4b200.h).

Every function has the name, argument order and meaning of the corresponding
wrapper in the reference's method table (Pf/pfmodule.c:2873-3020).  Handles are
plain Python ints (the reference returns C pointers as ``Py_BuildValue("l")``,
Pf/pfmodule.c:1405).  NumPy arrays whose buffers the reference borrows
(comp.val, gdasrv val/freqs/rates, bQETneedsReset, preOrder/postOrder,
partLikes, the var limits) are borrowed here too and are re-read on every
compute call; this module additionally pins a reference to each so a buffer
cannot be freed under the engine.

Only the hot-path subset exists.  Anything else raises
AttributeError naming the missing function.

There is no CPU fallback: compute calls fail loudly without a CUDA device.
"""
import ctypes as C
import os

import numpy as np

from . import _build

_HERE = os.path.dirname(os.path.abspath(__file__))


class P4bFatal(SystemExit):
    """An engine-level fatal error.

    The reference prints a message and calls exit(1) in these situations
    (e.g. Pf/p4_tree.c:436, 495; Pf/part.c:247).  Left uncaught this ends the
    process with status 1 as well; unlike exit() it can be caught by a test.
    """

    def __init__(self, message):
        super().__init__(1)
        self.message = message

    def __str__(self):
        return self.message


def _load():
    path = os.path.join(_HERE, "libp4b200.so")
    if not os.path.exists(path):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU fallback for this path." % path)
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


_lib = _load()
ACCEPTS_BYTES = True   # pokeSequences etc. take bytes as well as str (no 320 MB re-encode)
lib_path = os.path.join(_HERE, "libp4b200.so")

_vp, _i, _d, _cp = C.c_void_p, C.c_int, C.c_double, C.c_char_p
_ip, _dp = C.POINTER(C.c_int), C.POINTER(C.c_double)


def _sig(name, res, *args):
    f = getattr(_lib, name)
    f.restype = res
    f.argtypes = list(args)
    return f


_lastError = _sig("p4b_lastError", _cp)
_sig("p4b_version", _cp)
_sig("p4b_deviceCount", _i)
_sig("p4b_setDevice", _i, _i)
_sig("p4b_setShard", _i, _i, _i)
_sig("p4b_shardRangeFor", _i, _i, _i, _i, _ip, _ip)
_sig("p4b_commGetUniqueId", _i, C.c_char_p)
_sig("p4b_commInitRank", _i, C.c_char_p, _i, _i)
_sig("p4b_commDestroy", _i)
_sig("p4b_peerReduceState", _i)
_sig("p4b_kernelLaunchCount", C.c_longlong)
_sig("p4b_setFusedTreeKernel", None, _i)
_sig("p4b_setFusedVariant", _i, _i)
_sig("p4b_lastCLKernelName", _cp)
_sig("p4b_rngSize", C.c_long, _vp)
_sig("p4b_rngGetState", None, _vp, _vp)
_sig("p4b_rngSetState", None, _vp, _vp)
_sig("p4b_ranGamma", C.c_double, _vp, C.c_double, C.c_double)
_sig("p4b_ranDirichlet", None, _vp, _i, _vp, _vp)
_sig("p4b_ranDirichletLnPdf", C.c_double, _i, _vp, _vp)
_sig("p4b_sfLnGamma", C.c_double, C.c_double)
_sig("p4b_ranGammaPdf", C.c_double, C.c_double, C.c_double, C.c_double)
_sig("p4b_meanVariance", None, _vp, _i, _vp, _vp)
_OBJ = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_void_p)
_sig("p4b_allBrentPowellOptimize", C.c_long, _vp)
_sig("p4b_allBOBYQAOptimize", C.c_long, _vp, _i)
_sig("p4b_newtAndBrentPowellOpt", C.c_long, _vp)
_sig("p4b_newtAndBOBYQAOpt", C.c_long, _vp)
_sig("p4b_praxisMinimize", C.c_double, _i, _vp, C.c_double, C.c_double, _OBJ, _vp)
_sig("p4b_boundedMinimize", C.c_double, _i, _vp, _vp, _vp, C.c_double, C.c_double, C.c_long, _OBJ, _vp, _vp)
_sig("p4b_setTensorCoreKernel", None, _i)
_sig("p4b_setFusedTreeKernel20", None, _i)
_sig("p4b_setDeferredNodeCalls", None, _i)
_sig("p4b_setSharedCondLikes", None, _i)
_sig("p4b_setMemoize", None, _i)
_sig("p4b_treesPartLogLike", _i, _i, _vp, _i, _vp)
_sig("p4b_partLogLikeBegin", _i, _vp, _i)
_sig("p4b_setScalers", None, _i)
_sig("p4b_newData", _vp, _i, _i)
_sig("p4b_freeData", None, _vp)
_sig("p4b_pokePartInData", _i, _vp, _vp, _i)
_sig("p4b_newPart", _vp, _i, _i, _cp, _i, _cp, _i)
_sig("p4b_freePart", None, _vp)
_sig("p4b_pokeEquatesTable", _i, _vp, _cp)
_sig("p4b_pokeSequences", _i, _vp, _cp)
_sig("p4b_makePatterns", _i, _vp)
_sig("p4b_setGlobalInvarSitesVec", _i, _vp)
_sig("p4b_partPatternCount", _i, _vp)
_sig("p4b_getUnconstrainedLogLike", _i, _vp, _dp)
_sig("p4b_singleSequenceBaseCounts", _i, _vp, _i, _vp)
_sig("p4b_symbolSequences", _i, _vp, _vp)
_sig("p4b_partSequenceSitesCount", _i, _vp, _i)
_sig("p4b_pokePartTaxListAtIndex", _i, _vp, _i, _i)
_sig("p4b_partComposition", _i, _vp, _vp)
_sig("p4b_partMeanNCharsPerSite", _d, _vp)
_sig("p4b_partSimpleConstantSitesCount", _i, _vp)
_sig("p4b_partBigXSquared", _d, _vp)
_sig("p4b_getSiteLikes", _i, _vp, _vp, _i)
for _n in ("Sequences", "Patterns", "PatternCounts", "SequencePositionPatternIndex", "GlobalInvarSitesVec",
           "GlobalInvarSitesArray", "Equates"):
    _sig("p4b_part" + _n, _vp, _vp)
_sig("p4b_partNChar", _i, _vp)
_sig("p4b_partNTax", _i, _vp)
_sig("p4b_partDim", _i, _vp)
_sig("p4b_newModel", _vp, _i, _i, _i, _i, _i, *([_vp] * 15))
_sig("p4b_freeModel", None, _vp)
_sig("p4b_newModelPart", _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp)
_sig("p4b_newComp", _i, _vp, _i, _i, _i, _vp)
_sig("p4b_newRMatrix", _i, _vp, _i, _i, _i, _i)
_sig("p4b_newGdasrv", _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp)
_sig("p4b_gdasrvCalcRates", _i, _vp)
_sig("p4b_gdasrvCalcRates_np", _i, _i, _d, _vp, _vp)
_sig("p4b_setRMatrixBigR", _i, _vp, _i, _i, _i, _i, _d)
_sig("p4b_setKappa", _i, _vp, _i, _i, _d)
_sig("p4b_setPInvarVal", _i, _vp, _i, _d)
_sig("p4b_setRelRateVal", _i, _vp, _i, _d)
_sig("p4b_resetBQET", _i, _vp, _i, _i, _i)
_sig("p4b_getRelRate", _d, _vp, _i)
_sig("p4b_getBigQ", _i, _vp, _i, _i, _i, _vp)
_sig("p4b_getBigR", _i, _i, _vp)
_sig("p4b_getModelBigR", _i, _vp, _i, _i, _vp)
_sig("p4b_getEig", _i, _vp, _i, _i, _i, _vp, _vp, _vp)
_sig("p4b_newTree", _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp)
_sig("p4b_freeTree", None, _vp)
_sig("p4b_newNode", _vp, _i, _vp, _i, _i, _i)
_sig("p4b_freeNode", None, _vp)
_sig("p4b_setNodeRelation", _i, _vp, _i, _i)
_sig("p4b_setTreeRoot", _i, _vp, _vp)
_sig("p4b_setTreeCStuff", _i, _vp, _i, _vp, _vp, _vp, _vp, _i)
_sig("p4b_setBrLen", _i, _vp, _d)
_sig("p4b_setCompNum", _i, _vp, _i, _i)
_sig("p4b_setRMatrixNum", _i, _vp, _i, _i)
_sig("p4b_setGdasrvNum", _i, _vp, _i, _i)
_sig("p4b_getTreeLen", _d, _vp)
_sig("p4b_setPrams", _i, _vp, _i)
_sig("p4b_calculateBigPDecks", _i, _vp)
_sig("p4b_calculateAllBigPDecksAllParts", _i, _vp)
_sig("p4b_setConditionalLikelihoodsOfInternalNodePart", _i, _vp, _i)
_sig("p4b_partLogLike", _d, _vp, _vp, _i, _i)
_sig("p4b_treeLogLike", _d, _vp, _i)
_sig("p4b_countParameters", _i, _vp, _i)
_sig("p4b_windUpParameters", _i, _vp, _i, _vp, _vp, _vp)
_sig("p4b_unWindParameters", _i, _vp, _i, _vp)
_sig("p4b_logLikeForParameters", _d, _vp, _i, _vp)
_sig("p4b_getBrLens", _i, _vp, _vp)
_sig("p4b_optimizeBrLens", _d, _vp, _i, _d, C.POINTER(C.c_long))
_sig("p4b_treePassLimit", _i, _vp)
_sig("p4b_rngNew", _vp)
_sig("p4b_rngFree", None, _vp)
_sig("p4b_rngSet", None, _vp, C.c_ulong)
_sig("p4b_rngGet", C.c_ulong, _vp)
_sig("p4b_rngUniform", _d, _vp)
_sig("p4b_rngFillUniform", None, _vp, _vp, C.c_long)
_sig("p4b_simulate", _i, _vp, _vp, _vp)
_sig("p4b_drawAncState", _i, _vp, _i, _i, _vp)
_sig("p4b_bootstrapData", _i, _vp, _vp, _vp)
_sig("p4b_reseedCRandomizer", None, _i)
_sig("p4b_drawAncStateFromCL", _i, _vp, _i, _i, _d, _i, _vp, _vp, _vp)
_sig("p4b_expectedComposition", _i, _vp, _i, _vp)
_sig("p4b_expectedCompositionCounts", _i, _vp, _i, _vp)
_sig("p4b_newtSetup", _i, _vp)
_sig("p4b_newtAround", _d, _vp, _d, _d)
_sig("p4b_newtDerivs", _i, _vp, _vp)
_sig("p4b_getNodeCL2", _i, _vp, _i, _vp)
_sig("p4b_newtIterations", C.c_longlong, _vp)
_sig("p4b_treeNNodes", _i, _vp)
_sig("p4b_treeNLeaves", _i, _vp)
_sig("p4b_treeNParts", _i, _vp)
_sig("p4b_treePartDim", _i, _vp, _i)
_sig("p4b_copyCondLikes", _i, _vp, _vp, _i)
_sig("p4b_copyBigPDecks", _i, _vp, _vp, _i)
_sig("p4b_copyModelPrams", _i, _vp, _vp)
_sig("p4b_verifyIdentityOfTwoTrees", _i, _vp, _vp)
_sig("p4b_treeShardRange", _i, _vp, _i, _ip, _ip)
_sig("p4b_getNodeCL", _i, _vp, _i, _vp)
_sig("p4b_getNodeBigP", _i, _vp, _i, _vp)
_sig("p4b_setNodeBigP", _i, _vp, _i, _vp)
_sig("p4b_setTreeStoresCL", _i, _vp, _i)
_sig("p4b_treeSync", _i, _vp)
_sig("p4b_treeTimerBegin", _i, _vp)
_sig("p4b_treeTimerEnd", _d, _vp)
_sig("p4b_treeLastCLTiming", _i, _vp, _dp, _ip)
_sig("p4b_treeDeviceBytes", C.c_longlong, _vp)
_sig("p4b_flushL2", _i, _vp)

# object -> list of borrowed numpy buffers kept alive for it
_keep = {}


def _fatal():
    msg = (_lastError() or b"").decode("utf-8", "replace")
    print(msg)
    raise P4bFatal(msg)


def _ok(rc):
    if rc != 0:
        _fatal()


def _handle(h):
    if not h:
        _fatal()
    return int(h)


def _arr(a, dtype, what):
    """Raw pointer of a numpy array whose buffer the engine borrows."""
    if not isinstance(a, np.ndarray) or a.dtype != np.dtype(dtype) or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("%s must be a C-contiguous numpy array of %s" % (what, np.dtype(dtype)))
    return a.ctypes.data


def _bytes(s):
    return s if isinstance(s, (bytes, bytearray)) else s.encode("latin-1")


# ---- engine (additions; no counterpart in the reference) ---------------------
def version():
    return _lib.p4b_version().decode()


def deviceCount():
    return _lib.p4b_deviceCount()


def setDevice(device):
    _ok(_lib.p4b_setDevice(device))


def setShard(rank, world):
    _ok(_lib.p4b_setShard(rank, world))


def shardRangeFor(nPatterns, rank, world):
    lo, hi = C.c_int(), C.c_int()
    _ok(_lib.p4b_shardRangeFor(nPatterns, rank, world, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def commGetUniqueId():
    buf = C.create_string_buffer(128)
    _ok(_lib.p4b_commGetUniqueId(buf))
    return buf.raw


def commInitRank(uid, rank, world):
    assert len(uid) == 128
    _ok(_lib.p4b_commInitRank(uid, rank, world))


def commDestroy():
    _ok(_lib.p4b_commDestroy())


def peerReduceState():
    """1: shard sums are combined inside the kernel over NVLink peer mailboxes; -1: NCCL all-reduce; 0: undecided."""
    return int(_lib.p4b_peerReduceState())


def setFusedTreeKernel(on):
    _lib.p4b_setFusedTreeKernel(int(on))


def setFusedVariant(v):
    """Launch shape of the 4-state whole-tree kernel: -1 by shard size (default), 0/1/2 forced (include/p4b200.h)."""
    _ok(_lib.p4b_setFusedVariant(int(v)))


def lastCLKernelName():
    return (_lib.p4b_lastCLKernelName() or b"").decode()


def setDeferredNodeCalls(on):
    """0: node-level calls launch at once instead of queueing (see include/p4b200.h)."""
    _lib.p4b_setDeferredNodeCalls(int(on))


def setSharedCondLikes(on):
    """0: p4_copyCondLikes really copies instead of sharing buffers between twin trees (see include/p4b200.h)."""
    _lib.p4b_setSharedCondLikes(int(on))


def partLogLikeBegin(cTree, pNum):
    """Start p4_partLogLike(cTree, ., pNum, 0) on the GPU and return at once; collect with p4_partLogLike or treesPartLogLike."""
    _ok(_lib.p4b_partLogLikeBegin(cTree, int(pNum)))


def setMemoize(on):
    """0: node-level calls always do their full work (see include/p4b200.h p4b_setMemoize)."""
    _lib.p4b_setMemoize(int(on))


def treesPartLogLike(cTrees, pNum):
    """p4_partLogLike of several trees (sharing the data part) as one batched launch -> list of floats."""
    n = len(cTrees)
    arr = (C.c_void_p * n)(*cTrees)
    out = np.empty(n, dtype=np.float64)
    _ok(_lib.p4b_treesPartLogLike(n, arr, int(pNum), out.ctypes.data))
    return out.tolist()


def setFusedTreeKernel20(on):
    _lib.p4b_setFusedTreeKernel20(int(on))


def setTensorCoreKernel(on):
    _lib.p4b_setTensorCoreKernel(int(on))


def setScalers(on):
    """Per-pattern log-scalers for trees created afterwards (off by default; see include/p4b200.h)."""
    _lib.p4b_setScalers(int(on))


def kernelLaunchCount():
    return _lib.p4b_kernelLaunchCount()


# ---- data ---------------------------------------------------------------------
def newData(nTax, nParts):
    return _handle(_lib.p4b_newData(nTax, nParts))


def freeData(cData):
    _lib.p4b_freeData(cData)


def pokePartInData(cPart, cData, i):
    _ok(_lib.p4b_pokePartInData(cPart, cData, i))


def newPart(nTax, nChar, equateSymbols, nEquates, symbols, dim):
    return _handle(_lib.p4b_newPart(nTax, nChar, _bytes(equateSymbols), nEquates, _bytes(symbols), dim))


def freePart(cPart):
    _lib.p4b_freePart(cPart)


def pokeEquatesTable(cPart, theString):
    _ok(_lib.p4b_pokeEquatesTable(cPart, _bytes(theString)))


def pokeSequences(cPart, theString):
    b = _bytes(theString)
    need = _lib.p4b_partNTax(cPart) * _lib.p4b_partNChar(cPart)
    if len(b) < need:
        raise P4bFatal("pokeSequences: got %d characters, need nTax*nChar = %d" % (len(b), need))
    _ok(_lib.p4b_pokeSequences(cPart, b))


def makePatterns(cPart):
    _ok(_lib.p4b_makePatterns(cPart))


def setGlobalInvarSitesVec(cPart):
    _ok(_lib.p4b_setGlobalInvarSitesVec(cPart))


def partPatternCount(cPart):
    return _lib.p4b_partPatternCount(cPart)


def singleSequenceBaseCounts(cPart, seqNum):
    """pf.singleSequenceBaseCounts(cPart, seqNum) -> list of ints (Pf/pfmodule.c:219, Pf/part.c:556-602)."""
    out = np.zeros(_lib.p4b_partDim(cPart), dtype=np.int32)
    _ok(_lib.p4b_singleSequenceBaseCounts(cPart, int(seqNum), out.ctypes.data))
    return [int(v) for v in out]


def symbolSequences(cPart):
    """pf.symbolSequences(cPart) -> all sequences as one string of nTax*nChar symbols (Pf/pfmodule.c:233, Pf/part.c:604-680)."""
    n = _lib.p4b_partNTax(cPart) * _lib.p4b_partNChar(cPart)
    buf = C.create_string_buffer(n + 1)
    _ok(_lib.p4b_symbolSequences(cPart, C.cast(buf, C.c_void_p)))
    return buf.raw[:n].decode("latin-1")


def partSequenceSitesCount(cPart, seqNum):
    """pf.partSequenceSitesCount(cPart, seqNum): sites that are neither gap nor '?' (Pf/pfmodule.c:363, Pf/part.c:1068)."""
    v = _lib.p4b_partSequenceSitesCount(cPart, int(seqNum))
    if v < 0:
        _fatal()
    return v


def pokePartTaxListAtIndex(cPart, val, index):
    """pf.pokePartTaxListAtIndex(cPart, val, index) (Pf/pfmodule.c:326): select the sequences partComposition looks at."""
    _ok(_lib.p4b_pokePartTaxListAtIndex(cPart, int(val), int(index)))


def partComposition(cPart):
    """pf.partComposition(cPart) -> list of floats (Pf/pfmodule.c:350, Pf/part.c:850-1066)."""
    out = np.zeros(_lib.p4b_partDim(cPart), dtype=np.float64)
    _ok(_lib.p4b_partComposition(cPart, out.ctypes.data))
    return [float(v) for v in out]


def partMeanNCharsPerSite(cPart):
    """pf.partMeanNCharsPerSite(cPart) (Pf/pfmodule.c:261, Pf/part.c:1419)."""
    return _lib.p4b_partMeanNCharsPerSite(cPart)


def partSimpleConstantSitesCount(cPart):
    """pf.partSimpleConstantSitesCount(cPart) (Pf/pfmodule.c:275, Pf/part.c:1459)."""
    return _lib.p4b_partSimpleConstantSitesCount(cPart)


def partBigXSquared(cPart):
    """pf.partBigXSquared(cPart) (Pf/pfmodule.c:296, Pf/part.c:1490); -2.0 when the data hold gaps or ambiguities."""
    return _lib.p4b_partBigXSquared(cPart)


def getUnconstrainedLogLike(cPart):
    """pf.getUnconstrainedLogLike(cPart) (Pf/pfmodule.c:432, Pf/part.c:682-714)."""
    out = C.c_double(0.0)
    _ok(_lib.p4b_getUnconstrainedLogLike(cPart, C.byref(out)))
    return out.value


def getSiteLikes(cPart):
    n = _lib.p4b_partNChar(cPart)
    out = np.empty(n, dtype=np.float64)
    if _lib.p4b_getSiteLikes(cPart, out.ctypes.data, n) < 0:
        _fatal()
    return out.tolist()


def _part_view(cPart, name, shape):
    ptr = getattr(_lib, "p4b_part" + name)(cPart)
    if not ptr:
        return None
    n = int(np.prod(shape))
    return np.ctypeslib.as_array(C.cast(ptr, _ip), shape=(n,)).reshape(shape).copy()


def partArrays(cPart):
    """Copies of the part's host arrays (struct partStruct, Pf/pftypes.h:31-53) for parity checks."""
    nTax, nChar, dim = _lib.p4b_partNTax(cPart), _lib.p4b_partNChar(cPart), _lib.p4b_partDim(cPart)
    return {
        "nPatterns": _lib.p4b_partPatternCount(cPart),
        "sequences": _part_view(cPart, "Sequences", (nTax, nChar)),
        "patterns": _part_view(cPart, "Patterns", (nTax, nChar)),
        "patternCounts": _part_view(cPart, "PatternCounts", (nChar,)),
        "sequencePositionPatternIndex": _part_view(cPart, "SequencePositionPatternIndex", (nChar,)),
        "globalInvarSitesVec": _part_view(cPart, "GlobalInvarSitesVec", (nChar,)),
        "globalInvarSitesArray": _part_view(cPart, "GlobalInvarSitesArray", (dim, nChar)),
    }


# ---- model ----------------------------------------------------------------------
def p4_newModel(nParts, doRelRates, relRatesAreFree, nFreePrams, isHet, rMatrixNormalizeTo1,
                PINVAR_MIN, PINVAR_MAX, KAPPA_MIN, KAPPA_MAX, GAMMA_SHAPE_MIN, GAMMA_SHAPE_MAX,
                PIVEC_MIN, PIVEC_MAX, RATE_MIN, RATE_MAX, RELRATE_MIN, RELRATE_MAX, BRLEN_MIN, BRLEN_MAX):
    lims = [PINVAR_MIN, PINVAR_MAX, KAPPA_MIN, KAPPA_MAX, GAMMA_SHAPE_MIN, GAMMA_SHAPE_MAX, PIVEC_MIN, PIVEC_MAX,
            RATE_MIN, RATE_MAX, RELRATE_MIN, RELRATE_MAX, BRLEN_MIN, BRLEN_MAX]
    ptrs = [_arr(rMatrixNormalizeTo1, np.int32, "rMatrixNormalizeTo1")] + [_arr(a, np.float64, "limit") for a in lims]
    h = _handle(_lib.p4b_newModel(int(nParts), int(doRelRates), int(relRatesAreFree), int(nFreePrams), int(isHet), *ptrs))
    _keep[h] = [rMatrixNormalizeTo1] + lims
    return h


def p4_freeModel(cModel):
    _lib.p4b_freeModel(cModel)
    _keep.pop(cModel, None)


def p4_newModelPart(cModel, pNum, dim, nComps, nRMatrices, nGdasrvs, nCat, pInvarFree, bQETneedsReset):
    _ok(_lib.p4b_newModelPart(cModel, pNum, dim, nComps, nRMatrices, nGdasrvs, nCat, int(pInvarFree),
                              _arr(bQETneedsReset, np.int32, "bQETneedsReset")))
    _keep.setdefault(cModel, []).append(bQETneedsReset)


def p4_resetBQET(cModel, pNum, compNum, rMatrixNum):
    _ok(_lib.p4b_resetBQET(cModel, pNum, compNum, rMatrixNum))


def p4_newComp(cModel, pNum, mNum, free, val):
    _ok(_lib.p4b_newComp(cModel, pNum, mNum, int(free), _arr(val, np.float64, "comp.val")))
    _keep.setdefault(cModel, []).append(val)


def p4_newRMatrix(cModel, pNum, mNum, free, spec):
    _ok(_lib.p4b_newRMatrix(cModel, pNum, mNum, int(free), int(spec)))


def p4_newGdasrv(cModel, pNum, mNum, nCat, free, val, freqs, rates):
    h = _handle(_lib.p4b_newGdasrv(cModel, pNum, mNum, nCat, int(free), _arr(val, np.float64, "gdasrv.val"),
                                   _arr(freqs, np.float64, "gdasrv.freqs"), _arr(rates, np.float64, "gdasrv.rates")))
    _keep.setdefault(cModel, []).extend([val, freqs, rates])
    return h


def gdasrvCalcRates(cGdasrv):
    _ok(_lib.p4b_gdasrvCalcRates(cGdasrv))


def gdasrvCalcRates_np(nGammaCat, val, freqs, rates):
    _ok(_lib.p4b_gdasrvCalcRates_np(nGammaCat, float(val), _arr(freqs, np.float64, "freqs"), _arr(rates, np.float64, "rates")))


def p4_setRMatrixBigR(cModel, pNum, rNum, i, j, val):
    _ok(_lib.p4b_setRMatrixBigR(cModel, pNum, rNum, i, j, float(val)))


def p4_setKappa(cModel, pNum, rNum, val):
    _ok(_lib.p4b_setKappa(cModel, pNum, rNum, float(val)))


def p4_setPInvarVal(cModel, pNum, val):
    _ok(_lib.p4b_setPInvarVal(cModel, pNum, float(val)))


def p4_setRelRateVal(cModel, pNum, val):
    _ok(_lib.p4b_setRelRateVal(cModel, pNum, float(val)))


def p4_getRelRate(cModel, pNum):
    return _lib.p4b_getRelRate(cModel, pNum)


def getBigQ(cModel, dim, pNum, compNum, rMatrixNum, numpyBigQ):
    assert numpyBigQ.size >= dim * dim
    _ok(_lib.p4b_getBigQ(cModel, pNum, compNum, rMatrixNum, _arr(numpyBigQ, np.float64, "bigQ")))


def getBigR(spec, numpyBigR):
    assert numpyBigR.size >= 400
    _ok(_lib.p4b_getBigR(spec, _arr(numpyBigR, np.float64, "bigR")))


def getEig(cModel, dim, pNum, compNum, rMatrixNum):
    """(eigvecs, inverseEigvecs, eigvals) cached for one (comp, rMatrix) -- inspection only."""
    v, vi, lam = np.empty((dim, dim)), np.empty((dim, dim)), np.empty(dim)
    _ok(_lib.p4b_getEig(cModel, pNum, compNum, rMatrixNum, v.ctypes.data, vi.ctypes.data, lam.ctypes.data))
    return v, vi, lam


# ---- tree -------------------------------------------------------------------------
def p4_newTree(nNodes, nLeaves, preOrder, postOrder, newtAndBrentPowellOptPassLimit, partLikes, cData, cModel):
    h = _handle(_lib.p4b_newTree(nNodes, nLeaves, _arr(preOrder, np.int32, "preOrder"), _arr(postOrder, np.int32, "postOrder"),
                                 _arr(newtAndBrentPowellOptPassLimit, np.int32, "passLimit"),
                                 _arr(partLikes, np.float64, "partLikes"), cData, cModel))
    _keep[h] = [preOrder, postOrder, newtAndBrentPowellOptPassLimit, partLikes]
    return h


def p4_freeTree(cTree):
    _lib.p4b_freeTree(cTree)
    _keep.pop(cTree, None)


def p4_newNode(nodeNum, cTree, seqNum, isLeaf, inTree):
    return _handle(_lib.p4b_newNode(nodeNum, cTree, seqNum, int(isLeaf), int(inTree)))


def p4_freeNode(cNode):
    _lib.p4b_freeNode(cNode)


def p4_setNodeRelation(cNode, relation, relNum):
    _ok(_lib.p4b_setNodeRelation(cNode, relation, relNum))


def setTreeCStuff(cTree, parent, leftChild, sibling, brLen, rootNum):
    """Addition: the relations and branch lengths of every node in one call (int32 / float64 arrays by node number)."""
    _ok(_lib.p4b_setTreeCStuff(cTree, len(parent), _arr(parent, np.int32, "parent"), _arr(leftChild, np.int32, "leftChild"),
                               _arr(sibling, np.int32, "sibling"), _arr(brLen, np.float64, "brLen"), int(rootNum)))


def p4_setTreeRoot(cTree, cNode):
    _ok(_lib.p4b_setTreeRoot(cTree, cNode))


def p4_setBrLen(cNode, brLen):
    _ok(_lib.p4b_setBrLen(cNode, float(brLen)))


def p4_getTreeLen(cTree):
    return _lib.p4b_getTreeLen(cTree)


def p4_setCompNum(cNode, pNum, val):
    _ok(_lib.p4b_setCompNum(cNode, pNum, val))


def p4_setRMatrixNum(cNode, pNum, val):
    _ok(_lib.p4b_setRMatrixNum(cNode, pNum, val))


def p4_setGdasrvNum(cNode, pNum, val):
    _ok(_lib.p4b_setGdasrvNum(cNode, pNum, val))


# ---- the hot path -------------------------------------------------------------------
def p4_setPrams(cTree, pNum):
    _ok(_lib.p4b_setPrams(cTree, pNum))


def p4_calculateBigPDecks(cNode):
    _ok(_lib.p4b_calculateBigPDecks(cNode))


def p4_calculateAllBigPDecksAllParts(cTree):
    _ok(_lib.p4b_calculateAllBigPDecksAllParts(cTree))


def p4_setConditionalLikelihoodsOfInternalNodePart(cNode, pNum):
    _ok(_lib.p4b_setConditionalLikelihoodsOfInternalNodePart(cNode, pNum))


def p4_partLogLike(cTree, cPart, pNum, getSiteLikes):
    v = _lib.p4b_partLogLike(cTree, cPart, pNum, int(getSiteLikes))
    if v != v:
        _fatal()
    return v


def p4_treeLogLike(cTree, getSiteLikes):
    v = _lib.p4b_treeLogLike(cTree, int(getSiteLikes))
    if v != v and (_lastError() or b""):
        _fatal()
    return v


# ---- optimisers (callers of the hot path) ----------------------------
def windUpParameters(cTree, doBrLens=1):
    """(x, lower, upper): the reference's parameter vector and bounds (Pf/p4_treeOpt.c:17-120)."""
    n = _lib.p4b_countParameters(cTree, int(doBrLens))
    if n < 0:
        _fatal()
    x, lo, hi = np.zeros(n), np.zeros(n), np.zeros(n)
    if _lib.p4b_windUpParameters(cTree, int(doBrLens), x.ctypes.data, lo.ctypes.data, hi.ctypes.data) != n:
        _fatal()
    return x, lo, hi


def logLikeForParameters(cTree, doBrLens, x):
    """The optimisers' objective: unwind, p4_setPrams, p4_treeLogLike (Pf/p4_treeOpt.c:579-615)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    v = _lib.p4b_logLikeForParameters(cTree, int(doBrLens), x.ctypes.data)
    if v != v:
        _fatal()
    return v


def p4_getBrLens(cTree):
    """pf.p4_getBrLens(cTree) -> list of branch lengths by node number (-1.0 for the root)."""
    out = np.zeros(_lib.p4b_treeNNodes(cTree))
    _ok(_lib.p4b_getBrLens(cTree, out.ctypes.data))
    return [float(v) for v in out]


def p4_getFreePrams(cTree):
    """pf.p4_getFreePrams(cTree) -> list: the model part of the parameter vector."""
    return [float(v) for v in windUpParameters(cTree, 0)[0]]


def _optResult(n):
    if n < 0:
        _fatal()
    return int(n)


def p4_allBOBYQAOptimize(cTree, doBrLens=1, verbose=0):
    """pf.p4_allBOBYQAOptimize(cTree, doBrLens) (Pf/pfmodule.c:2212 -> Pf/p4_treeOpt.c:617-753): maximise lnL over the
    free model parameters and, with doBrLens, all branch lengths, inside the reference's box of bounds.

    Native (csrc/opt.cpp p4b_allBOBYQAOptimize): the reference's two-pass schedule on the reference's objective
    (p4_unWindParameters -> p4_setPrams -> p4_treeLogLike, evaluated on the GPU).  Where the reference calls nlopt's
    BOBYQA -- a third-party library this engine does not link -- a bounded Powell method runs (csrc/praxis.cpp): the
    contract kept is the box and the optimum, not BOBYQA's trajectory.  Returns the number of likelihood evaluations."""
    n = _optResult(_lib.p4b_allBOBYQAOptimize(cTree, int(doBrLens)))
    if verbose:
        print("p4_allBOBYQAOptimize: %d likelihood evaluations, lnL %.6f" % (n, p4_treeLogLike(cTree, 0)))
    return n


def p4_allBrentPowellOptimize(cTree, verbose=0):
    """pf.p4_allBrentPowellOptimize(cTree) (Pf/pfmodule.c:2244 -> Pf/p4_treeOpt.c:996-1180): Brent's praxis over model
    parameters and branch lengths, the reference's schedule (tol 1e-4, h 0.1 until a call gains < 1e-6, then h 0.05).
    Native: csrc/praxis.cpp restates Brent's published algorithm with the reference's settings."""
    n = _optResult(_lib.p4b_allBrentPowellOptimize(cTree))
    if verbose:
        print("p4_allBrentPowellOptimize: %d likelihood evaluations, lnL %.6f" % (n, p4_treeLogLike(cTree, 0)))
    return n


def optimizeBrLens(cTree, maxPasses=1, tol=1e-6):
    """All branch lengths, one at a time, each through the dirty path (include/p4b200.h p4b_optimizeBrLens).
    Returns (lnL, number of likelihood evaluations)."""
    n = C.c_long(0)
    v = _lib.p4b_optimizeBrLens(cTree, int(maxPasses), float(tol), C.byref(n))
    if v != v:
        _fatal()
    return v, n.value


def p4_newtSetup(cTree):
    """pf.p4_newtSetup(cTree) (Pf/pfmodule.c p4_newtSetup -> Pf/p4_treeNewt.c:11-75): allocate cl2 -- per node,
    the conditional likelihoods of everything on the far side of its branch -- and the work space of the
    derivative P decks on the device.  Idempotent."""
    if not cTree:
        _fatal()
    _ok(_lib.p4b_newtSetup(cTree))


def newtAround(cTree, epsilon, likeDelta):
    """p4_newtAround(aTree, epsilon, likeDelta) (Pf/p4_treeNewt.c:78-205): Newton-Raphson on every branch
    length in post-order, round after round until lnL moves by less than likeDelta.  Returns lnL."""
    v = _lib.p4b_newtAround(cTree, float(epsilon), float(likeDelta))
    if v != v:
        _fatal()
    return v


def newtDerivs(cNode):
    """(lnL, d lnL/dv, d2 lnL/dv2) in the length v of the node's branch, as p4_newtNode forms them."""
    out = np.empty(3, dtype=np.float64)
    _ok(_lib.p4b_newtDerivs(cNode, out.ctypes.data))
    return float(out[0]), float(out[1]), float(out[2])


def getNodeCL2(cTree, cNode, pNum, nCat, dim):
    """cl2[cat][state][pattern] of one node for the patterns resident on this device."""
    lo, hi = treeShardRange(cTree, pNum)
    out = np.empty((nCat, dim, hi - lo), dtype=np.float64)
    _ok(_lib.p4b_getNodeCL2(cNode, pNum, out.ctypes.data))
    return out


def newtIterations(cTree):
    return int(_lib.p4b_newtIterations(cTree))


def p4_newtAndBrentPowellOpt(cTree, verbose=0):
    """pf.p4_newtAndBrentPowellOpt(cTree) (Pf/pfmodule.c:2259 -> Pf/p4_treeOpt.c:1182-1330): branch lengths by
    Newton-Raphson (p4_newtAround on the device, the reference's schedule of tolerances), free model parameters by
    Brent's praxis on the reference's parameter vector, alternating until a round gains less than 1e-6 or
    var.newtAndBrentPowellOptPassLimit rounds have run.  With no free parameter it is exactly the reference's four
    p4_newtAround calls (:1214-1226); with one, Newton + a one-dimensional Brent search (:1395-1436).  Native
    (csrc/opt.cpp).  Returns the number of objective evaluations of the model step."""
    n = _optResult(_lib.p4b_newtAndBrentPowellOpt(cTree))
    if verbose:
        print("p4_newtAndBrentPowellOpt: %d likelihood evaluations, lnL %.6f" % (n, p4_treeLogLike(cTree, 0)))
    return n


def p4_newtAndBOBYQAOpt(cTree, verbose=0):
    """pf.p4_newtAndBOBYQAOpt(cTree) (Pf/pfmodule.c:2227 -> Pf/p4_treeOpt.c:755-945): as above with the bounded method
    for the model step (pass limit 50).  Native (csrc/opt.cpp)."""
    n = _optResult(_lib.p4b_newtAndBOBYQAOpt(cTree))
    if verbose:
        print("p4_newtAndBOBYQAOpt: %d likelihood evaluations, lnL %.6f" % (n, p4_treeLogLike(cTree, 0)))
    return n


def praxisMinimize(fn, x0, tol=1.0e-4, h=1.0):
    """Brent's principal-axis minimiser (csrc/praxis.cpp) on a Python objective: (minimum, x).  For tests."""
    x = np.array(x0, dtype=np.float64)
    cb = _OBJ(lambda p, ctx: float(fn(np.ctypeslib.as_array(p, shape=(len(x),)))))
    v = _lib.p4b_praxisMinimize(len(x), x.ctypes.data, float(tol), float(h), cb, None)
    return v, x


def boundedMinimize(fn, x0, lo, hi, xtol=1.0e-7, ftol=1.0e-12, maxEvals=100000):
    """Powell's method inside the box (csrc/praxis.cpp) on a Python objective: (minimum, x, evaluations).  For tests."""
    x = np.array(x0, dtype=np.float64)
    lo = np.ascontiguousarray(lo, dtype=np.float64)
    hi = np.ascontiguousarray(hi, dtype=np.float64)
    cb = _OBJ(lambda p, ctx: float(fn(np.ctypeslib.as_array(p, shape=(len(x),)))))
    n = C.c_long(0)
    v = _lib.p4b_boundedMinimize(len(x), x.ctypes.data, lo.ctypes.data, hi.ctypes.data, float(xtol), float(ftol), int(maxEvals), cb, None, C.byref(n))
    return v, x, n.value


# ---- consumers of the P decks beside the likelihood -----------------------------------------
def _expected(cTree, pNum, fn):
    nTax, dim = _lib.p4b_treeNLeaves(cTree), _lib.p4b_treePartDim(cTree, pNum)
    if nTax < 0 or dim < 0:
        _fatal()
    out = np.zeros((nTax, dim), dtype=np.float64)
    _ok(fn(cTree, int(pNum), out.ctypes.data))
    return tuple(tuple(float(v) for v in row) for row in out)


def gsl_rng_get():
    """pf.gsl_rng_get() -> handle of a new mt19937 stream (Pf/pfmodule.c:674; GSL's default generator and seeding)."""
    return _lib.p4b_rngNew()


def gsl_rng_free(g):
    _lib.p4b_rngFree(g)


def gsl_rng_set(g, seed):
    """pf.gsl_rng_set(g, seed) (Pf/pfmodule.c:708)."""
    _lib.p4b_rngSet(g, int(seed) & 0xFFFFFFFFFFFFFFFF)


def gsl_rng_uniform(g):
    """pf.gsl_rng_uniform(g) -> float in [0, 1) (Pf/pfmodule.c:739)."""
    return _lib.p4b_rngUniform(g)


def gsl_rng_uniform_array(g, n):
    """The next n uniforms of the stream as a numpy array (an addition: what pf.p4_simulate draws from, in bulk)."""
    out = np.empty(int(n), dtype=np.float64)
    _lib.p4b_rngFillUniform(g, out.ctypes.data, int(n))
    return out


def gsl_rng_size(g):
    """pf.gsl_rng_size(g) (Pf/pfmodule.c:724): bytes of generator state (what Mcmc checkpoints pickle)."""
    return int(_lib.p4b_rngSize(g))


def gsl_rng_getstate(g, byteArray):
    """pf.gsl_rng_getstate(g, numpy byte array) (Pf/pfmodule.c:752)."""
    a = np.asarray(byteArray)
    if a.nbytes < gsl_rng_size(g):
        raise ValueError("gsl_rng_getstate: the array is smaller than the generator state")
    _lib.p4b_rngGetState(g, a.ctypes.data)


def gsl_rng_setstate(g, byteArray):
    """pf.gsl_rng_setstate(g, numpy byte array) (Pf/pfmodule.c:779)."""
    a = np.ascontiguousarray(byteArray)
    if a.nbytes < gsl_rng_size(g):
        raise ValueError("gsl_rng_setstate: the array is smaller than the generator state")
    _lib.p4b_rngSetState(g, a.ctypes.data)


def gsl_ran_gamma(g, a, b):
    """pf.gsl_ran_gamma(g, a, b) -> a draw from the gamma distribution (Pf/pfmodule.c:827)."""
    return _lib.p4b_ranGamma(g, float(a), float(b))


def gsl_ran_dirichlet(g, k, alpha, theta):
    """pf.gsl_ran_dirichlet(g, k, alpha, theta): theta is overwritten with the draw (Pf/pfmodule.c:940)."""
    _lib.p4b_ranDirichlet(g, int(k), _arr(alpha, np.float64, "alpha"), _arr(theta, np.float64, "theta"))


def gsl_ran_dirichlet_lnpdf(k, alpha, theta):
    """pf.gsl_ran_dirichlet_lnpdf(k, alpha, theta) -> float (Pf/pfmodule.c:989)."""
    return _lib.p4b_ranDirichletLnPdf(int(k), _arr(alpha, np.float64, "alpha"), _arr(theta, np.float64, "theta"))


def gsl_ran_dirichlet_pdf(k, alpha, theta):
    return float(np.exp(gsl_ran_dirichlet_lnpdf(k, alpha, theta)))


def gsl_sf_lngamma(a):
    """pf.gsl_sf_lngamma(a) (Pf/pfmodule.c:887)."""
    return _lib.p4b_sfLnGamma(float(a))


def gsl_ran_gamma_pdf(x, a, b):
    """pf.gsl_ran_gamma_pdf(x, a, b) (Pf/pfmodule.c:848)."""
    return _lib.p4b_ranGammaPdf(float(x), float(a), float(b))


def gsl_meanVariance(seq, seqLen, mean, variance):
    """pf.gsl_meanVariance(seq, N, mean, variance): the two 1-element arrays are filled (Pf/pfmodule.c:1025)."""
    _lib.p4b_meanVariance(_arr(seq, np.float64, "seq"), int(seqLen), _arr(mean, np.float64, "mean"), _arr(variance, np.float64, "variance"))


_mcmcTreeCallbacks = {}


def setMcmcTreeCallback(cTree, fn):
    """pf.setMcmcTreeCallback(cTree, callable) (Pf/pfmodule.c:2727): the reference stores the callable on the tree and
    only ever calls it from code that is compiled out (Pf/p4_node.c:420-438); it is kept here for the same lifetime."""
    if not callable(fn):
        raise TypeError("pf_setMcmcTreeCallback(): parameter must be callable")
    _mcmcTreeCallbacks[cTree] = fn


def unsetMcmcTreeCallback(cTree):
    _mcmcTreeCallbacks.pop(cTree, None)


def p4_simulate(cTree, cRefTree, g):
    """pf.p4_simulate(cTree, cRefTree|0, gsl_rng) (Pf/pfmodule.c:2333, Pf/p4_treeSim.c:14-420): simulate new
    sequences down the tree on the device; the same seed gives the reference's sequences."""
    _ok(_lib.p4b_simulate(cTree, cRefTree if cRefTree else None, g))


def p4_drawAncState(cTree, partNum, seqPos, draw):
    """pf.p4_drawAncState(cTree, partNum, seqPos, draw) (Pf/pfmodule.c:2353): fills the int32 array ``draw`` with
    {chStNum, catNum, isInvar, invarChNum} for one draw from the root's posterior at that site."""
    _ok(_lib.p4b_drawAncState(cTree, int(partNum), int(seqPos), _arr(draw, np.int32, "draw")))


def bootstrapData(cDataReference, cDataToFill, g):
    """pf.bootstrapData(referenceData, toFillData, gsl_rng) (Pf/pfmodule.c:90, Pf/data.c:107-139)."""
    _ok(_lib.p4b_bootstrapData(cDataReference, cDataToFill, g))


def drawAncStateFromCL(cPart, seqPos, nCat, pInvar, pInvarFree, pi, rootCL):
    """Test hook (include/p4b200.h p4b_drawAncStateFromCL): p4_drawAncState's draw for a given root CL [cat][state][pattern]."""
    pi = np.ascontiguousarray(pi, dtype=np.float64)
    cl = np.ascontiguousarray(rootCL, dtype=np.float64)
    d = np.empty(4, dtype=np.int32)
    _ok(_lib.p4b_drawAncStateFromCL(cPart, int(seqPos), int(nCat), float(pInvar), int(pInvarFree), pi.ctypes.data, cl.ctypes.data, d.ctypes.data))
    return [int(v) for v in d]


def reseedCRandomizer(seed):
    """pf.reseedCRandomizer(seed) (Pf/pfmodule.c:472): srandom(seed)."""
    _lib.p4b_reseedCRandomizer(int(seed))


def p4_expectedComposition(cTree):
    """pf.p4_expectedComposition(cTree) -> tuple over parts of tuple over sequences of tuple over states
    (Pf/pfmodule.c:2404, Pf/p4_treeSim.c:951-1045): the composition the model expects at every tip."""
    return tuple(_expected(cTree, p, _lib.p4b_expectedComposition) for p in range(_lib.p4b_treeNParts(cTree)))


def p4_expectedCompositionCounts(cTree, partNum):
    """pf.p4_expectedCompositionCounts(cTree, partNum) (Pf/pfmodule.c:2384, Pf/p4_treeSim.c:859-949)."""
    return _expected(cTree, partNum, _lib.p4b_expectedCompositionCounts)


# ---- cur/prop state transfer ------------------------------------------------------------
def p4_copyCondLikes(cTreeA, cTreeB, doAll):
    _ok(_lib.p4b_copyCondLikes(cTreeA, cTreeB, int(doAll)))


def p4_copyBigPDecks(cTreeA, cTreeB, doAll):
    _ok(_lib.p4b_copyBigPDecks(cTreeA, cTreeB, int(doAll)))


def p4_copyModelPrams(cTreeA, cTreeB):
    _ok(_lib.p4b_copyModelPrams(cTreeA, cTreeB))


def p4_verifyIdentityOfTwoTrees(cTreeA, cTreeB):
    r = _lib.p4b_verifyIdentityOfTwoTrees(cTreeA, cTreeB)
    if r < 0:
        _fatal()
    return r


# ---- inspection (additions) ----------------------------------------------------------------
def treeShardRange(cTree, pNum):
    lo, hi = C.c_int(), C.c_int()
    _ok(_lib.p4b_treeShardRange(cTree, pNum, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def getNodeCL(cTree, cNode, pNum, nCat, dim):
    """cl[cat][state][pattern] of one node for the patterns resident on this device."""
    lo, hi = treeShardRange(cTree, pNum)
    out = np.empty((nCat, dim, hi - lo), dtype=np.float64)
    _ok(_lib.p4b_getNodeCL(cNode, pNum, out.ctypes.data))
    return out


def getNodeBigP(cNode, pNum, nCat, dim):
    out = np.empty((nCat, dim, dim), dtype=np.float64)
    _ok(_lib.p4b_getNodeBigP(cNode, pNum, out.ctypes.data))
    return out


def setNodeBigP(cNode, pNum, bigP):
    """Test hook: overwrite one node's P deck (nCat, dim, dim)."""
    a = np.ascontiguousarray(bigP, dtype=np.float64)
    _ok(_lib.p4b_setNodeBigP(cNode, pNum, a.ctypes.data))


def setTreeStoresCL(cTree, on):
    """0: whole-tree evaluations keep only the CLs they re-read (lnL-only mode); see include/p4b200.h."""
    _ok(_lib.p4b_setTreeStoresCL(cTree, int(on)))


def treeSync(cTree):
    _ok(_lib.p4b_treeSync(cTree))


def treeTimerBegin(cTree):
    _ok(_lib.p4b_treeTimerBegin(cTree))


def treeTimerEnd(cTree):
    ms = _lib.p4b_treeTimerEnd(cTree)
    if ms < 0:
        _fatal()
    return ms


def treeLastCLTiming(cTree):
    ms, n = C.c_double(), C.c_int()
    _ok(_lib.p4b_treeLastCLTiming(cTree, C.byref(ms), C.byref(n)))
    return ms.value, n.value


def treeDeviceBytes(cTree):
    return _lib.p4b_treeDeviceBytes(cTree)


def flushL2(cTree):
    _ok(_lib.p4b_flushL2(cTree))


# ---- fast bindings of the per-node calls ------------------------------------------------------------
# csrc/pfhot.c holds the same wrappers as METH_FASTCALL C functions (about 0.1 us per call instead of 0.5 us
# through ctypes; p4 issues 4 per node before every likelihood calculation).  Same names, arguments, errors.
_hot = None
_HOT_NAMES = ("p4_setNodeRelation", "p4_setTreeRoot", "p4_setBrLen", "p4_setCompNum", "p4_setRMatrixNum", "p4_setGdasrvNum",
              "p4_setRMatrixBigR", "p4_setKappa", "p4_setPInvarVal", "p4_setRelRateVal", "p4_setPrams", "p4_calculateBigPDecks",
              "p4_setConditionalLikelihoodsOfInternalNodePart", "p4_partLogLike", "p4_treeLogLike",
              "p4_copyCondLikes", "p4_copyBigPDecks", "p4_copyModelPrams")
_ctypes_versions = {}


def use_fast_bindings(on=True):
    """Install (or remove) the C bindings of csrc/pfhot.c over the ctypes wrappers.  Returns True if they are active."""
    global _hot
    g = globals()
    if on and _hot is None:
        try:
            import importlib.util
            path = _build.hot_path()
            if not os.path.exists(path):
                return False
            spec = importlib.util.spec_from_file_location("_pfhot", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.set_fatal(P4bFatal)
            _hot = mod
        except Exception:      # binding layer only: the ctypes wrappers call the same engine
            _hot = None
            return False
    for n in _HOT_NAMES:
        _ctypes_versions.setdefault(n, g[n])
        g[n] = getattr(_hot, n) if (on and _hot is not None) else _ctypes_versions[n]
    return bool(on and _hot is not None)


use_fast_bindings(True)

_passthrough = None


def set_passthrough(module):
    """Delegate every name this module does not define (GSL wrappers, statistics, simulation, ...
    -- everything outside the likelihood path) to ``module``, normally the stock ``p4.pf``.
    Objects created by one engine must not be handed to the other."""
    global _passthrough
    _passthrough = module


def __getattr__(name):
    if _passthrough is not None and hasattr(_passthrough, name):
        return getattr(_passthrough, name)
    raise AttributeError("p4b200 pf mirror has no '%s': only the likelihood hot path of p4.pf is provided "
                         "(SURVEY.md section 8b)" % name)
