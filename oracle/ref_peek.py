"""oracle/ref_peek.py -- TEST INFRASTRUCTURE ONLY.

Reads arrays out of the reference engine's C structs so that parity can be
checked array by array, not only on the final log-likelihood.  The reference's
``pf`` module hands out raw struct addresses as Python ints
(Pf/pfmodule.c:1405) but has no accessors for patterns, conditional
likelihoods or transition matrices; these ctypes mirrors of the struct heads
(Pf/pftypes.h:31-53 partStruct, :116-146 p4_nodeStruct) let tests read them.
"""
import ctypes as C

import numpy as np

_ipp = C.POINTER(C.POINTER(C.c_int))
_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)
_dpp = C.POINTER(_dp)
_dppp = C.POINTER(_dpp)
_dpppp = C.POINTER(_dppp)


class PartStruct(C.Structure):          # Pf/pftypes.h:31-53
    _fields_ = [("data", C.c_void_p), ("dim", C.c_int), ("nTax", C.c_int), ("nChar", C.c_int),
                ("symbols", C.c_char_p), ("patterns", _ipp), ("patternCounts", _ip),
                ("sequencePositionPatternIndex", _ip), ("nPatterns", C.c_int), ("sequences", _ipp),
                ("nEquates", C.c_int), ("equates", _ipp), ("equateSymbols", C.c_char_p),
                ("globalInvarSitesVec", _ip), ("globalInvarSitesArray", _ipp), ("siteLikes", _dp)]


class NodeStruct(C.Structure):          # Pf/pftypes.h:116-146
    _fields_ = [("nodeNum", C.c_int), ("tree", C.c_void_p), ("parent", C.c_void_p), ("leftChild", C.c_void_p),
                ("sibling", C.c_void_p), ("seqNum", C.c_int), ("isLeaf", C.c_int), ("nParts", C.c_int),
                ("brLen", _dp), ("savedBrLen", C.c_double), ("compNums", _ip), ("rMatrixNums", _ip),
                ("gdasrvNums", _ip), ("bigPDecks", _dpppp), ("bigPDecks_1stD", _dpppp), ("bigPDecks_2ndD", _dpppp),
                ("cl", _dpppp), ("cl2", _dpppp), ("pickerDecks", _dpppp), ("clNeedsUpdating", C.c_int)]


def _rows(pp, nRows, nCols, dtype):
    # pimatrix / pdmatrix: row pointers over one contiguous block (Pf/pmatrices.c:6-27, 163)
    base = pp[0]
    return np.ctypeslib.as_array(base, shape=(nRows * nCols,)).reshape(nRows, nCols).astype(dtype, copy=True)


def part_arrays(cPart):
    p = PartStruct.from_address(cPart)
    n = p.nChar
    out = {"nPatterns": p.nPatterns, "dim": p.dim, "nTax": p.nTax, "nChar": n,
           "sequences": _rows(p.sequences, p.nTax, n, np.int32),
           "patterns": _rows(p.patterns, p.nTax, n, np.int32),
           "patternCounts": np.ctypeslib.as_array(p.patternCounts, shape=(n,)).copy(),
           "sequencePositionPatternIndex": np.ctypeslib.as_array(p.sequencePositionPatternIndex, shape=(n,)).copy(),
           "globalInvarSitesVec": None, "globalInvarSitesArray": None}
    if p.globalInvarSitesVec:
        out["globalInvarSitesVec"] = np.ctypeslib.as_array(p.globalInvarSitesVec, shape=(n,)).copy()
    if p.globalInvarSitesArray:
        out["globalInvarSitesArray"] = _rows(p.globalInvarSitesArray, p.dim, n, np.int32)
    return out


def node_cl(cNode, pNum, nCat, dim, nChar, nPatterns):
    """cl[cat][state][pattern] of one reference node (patterns < nPatterns)."""
    nd = NodeStruct.from_address(cNode)
    out = np.empty((nCat, dim, nPatterns))
    for c in range(nCat):
        m = nd.cl[pNum][c]
        out[c] = np.ctypeslib.as_array(m[0], shape=(dim * nChar,)).reshape(dim, nChar)[:, :nPatterns]
    return out


def node_bigP(cNode, pNum, nCat, dim):
    nd = NodeStruct.from_address(cNode)
    out = np.empty((nCat, dim, dim))
    for c in range(nCat):
        out[c] = np.ctypeslib.as_array(nd.bigPDecks[pNum][c][0], shape=(dim * dim,)).reshape(dim, dim)
    return out


def node_cl2(cNode, pNum, nCat, dim, nChar, nPatterns):
    """cl2[cat][state][pattern] of one reference node (after p4_newtSetup; Pf/pftypes.h:134)."""
    nd = NodeStruct.from_address(cNode)
    out = np.empty((nCat, dim, nPatterns))
    for c in range(nCat):
        m = nd.cl2[pNum][c]
        out[c] = np.ctypeslib.as_array(m[0], shape=(dim * nChar,)).reshape(dim, nChar)[:, :nPatterns]
    return out


def node_brlen(cNode):
    return float(NodeStruct.from_address(cNode).brLen[0])


_newt_lib = None


def newt_lib():
    """The reference engine's Newton-Raphson entry points, which its pf module does not wrap one by one
    (Pf/p4_tree.h:50-51, Pf/p4_treeNewt.c): p4_newtAround, p4_newtNode, p4_setNodeCL2."""
    global _newt_lib
    if _newt_lib is None:
        import ref_loader
        lib = C.CDLL(ref_loader.ref_pf_path())
        lib.p4_newtAround.argtypes = [C.c_void_p, C.c_double, C.c_double]
        lib.p4_newtAround.restype = None
        lib.p4_newtNode.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        lib.p4_newtNode.restype = None
        lib.p4_setNodeCL2.argtypes = [C.c_void_p, C.c_void_p]
        lib.p4_setNodeCL2.restype = None
        _newt_lib = lib
    return _newt_lib
