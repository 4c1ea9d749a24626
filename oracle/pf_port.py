"""oracle/pf_port.py -- TEST INFRASTRUCTURE ONLY.

ctypes front end of the plain-C restatement in pf_port.c (libpforacle.so).  It
takes the host-side objects of ``p4_phylogenetics_b200.host`` (tree, data as
character strings, model) and evaluates them entirely on the CPU with the
oracle's own code: its own character coding, pattern compression, constant-site
masks, gamma rates, Q, P(t), CL recursion and log-likelihood.  Nothing from the
product library is called.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)


class _Tree(C.Structure):
    _fields_ = [("nNodes", C.c_int), ("root", C.c_int), ("nPost", C.c_int), ("parent", _ip), ("leftChild", _ip),
                ("sibling", _ip), ("isLeaf", _ip), ("seqNum", _ip), ("brLen", _dp), ("postOrder", _ip),
                ("compNum", _ip), ("rMatrixNum", _ip), ("gdasrvNum", _ip)]


class _Part(C.Structure):
    _fields_ = [("nTax", C.c_int), ("nPatterns", C.c_int), ("stride", C.c_int), ("patterns", _ip), ("patternCounts", _ip),
                ("nEquates", C.c_int), ("equates", _ip), ("invarVec", _ip), ("invarArray", _ip)]


class _Model(C.Structure):
    _fields_ = [("dim", C.c_int), ("nCat", C.c_int), ("nComps", C.c_int), ("nRMatrices", C.c_int), ("nGdasrvs", C.c_int),
                ("comps", _dp), ("bigR", _dp), ("rates", _dp), ("pInvar", C.c_double), ("relRate", C.c_double)]


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libpforacle.so")
        if not os.path.exists(path):
            raise ImportError("oracle/libpforacle.so not built: run `make -C oracle libpforacle.so`")
        _lib = C.CDLL(path)
        _lib.pfport_part_loglike.restype = C.c_double
        _lib.pfport_part_loglike.argtypes = [C.POINTER(_Tree), C.POINTER(_Part), C.POINTER(_Model), _dp, _dp, _dp]
        _lib.pfport_part_loglike_ld.restype = C.c_double
        _lib.pfport_part_loglike_ld.argtypes = _lib.pfport_part_loglike.argtypes
        _lib.pfport_make_patterns.restype = C.c_int
    return _lib


def _i(a):
    return a.ctypes.data_as(_ip)


def _d(a):
    return a.ctypes.data_as(_dp)


def compress(sequences, symbols, equates):
    """Character coding + pattern compression + constant-site masks of one alignment (Pf/part.c)."""
    L = lib()
    nTax, nChar, dim = len(sequences), len(sequences[0]), len(symbols)
    eqSymb = "".join(sorted(equates.keys()))
    joined = b"".join(s if isinstance(s, (bytes, bytearray)) else s.encode("latin-1") for s in sequences)
    seq = np.zeros((nTax, nChar), dtype=np.int32)
    rc = L.pfport_poke_sequences(joined, nTax, nChar, symbols.encode(), dim, eqSymb.encode(), len(eqSymb), _i(seq))
    if rc:
        raise ValueError("a character is neither in the symbols nor in the equates")
    eq = np.zeros((max(len(eqSymb), 1), dim), dtype=np.int32)
    for i, e in enumerate(eqSymb):
        for j, s in enumerate(symbols):
            eq[i, j] = 1 if s in equates[e] else 0
    pat = np.zeros((nTax, nChar), dtype=np.int32)
    counts = np.zeros(nChar, dtype=np.int32)
    index = np.zeros(nChar, dtype=np.int32)
    nPat = L.pfport_make_patterns(_i(seq), nTax, nChar, _i(pat), _i(counts), _i(index))
    vec = np.zeros(nChar, dtype=np.int32)
    arr = np.zeros((dim, nChar), dtype=np.int32)
    L.pfport_invar_sites(_i(pat), nTax, nChar, nPat, dim, _i(eq), _i(vec), _i(arr))
    return {"nPatterns": nPat, "sequences": seq, "patterns": pat, "patternCounts": counts,
            "sequencePositionPatternIndex": index, "globalInvarSitesVec": vec, "globalInvarSitesArray": arr,
            "equates": eq, "nEquates": len(eqSymb), "dim": dim, "nTax": nTax, "nChar": nChar}


def discrete_gamma(alpha, K):
    f, r = np.zeros(K), np.zeros(K)
    lib().pfport_discrete_gamma(C.c_double(alpha), K, _d(f), _d(r))
    return f, r


def big_q(R, pi):
    dim = len(pi)
    Q = np.zeros((dim, dim))
    lib().pfport_big_q(_d(np.ascontiguousarray(R, dtype=np.float64)), _d(np.ascontiguousarray(pi, dtype=np.float64)), dim, _d(Q))
    return Q


def _big_r(rm, dim):
    """The exchangeability matrix the reference would hold for one rMatrix (Pf/p4_model.c:346-435, p4/model.py:185-207)."""
    R = np.ones((dim, dim))
    if rm.spec == "2p":
        a, b = 1.0 / 3.0, (1.0 / 3.0) * float(rm.val[0])
        R = np.array([[0, a, b, a], [a, 0, a, b], [b, a, 0, a], [a, b, a, 0]], dtype=np.float64)
    elif rm.spec == "specified" or (rm.free and rm.spec not in ("ones",) and rm.val is not None and rm.spec in ("optimized",)):
        k = 0
        for i in range(dim - 1):
            for j in range(i + 1, dim):
                R[i, j] = R[j, i] = rm.val[k]
                k += 1
    elif rm.spec == "ones":
        if rm.free and rm.val is not None:
            k = 0
            for i in range(dim - 1):
                for j in range(i + 1, dim):
                    R[i, j] = R[j, i] = rm.val[k]
                    k += 1
    else:
        R = protein_big_r(rm.spec)
    return R


_protein_cache = {}


def protein_big_r(spec):
    """Published empirical exchangeabilities, from the fixture written by tools/gen_protein_tables.py."""
    if not _protein_cache:
        path = os.path.join(_HERE, "..", "p4-phylogenetics_b200", "csrc", "protein_rmatrices.inc")
        txt = open(path).read()
        names = txt.split("kProteinSpecNames[] = {")[1].split("}")[0].replace('"', "").replace(" ", "").split(",")
        body = txt.split("kProteinBigR[][400] = {")[1]
        blocks = body.split("{ //")[1:]
        for nm, blk in zip(names, blocks):
            nums = blk.split("\n", 1)[1].split("}")[0].replace("\n", " ").split(",")
            vals = [float(x) for x in nums if x.strip()]
            _protein_cache[nm] = np.array(vals[:400]).reshape(20, 20)
    return _protein_cache[spec]


def _structs(tree):
    """Per part: the C structs of pf_port.h for a ``host.Tree`` (data and model attached), and the arrays they point into."""
    nodes = tree.nodes
    nN = len(nodes)
    num = lambda x: x.nodeNum if x is not None else -1
    parent = np.array([num(n.parent) for n in nodes], dtype=np.int32)
    left = np.array([num(n.leftChild) for n in nodes], dtype=np.int32)
    sib = np.array([num(n.sibling) for n in nodes], dtype=np.int32)
    isLeaf = np.array([int(n.isLeaf) for n in nodes], dtype=np.int32)
    seqNum = np.array([int(n.seqNum) for n in nodes], dtype=np.int32)
    brLen = np.array([float(n.br.len) for n in nodes], dtype=np.float64)
    if not tree.preAndPostOrderAreValid:
        tree.setPreAndPostOrder()
    post = np.ascontiguousarray(tree.postOrder, dtype=np.int32)
    out = []
    for pNum, (aln, mp) in enumerate(zip(tree.data.alignments, tree.model.parts)):
        comp = compress(aln.sequences, aln.symbols, aln.equates)
        dim, nCat = mp.dim, mp.nGammaCat
        compNum = np.array([n.parts[pNum].compNum for n in nodes], dtype=np.int32)
        rNum = np.array([n.br.parts[pNum].rMatrixNum for n in nodes], dtype=np.int32)
        gNum = np.array([n.br.parts[pNum].gdasrvNum for n in nodes], dtype=np.int32)
        comps = np.ascontiguousarray(np.stack([c.val for c in mp.comps]), dtype=np.float64)
        bigR = np.ascontiguousarray(np.stack([_big_r(r, dim) for r in mp.rMatrices]), dtype=np.float64)
        rates = np.ones((max(len(mp.gdasrvs), 1), nCat))
        for gi, g in enumerate(mp.gdasrvs):
            rates[gi] = discrete_gamma(float(g.val[0]), nCat)[1]
        T = _Tree(nN, tree.root.nodeNum, len(post), _i(parent), _i(left), _i(sib), _i(isLeaf), _i(seqNum), _d(brLen), _i(post),
                  _i(compNum), _i(rNum), _i(gNum))
        D = _Part(comp["nTax"], comp["nPatterns"], comp["nChar"], _i(comp["patterns"]), _i(comp["patternCounts"]),
                  comp["nEquates"], _i(comp["equates"]), _i(comp["globalInvarSitesVec"]), _i(comp["globalInvarSitesArray"]))
        M = _Model(dim, nCat, len(mp.comps), len(mp.rMatrices), len(mp.gdasrvs), _d(comps), _d(bigR), _d(rates),
                   float(mp.pInvar.val), float(mp.relRate))
        keep = (parent, left, sib, isLeaf, seqNum, brLen, post, compNum, rNum, gNum, comps, bigR, rates, comp)
        out.append((T, D, M, comp, rates, keep))
    return out


def branch_derivs(tree):
    """{nodeNum: (lnL, d lnL/dv, d2 lnL/dv2)} in every branch length v, summed over the parts: the three sums of one
    iteration of p4_newtNode (Pf/p4_treeNewt.c:238-520), by the oracle port alone (pfport_branch_derivs)."""
    L = lib()
    L.pfport_branch_derivs.restype = C.c_int
    parts = _structs(tree)
    res = {}
    for n in tree.nodes:
        if n is tree.root or n.parent is None:
            continue
        tot = np.zeros(3)
        for (T, D, M, comp, rates, keep) in parts:
            o = np.zeros(3)
            if L.pfport_branch_derivs(C.byref(T), C.byref(D), C.byref(M), int(n.nodeNum), _d(o)):
                raise ValueError("pfport_branch_derivs: bad node %d" % n.nodeNum)
            tot += o
        res[n.nodeNum] = (float(tot[0]), float(tot[1]), float(tot[2]))
    return res


def newt_node(tree, node, epsilon, brLenMin=1.0e-8, brLenMax=3.0):
    """p4_newtNode (Pf/p4_treeNewt.c:210-600) for one node of a ``host.Tree``: Newton-Raphson on its branch length with the
    reference's guards, the derivatives from the oracle port.  Sets and returns node.br.len."""
    old = cur = float(node.br.len)
    it = 0
    while True:
        node.br.len = cur
        _, first, second = branch_derivs_of(tree, node)
        nxt = cur - (first / second) if second != 0.0 else float("inf")
        if second >= 0.0:
            nxt = cur / 5.0
        if nxt < brLenMin:
            cur = brLenMin
            break
        if nxt >= 5.0 * old:
            cur = 5.0 * old
            break
        if nxt > brLenMax:
            cur = brLenMax
            break
        if it > 20:
            break
        it += 1
        if abs(first) < epsilon:
            break
        cur = nxt
    node.br.len = cur
    return cur


def branch_derivs_of(tree, node):
    L = lib()
    L.pfport_branch_derivs.restype = C.c_int
    tot = np.zeros(3)
    for (T, D, M, comp, rates, keep) in _structs(tree):
        o = np.zeros(3)
        if L.pfport_branch_derivs(C.byref(T), C.byref(D), C.byref(M), int(node.nodeNum), _d(o)):
            raise ValueError("pfport_branch_derivs: bad node %d" % node.nodeNum)
        tot += o
    return float(tot[0]), float(tot[1]), float(tot[2])


def tree_loglike(tree, want_arrays=False, long_double=False):
    """lnL of a ``host.Tree`` (data and model attached) computed by the oracle port alone.

    Returns lnL, or (lnL, partLikes, per-part dict of arrays) with ``want_arrays``.
    ``long_double`` carries the CL recursion in 80-bit floats (no underflow; checks the scalers)."""
    L = lib()
    nodes = tree.nodes
    nN = len(nodes)
    num = lambda x: x.nodeNum if x is not None else -1
    parent = np.array([num(n.parent) for n in nodes], dtype=np.int32)
    left = np.array([num(n.leftChild) for n in nodes], dtype=np.int32)
    sib = np.array([num(n.sibling) for n in nodes], dtype=np.int32)
    isLeaf = np.array([int(n.isLeaf) for n in nodes], dtype=np.int32)
    seqNum = np.array([int(n.seqNum) for n in nodes], dtype=np.int32)
    brLen = np.array([float(n.br.len) for n in nodes], dtype=np.float64)
    if not tree.preAndPostOrderAreValid:
        tree.setPreAndPostOrder()
    post = np.ascontiguousarray(tree.postOrder, dtype=np.int32)
    total, partLikes, extra = 0.0, [], []
    for pNum, (aln, mp) in enumerate(zip(tree.data.alignments, tree.model.parts)):
        comp = compress(aln.sequences, aln.symbols, aln.equates)
        dim, nCat = mp.dim, mp.nGammaCat
        compNum = np.array([n.parts[pNum].compNum for n in nodes], dtype=np.int32)
        rNum = np.array([n.br.parts[pNum].rMatrixNum for n in nodes], dtype=np.int32)
        gNum = np.array([n.br.parts[pNum].gdasrvNum for n in nodes], dtype=np.int32)
        comps = np.ascontiguousarray(np.stack([c.val for c in mp.comps]), dtype=np.float64)
        bigR = np.ascontiguousarray(np.stack([_big_r(r, dim) for r in mp.rMatrices]), dtype=np.float64)
        rates = np.ones((max(len(mp.gdasrvs), 1), nCat))
        for gi, g in enumerate(mp.gdasrvs):
            rates[gi] = discrete_gamma(float(g.val[0]), nCat)[1]
        T = _Tree(nN, tree.root.nodeNum, len(post), _i(parent), _i(left), _i(sib), _i(isLeaf), _i(seqNum), _d(brLen), _i(post),
                  _i(compNum), _i(rNum), _i(gNum))
        D = _Part(comp["nTax"], comp["nPatterns"], comp["nChar"], _i(comp["patterns"]), _i(comp["patternCounts"]),
                  comp["nEquates"], _i(comp["equates"]), _i(comp["globalInvarSitesVec"]), _i(comp["globalInvarSitesArray"]))
        M = _Model(dim, nCat, len(mp.comps), len(mp.rMatrices), len(mp.gdasrvs), _d(comps), _d(bigR), _d(rates),
                   float(mp.pInvar.val), float(mp.relRate))
        nPat = comp["nPatterns"]
        cl = P = pl = None
        if want_arrays:
            cl = np.zeros((nN, nCat, dim, nPat))
            P = np.zeros((nN, nCat, dim, dim))
            pl = np.zeros(nPat)
        fn = L.pfport_part_loglike_ld if long_double else L.pfport_part_loglike
        v = fn(C.byref(T), C.byref(D), C.byref(M), _d(cl) if want_arrays else None,
                                  _d(P) if want_arrays else None, _d(pl) if want_arrays else None)
        partLikes.append(v)
        total += v
        extra.append({"cl": cl, "P": P, "patLikes": pl, "compress": comp, "rates": rates})
    if want_arrays:
        return total, partLikes, extra
    return total
