"""host -- the host-side callers of the likelihood path, mirroring p4's own.

These classes issue exactly the sequences of ``pf.*`` calls that the reference's
Python layer issues around the hot path, with the same method names:

    Alignment._initParts       p4/alignment.py:5784-5877
    Data._setCStuff            p4/data.py:261-279
    Model.allocCStuff          p4/model.py:735-824
    ModelPart.setCStuff        p4/model.py:153-207
    Tree._allocCStuff          p4/tree.py:9255-9319
    Tree.setCStuff             p4/tree.py:9322-9377
    Tree._commonCStuff         p4/tree.py:9379-9404
    Tree.calcLogLike           p4/tree.py:9406-9415
    Tree.optLogLike            p4/tree.py:9417-9499
    Tree.getSiteLikes          p4/tree.py:9679-9700
    Chain.proposeSp dirty path p4/chain.py:668-688
    Chain.__init__ / gen state transfer   p4/chain.py:43-46, 1497-1531

They are engine-agnostic: every object takes the ``pf`` module to drive, so
the same code runs this repository's CUDA engine (``p4_phylogenetics_b200.pf``)
and, in tests, the reference's own Pf engine (``oracle/_ref``), which is how
parity is checked call for call.  They hold no likelihood arithmetic.
"""
import numpy as np

NO_ORDER = -10000   # p4/var.py:109

DNA_SYMBOLS = "acgt"
DNA_EQUATES = {"n": "acgt", "m": "ac", "k": "gt", "h": "act", "y": "ct", "v": "acg",
               "w": "at", "d": "agt", "b": "cgt", "r": "ag", "s": "cg"}           # p4/alignment.py:431-434
PROTEIN_SYMBOLS = "arndcqeghilkmfpstwyv"
PROTEIN_EQUATES = {"b": "dn", "x": "arndcqeghilkmfpstwyv", "z": "eq"}               # p4/alignment.py:439-440

RMATRIX_PROTEIN_SPEC = {"cpREV": 101, "d78": 102, "jtt": 103, "mtREV24": 104, "mtmam": 105, "wag": 106,
                        "blosum62": 107, "rtRev": 110, "tmjtt94": 111, "tmlg99": 112, "lg": 113, "hivb": 114,
                        "mtart": 115, "mtzoa": 116, "gcpREV": 117, "stmtREV": 118, "vt": 119, "pmb": 120}   # p4/var.py:245-263


class Var:
    """The numeric limits p4 shares with C as 1-element arrays (p4/var.py:287-305)."""

    def __init__(self):
        f = lambda v: np.array([v], dtype=np.float64)
        self._rMatrixNormalizeTo1 = np.array([1], np.int32)
        self._PIVEC_MIN, self._PIVEC_MAX = f(1.0e-13), f(0.999)
        self._RATE_MIN, self._RATE_MAX = f(1.0e-13), f(0.9999999)
        self._GAMMA_SHAPE_MIN, self._GAMMA_SHAPE_MAX = f(0.1), f(300.0)
        self._PINVAR_MIN, self._PINVAR_MAX = f(0.0), f(0.99)
        self._RELRATE_MIN, self._RELRATE_MAX = f(1.0e-8), f(1.0e8)
        self._KAPPA_MIN, self._KAPPA_MAX = f(0.000001), f(100.0)
        self._BRLEN_MIN, self._BRLEN_MAX = f(1.0e-8), f(3.0)
        self._newtAndBrentPowellOptPassLimit = np.array([50], dtype=np.int32)


var = Var()


# ------------------------------------------------------------------------------
# Data
# ------------------------------------------------------------------------------
class Part:
    def __init__(self):
        self.cPart = None
        self.nTax = self.nChar = self.dim = 0
        self.symbols = ""
        self.equates = {}


class Alignment:
    """Sequences of one datatype; ``_initParts`` builds the C part (one part per alignment)."""

    def resetSequencesFromParts(self):
        """Alignment.resetSequencesFromParts() (p4/alignment.py:5930-5970, one part per alignment here): bring the part's
        sequences -- e.g. freshly simulated ones -- back as strings."""
        p = self.parts[0]
        allSeq = self.pf.symbolSequences(p.cPart)
        asBytes = isinstance(self.sequences[0], (bytes, bytearray))
        self.sequences = [(allSeq[i * p.nChar:(i + 1) * p.nChar].encode("latin-1") if asBytes else allSeq[i * p.nChar:(i + 1) * p.nChar])
                          for i in range(p.nTax)]

    def composition(self, sequenceNumberList=None):
        """Part.composition() (p4/part.py:41-70): composition over the chosen sequences (all by default), ambiguities shared out."""
        p = self.parts[0]
        chosen = set(range(p.nTax)) if sequenceNumberList is None else set(sequenceNumberList)
        for i in range(p.nTax):
            self.pf.pokePartTaxListAtIndex(p.cPart, 1 if i in chosen else 0, i)
        return self.pf.partComposition(p.cPart)

    def __init__(self, pf, sequences, symbols, equates):
        self.pf = pf
        self.sequences = sequences       # list of str or bytes, all the same length
        self.symbols = symbols
        self.equates = equates
        self.dim = len(symbols)
        self.length = len(sequences[0])
        self.parts = []

    def _initParts(self):
        pf = self.pf
        eqSymb = "".join(sorted(self.equates.keys()))
        p = Part()
        p.dim, p.symbols, p.equates = self.dim, self.symbols, self.equates
        p.nTax, p.nChar = len(self.sequences), self.length
        p.cPart = pf.newPart(p.nTax, p.nChar, eqSymb, len(eqSymb), self.symbols, len(self.symbols))
        table = []
        for e in eqSymb:
            for s in self.symbols:
                table.append("1" if s in self.equates[e] else "0")
        pf.pokeEquatesTable(p.cPart, "".join(table))
        if isinstance(self.sequences[0], (bytes, bytearray)):
            joined = b"".join(self.sequences)
            if not getattr(pf, "ACCEPTS_BYTES", False):   # the reference's wrapper parses "s": it wants str
                joined = joined.decode("latin-1")
            pf.pokeSequences(p.cPart, joined)
        else:
            pf.pokeSequences(p.cPart, "".join(self.sequences))
        pf.makePatterns(p.cPart)
        pf.setGlobalInvarSitesVec(p.cPart)
        self.parts = [p]
        return p


class Data:
    """A list of parts, one per alignment (p4/data.py)."""

    def __init__(self, pf, alignments):
        self.pf = pf
        self.alignments = alignments
        self.parts = []
        for a in alignments:
            if pf is not None and not a.parts:      # pf None: inputs only (the oracle port reads the strings itself)
                a._initParts()
            self.parts.extend(a.parts)
        self.nParts = len(alignments)
        self.nTax = len(alignments[0].sequences)
        self.cData = None

    def _setCStuff(self):
        assert self.cData is None
        self.cData = self.pf.newData(self.nTax, self.nParts)
        for i, p in enumerate(self.parts):
            self.pf.pokePartInData(p.cPart, self.cData, i)

    def free(self):
        for p in self.parts:
            if p.cPart:
                self.pf.freePart(p.cPart)
                p.cPart = None
        if self.cData:
            self.pf.freeData(self.cData)
            self.cData = None


# ------------------------------------------------------------------------------
# Model
# ------------------------------------------------------------------------------
class Comp:
    def __init__(self, val, free=0):
        self.val = np.array(val, dtype=np.float64)
        self.free = free


class RMatrix:
    """spec: 'ones', 'specified', '2p' or a protein matrix name (p4/model.py:792-810)."""

    def __init__(self, spec="ones", val=None, free=0):
        self.spec = spec
        self.val = None if val is None else np.array(val, dtype=np.float64)
        self.free = free


class Gdasrv:
    def __init__(self, nGammaCat, val, free=0):
        self.nGammaCat = nGammaCat
        self.val = np.array([val], dtype=np.float64)
        self.freqs = np.zeros(nGammaCat, dtype=np.float64)
        self.rates = np.zeros(nGammaCat, dtype=np.float64)
        self.free = free
        self.c = None


class PInvar:
    def __init__(self, val=0.0, free=0):
        self.val = val
        self.free = free


class ModelPart:
    def __init__(self, num, dim, nGammaCat=1):
        self.num = num
        self.dim = dim
        self.comps, self.rMatrices, self.gdasrvs = [], [], []
        self.nGammaCat = nGammaCat
        self.pInvar = PInvar()
        self.relRate = 1.0
        self.isHet = 0
        self.bQETneedsReset = None

    nComps = property(lambda s: len(s.comps))
    nRMatrices = property(lambda s: len(s.rMatrices))
    nGdasrvs = property(lambda s: len(s.gdasrvs))

    def setCStuff(self, model):
        pf = model.pf
        for mt in self.comps:
            assert np.min(mt.val) >= var._PIVEC_MIN[0]
        for mNum, mt in enumerate(self.rMatrices):
            if mt.spec == "2p":
                pf.p4_setKappa(model.cModel, self.num, mNum, float(mt.val[0]))
            elif mt.free or mt.spec == "specified":
                k = 0
                lim = self.dim - 1 if var._rMatrixNormalizeTo1[0] else self.dim - 2
                for i in range(lim):
                    for j in range(i + 1, self.dim):
                        pf.p4_setRMatrixBigR(model.cModel, self.num, mNum, i, j, float(mt.val[k]))
                        k += 1
        pf.p4_setPInvarVal(model.cModel, self.num, float(self.pInvar.val))
        pf.p4_setRelRateVal(model.cModel, self.num, float(self.relRate))


class Model:
    def __init__(self, pf, parts):
        self.pf = pf
        self.parts = parts
        self.nParts = len(parts)
        self.doRelRates = 0
        self.relRatesAreFree = 0
        self.nFreePrams = 0
        self.isHet = int(any(p.isHet for p in parts))
        self.cModel = None

    def allocCStuff(self):
        pf = self.pf
        assert self.cModel is None
        self.cModel = pf.p4_newModel(self.nParts, self.doRelRates, self.relRatesAreFree, int(self.nFreePrams), self.isHet,
                                     var._rMatrixNormalizeTo1, var._PINVAR_MIN, var._PINVAR_MAX, var._KAPPA_MIN, var._KAPPA_MAX,
                                     var._GAMMA_SHAPE_MIN, var._GAMMA_SHAPE_MAX, var._PIVEC_MIN, var._PIVEC_MAX,
                                     var._RATE_MIN, var._RATE_MAX, var._RELRATE_MIN, var._RELRATE_MAX,
                                     var._BRLEN_MIN, var._BRLEN_MAX)
        for pNum, mp in enumerate(self.parts):
            mp.bQETneedsReset = np.ones((mp.nComps, mp.nRMatrices), np.int32)
            pf.p4_newModelPart(self.cModel, pNum, mp.dim, mp.nComps, mp.nRMatrices, mp.nGdasrvs, mp.nGammaCat,
                               mp.pInvar.free, mp.bQETneedsReset)
            for mNum, mt in enumerate(mp.comps):
                pf.p4_newComp(self.cModel, pNum, mNum, mt.free, mt.val)
            for mNum, mt in enumerate(mp.rMatrices):
                if mt.spec == "ones":
                    spec = 100
                elif mt.spec in ("specified", "optimized"):
                    spec = 20
                elif mt.spec == "2p":
                    spec = 5
                else:
                    spec = RMATRIX_PROTEIN_SPEC[mt.spec]
                pf.p4_newRMatrix(self.cModel, pNum, mNum, mt.free, spec)
            for mNum, mt in enumerate(mp.gdasrvs):
                mt.c = pf.p4_newGdasrv(self.cModel, pNum, mNum, mt.nGammaCat, mt.free, mt.val, mt.freqs, mt.rates)
                pf.gdasrvCalcRates(mt.c)     # Gdasrv.calcRates(), p4/model.py

    def setCStuff(self, partNum=None):
        for pNum, mp in enumerate(self.parts):
            if partNum is None or pNum == partNum:
                mp.setCStuff(self)

    def restoreFreePrams(self, prams):
        """After an optimisation: copy the optimised values back (p4/model.py:835-960).  Comps and
        gdasrv shapes are shared arrays the engine already updated in place."""
        pos = 0
        for mp in self.parts:
            for mt in mp.comps:
                if mt.free:
                    pos += mp.dim - 1
            for mt in mp.rMatrices:
                if mt.free:
                    if mt.spec == "2p":
                        mt.val[0] = prams[pos]
                        pos += 1
                    else:
                        mt.spec = "optimized"
                        n = mp.dim * (mp.dim - 1) // 2
                        if mt.val is None or len(mt.val) != n:
                            mt.val = np.zeros(n)
                        for k in range(n - 1):
                            mt.val[k] = prams[pos]
                            pos += 1
                        mt.val[-1] = 1.0 - float(np.sum(mt.val[:-1])) if var._rMatrixNormalizeTo1[0] else 1.0
            for mt in mp.gdasrvs:
                if mt.free:
                    mt.val[0] = prams[pos]
                    pos += 1
            if mp.pInvar.free:
                mp.pInvar.val = prams[pos]
                pos += 1
        if self.relRatesAreFree:
            for mp in self.parts[:-1]:
                mp.relRate = prams[pos]
                pos += 1

    def free(self):
        if self.cModel:
            self.pf.p4_freeModel(self.cModel)
            self.cModel = None


# ------------------------------------------------------------------------------
# Tree
# ------------------------------------------------------------------------------
class NodePart:
    def __init__(self):
        self.compNum = 0


class BranchPart:
    def __init__(self):
        self.rMatrixNum = 0
        self.gdasrvNum = 0


class Branch:
    def __init__(self):
        self.len = 0.1
        self.lenChanged = False
        self.parts = []


class Node:
    def __init__(self, nodeNum):
        self.nodeNum = nodeNum
        self.parent = self.leftChild = self.sibling = None
        self.isLeaf = 0
        self.seqNum = -1
        self.br = Branch()
        self.parts = []
        self.cNode = None
        self.flag = 0

    def iterChildren(self):
        c = self.leftChild
        while c:
            yield c
            c = c.sibling


class Tree:
    def __init__(self, pf, nodes, root):
        self.pf = pf
        self.nodes = nodes
        self.root = root
        self.data = None
        self.model = None
        self.cTree = None
        self.partLikes = None
        self.logLike = None
        self.preOrder = np.full(len(nodes), NO_ORDER, dtype=np.int32)
        self.postOrder = np.full(len(nodes), NO_ORDER, dtype=np.int32)
        self.preAndPostOrderAreValid = False
        self.bulkSetCStuff = False     # True: setCStuff uses pf.setTreeCStuff where the engine offers it

    # -- traversal (p4/tree.py setPreAndPostOrder writes the arrays in place) --
    def setPreAndPostOrder(self):
        pre, post = [], []
        stack = [(self.root, False)]
        while stack:
            n, done = stack.pop()
            if done:
                post.append(n.nodeNum)
                continue
            pre.append(n.nodeNum)
            stack.append((n, True))
            for c in reversed(list(n.iterChildren())):
                stack.append((c, False))
        self.preOrder[:] = NO_ORDER
        self.postOrder[:] = NO_ORDER
        self.preOrder[:len(pre)] = pre
        self.postOrder[:len(post)] = post
        self.preAndPostOrderAreValid = True

    def iterNodes(self):
        for n in self.nodes:
            if n.nodeNum != NO_ORDER:
                yield n

    def iterNodesNoRoot(self):
        for n in self.iterNodes():
            if n is not self.root:
                yield n

    def iterInternalsPostOrder(self):
        for i in self.postOrder:
            if i != NO_ORDER and not self.nodes[i].isLeaf:
                yield self.nodes[i]

    def attach(self, data, model):
        self.data, self.model = data, model
        for n in self.nodes:
            n.parts = [NodePart() for _ in range(model.nParts)]
            n.br.parts = [BranchPart() for _ in range(model.nParts)]

    # -- C glue -------------------------------------------------------------------
    def _allocCStuff(self):
        pf = self.pf
        if not self.preAndPostOrderAreValid:
            self.setPreAndPostOrder()
        if not self.data.cData:
            self.data._setCStuff()
        if not self.model.cModel:
            self.model.allocCStuff()
        nLeaves = sum(1 for n in self.iterNodes() if n.isLeaf)
        self.partLikes = np.zeros(self.model.nParts, dtype=np.float64)
        self.cTree = pf.p4_newTree(len(list(self.iterNodes())), nLeaves, self.preOrder, self.postOrder,
                                   var._newtAndBrentPowellOptPassLimit, self.partLikes, self.data.cData, self.model.cModel)
        inTree = set(int(i) for i in self.preOrder)
        for i, n in enumerate(self.nodes):
            if n.nodeNum == NO_ORDER:
                continue
            n.cNode = pf.p4_newNode(n.nodeNum, self.cTree, n.seqNum, n.isLeaf, 1 if i in inTree else 0)

    def setCStuff(self):
        pf = self.pf
        if self.bulkSetCStuff and hasattr(pf, "setTreeCStuff") and all(n.nodeNum != NO_ORDER for n in self.nodes):
            # the engine's one-call form of the loops below (include/p4b200.h p4b_setTreeCStuff)
            nn = len(self.nodes)
            rel = np.full((3, nn), -1, dtype=np.int32)
            brl = np.zeros(nn, dtype=np.float64)
            for n in self.nodes:
                i = n.nodeNum
                if n.parent:
                    rel[0, i] = n.parent.nodeNum
                    brl[i] = n.br.len
                if n.leftChild:
                    rel[1, i] = n.leftChild.nodeNum
                if n.sibling:
                    rel[2, i] = n.sibling.nodeNum
            pf.setTreeCStuff(self.cTree, rel[0], rel[1], rel[2], brl, self.root.nodeNum)
        else:
            for n in self.iterNodes():
                pf.p4_setNodeRelation(n.cNode, 0, n.parent.nodeNum if n.parent else -1)
                pf.p4_setNodeRelation(n.cNode, 1, n.leftChild.nodeNum if n.leftChild else -1)
                pf.p4_setNodeRelation(n.cNode, 2, n.sibling.nodeNum if n.sibling else -1)
            pf.p4_setTreeRoot(self.cTree, self.root.cNode)
            for n in self.iterNodesNoRoot():
                pf.p4_setBrLen(n.cNode, n.br.len)
        if self.model.isHet:
            for pNum in range(self.model.nParts):
                if self.model.parts[pNum].isHet:
                    for n in self.iterNodes():
                        pf.p4_setCompNum(n.cNode, pNum, n.parts[pNum].compNum)
                        if n is not self.root:
                            pf.p4_setRMatrixNum(n.cNode, pNum, n.br.parts[pNum].rMatrixNum)
                            pf.p4_setGdasrvNum(n.cNode, pNum, n.br.parts[pNum].gdasrvNum)
        if not self.preAndPostOrderAreValid:
            self.setPreAndPostOrder()

    def _commonCStuff(self):
        if not self.cTree:
            self._allocCStuff()
        self.model.setCStuff()
        self.setCStuff()
        self.pf.p4_setPrams(self.cTree, -1)

    def calcLogLike(self, verbose=0):
        self._commonCStuff()
        self.logLike = self.pf.p4_treeLogLike(self.cTree, 0)
        if verbose:
            print("Tree.calcLogLike(). %f" % self.logLike)
        return self.logLike

    def optLogLike(self, verbose=0, method="BOBYQA", optBrLens=True):
        """Maximise the likelihood over the free parameters (p4/tree.py:9417-9499)."""
        pf = self.pf
        self._commonCStuff()
        if method == "BOBYQA":
            pf.p4_allBOBYQAOptimize(self.cTree, 1 if optBrLens else 0)
        elif method == "allBrentPowell":
            pf.p4_allBrentPowellOptimize(self.cTree)
        elif method == "newtAndBrentPowell":
            pf.p4_newtSetup(self.cTree)
            pf.p4_newtAndBrentPowellOpt(self.cTree)
        elif method == "newtAndBOBYQA":
            pf.p4_newtSetup(self.cTree)
            pf.p4_newtAndBOBYQAOpt(self.cTree)
        else:
            raise ValueError('method should be one of "newtAndBrentPowell", "allBrentPowell", "newtAndBOBYQA", or "BOBYQA"')
        self.logLike = pf.p4_treeLogLike(self.cTree, 0)
        brLens = pf.p4_getBrLens(self.cTree)
        for n in self.iterNodesNoRoot():
            n.br.len = brLens[n.nodeNum]
        self.model.restoreFreePrams(pf.p4_getFreePrams(self.cTree))
        if verbose:
            print("optLogLike = %f" % self.logLike)
        return self.logLike

    def simulate(self, seed=None, calculatePatterns=True, refTree=None, resetSequences=True):
        """Tree.simulate() (p4/tree.py:9527-9637): new data down this tree with its model, into its own data parts; with a
        refTree (same tree and model, its own data, likelihood calculated) the root states come from its posterior.
        ``seed`` (an addition) re-seeds the stream kept on the tree, p4 keeps it in var.gsl_rng."""
        pf = self.pf
        if getattr(self, "gsl_rng", None) is None:
            self.gsl_rng = pf.gsl_rng_get()
            if seed is None:
                import time
                pf.gsl_rng_set(self.gsl_rng, int(time.time()))
        if seed is not None:
            pf.gsl_rng_set(self.gsl_rng, int(seed))
        self._commonCStuff()
        if refTree is not None:
            if not refTree.cTree:
                refTree.calcLogLike()
            assert refTree.data.cData != self.data.cData
        pf.p4_simulate(self.cTree, refTree.cTree if refTree is not None else 0, self.gsl_rng)
        if calculatePatterns:
            for p in self.data.parts:
                pf.makePatterns(p.cPart)
                pf.setGlobalInvarSitesVec(p.cPart)
        if resetSequences:                       # Data.resetSequencesFromParts, p4/tree.py:9627-9628
            for a in self.data.alignments:
                a.resetSequencesFromParts()

    def ancestralStateDraw(self):
        """Tree.ancestralStateDraw() (p4/tree.py:9640-9677): one draw of the root's state at every site, as a string."""
        pf = self.pf
        self._commonCStuff()
        self.logLike = pf.p4_treeLogLike(self.cTree, 0)
        draw = np.empty(4, dtype=np.int32)
        out = []
        for pNum, dp in enumerate(self.data.parts):
            for seqPos in range(dp.nChar):
                pf.p4_drawAncState(self.cTree, pNum, seqPos, draw)
                if draw[1] >= 0:
                    out.append(dp.symbols[draw[0]])
                elif draw[2]:
                    out.append(dp.symbols[draw[3]])
                else:
                    raise RuntimeError("Tree.ancestralStateDraw(). Problem with returned draw.  Got %s" % draw)
        return "".join(out)

    def getSiteLikes(self):
        self._commonCStuff()
        self.logLike = self.pf.p4_treeLogLike(self.cTree, 1)
        self.siteLikes = []
        for p in self.data.parts:
            self.siteLikes += self.pf.getSiteLikes(p.cPart)
        return self.siteLikes     # (the reference returns None and leaves the list in self.siteLikes)

    # -- the dirty path of Chain.proposeSp (p4/chain.py:668-688) -------------------
    def recalcAfterBranchChange(self):
        """Recompute P for branches with ``br.lenChanged`` and the CLs from there to the root."""
        pf = self.pf
        self.setCStuff()
        for n in self.iterNodesNoRoot():
            if n.br.lenChanged:
                pf.p4_calculateBigPDecks(n.cNode)
                p = n
                while p.parent:
                    p = p.parent
                    p.flag = 1
                n.br.lenChanged = False
        for pNum in range(self.model.nParts):
            for n in self.iterInternalsPostOrder():
                if n.flag:
                    pf.p4_setConditionalLikelihoodsOfInternalNodePart(n.cNode, pNum)
        for n in self.iterNodes():
            n.flag = 0
        for pNum in range(self.model.nParts):
            pf.p4_partLogLike(self.cTree, self.data.parts[pNum].cPart, pNum, 0)
        self.logLike = float(sum(self.partLikes))
        return self.logLike

    def deleteCStuff(self):
        pf = self.pf
        if self.cTree:
            for n in self.nodes:     # nodes, then tree, then model (p4/tree.py:9202-9253)
                if n.cNode:
                    pf.p4_freeNode(n.cNode)
                    n.cNode = None
            pf.p4_freeTree(self.cTree)
            self.cTree = None

    def dupe(self):
        """Same topology, branch lengths and model usage, new Node objects, no C stuff."""
        nn = [Node(n.nodeNum) for n in self.nodes]
        for a, b in zip(self.nodes, nn):
            b.isLeaf, b.seqNum = a.isLeaf, a.seqNum
            b.br.len = a.br.len
            b.parent = nn[a.parent.nodeNum] if a.parent else None
            b.leftChild = nn[a.leftChild.nodeNum] if a.leftChild else None
            b.sibling = nn[a.sibling.nodeNum] if a.sibling else None
        t = Tree(self.pf, nn, nn[self.root.nodeNum])
        t.bulkSetCStuff = self.bulkSetCStuff
        if self.model is not None:
            for a, b in zip(self.nodes, nn):
                b.parts = [NodePart() for _ in a.parts]
                b.br.parts = [BranchPart() for _ in a.br.parts]
                for x, y in zip(a.parts, b.parts):
                    y.compNum = x.compNum
                for x, y in zip(a.br.parts, b.br.parts):
                    y.rMatrixNum, y.gdasrvNum = x.rMatrixNum, x.gdasrvNum
        return t


# ------------------------------------------------------------------------------
# Twins: the same inputs handed to another engine (parity tests, cur/prop trees)
# ------------------------------------------------------------------------------
def clone_model(model, pf):
    parts = []
    for mp in model.parts:
        q = ModelPart(mp.num, mp.dim, mp.nGammaCat)
        q.comps = [Comp(c.val.copy(), c.free) for c in mp.comps]
        q.rMatrices = [RMatrix(r.spec, None if r.val is None else r.val.copy(), r.free) for r in mp.rMatrices]
        q.gdasrvs = [Gdasrv(g.nGammaCat, float(g.val[0]), g.free) for g in mp.gdasrvs]
        q.pInvar = PInvar(mp.pInvar.val, mp.pInvar.free)
        q.relRate = mp.relRate
        q.isHet = mp.isHet
        parts.append(q)
    m = Model(pf, parts)
    m.doRelRates, m.relRatesAreFree, m.nFreePrams = model.doRelRates, model.relRatesAreFree, model.nFreePrams
    return m


def clone_data(data, pf):
    return Data(pf, [Alignment(pf, a.sequences, a.symbols, a.equates) for a in data.alignments])


def clone_tree(tree, pf, data=None):
    """The same tree, data and model driven through another ``pf`` (or sharing ``data``)."""
    t = tree.dupe()
    t.pf = pf
    t.data = data if data is not None else clone_data(tree.data, pf)
    t.model = clone_model(tree.model, pf)
    return t
