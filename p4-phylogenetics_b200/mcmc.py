"""mcmc -- a host-side driver that issues the ``pf.*`` call sequences of p4's MCMC.

p4's ``Mcmc.run`` (p4/mcmc.py:2830-2975) runs, per generation, every chain's
``Chain.gen`` (p4/chain.py:1138-1536): propose on the chain's *prop* tree, send
the change to C, recompute what the change made stale (``Chain.proposeSp``,
p4/chain.py:305-1000), read the part log-likelihoods, accept or reject, and
copy the winning tree's numerical state over the loser's (``p4_copyCondLikes``,
``p4_copyBigPDecks``, ``p4_copyModelPrams``, p4/chain.py:1497-1531); then one
swap between two Metropolis-coupled chains is proposed (p4/mcmc.py:3299-3353).

This module restates that *call protocol* for a representative proposal set so
the engine can be driven -- and measured -- the way p4 drives it:

    brLen, allBrLens            p4/chain.py:694-741   one / all branch lengths
    local                       p4/chain.py:609-690   three branch lengths + an NNI (stand-in for LOCAL)
    eTBR                        p4/chain.py:743-846   prune-and-regraft (stand-in for the extending TBR)
    allCompsDir, allRMatricesDir, gdasrv, pInvar   p4/chain.py:305-380   p4_setPrams + whole part
    relRate                     p4/chain.py:974-990

The proposal *distributions* of ``local`` and ``eTBR`` are simpler than the
reference's (they are host-side tree surgery, not on the likelihood path); what
is kept exactly is which pf calls follow: ``setCStuff``, ``p4_calculateBigPDecks``
for every changed branch, ``p4_setConditionalLikelihoodsOfInternalNodePart`` for
the flagged internals in post-order, ``p4_partLogLike``.

It is engine-agnostic like ``host``: with the reference's own Pf engine
(``oracle/_ref``) it reproduces the same chain; tests compare the two
trajectories.  With this repository's engine ``Mcmc.run(batched=True)``
issues the proposals of all chains first and reads all prop-tree
log-likelihoods through ONE ``pf.treesPartLogLike`` call per part (north_star:
"several Metropolis-coupled chains are batched per launch").
"""
import math

import numpy as np
from scipy.special import gammaln

from . import host

var = host.var


class Proposal:
    def __init__(self, name, weight, tuning=None, pNum=-1):
        self.name = name
        self.weight = float(weight)
        self.tuning = tuning
        self.pNum = pNum
        self.nProposals = 0
        self.nAcceptances = 0


def _children(n):
    return list(n.iterChildren())


def _set_children(parent, kids):
    parent.leftChild = kids[0] if kids else None
    for a, b in zip(kids, kids[1:]):
        a.sibling = b
    if kids:
        kids[-1].sibling = None
    for k in kids:
        k.parent = parent


def _subtree_nums(n):
    out, stack = set(), [n]
    while stack:
        x = stack.pop()
        out.add(x.nodeNum)
        stack.extend(_children(x))
    return out


def _log_dirichlet_pdf(x, alpha):
    return float(gammaln(np.sum(alpha)) - np.sum(gammaln(alpha)) + np.sum((alpha - 1.0) * np.log(x)))


def copy_tree_to(a, b):
    """Tree.copyToTree (p4/tree.py:2513-2537): relations, branch lengths, traversal orders, model usage."""
    for na, nb in zip(a.nodes, b.nodes):
        nb.parent = b.nodes[na.parent.nodeNum] if na.parent else None
        nb.leftChild = b.nodes[na.leftChild.nodeNum] if na.leftChild else None
        nb.sibling = b.nodes[na.sibling.nodeNum] if na.sibling else None
        nb.br.len = na.br.len
        nb.br.lenChanged = False
        nb.flag = 0
        for x, y in zip(na.parts, nb.parts):
            y.compNum = x.compNum
        for x, y in zip(na.br.parts, nb.br.parts):
            y.rMatrixNum, y.gdasrvNum = x.rMatrixNum, x.gdasrvNum
    b.root = b.nodes[a.root.nodeNum]
    b.preOrder[:] = a.preOrder          # in place: the engine borrowed these buffers (Pf/p4_tree.c:46-47)
    b.postOrder[:] = a.postOrder
    b.preAndPostOrderAreValid = True


def copy_model_part_vals(a, b):
    """ModelPart.copyValsTo (p4/model.py): in place where the engine borrowed the buffer."""
    for x, y in zip(a.comps, b.comps):
        y.val[:] = x.val
    for x, y in zip(a.rMatrices, b.rMatrices):
        if x.val is not None:
            if y.val is None or len(y.val) != len(x.val):
                y.val = x.val.copy()
            else:
                y.val[:] = x.val
        y.spec = x.spec
    for x, y in zip(a.gdasrvs, b.gdasrvs):
        y.val[:] = x.val
        y.freqs[:] = x.freqs
        y.rates[:] = x.rates
    b.pInvar.val = a.pInvar.val
    b.relRate = a.relRate


class Chain:
    """One chain: a cur tree and a prop tree sharing the data (p4/chain.py:13-71)."""

    def __init__(self, mcmc, tempNum):
        self.mcmc = mcmc
        self.pf = pf = mcmc.pf
        self.tempNum = tempNum
        t = mcmc.tree
        self.curTree = host.clone_tree(t, pf, data=t.data)
        self.propTree = host.clone_tree(t, pf, data=t.data)
        self.curTree.calcLogLike()
        self.propTree.calcLogLike()
        pf.p4_copyCondLikes(self.curTree.cTree, self.propTree.cTree, 1)
        pf.p4_copyBigPDecks(self.curTree.cTree, self.propTree.cTree, 1)
        pf.p4_copyModelPrams(self.curTree.cTree, self.propTree.cTree)
        self.propTree.calcLogLike()
        if pf.p4_verifyIdentityOfTwoTrees(self.curTree.cTree, self.propTree.cTree) != 0:
            raise RuntimeError("Chain.__init__: the prop tree should be identical to the cur tree, and it is not")
        self.logProposalRatio = 0.0
        self.logPriorRatio = 0.0
        self.parts_to_read = []

    # ---- proposals: change the prop tree / model on the host ---------------------
    def _reflect(self, v, lo, hi):
        while v < lo or v > hi:
            if v < lo:
                v = 2.0 * lo - v
            if v > hi:
                v = 2.0 * hi - v
        return v

    def _new_len(self, old, tuning):
        rng = self.mcmc.rng
        new = old * math.exp(tuning * (rng.random() - 0.5))
        lo, hi = var._BRLEN_MIN[0], var._BRLEN_MAX[0]
        if new < lo or new > hi:          # reflect in log space keeps the multiplier's Hastings ratio
            new = math.exp(self._reflect(math.log(new), math.log(lo), math.log(hi)))
        return new

    def _change_len(self, n, new):
        lam = self.mcmc.brLenPriorLambda
        self.logProposalRatio += math.log(new / n.br.len)
        self.logPriorRatio += -lam * (new - n.br.len)
        n.br.len = new
        n.br.lenChanged = True

    def proposeBrLen(self, p):
        t = self.propTree
        cand = [n for n in t.iterNodesNoRoot()]
        n = cand[int(self.mcmc.rng.integers(len(cand)))]
        self._change_len(n, self._new_len(n.br.len, p.tuning))

    def proposeAllBrLens(self, p):
        for n in self.propTree.iterNodesNoRoot():
            self._change_len(n, self._new_len(n.br.len, p.tuning))

    def proposeLocal(self, p):
        """Three branch lengths around an internal edge scaled by one factor and, half of the time, a
        nearest-neighbour interchange across it."""
        t, rng = self.propTree, self.mcmc.rng
        cand = [n for n in t.iterNodesNoRoot() if not n.isLeaf]
        v = cand[int(rng.integers(len(cand)))]
        u = v.parent
        kids = _children(v)
        c = kids[int(rng.integers(len(kids)))]
        sibs = [x for x in _children(u) if x is not v]
        s = sibs[int(rng.integers(len(sibs)))]
        m = math.exp(p.tuning * (rng.random() - 0.5))
        lo, hi = var._BRLEN_MIN[0], var._BRLEN_MAX[0]
        for n in (v, c, s):
            new = n.br.len * m
            if new < lo or new > hi:
                new = min(max(new, lo), hi)
            self._change_len(n, new)
        if rng.random() < 0.5:
            uk, vk = _children(u), _children(v)
            uk[uk.index(s)] = c
            vk[vk.index(c)] = s
            _set_children(u, uk)
            _set_children(v, vk)
            t.preAndPostOrderAreValid = False
            p.topologyChanged = True

    def proposeETBR(self, p):
        """Prune the subtree below a branch and regraft it on another branch, at a uniform point."""
        t, rng = self.propTree, self.mcmc.rng
        cand = [n for n in t.iterNodesNoRoot() if n.parent is not t.root and len(_children(n.parent)) == 2]
        if not cand:
            return self.proposeBrLen(p)
        x = cand[int(rng.integers(len(cand)))]
        pn = x.parent
        g = pn.parent
        y = [k for k in _children(pn) if k is not x][0]
        inside = _subtree_nums(x)
        targets = [n for n in t.iterNodesNoRoot() if n.nodeNum not in inside and n is not pn and n is not y]
        if not targets:
            return self.proposeBrLen(p)
        z = targets[int(rng.integers(len(targets)))]
        lo = var._BRLEN_MIN[0]
        merged = min(y.br.len + pn.br.len, var._BRLEN_MAX[0])
        # take pn (with x) out: y takes its place under g
        gk = _children(g)
        gk[gk.index(pn)] = y
        _set_children(g, gk)
        # put pn on the branch above z
        w = z.parent
        wk = _children(w)
        wk[wk.index(z)] = pn
        _set_children(w, wk)
        _set_children(pn, [x, z])
        Lz = z.br.len
        r = rng.random()
        up, down = max(Lz * r, lo), max(Lz * (1.0 - r), lo)
        self.logProposalRatio += math.log(Lz / merged)
        lam = self.mcmc.brLenPriorLambda
        self.logPriorRatio += -lam * ((merged + up + down) - (y.br.len + pn.br.len + Lz))
        y.br.len, pn.br.len, z.br.len = merged, up, down
        y.br.lenChanged = pn.br.lenChanged = z.br.lenChanged = True
        t.preAndPostOrderAreValid = False
        p.topologyChanged = True

    def _dirichlet_move(self, old, tuning, floor):
        rng = self.mcmc.rng
        for _ in range(100):
            new = rng.dirichlet(tuning * old)
            if new.min() > floor:
                break
        else:
            new = old.copy()
        new = new / new.sum()
        i = int(np.argmax(new))
        new[i] += 1.0 - new.sum()
        self.logProposalRatio += _log_dirichlet_pdf(old, tuning * new) - _log_dirichlet_pdf(new, tuning * old)
        return new

    def proposeAllCompsDir(self, p):
        mp = self.propTree.model.parts[p.pNum]
        for c in mp.comps:
            c.val[:] = self._dirichlet_move(c.val.copy(), p.tuning, 10.0 * var._PIVEC_MIN[0])   # in place: shared with C

    def proposeAllRMatricesDir(self, p):
        mp = self.propTree.model.parts[p.pNum]
        for r in mp.rMatrices:
            r.val[:] = self._dirichlet_move(r.val.copy(), p.tuning, 10.0 * var._RATE_MIN[0])

    def proposeGdasrv(self, p):
        mp = self.propTree.model.parts[p.pNum]
        g = mp.gdasrvs[0]
        old = float(g.val[0])
        new = old * math.exp(p.tuning * (self.mcmc.rng.random() - 0.5))
        lo, hi = var._GAMMA_SHAPE_MIN[0], var._GAMMA_SHAPE_MAX[0]
        if new < lo or new > hi:
            new = math.exp(self._reflect(math.log(new), math.log(lo), math.log(hi)))
        self.logProposalRatio += math.log(new / old)
        g.val[0] = new            # shared with C (Pf/p4_model.c:504)

    def proposePInvar(self, p):
        mp = self.propTree.model.parts[p.pNum]
        new = mp.pInvar.val + p.tuning * (self.mcmc.rng.random() - 0.5)
        mp.pInvar.val = self._reflect(new, var._PINVAR_MIN[0] + 1e-6, var._PINVAR_MAX[0])

    def proposeRelRate(self, p):
        mps = self.propTree.model.parts
        i = int(self.mcmc.rng.integers(len(mps)))
        old = mps[i].relRate
        new = old * math.exp(p.tuning * (self.mcmc.rng.random() - 0.5))
        self.logProposalRatio += math.log(new / old)
        mps[i].relRate = new
        # keep the weighted mean rate at 1 (p4 normalises relRates by part length)
        lens = np.array([pt.nChar for pt in self.propTree.data.parts], dtype=float)
        rates = np.array([m.relRate for m in mps])
        f = float(np.sum(lens) / np.sum(lens * rates))
        for m in mps:
            m.relRate *= f

    # ---- Chain.proposeSp: the pf calls that follow a proposal ----------------------
    def _all_cl(self, pNum):
        pf, t = self.pf, self.propTree
        for n in t.iterInternalsPostOrder():
            pf.p4_setConditionalLikelihoodsOfInternalNodePart(n.cNode, pNum)

    def proposeSp(self, p):
        """Everything of Chain.proposeSp up to, not including, the p4_partLogLike calls.  Leaves the list of
        parts whose log-likelihood must be read in ``self.parts_to_read``."""
        pf, t = self.pf, self.propTree
        self.logProposalRatio = 0.0
        self.logPriorRatio = 0.0
        p.topologyChanged = False
        allParts = list(range(t.model.nParts))
        if p.name in ("allCompsDir", "allRMatricesDir", "gdasrv", "pInvar"):
            getattr(self, {"allCompsDir": "proposeAllCompsDir", "allRMatricesDir": "proposeAllRMatricesDir",
                           "gdasrv": "proposeGdasrv", "pInvar": "proposePInvar"}[p.name])(p)
            if p.name in ("allRMatricesDir", "pInvar"):      # values that are not shared numpy buffers
                t.model.setCStuff(partNum=p.pNum)
            pf.p4_setPrams(t.cTree, p.pNum)
            self._all_cl(p.pNum)
            self.parts_to_read = [p.pNum]
        elif p.name == "relRate":
            self.proposeRelRate(p)
            for mp in t.model.parts:
                pf.p4_setRelRateVal(t.model.cModel, mp.num, mp.relRate)
            pf.p4_calculateAllBigPDecksAllParts(t.cTree)
            for pNum in allParts:
                self._all_cl(pNum)
            self.parts_to_read = allParts
        elif p.name == "allBrLens":
            self.proposeAllBrLens(p)
            t.setCStuff()
            for n in t.iterNodesNoRoot():
                pf.p4_calculateBigPDecks(n.cNode)
                n.br.lenChanged = False
            for pNum in allParts:
                self._all_cl(pNum)
            self.parts_to_read = allParts
        elif p.name in ("brLen", "local", "eTBR"):
            getattr(self, {"brLen": "proposeBrLen", "local": "proposeLocal", "eTBR": "proposeETBR"}[p.name])(p)
            if not t.preAndPostOrderAreValid:
                t.setPreAndPostOrder()
            t.setCStuff()
            for n in t.iterNodesNoRoot():
                if n.br.lenChanged:
                    pf.p4_calculateBigPDecks(n.cNode)
                    q = n
                    while q is not t.root:
                        q = q.parent
                        q.flag = 1
                    n.br.lenChanged = False
            for n in t.iterInternalsPostOrder():
                if n.flag:
                    for pNum in allParts:
                        pf.p4_setConditionalLikelihoodsOfInternalNodePart(n.cNode, pNum)
                n.flag = 0
            self.parts_to_read = allParts
        else:
            raise ValueError("Unlisted proposal.name=%s" % p.name)

    def readLikes(self):
        """The p4_partLogLike calls that end Chain.proposeSp, for this chain alone."""
        pf, t = self.pf, self.propTree
        for pNum in self.parts_to_read:
            pf.p4_partLogLike(t.cTree, t.data.parts[pNum].cPart, pNum, 0)

    # ---- Chain.gen after the likelihood is known (p4/chain.py:1283-1531) ------------
    def finish(self, p):
        pf = self.pf
        cur, prop = self.curTree, self.propTree
        prop.logLike = float(sum(prop.partLikes))
        logLikeRatio = prop.logLike - cur.logLike
        logPriorRatio = self.logPriorRatio
        if self.mcmc.nChains > 1:
            heatBeta = 1.0 / (1.0 + self.mcmc.chainTemp * self.tempNum)
            logLikeRatio *= heatBeta
            logPriorRatio *= heatBeta
        pRet = logLikeRatio + self.logProposalRatio + logPriorRatio
        if pRet < -100.0:
            r = 0.0
        elif pRet >= 0.0:
            r = 1.0
        else:
            r = math.exp(pRet)
        accept = r == 1.0 or self.mcmc.rng.random() < r
        p.nProposals += 1
        if accept:
            p.nAcceptances += 1
        a, b = (prop, cur) if accept else (cur, prop)
        b.logLike = a.logLike
        if p.name in ("allCompsDir", "allRMatricesDir", "gdasrv", "pInvar"):
            b.partLikes[p.pNum] = a.partLikes[p.pNum]
            copy_model_part_vals(a.model.parts[p.pNum], b.model.parts[p.pNum])
            if p.name not in ("allCompsDir", "gdasrv"):
                b.model.setCStuff(partNum=p.pNum)
            if not (a.model.parts[p.pNum].bQETneedsReset == b.model.parts[p.pNum].bQETneedsReset).all():
                b.model.parts[p.pNum].bQETneedsReset[:] = a.model.parts[p.pNum].bQETneedsReset
        elif p.name == "relRate":
            b.partLikes[:] = a.partLikes
            for x, y in zip(a.model.parts, b.model.parts):
                copy_model_part_vals(x, y)
        else:
            b.partLikes[:] = a.partLikes
            copy_tree_to(a, b)
            for x, y in zip(a.model.parts, b.model.parts):
                y.bQETneedsReset[:] = x.bQETneedsReset
            b.setCStuff()
        pf.p4_copyCondLikes(a.cTree, b.cTree, 1)
        pf.p4_copyBigPDecks(a.cTree, b.cTree, 1)
        pf.p4_copyModelPrams(a.cTree, b.cTree)
        return accept

    def gen(self, p):
        self.proposeSp(p)
        self.readLikes()
        return self.finish(p)

    def free(self):
        for t in (self.curTree, self.propTree):
            t.deleteCStuff()
            t.model.free()


class Mcmc:
    """``Mcmc(tree, nChains).run(nGens)``: the generation loop of p4/mcmc.py:2830-2975 with the default
    proposal mix (p4/mcmc.py:132-163) and its weights (p4/mcmc.py:1300-1900)."""

    def __init__(self, tree, nChains=1, seed=0, chainTemp=0.15, proposals=None, brLenPriorLambda=10.0):
        self.tree = tree
        self.pf = tree.pf
        self.nChains = nChains
        self.chainTemp = chainTemp
        self.brLenPriorLambda = brLenPriorLambda
        self.rng = np.random.Generator(np.random.PCG64(seed))
        self.gen = 0
        self.nSwapAttempts = self.nSwaps = 0
        if tree.cTree is None:
            tree.calcLogLike()
        self.proposals = proposals if proposals is not None else self._makeProposals()
        w = np.array([p.weight for p in self.proposals])
        self._cum = np.cumsum(w / w.sum())
        self.chains = [Chain(self, i) for i in range(nChains)]
        self.trace = []           # (gen, [cur lnL by tempNum])

    def _makeProposals(self):
        t = self.tree
        nBr = len(t.nodes) - 1
        props = [Proposal("local", nBr, 0.3), Proposal("eTBR", nBr, 0.3), Proposal("allBrLens", nBr, 0.02)]
        for pNum, mp in enumerate(t.model.parts):
            if mp.comps and mp.comps[0].free:
                props.append(Proposal("allCompsDir", (mp.dim - 1) * mp.nComps, 1000.0 * mp.dim, pNum))
            if mp.rMatrices and mp.rMatrices[0].free and mp.rMatrices[0].spec != "2p":
                props.append(Proposal("allRMatricesDir", mp.nRMatrices * ((mp.dim * mp.dim - mp.dim) / 2 - 1), 2000.0, pNum))
            if mp.gdasrvs and mp.gdasrvs[0].free:
                props.append(Proposal("gdasrv", 1.0, 0.5, pNum))
            if mp.pInvar.free:
                props.append(Proposal("pInvar", 1.0, 0.1, pNum))
        if t.model.nParts > 1 and t.model.relRatesAreFree:
            props.append(Proposal("relRate", t.model.nParts, 0.2))
        return props

    def _choose(self):
        return self.proposals[int(np.searchsorted(self._cum, self.rng.random(), side="right").clip(0, len(self.proposals) - 1))]

    def _proposeSwap(self):
        if self.nChains < 2:
            return
        t1 = int(self.rng.integers(self.nChains - 1))
        c1 = [c for c in self.chains if c.tempNum == t1][0]
        c2 = [c for c in self.chains if c.tempNum == t1 + 1][0]
        b1, b2 = 1.0 / (1.0 + self.chainTemp * t1), 1.0 / (1.0 + self.chainTemp * (t1 + 1))
        lnR = b1 * c2.curTree.logLike + b2 * c1.curTree.logLike - b1 * c1.curTree.logLike - b2 * c2.curTree.logLike
        r = 0.0 if lnR < -100.0 else (1.0 if lnR >= 0.0 else math.exp(lnR))
        self.nSwapAttempts += 1
        if self.rng.random() < r:
            self.nSwaps += 1
            c1.tempNum, c2.tempNum = c2.tempNum, c1.tempNum

    def run(self, nGens, batched=None):
        """``batched``: issue all chains' proposals, then read every prop-tree log-likelihood with one
        ``pf.treesPartLogLike`` per part (one launch for all chains).  ``batched="pipelined"``: as soon as a
        chain's proposal is issued its evaluation is started with ``pf.partLogLikeBegin`` -- the GPU works on it
        while the host prepares the next chain -- and the values are collected chain by chain, each wait on that
        evaluation's own event.  Default: pipelined when
        the engine offers it.  The random stream is consumed in the same order in every mode, so
        all modes give the same chain."""
        pf = self.pf
        if batched is None:
            batched = "pipelined" if hasattr(pf, "partLogLikeBegin") else hasattr(pf, "treesPartLogLike")
        pipelined = batched == "pipelined"     # each chain's evaluation starts as soon as its proposal is issued
        for _ in range(nGens):
            chosen = [self._choose() for _ in self.chains]
            if batched:
                for ch, p in zip(self.chains, chosen):
                    ch.proposeSp(p)
                    if pipelined:
                        for pNum in ch.parts_to_read:
                            pf.partLogLikeBegin(ch.propTree.cTree, pNum)
                if pipelined:
                    # collect chain by chain: each wait is on that chain's own event, so a chain's accept / reject
                    # and cur/prop transfer run on the host while the GPU is still evaluating the later chains
                    for ch, p in zip(self.chains, chosen):
                        ch.readLikes()
                        ch.finish(p)
                else:
                    parts = sorted(set(q for ch in self.chains for q in ch.parts_to_read))
                    for pNum in parts:
                        who = [ch for ch in self.chains if pNum in ch.parts_to_read]
                        pf.treesPartLogLike([ch.propTree.cTree for ch in who], pNum)
                    for ch, p in zip(self.chains, chosen):
                        ch.finish(p)
            else:
                # the reference's order: a chain finishes its generation before the next one starts;
                # the accept draws are taken after all proposals in batched mode, so draw them in the
                # same order here: proposals first, then accept/reject
                for ch, p in zip(self.chains, chosen):
                    ch.proposeSp(p)
                    ch.readLikes()
                for ch, p in zip(self.chains, chosen):
                    ch.finish(p)
            self._proposeSwap()
            self.gen += 1
            byTemp = sorted(self.chains, key=lambda c: c.tempNum)
            self.trace.append((self.gen, [c.curTree.logLike for c in byTemp]))
        return self.trace

    def free(self):
        for c in self.chains:
            c.free()
