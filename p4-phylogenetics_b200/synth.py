"""synth -- seeded synthetic inputs for the BASELINE configs.

Everything derives from one ``numpy.random.Generator(PCG64(seed))``:

  tree       random unrooted binary topology by sequential random edge insertion
             under a trifurcating root; branch lengths Exp(mean 0.05) clipped to
             [1e-4, 0.5]
  DNA model  pi ~ Dirichlet(10), GTR rates ~ Dirichlet(5) normalised to sum 1,
             gamma shape 0.5 with 4 categories, optional pInvar
  protein    LG exchangeabilities with LG pi (or per-node comps ~ Dirichlet(50 pi))
  alignment  simulated down the tree under the same model; `nPatterns` distinct
             columns, each repeated Poisson(1)+1 times and shuffled; 1 % gaps and
             0.5 % ambiguity codes sprinkled in before the repetition

The module builds inputs only; it contains no likelihood code and is used the
same way with this repository's engine and with the reference engine.
"""
import json
import os

import numpy as np

from . import host

_HERE = os.path.dirname(os.path.abspath(__file__))


def protein_comp(spec="lg"):
    with open(os.path.join(_HERE, "data", "protein_comps.json")) as f:
        v = np.array(json.load(f)[spec], dtype=np.float64)
    return normalise_comp(v)


def normalise_comp(v):
    """Scale to sum 1 within the 1e-14 the engine insists on (Pf/p4_tree.c:445)."""
    v = np.asarray(v, dtype=np.float64)
    v = v / v.sum()
    i = int(np.argmax(v))
    v[i] += 1.0 - v.sum()
    return v


# ------------------------------------------------------------------------------
# trees
# ------------------------------------------------------------------------------
def random_tree(pf, nTax, rng, root_is_leaf=False):
    """Random binary topology with a trifurcating root (or, unusually, a leaf root)."""
    class _N:
        pass
    root = _N()
    root.children, root.parent, root.tax = [], None, -1
    all_nodes = [root]

    def add_leaf(parent, tax):
        n = _N()
        n.children, n.parent, n.tax = [], parent, tax
        parent.children.append(n)
        all_nodes.append(n)
        return n
    for k in range(min(3, nTax)):
        add_leaf(root, k)
    for k in range(3, nTax):
        cands = [n for n in all_nodes if n.parent is not None]
        target = cands[int(rng.integers(len(cands)))]
        par = target.parent
        mid = _N()
        mid.children, mid.parent, mid.tax = [target], par, -1
        par.children[par.children.index(target)] = mid
        target.parent = mid
        all_nodes.append(mid)
        add_leaf(mid, k)
    if root_is_leaf:
        # re-root on the leaf carrying taxon 0: that leaf becomes the root and its
        # old parent its only child (a root that is a leaf, Pf/p4_tree.c:1199).
        leaf = [n for n in all_nodes if n.tax == 0][0]
        path = []
        n = leaf
        while n is not None:
            path.append(n)
            n = n.parent
        for child, par in zip(path[:-1], path[1:]):
            par.children.remove(child)
        for child, par in zip(path[:-1], path[1:]):
            child.children.append(par)
            par.parent = child
        leaf.parent = None
        old_root = path[-1]
        if len(old_root.children) == 1:           # splice out a degree-2 old root
            only = old_root.children[0]
            gp = old_root.parent
            gp.children[gp.children.index(old_root)] = only
            only.parent = gp
        root = leaf
    # number in pre-order, root = 0
    order = []
    stack = [root]
    while stack:
        n = stack.pop()
        order.append(n)
        for c in reversed(n.children):
            stack.append(c)
    nodes = [host.Node(i) for i in range(len(order))]
    index = {id(n): i for i, n in enumerate(order)}
    for i, n in enumerate(order):
        h = nodes[i]
        h.isLeaf = 1 if n.tax >= 0 else 0
        h.seqNum = n.tax
        h.parent = nodes[index[id(n.parent)]] if n.parent is not None else None
        kids = [nodes[index[id(c)]] for c in n.children]
        h.leftChild = kids[0] if kids else None
        for a, b in zip(kids[:-1], kids[1:]):
            a.sibling = b
        h.br.len = float(np.clip(rng.exponential(0.05), 1e-4, 0.5))
    t = host.Tree(pf, nodes, nodes[0])
    t.setPreAndPostOrder()
    return t


# ------------------------------------------------------------------------------
# models
# ------------------------------------------------------------------------------
def dna_model_part(num, rng, nGammaCat=4, pInvar=0.0, alpha=0.5, free=0):
    mp = host.ModelPart(num, 4, nGammaCat)
    mp.comps.append(host.Comp(normalise_comp(rng.dirichlet(10.0 * np.ones(4))), free=free))
    r = rng.dirichlet(5.0 * np.ones(6))
    mp.rMatrices.append(host.RMatrix("specified", r / r.sum(), free=free))
    if nGammaCat > 1:
        mp.gdasrvs.append(host.Gdasrv(nGammaCat, alpha, free=free))
    mp.pInvar = host.PInvar(pInvar, free=0)
    return mp


def protein_model_part(num, rng, spec="lg", nGammaCat=4, alpha=0.5, nComps=1):
    mp = host.ModelPart(num, 20, nGammaCat)
    base = protein_comp(spec)
    if nComps == 1:
        mp.comps.append(host.Comp(base))
    else:
        for _ in range(nComps):
            mp.comps.append(host.Comp(normalise_comp(rng.dirichlet(50.0 * base)), free=1))
        mp.isHet = 1
    mp.rMatrices.append(host.RMatrix(spec))
    if nGammaCat > 1:
        mp.gdasrvs.append(host.Gdasrv(nGammaCat, alpha))
    return mp


# ------------------------------------------------------------------------------
# alignments
# ------------------------------------------------------------------------------
def _rate_matrix(mp, compIdx=0):
    dim = mp.dim
    pi = mp.comps[compIdx].val
    rm = mp.rMatrices[0]
    R = np.ones((dim, dim))
    if rm.spec == "specified":
        k = 0
        for i in range(dim - 1):
            for j in range(i + 1, dim):
                R[i, j] = R[j, i] = rm.val[k]
                k += 1
    elif rm.spec in host.RMATRIX_PROTEIN_SPEC:
        # the generator only needs *a* plausible reversible process to draw data
        # from; a flat exchangeability matrix with the right pi is enough.
        pass
    Q = R * pi[None, :]
    np.fill_diagonal(Q, 0.0)
    np.fill_diagonal(Q, -Q.sum(1))
    Q /= -(pi * np.diag(Q)).sum()
    return Q, pi


def _gamma_rates(alpha, nCat):
    if nCat == 1:
        return np.ones(1)
    # mean-of-quantile-bin rates by Monte Carlo-free quadrature: good enough to simulate from
    from scipy.stats import gamma as _g
    edges = _g.ppf(np.linspace(0, 1, nCat + 1), alpha, scale=1.0 / alpha)
    cdf1 = _g.cdf(edges, alpha + 1, scale=1.0 / alpha)
    return (cdf1[1:] - cdf1[:-1]) * nCat


def simulate_columns(tree, mp, nSites, rng):
    """States (nTax, nSites) uint8 simulated down ``tree`` under model part ``mp``."""
    dim, nCat = mp.dim, mp.nGammaCat
    Q, pi = _rate_matrix(mp)
    w, V = np.linalg.eig(Q)
    Vi = np.linalg.inv(V)
    alpha = float(mp.gdasrvs[0].val[0]) if mp.gdasrvs else 1.0
    rates = _gamma_rates(alpha, nCat)
    pinv = float(mp.pInvar.val)
    cats = rng.integers(nCat, size=nSites).astype(np.intp)
    invariant = rng.random(nSites) < pinv
    nTax = sum(1 for n in tree.nodes if n.isLeaf)
    out = np.zeros((nTax, nSites), dtype=np.uint8)
    states = {}
    cum_pi = np.cumsum(pi)
    cum_pi[-1] = 1.0
    root = tree.root
    states[root.nodeNum] = np.minimum((rng.random(nSites)[:, None] > cum_pi[None, :]).sum(1), dim - 1).astype(np.uint8)
    if root.isLeaf:
        out[root.seqNum] = states[root.nodeNum]
    for i in tree.preOrder:
        if i == host.NO_ORDER or i == root.nodeNum:
            continue
        n = tree.nodes[i]
        par = states[n.parent.nodeNum]
        cumP = np.empty((nCat, dim, dim))
        for c in range(nCat):
            t = n.br.len * rates[c] * mp.relRate / (1.0 - pinv)
            P = np.real((V * np.exp(w * t)[None, :]) @ Vi)
            P = np.clip(P, 0.0, None)
            P /= P.sum(1, keepdims=True)
            cumP[c] = np.cumsum(P, axis=1)
            cumP[c][:, -1] = 1.0
        st = np.empty(nSites, dtype=np.uint8)
        CH = 1 << 18
        for lo in range(0, nSites, CH):
            hi = min(nSites, lo + CH)
            cp = cumP[cats[lo:hi], par[lo:hi]]
            st[lo:hi] = np.minimum((rng.random(hi - lo)[:, None] > cp).sum(1), dim - 1)
        st[invariant] = par[invariant]
        if n.isLeaf:
            out[n.seqNum] = st
        else:
            states[i] = st
        if n.sibling is None:       # pre-order: the parent's last child has now read it
            del states[n.parent.nodeNum]
    return out


def distinct_columns(tree, mp, nPatterns, rng):
    """First ``nPatterns`` distinct simulated columns, in order of first appearance."""
    nTax = sum(1 for n in tree.nodes if n.isLeaf)
    weights = rng.integers(1, 2 ** 62, size=nTax, dtype=np.uint64) | np.uint64(1)
    got = np.zeros((nTax, 0), dtype=np.uint8)
    seen = np.zeros(0, dtype=np.uint64)
    draw = int(nPatterns * 1.05) + 16
    for _ in range(200):
        cols = simulate_columns(tree, mp, draw, rng)
        with np.errstate(over="ignore"):
            h = (cols.astype(np.uint64) * weights[:, None]).sum(0, dtype=np.uint64)
        _, first = np.unique(h, return_index=True)
        first.sort()
        fresh = first[~np.isin(h[first], seen)]
        got = np.concatenate([got, cols[:, fresh]], axis=1)
        seen = np.concatenate([seen, h[fresh]])
        if got.shape[1] >= nPatterns:
            return got[:, :nPatterns]
        draw = max(16, int((nPatterns - got.shape[1]) * 1.5))
    raise RuntimeError("could not draw %d distinct columns (tree too short for that many patterns?)" % nPatterns)


def make_alignment(pf, tree, mp, nPatterns, rng, datatype, gap_frac=0.01, ambig_frac=0.005, repeat=True):
    """An ``host.Alignment`` whose compressed form has exactly ``nPatterns`` patterns (w.h.p.)."""
    if datatype == "dna":
        symbols, equates, ambig = host.DNA_SYMBOLS, host.DNA_EQUATES, "ryn"
    elif datatype == "protein":
        symbols, equates, ambig = host.PROTEIN_SYMBOLS, host.PROTEIN_EQUATES, "xb"
    else:
        raise ValueError(datatype)
    cols = distinct_columns(tree, mp, nPatterns, rng)
    nTax = cols.shape[0]
    lut = np.frombuffer(symbols.encode(), dtype=np.uint8)
    chars = lut[cols]
    if gap_frac > 0:
        chars[rng.random(chars.shape) < gap_frac] = ord("-")
    if ambig_frac > 0:
        m = rng.random(chars.shape) < ambig_frac
        chars[m] = np.frombuffer(ambig.encode(), dtype=np.uint8)[rng.integers(len(ambig), size=int(m.sum()))]
    if repeat:
        reps = rng.poisson(1.0, size=chars.shape[1]) + 1
        idx = np.repeat(np.arange(chars.shape[1]), reps)
        rng.shuffle(idx)
        chars = chars[:, idx]
    seqs = [np.ascontiguousarray(chars[i]).tobytes() for i in range(nTax)]
    return host.Alignment(pf, seqs, symbols, equates)


# ------------------------------------------------------------------------------
# the BASELINE configs
# ------------------------------------------------------------------------------
def build_config(pf, cfg, nTax=None, nPatterns=None, seed=None):
    """Tree + data + model for BASELINE.json config ``cfg`` (1..5), optionally scaled down."""
    seed = 20240 + cfg if seed is None else seed
    rng = np.random.Generator(np.random.PCG64(seed))
    if cfg == 1:
        nTax, nPatterns = nTax or 32, nPatterns or 10000
        tree = random_tree(pf, nTax, rng)
        mps = [dna_model_part(0, rng, 4, pInvar=0.2)]
        alns = [make_alignment(pf, tree, mps[0], nPatterns, rng, "dna")]
    elif cfg == 2:
        nTax, nPatterns = nTax or 200, nPatterns or 1000000
        tree = random_tree(pf, nTax, rng)
        mps = [dna_model_part(0, rng, 4, pInvar=0.0)]
        alns = [make_alignment(pf, tree, mps[0], nPatterns, rng, "dna")]
    elif cfg == 3:
        nTax, nPatterns = nTax or 100, nPatterns or 200000
        tree = random_tree(pf, nTax, rng)
        mps = [protein_model_part(0, rng, "lg", 4)]
        alns = [make_alignment(pf, tree, mps[0], nPatterns, rng, "protein")]
    elif cfg == 4:
        nTax, nPatterns = nTax or 60, nPatterns or 50000
        tree = random_tree(pf, nTax, rng)
        nNodes = len(tree.nodes)
        mps = [protein_model_part(p, rng, "lg", 4, nComps=nNodes) for p in range(4)]
        sim = protein_model_part(0, rng, "lg", 4)
        alns = [make_alignment(pf, tree, sim, nPatterns, rng, "protein") for _ in range(4)]
    elif cfg == 5:   # the MCMC config: every model parameter free
        nTax, nPatterns = nTax or 100, nPatterns or 500000
        tree = random_tree(pf, nTax, rng)
        mps = [dna_model_part(0, rng, 4, pInvar=0.0, free=1)]
        alns = [make_alignment(pf, tree, mps[0], nPatterns, rng, "dna")]
    else:
        raise ValueError("config %r" % cfg)
    data = host.Data(pf, alns)
    model = host.Model(pf, mps)
    tree.attach(data, model)
    if cfg == 4:   # NDCH2: one composition per node (share/Examples/W_recipes/sMcmcNDCH2.py pattern)
        for n in tree.nodes:
            for p in range(4):
                n.parts[p].compNum = n.nodeNum
    return tree


# ------------------------------------------------------------------------------
# any state count (61-state codon-like data: a 'standard' datatype with 61 symbols, SURVEY.md section 2 note)
# ------------------------------------------------------------------------------
SYMBOLS_61 = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ012345678"


def evolve_states(rng, tree, dim, nSites, mut=0.35):
    """Leaf state rows with phylogenetic signal: states copied down the tree, changed with a per-branch probability."""
    tree.setPreAndPostOrder()
    states = {tree.root.nodeNum: rng.integers(dim, size=nSites)}
    out = {}
    for i in tree.preOrder:
        if i < 0 or i == tree.root.nodeNum:
            continue
        n = tree.nodes[i]
        s = states[n.parent.nodeNum].copy()
        m = rng.random(nSites) < min(0.9, mut * (0.3 + 5.0 * n.br.len))
        s[m] = rng.integers(dim, size=int(m.sum()))
        states[i] = s
        if n.isLeaf:
            out[n.seqNum] = s
    return [out[k] for k in sorted(out)]


def build_generic(pf, symbols, nTax, nSites, nCat, seed, equates=None, pInvar=0.0, tree=None):
    """Tree + data + model for a datatype with len(symbols) states: random composition and exchangeabilities, gamma rates."""
    rng = np.random.Generator(np.random.PCG64(seed))
    dim = len(symbols)
    if tree is None:
        tree = random_tree(pf, nTax, rng)
    lut = np.frombuffer(symbols.encode(), dtype=np.uint8)
    seqs = []
    eqChars = sorted((equates or {}).keys())
    for s in evolve_states(rng, tree, dim, nSites):
        chars = lut[s].copy()
        chars[rng.random(nSites) < 0.02] = ord("-")
        for e in eqChars:
            chars[rng.random(nSites) < 0.01] = ord(e)
        seqs.append(chars.tobytes())
    aln = host.Alignment(pf, seqs, symbols, equates or {})
    mp = host.ModelPart(0, dim, nCat)
    mp.comps.append(host.Comp(normalise_comp(rng.dirichlet(20.0 * np.ones(dim)))))
    r = rng.dirichlet(3.0 * np.ones(dim * (dim - 1) // 2))
    mp.rMatrices.append(host.RMatrix("specified", r / r.sum()))
    if nCat > 1:
        mp.gdasrvs.append(host.Gdasrv(nCat, 0.7))
    mp.pInvar = host.PInvar(pInvar)
    tree.attach(host.Data(pf, [aln]), host.Model(pf, [mp]))
    return tree
