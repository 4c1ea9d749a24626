"""tests/dropin/p4_data_side.py -- run by tests/test_p4_dropin.py in a process of its own (build container only).

Imports the reference's REAL Python package ``p4`` from /root/reference with either the reference's own Pf engine
(argument "ref") or this repository's ``pf`` module (argument "mine") installed as ``p4.pf`` -- the substitution
INTEGRATION.md section 1 describes -- and runs p4's own data-side code on its own example alignment: reading, parts and
pattern compression, Part.composition, Data.simpleBigXSquared / simpleConstantSitesCount / unconstrained log-likelihood,
Data.bootstrap on p4's mt19937 stream.  Prints one JSON line.  Nothing here needs a GPU."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import ref_loader
which = sys.argv[1]
if which == "ref":
    pfm = ref_loader.load_ref_pf()
else:
    import p4_phylogenetics_b200 as P
    pfm = P.pf
p4 = ref_loader.load_ref_p4(pf_module=pfm)
from p4 import read, var, Data, func
var.verboseRead = 0
EX = os.path.join(ref_loader.REF_ROOT, "share", "Examples", "L_mcmc")
read(os.path.join(EX, "d.nex"))
d = Data()
out = {}
out["nPatterns"] = [pfm.partPatternCount(p.cPart) for p in d.parts]
out["comp"] = list(d.parts[0].composition())
out["comp_sub"] = list(d.parts[0].composition([0, 2]))
out["bigX"] = d.simpleBigXSquared()
out["const"] = d.simpleConstantSitesCount()
try:
    d.calcUnconstrainedLogLikelihood1()
    out["unc"] = d.unconstrainedLogLikelihood
except SystemExit as e:
    out["unc"] = "fatal"
var.gsl_rng = pfm.gsl_rng_get()
pfm.gsl_rng_set(var.gsl_rng, 5)
b = d.bootstrap()
out["boot"] = [s.sequence for s in b.alignments[0].sequences]
out["boot_nPat"] = [pfm.partPatternCount(p.cPart) for p in b.parts]
# two character partitions of the same alignment (share/Examples/H_calcLike/C_3_data_partitions pattern), and a recoded datatype
a = var.alignments[0]
half = a.length // 2
read("#nexus\nbegin sets;\n charset c1 = 1-%d;\n charset c2 = %d-.;\n charpartition cp1 = c1:c1, c2:c2;\nend;\n" % (half, half + 1))
a.setCharPartition("cp1")
d2 = Data()
out["two_parts_nPatterns"] = [pfm.partPatternCount(p.cPart) for p in d2.parts]
out["two_parts_comp"] = [list(p.composition()) for p in d2.parts]
out["two_parts_sites_seq0"] = [pfm.partSequenceSitesCount(p.cPart, 0) for p in d2.parts]
out["two_parts_counts_seq1"] = [list(pfm.singleSequenceBaseCounts(p.cPart, 1)) for p in d2.parts]
var.alignments = []
read(os.path.join(EX, "B_grouped_aa", "protein.nex"))
pa = var.alignments[0]
pa.recodeDayhoff()
d3 = Data()
out["dayhoff_nPatterns"] = [pfm.partPatternCount(p.cPart) for p in d3.parts]
out["dayhoff_comp"] = list(d3.parts[0].composition())
out["dayhoff_symbols"] = pfm.symbolSequences(d3.parts[0].cPart)[:200]
print("RESULT" + json.dumps(out))
