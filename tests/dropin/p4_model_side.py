"""tests/dropin/p4_model_side.py -- run by tests/test_p4_dropin.py in a process of its own (build container only).

The reference's REAL ``p4`` package with either pf module installed as ``p4.pf`` builds a model through p4's own Tree API
(newComp with empirical frequencies, newRMatrix, setNGammaCat, newGdasrv, setPInvar), runs modelSanityCheck,
setEmpiricalComps, Model.allocCStuff and Model.setCStuff (p4/model.py:735-833, 153-207) -- the host side of the engine: no
device is touched before p4_newTree -- and reads back what the engine holds: gamma rates, the normalised Q, relRate."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_loader
import numpy as np
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
t = func.randomTree(taxNames=d.taxNames)
t.data = d
t.newComp(free=1, spec="empirical")
t.newRMatrix(free=1, spec="specified", val=[1.2, 3.1, 0.8, 0.9, 3.5, 1.0])
t.setNGammaCat(nGammaCat=4)
t.newGdasrv(free=1, val=0.6)
t.setPInvar(free=1, val=0.15)
t.modelSanityCheck()
t.setEmpiricalComps()
t.model.allocCStuff()
t.model.setCStuff()
mp = t.model.parts[0]
pfm.p4_resetBQET(t.model.cModel, 0, 0, 0)
Q = np.zeros((4, 4))
pfm.getBigQ(t.model.cModel, 4, 0, 0, 0, Q)
out = {"comp": [float(v) for v in mp.comps[0].val], "rates": [float(v) for v in mp.gdasrvs[0].rates], "Q": Q.ravel().tolist(),
       "nFreePrams": int(t.model.nFreePrams), "relRate": float(pfm.p4_getRelRate(t.model.cModel, 0))}
print("RESULT" + json.dumps(out))
