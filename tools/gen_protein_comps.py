"""tools/gen_protein_comps.py -- regenerate p4-phylogenetics_b200/data/protein_comps.json.

The equilibrium frequencies that go with the empirical protein matrices (LG,
WAG, ...) are published data.  The reference returns them from
p4.func.getProteinEmpiricalModelComp(spec) (p4/func.py:3560); this script asks
the reference (build container only) and writes the numbers as JSON, used by
the synthetic-input generator ("LG exchangeabilities with
LG pi").  Usage: python tools/gen_protein_comps.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_loader  # noqa: E402


def main():
    p4 = ref_loader.load_ref_p4()
    out = {}
    for spec in p4.var.rMatrixProteinSpecs:
        out[spec] = [float(x) for x in p4.func.getProteinEmpiricalModelComp(spec)]
        assert abs(sum(out[spec]) - 1.0) < 1e-12
    path = os.path.join(ROOT, "p4-phylogenetics_b200", "data", "protein_comps.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path)


if __name__ == "__main__":
    main()
