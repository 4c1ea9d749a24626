"""CPU: the reference's REAL p4 Python package with this repository's ``pf`` module installed as ``p4.pf`` (the drop-in
substitution of INTEGRATION.md) runs its own data-side code -- read, Data, parts, Part.composition, X^2, constant sites,
unconstrained log-likelihood, Data.bootstrap -- and gets the values it gets on the reference's own Pf engine.
Build container only (needs /root/reference); the likelihood side of the same substitution is what tests/test_trace.py replays."""
import json
import os
import subprocess
import sys

import pytest

import ref_loader

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(which, script="p4_data_side.py"):
    r = subprocess.run([sys.executable, os.path.join(HERE, "dropin", script), which], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][-1]
    return json.loads(line[len("RESULT"):])


def test_real_p4_data_side_runs_on_this_pf_module(pkg):
    if not ref_loader.have_ref_p4():
        pytest.skip("the reference's p4 package is not present (it never is on the GPU box)")
    want, got = _run("ref"), _run("mine")
    assert got.keys() == want.keys()
    for k in want:
        assert got[k] == want[k], k
    assert got["nPatterns"][0] > 10 and len(got["boot"]) >= 4


def test_real_p4_model_side_runs_on_this_pf_module(pkg):
    """p4's own Tree model API and Model.allocCStuff / setCStuff on this pf module: empirical composition, gamma rates and the
    normalised rate matrix the engine then holds are the reference engine's, bit for bit."""
    if not ref_loader.have_ref_p4():
        pytest.skip("the reference's p4 package is not present (it never is on the GPU box)")
    want, got = _run("ref", "p4_model_side.py"), _run("mine", "p4_model_side.py")
    assert got == want
    assert len(got["Q"]) == 16 and got["nFreePrams"] == 3 + 5 + 1 + 1
