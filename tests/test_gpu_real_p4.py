"""GPU: the reference's REAL p4 package end to end on the B200 engine (missing item 3).

tests/dropin/p4_like_side.py imports p4 itself -- p4/tree.py, p4/model.py, p4/data.py, p4/alignment.py, p4/mcmc.py,
p4/chain.py, unmodified, staged under oracle/_ref/p4 by `make -C oracle` -- with a pf module installed as ``p4.pf``,
and drives it through p4's own public API: Tree.calcLogLike (p4/tree.py:9406), Tree.optLogLike (:9417),
Mcmc(...).run(n) (p4/mcmc.py:2496) with p4's own proposal code (p4/chain.py:313-1136).  The script runs twice, in
processes of their own: once on the reference's Pf engine (oracle/_ref/pf*.so), once on this repository's pf module
with NO pass-through -- every pf call p4 makes, the GSL wrappers of its proposals included, lands in libp4b200.so.

  * Tree.calcLogLike: lnL within 1e-9 relative;
  * Tree.optLogLike("newtAndBrentPowell") and ("allBrentPowell"): the optimum within 1e-6 relative, every branch
    length and model parameter near the reference's;
  * Mcmc.run: the SAME chain -- every sampled log-likelihood within 1e-9 relative and identical acceptance counts for
    every proposal over 200 generations of 4 Metropolis-coupled chains (the engines agree to ~1e-14, the random
    streams are the same, so the accept / reject decisions are the same).
"""
import json
import os
import subprocess
import sys

import pytest

import ref_loader

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(which, *extra):
    r = subprocess.run([sys.executable, os.path.join(HERE, "dropin", "p4_like_side.py"), which] + list(extra),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][-1]
    return json.loads(line[len("RESULT"):])


@pytest.fixture(scope="module", autouse=True)
def _need_p4(pkg):
    if not ref_loader.have_ref_p4():
        pytest.skip("the reference's p4 package is not staged (oracle/_ref/p4: run `make -C oracle` where /root/reference exists)")


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def test_real_p4_calcloglike():
    want = _run("ref", "--taxa", "24", "--patterns", "3000", "--gens", "0", "--skip-opt")
    got = _run("mine", "--taxa", "24", "--patterns", "3000", "--gens", "0", "--skip-opt")
    assert got["nPatterns"] == want["nPatterns"] == 3000
    assert _rel(got["lnL0"], want["lnL0"]) <= 1e-9


@pytest.mark.parametrize("method", ["newtAndBrentPowell", "allBrentPowell"])
def test_real_p4_optloglike(method):
    args = ("--taxa", "9", "--patterns", "400", "--gens", "0", "--opt-method", method)
    want = _run("ref", *args)
    got = _run("mine", *args)
    assert _rel(got["lnL0"], want["lnL0"]) <= 1e-9
    assert got["lnLopt"] > got["lnL0"] + 1.0
    assert _rel(got["lnLopt"], want["lnLopt"]) <= 1e-6, (got["lnLopt"], want["lnLopt"])
    for a, b in zip(got["brLens"], want["brLens"]):
        assert abs(a - b) <= 1e-3 + 2e-2 * b
    for key in ("comp", "rMatrix"):
        for a, b in zip(got[key], want[key]):
            assert abs(a - b) <= 2e-3 + 2e-2 * abs(b), key
    assert abs(got["shape"] - want["shape"]) <= 2e-2 * want["shape"] + 1e-3
    assert abs(got["pInvar"] - want["pInvar"]) <= 5e-3


def test_real_p4_mcmc_run_is_the_same_chain():
    args = ("--taxa", "12", "--patterns", "600", "--gens", "200", "--chains", "4", "--skip-opt")
    want = _run("ref", *args)
    got = _run("mine", *args)
    assert len(got["mcmc_likes"]) == len(want["mcmc_likes"]) >= 20
    for a, b in zip(got["mcmc_likes"], want["mcmc_likes"]):
        assert _rel(a, b) <= 1e-9
    for a, b in zip(got["final_likes"], want["final_likes"]):
        assert _rel(a, b) <= 1e-9
    assert got["accepted"] == want["accepted"]
