"""CPU: the native minimisers behind the optimiser entry points (csrc/praxis.cpp), through the C ABI.

  * Brent's praxis restated -- against the reference's own praxis (Pf/brent.c, compiled unmodified into oracle/_ref)
    on the same analytic objectives with the same libc random() seed: same minimum, same minimiser, and a
    comparable number of function evaluations (identical on most: the restatement walks the same trajectory);
  * Powell's method in a box -- known constrained and unconstrained minima.
"""
import ctypes as C
import glob
import math
import os

import numpy as np
import pytest


def rosen(x):
    return sum(100.0 * (x[i + 1] - x[i] ** 2) ** 2 + (1.0 - x[i]) ** 2 for i in range(len(x) - 1))


def quad(x):
    return sum((i + 1) * (x[i] - 0.3 * i) ** 2 for i in range(len(x))) + 0.5 * x[0] * x[1]


def helical(x):
    th = math.atan2(x[1], x[0]) / (2 * math.pi)
    return 100.0 * ((x[2] - 10.0 * th) ** 2 + (math.hypot(x[0], x[1]) - 1.0) ** 2) + x[2] ** 2


CASES = [("rosen2", rosen, [-1.2, 1.0]), ("rosen5", rosen, [-1.2, 1.0, -0.5, 0.8, 1.1]), ("quad6", quad, [1.0] * 6),
         ("helical", helical, [-1.0, 0.0, 0.0])]


@pytest.fixture(scope="module")
def ref_praxis():
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sos = glob.glob(os.path.join(here, "oracle", "_ref", "pf*.so"))
    if not sos:
        pytest.skip("oracle/_ref/pf*.so not built (needs /root/reference at build time)")
    lib = C.CDLL(sos[0])
    FN = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double))
    lib.newBrent.restype = C.c_void_p
    lib.newBrent.argtypes = [C.c_int]
    lib.praxis.restype = C.c_double
    lib.praxis.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double), FN]

    def run(f, x0, tol, h):
        n = len(x0)
        count = [0]

        def cb(p):
            count[0] += 1
            return f([p[i] for i in range(n)])
        x = (C.c_double * n)(*x0)
        v = lib.praxis(lib.newBrent(n), tol, h, n, x, FN(cb))
        return v, [x[i] for i in range(n)], count[0]
    return run


@pytest.mark.parametrize("name,f,x0", CASES)
@pytest.mark.parametrize("tol,h", [(1e-4, 0.1), (1e-4, 1.0), (1e-4, 0.05)])
def test_praxis_walks_like_the_reference(pf, ref_praxis, name, f, x0, tol, h):
    libc = C.CDLL(None)
    libc.srandom(1)
    vr, xr, nr = ref_praxis(f, x0, tol, h)
    count = [0]

    def g(p):
        count[0] += 1
        return f(p)
    libc.srandom(1)
    vm, xm = pf.praxisMinimize(g, x0, tol, h)
    assert abs(vm - vr) <= 1e-12 * max(1.0, abs(vr)) + 1e-15
    assert np.max(np.abs(np.array(xm) - np.array(xr))) <= 1e-7
    assert abs(count[0] - nr) <= 0.15 * nr


def test_praxis_identical_trajectory_on_rosenbrock(pf, ref_praxis):
    """Where no rounding tie intervenes the restatement is the reference's algorithm step for step: same evaluations, same bits."""
    libc = C.CDLL(None)
    libc.srandom(1)
    vr, xr, nr = ref_praxis(rosen, [-1.2, 1.0], 1e-4, 0.1)
    count = [0]
    libc.srandom(1)
    vm, xm = pf.praxisMinimize(lambda p: (count.__setitem__(0, count[0] + 1), rosen(p))[1], [-1.2, 1.0], 1e-4, 0.1)
    assert (vm, list(xm), count[0]) == (vr, xr, nr)


def test_bounded_powell(pf):
    v, x, n = pf.boundedMinimize(quad, [1.0] * 6, [-10.0] * 6, [10.0] * 6)
    # the minimum of the quadratic, from its normal equations
    A = np.diag([2.0 * (i + 1) for i in range(6)])
    A[0, 1] = A[1, 0] = 0.5
    b = np.array([2.0 * (i + 1) * 0.3 * i for i in range(6)])
    xs = np.linalg.solve(A, b)
    assert np.max(np.abs(x - xs)) <= 1e-5
    assert abs(v - quad(xs)) <= 1e-10
    # an active bound: every coordinate's free minimum below 0.5 is clipped to it
    v, x, n = pf.boundedMinimize(quad, [1.0] * 6, [0.5] * 6, [3.0] * 6)
    assert np.all(x >= 0.5) and np.all(x <= 3.0)
    assert np.max(np.abs(x - np.array([0.5, 0.5, 0.6, 0.9, 1.2, 1.5]))) <= 1e-5
    # Rosenbrock with x0 confined to <= 0.8: the constrained minimiser lies on the face x0 = 0.8
    v, x, n = pf.boundedMinimize(rosen, [-1.2, 1.0, 0.5], [-2.0] * 3, [0.8, 2.0, 2.0], ftol=1e-14)
    assert abs(x[0] - 0.8) <= 1e-9
    assert v < rosen([0.8, 0.64, 0.4096]) + 1e-9
    # starts outside the box are moved inside
    v, x, n = pf.boundedMinimize(quad, [50.0] * 6, [0.0] * 6, [2.0] * 6)
    assert np.all(x >= 0.0) and np.all(x <= 2.0)
