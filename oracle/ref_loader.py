"""oracle/ref_loader.py -- TEST INFRASTRUCTURE ONLY.

Loads the reference's own Pf engine (``oracle/_ref/pf*.so``, built by
``oracle/Makefile`` from the unmodified sources under /root/reference/Pf) as a
Python module, and the reference's pure-Python ``p4`` package on top of it
(from /root/reference in the build container, from the copy staged under
oracle/_ref/p4 on the GPU box).

Only ``tests/``, ``__graft_entry__.smoke()`` and the cpu_baseline / ``--impl
reference`` legs of ``bench.py`` may import this file.  The product package
never does.
"""
import glob
import importlib.machinery
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("P4_REFERENCE_ROOT", "/root/reference")


def ref_pf_path():
    hits = sorted(glob.glob(os.path.join(_HERE, "_ref", "pf*.so")))
    return hits[0] if hits else None


def have_ref_pf():
    return ref_pf_path() is not None


def load_ref_pf(as_name="pf_ref"):
    """Return the reference ``pf`` extension module (Pf/pfmodule.c:3033 PyInit_pf)."""
    if as_name in sys.modules:
        return sys.modules[as_name]
    path = ref_pf_path()
    if path is None:
        raise ImportError("oracle/_ref/pf*.so not built; run `make -C oracle` where /root/reference exists")
    # The init symbol is PyInit_pf, so the spec name must end in "pf".
    loader = importlib.machinery.ExtensionFileLoader("pf", path)
    spec = importlib.util.spec_from_loader("pf", loader, origin=path)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    sys.modules[as_name] = mod
    return mod


def ref_p4_parent():
    """Directory that holds the reference's ``p4`` package: /root/reference in the build container, else the copy
    ``make -C oracle`` staged under oracle/_ref/ (git-ignored; it travels to the GPU box with the snapshot)."""
    if os.path.isdir(os.path.join(REF_ROOT, "p4")):
        return REF_ROOT
    if os.path.isfile(os.path.join(_HERE, "_ref", "p4", "__init__.py")):
        return os.path.join(_HERE, "_ref")
    return None


def have_ref_p4():
    return ref_p4_parent() is not None and have_ref_pf()


def load_ref_p4(pf_module=None):
    """Import the reference's ``p4`` package (build container only).

    ``pf_module`` is installed as ``p4.pf`` (default: the reference's own Pf
    engine).  Passing this repository's ``pf`` mirror instead is exactly the
    drop-in substitution INTEGRATION.md describes.
    """
    if "p4" in sys.modules:
        return sys.modules["p4"]
    parent = ref_p4_parent()
    if parent is None:
        raise ImportError("the reference's p4 package is neither at %s/p4 nor staged under oracle/_ref/p4" % REF_ROOT)
    if pf_module is None:
        pf_module = load_ref_pf()
    sys.modules["p4.pf"] = pf_module
    if "bitarray" not in sys.modules:
        # p4/stmcmc.py:25 hard-imports bitarray (supertree code, out of scope).
        stub = types.ModuleType("bitarray")
        stub.bitarray = type("bitarray", (), {})
        sys.modules["bitarray"] = stub
    sys.path.insert(0, parent)
    try:
        import p4  # noqa: F401
    finally:
        sys.path.remove(parent)
    p4 = sys.modules["p4"]
    p4.pf = pf_module
    return p4
