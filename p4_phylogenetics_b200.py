"""Import shim: ``import p4_phylogenetics_b200`` -> the package in ``p4-phylogenetics_b200/``.

The package directory carries the project's name, which is not a valid Python
identifier.  Giving this module a ``__path__`` makes the import system treat it
as that package, so ``p4_phylogenetics_b200.pf`` etc. resolve to the files there.
"""
import os as _os

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "p4-phylogenetics_b200")
__path__ = [_dir]
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
with open(_os.path.join(_dir, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_dir, "__init__.py"), "exec"))
