"""p4-phylogenetics_b200 -- B200-native Felsenstein-pruning likelihood engine for p4.

One path of pgfoster/p4-phylogenetics, rebuilt for sm_100a: transition-matrix
build, conditional-likelihood recursion and per-part log-likelihood reduction,
behind the reference's own ``pf`` extension-module interface.

  pf      drop-in mirror of the reference's ``p4.pf`` for the likelihood path
          (ctypes over libp4b200.so, C ABI in include/p4b200.h)
  host    the callers of that path, mirroring p4's Tree / Model / Data glue
  synth   seeded synthetic inputs for the BASELINE configs
  mcmc    the pf call protocol of p4's Mcmc / Chain (proposals, dirty path, cur/prop transfer)
  _build  nvcc recipe for libp4b200.so

The directory name is not a Python identifier; import it through the
``p4_phylogenetics_b200`` shim at the repository root.  Submodules load on first
use, so ``_build`` can be imported before the library exists; ``pf`` raises
ImportError when libp4b200.so is missing -- there is no CPU fallback.
"""
import importlib

__all__ = ["pf", "host", "synth", "mcmc", "_build"]


def __getattr__(name):
    if name in __all__:
        return importlib.import_module(__name__ + "." + name)
    raise AttributeError(name)
