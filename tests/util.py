"""Shared helpers for the parity tests."""
import numpy as np


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def build_pair(pkg, ref_pf, cfg, **kw):
    """The same synthetic config through this engine and through the reference engine."""
    mine = pkg.synth.build_config(pkg.pf, cfg, **kw)
    twin = pkg.host.clone_tree(mine, ref_pf)
    return mine, twin


def max_rel_err(a, b, floor=1e-300):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))
