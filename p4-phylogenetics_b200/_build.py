"""Build recipe for libp4b200.so (the C-ABI library declared in include/p4b200.h).

nvcc cross-compiles for sm_100a without a GPU; the library is built in-tree so
that it travels with the repository snapshot to the GPU box.  CUDA runtime is
linked statically; NCCL is bound at run time (csrc/comm.cpp).
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libp4b200.so")
SOURCES = ["capi.cpp", "data.cpp", "model.cpp", "comm.cpp", "opt.cpp", "praxis.cpp", "sim.cpp", "partstats.cpp", "tree.cu"]
HEADERS = ["engine.h", "optim.h", "kernels.cuh", "tree_dna.cuh", "tree_aa.cuh", "tree_dmma.cuh", "newt.cuh", "protein_rmatrices.inc", os.path.join("..", "..", "include", "p4b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=true",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fvisibility=default",
    "-shared", "-cudart", "static",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def hot_path():
    import sysconfig
    return os.path.join(HERE, "_pfhot" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_hot(force=False, verbose=False):
    """_pfhot: CPython METH_FASTCALL bindings of the per-node calls (csrc/pfhot.c), linked against libp4b200.so."""
    import sysconfig
    out, src = hot_path(), os.path.join(CSRC, "pfhot.c")
    deps = [src, LIB, os.path.join(HERE, "..", "include", "p4b200.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-fPIC", "-shared", "-Wall", "-I" + sysconfig.get_paths()["include"],
           "-o", out, src, "-L" + HERE, "-l:libp4b200.so", "-Wl,-rpath,$ORIGIN"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return out


def build(force=False, verbose=False, extra=()):
    if force or needs_build():
        cmd = [nvcc_path()] + NVCC_FLAGS + list(extra) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    build_hot(force=force, verbose=verbose)
    return LIB


if __name__ == "__main__":
    import sys
    build(force=True, verbose=True, extra=sys.argv[1:])
    print(LIB)
