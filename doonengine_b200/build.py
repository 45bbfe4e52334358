"""Builds doonengine_b200/libdoon_b200.so in-tree with nvcc for sm_100a.

    python -m doonengine_b200.build [--force] [--verbose]

The library is one C-ABI shared object: the DN_* API of include/DoonEngine/voxel.h plus the additive DN_b200_*
entry points of include/DoonEngine/b200.h.  Parity-critical flags: -fmad=false (no FMA contraction in device
code), IEEE division / square root (nvcc defaults, no --use_fast_math), -ffp-contract=off for the host code.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdoon_b200.so")
SCENES_LIB = os.path.join(HERE, "libdoon_scenes.so")  # native generators of the large synthetic maps (bench / tests only)

SOURCES = ["draw.cu", "light.cu", "compact.cu", "upload.cu", "peer.cu", "pick.cu", "engine.cpp", "volume_host.cpp"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".h", ".cuh")))  # every header: a stale library is worse than a slow build
PUBLIC_HEADERS = ["DoonEngine/voxel.h", "DoonEngine/b200.h", "DoonEngine/globals.h", "DoonEngine/mathtypes.h"]

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
                     "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall,-Wno-unused-function,-pthread"]


def nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _inputs():
    files = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    files += [os.path.join(ROOT, "include", h) for h in PUBLIC_HEADERS]
    files.append(os.path.abspath(__file__))
    return files


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(f) <= t for f in _inputs())


def build_scenes(force=False):
    """gcc -> libdoon_scenes.so: csrc/scenegen.c, the native twin of scenes.sparse_balls / dense_corridors."""
    src = os.path.join(CSRC, "scenegen.c")
    if not force and os.path.exists(SCENES_LIB) and os.path.getmtime(SCENES_LIB) >= os.path.getmtime(src):
        return SCENES_LIB
    subprocess.check_call(["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", "-Wall", "-o", SCENES_LIB, src])
    return SCENES_LIB


def build(force=False, verbose=False, defines=(), out=None):
    """compile (if stale) and return the path of the shared library.  `defines` / `out` build an experimental variant
    (e.g. defines=["FLAT_FIRST_STEP=1"], out="libdoon_b200_v1.so"; load it with $DN_B200_LIB)."""
    global LIB
    if out:
        return _build_variant(defines, out, verbose)
    build_scenes(force)
    if not force and up_to_date():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    objs = []
    procs = []
    for s in SOURCES:
        obj = os.path.join(objdir, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        cmd = [nvcc()] + NVCC_FLAGS + inc + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stdout.write(out)
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xcompiler", "-pthread", "-cudart", "static"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


def _build_variant(defines, out, verbose):
    objdir = os.path.join(HERE, "build", os.path.splitext(out)[0])
    os.makedirs(objdir, exist_ok=True)
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    objs, procs = [], []
    for s in SOURCES:
        obj = os.path.join(objdir, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        cmd = [nvcc()] + NVCC_FLAGS + inc + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", os.path.join(CSRC, s), "-o", obj]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        o, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stdout.write(o)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed")
    path = os.path.join(HERE, out)
    subprocess.check_call([nvcc()] + ARCH + ["-shared", "-o", path] + objs + ["-Xcompiler", "-pthread", "-cudart", "static"])
    return path


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
