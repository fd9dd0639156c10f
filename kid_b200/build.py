"""Build kid_b200/libkidmp.so (the C-ABI library of include/kidmp.h) in-tree with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the repo snapshot.  Flags that matter for
parity (DESIGN.md "Arithmetic"): -fmad=false (no FMA contraction, like gfortran on x86-64),
IEEE division / sqrt and no flush-to-zero (nvcc defaults, restated explicitly).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkidmp.so")
SOURCES = ["kidmp_api.cu"]
DEPS = ["kidmp_api.cu", "kidmp_column.cuh", "kidmp_tables.cuh", "kidmp_math.cuh", "kidmp_fastmath.h", "kidmp_kid.cuh", "kidmp_wrf.cuh", "kidmp_cells.cuh", "kidmp_hostinit.h",
        "kidmp_internal.h", os.path.join("..", "..", "include", "kidmp.h")]


def nvcc_path():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def source_id():
    """SHA-256 over every source the library is built from (DEPS, in order) and the flags: the build id."""
    import hashlib
    h = hashlib.sha256()
    for d in DEPS:
        with open(os.path.join(CSRC, d), "rb") as f:
            h.update(d.encode() + b"\0" + f.read())
    h.update(" ".join(BASE_FLAGS).encode())
    return h.hexdigest()[:32]


def library_id(path=None):
    """The build id embedded in a built library (kidmp_build_id()), read from the file without loading it."""
    import re
    try:
        with open(path or LIB, "rb") as f:
            m = re.search(rb"KIDMP_BUILD_ID=([0-9a-f]{32})", f.read())
        return m.group(1).decode() if m else None
    except OSError:
        return None


BASE_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
              "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fno-fast-math", "-shared", "-lcudart", "-ldl"]


def flags(extra=()):
    return [*BASE_FLAGS, "-DKIDMP_BUILD_ID=\"%s\"" % source_id(), *extra]


def needs_build():
    """The library is tied to the tree by its build id, not by file times: a shipped binary of other sources is rebuilt."""
    return library_id() != source_id()


def build(force=False, verbose=False, extra=(), out=None):
    """Compile if any source is newer than the library.  Returns the library path.
    `extra` / `out`: tuning builds with other -D flags into another file (tools/variants.py; loaded through KIDMP_LIB)."""
    if out is None and not force and not needs_build():
        return LIB
    out = out or LIB
    tmp = "%s.tmp%d" % (out, os.getpid())          # several ranks may build at once: each writes its own file, the rename is atomic
    cmd = [nvcc_path(), *flags(extra), "-o", tmp, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    try:
        subprocess.check_call(cmd)
        os.replace(tmp, out)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    return out


HOST_DIR = os.path.join(HERE, "host")
HOST_EXE = os.path.join(HOST_DIR, "kid_driver")


def build_host(force=False):
    """g++ build of the C++ twin of KiD's interface module and its little driver (kid_b200/host/)."""
    srcs = [os.path.join(HOST_DIR, f) for f in ("kid_driver.cpp", "mphys_thompson09n.cpp")]
    deps = srcs + [os.path.join(HOST_DIR, "mphys_thompson09n.hpp"), LIB]
    if not force and os.path.exists(HOST_EXE) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_EXE) for d in deps):
        return HOST_EXE
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", HOST_EXE, *srcs, "-L" + HERE, "-lkidmp",
                           "-Wl,-rpath,$ORIGIN/.."])
    return HOST_EXE


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    build_host(force=True)
    print(LIB)
