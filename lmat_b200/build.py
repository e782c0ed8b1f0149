"""Build libkmat.so (CUDA, sm_100a) and the read_label host binary in-tree with nvcc/g++.

`python -m lmat_b200.build` or lmat_b200.build.build_all().  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkmat.so")
BIN = os.path.join(HERE, "bin", "read_label")

CUDA_SRCS = ["kmat_db.cu", "kmat_label.cu"]
HOST_SRCS = ["kmat_host.cpp", "kmat_reader.cpp", "kmat_build.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _deps():
    out = [os.path.join(ROOT, "include", "kmat.h")]
    for f in os.listdir(CSRC):
        if f.endswith((".h", ".cuh", ".cu", ".cpp")):
            out.append(os.path.join(CSRC, f))
    return out


def build_lib(force=False, verbose=False):
    if not force and not _newer(LIB, _deps()):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    log = []
    for src in CUDA_SRCS:
        obj = os.path.join(HERE, "build", src + ".o")
        # KMAT_NVCC_DEFINES="-DKMAT_K4_PACKED_DEPTH=1 ...": compile-time experiments (see the #ifndef blocks in csrc/)
        extra = os.environ.get("KMAT_NVCC_DEFINES", "").split()
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-I" + os.path.join(ROOT, "include"), "-c", os.path.join(CSRC, src), "-o", obj]
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log.append(p.stdout)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{p.stdout}")
        objs.append(obj)
    for src in HOST_SRCS:
        obj = os.path.join(HERE, "build", src + ".o")
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-c", os.path.join(CSRC, src), "-o", obj]
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log.append(p.stdout)
        if p.returncode != 0:
            raise RuntimeError(f"g++ failed for {src}:\n{p.stdout}")
        objs.append(obj)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-lz", "-Xcompiler", "-fPIC"]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"link failed:\n{p.stdout}")
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


def build_cli(force=False):
    src = os.path.join(CSRC, "read_label_main.cpp")
    if not os.path.exists(src):
        return None
    if not force and not _newer(BIN, [src, LIB]):
        return BIN
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), src, "-o", BIN, "-L" + HERE, "-lkmat",
           "-Wl,-rpath,$ORIGIN/..", "-lpthread"]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"g++ failed for read_label_main.cpp:\n{p.stdout}")
    return BIN


MDT_BIN = os.path.join(HERE, "bin", "make_db_table")


GL_BIN = os.path.join(HERE, "bin", "gene_label")
RRL_BIN = os.path.join(HERE, "bin", "rand_read_label")
CS_BIN = os.path.join(HERE, "bin", "content_summ")


def build_tools(force=False):
    """The other host binaries over libkmat (drop-ins for the reference tools of the same name)."""
    for name, out in (("make_db_table_main.cpp", MDT_BIN), ("gene_label_main.cpp", GL_BIN), ("rand_read_label_main.cpp", RRL_BIN),
                      ("content_summ_main.cpp", CS_BIN)):
        src = os.path.join(CSRC, name)
        if force or _newer(out, [src, LIB]):
            os.makedirs(os.path.dirname(out), exist_ok=True)
            cmd = ["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), src, "-o", out, "-L" + HERE, "-lkmat",
                   "-Wl,-rpath,$ORIGIN/..", "-lpthread", "-lz"]
            p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if p.returncode != 0:
                raise RuntimeError(f"g++ failed for {name}:\n{p.stdout}")
    return MDT_BIN


def build_all(force=False, verbose=False):
    lib = build_lib(force=force, verbose=verbose)
    cli = build_cli(force=force)
    build_tools(force=force)
    return lib, cli


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose=True))
