"""In-tree build of libckks_b200.so (hand-written sm_100a kernels + C ABI).

nvcc cross-compiles without a GPU.  The shared object lands next to this file so it travels
to the GPU box with the repository snapshot; it is git-ignored.
"""
import fcntl
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libckks_b200.so")
SOURCES = ["engine.cu", "tables.cpp"]
HEADERS = ["kernels.cuh", "ntt_passes.cuh", "modarith.cuh", "encoder.cuh", "tables.h",
           os.path.join("..", "..", "include", "ckks_b200.h")]
STAMP = LIB + ".srchash"      # hash of everything the shared object was built from (travels with it)
LOCK = LIB + ".lock"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CKKS engine cannot be built")


def _source_hash():
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


def is_stale():
    """Staleness is decided by content, not by mtime: a snapshot copied to another machine (gpurun) does not
    keep the relative modification times of the sources and the shared object."""
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != _source_hash()


def build(force=False, verbose=False):
    """Compile the engine for sm_100a.  Returns the path of the shared object.

    Safe under concurrent callers (one process per GPU all importing the package): the build runs under an
    exclusive file lock, into a temporary file that is renamed over the target, so no process ever maps a
    half-written library."""
    if not force and not is_stale():
        return LIB
    with open(LOCK, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():      # another process built it while this one waited
                return LIB
            tmp = "%s.tmp.%d" % (LIB, os.getpid())
            cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
            res = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + res.stdout)
            os.replace(tmp, LIB)
            with open(STAMP + ".tmp", "w") as fh:
                fh.write(_source_hash() + "\n")
            os.replace(STAMP + ".tmp", STAMP)
            if verbose:
                print(res.stdout)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB
