"""The in-tree build must be safe when several processes need the library at once (one process per GPU all
import the package, bench.py under torchrun): content-hash staleness, an exclusive lock and an atomic rename.
Regression test for 'libckks_b200.so: file too short' seen on an 8-GPU run."""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "seal-fyp-logistic-regression_b200"

WORKER = r"""
import ctypes, importlib, sys
sys.path.insert(0, %r)
b = importlib.import_module(%r + "._build")
path = b.build()
lib = ctypes.CDLL(path)            # a half-written file would fail here
lib.ckks_version.restype = ctypes.c_char_p
print(lib.ckks_version().decode())
"""


def test_staleness_is_decided_by_content():
    b = importlib.import_module(PKG + "._build")
    b.build()
    assert not b.is_stale()
    # touching a source (what a snapshot copy to another machine does) must not trigger a rebuild
    src = os.path.join(b.CSRC, "engine.cu")
    os.utime(src, None)
    assert not b.is_stale()
    assert "encoder.cuh" in b.HEADERS and "kernels.cuh" in b.HEADERS


def test_concurrent_builders_never_see_a_partial_library():
    b = importlib.import_module(PKG + "._build")
    b.build()
    os.remove(b.STAMP)                      # every worker now finds the library stale
    assert b.is_stale()
    procs = [subprocess.Popen([sys.executable, "-c", WORKER % (ROOT, PKG)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for _ in range(3)]
    outs = [p.communicate(timeout=900)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ckks_b200" in o for o in outs), outs
    assert not b.is_stale()
    assert not [f for f in os.listdir(b.HERE) if ".so.tmp." in f]
