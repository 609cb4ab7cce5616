"""CPU: the C-ABI library builds, loads and exports every symbol include/ckks_b200.h declares.
No compute call is made here (there is no GPU in this container)."""
import ctypes
import importlib
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "seal-fyp-logistic-regression_b200"


def _declared():
    text = open(os.path.join(ROOT, "include", "ckks_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ckks_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    capi = importlib.import_module(PKG + ".capi")
    lib = capi.load()
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), n
    # the Python binding lists exactly the declared functions
    assert sorted(capi.SIGNATURES) == names


def test_no_device_fails_loudly(pkg):
    """without a CUDA device the engine refuses to create a context (no CPU fallback)"""
    import torch
    if torch.cuda.is_available():
        return
    capi = importlib.import_module(PKG + ".capi")
    lib = capi.load()
    primes = (ctypes.c_uint64 * 2)(0xffffee001, 0xffffc4001)
    h = ctypes.c_void_p()
    rc = lib.ckks_ctx_create(12, 2, primes, 0, ctypes.byref(h))
    assert rc == capi.CKKS_ERR_CUDA
    assert b"no CUDA device" in lib.ckks_last_error() or b"cuda" in lib.ckks_last_error().lower()


def test_invalid_parameters_rejected(pkg):
    capi = importlib.import_module(PKG + ".capi")
    lib = capi.load()
    h = ctypes.c_void_p()
    bad = (ctypes.c_uint64 * 2)(0xffffee001, 0xffffee003)   # second is not an NTT prime
    assert lib.ckks_ctx_create(12, 2, bad, 0, ctypes.byref(h)) == capi.CKKS_ERR_INVALID
    ok = (ctypes.c_uint64 * 2)(0xffffee001, 0xffffc4001)
    assert lib.ckks_ctx_create(11, 2, ok, 0, ctypes.byref(h)) == capi.CKKS_ERR_INVALID
    dup = (ctypes.c_uint64 * 2)(0xffffee001, 0xffffee001)
    assert lib.ckks_ctx_create(12, 2, dup, 0, ctypes.byref(h)) == capi.CKKS_ERR_INVALID


def test_product_does_not_import_oracle():
    """the product path must never route through oracle/ (test infrastructure)"""
    pat = re.compile(r"(from\s+oracle|import\s+oracle|pyoracle|ckks_oracle|oracle/|liboracle|orc_[a-z]+\()")
    pkg_dir = os.path.join(ROOT, PKG)
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), (dirpath, f)
