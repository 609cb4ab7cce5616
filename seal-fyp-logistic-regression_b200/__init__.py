"""B200-native CKKS evaluation engine behind the SEAL Evaluator surface used by
MarwanNour/SEAL-FYP-Logistic-Regression (hot path only).

    csrc/        hand-written sm_100a kernels + the C ABI (include/ckks_b200.h)
    capi.py      ctypes declarations of the C ABI
    engine.py    batched Python mirror of the SEAL Evaluator members (torch = plumbing)

The package directory name contains '-', so import it with
`importlib.import_module("seal-fyp-logistic-regression_b200")`.
"""
from ._build import build, LIB  # noqa: F401


def load_engine():
    """import the torch-backed host layer (needs libckks_b200.so to be built)"""
    from . import engine
    return engine
