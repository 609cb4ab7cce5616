import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "seal-fyp-logistic-regression_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def po():
    """the CPU oracle binding (test infrastructure)"""
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def pkg():
    p = importlib.import_module(PKG)
    p.build()
    return p


@pytest.fixture(scope="session")
def eng(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return pkg.load_engine()


class Fixture:
    """one parameter set with oracle-generated keys mirrored on the device"""

    def __init__(self, po, eng, log_n, bits, steps=(1, -1, 2, 4, -8), seed=100):
        import numpy as np
        self.np = np
        self.po, self.eng = po, eng
        self.log_n = log_n
        self.primes = po.coeff_modulus_create(log_n, list(bits)) if not isinstance(bits[0], int) or bits[0] < 64 else list(bits)
        self.orc = po.Oracle(log_n, self.primes)
        self.ctx = eng.Context(log_n, self.primes)
        self.ev = eng.Evaluator(self.ctx)
        self.sk = self.orc.gen_secret(seed)
        self.pk = self.orc.gen_public(seed + 1, self.sk)
        self.rlk = self.orc.gen_relin_key(seed + 2, self.sk)
        self.gks = self.orc.gen_galois_keys(seed + 3, self.sk, steps=steps)
        self.keys = eng.KeySet(self.ctx)
        self.keys.set_relin(self.ctx.upload_key(self.rlk))
        for g, k in self.gks.items():
            self.keys.set_galois(g, self.ctx.upload_key(k))
        self.n = 1 << log_n
        self.L = len(self.primes) - 1

    def random_ct(self, rng, batch, size, limbs):
        """uniform random residues: a valid input for bit-exactness checks of any evaluator op"""
        np = self.np
        out = np.empty((batch, size, limbs, self.n), dtype=np.uint64)
        for j in range(limbs):
            out[:, :, j, :] = rng.integers(0, self.primes[j], size=(batch, size, self.n), dtype=np.uint64)
        return out


@pytest.fixture(scope="session")
def make_fixture(po, eng):
    cache = {}

    def get(log_n, bits, steps=(1, -1, 2, 4, -8)):
        key = (log_n, tuple(bits), tuple(steps))
        if key not in cache:
            cache[key] = Fixture(po, eng, log_n, bits, steps)
        return cache[key]

    return get
