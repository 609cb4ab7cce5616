"""Micro-benchmark: batched CKKSEncoder::encode / decode on the device.
usage: python profiles/encoder_microbench.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "seal-fyp-logistic-regression_b200"
pkg = importlib.import_module(PKG)
eng = pkg.load_engine()
params = importlib.import_module(PKG + ".params")
for log_n, bits, batch in ((14, [60, 40, 40, 60], 128), (15, [60] + [40] * 8 + [60], 64)):
    primes = params.coeff_modulus_create(log_n, bits)
    ctx = eng.Context(log_n, primes)
    ev = eng.Evaluator(ctx)
    L, n = len(primes) - 1, 1 << log_n
    x = torch.rand((batch, n // 2), dtype=torch.float64, device="cuda")
    scale = 2.0 ** 40
    pt = ev.encode(x, scale)
    back = ev.decode(pt)
    assert (back - x).abs().max().item() < 2.0 ** -20
    for name, fn in (("encode", lambda: ev.encode(x, scale)), ("decode", lambda: ev.decode(pt))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("N=%d L=%d batch=%d %s: %.3f ms per batch, %.0f plaintexts/s, %.0f GB/s of plaintext limbs" % (
            n, L, batch, name, ms, batch / ms * 1e3, batch * L * n * 8 / ms / 1e6))
