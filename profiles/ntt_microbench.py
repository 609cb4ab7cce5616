"""Standalone batched NTT / INTT throughput (the (a) kernels of the north star) for a 60-bit prime
(integer butterflies) and a 40-bit prime (FP64 butterflies).  Algorithmic bytes = 16 N per limb."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "seal-fyp-logistic-regression_b200"
pkg = importlib.import_module(PKG)
eng = pkg.load_engine()
params = importlib.import_module(PKG + ".params")

for log_n in ((14,) if os.environ.get("CKKS_NTT_LIMB") else (14, 15)):
    primes = params.coeff_modulus_create(log_n, [60, 40, 40, 60])
    ctx = eng.Context(log_n, primes)
    ev = eng.Evaluator(ctx)
    n = ctx.n
    limbs = (1 << 31) // (n * 8) // 2          # 1 GiB of limbs: larger than L2
    for first_prime, name in ((0, "60-bit prime (integer path)"), (1, "40-bit prime (FP64 path)")):
        t = torch.randint(0, 1 << 39, (limbs, 1, n), dtype=torch.int64, device="cuda")
        for direction, fn in (("forward", ev.ntt_forward), ("inverse", ev.ntt_inverse)):
            for _ in range(2):
                fn(t, first_prime=first_prime)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            reps = 5
            for _ in range(reps):
                fn(t, first_prime=first_prime)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            gbs = limbs * 16 * n / ms / 1e6
            print("N=%d %s %s: %d limbs in %.3f ms = %.2f M limbs/s, %.0f GB/s algorithmic (%.2f of 6537)" % (
                n, name, direction, limbs, ms, limbs / ms / 1e3, gbs, gbs / 6537))
