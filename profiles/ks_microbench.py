"""Micro-benchmark / profiling target: batched Galois key switch (the dominant op of the LR epoch).
usage: python profiles/ks_microbench.py [log_n] [data_limbs_total K-1] [level L] [batch] [reps]"""
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
client = importlib.import_module(PKG + ".client")

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 15
top = int(sys.argv[2]) if len(sys.argv) > 2 else 9
L = int(sys.argv[3]) if len(sys.argv) > 3 else 3
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 32
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 20

primes = params.coeff_modulus_create(log_n, [60] + [40] * (top - 1) + [60])
ctx = eng.Context(log_n, primes)
if len(sys.argv) > 6:
    ctx.set_workspace_cap(int(float(sys.argv[6]) * (1 << 20)))
ev = eng.Evaluator(ctx)
keys = client.KeyGenerator(ctx, seed=1).keyset(steps=[1])
a = ctx.empty(batch, 2, L, cap=top)
a.data.random_(0, 1 << 39)
b = a.like()
g = ctx.galois_elt(1)
for _ in range(3):
    ev.apply_galois(a, g, keys, out=b)
torch.cuda.synchronize()
torch.cuda.profiler.start()          # ncu --profile-from-start off captures only this region
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ev.apply_galois(a, g, keys, out=b)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
ms = e0.elapsed_time(e1) / reps
alg = batch * (2 * L * L + 6 * L) * 8 * ctx.n
print("N=%d L=%d batch=%d: %.1f us per batched key switch, %.0f ks/s, %.0f GB/s algorithmic" % (
    ctx.n, L, batch, ms * 1e3, batch / ms * 1e3, alg / ms / 1e6))
