"""Profiling target for the element-wise kernels and rescale (op_rooflines of bench.py): add, multiply_plain, multiply,
rescale_to_next at N = 32768, L = 9, batch 128 (larger than L2), one call each between cudaProfilerStart/Stop.
usage: ncu --set full --profile-from-start off ... python profiles/ew_microbench.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "seal-fyp-logistic-regression_b200"
eng = importlib.import_module(PKG).load_engine()
params = importlib.import_module(PKG + ".params")
ctx = eng.Context(15, params.coeff_modulus_create(15, [60] + [40] * 8 + [60]))
ev = eng.Evaluator(ctx)
L, batch = ctx.top_limbs, 128
a = ctx.empty(batch, 2, L)
a.data.random_(0, 1 << 39)
b = a.clone()
pt = ctx.empty(batch, 1, L)
pt.data.random_(0, 1 << 39)
o2, o3 = a.like(), a.like(size=3)
ops = [lambda: ev.add(a, b, out=o2), lambda: ev.multiply_plain(a, pt, out=o2), lambda: ev.multiply(a, b, out=o3),
       lambda: ev.rescale_to_next(a, out=o2)]
for f in ops:
    a.scale = b.scale = pt.scale = 1.0
    f()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for f in ops:
    a.scale = b.scale = pt.scale = 1.0
    f()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
