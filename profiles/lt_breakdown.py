"""Where the sharded Linear_Transform_Plain (config 3, d = 128, N = 16384) spends its time on ONE rank of an 8-GPU run:
times the rank's pieces separately on one GPU (no collective).  usage: python profiles/lt_breakdown.py [world=8]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "seal-fyp-logistic-regression_b200"
eng = importlib.import_module(PKG).load_engine()
params = importlib.import_module(PKG + ".params")
client = importlib.import_module(PKG + ".client")
wl = importlib.import_module(PKG + ".workloads")
par = importlib.import_module(PKG + ".parallel")
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
d, SCALE = 128, 2.0 ** 40
ctx = eng.Context(14, params.coeff_modulus_create(14, [60, 40, 40, 60]))
ev = eng.Evaluator(ctx)
enc = client.CKKSEncoder(ctx)
kg = client.KeyGenerator(ctx, seed=77)
keys = kg.keyset(steps=[s for i in range(8) for s in (1 << i, -(1 << i))])
encr = client.Encryptor(ctx, kg.public_key(), seed=78)
rng = np.random.default_rng(79)
U, v = rng.uniform(0, 1, (d, d)), rng.uniform(0, 1, d)
ct = encr.encrypt(enc.encode(v, SCALE))
diags = enc.encode(wl.all_diagonals(U), SCALE)
plans = wl.PlanCache(ctx, keys)
mine = par.shard_rotations_shared(list(range(d)), 0, world)
plan = plans.get(mine)
dl = eng.Ciphertext(ctx, diags.data[torch.tensor(mine, device=ctx.device)].contiguous(), diags.limbs, diags.scale)


def timed(fn, reps=50):
    for _ in range(5):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, out


t_dup, dup = timed(lambda: wl.duplicate_fill(ev, ct, d, keys))
t_rot, rots = timed(lambda: ev.rotate_plan(dup, plan))
t_mps, part = timed(lambda: ev.multiply_plain_sum(rots, dl))
g = eng.Ciphertext(ctx, part.data.repeat(world, 1, 1, 1), part.limbs, part.scale)
t_add, _ = timed(lambda: ev.add_many(g))
t_all, _ = timed(lambda: ev.multiply_plain_sum(ev.rotate_plan(wl.duplicate_fill(ev, ct, d, keys), plan), dl))
ctx.reserve(len(mine), ctx.top_limbs)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    out = ev.multiply_plain_sum(ev.rotate_plan(wl.duplicate_fill(ev, ct, d, keys), plan), dl)
t_graph, _ = timed(graph.replay)
print("rank 0 of %d: %d diagonals, %d key switches in %d rounds (round sizes via plan)" % (world, len(mine), plan.keyswitches_shared, plan.rounds))
print("duplicate_fill (1 key switch + add) %.1f us | rotate_plan %.1f us | multiply_plain_sum %.1f us | add_many over %d partials %.1f us" % (
    t_dup, t_rot, t_mps, world, t_add))
print("local part end to end: eager %.1f us, one replayed CUDA graph %.1f us" % (t_all, t_graph))
