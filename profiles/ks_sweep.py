"""Key-switch timing sweep (one process, CUDA events): batched Galois key switch and the fused rotate-and-sum chain
at the shapes the bench and the north-star metric name, over several batch sizes.  Prints one JSON line per case;
used to A/B engine variants selected by environment variables (CKKS_FUSE, CKKS_NO_PDL, ...).
usage: python profiles/ks_sweep.py [--quick] [--tag NAME] [--check]"""
import importlib
import json
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

quick = "--quick" in sys.argv
tag = sys.argv[sys.argv.index("--tag") + 1] if "--tag" in sys.argv else os.environ.get("CKKS_VARIANT", "default")
HBM = 6537.0


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


def digest(t):
    return int(t.data.view(-1)[:: max(1, t.data.numel() // 4096)].sum().item()) & 0xFFFFFFFFFFFF


def run(log_n, top, cases):
    primes = params.coeff_modulus_create(log_n, [60] + [40] * (top - 1) + [60])
    ctx = eng.Context(log_n, primes)
    ev = eng.Evaluator(ctx)
    keys = client.KeyGenerator(ctx, seed=1).keyset(steps=[1])
    g = ctx.galois_elt(1)
    for L, batch, reps in cases:
        gen = torch.Generator(device="cuda").manual_seed(1234 + L + batch)
        a = ctx.empty(batch, 2, L, cap=top)
        a.data.random_(0, 1 << 39, generator=gen)
        b = a.like()
        us = timed(lambda: ev.apply_galois(a, g, keys, out=b), reps)
        alg = batch * (2 * L * L + 6 * L) * 8 * ctx.n
        rec = {"tag": tag, "op": "apply_galois", "N": ctx.n, "L": L, "batch": batch, "us": round(us, 1),
               "ks_per_s": round(batch / us * 1e6), "hbm_frac": round(alg / us / 1e3 / HBM, 4), "digest": digest(b)}
        print(json.dumps(rec), flush=True)
        # dependent chain: 64 steps of rotate + add (graph replay), per-step latency
        dup, acc = a.clone(), a.clone()
        steps = 64
        us = timed(lambda: ev.rotate_sum_chain(dup, acc, 1, steps, keys), max(2, reps // 8)) / steps
        rec = {"tag": tag, "op": "chain_step", "N": ctx.n, "L": L, "batch": batch, "us": round(us, 1),
               "ks_per_s": round(batch / us * 1e6), "hbm_frac": round(alg / us / 1e3 / HBM, 4), "digest": digest(acc)}
        print(json.dumps(rec), flush=True)
        # relinearize
        a3 = ctx.empty(batch, 3, L, cap=top)
        a3.data.random_(0, 1 << 39, generator=gen)
        us = timed(lambda: ev.relinearize(a3, keys, out=b), reps)
        algr = batch * (2 * L * L + 7 * L) * 8 * ctx.n
        rec = {"tag": tag, "op": "relinearize", "N": ctx.n, "L": L, "batch": batch, "us": round(us, 1),
               "ks_per_s": round(batch / us * 1e6), "hbm_frac": round(algr / us / 1e3 / HBM, 4), "digest": digest(b)}
        print(json.dumps(rec), flush=True)
        # rescale (S = 2)
        if L > 1:
            o = a.like()
            us = timed(lambda: ev.rescale_to_next(a, out=o), reps)
            algs = batch * 8 * 2 * (2 * L - 1) * ctx.n
            rec = {"tag": tag, "op": "rescale", "N": ctx.n, "L": L, "batch": batch, "us": round(us, 1),
                   "ops_per_s": round(batch / us * 1e6), "hbm_frac": round(algs / us / 1e3 / HBM, 4), "digest": digest(o)}
            print(json.dumps(rec), flush=True)
    del ctx


if "--small" in sys.argv:      # the per-GPU batches of the strong-scaling run (32 chains over 2 / 4 / 8 GPUs) and below
    run(15, 9, [(3, 16, 30), (3, 8, 30), (3, 4, 30), (3, 2, 30)])
elif quick:
    run(15, 9, [(3, 32, 20), (3, 4, 20), (9, 32, 10)])
else:
    run(15, 9, [(3, 32, 30), (3, 16, 30), (3, 8, 30), (3, 4, 30), (3, 1, 30), (9, 32, 10), (9, 4, 10)])
    run(14, 8, [(8, 64, 10), (3, 256, 10), (3, 4, 20)])
