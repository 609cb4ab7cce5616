"""Mini-workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every kernel family of the engine
once, at small degrees, both arithmetic paths (50-/60-bit integer limbs and 40-bit FP64 limbs).
usage: compute-sanitizer --tool racecheck python profiles/sanitize_driver.py [LOG_N ...]"""
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
client = importlib.import_module(PKG + ".client")

for log_n in [int(a) for a in sys.argv[1:]] or [12, 13]:
    primes = params.coeff_modulus_create(log_n, [60, 40, 40, 60])
    ctx = eng.Context(log_n, primes)
    ev = eng.Evaluator(ctx)
    kg = client.KeyGenerator(ctx, seed=1)
    steps = [1, 2, 4, -1]
    keys = kg.keyset(steps=steps)
    enc = client.Encryptor(ctx, kg.public_key(), seed=2)
    dec = client.Decryptor(ctx, kg.secret_key())
    cod = client.CKKSEncoder(ctx)
    scale = 2.0 ** 40
    slots = ctx.n // 2
    rng = np.random.default_rng(log_n)
    x = rng.uniform(-1, 1, (3, slots))
    ct = enc.encrypt(cod.encode(x, scale))                       # encoder, sampler, element-wise, forward NTT
    pt = cod.encode(rng.uniform(-1, 1, (3, slots)), scale)
    prod = ev.rescale_to_next(ev.relinearize(ev.multiply(ct, ct), keys))     # multiply, relinearize, rescale
    pp = ev.rescale_to_next(ev.multiply_plain(ct, pt))
    s = ev.add(prod, pp)
    rot = ev.rotate_vector(ct, 3, keys)                          # NAF rounds (1 + 2)
    plan = eng.RotPlan(ctx, keys, [1, 2, 3, -1, 5])
    one = ct[0:1]
    r0 = ev.rotate_vector(one, 5, keys)                          # single ciphertext: thread-block-cluster inner product
    q0 = ev.relinearize(ev.multiply(one, one), keys)
    r1 = ev.rotate_plan(one, plan)                               # batched plan, shared NAF prefixes
    hplan = eng.RotPlan(ctx, keys, [1, 2, 4, -1])
    r2 = ev.rotate_plan_hoisted(one, hplan)                      # hoisted kernels
    dup, acc = ct.clone(), ct.clone()
    ev.rotate_sum_chain(dup, acc, 1, 19, keys)                   # graph-replayed chain, fused add
    ms = ev.multiply_plain_sum(ct, pt)
    mm = ev.multiply_sum(ct, ct)
    low = ev.rotate_vector(s, 1, keys)                           # key switch below the top level
    out = cod.decode(dec.decrypt(low))
    torch.cuda.synchronize()
    got = np.asarray(out)[:, :8].real
    print("log_n", log_n, "launches", ctx.launch_count(), "decoded sample", got[0, :3])
print("sanitize driver done")
