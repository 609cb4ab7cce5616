"""Times the BASELINE.json configurations other than the bench workload on one B200 and writes
gpurun_out/r02_configs.json (GPU wall times with CUDA events, after warm-up; decrypted results checked
against plaintext math).  usage: python profiles/run_configs.py [--skip-matmul64]"""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "seal-fyp-logistic-regression_b200"
pkg = importlib.import_module(PKG)
eng = pkg.load_engine()
params = importlib.import_module(PKG + ".params")
client = importlib.import_module(PKG + ".client")
wl = importlib.import_module(PKG + ".workloads")
SCALE = 2.0 ** 40


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def setup(log_n, bits, steps, seed=1):
    ctx = eng.Context(log_n, params.coeff_modulus_create(log_n, bits))
    ev = eng.Evaluator(ctx)
    enc = client.CKKSEncoder(ctx)
    kg = client.KeyGenerator(ctx, seed=seed)
    keys = kg.keyset(steps=steps)
    encr = client.Encryptor(ctx, kg.public_key(), seed=seed + 1)
    decr = client.Decryptor(ctx, kg.secret_key())
    return ctx, ev, enc, keys, encr, decr


def config1(results):
    """config 1: the reference's own LR program (row layout, degree-3 Horner sigmoid, lr 0.1,
    logistic_regression_ckks.cpp:208-345) with repairs R1-R6 on the first 2000 rows of pulsar_stars.csv (tests/golden copy),
    standardised with the reference's scaler, the reference program's initial weights; N = 32768, {60, 40 x 8, 60}"""
    lrm = importlib.import_module(PKG + ".lr")
    pulsar = importlib.import_module(PKG + ".pulsar")
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "pulsar_plain_lr.json")))
    pow2 = [s for i in range(13) for s in (1 << i, -(1 << i))]
    ctx, ev, enc, keys, encr, decr = setup(15, [60] + [40] * 8 + [60], pow2, seed=7)
    X, y = pulsar.load_csv()
    X, y = pulsar.standard_scaler(X).astype(np.float64), y.astype(np.float64)
    R, C = X.shape
    w0 = np.array(gold["initial_weights"])
    lay = lrm.RowLayout(R, C, ctx.n // 2)
    rows = encr.encrypt(enc.encode(lay.rows(X), SCALE))
    cols = encr.encrypt(enc.encode(lay.columns(X), SCALE))
    labs = encr.encrypt(enc.encode(lay.labels(y), SCALE))
    wct = encr.encrypt(enc.encode(lay.weights(w0), SCALE))
    ctx.reset_launch_count()
    ms, neww = timed(lambda: lrm.update_weights(ev, rows, cols, labs, wct, 0.1, SCALE, keys, enc, encr, degree=3,
                                                method="horner"), reps=1, warm=1)
    launches = ctx.launch_count() // 2
    got = enc.decode(decr.decrypt(neww))[0, :C]
    want = lrm.plain_epoch(X, y, w0, 0.1, 3)
    err = float(np.abs(got - want).max())
    results["config1_lr_row_layout_pulsar_R%d_N32768" % R] = {
        "ms_per_iteration": ms, "iterations_per_s": 1e3 / ms, "kernel_launches": int(launches), "max_abs_err_vs_plain_lr": err,
        "key_switches": R * (1 + 1 + C - 1) + C * (1 + 13 + R - 1) + 3,
        "cost_after_step": pulsar.cost_function(X.astype(np.float32), y, got),
        "reference_program_cost_after_iteration_0": gold["cost_after_iteration_0"],
        "note": "one update_weights = one training iteration over all R rows; includes encoding the R one-hot masks on the device"}
    print("config1 R=%d: %.1f ms per iteration, err %.2e" % (R, ms, err), flush=True)


def config3(results):
    pow2 = [s for i in range(13) for s in (1 << i, -(1 << i))]
    ctx, ev, enc, keys, encr, decr = setup(14, [60, 40, 40, 60], pow2)
    plans = wl.PlanCache(ctx, keys)
    rng = np.random.default_rng(3)
    for d in (64, 128):
        U, v = rng.uniform(0, 1, (d, d)), rng.uniform(0, 1, d)
        ct = encr.encrypt(enc.encode(v, SCALE))
        diags = enc.encode(wl.all_diagonals(U), SCALE)
        ms, out = timed(lambda: wl.linear_transform_plain(ev, ct, diags, keys, plans))
        err = float(np.abs(enc.decode(decr.decrypt(out))[0, :d] - U @ v).max())
        ks = plans.get(range(d)).keyswitches_shared + 1
        results["config3_linear_transform_d%d_N16384" % d] = {"ms": ms, "key_switches": ks, "max_abs_err": err,
                                                              "key_switches_without_prefix_sharing": plans.get(range(d)).keyswitches + 1}
        print("config3 d=%d: %.2f ms (%d key switches), err %.2e" % (d, ms, ks, err), flush=True)


def config4(results, d):
    pow2 = [s for i in range(13) for s in (1 << i, -(1 << i))]
    ctx, ev, enc, keys, encr, decr = setup(14, [60, 40, 40, 40, 40, 60], pow2, seed=5)
    plans = wl.PlanCache(ctx, keys)
    rng = np.random.default_rng(4)
    A, B = rng.uniform(0, 1, (d, d)), rng.uniform(0, 1, (d, d))
    eps = 1e-8
    t0 = time.time()
    mk = lambda U: wl.DiagonalSet.from_matrix(U, eps, SCALE, enc)
    sigma, tau = mk(wl.u_sigma(d)), mk(wl.u_tau(d))
    V = [mk(wl.v_k(d, k)) for k in range(1, d)]
    W = [mk(wl.w_k(d, k)) for k in range(1, d)]
    t_diag = time.time() - t0
    ctA = encr.encrypt(enc.encode(A.reshape(-1), SCALE))     # already in C_Matrix_Encode (row-major) form
    ctB = encr.encrypt(enc.encode(B.reshape(-1), SCALE))
    ms, out = timed(lambda: wl.cc_matrix_multiplication_sparse(ev, ctA, ctB, d, sigma, tau, V, W, keys, plans), reps=1, warm=1)
    got = enc.decode(decr.decrypt(out))[0, : d * d].reshape(d, d)
    err = float(np.abs(got - A @ B).max())
    ks = 4 * (plans.get(range(d * d)).keyswitches_shared + 1)
    ks_plain = 4 * (plans.get(range(d * d)).keyswitches + 1)
    ctx.reset_launch_count()
    ms_nz, out_nz = timed(lambda: wl.cc_matrix_multiplication_nonzero(ev, ctA, ctB, d, sigma, tau, V, W, keys, plans), reps=3, warm=1)
    launches_nz = ctx.launch_count() // 4
    err_nz = float(np.abs(enc.decode(decr.decrypt(out_nz))[0, : d * d].reshape(d, d) - A @ B).max())
    steps_nz = [sigma.index, tau.index, sorted({l for s in V for l in s.index}), sorted({l for s in W for l in s.index})]
    ks_nz = sum(plans.get(st).keyswitches_shared for st in steps_nz) + 4
    results["config4_matrix_multiplication_d%d_N16384" % d] = {
        "ms": ms, "galois_key_switches": ks, "galois_key_switches_without_prefix_sharing": ks_plain,
        "reference_key_switches": (2 + 2 * (d - 1)) * (ks_plain // 4),
        "max_abs_err": err, "host_diagonal_setup_s": t_diag,
        "nonzero_diagonals_mode": {"ms": ms_nz, "galois_key_switches": int(ks_nz), "kernel_launches": int(launches_nz), "max_abs_err": err_nz,
                                   "note": "SURVEY 8(f4) tolerance mode: the all-epsilon diagonals are skipped; ciphertexts differ from the "
                                           "reference sequence, the decrypted product is closer to A @ B (no epsilon error term)"}}
    print("config4 d=%d non-empty diagonals only: %.2f ms (%d key switches), err %.2e" % (d, ms_nz, ks_nz, err_nz), flush=True)
    print("config4 d=%d: %.1f ms (%d key switches on device; the reference performs %d), err %.2e" % (
        d, ms, ks, (2 + 2 * (d - 1)) * (ks_plain // 4), err), flush=True)


def config2(results):
    for log_n in (12, 13, 14):
        primes = params.bfv_default(log_n)
        ctx = eng.Context(log_n, primes)
        ev = eng.Evaluator(ctx)
        keys = client.KeyGenerator(ctx, seed=2).keyset(steps=[1, -1, -4, 16])
        L = ctx.top_limbs
        for batch in (1, 1024):
            a = ctx.empty(batch, 2, L)
            a.data.random_(0, 1 << 35)
            b = a.clone()
            pt = ctx.empty(1, 1, L)
            pt.data.random_(0, 1 << 35)
            a3 = ctx.empty(batch, 3, L)
            a3.data.random_(0, 1 << 35)
            o2, o3, sc = a.like(), a3.like(), a.like()
            ops = {
                "multiply_plain": lambda: ev.multiply_plain(a, pt, out=o2),
                "multiply+relinearize": lambda: ev.relinearize(ev.multiply(a, b, out=o3), keys, out=o2),
                "rotate_vector(1)": lambda: ev.rotate_vector(a, 1, keys, out=o2),
                "rotate_vector(11)": lambda: ev.rotate_vector(a, 11, keys, out=o2, scratch=sc),
                "rescale_to_next": lambda: ev.rescale_to_next(a, out=o2),
            }
            for name, fn in ops.items():
                a.scale = b.scale = pt.scale = a3.scale = 1.0
                ms, _ = timed(fn, reps=10, warm=2)
                results.setdefault("config2_op_sweep", {})["N=%d,L=%d,batch=%d,%s" % (ctx.n, L, batch, name)] = {
                    "us_per_call": ms * 1e3, "ops_per_s": batch / (ms * 1e-3)}
        print("config2 N=%d done" % ctx.n, flush=True)


def main():
    results = {}
    if "--only-config1" in sys.argv:
        config1(results)
        json.dump(results, open(os.path.join(ROOT, "gpurun_out", "r02_config1.json"), "w"), indent=1)
        return
    config1(results)
    config2(results)
    config3(results)
    config4(results, 5)
    if "--skip-matmul64" not in sys.argv:
        config4(results, 64)
    out = os.path.join(ROOT, "gpurun_out", "r02_configs.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(results, open(out, "w"), indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
