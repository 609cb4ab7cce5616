#!/usr/bin/env python
"""Replay SEAL-produced test vectors on this engine and compare bit for bit.

    python tools/seal_replay.py DIR [--backend cuda|oracle|both]

DIR holds what tools/seal_dump_vectors.cpp wrote with a real SEAL build: parms.bin, relin_keys.bin, galois_keys.bin,
in_x.ct, in_y.ct, in_y.pt and the out_*.ct files.  For every op the inputs are loaded in SEAL's binary format
(sealio.py), evaluated here, and the resulting polynomials are compared with SEAL's own output word for word.

`--self-test` needs no SEAL: the CPU oracle plays SEAL's role, writes the same file set into a temporary directory
in SEAL's format and the replay then runs over those files -- it proves the tool chain (file format, key layout,
op sequence), not parity with SEAL."""
import argparse
import importlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "seal-fyp-logistic-regression_b200"
sio = importlib.import_module(PKG + ".sealio")


def _load(path, fn, *a):
    with open(path, "rb") as f:
        return fn(f, *a)


class OracleBackend:
    name = "oracle"

    def __init__(self, p, rlk, gks, rounding=1):
        from oracle import pyoracle as po
        self.o = po.Oracle(p.log_n, p.primes)
        self.o.set_rounding(rounding)
        self.rlk, self.gks = rlk, gks

    def add(self, a, b): return self.o.add(a, b)
    def sub(self, a, b): return self.o.sub(a, b)
    def multiply(self, a, b): return self.o.multiply(a, b)
    def multiply_plain(self, a, p): return self.o.multiply_plain(a, p)
    def relinearize(self, a): return self.o.relinearize(a, self.rlk)
    def rescale(self, a): return self.o.rescale(a)
    def rotate(self, a, steps): return self.o.rotate(a, steps, self.gks)
    def mod_switch(self, a): return np.ascontiguousarray(a[:, :-1])


class CudaBackend:
    name = "cuda"

    def __init__(self, p, rlk, gks, rounding=1):
        eng = importlib.import_module(PKG).load_engine()
        self.eng = eng
        self.ctx = eng.Context(p.log_n, p.primes)
        self.ctx.set_rounding(rounding)
        self.ev = eng.Evaluator(self.ctx)
        self.keys = eng.KeySet(self.ctx)
        self.keys.set_relin(self.ctx.upload_key(rlk))
        for g, k in gks.items():
            self.keys.set_galois(g, self.ctx.upload_key(k))
        self.top = self.ctx.top_limbs

    def _up(self, a):
        return self.ctx.upload(a, cap=self.top)

    def add(self, a, b): return self.ev.add(self._up(a), self._up(b)).numpy()[0]
    def sub(self, a, b): return self.ev.sub(self._up(a), self._up(b)).numpy()[0]
    def multiply(self, a, b): return self.ev.multiply(self._up(a), self._up(b)).numpy()[0]
    def multiply_plain(self, a, p): return self.ev.multiply_plain(self._up(a), self.ctx.upload_plain(p, cap=self.top)).numpy()[0]
    def relinearize(self, a): return self.ev.relinearize(self._up(a), self.keys).numpy()[0]
    def rescale(self, a): return self.ev.rescale_to_next(self._up(a)).numpy()[0]
    def rotate(self, a, steps): return self.ev.rotate_vector(self._up(a), steps, self.keys).numpy()[0]
    def mod_switch(self, a): return self.ev.mod_switch_to(self._up(a), a.shape[1] - 1).numpy()[0]


def replay(d, backends, rounding=1):
    p = _load(os.path.join(d, "parms.bin"), sio.load_params)
    n = p.n
    rlk = _load(os.path.join(d, "relin_keys.bin"), sio.load_kswitch_keys, p)[0][0]
    gk_idx = _load(os.path.join(d, "galois_keys.bin"), sio.load_kswitch_keys, p)[0]
    gks = {2 * i + 1: k for i, k in gk_idx.items()}
    ct = lambda name: _load(os.path.join(d, name), sio.load_ciphertext, p)
    x, y = ct("in_x.ct"), ct("in_y.ct")
    py = _load(os.path.join(d, "in_y.pt"), sio.load_plaintext, p)[0]
    hash_ok = sio.level_of(x, p)[1]
    print("parameters: N=%d, %d primes, SEAL %d.%d stream; parms_id hash convention %s" % (
        n, len(p.primes), p.version[0], p.version[1], "confirmed" if hash_ok else "NOT reproduced (level taken from coeff_mod_count)"))
    failures = 0
    for be_cls in backends:
        be = be_cls(p, rlk, gks, rounding)
        prod = be.multiply(x.data, y.data)
        rel = be.relinearize(prod)
        res = be.rescale(rel)
        cases = [
            ("add", be.add(x.data, y.data)), ("sub", be.sub(x.data, y.data)),
            ("multiply_plain", be.multiply_plain(x.data, py)), ("multiply", prod), ("relinearize", rel), ("rescale", res),
            ("rotate_1", be.rotate(x.data, 1)), ("rotate_3", be.rotate(x.data, 3)),
            ("rotate_low_m8", be.rotate(ct("out_rescale.ct").data, -8)), ("mod_switch", be.mod_switch(x.data)),
        ]
        for name, got in cases:
            want = ct("out_%s.ct" % name).data
            ok = got.shape == want.shape and np.array_equal(got, want)
            if not ok:
                failures += 1
                where = "shape %s vs %s" % (got.shape, want.shape) if got.shape != want.shape else \
                    "first difference at [poly, limb, coeff] = %s" % (tuple(int(v) for v in np.argwhere(got != want)[0]),)
            print("%-7s %-16s %s" % (be.name, name, "bit-identical to SEAL's output" if ok else "DIFFERS: " + where))
    return failures


def self_test_files(d, log_n=12):
    """the oracle writes what seal_dump_vectors.cpp would (tool-chain test only)"""
    from oracle import pyoracle as po
    primes = po.coeff_modulus_create(log_n, [60, 40, 40, 60])
    o = po.Oracle(log_n, primes)
    p = sio.Params(sio.SCHEME_CKKS, 1 << log_n, primes)
    sk = o.gen_secret(1)
    pk = o.gen_public(2, sk)
    rlk = o.gen_relin_key(3, sk)
    gks = o.gen_galois_keys(4, sk, steps=[1, -1, 4, -8])
    scale = 2.0 ** 40
    xv = 0.01 * np.arange(64) - 0.3
    yv = 1.0 / (np.arange(64) + 1)
    py = o.encode(yv, scale)
    x, y = o.encrypt(5, pk, o.encode(xv, scale)), o.encrypt(6, pk, py)
    w = lambda name, fn, *a: fn(open(os.path.join(d, name), "wb"), *a)
    w("parms.bin", sio.save_params, p)
    w("relin_keys.bin", sio.save_kswitch_keys, {0: rlk}, p, 1)
    w("galois_keys.bin", sio.save_kswitch_keys, {sio.galois_index(g): k for g, k in gks.items()}, p, 1 << log_n, True)
    w("in_x.ct", sio.save_ciphertext, x, scale, p)
    w("in_y.ct", sio.save_ciphertext, y, scale, p)
    w("in_y.pt", sio.save_plaintext, py, scale, p)
    prod = o.multiply(x, y)
    rel = o.relinearize(prod, rlk)
    res = o.rescale(rel)
    outs = {"add": o.add(x, y), "sub": o.sub(x, y), "multiply_plain": o.multiply_plain(x, py), "multiply": prod,
            "relinearize": rel, "rescale": res, "rotate_1": o.rotate(x, 1, gks), "rotate_3": o.rotate(x, 3, gks),
            "rotate_low_m8": o.rotate(res, -8, gks), "mod_switch": np.ascontiguousarray(x[:, :-1])}
    for name, a in outs.items():
        w("out_%s.ct" % name, sio.save_ciphertext, a, scale, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dir", nargs="?")
    ap.add_argument("--backend", default="both", choices=["cuda", "oracle", "both"])
    ap.add_argument("--self-test", action="store_true")
    ap.add_argument("--rounding", default="auto", choices=["auto", "0", "1", "2", "3"],
                    help="divide-by-last-prime convention (ckks_ctx_set_rounding): 1 = round in key switch and rescale (default), "
                         "0 = floor in both, 2 = round in key switch only, 3 = round in rescale only; auto tries 1, then the others, "
                         "and reports which convention reproduces the files")
    args = ap.parse_args()
    backends = {"cuda": [CudaBackend], "oracle": [OracleBackend], "both": [OracleBackend, CudaBackend]}[args.backend]
    if args.self_test:
        with tempfile.TemporaryDirectory() as d:
            self_test_files(d)
            failures = replay(d, backends)
    else:
        if not args.dir:
            ap.error("DIR (or --self-test) is required")
        modes = [1, 2, 3, 0] if args.rounding == "auto" else [int(args.rounding)]
        for mode in modes:
            print("---- rounding mode %d" % mode)
            failures = replay(args.dir, backends, mode)
            if failures == 0:
                print("rounding mode %d reproduces the files%s" % (mode, "" if mode == 1 else
                      " -- NOT the engine default: set ckks_ctx_set_rounding(ctx, %d) / flip the default" % mode))
                break
    print("%d op(s) differ" % failures)
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
