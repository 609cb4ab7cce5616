import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
PKG = bench.PKG
pkg = importlib.import_module(PKG); eng = pkg.load_engine()
client = importlib.import_module(PKG + ".client"); lr = importlib.import_module(PKG + ".lr"); wl = importlib.import_module(PKG + ".workloads")
ctx = eng.Context(bench.LOG_N, bench._primes()); ev = eng.Evaluator(ctx); enc = client.CKKSEncoder(ctx)
kg = client.KeyGenerator(ctx, seed=1); keys = kg.keyset(steps=[1, -bench.B_MINI]); encr = client.Encryptor(ctx, kg.public_key(), seed=2)
X, y = bench.synthetic_shard(3); slots = ctx.n // 2
lay = lr.ColumnLayout(bench.R_PER_GPU, 8, bench.B_MINI, slots)
cols = encr.encrypt(enc.encode(lay.columns(X), bench.SCALE)); labs = encr.encrypt(enc.encode(lay.labels(y), bench.SCALE))
w0 = np.random.default_rng(5).uniform(-1, 1, 8)
wb = encr.encrypt(enc.encode(np.repeat(w0[:, None], slots, axis=1), bench.SCALE))
import types
T = {}
def timed(name, fn, *a, **k):
    torch.cuda.synchronize(); t = time.time(); r = fn(*a, **k); torch.cuda.synchronize(); T[name] = T.get(name, 0) + time.time() - t; return r
orig_dot = lr.cipher_dot_product; orig_poly = lr.tree_cipher; orig_enc = enc.encode; orig_encrypt = encr.encrypt
lr.cipher_dot_product = lambda *a, **k: timed("cipher_dot_product", orig_dot, *a, **k)
lr.tree_cipher = lambda *a, **k: timed("tree_cipher", orig_poly, *a, **k)
enc.encode = lambda *a, **k: timed("encode(host)", orig_enc, *a, **k)
encr.encrypt = lambda *a, **k: timed("encrypt(host sampling)", orig_encrypt, *a, **k)
for it in range(2):
    T.clear()
    torch.cuda.synchronize(); t = time.time()
    g = lr.column_epoch_gradient(ev, cols, labs, wb, 8, bench.B_MINI, bench.SCALE, keys, enc, encr, degree=7, method="tree")
    torch.cuda.synchronize(); tot = time.time() - t
print("total %.3f s" % tot, {k: round(v, 3) for k, v in T.items()})
