"""GPU parity tests: every evaluator op of the CUDA path, through the C ABI, bit-exact against
the CPU oracle on identical seeded inputs and keys.  Uniform random residues are valid inputs
for every op (the evaluator is a deterministic function of limbs and keys), so bit-exactness is
checked on those; decrypt/decode semantics are checked separately within a stated tolerance."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CHAINS = {
    12: [50, 40, 40, 50],
    13: [60, 40, 40, 60],
    14: [60, 40, 40, 40, 40, 60],
    15: [60, 40, 40, 40, 60],
}


@pytest.mark.parametrize("log_n", [12, 13, 14, 15])
def test_ntt_bit_exact_all_primes(make_fixture, eng, log_n):
    import torch
    fx = make_fixture(log_n, CHAINS[log_n])
    rng = np.random.default_rng(log_n)
    K = len(fx.primes)
    a = np.stack([rng.integers(0, p, size=(3, fx.n), dtype=np.uint64) for p in fx.primes], axis=1)  # [3][K][N]
    a[0, :, :] = 0
    a[0, :, 1] = 1                      # x -> twiddle pattern
    a[1] = np.array(fx.primes, dtype=np.uint64)[:, None] - 1   # maximum residues
    want = np.stack([np.stack([fx.orc.ntt(j, a[i, j]) for j in range(K)]) for i in range(3)])
    t = torch.from_numpy(a.view(np.int64).copy()).cuda()
    fx.ev.ntt_forward(t)
    got = t.cpu().numpy().view(np.uint64)
    assert np.array_equal(got, want)
    fx.ev.ntt_inverse(t)
    assert np.array_equal(t.cpu().numpy().view(np.uint64), a)
    # inverse alone vs oracle
    t2 = torch.from_numpy(a.view(np.int64).copy()).cuda()
    fx.ev.ntt_inverse(t2)
    want_i = np.stack([np.stack([fx.orc.intt(j, a[i, j]) for j in range(K)]) for i in range(3)])
    assert np.array_equal(t2.cpu().numpy().view(np.uint64), want_i)


def test_ntt_limb_per_cta_variant_bit_exact():
    """the north-star NTT variant (one RNS limb per CTA, limb resident in shared memory, CKKS_NTT_LIMB=1; N <= 16384) gives
    the same bits as the oracle: the NTT parity tests above re-run in a child process with the variant selected"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CKKS_NTT_LIMB="1")
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "-x",
                          "-k", "test_ntt_bit_exact_all_primes or test_ntt_bfv_default_primes"], env=env, cwd=root,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0 and "passed" in res.stdout, res.stdout[-2000:]


def test_cluster_inner_product_variant_bit_exact():
    """the thread-block-cluster inner product (k_ks_mac_cl: one CTA per digit, partial products summed through distributed
    shared memory; default only for single ciphertexts at N <= 16384) forced for every batch size and degree
    (CKKS_CLUSTER_BELOW=1000): the key-switch parity tests re-run in a child process must still match the oracle bit for bit"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CKKS_CLUSTER_BELOW="1000")
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "-x",
                          "-k", "test_relinearize_bit_exact or test_apply_galois_bit_exact or test_rotate_vector_naf_chain or "
                                "test_rotate_sum_chain_matches_sequential or test_extreme_modulus_chains or test_random_mixed_chains"],
                         env=env, cwd=root, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0 and "passed" in res.stdout, res.stdout[-2000:]


def test_ntt_bfv_default_primes(po, eng):
    """config 2 chains: SEAL's BFVDefault primes for N = 4096 / 8192 (benchmark.cpp:137)"""
    import torch
    for log_n in (12, 13):
        primes = po.bfv_default(log_n)
        orc = po.Oracle(log_n, primes)
        ctx = eng.Context(log_n, primes)
        ev = eng.Evaluator(ctx)
        rng = np.random.default_rng(7)
        a = np.stack([rng.integers(0, p, size=(1 << log_n), dtype=np.uint64) for p in primes])[None]
        t = torch.from_numpy(a.view(np.int64).copy()).cuda()
        ev.ntt_forward(t)
        want = np.stack([orc.ntt(j, a[0, j]) for j in range(len(primes))])[None]
        assert np.array_equal(t.cpu().numpy().view(np.uint64), want)


@pytest.mark.parametrize("log_n", [12, 13])
def test_elementwise_bit_exact(make_fixture, log_n):
    fx = make_fixture(log_n, CHAINS[log_n])
    rng = np.random.default_rng(11)
    o, ev, ctx = fx.orc, fx.ev, fx.ctx
    for L in (fx.L, 1):
        a = fx.random_ct(rng, 3, 2, L)
        b = fx.random_ct(rng, 3, 2, L)
        a[0, 0, 0, :4] = 0                               # zero / negate edge
        b[0, 0, 0, :4] = np.uint64(fx.primes[0] - 1)
        da, db = ctx.upload(a), ctx.upload(b)
        assert np.array_equal(ev.add(da, db).numpy(), np.stack([o.add(a[i], b[i]) for i in range(3)]))
        assert np.array_equal(ev.sub(da, db).numpy(), np.stack([o.sub(a[i], b[i]) for i in range(3)]))
        assert np.array_equal(ev.negate_inplace(da.clone()).numpy(), np.stack([o.negate(a[i]) for i in range(3)]))
        assert np.array_equal(ev.multiply(da, db).numpy(), np.stack([o.multiply(a[i], b[i]) for i in range(3)]))
        # size 3 x size 2 and 3 x 3 (Linear_Transform_Cipher sums size-3 products, helper.h:222-233)
        a3 = fx.random_ct(rng, 3, 3, L)
        d3 = ctx.upload(a3)
        assert np.array_equal(ev.multiply(d3, db).numpy(), np.stack([o.multiply(a3[i], b[i]) for i in range(3)]))
        assert np.array_equal(ev.add(d3, d3).numpy(), np.stack([o.add(a3[i], a3[i]) for i in range(3)]))
        # plaintext ops: broadcast and per-entry plaintexts
        p1 = fx.random_ct(rng, 1, 1, L)
        p3 = fx.random_ct(rng, 3, 1, L)
        dp1, dp3 = ctx.upload(p1), ctx.upload(p3)
        assert np.array_equal(ev.multiply_plain(da, dp1).numpy(), np.stack([o.multiply_plain(a[i], p1[0, 0]) for i in range(3)]))
        assert np.array_equal(ev.multiply_plain(da, dp3).numpy(), np.stack([o.multiply_plain(a[i], p3[i, 0]) for i in range(3)]))
        assert np.array_equal(ev.add_plain(da, dp1).numpy(), np.stack([o.add_plain(a[i], p1[0, 0]) for i in range(3)]))
        # add_many = sequential adds
        want = a[0]
        for i in range(1, 3):
            want = o.add(want, a[i])
        assert np.array_equal(ev.add_many(da).numpy()[0], want)


@pytest.mark.parametrize("log_n", [12, 13, 14, 15])
def test_relinearize_bit_exact(make_fixture, log_n):
    fx = make_fixture(log_n, CHAINS[log_n])
    rng = np.random.default_rng(21)
    levels = (fx.L, 1) if log_n > 12 else tuple(range(fx.L, 0, -1))
    for L in levels:
        a = fx.random_ct(rng, 2, 3, L)
        got = fx.ev.relinearize(fx.ctx.upload(a), fx.keys).numpy()
        want = np.stack([fx.orc.relinearize(a[i], fx.rlk) for i in range(2)])
        assert np.array_equal(got, want), (log_n, L)


@pytest.mark.parametrize("log_n", [12, 13, 14, 15])
def test_apply_galois_bit_exact(make_fixture, log_n):
    fx = make_fixture(log_n, CHAINS[log_n])
    rng = np.random.default_rng(22)
    levels = (fx.L, 1) if log_n > 12 else tuple(range(fx.L, 0, -1))
    for L in levels:
        a = fx.random_ct(rng, 2, 2, L)
        d = fx.ctx.upload(a)
        for step in ((1, -8) if log_n > 12 else (1, -1, 2, 4, -8)):
            g = fx.orc.galois_elt(step)
            got = fx.ev.apply_galois(d, g, fx.keys).numpy()
            want = np.stack([fx.orc.apply_galois(a[i], g, fx.gks[g]) for i in range(2)])
            assert np.array_equal(got, want), (log_n, L, step)


def test_rotate_vector_naf_chain(make_fixture):
    """composite steps go through SEAL's NAF decomposition, least-significant term first"""
    fx = make_fixture(12, CHAINS[12], steps=(1, -1, 2, -2, 4, -4, 8, -8, 16))
    rng = np.random.default_rng(23)
    a = fx.random_ct(rng, 2, 2, fx.L)
    d = fx.ctx.upload(a)
    for step in (3, 7, -3, 5, 11, 13, -6, 0):
        got = fx.ev.rotate_vector(d, step, fx.keys).numpy()
        want = np.stack([fx.orc.rotate(a[i], step, fx.gks) for i in range(2)])
        assert np.array_equal(got, want), step
    from importlib import import_module
    capi = import_module("seal-fyp-logistic-regression_b200.capi")
    with pytest.raises(capi.CkksInvalidArgument):
        fx.ev.rotate_vector(d, 32, fx.keys)          # power of two without a key: "Galois key not present"
    with pytest.raises(capi.CkksInvalidArgument):
        fx.ev.rotate_vector(d, fx.n // 2, fx.keys)   # "step count too large"


def test_rotation_plan_shares_naf_prefixes_bit_exact(make_fixture, eng):
    """a rotation plan applied to ONE ciphertext runs the common leading NAF terms of its rotations once
    (ckks_rotplan_keyswitches_shared): every output must still equal the oracle's own rotate_vector of that step, bit for
    bit -- including a step that is a proper prefix of another (5 = [1,4] and 21 = [1,4,16]), repeated steps, step 0, a
    level below the top, and a plan whose pure intermediates outnumber its entries (kept in the plain form)"""
    fx = make_fixture(12, CHAINS[12], steps=(1, -1, 2, -2, 4, -4, 8, -8, 16, -16, 32, 64))
    rng = np.random.default_rng(29)
    steps = [5, 21, 85, 3, 0, 5, -3, 13, 19, 11, 27, 7, 1, 64, 43, 21]
    plan = eng.RotPlan(fx.ctx, fx.keys, steps)
    assert plan.keyswitches_shared < plan.keyswitches
    for L in (fx.L, fx.L - 1):
        a = fx.random_ct(rng, 1, 2, L)
        d = fx.ctx.upload(a)
        got = fx.ev.rotate_plan(d, plan).numpy()
        for i, st in enumerate(steps):
            assert np.array_equal(got[i], fx.orc.rotate(a[0], st, fx.gks)), (L, st)
    # distinct inputs per entry: nothing to share, same plan object
    a = fx.random_ct(rng, len(steps), 2, fx.L)
    got = fx.ev.rotate_plan(fx.ctx.upload(a), plan).numpy()
    for i, st in enumerate(steps):
        assert np.array_equal(got[i], fx.orc.rotate(a[i], st, fx.gks)), st
    # two entries, four pure intermediates (85 = [1, 4, 16, 64], 81 = [1, 16, 64] share only [1]): more than the scratch holds
    deep = eng.RotPlan(fx.ctx, fx.keys, [85, 81])
    assert deep.keyswitches_shared == deep.keyswitches == 7
    a = fx.random_ct(rng, 1, 2, fx.L)
    got = fx.ev.rotate_plan(fx.ctx.upload(a), deep).numpy()
    assert np.array_equal(got[0], fx.orc.rotate(a[0], 85, fx.gks)) and np.array_equal(got[1], fx.orc.rotate(a[0], 81, fx.gks))
    # 85 and 21 = [1, 4, 16]: one chain, the shorter rotation is an inner node of the longer one
    chain = eng.RotPlan(fx.ctx, fx.keys, [85, 21])
    assert chain.keyswitches_shared == 4 and chain.keyswitches == 7
    got = fx.ev.rotate_plan(fx.ctx.upload(a), chain).numpy()
    assert np.array_equal(got[0], fx.orc.rotate(a[0], 85, fx.gks)) and np.array_equal(got[1], fx.orc.rotate(a[0], 21, fx.gks))


@pytest.mark.parametrize("log_n", [12, 13, 14, 15])
def test_rescale_bit_exact(make_fixture, log_n):
    fx = make_fixture(log_n, CHAINS[log_n])
    rng = np.random.default_rng(24)
    for L in range(fx.L, 1, -1):
        for S in (2, 3):
            a = fx.random_ct(rng, 2, S, L)
            got = fx.ev.rescale_to_next(fx.ctx.upload(a))
            want = np.stack([fx.orc.rescale(a[i]) for i in range(2)])
            assert got.limbs == L - 1
            assert np.array_equal(got.numpy(), want), (log_n, L, S)
    # in place, with limb capacity kept
    a = fx.random_ct(rng, 1, 2, fx.L)
    d = fx.ctx.upload(a)
    fx.ev.rescale_to_next_inplace(d)
    assert np.array_equal(d.numpy()[0], fx.orc.rescale(a[0]))


def test_rounding_switch_matches_oracle(make_fixture):
    """the two unconfirmed SEAL conventions (round vs floor in the key switch's mod-down and in rescale) can be flipped
    independently, on the engine and on the oracle alike; every combination stays bit-exact, and each switch changes
    exactly the op it names"""
    fx = make_fixture(12, CHAINS[12])
    rng = np.random.default_rng(25)
    a = fx.random_ct(rng, 1, 2, fx.L)
    a3 = fx.random_ct(rng, 1, 3, fx.L)
    seen = {}
    try:
        for mode in (0, 1, 2, 3):
            fx.ctx.set_rounding(mode)
            fx.orc.set_rounding(mode)
            rs = fx.ev.rescale_to_next(fx.ctx.upload(a)).numpy()[0]
            rl = fx.ev.relinearize(fx.ctx.upload(a3), fx.keys).numpy()[0]
            assert np.array_equal(rs, fx.orc.rescale(a[0])), mode
            assert np.array_equal(rl, fx.orc.relinearize(a3[0], fx.rlk)), mode
            seen[mode] = (rs, rl)
    finally:
        fx.ctx.set_rounding(1)
        fx.orc.set_rounding(1)
    assert np.array_equal(seen[1][0], seen[3][0]) and np.array_equal(seen[0][0], seen[2][0])     # rescale: modes 1, 3 round
    assert np.array_equal(seen[1][1], seen[2][1]) and np.array_equal(seen[0][1], seen[3][1])     # key switch: modes 1, 2 round
    assert not np.array_equal(seen[0][0], seen[1][0]) and not np.array_equal(seen[0][1], seen[1][1])


def test_batched_equals_single_and_chunked_workspace(make_fixture):
    """a batch is processed exactly like its entries one by one, also when the workspace cap
    forces the key switch to run in chunks"""
    fx = make_fixture(12, CHAINS[12])
    rng = np.random.default_rng(26)
    a = fx.random_ct(rng, 7, 2, fx.L)
    d = fx.ctx.upload(a)
    g = fx.orc.galois_elt(1)
    full = fx.ev.apply_galois(d, g, fx.keys).numpy()
    fx.ctx.set_workspace_cap(3 * 8 * fx.n * 40)      # room for ~2 ciphertexts per chunk
    try:
        chunked = fx.ev.apply_galois(d, g, fx.keys).numpy()
    finally:
        fx.ctx.set_workspace_cap(1 << 30)
    assert np.array_equal(full, chunked)
    for i in range(7):
        assert np.array_equal(fx.ev.apply_galois(d[i], g, fx.keys).numpy()[0], full[i])


def test_end_to_end_semantics_and_errors(make_fixture, po):
    """encrypt -> GPU evaluate -> decrypt/decode vs plaintext math, |err| < 2^-20 relative at
    scale 2^40 (north star tolerance); plus SEAL's error behaviour"""
    from importlib import import_module
    capi = import_module("seal-fyp-logistic-regression_b200.capi")
    fx = make_fixture(13, CHAINS[13])
    o, ev, ctx = fx.orc, fx.ev, fx.ctx
    rng = np.random.default_rng(27)
    scale = 2.0 ** 40
    x, y = rng.uniform(-1, 1, 128), rng.uniform(-1, 1, 128)
    cx = o.encrypt(30, fx.pk, o.encode(x, scale))
    cy = o.encrypt(31, fx.pk, o.encode(y, scale))
    dx, dy = ctx.upload(cx, scale=scale), ctx.upload(cy, scale=scale)
    prod = ev.rescale_to_next(ev.relinearize(ev.multiply(dx, dy), fx.keys))
    dec = o.decode(o.decrypt(fx.sk, prod.numpy()[0]), prod.scale)[:128]
    tol = 2.0 ** -20
    assert np.abs(dec - x * y).max() < tol * max(1.0, np.abs(x * y).max())
    rot = ev.rotate_vector(dx, 1, fx.keys)
    full = np.zeros(fx.n // 2)
    full[:128] = x
    dec = o.decode(o.decrypt(fx.sk, rot.numpy()[0]), scale)
    assert np.abs(dec - np.roll(full, -1)).max() < tol
    # scale mismatch / level mismatch / scale out of bounds / end of chain / transparent
    with pytest.raises(capi.CkksInvalidArgument, match="scale mismatch"):
        ev.add(dx, ev.multiply_plain(dy, ctx.upload_plain(o.encode(1.0, scale), scale=scale)))
    low = ev.mod_switch_to(dy, dy.limbs - 1)
    with pytest.raises(capi.CkksInvalidArgument, match="parameter mismatch"):
        ev.add(dx, low)
    big = ev.multiply(dx, dy)
    big2 = ev.relinearize(big, fx.keys)
    with pytest.raises(capi.CkksInvalidArgument, match="scale out of bounds"):
        ev.multiply(ev.multiply(big2, big2), big2)
    one = ev.mod_switch_to(dx, 1)
    with pytest.raises(capi.CkksInvalidArgument, match="end of modulus switching chain"):
        ev.rescale_to_next(one)
    zero = ctx.upload_plain(np.zeros((fx.L, fx.n), dtype=np.uint64), scale=scale)
    with pytest.raises(capi.CkksLogicError, match="transparent"):
        ev.multiply_plain(dx, zero, check_transparent=True)
    with pytest.raises(capi.CkksInvalidArgument):
        ev.rotate_vector(dx, 3 * 1024 + 5, eng_keys_without(fx))


def eng_keys_without(fx):
    return fx.eng.KeySet(fx.ctx)


def test_full_size_properties_n32768(make_fixture):
    """BASELINE config 5 size (N = 32768): besides the direct oracle comparison above, a
    size-independent property -- for ANY input polynomials, Dec(apply_galois(ct)) equals the
    automorphism of Dec(ct) up to key-switch noise (a few thousand units against 2^60 moduli)."""
    fx = make_fixture(15, CHAINS[15])
    rng = np.random.default_rng(28)
    L = fx.L
    a = fx.random_ct(rng, 2, 2, L)
    da = fx.ctx.upload(a)
    for step in (1, -8):
        g = fx.orc.galois_elt(step)
        ra = fx.ev.apply_galois(da, g, fx.keys).numpy()
        for i in range(2):
            m = fx.orc.decrypt(fx.sk, a[i])
            mr = fx.orc.decrypt(fx.sk, ra[i])
            for j in range(L):
                p = fx.primes[j]
                want = fx.orc.galois_permute(g, m[j])
                diff = (mr[j].astype(object) - want.astype(object)) % p
                c = fx.orc.intt(j, np.array(diff, dtype=np.uint64)).astype(object)
                c = np.where(c > p // 2, c - p, c)
                assert np.abs(c).max() < 2 ** 24, (step, i, j)


@pytest.mark.parametrize("batch,count", [(1, 1), (3, 4), (5, 9), (16, 6), (19, 7), (33, 12), (2, 19), (17, 35)])
def test_rotate_sum_chain_matches_sequential(make_fixture, batch, count):
    """ckks_rotate_sum_chain (fused add, CUDA-graph replay -- graphs of 8 ping-pong pairs for long chains, of 1 pair and
    single eager steps for the remainder --, two concurrent lanes for batches >= 16)
    equals `count` sequential rotate_vector + add_inplace steps of the oracle (helper.h:472-476)"""
    fx = make_fixture(12, CHAINS[12])
    rng = np.random.default_rng(100 + batch)
    L = 2
    dup0 = fx.random_ct(rng, batch, 2, L)
    acc0 = fx.random_ct(rng, batch, 2, L)
    g = fx.orc.galois_elt(1)
    want_acc, want_dup = [], []
    for b in range(batch):
        d, a = dup0[b], acc0[b]
        for _ in range(count):
            d = fx.orc.apply_galois(d, g, fx.gks[g])
            a = fx.orc.add(a, d)
        want_acc.append(a)
        want_dup.append(d)
    dup, acc = fx.ctx.upload(dup0), fx.ctx.upload(acc0)
    for _ in range(2):                      # second round replays the cached graphs on fresh data
        dup.data.copy_(fx.ctx.upload(dup0).data)
        acc.data.copy_(fx.ctx.upload(acc0).data)
        last = fx.ev.rotate_sum_chain(dup, acc, 1, count, fx.keys)
        assert np.array_equal(acc.numpy(), np.stack(want_acc))
        assert np.array_equal(last.numpy(), np.stack(want_dup))


def test_large_batch_two_lanes_matches_oracle(make_fixture):
    """batches of 16+ run as two concurrent half-batch pipelines; every entry still equals the oracle"""
    fx = make_fixture(12, CHAINS[12])
    rng = np.random.default_rng(77)
    a = fx.random_ct(rng, 37, 2, fx.L)
    a3 = fx.random_ct(rng, 21, 3, fx.L)
    g = fx.orc.galois_elt(-8)
    got = fx.ev.apply_galois(fx.ctx.upload(a), g, fx.keys).numpy()
    for i in range(37):
        assert np.array_equal(got[i], fx.orc.apply_galois(a[i], g, fx.gks[g])), i
    got = fx.ev.relinearize(fx.ctx.upload(a3), fx.keys).numpy()
    for i in range(21):
        assert np.array_equal(got[i], fx.orc.relinearize(a3[i], fx.rlk)), i


@pytest.mark.parametrize("log_n,bits", [
    (12, [36, 36]),                          # shortest chain: one data prime + the special prime (L = 1)
    (15, [60] + [40] * 19 + [60]),           # deepest chain SEAL allows at N = 32768 with 40-bit primes (880 of 881 bits)
    (15, [27] * 32),                         # most primes the engine accepts (K = 32), all on the FP64 path
    (13, [60, 60, 60]),                      # only large primes: everything on the integer path
    (13, [50, 30, 30, 30, 30, 40]),          # large first prime, SMALL special prime: its limb comes from the FP64 kernel
    (13, [40, 60, 40]),                      # a large prime between small ones, small special prime
])
def test_extreme_modulus_chains(make_fixture, log_n, bits):
    """maximum / minimum sizes: relinearize, one Galois step and rescale stay bit-exact at the top level
    and at level 1 of the shortest, the deepest and the widest chains"""
    fx = make_fixture(log_n, bits, steps=(1,))
    rng = np.random.default_rng(len(bits))
    g = fx.orc.galois_elt(1)
    for L in sorted({fx.L, 1}, reverse=True):
        a3 = fx.random_ct(rng, 1, 3, L)
        assert np.array_equal(fx.ev.relinearize(fx.ctx.upload(a3, cap=fx.L), fx.keys).numpy()[0], fx.orc.relinearize(a3[0], fx.rlk)), L
        a2 = fx.random_ct(rng, 1, 2, L)
        d = fx.ctx.upload(a2, cap=fx.L)
        assert np.array_equal(fx.ev.apply_galois(d, g, fx.keys).numpy()[0], fx.orc.apply_galois(a2[0], g, fx.gks[g])), L
        if L > 1:
            assert np.array_equal(fx.ev.rescale_to_next(d).numpy()[0], fx.orc.rescale(a2[0])), L


def _random_chains(count, seed=2024):
    """prime-size patterns mixing both arithmetic paths (primes below 2^41 run on the FP64 pipe, larger
    ones on the integer pipe; 41 and 42 bits straddle the boundary), total <= 218 bits (N = 8192)"""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        K = int(rng.integers(2, 7))
        bits = [int(b) for b in rng.choice([28, 33, 36, 40, 41, 42, 47, 50, 60], size=K)]
        if sum(bits) <= 218 and bits not in out:
            out.append(bits)
    return out


@pytest.mark.parametrize("bits", _random_chains(10), ids=lambda b: "-".join(map(str, b)))
def test_random_mixed_chains(make_fixture, bits):
    """every placement of small / large primes (first, middle, special) keeps the key switch, the rotate-and-sum
    chain and rescale bit-exact; batch 17 exercises the two-lane split"""
    fx = make_fixture(13, bits, steps=(1,))
    rng = np.random.default_rng(sum(bits))
    g = fx.orc.galois_elt(1)
    L = fx.L
    a3 = fx.random_ct(rng, 1, 3, L)
    assert np.array_equal(fx.ev.relinearize(fx.ctx.upload(a3), fx.keys).numpy()[0], fx.orc.relinearize(a3[0], fx.rlk))
    a2 = fx.random_ct(rng, 17, 2, L)
    d = fx.ctx.upload(a2)
    got = fx.ev.apply_galois(d, g, fx.keys).numpy()
    for b in (0, 8, 16):
        assert np.array_equal(got[b], fx.orc.apply_galois(a2[b], g, fx.gks[g])), b
    if L > 1:
        assert np.array_equal(fx.ev.rescale_to_next(d).numpy()[3], fx.orc.rescale(a2[3]))
    # two steps of the fused rotate-and-sum chain (graph replay) on the first entries
    dup, acc = fx.ctx.upload(a2[:2]), fx.ctx.upload(a2[:2])
    fx.ev.rotate_sum_chain(dup, acc, 1, 2, fx.keys)
    want = a2[0]
    r = a2[0]
    for _ in range(2):
        r = fx.orc.apply_galois(r, g, fx.gks[g])
        want = fx.orc.add(want, r)
    assert np.array_equal(acc.numpy()[0], want)


def test_rotation_plan_survives_key_replacement(make_fixture):
    """a rotation plan compiled before one of its Galois keys is replaced (KeySet.set_galois with a new buffer for the same
    element) must use the NEW key at its next run -- the engine-owned copy of the old key is freed on replacement, and
    plans used to keep its address"""
    import importlib
    wl = importlib.import_module("seal-fyp-logistic-regression_b200.workloads")
    fx = make_fixture(12, CHAINS[12], steps=(1, -1, 2, -2, 4, 8))
    eng, ctx, orc = fx.eng, fx.ctx, fx.orc
    keys = eng.KeySet(ctx)                               # a private key set: the fixture's is shared by other tests
    gks = dict(fx.gks)
    for g, k in gks.items():
        keys.set_galois(g, ctx.upload_key(k))
    rng = np.random.default_rng(91)
    a = fx.random_ct(rng, 1, 2, fx.L)
    d = ctx.upload(a)
    plan = wl.PlanCache(ctx, keys).get([1, 3, 6])
    before = fx.ev.rotate_plan(d, plan).numpy()
    for b, st in enumerate((1, 3, 6)):
        assert np.array_equal(before[b], orc.rotate(a[0], st, gks))
    # a different key for step 4 (another sample of the same key switch): replace it, run the SAME plan.
    # NAF: 3 = {-1, 4} uses it, 1 and 6 = {-2, 8} do not
    g4 = orc.galois_elt(4)
    gks[g4] = orc.gen_galois_key(4242, fx.sk, g4)
    assert not np.array_equal(gks[g4], fx.gks[g4])
    keys.set_galois(g4, ctx.upload_key(gks[g4]))
    after = fx.ev.rotate_plan(d, plan).numpy()
    for b, st in enumerate((1, 3, 6)):
        assert np.array_equal(after[b], orc.rotate(a[0], st, gks)), st
    assert np.array_equal(after[0], before[0]) and np.array_equal(after[2], before[2]) and not np.array_equal(after[1], before[1])
