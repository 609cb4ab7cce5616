"""The oracle's RNS arithmetic against exact integer arithmetic (CPU only, Python big integers).

SEAL 3.4.x performs rescaling and the mod-down of key switching residue by residue (Evaluator::mod_switch_scale_to_next,
Evaluator::switch_key_inplace).  Over the integers both are ONE statement:

    rescale      : c  ->  floor((c + floor(q_last / 2)) / q_last)                  coefficient-wise, then mod q_j
    key switch   : c2 ->  floor((sum_i d_i * ksk_i + floor(P / 2)) / P)            with d_i = [c2]_{q_i} lifted to [0, q_i)
    rotation     : (c0, c1) -> (c0(x^g), 0) + key switch of c1(x^g)                (Evaluator::apply_galois_inplace)

(floor mode of the rounding switch drops the "+ half").  These tests CRT-compose every coefficient, evaluate that statement
with Python integers and compare the oracle's output word for word -- an implementation-independent pin of the lazy
reductions, the digit lift, the mod-up and the mod-down of oracle/ckks_oracle.c (the polynomial products go through the
oracle's NTT, which test_oracle.py::test_ntt_matches_definition pins against the transform's definition)."""
import numpy as np
import pytest

N_LOG = 12
BITS = [50, 40, 40, 50]


def _crt_basis(primes):
    M = 1
    for p in primes:
        M *= p
    basis = []
    for p in primes:
        Mj = M // p
        basis.append(Mj * pow(Mj % p, -1, p))
    return M, basis


def _compose(rows, primes):
    """rows[j][n] residues -> list of integers in [0, prod primes)"""
    M, basis = _crt_basis(primes)
    n = len(rows[0])
    cols = [[int(v) for v in r] for r in rows]
    return [sum(cols[j][i] * basis[j] for j in range(len(primes))) % M for i in range(n)]


def _limb(values, p):
    return np.array([v % p for v in values], dtype=np.uint64)


@pytest.fixture(scope="module")
def orc(po):
    return po.Oracle(N_LOG, po.coeff_modulus_create(N_LOG, BITS))


@pytest.mark.parametrize("mode", [1, 0])
@pytest.mark.parametrize("L", [3, 2])
def test_rescale_is_integer_division(orc, mode, L):
    rng = np.random.default_rng(100 + L)
    q = orc.primes
    ct = np.stack([np.stack([rng.integers(0, q[j], orc.n, dtype=np.uint64) for j in range(L)]) for _ in range(2)])
    orc.set_rounding(mode)
    try:
        got = orc.rescale(ct)
    finally:
        orc.set_rounding(1)
    half = (q[L - 1] >> 1) if mode else 0
    for s in range(2):
        X = _compose([orc.intt(j, ct[s, j]) for j in range(L)], q[:L])
        Y = [(x + half) // q[L - 1] for x in X]
        for j in range(L - 1):
            assert np.array_equal(got[s, j], orc.ntt(j, _limb(Y, q[j]))), (mode, L, s, j)


@pytest.mark.parametrize("mode", [1, 0])
@pytest.mark.parametrize("L", [3, 2])
def test_key_switch_is_integer_inner_product_and_division(orc, mode, L):
    rng = np.random.default_rng(200 + L)
    q, K, n = orc.primes, orc.K, orc.n
    P = q[K - 1]
    sk = orc.gen_secret(7)
    rlk = orc.gen_relin_key(8, sk)                                   # [K-1][2][K][N]
    ct3 = np.stack([np.stack([rng.integers(0, q[j], n, dtype=np.uint64) for j in range(L)]) for _ in range(3)])
    orc.set_rounding(mode)
    try:
        got = orc.relinearize(ct3, rlk)
    finally:
        orc.set_rounding(1)
    kp = list(range(L)) + [K - 1]                                    # primes of the extended basis at this level
    digits = [[int(v) for v in orc.intt(i, ct3[2, i])] for i in range(L)]          # d_i in [0, q_i)
    acc = [[None] * len(kp) for _ in range(2)]
    for a, j in enumerate(kp):
        t = [[int(v) for v in orc.ntt(j, _limb(digits[i], q[j]))] for i in range(L)]   # NTT_{p_j}(d_i mod p_j)
        for k in range(2):
            key = [[int(v) for v in rlk[i, k, j]] for i in range(L)]
            acc[k][a] = np.array([sum(t[i][m] * key[i][m] for i in range(L)) % q[j] for m in range(n)], dtype=np.uint64)
    half = (P >> 1) if mode else 0
    for k in range(2):
        A = _compose([orc.intt(j, acc[k][a]) for a, j in enumerate(kp)], [q[j] for j in kp])
        Y = [(x + half) // P for x in A]
        for j in range(L):
            want = (orc.ntt(j, _limb(Y, q[j])).astype(object) + ct3[k, j].astype(object)) % q[j]
            assert np.array_equal(got[k, j], want.astype(np.uint64)), (mode, L, k, j)


def _automorphism(coeffs, g, n, p):
    """a(x) -> a(x^g) mod (x^n + 1, p) on a coefficient vector, straight from the definition"""
    out = [0] * n
    for i, a in enumerate(coeffs):
        e = (i * g) % (2 * n)
        if e < n:
            out[e] = int(a) % p
        else:
            out[e - n] = (-int(a)) % p
    return out


@pytest.mark.parametrize("step", [1, -3])
def test_rotation_is_automorphism_plus_integer_key_switch(orc, step):
    """the headline op (rotate_vector with a direct key): the Galois automorphism applied coefficient-wise by its
    definition, then the integer statement of the key switch above"""
    rng = np.random.default_rng(300 + step)
    q, K, n = orc.primes, orc.K, orc.n
    L, P = K - 1, q[K - 1]
    g = orc.galois_elt(step)
    sk = orc.gen_secret(17)
    gk = orc.gen_galois_key(18, sk, g)
    ct = np.stack([np.stack([rng.integers(0, q[j], n, dtype=np.uint64) for j in range(L)]) for _ in range(2)])
    got = orc.apply_galois(ct, g, gk)
    # sigma_g on both polynomials, coefficient domain, residue by residue (the map is linear, so it commutes with the CRT)
    sig = [[_automorphism(orc.intt(j, ct[s, j]), g, n, q[j]) for j in range(L)] for s in range(2)]
    kp = list(range(L)) + [K - 1]
    acc = [[None] * len(kp) for _ in range(2)]
    for a, j in enumerate(kp):
        t = [[int(v) for v in orc.ntt(j, _limb(sig[1][i], q[j]))] for i in range(L)]     # digits of sigma_g(c1), lifted to [0, q_i)
        for k in range(2):
            key = [[int(v) for v in gk[i, k, j]] for i in range(L)]
            acc[k][a] = np.array([sum(t[i][m] * key[i][m] for i in range(L)) % q[j] for m in range(n)], dtype=np.uint64)
    for k in range(2):
        A = _compose([orc.intt(j, acc[k][a]) for a, j in enumerate(kp)], [q[j] for j in kp])
        Y = [(x + (P >> 1)) // P for x in A]
        for j in range(L):
            want = orc.ntt(j, _limb(Y, q[j])).astype(object)
            if k == 0:
                want = want + orc.ntt(j, np.array(sig[0][j], dtype=np.uint64)).astype(object)
            assert np.array_equal(got[k, j], (want % q[j]).astype(np.uint64)), (step, k, j)
