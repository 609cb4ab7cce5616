"""Generates tests/golden/*.json from the CPU oracle.

The reference holds no golden vectors for this path and SEAL is not installed here (SURVEY.md
8(c)), so these fixtures pin (a) CoeffModulus::Create outputs derived from SEAL's published prime
search rule and (b) SHA-256 digests of oracle outputs on seeded inputs, as a regression guard.
Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CREATE_CASES = [
    (13, [60, 40, 40, 60]),                    # README.md:176
    (14, [60, 40, 40, 60]),
    (14, [60, 40, 40, 40, 40, 60]),            # matrix_multiplication.cpp:147
    (14, [60, 40, 40, 40, 40, 40, 40, 40, 60]),  # logistic_regression_ckks.cpp:421
    (15, [60, 40, 40, 40, 40, 40, 40, 40, 40, 60]),  # repaired LR chain
    (13, [50, 30, 30, 50, 50]),                # 3_levels.cpp:16
]


def _h(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def compute_digests(po):
    log_n = 12
    primes = po.coeff_modulus_create(log_n, [50, 40, 40, 50])
    o = po.Oracle(log_n, primes)
    sk = o.gen_secret(1)
    pk = o.gen_public(2, sk)
    rlk = o.gen_relin_key(3, sk)
    g = o.galois_elt(3)
    gk = o.gen_galois_key(4, sk, g)
    x = np.arange(64) / 64.0
    ct = o.encrypt(5, pk, o.encode(x, 2.0 ** 40))
    m = o.multiply(ct, ct)
    r = o.relinearize(m, rlk)
    return {
        "psi": [hex(o.psi(j)) for j in range(o.K)],
        "encode": _h(o.encode(x, 2.0 ** 40)),
        "encrypt": _h(ct),
        "multiply": _h(m),
        "relinearize": _h(r),
        "rescale": _h(o.rescale(r)),
        "apply_galois": _h(o.apply_galois(ct, g, gk)),
    }


def main():
    from oracle import pyoracle as po
    create = [{"log_n": ln, "bits": bits, "primes": [hex(p) for p in po.coeff_modulus_create(ln, bits)]}
              for ln, bits in CREATE_CASES]
    json.dump({"create": create}, open(os.path.join(HERE, "coeff_modulus.json"), "w"), indent=1)
    json.dump(compute_digests(po), open(os.path.join(HERE, "oracle_digests.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
