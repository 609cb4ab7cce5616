"""CPU: the product's CoeffModulus helpers agree with the oracle restatement and the golden file."""
import importlib
import json
import os

PKG = "seal-fyp-logistic-regression_b200"
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_params_match_oracle_and_golden(po, pkg):
    params = importlib.import_module(PKG + ".params")
    gold = json.load(open(os.path.join(GOLD, "coeff_modulus.json")))
    for case in gold["create"]:
        got = params.coeff_modulus_create(case["log_n"], case["bits"])
        assert [hex(p) for p in got] == case["primes"]
        assert got == po.coeff_modulus_create(case["log_n"], case["bits"])
    for log_n in (12, 13, 14, 15):
        assert params.bfv_default(log_n) == po.bfv_default(log_n)
        assert params.max_bit_count(log_n) == po.max_bit_count(log_n)
    assert all(params.is_prime(p) == po.is_prime(p) for p in list(range(2, 200)) + [0xffffee001, 0xffffee003, 2**61 - 1])


def test_bench_op_inventory():
    """bench.py's per-epoch op inventory: the key-switch count of the column-layout epoch"""
    import bench
    ops = bench.epoch_op_counts()
    rot = sum(c for op, L, c in ops if op == "rotate")
    assert rot == 4 * 8 * 8192
    relin = sum(c for op, L, c in ops if op == "relinearize")
    assert relin == 4 + 4 * 6 + 32
    assert bench.ks_bytes(3, 16384) == 4718592 and bench.ks_bytes(3, 16384, relin=True) == 5111808


def test_engine_host_tables_match_oracle(po, tmp_path):
    """csrc/tables.cpp (plain C++, no CUDA) against the oracle: minimal primitive 2N-th roots, the twiddle trees,
    N^-1, cross-prime inverses and half-moduli, the NTT-domain Galois permutations, step -> Galois element, NAF"""
    import os
    import subprocess
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "seal-fyp-logistic-regression_b200", "csrc")
    exe = str(tmp_path / "tables_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", csrc, os.path.join(root, "tests", "cpp", "tables_check.cpp"),
                    os.path.join(csrc, "tables.cpp"), "-o", exe], check=True)
    log_n = 12
    primes = po.coeff_modulus_create(log_n, [50, 36, 41, 42, 50])
    out = subprocess.run([exe, str(log_n)] + [str(p) for p in primes], stdout=subprocess.PIPE, text=True, check=True).stdout
    orc = po.Oracle(log_n, primes)
    n, K = 1 << log_n, len(primes)
    seen = {"psi": 0, "ninv": 0, "twsum": 0, "inv": 0, "elt": 0, "perm": 0, "naf": 0}
    for line in out.strip().splitlines():
        f = line.split()
        seen[f[0]] += 1
        if f[0] == "psi":
            assert int(f[2]) == orc.psi(int(f[1]))
        elif f[0] == "ninv":
            j = int(f[1])
            assert int(f[2]) * n % primes[j] == 1
        elif f[0] == "twsum":
            # forward tree node bitrev(e) = psi^e, inverse tree = its inverses: rebuild the checksum from psi
            j = int(f[1])
            p, psi = primes[j], orc.psi(j)
            pw = [1] * n
            for e in range(1, n):
                pw[e] = pw[e - 1] * psi % p
            br = lambda v: int(format(v, "0%db" % log_n)[::-1], 2)
            fw = [0] * n
            for e in range(n):
                fw[br(e)] = pw[e]
            acc = 0
            for e in range(n):
                acc = (acc * 1000003 + fw[e] + 7 * pow(fw[e], -1, p)) % (1 << 64)
            assert acc == int(f[2])
        elif f[0] == "inv":
            a, j = int(f[1]), int(f[2])
            assert int(f[3]) * (primes[a] % primes[j]) % primes[j] == 1
            assert int(f[4]) == (primes[a] >> 1) % primes[j]
        elif f[0] == "elt":
            steps = int(f[1])
            assert int(f[2]) == orc.galois_elt(steps)
        elif f[0] == "perm":
            steps = int(f[1])
            g = orc.galois_elt(steps)
            ident = np.arange(n, dtype=np.uint64)
            perm = orc.galois_permute(g, ident)          # out[i] = in[perm[i]] applied to the identity
            acc = 0
            for v in perm.tolist():
                acc = (acc * 1000003 + int(v)) % (1 << 64)
            assert acc == int(f[2]) and [int(perm[0]), int(perm[1]), int(perm[-1])] == [int(x) for x in f[3:6]]
        elif f[0] == "naf":
            steps = int(f[1])
            assert [int(x) for x in f[2:]] == po.naf(steps)
            assert sum(int(x) for x in f[2:]) == steps
    assert seen == {"psi": K, "ninv": K, "twsum": K, "inv": K * (K - 1), "elt": 7, "perm": 7, "naf": 11}
