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
