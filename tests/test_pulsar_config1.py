"""config 1 data path, pinned on the reference's own output: tests/golden/pulsar_plain_lr.json is the stdout of
/root/reference/logistic_regression.cpp compiled unchanged and run on pulsar_stars.csv
(tests/golden/make_pulsar_golden.py).  CPU-only."""
import importlib
import json
import os

import numpy as np

PKG = "seal-fyp-logistic-regression_b200"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load():
    pulsar = importlib.import_module(PKG + ".pulsar")
    with open(os.path.join(GOLD, "pulsar_plain_lr.json")) as fh:
        gold = json.load(fh)
    return pulsar, gold


def test_csv_shape_and_labels():
    pulsar, gold = _load()
    X, y = pulsar.load_csv()
    assert X.shape == (gold["rows"], gold["cols"]) == (2000, 8)
    assert set(np.unique(y)) == {0.0, 1.0} and int(y.sum()) == 164          # 164 positives
    assert abs(float(X[0, 0]) - 140.5625) < 1e-4 and y[0] == 0.0


def test_plain_lr_matches_reference_program_output():
    """scaler + update + cost reproduce what the reference program printed (6 significant digits, float32 vs
    float64 accumulation: tolerance 5e-5 on weights of magnitude ~1, 2e-5 on the cost)"""
    pulsar, gold = _load()
    X, y = pulsar.load_csv()
    Xs = pulsar.standard_scaler(X)
    assert np.abs(Xs.mean(axis=0)).max() < 1e-4 and np.abs(Xs.std(axis=0) - 1).max() < 1e-4
    w0 = np.array(gold["initial_weights"])
    w1 = pulsar.update_weights(Xs, y, w0, gold["learning_rate"])
    assert np.abs(w1 - np.array(gold["weights_after_iteration_0"])).max() < 5e-5
    assert abs(pulsar.cost_function(Xs, y, w1) - gold["cost_after_iteration_0"]) < 2e-5     # 0.665766
    wN, hist = pulsar.train(Xs, y, w0, gold["learning_rate"], gold["iterations"])
    assert np.abs(np.array(hist) - np.array(gold["cost_history"])).max() < 5e-5
    assert np.abs(wN - np.array(gold["final_weights"])).max() < 2e-4
