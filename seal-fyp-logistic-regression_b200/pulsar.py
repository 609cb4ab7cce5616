"""Host-side data path of config 1 (BASELINE.json configs[0]): the pulsar CSV, the standard scaler and the
plain logistic regression the reference ships beside its encrypted one.

Mirrors the reference's host utilities -- CSVtoMatrix / stringToFloatMatrix (logistic_regression.cpp:236-270,
helper.h:550-686), getMean / getStandardDev / standard_scaler (logistic_regression.cpp:272-338: population
standard deviation, float32), sigmoid / predict / cost_function / update_weights / train (:71-229) -- so that the
encrypted training run (lr.update_weights on the GPU engine) can be driven with the reference's data and checked
against the reference's own plaintext program (tests/golden/pulsar_plain_lr.json, produced by running that
program).  Nothing here is on the GPU hot path; it is the data loader and the semantic reference of config 1.
"""
import os

import numpy as np

DEFAULT_CSV = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pulsar_stars.csv")


def load_csv(path=DEFAULT_CSV, rows=None):
    """-> (features float32 [R][8], labels float32 [R]); first line is the header (skipped, :249-253)"""
    data = np.loadtxt(path, delimiter=",", skiprows=1, dtype=np.float64)
    if rows is not None:
        data = data[:rows]
    return data[:, :-1].astype(np.float32), data[:, -1].astype(np.float32)


def standard_scaler(X):
    """standard_scaler (logistic_regression.cpp:301-338): (x - mean) / population-stddev per column.
    The reference accumulates in float32 in index order; float32 pairwise sums agree to ~1e-6 relative."""
    X = np.asarray(X, dtype=np.float32)
    mean = np.zeros(X.shape[1], dtype=np.float32)
    std = np.zeros(X.shape[1], dtype=np.float32)
    for j in range(X.shape[1]):
        col = X[:, j]
        m = np.float32(0)
        for v in col:                       # getMean: sequential float32 accumulation (:273-283)
            m = np.float32(m + v)
        m = np.float32(m / np.float32(len(col)))
        # getStandardDev (:286-297): pow(float, int) promotes to double, the running sum stays float32
        var = np.float32(0)
        for v in col:
            var = np.float32(var + np.float32((np.float64(np.float32(v - m))) ** 2))
        var = np.float32(var / np.float32(len(col)))
        mean[j], std[j] = m, np.float32(np.sqrt(np.float64(var)))
    return ((X - mean) / std).astype(np.float32)


def sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def cost_function(X, y, w):
    """cost_function (logistic_regression.cpp:98-146): mean cross entropy"""
    p = sigmoid(X.astype(np.float64) @ w.astype(np.float64))
    p = np.where(p == 1.0, p - 1e-4, p)
    return float(np.mean(-y * np.log(p) - (1.0 - y) * np.log(1.0 - p)))


def update_weights(X, y, w, lr):
    """update_weights (logistic_regression.cpp:149-180): w - lr/R * X^T (sigmoid(Xw) - y)"""
    p = sigmoid(X.astype(np.float64) @ w.astype(np.float64))
    return w - lr * (X.T.astype(np.float64) @ (p - y)) / X.shape[0]


def train(X, y, w, lr, iters):
    """train (logistic_regression.cpp:183-229) -> (weights, cost history)"""
    hist = []
    w = np.asarray(w, dtype=np.float64)
    for _ in range(iters):
        w = update_weights(X, y, w, lr)
        hist.append(cost_function(X, y, w))
    return w, hist
