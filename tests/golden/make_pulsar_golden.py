"""Generates the config-1 fixtures (BASELINE.json configs[0]: logistic regression on the pulsar subset).

  tests/golden/pulsar_stars.csv     -- a verbatim copy of the reference's data file (header + 2000 rows x
                                       (8 features + 0/1 label)); data, not code.  The GPU box has no
                                       /root/reference, so the tests and bench.py read this copy.
  tests/golden/pulsar_plain_lr.json -- what the REFERENCE'S OWN plain logistic regression prints on that file:
                                       initial weights (glibc rand(), default seed), the weights and cost after
                                       iteration 0, the final weights and the 100-entry cost history.  Produced by
                                       compiling /root/reference/logistic_regression.cpp unchanged
                                       (`make -C oracle ref` -> oracle/_ref/logistic_regression) and running it.
                                       It is the one golden vector in this repository that comes from reference
                                       code executed here (the CKKS programs need Microsoft SEAL and cannot run).

Run in the build container (needs /root/reference):  python tests/golden/make_pulsar_golden.py
"""
import json
import os
import re
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def main():
    shutil.copyfile(os.path.join(REF, "pulsar_stars.csv"), os.path.join(HERE, "pulsar_stars.csv"))
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref"])
    exe = os.path.join(ROOT, "oracle", "_ref", "logistic_regression")
    with tempfile.TemporaryDirectory() as tmp:
        shutil.copyfile(os.path.join(HERE, "pulsar_stars.csv"), os.path.join(tmp, "pulsar_stars.csv"))
        out = subprocess.run([exe], cwd=tmp, stdout=subprocess.PIPE, text=True, check=True).stdout
    num = r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?"
    w0 = [float(x) for x in re.findall(r"weights\[i\] = (%s)" % num, out)]
    it0 = re.search(r"Iteration:\s+0\s+(%s)\s*\nWeights: ((?:%s, )+)" % (num, num), out)
    cost0 = float(it0.group(1))
    w1 = [float(x) for x in re.findall(num, it0.group(2))]
    new = re.search(r"NEW WEIGHTS\n-+\n((?:%s, )+)" % num, out)
    wN = [float(x) for x in re.findall(num, new.group(1))]
    hist = re.search(r"COST HISTORY\n-+\n(.*?)\n\nACCURACY", out, re.S)
    costs = [float(x) for x in re.findall(num, hist.group(1))]
    rows = int(re.search(r"Number of rows\s+= (\d+)", out).group(1))
    cols = int(re.search(r"Number of cols\s+= (\d+)", out).group(1))
    assert len(w0) == cols == 8 and len(w1) == 8 and len(wN) == 8 and len(costs) == 100 and rows == 2000
    doc = {
        "source": "stdout of /root/reference/logistic_regression.cpp compiled unchanged with g++ -O2 and run on pulsar_stars.csv "
                  "(float32 arithmetic, learning rate 0.1, 100 iterations, logistic_regression.cpp:493)",
        "rows": rows, "cols": cols, "learning_rate": 0.1, "iterations": 100,
        "initial_weights": w0, "cost_after_iteration_0": cost0, "weights_after_iteration_0": w1,
        "final_weights": wN, "cost_history": costs,
        "printed_precision": "6 significant digits (operator<< default)",
    }
    with open(os.path.join(HERE, "pulsar_plain_lr.json"), "w") as fh:
        json.dump(doc, fh, indent=1)
    print("rows %d cols %d cost0 %.6f" % (rows, cols, cost0))


if __name__ == "__main__":
    main()
