"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel shares.
usage: python profiles/launch_shares.py gpurun_out/bench_launches.csv profiles/r01_bench_launch_shares.md"""
import collections
import csv
import re
import sys


def main(src, dst, tag="r02"):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if len(r) > 5]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[iu], 1.0)
        name = re.sub(r"\(int\)|\(bool\)", "", r[ik].split("(")[0] if "<" not in r[ik] else r[ik][:r[ik].index(">") + 1])
        name = name.replace("void ", "")[:70]
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    with open(dst, "w") as f:
        f.write("# %s: ncu launch list of `bench.py --steps 1 --warmup 1 --ncu` (first %d launches of the timed epoch)\n\n" % (tag, sum(cnt.values())))
        f.write("command: ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 "
                "python bench.py --steps 1 --warmup 1 --no-cpu --no-sweep --ncu\n")
        f.write("per-launch times are cold-cache and serialised (the two lanes and the FP64/integer inner-product kernels "
                "overlap in a real run): compare SHARES.  Full list: %s_bench_launches.csv.gz\n\n" % tag)
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, v in tot.most_common():
            f.write("| %s | %d | %.1f | %.1f%% |\n" % (k, cnt[k], v, 100 * v / total))
        f.write("| **total** | %d | %.1f | 100%% |\n" % (sum(cnt.values()), total))


if __name__ == "__main__":
    main(*sys.argv[1:4])
