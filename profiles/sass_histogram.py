"""SASS opcode histogram per kernel of libckks_b200.so (cuobjdump -sass; static instruction counts).
usage: python profiles/sass_histogram.py [OUT.md]   -- proves what the hot path is made of: integer multiply-adds
(IMAD / IMAD.WIDE / IMAD.HI), FP64 (DFMA / DADD / DMUL), TMA bulk copies (UBLKCP) with mbarriers (SYNCS), no tensor-core
instruction (HMMA / IMMA / UTCxMMA) anywhere."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "seal-fyp-logistic-regression_b200", "libckks_b200.so")
GROUPS = [("IMAD.WIDE", r"^U?IMAD\.WIDE"), ("IMAD.HI", r"^IMAD\.HI"), ("IMAD (lo/other)", r"^U?IMAD"), ("IADD3/LOP3/SHF/SEL/ISETP", r"^(U?IADD3|LOP3|U?SHF|SEL|ISETP|VIADD|LEA|U?MOV|PRMT|PLOP3)"),
          ("DFMA", r"^DFMA"), ("DADD", r"^DADD"), ("DMUL", r"^DMUL"), ("F2F/I2F/conv", r"^(F2F|I2F|F2I|DSETP|FSEL)"),
          ("LDG/STG", r"^(LDG|STG)"), ("LDS/STS", r"^(LDS|STS)"), ("UBLKCP (TMA bulk)", r"^UBLKCP"), ("SYNCS (mbarrier)", r"^SYNCS"),
          ("BAR", r"^BAR"), ("tensor core (HMMA/IMMA/UTC*MMA)", r"^(HMMA|IMMA|DMMA|UTC.*MMA|QMMA)")]


def main(out):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", name).replace("void ", "")
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    arch = set(re.findall(r"arch = (sm_\w+)", txt))
    with open(out, "w") as f:
        f.write("# SASS opcode histogram of libckks_b200.so (static counts per kernel; `cuobjdump -sass`, arch %s, %d kernels)\n\n" % (
            ", ".join(sorted(arch)), len(kernels)))
        f.write("| kernel | total | " + " | ".join(g for g, _ in GROUPS) + " |\n|---|---:|" + "---:|" * len(GROUPS) + "\n")
        tot = collections.Counter()
        for k, c in kernels.items():
            row, seen = [], set()
            for g, pat in GROUPS:
                n = sum(v for op, v in c.items() if re.match(pat, op) and op not in seen)
                seen |= {op for op in c if re.match(pat, op)}
                row.append(n)
                tot[g] += n
            tot["total"] += sum(c.values())
            if sum(c.values()) >= 150:
                f.write("| %s | %d | %s |\n" % (k[:60], sum(c.values()), " | ".join(str(n) for n in row)))
        f.write("| **all %d kernels** | %d | %s |\n" % (len(kernels), tot["total"], " | ".join(str(tot[g]) for g, _ in GROUPS)))
        f.write("\n(kernels under 150 instructions omitted from the rows, included in the total)\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_histogram.md"))
