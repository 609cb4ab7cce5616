"""Turn `ncu -i x.ncu-rep --page raw --csv` output (or the .ncu-rep itself) into a compact per-launch CSV for profiles/,
plus (optionally) the launch-group totals bench.py reports (DRAM traffic, time-weighted pipe utilisation).
usage: python profiles/summarize_ncu.py RAW.csv|X.ncu-rep OUT.csv [--group N_KERNELS --json OUT.json --note TEXT]"""
import csv
import json
import subprocess
import sys

COLS = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
    ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_registers", "ctas_per_sm_by_regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "int_multiplier_pipe_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_pipe_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible_warps_per_cycle"),
]
SCALE_B = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
SCALE_T = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def load(path):
    if path.endswith(".ncu-rep"):
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
        return list(csv.reader(raw.splitlines()))
    return list(csv.reader(open(path)))


def main(argv):
    rows = load(argv[0])
    hdr, units = rows[0], rows[1]
    stall = [i for i, h in enumerate(hdr) if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    out_rows = []
    for r in rows[2:]:
        rec = {}
        for h, n in COLS:
            v = r[hdr.index(h)] if h in hdr else ""
            u = units[hdr.index(h)] if h in hdr else ""
            if n.endswith("_MB") and v:
                v = "%.3f" % (float(v) * SCALE_B.get(u, 1.0))
            elif n == "time_us" and v:
                v = "%.2f" % (float(v) * SCALE_T.get(u, 1.0))
            elif n not in ("kernel", "grid", "regs", "warp_insts") and v:
                v = "%.1f" % float(v)
            rec[n] = v[:70]
        st = sorted(((float(r[i]) if r[i] else 0.0, hdr[i].replace("smsp__average_warps_issue_stalled_", "")
                      .replace("_per_issue_active.ratio", "")) for i in stall), reverse=True)[:4]
        rec["top_stalls"] = " ".join("%s=%.2f" % (n, v) for v, n in st)
        out_rows.append(rec)
    with open(argv[1], "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=[n for _, n in COLS] + ["top_stalls"])
        w.writeheader()
        w.writerows(out_rows)
    if "--group" in argv:
        g = int(argv[argv.index("--group") + 1])
        grp = out_rows[:g]
        t = sum(float(r["time_us"]) for r in grp)
        wavg = lambda k: sum(float(r[k] or 0) * float(r["time_us"]) for r in grp) / t
        doc = {
            "bytes_per_launch_group": int(1e6 * sum(float(r["dram_read_MB"]) + float(r["dram_write_MB"]) for r in grp)),
            "kernels": g, "sum_time_us_under_ncu": round(t, 2),
            "pipe_utilisation": {"issue_slots_pct": round(wavg("issue_active_pct"), 1),
                                 "integer_multiplier_fmaheavy_pct": round(wavg("int_multiplier_pipe_pct"), 1),
                                 "fp64_pct": round(wavg("fp64_pipe_pct"), 1), "alu_pct": round(wavg("alu_pipe_pct"), 1),
                                 "lsu_pct": round(wavg("lsu_pipe_pct"), 1),
                                 "how": "time-weighted over the launch group's kernels, each measured alone by ncu"},
            "source": argv[argv.index("--note") + 1] if "--note" in argv else argv[0],
        }
        with open(argv[argv.index("--json") + 1], "w") as f:
            json.dump(doc, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1:])
