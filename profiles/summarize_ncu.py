"""Turn an .ncu-rep (ncu --set full) into a compact per-launch CSV for profiles/.
usage: python profiles/summarize_ncu.py gpurun_out/x.ncu-rep profiles/r01_x.csv"""
import csv
import subprocess
import sys

COLS = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
    ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_registers", "ctas_per_sm_by_regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    stall = [i for i, h in enumerate(hdr) if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([n for _, n in COLS] + ["top_stalls"])
        for r in rows[2:]:
            vals = []
            for h, n in COLS:
                v = r[hdr.index(h)] if h in hdr else ""
                if n.endswith("_MB") and v:
                    u = units[hdr.index(h)]
                    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                    v = "%.3f" % (float(v) * scale)
                if n == "time_us" and v:
                    u = units[hdr.index(h)]
                    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                    v = "%.2f" % (float(v) * scale)
                vals.append(v[:70])
            st = sorted(((float(r[i]) if r[i] else 0.0, hdr[i].replace("smsp__average_warps_issue_stalled_", "")
                          .replace("_per_issue_active.ratio", "")) for i in stall), reverse=True)[:4]
            vals.append(" ".join("%s=%.2f" % (n, v) for v, n in st))
            w.writerow(vals)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
