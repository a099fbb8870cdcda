#!/usr/bin/env python
"""Summarise a .ncu-rep (ncu --set full capture) into a small text file for profiles/ and print the
per-launch DRAM traffic. Usage: tools/ncu_summary.py X.ncu-rep [out.txt]"""
import csv
import io
import subprocess
import sys

KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        out.append("== launch ==")
        vals = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.append("%-70s %s %s" % (k, r[i], units[i]))
                vals[k] = (r[i], units[i])
        # stall reasons (warp-sampling totals)
        st = [(hdr[i], r[i]) for i in range(len(hdr)) if hdr[i].startswith("smsp__pcsamp_warps_issue_stalled_") and not hdr[i].endswith("_not_issued")]
        tot = sum(float(v) for _, v in st if v) or 1.0
        st = sorted(((float(v or 0) / tot, k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k, v in st), reverse=True)[:8]
        out.append("stall samples: " + ", ".join("%s %.1f%%" % (k, 100 * f) for f, k in st))

        def tobytes(k):
            v, u = vals.get(k, ("0", "byte"))
            m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            return float(v) * m
        out.append("dram traffic per launch: %.0f bytes" % (tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")))
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
