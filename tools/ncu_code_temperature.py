#!/usr/bin/env python
"""Warp-stall samples of a kernel grouped by how often each SASS instruction runs ("code temperature"): the view that
showed k_bake_stream's cold code (hit shading, sky lookup, projection) stalling on instruction fetch (DESIGN 4.2,
profiles/r02_bake_icache_ab.log). Usage: ncu -i X.ncu-rep --page source --csv > s.csv; tools/ncu_code_temperature.py s.csv"""
import csv
import sys

EDGES = [(0, 0, "never executed"), (1e-9, 1, "< 1 M"), (1, 5, "1-5 M"), (5, 20, "5-20 M"), (20, 60, "20-60 M"), (60, 400, "60-400 M"),
         (400, 1e18, ">= 400 M")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, IndexError, KeyError):
            return 0.0

    total = sum(f(r, "# Samples") for r in data) or 1.0
    acc = {name: [0, 0.0, 0.0, 0.0, 0.0] for _, _, name in EDGES}
    for r in data:
        e = f(r, "Instructions Executed") / 1e6
        for lo, hi, name in EDGES:
            if (e == 0 and hi == 0) or (e > 0 and lo <= e < hi):
                a = acc[name]
                a[0] += 1; a[1] += f(r, "# Samples"); a[2] += f(r, "stall_no_inst"); a[3] += f(r, "stall_long_sb"); a[4] += e
                break
    print("%d instructions (%.1f KB), %d stall samples" % (len(data), len(data) * 16 / 1024.0, total))
    print("%-16s %6s %8s %9s %10s %10s %12s" % ("executions", "instr", "KB", "samples", "no_inst", "long_sb", "warp-instr M"))
    for _, _, name in EDGES:
        a = acc[name]
        print("%-16s %6d %8.1f %8.1f%% %9.1f%% %9.1f%% %12.0f" % (name, a[0], a[0] * 16 / 1024.0, 100 * a[1] / total,
                                                               100 * a[2] / total, 100 * a[3] / total, a[4]))


if __name__ == "__main__":
    main()
