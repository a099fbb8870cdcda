#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: top SASS instructions by warp-stall samples with
the dominant stall reason. Usage: ncu -i X.ncu-rep --page source --csv > s.csv; ncu_top.py s.csv [N]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    kern = 0
    i = 0
    while i < len(rows):
        r = rows[i]
        if 'Source' in r and any('Sampling' in c for c in r):
            hdr = r
            kern += 1
            j = i + 1
            data = []
            while j < len(rows) and not ('Source' in rows[j] and any('Sampling' in c for c in rows[j])):
                data.append(rows[j])
                j += 1
            report(kern, hdr, data, n)
            i = j
        else:
            i += 1


def report(kern, hdr, data, n):
    si = hdr.index('Source')
    samp = hdr.index('Warp Stall Sampling (All Samples)')
    stall_cols = [(k, c) for k, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
    out = []
    tot = 0
    reason_tot = {}
    for idx, r in enumerate(data):
        if len(r) <= samp:
            continue
        try:
            s = int(r[samp])
        except ValueError:
            continue
        tot += s
        reasons = []
        for k, c in stall_cols:
            try:
                v = int(r[k])
            except (ValueError, IndexError):
                v = 0
            if v:
                reasons.append((v, c[6:]))
                reason_tot[c[6:]] = reason_tot.get(c[6:], 0) + v
        reasons.sort(reverse=True)
        out.append((s, idx, r[si][:90], " ".join("%s:%d" % (c, v) for v, c in reasons[:3])))
    print("=== kernel #%d: %d samples, %d instructions" % (kern, tot, len(data)))
    print("    reasons:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(tot, 1)) for k, v in sorted(reason_tot.items(), key=lambda x: -x[1])[:8]))
    for s, idx, src, why in sorted(out, reverse=True)[:n]:
        print("%6d %5.1f%% @%-5d %-90s %s" % (s, 100.0 * s / max(tot, 1), idx, src, why))


if __name__ == "__main__":
    main()
