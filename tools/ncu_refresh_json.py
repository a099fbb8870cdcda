#!/usr/bin/env python
"""Refresh profiles/ncu_pipes.json and profiles/ncu_traffic.json (the static ncu figures bench.py quotes in its `roofline`
block) from an `ncu --set full` capture. Usage: tools/ncu_refresh_json.py X.ncu-rep <kernel key> <summary path> [note]
e.g. tools/ncu_refresh_json.py gpurun_out/prof_bake_c3_r2.ncu-rep 'vlb::k_bake_stream<9,false,false,false>:c3' profiles/r02_ncu_full_k_bake_stream_c3.txt"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PIPES = {"l1_data_pipe_lsu_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
         "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
         "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
         "lsu_inst_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
         "issue_slots_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
         "l2_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
         "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
         "l1_hit_pct": "l1tex__t_sector_hit_rate.pct", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
         "active_threads_per_warp_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
         "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
         "warp_instructions": "smsp__inst_executed.sum", "kernel_ms_under_ncu": "gpu__time_duration.sum"}


def main():
    rep, key, summary = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]

    def val(name):
        for cand in (name, name.replace("_active", "_elapsed"), name.replace("_elapsed", "_active")):
            if cand in hdr:
                i = hdr.index(cand)
                v = float(r[i].replace(",", ""))
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[i], 1.0)
                return v * scale
        return None

    pipes_path = os.path.join(ROOT, "profiles", "ncu_pipes.json")
    traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    pipes = json.load(open(pipes_path))
    traffic = json.load(open(traffic_path))
    entry = {"capture": summary}
    for k, m in PIPES.items():
        v = val(m)
        if v is not None:
            entry[k] = round(v, 2) if k != "warp_instructions" else int(v)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    entry["dram_read_bytes"], entry["dram_write_bytes"] = int(rd), int(wr)
    if len(sys.argv) > 4:
        entry["note"] = sys.argv[4]
    pipes[key] = entry
    traffic[key] = int(rd + wr)
    json.dump(pipes, open(pipes_path, "w"), indent=1)
    json.dump(traffic, open(traffic_path, "w"), indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
