#!/usr/bin/env python3
"""Turns raw Nsight Compute output into the summaries kept under profiles/.

  launches  <ncu --csv --log-file launches.csv of `--metrics gpu__time_duration.sum --clock-control none`> <out_prefix> "<note>"
            -> <out_prefix>_launches.csv   (one line per launch: kernel, grid, block, microseconds)
               <out_prefix>_launch_shares.csv (per kernel: launches, total ms, share of all kernel time)
  full      <`ncu -i prof.ncu-rep --page raw --csv` output> <out_txt> <traffic_json> <streams_in_capture> "<note>"
            -> per-kernel table of the counters DESIGN.md argues from, and DRAM bytes per stream for bench.py's
               roofline.traffic (dram__bytes_read.sum + dram__bytes_write.sum of the launch / streams in the capture)
"""
import csv
import json
import re
import sys
from collections import OrderedDict

csv.field_size_limit(10 ** 9)

KEEP = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct",
]
UNIT_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def short(name):
    m = re.match(r"(?:cct::)?(\w+)\(", name)
    return m.group(1) if m else name


def launches(src, prefix, note):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    head, rows = rows[0], rows[1:]
    ik, ig, ib, iv, iu = (head.index(x) for x in ("Kernel Name", "Grid Size", "Block Size", "Metric Value", "Metric Unit"))
    per = OrderedDict()
    with open(prefix + "_launches.csv", "w") as f:
        f.write(f"# {note}\nkernel,grid,block,us\n")
        for r in rows:
            us = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1e-3)
            k = short(r[ik])
            f.write(f'{k},"{r[ig]}","{r[ib]}",{us:.3f}\n')
            a = per.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += us
    total = sum(v[1] for v in per.values())
    with open(prefix + "_launch_shares.csv", "w") as f:
        f.write(f"# {note}\n# (every launch of the process: warm-up + parity check + timed steps; cold-cache, serialised: compare SHARES)\n")
        f.write("kernel,launches,total_ms,share\n")
        for k, v in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{v[0]},{v[1] / 1e3:.3f},{v[1] / total:.3f}\n")
    print(f"{len(rows)} launches, {total / 1e3:.1f} ms of kernels -> {prefix}_launches.csv, {prefix}_launch_shares.csv")


def full(src, out_txt, traffic_json, streams, note):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    head, units, rows = rows[0], rows[1], rows[2:]
    ik = head.index("Kernel Name")
    traffic = OrderedDict()
    with open(out_txt, "w") as f:
        f.write(f"# {note}\n# read with: ncu -i prof.ncu-rep --page raw --csv | profiles/summarize_ncu.py full ...\n")
        for r in rows:
            f.write(f"\n## {r[ik]}\n")
            for m in KEEP:
                if m in head:
                    i = head.index(m)
                    f.write(f"{m:<86}{r[i]} {units[i]}\n")
            by = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = head.index(m)
                by += float(r[i].replace(",", "")) * UNIT_BYTES.get(units[i], 1.0)
            t = traffic.setdefault(short(r[ik]), {"dram_bytes_per_stream": 0.0, "streams_in_capture": streams, "launches": 0})
            t["dram_bytes_per_stream"] += by / streams  # a kernel launched once per half batch: the halves add up
            t["launches"] += 1
    json.dump({"source": note, "kernels": traffic}, open(traffic_json, "w"), indent=1)
    print(f"{len(rows)} kernels -> {out_txt}, {traffic_json}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    elif sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), sys.argv[6])
    else:
        sys.exit(__doc__)
