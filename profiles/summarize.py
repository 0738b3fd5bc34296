#!/usr/bin/env python
"""Turn an .ncu-rep (brought back from the GPU box in gpurun_out/) into the small text summary kept under profiles/.
usage: python profiles/summarize.py gpurun_out/apply_r1b.ncu-rep > profiles/apply_r1b.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__waves_per_multiprocessor", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("kernel:", name)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-70s %s %s" % (k, r[i], units[i]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    h = rows[1]
    ix = {x: i for i, x in enumerate(h)}
    data = rows[2:]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    print("warp-stall samples: %d over %d SASS instructions" % (tot, len(data)))
    agg = {x: sum(int(r[ix[x]]) for r in data) for x in h if x.startswith("stall_") and "Not Issued" not in x}
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
        print("  %-28s %7d  %5.1f%%" % (k, v, 100.0 * v / max(tot, 1)))
    print("top instructions by samples:")
    for i in sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:10]:
        print("  %6s samples  %10s exec  %s" % (data[i][ix["# Samples"]], data[i][ix["Instructions Executed"]], data[i][1].strip()))


if __name__ == "__main__":
    main(sys.argv[1])
