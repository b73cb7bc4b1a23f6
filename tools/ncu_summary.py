#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu -i ... --page raw --csv) or a launch-list csv into the few numbers DESIGN.md / bench.py
quote: per-launch duration, DRAM bytes, DRAM and tensor-pipe utilisation, registers, occupancy limits.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep   |   python tools/ncu_summary.py --launches gpurun_out/x.csv
"""
import collections
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def unit_scale(u):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
            "msecond": 1e3, "nsecond": 1e-3}.get(u, 1.0)


def from_rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:70], "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                try:
                    d[w] = float(r[i].replace(",", "")) * unit_scale(units[i])
                except ValueError:
                    d[w] = r[i]
        yield d


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    cur = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = cur.setdefault(row["ID"], {"kernel": row["Kernel Name"][:60], "grid": row["Grid Size"]})
        v = float(row["Metric Value"].replace(",", "")) * unit_scale(row["Metric Unit"])
        d[row["Metric Name"]] = v
    agg = collections.OrderedDict()
    for d in cur.values():
        a = agg.setdefault((d["kernel"], d["grid"]), [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    print("launches %d, total %.1f us (cold-cache, serialised: compare shares)" % (len(cur), tot))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s %-14s n=%4d avg=%8.1f us share=%5.1f%% dram/launch=%8.2f MB" % (k[0], k[1], a[0], a[1] / a[0], 100 * a[1] / tot, a[2] / a[0] / 1e6))


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        for d in from_rep(sys.argv[1]):
            t = d.get("gpu__time_duration.sum", 0.0)
            by = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
            print("%s grid=%s block=%s" % (d["kernel"], d["grid"], d["block"]))
            print("   %.2f us  dram %.2f MB (%.0f GB/s, %.1f%% of peak)  tensor-pipe active %.1f%%  warps active %.1f%%  regs %s  "
                  "dyn smem %.1f KB  occ limit smem/regs %s/%s  waves/SM %s" % (
                      t, by / 1e6, by / 1e3 / t if t else 0, d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0),
                      d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0),
                      d.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0), d.get("launch__registers_per_thread"),
                      d.get("launch__shared_mem_per_block_dynamic", 0) / 1e3, d.get("launch__occupancy_limit_shared_mem"),
                      d.get("launch__occupancy_limit_registers"), d.get("launch__waves_per_multiprocessor")))
