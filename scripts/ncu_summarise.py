#!/usr/bin/env python
"""Summarises ncu output into the short text files kept under profiles/.

  python scripts/ncu_summarise.py launches gpurun_out/launches_v3.csv      # per-kernel time shares of a launch list
  python scripts/ncu_summarise.py full gpurun_out/prof_trunk_v3.ncu-rep    # selected --set full metrics per launch
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
]


def short(name):
    return name.split("(")[0].split("<")[0].replace("crl::", "").strip()


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    kn, mn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        v = float(r[mv].replace(",", ""))
        us = v / 1e3 if r[mu] in ("ns", "nsecond") else (v if r[mu] in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(short(r[kn]), [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s n=%5d total=%10.3f ms  avg=%9.1f us  share=%5.1f%%" % (k, n, us / 1e3, us / n, 100 * us / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("%s  grid=%s" % (short(r[kn]), r[hdr.index("launch__grid_size")] if "launch__grid_size" in hdr else "?"))
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print("  %s [%s] = %s" % (m, units[i], r[i]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
