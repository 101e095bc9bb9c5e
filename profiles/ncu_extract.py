#!/usr/bin/env python
"""Compact text summary of an .ncu-rep (run where the report lies — the GPU box — so that only the summary
travels): selected raw metrics per captured kernel + the most-sampled SASS lines of each.
usage: ncu_extract.py report.ncu-rep > summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp", "smsp__warp_issue_stalled",
        "launch__shared_mem_per_block", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct"]


def run(args):
    return subprocess.run(["ncu", "-i", sys.argv[1]] + args, capture_output=True, text=True).stdout


def main():
    raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
    hdr = raw[0]
    name_i = hdr.index("Kernel Name")
    cols = [i for i, h in enumerate(hdr) if any(k in h for k in KEYS)]
    for row in raw[2:]:
        print("=== %s" % row[name_i][:110])
        for i in cols:
            if row[i] not in ("", "n/a"):
                print("  %-78s %s %s" % (hdr[i][:78], row[i], raw[1][i]))
    src = run(["--page", "source", "--csv"])
    kernel, rows, header = None, [], None
    def flush():
        if kernel and rows:
            si, ss = header.index("# Samples"), header.index("Source")
            tot = sum(int(r[si]) for r in rows if r[si].isdigit())
            print("--- top sampled SASS of %s (total samples %d)" % (kernel[:90], tot))
            for r in sorted((r for r in rows if r[si].isdigit()), key=lambda r: -int(r[si]))[:22]:
                print("  %6s  %s" % (r[si], r[ss].strip()[:120]))
    for r in csv.reader(io.StringIO(src)):
        if len(r) >= 2 and r[0] == "Kernel Name":
            flush()
            kernel, rows, header = r[1], [], None
        elif r and r[0] == "Address":
            header = r
        elif header and len(r) == len(header):
            rows.append(r)
    flush()


if __name__ == "__main__":
    main()
