#!/usr/bin/env python
"""Markdown table of every kernel in an ncu report: time, DRAM bytes, achieved DRAM GB/s and its fraction of
the measured peak, SM / issue / L1-data-pipe utilisation.  usage: ncu_table.py report.ncu-rep [peak_GBs]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6543.4
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]


def col(r, k, default=""):
    return r[h.index(k)] if k in h else default


def num(r, k):
    try:
        return float(col(r, k, "nan").replace(",", ""))
    except ValueError:
        return float("nan")


def to_bytes(r, k):
    v, unit = num(r, k), u[h.index(k)] if k in h else ""
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def to_ms(r, k):
    v, unit = num(r, k), u[h.index(k)] if k in h else ""
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1)


print("| kernel | grid x block | regs | time ms | DRAM read + write MB | DRAM GB/s | % of measured peak | SM % | issue % | L1 data pipe % |")
print("|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    name = col(r, "Kernel Name")
    short = name.split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    ms = to_ms(r, "gpu__time_duration.sum")
    rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
    gbs = (rd + wr) / (ms * 1e-3) / 1e9 if ms > 0 else float("nan")
    print(f"| `{short}` | {col(r, 'launch__grid_size')} x {col(r, 'launch__block_size')} | {col(r, 'launch__registers_per_thread')} "
          f"| {ms:.3f} | {rd / 1e6:.1f} + {wr / 1e6:.1f} | {gbs:.0f} | {100 * gbs / peak:.1f} "
          f"| {num(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} "
          f"| {num(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} "
          f"| {num(r, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):.0f} |")
