#!/usr/bin/env python3
"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (tracked).

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python tools/ncu_summary.py report   gpurun_out/prof.ncu-rep  profiles/r01_kernels.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0]
        v = float(r[iv].replace(",", ""))
        v = v / 1000.0 if r[iu] in ("ns", "nsecond") else (v * 1000.0 if r[iu] in ("ms", "msecond") else v)  # -> us
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}): gpu__time_duration.sum, --clock-control none\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f} % |\n")
    print(open(dst).read())


def report(src, dst):
    out = subprocess.check_output(["ncu", "-i", src, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src}), --clock-control none\n\n")
        for r in rows[2:]:
            f.write(f"## `{r[hdr.index('Kernel Name')].split('(')[0]}`  (launch id {r[0]})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    f.write(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")
            f.write("\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
