#!/usr/bin/env python3
"""Per-CUDA-line shares of a kernel's executed warp instructions and warp-stall samples from an ncu report's source page.

  python tools/ncu_source_lines.py gpurun_out/r02_fq.ncu-rep k_fast_cells_v2 [min_pct [nth-matching-launch]] >> profiles/r02_source_lines.md
"""
import csv
import io
import os
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"] + (["--launch-skip", sys.argv[4], "--launch-count", "1"] if len(sys.argv) > 4 else []),
                                  text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(out)))
    agg, cur_file, hdr, func = {}, None, None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = os.path.basename(r[1]); hdr = None
        elif r[0] == "Function Name":
            func = r[1]
        elif r[0] == "Line No":
            hdr = r
            ii, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), 1
            stall_cols = {n: i for i, n in enumerate(hdr) if n.startswith("stall_") and "(Not Issued)" not in n}
        elif hdr is not None and cur_file is not None and len(r) > ii:
            try:
                ins, smp = float(r[ii] or 0), float(r[isamp] or 0)
            except ValueError:
                continue
            if not r[0].strip().isdigit():
                continue                      # SASS rows under a CUDA line: already counted in the line's own row
            a = agg.setdefault((cur_file, r[0]), [0.0, 0.0, r[isrc], {}])
            a[0] += ins; a[1] += smp
            for n, i in stall_cols.items():
                try:
                    a[3][n] = a[3].get(n, 0.0) + float(r[i] or 0)
                except (ValueError, IndexError):
                    pass
    ti, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    print(f"## `{kern}`  ({ti / 1e6:.1f} M warp instructions, {ts:.0f} stall samples; {func})\n")
    tot_stall = {}
    for a in agg.values():
        for n, v in a[3].items():
            tot_stall[n] = tot_stall.get(n, 0.0) + v
    top = sorted(tot_stall.items(), key=lambda x: -x[1])[:6]
    print("stall reasons (share of samples): " + ", ".join(f"{n[6:]} {100 * v / max(ts, 1):.0f} %" for n, v in top) + "\n")
    print("| file:line | instructions | stall samples | top stall | source |\n|---|---:|---:|---|---|")
    def key(k):
        try:
            return (k[0], int(k[1]))
        except ValueError:
            return (k[0], 0)
    for k in sorted(agg, key=key):
        a = agg[k]
        pi, ps = 100 * a[0] / max(ti, 1), 100 * a[1] / max(ts, 1)
        if pi >= min_pct or ps >= min_pct:
            st = max(a[3].items(), key=lambda x: x[1])[0][6:] if a[3] and max(a[3].values()) > 0 else ""
            print(f"| {k[0]}:{k[1]} | {pi:.1f} % | {ps:.1f} % | {st} | `{a[2].strip()[:110].replace('|', '¦')}` |")
    print()


if __name__ == "__main__":
    main()
