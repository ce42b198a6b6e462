#!/usr/bin/env python
"""
Per-source-line stall samples of one kernel from an ncu report (needs -lineinfo, --import-source on).

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_tile_ccl [top=25]

Runs `ncu -i REP --page source --csv --print-source cuda,sass --kernel-name regex:NAME`, keeps the
per-line aggregate rows and prints the lines with the most warp-stall samples and their main stall reasons.
"""
import csv
import io
import subprocess
import sys


def main():
    rep, name = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", f"regex:{name}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, fname, lines = None, "", []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Name":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == "-" and r[0].isdigit():
            d = dict(zip(hdr, r))
            d["_file"], d["_src"] = fname, r[1]
            lines.append(d)
    if not lines:
        print("no per-line rows found")
        return
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(d["# Samples"] or 0) for d in lines)
    print(f"{name}: {tot} samples over {len(lines)} source lines")
    lines.sort(key=lambda d: -int(d["# Samples"] or 0))
    for d in lines[:top]:
        n = int(d["# Samples"] or 0)
        st = sorted(((int(d[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
        sts = " ".join(f"{c}={v}" for (v, c) in st if v)
        print(f"{100.0 * n / tot:5.1f}%  {d['_file']}:{d['Line No']:>4}  {d['_src'].strip()[:80]:80s} | {sts}")


if __name__ == "__main__":
    main()
