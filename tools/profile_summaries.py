#!/usr/bin/env python
"""
Turns the two ncu outputs of a round into the text summaries kept under profiles/:

  python tools/profile_summaries.py TAG gpurun_out/TAG_launches.csv gpurun_out/prof_TAG.ncu-rep

  profiles/TAG_launches.csv        copy of the launch list (ncu --metrics gpu__time_duration.sum)
  profiles/TAG_launch_shares.txt   per-kernel mean time and share of one step
  profiles/TAG_top_summary.txt     one line per kernel of the --set full capture (DRAM bytes, issue %, ...)
"""
import collections
import csv
import io
import re
import shutil
import subprocess
import sys


def main():
    tag, launches, rep = sys.argv[1:4]
    shutil.copy(launches, f"profiles/{tag}_launches.csv")
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            agg.setdefault(re.sub(r"\(.*", "", r[ki]), []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    nstep = len(next(v for (k, v) in agg.items() if "k_stream" in k))
    tot = sum(sum(v) for v in agg.values()) / nstep
    out = [f"ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --no-e2e --no-cpu-baseline "
           f"--steps 3 --warmup 3 (512^3, fused chunk_ccs+overlaps), {tag}"]
    for (k, v) in agg.items():
        per = sum(v) / nstep
        out.append(f"{k:48s} n={len(v):3d} mean={sum(v) / len(v) / 1e3:9.1f}us per-step={per / 1e3:8.1f}us "
                   f"share={100 * per / tot:5.1f}%")
    out.append(f"sum of kernel time per step (serialised, cold cache): {tot / 1e3:.1f} us")
    open(f"profiles/{tag}_launch_shares.txt", "w").write("\n".join(out) + "\n")
    print("\n".join(out))

    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]

    def g(r, n):
        return r[hdr.index(n)].replace(",", "")
    ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
    out = [f"ncu --set full --clock-control none --import-source on (python bench.py --no-e2e --no-cpu-baseline "
           f"--steps 3 --warmup 3), one launch per kernel of one 512^3 step; {tag}",
           "dur us; dram read/write as printed by ncu; issue% = smsp__issue_active; inst = warp instructions executed"]
    for r in rows[2:]:
        nm = re.sub(r"\(.*", "", g(r, "Kernel Name"))
        out.append(
            f"{nm:46s} dur={float(g(r, 'gpu__time_duration.sum')):8.1f} dram_rd={g(r, 'dram__bytes_read.sum')}{ur} "
            f"dram_wr={g(r, 'dram__bytes_write.sum')}{uw} "
            f"dram%={float(g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):5.1f} "
            f"issue%={float(g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')):5.1f} "
            f"inst={float(g(r, 'smsp__inst_executed.sum')) / 1e6:7.2f}M "
            f"warps%={float(g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')):5.1f} "
            f"regs={g(r, 'launch__registers_per_thread')} grid={g(r, 'launch__grid_size')} "
            f"L2hit%={float(g(r, 'lts__t_sector_hit_rate.pct')):5.1f}")
    open(f"profiles/{tag}_top_summary.txt", "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
