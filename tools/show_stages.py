#!/usr/bin/env python
"""Prints ms/step and the per-kernel stage times of a bench.py JSON line (file argument)."""
import json
import sys

d = json.load(open(sys.argv[1]))
print(f"ms_per_step={d['ms_per_step']:.4f}  value={d['value']:.1f} {d['unit']}  pipeline_frac={d['roofline']['pipeline']['frac']:.3f}")
a = d["roofline"]["stage_ms"]
for k in a:
    print(f"  {k:18s} {a[k]*1e3:8.1f} us")
print(f"  {'sum (with events)':18s} {sum(a.values())*1e3:8.1f} us;  roofline kernel frac {d['roofline']['frac']:.3f}")
