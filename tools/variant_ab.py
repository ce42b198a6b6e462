#!/usr/bin/env python
"""
A/B of library builds on the single-chunk configs: one subprocess per build (SYNAPTOR_B200_LIB), each running
bench.bench_chunk_config for c2 and c3 with the full-size parity check; synthetic inputs are generated once and
cached under /tmp.

  python tools/variant_ab.py build_variants/libsyn_base.so build_variants/libsyn_v1.so ... > gpurun_out/ab.jsonl
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CACHE = "/tmp/syn_ab_cache"


def cached(fn, name):
    def wrapped(shape, **kw):
        key = os.path.join(CACHE, f"{name}_{'x'.join(map(str, shape))}_{'_'.join(f'{k}{v}' for k, v in sorted(kw.items()))}.npy")
        if os.path.exists(key):
            return np.load(key)
        out = fn(shape, **kw)
        os.makedirs(CACHE, exist_ok=True)
        np.save(key, out)
        return out
    return wrapped


def child(configs, steps):
    import argparse
    import torch
    import bench
    from oracle import oracle as orc
    orc.synth_pred = cached(orc.synth_pred, "pred")
    orc.synth_base = cached(orc.synth_base, "base")
    dev = torch.device("cuda", 0)
    args = argparse.Namespace(steps=steps, warmup=5)
    for name in configs:
        r = bench.bench_chunk_config(torch, dev, name, args, True)
        keep = {k: r[k] for k in ("ms_per_step", "ms_per_step_two_chunks_in_flight", "stage_ms")}
        keep["parity"] = r.get("parity", {}).get("bit_exact_vs_oracle")
        print(json.dumps({"lib": os.path.basename(os.environ.get("SYNAPTOR_B200_LIB", "default")), "config": name, **keep}),
              flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2].split(","), int(sys.argv[3]))
    else:
        libs = sys.argv[1:]
        for lib in libs:
            env = dict(os.environ)
            if lib != "default":
                env["SYNAPTOR_B200_LIB"] = os.path.abspath(lib)
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "c2,c3", "20"], env=env,
                               capture_output=True, text=True, timeout=300)
            sys.stdout.write(p.stdout)
            if p.returncode != 0:
                sys.stdout.write(json.dumps({"lib": lib, "error": p.stderr[-600:]}) + "\n")
            sys.stdout.flush()
