#!/usr/bin/env python
"""
Times the device pipeline (inputs resident, CUDA events) on the BASELINE.json configs other than the
headline one, and prints one JSON line per case.  Parity for these shapes is covered by tests/ at
reduced size; here only counts are sanity-checked.

  python tools/config_sweep.py [--cases c1,c3,c5] [--iters 5]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_case(name, shape, sigma, fg, conn, dil, with_base, iters, torch, D, orc, C):
    dev = torch.device("cuda", torch.cuda.current_device())
    pred = torch.from_numpy(orc.synth_pred(shape, seed=0, sigma=sigma, fg=fg)).to(dev)
    base = torch.from_numpy(orc.synth_base(shape, seed=1).view(np.int64)).to(dev) if with_base else None
    res = D.chunk_ccs_device(pred, np.float32(0.5), 100, connectivity=conn, dil_k=dil, base=base, sync=True)
    counts = dict(res.counts)
    for _ in range(2):
        D.chunk_ccs_device(pred, np.float32(0.5), 100, connectivity=conn, dil_k=dil, base=base, sync=False, out=res)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        D.chunk_ccs_device(pred, np.float32(0.5), 100, connectivity=conn, dil_k=dil, base=base, sync=False, out=res)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    L = C.lib()
    C.check(L.syn_profile_enable(1))
    D.chunk_ccs_device(pred, np.float32(0.5), 100, connectivity=conn, dil_k=dil, base=base, sync=False, out=res)
    stages = C.profile_read()
    C.check(L.syn_profile_enable(0))
    after = D.read_counts(res)
    nvox = int(np.prod(shape))
    bpv = 16 if with_base else 8
    line = {"case": name, "shape": list(shape), "sigma": sigma, "fg": fg, "connectivity": conn, "dil_k": dil,
            "overlaps": with_base, "ms": ms, "gvox_s": nvox / ms / 1e6, "gb_s_algorithmic": bpv * nvox / ms / 1e6,
            "runs": counts["runs"], "components": counts["components"], "kept": counts["kept"],
            "pairs": counts["pairs"], "errflags": int(after.get("errflags", 0)), "flagged_tiles": after["flagged_tiles"], "tile_records": after["tile_records"],
            "stage_us": {n: round(1e3 * float(v), 1) for (n, v) in stages.items()}}
    print(json.dumps(line), flush=True)
    del pred, base, res
    torch.cuda.empty_cache()
    D.release_workspaces()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="c1,c1_26,c2,c2_ccs_only,c3,c5_128,c5_256,c5_256_k2,c5_256_k5,c5_512_k5,c5_1024")
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    import torch
    from oracle import oracle as orc            # synthetic input generator only
    from synaptor_b200 import _cabi as C
    from synaptor_b200 import _device as D
    cases = {
        "c1": ((256, 256, 256), 2.0, 0.03, 6, 0, False),
        "c1_26": ((256, 256, 256), 2.0, 0.03, 26, 0, False),
        "c2": ((512, 512, 512), 2.0, 0.03, 6, 0, True),
        "c2_ccs_only": ((512, 512, 512), 2.0, 0.03, 6, 0, False),
        "c3": ((1024, 1024, 128), 1.0, 0.10, 6, 0, True),
        "c5_128": ((128, 128, 128), 2.0, 0.03, 6, 0, False),
        "c5_256": ((256, 256, 256), 2.0, 0.03, 6, 0, False),
        "c5_256_k2": ((256, 256, 256), 2.0, 0.03, 6, 2, False),
        "c5_256_k5": ((256, 256, 256), 2.0, 0.03, 6, 5, False),
        "c5_512_k5": ((512, 512, 512), 2.0, 0.03, 6, 5, False),
        "c5_1024": ((1024, 1024, 1024), 2.0, 0.03, 6, 0, False),
    }
    for name in a.cases.split(","):
        run_case(name, *cases[name], a.iters, torch, D, orc, C)


if __name__ == "__main__":
    main()
