"""bench.py --impl reference (the CPU arm the driver runs next to ours): one JSON line with the contract's keys,
ranks other than 0 exit silently.  CPU only; here on a reduced volume (the default is the 2048 x 2048 x 256 job)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_reference_arm_contract():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                        "--steps", "1", "--warmup", "1", "--volume", "128", "128", "64", "--chunk", "64", "64", "32"],
                       capture_output=True, text=True, env=env, cwd=ROOT, timeout=580)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gvoxel/s" and d["higher_is_better"] is True
    assert d["metric"] == "Gvoxels/s chunk_ccs+overlaps" and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("configs[3]: 2048x2048x256 volume") and d["config"]["merge_ccs_in_step"]
    assert d["scaling"] == "strong" and d["config"]["grid"] == [2, 2, 2]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # under torchrun only rank 0 works and prints
    env1 = dict(os.environ, RANK="1", WORLD_SIZE="2")
    q = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env1, cwd=ROOT, timeout=120)
    assert q.returncode == 0 and q.stdout.strip() == ""
