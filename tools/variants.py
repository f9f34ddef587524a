"""Kernel-time comparison of the engine's tuning knobs on one GPU (development aid).
usage: python tools/variants.py [cells]"""
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cells = sys.argv[1] if len(sys.argv) > 1 else "64"
variants = [
    dict(),
]
if len(sys.argv) > 2:
    variants = variants[: int(sys.argv[2])]
for v in variants:
    env = dict(os.environ, **v)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--cells", cells, "--steps", "6", "--warmup", "2",
                          "--no-cpu-baseline", "--no-e2e", "--no-fdm-bench", "--grid", "32"], env=env, capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
        k = j["roofline"]["kernels_ms"]
        print(json.dumps(dict(v=v, ms_per_step=round(j["ms_per_step"], 3), value=round(j["value"] / 1e6, 1), kernels=k)), flush=True)
    except Exception as e:
        print("FAILED", v, out.stderr[-500:], flush=True)
