#!/bin/bash
# round 2, final multi-GPU call: bash tools/gpu/r2_finalN.sh N [static]
N=$1
mkdir -p gpurun_out
run() { # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N $2 > gpurun_out/fN_$1.json 2> gpurun_out/fN_$1.err
  echo "$1 rc=$?"; grep -v OMP_NUM gpurun_out/fN_$1.err | grep -v '^\*\*\*' | tail -2
}
run ${N}gpu ""
if [ "$2" = "static" ]; then run ${N}gpu_static "--mode static --no-e2e --no-check"; fi
python - $N <<'PY'
import json,sys
for f in ("%sgpu"%sys.argv[1], "%sgpu_static"%sys.argv[1]):
    try:
        d=json.loads(open("gpurun_out/fN_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["config"].get("ghost_exchange_transport"), (d.get("parity_vs_n1") or {}).get("max_rel_dev"), (d.get("e2e") or {}).get("value"))
        print("    ", d["roofline"]["kernels_ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
