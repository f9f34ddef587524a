#!/bin/bash
# round 2, GPU call B: A/B of force-pass pipelining / register budgets and the shared-memory table in the density pass
mkdir -p gpurun_out
B="--steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-fdm-bench"
for v in f0 f1 f2 f3 d1; do
  for L in 2 4; do
    EPH_B200_LANES=$L EPH_B200_ENGINE_LIB=$PWD/tools/gpu/variants/libeph_b200_$v.so python bench.py $B > gpurun_out/b_${v}_l$L.json 2> gpurun_out/b_${v}_l$L.err
  done
done
for f in gpurun_out/b_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["roofline"]["kernels_ms"])
except Exception as e:
    print("unreadable", e)
PY
done
