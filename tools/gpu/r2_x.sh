#!/bin/bash
# round 2, GPU call X: A/B of the device-side list construction (bench --neigh device, build_neighbors timer)
mkdir -p gpurun_out
for v in neigh_old neigh_unroll4 neigh_old neigh_unroll4; do
  EPH_B200_ENGINE_LIB=$PWD/tools/gpu/variants/libeph_b200_$v.so python bench.py --neigh device --no-extras --no-cpu-baseline --no-fdm-bench --no-e2e > gpurun_out/x_$v.json 2> gpurun_out/x_$v.err
  python - $v <<'PY'
import json,sys
d=json.loads(open("gpurun_out/x_%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["ms_per_step"], d["roofline"]["kernels_ms"].get("build_neighbors"))
PY
done
