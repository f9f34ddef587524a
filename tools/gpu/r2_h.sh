#!/bin/bash
# round 2, GPU call H: force pass shaped like the density pass (two slots per iteration, loads first)
mkdir -p gpurun_out
B="--steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-fdm-bench"
for v in s256x4 d256x3 d128x6 d128x5 d256x4; do
  for L in 2 4; do
    EPH_B200_LANES=$L EPH_B200_ENGINE_LIB=$PWD/tools/gpu/variants/libeph_b200_$v.so python bench.py $B > gpurun_out/h_${v}_l$L.json 2> gpurun_out/h_${v}_l$L.err
  done
done
for f in gpurun_out/h_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernels_ms"]
    print(d["ms_per_step"], "dens", k["density_sweep"], "force", k["force_sweep"])
except Exception as e:
    print("unreadable", e)
PY
done
