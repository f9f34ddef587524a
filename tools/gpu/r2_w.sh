#!/bin/bash
# round 2, GPU call W (1 GPU): device-side list construction with four candidates per trip
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "neigh or list or 131k or build" > gpurun_out/w_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/w_pytest.log
python bench.py --neigh device --no-extras --no-cpu-baseline --no-fdm-bench > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/w_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/w_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["kernels_ms"])
e=d["e2e"]; print("e2e", e["value"], e["ms_per_step"], "host mode", e.get("host_mode",{}).get("value"))
PY
