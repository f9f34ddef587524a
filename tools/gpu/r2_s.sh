#!/bin/bash
# round 2, GPU call S (1 GPU): resident GPU test + default bench after the faster inner-list build and the resident-mode force pass that no longer waits for the pair forces
mkdir -p gpurun_out
python -m pytest tests/test_zz_gpu_late_additions.py -m gpu -x -q > gpurun_out/s_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/s_pytest.log
python bench.py > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/s_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["step"]["frac"])
print(d["roofline"]["kernels_ms"])
e=d["e2e"]; print("e2e", e["value"], e["ms_per_step"], "host mode", e.get("host_mode",{}).get("value"))
print(d["extras"]); print(d["cpu_baseline"]); print(d["fdm"])
PY
