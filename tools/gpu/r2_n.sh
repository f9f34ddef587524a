#!/bin/bash
# round 2, GPU call N: late GPU tests + default bench (e2e legs)
mkdir -p gpurun_out
python -m pytest tests/test_zz_gpu_late_additions.py -m gpu -x -q > gpurun_out/n_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/n_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-fdm-bench --no-extras > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/n_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/n_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"])
e=d["e2e"]; print("e2e", e["value"], e["ms_per_step"], e.get("mode"), "host mode", e.get("host_mode",{}).get("value"))
PY
