#!/bin/bash
# round 2, GPU call T (1 GPU): inner-list build with tiles staged in shared memory
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/t_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/t_pytest.log
python bench.py --no-e2e --no-extras --no-cpu-baseline --no-fdm-bench > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/t_bench.err
python bench.py --no-e2e --no-extras --no-cpu-baseline --no-fdm-bench --cells 50 --steps 40 > gpurun_out/t_bench50.json 2> gpurun_out/t_bench50.err
python - <<'PY'
import json
for f in ("t_bench","t_bench50"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["roofline"]["kernels_ms"])
PY
