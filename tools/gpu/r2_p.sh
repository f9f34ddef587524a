#!/bin/bash
# round 2, GPU call P: GPU tests after the stand-by / reciprocal / reduce-scatter changes, step breakdown at two brick sizes
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/p_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/p_pytest.log
python tools/gpu/step_parts.py --cells 50 > gpurun_out/p_parts_50.json 2> gpurun_out/p_parts_50.err; echo "parts50 rc=$?"
python tools/gpu/step_parts.py --cells 100 > gpurun_out/p_parts_100.json 2> gpurun_out/p_parts_100.err; echo "parts100 rc=$?"
python bench.py --cells 50 --steps 40 --warmup 5 --no-cpu-baseline --no-fdm-bench --no-extras --no-e2e > gpurun_out/p_bench_50.json 2> gpurun_out/p_bench_50.err
echo "bench50 rc=$?"
python bench.py --cells 50 --mode static --steps 40 --warmup 5 --no-cpu-baseline --no-fdm-bench --no-extras --no-e2e > gpurun_out/p_bench_50s.json 2> gpurun_out/p_bench_50s.err
python - <<'PY'
import json
for f in ("p_parts_50","p_parts_100"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["nlocal"], d["nghost"], "loop", round(d["loop_ms_per_step"],4), "host", round(d["loop_host_ms_per_step"],4))
    for k,v in d["parts_ms"].items(): print("   %-32s %8.4f  host %8.4f"%(k, v["total"], v["host_enqueue"]))
for f in ("p_bench_50","p_bench_50s"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["roofline"]["kernels_ms_per_step"])
PY
