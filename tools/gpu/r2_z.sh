#!/bin/bash
# round 2, GPU call Z (1 GPU): the round-end sequence on the committed tree: GPU tests, smoke, default bench, reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/z_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/z_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/z_smoke.log
python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/z_bench.err
python bench.py --impl reference > gpurun_out/z_bench_reference.json 2> gpurun_out/z_bench_reference.err
echo "reference rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/z_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["step"]["frac"], d["gpu_launches"])
e=d["e2e"]; print("e2e", e["value"], e["ms_per_step"], "host mode", e.get("host_mode",{}).get("value"))
r=json.loads(open("gpurun_out/z_bench_reference.json").read().strip().splitlines()[-1])
print("reference", r["value"], r["ms_per_step"], r["steps"], r["cpu_baseline"]["sample"][:120])
PY
