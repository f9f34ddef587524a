#!/bin/bash
# round 2, final 1-GPU call: GPU tests, smoke, default bench, static mode, reference arm, ncu launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/f1_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/f1_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/f1_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/f1_smoke.log
python bench.py > gpurun_out/f1_bench.json 2> gpurun_out/f1_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/f1_bench.err
python bench.py --mode static --no-e2e --no-extras --no-cpu-baseline --no-fdm-bench > gpurun_out/f1_bench_static.json 2> gpurun_out/f1_bench_static.err
echo "static rc=$?"
python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/f1_bench_reference.json 2> gpurun_out/f1_bench_reference.err
echo "reference rc=$?"; tail -c 600 gpurun_out/f1_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/f1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fdm-bench --no-extras > gpurun_out/f1_ncu_bench.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/f1_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["step"]["frac"], d["gpu_launches"])
print(d["roofline"]["kernels_ms"])
e=d["e2e"]; print("e2e", e["value"], e["ms_per_step"], "host mode", e.get("host_mode",{}).get("value"))
print(d["extras"]["parity_vs_reference"]["max_rel_dev"], d["extras"]["C4_NiCoCrFe"]["value"]); print(d["cpu_baseline"]["value"])
s=json.loads(open("gpurun_out/f1_bench_static.json").read().strip().splitlines()[-1])
print("static", s["value"], s["ms_per_step"], s["roofline"]["step"]["frac"])
PY
